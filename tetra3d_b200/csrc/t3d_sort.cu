// t3d_sort.cu -- the per-draw 512-bin depth sort for draws that span several tiles (sm_100a).
//
// The reference's sortingTriangleBucket (sortingTriangles.go:38-125) is a counting sort on a 9-bit bin
// with arrival order preserved inside a bin, run once per MeshPart.  Triangles of a frame already sit in
// the arena draw after draw, so the frame-wide order is one *segmented* counting sort pass:
//
//   k_bin_hist   per tile: bin every survivor (the draw's min / max are final after k_cull_large), count
//                the tile's bins; the last tile of a draw to finish turns the draw's [tiles x 512] counts
//                into offsets (bins in visiting order, tiles in arena order inside a bin).
//   k_scatter    per tile: stable in-tile rank of every survivor inside its bin (warp match + per-warp
//                counts), then sorted[tbase + offset + rank] = triangle slot.
//   k_tile_scan  exclusive scan of the per-tile survivor counts: where each tile's triangles start in
//                the output lists, and the frame's total.
//
// Draws that fit inside one tile never come here: k_tile_packed (t3d_kernels.cu) culls and sorts them in
// shared memory.  Integer work only, apart from the bin expression itself.
#include "t3d_internal.h"
#include "t3d_scan.cuh"
#include "../../include/t3dgeo.h"

namespace t3d {

static const int SORT_THREADS = 256;
static const int SORT_ROUNDS = TILE_TRIS / SORT_THREADS;
static const int BINS_PER_THREAD = BIN_COUNT / SORT_THREADS;

// Stand-alone tile scan, one block per chunk (frames whose scan does not ride along in k_bin_hist's extra block).
__global__ void __launch_bounds__(1024)
k_tile_scan(const TriTile* __restrict__ tiles, uint32_t nTiles, const DrawStat* __restrict__ stats,
            uint32_t* __restrict__ tileSurv, uint32_t* __restrict__ tileOut, uint32_t* __restrict__ chunkSum) {
    __shared__ uint32_t sWarp[32];
    pdl_wait();
    pdl_launch();
    tile_scan_chunk<1024>(tiles, nTiles, blockIdx.x, stats, tileSurv, tileOut, chunkSum, sWarp);
}

// Histogram rows: tile j of a large draw owns row T.row0 + j; row T.row0 + nTilesOfDraw receives the bin bases.
__global__ void __launch_bounds__(SORT_THREADS)
k_bin_hist(const TriTile* __restrict__ tiles, const uint32_t* __restrict__ largeList, uint32_t nLarge, uint32_t nTiles,
           DrawStat* __restrict__ stats, uint32_t* __restrict__ tdepth, uint32_t* __restrict__ hist,
           uint32_t* __restrict__ tileSurv, uint32_t* __restrict__ tileOut, uint32_t* __restrict__ chunkSum) {
    __shared__ uint32_t sHist[BIN_COUNT];
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sLast;
    if (blockIdx.x == nLarge) {     // the extra block (launched only when nTiles is valid): where every tile starts in the lists (needs only the survivor counts)
        pdl_wait();
        pdl_launch();
        tile_scan_chunk<SORT_THREADS>(tiles, nTiles, 0, stats, tileSurv, tileOut, chunkSum, sWarp);      // nTiles <= SCAN_CHUNK
        return;
    }
    const TriTile T = tiles[__ldg(&largeList[blockIdx.x])];
    const uint32_t d = T.dFirst;
    const uint32_t j = T.info & ~TILE_LARGE;
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) sHist[threadIdx.x + q * SORT_THREADS] = 0;
    pdl_wait();
    pdl_launch();
    uint32_t bits[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = r * SORT_THREADS + threadIdx.x;
        bits[r] = (i < T.count) ? tdepth[T.start + i] : KEY_CULLED;
    }
    const uint32_t minBits = stats[d].minBits, maxBits = stats[d].maxBits;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        if (bits[r] != KEY_CULLED) {
            const uint32_t key = bin_key(bits[r], minBits, maxBits, T.sortMode);
            atomicAdd(&sHist[key], 1u);
            tdepth[T.start + r * SORT_THREADS + threadIdx.x] = key;   // the depth is not needed again; k_scatter reads the bin rank
        }
    }
    __syncthreads();
    // thread owns bins 2*tid, 2*tid+1 from here on (block scans then run in bin order)
    uint2* myRow = reinterpret_cast<uint2*>(hist + (size_t)(T.row0 + j) * BIN_COUNT);
    myRow[threadIdx.x] = make_uint2(sHist[2 * threadIdx.x], sHist[2 * threadIdx.x + 1]);
    if (T.nTilesOfDraw > HUGE_TILES) return;            // k_hist_scan_huge turns this draw's counts into offsets
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sLast = (atomicAdd(&stats[d].done, 1u) == T.nTilesOfDraw - 1) ? 1u : 0u;
    __syncthreads();
    if (!sLast) return;
    __threadfence();

    // Last tile of the draw: per-tile counts -> exclusive prefix over the draw's tiles (arena order inside a bin),
    // totals -> exclusive prefix over bins in visiting order (the base row).
    uint2* rows = reinterpret_cast<uint2*>(hist + (size_t)T.row0 * BIN_COUNT);
    uint2 run = make_uint2(0u, 0u);
    const uint32_t nT = T.nTilesOfDraw;
    for (uint32_t t0 = 0; t0 < nT; t0 += 8) {
        uint2 c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) c[u] = (t0 + u < nT) ? __ldcg(&rows[(size_t)(t0 + u) * (BIN_COUNT / 2) + threadIdx.x]) : make_uint2(0u, 0u);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (t0 + u < nT) rows[(size_t)(t0 + u) * (BIN_COUNT / 2) + threadIdx.x] = run;
            run.x += c[u].x; run.y += c[u].y;
        }
    }
    const uint32_t base = block_exclusive_scan_256(run.x + run.y, sWarp, nullptr);
    rows[(size_t)nT * (BIN_COUNT / 2) + threadIdx.x] = make_uint2(base, base + run.x);
}

// Draws of more than HUGE_TILES tiles: sixteen blocks per draw, each owning 32 of the 512 bins (a 128-byte column of
// the draw's [tiles x 512] counts), turn the counts into per-tile offsets.  Warp w owns a contiguous run of the
// draw's tiles, lane l one bin: per-run sums, an exclusive scan of the 8 run sums per bin in shared memory, then a
// second sweep writes offset(tile, bin) = earlier runs + earlier tiles of this run.  The base row receives the bin
// TOTALS; k_scatter scans them itself (512 values).  (For a draw of 1 M triangles the last tile's serial scan in
// k_bin_hist would be a 45 us tail.)
static const int HUGE_GROUPS = BIN_COUNT / 32;
__global__ void __launch_bounds__(SORT_THREADS)
k_hist_scan_huge(const TriTile* __restrict__ tiles, const uint32_t* __restrict__ hugeList, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sSeg[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();
    pdl_launch();
    const TriTile T = tiles[__ldg(&hugeList[blockIdx.x / HUGE_GROUPS])];
    const uint32_t nT = T.nTilesOfDraw;
    uint32_t* col = hist + (size_t)T.row0 * BIN_COUNT + (blockIdx.x % HUGE_GROUPS) * 32 + lane;    // this lane's bin in row 0
    const uint32_t per = (nT + 7) / 8;
    const uint32_t t0 = min(warp * per, nT), t1 = min(t0 + per, nT);
    uint32_t sum = 0;
    {
        uint32_t t = t0;
        for (; t + 8 <= t1; t += 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldcg(col + (size_t)(t + u) * BIN_COUNT);
#pragma unroll
            for (int u = 0; u < 8; u++) sum += v[u];
        }
        for (; t < t1; t++) sum += __ldcg(col + (size_t)t * BIN_COUNT);
    }
    sSeg[warp][lane] = sum;
    __syncthreads();
    if (warp == 0) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t c = sSeg[w][lane]; sSeg[w][lane] = run; run += c; }
        col[(size_t)nT * BIN_COUNT] = run;              // base row: the bin total
    }
    __syncthreads();
    uint32_t run = sSeg[warp][lane];
    {
        uint32_t t = t0;
        for (; t + 8 <= t1; t += 8) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldcg(col + (size_t)(t + u) * BIN_COUNT);
#pragma unroll
            for (int u = 0; u < 8; u++) { col[(size_t)(t + u) * BIN_COUNT] = run; run += v[u]; }
        }
        for (; t < t1; t++) { const uint32_t v = __ldcg(col + (size_t)t * BIN_COUNT); col[(size_t)t * BIN_COUNT] = run; run += v; }
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
k_scatter(const TriTile* __restrict__ tiles, const uint32_t* __restrict__ largeList,
          const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ hist, uint32_t* __restrict__ sorted) {
    // Warp w owns the tile's slots [w * 128, w * 128 + 128) and walks them in SORT_ROUNDS rounds of 32, so
    // arrival order is (warp, round, lane).  sCnt[w][bin] accumulates the warp's count of a bin round by round
    // (the value before a round's update is the rank base of that round's group), then becomes the exclusive
    // prefix over warps.
    __shared__ __align__(16) uint16_t sCnt[8][BIN_COUNT];
    __shared__ uint32_t sBase[BIN_COUNT];
    __shared__ uint32_t sWarp[8];
    const unsigned full = 0xFFFFFFFFu;
    const TriTile T = tiles[__ldg(&largeList[blockIdx.x])];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        uint4* z = reinterpret_cast<uint4*>(&sCnt[0][0]);
        static_assert(sizeof(sCnt) == 2 * SORT_THREADS * sizeof(uint4), "two 128-bit stores per thread clear the counts");
        z[threadIdx.x] = make_uint4(0, 0, 0, 0); z[threadIdx.x + SORT_THREADS] = make_uint4(0, 0, 0, 0);
    }
    pdl_wait();             // bin ranks and offsets are k_bin_hist's
    pdl_launch();
    uint32_t key[SORT_ROUNDS], rank[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = warp * (SORT_ROUNDS * 32) + r * 32 + lane;
        key[r] = (i < T.count) ? __ldg(&tkeys[T.start + i]) : KEY_CULLED;
    }
    {
        const uint32_t* offRow = hist + (size_t)(T.row0 + (T.info & ~TILE_LARGE)) * BIN_COUNT;
        const uint32_t* baseRow = hist + (size_t)(T.row0 + T.nTilesOfDraw) * BIN_COUNT;
        if (T.nTilesOfDraw > HUGE_TILES) {              // the base row holds bin totals (k_hist_scan_huge): scan them here
            const uint2 c = __ldg(reinterpret_cast<const uint2*>(baseRow) + threadIdx.x);
            const uint2 o = __ldg(reinterpret_cast<const uint2*>(offRow) + threadIdx.x);
            const uint32_t base = block_exclusive_scan_256(c.x + c.y, sWarp, nullptr);
            sBase[2 * threadIdx.x] = T.tbase + o.x + base;
            sBase[2 * threadIdx.x + 1] = T.tbase + o.y + base + c.x;
        } else {
#pragma unroll
            for (int q = 0; q < BINS_PER_THREAD; q++) {
                const uint32_t b = threadIdx.x + q * SORT_THREADS;
                sBase[b] = T.tbase + __ldg(&offRow[b]) + __ldg(&baseRow[b]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const unsigned peers = __match_any_sync(full, key[r]);
        const uint32_t inGroup = __popc(peers & ((1u << lane) - 1u));
        uint32_t before = 0;
        if (key[r] != KEY_CULLED && inGroup == 0) {
            before = sCnt[warp][key[r]];
            sCnt[warp][key[r]] = (uint16_t)(before + __popc(peers));
        }
        rank[r] = __shfl_sync(full, before, __ffs(peers) - 1) + inGroup;
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over the 8 warps, two adjacent bins per thread packed in one 32-bit word (counts <= 1024)
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint32_t* word = reinterpret_cast<uint32_t*>(&sCnt[w][0]) + threadIdx.x;
            const uint32_t c = *word;
            *word = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        if (key[r] != KEY_CULLED)
            sorted[sBase[key[r]] + sCnt[warp][key[r]] + rank[r]] = T.start + warp * (SORT_ROUNDS * 32) + r * 32 + lane;
    }
}

void launch_bin_hist(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge, uint32_t nTiles,
                     DrawStat* stats, uint32_t* tdepth, uint32_t* hist, uint32_t* tileSurv, uint32_t* tileOut, uint32_t* chunkSum) {
    if (!nLarge) return;
    launch_kernel(k_bin_hist, nLarge + (nTiles != 0xFFFFFFFFu ? 1 : 0), SORT_THREADS, 0, s, pdl, tiles, largeList, nLarge, nTiles, stats, tdepth, hist, tileSurv, tileOut, chunkSum);
}

void launch_hist_scan_huge(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* hugeList, uint32_t nHuge, uint32_t* hist) {
    if (!nHuge) return;
    launch_kernel(k_hist_scan_huge, nHuge * HUGE_GROUPS, SORT_THREADS, 0, s, pdl, tiles, hugeList, hist);
}

void launch_scatter(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                    const uint32_t* tkeys, const uint32_t* hist, uint32_t* sorted) {
    if (!nLarge) return;
    launch_kernel(k_scatter, nLarge, SORT_THREADS, 0, s, pdl, tiles, largeList, tkeys, hist, sorted);
}

void launch_tile_scan(cudaStream_t s, bool pdl, const TriTile* tiles, uint32_t nTiles, const DrawStat* stats, uint32_t* tileSurv,
                      uint32_t* tileOut, uint32_t* chunkSum) {
    launch_kernel(k_tile_scan, (nTiles + SCAN_CHUNK - 1) / SCAN_CHUNK, 1024, 0, s, pdl, tiles, nTiles, stats, tileSurv, tileOut, chunkSum);
}

}  // namespace t3d
