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
#include "../../include/t3dgeo.h"

namespace t3d {

static const int SORT_THREADS = 256;
static const int SORT_ROUNDS = TILE_TRIS / SORT_THREADS;
static const int BINS_PER_THREAD = BIN_COUNT / SORT_THREADS;

// Exclusive scan of one value per thread over a 256-thread block.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* sWarp /*8*/, uint32_t* total) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(full, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t wsum = (lane < 8) ? sWarp[lane] : 0;
    uint32_t winc = wsum;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        uint32_t n = __shfl_up_sync(full, winc, o);
        if (lane >= o) winc += n;
    }
    const uint32_t warpBase = __shfl_sync(full, winc - wsum, warp);
    if (total) *total = __shfl_sync(full, winc, 7);
    __syncthreads();
    return warpBase + inc - v;
}

// hist row r belongs to large tile largeList[r]; the rows of one draw are consecutive.
__global__ void __launch_bounds__(SORT_THREADS)
k_bin_hist(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ largeList,
           DrawStat* __restrict__ stats, uint32_t* __restrict__ tdepth, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sHist[BIN_COUNT];
    __shared__ uint32_t sWarp[8];
    __shared__ uint32_t sLast;
    const uint32_t row = blockIdx.x;
    const TriTile T = tiles[__ldg(&largeList[row])];
    const uint32_t d = T.dFirst;
    const uint32_t sortMode = draws[d].sortMode;
    const uint32_t nt = draws[d].nt;
    const uint32_t minBits = stats[d].minBits, maxBits = stats[d].maxBits;
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) sHist[threadIdx.x + q * SORT_THREADS] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = r * SORT_THREADS + threadIdx.x;
        if (i < T.count) {
            const uint32_t bits = tdepth[T.start + i];
            if (bits != KEY_CULLED) {
                const uint32_t key = bin_key(bits, minBits, maxBits, sortMode);
                atomicAdd(&sHist[key], 1u);
                tdepth[T.start + i] = key;          // the depth is not needed again; k_scatter reads the bin rank
            }
        }
    }
    __syncthreads();
    uint32_t* myRow = hist + (size_t)row * BIN_COUNT;
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) myRow[threadIdx.x + q * SORT_THREADS] = sHist[threadIdx.x + q * SORT_THREADS];
    __threadfence();
    __syncthreads();
    const uint32_t nTilesOfDraw = (nt + TILE_TRIS - 1) / TILE_TRIS;
    if (threadIdx.x == 0) sLast = (atomicAdd(&stats[d].done, 1u) == nTilesOfDraw - 1) ? 1u : 0u;
    __syncthreads();
    if (!sLast) return;
    __threadfence();

    // Last tile of the draw: counts -> offsets.  offset(tile, bin) = sum of the draw's earlier bins over all
    // tiles + the same bin in earlier tiles.
    uint32_t* rows = hist + (size_t)(row - (T.info & ~TILE_LARGE)) * BIN_COUNT;
    // thread owns bins [BINS_PER_THREAD * tid, +BINS_PER_THREAD) so that the block scan runs in bin order
    uint32_t tot[BINS_PER_THREAD];
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) {
        const uint32_t b = threadIdx.x * BINS_PER_THREAD + q;
        uint32_t run = 0;
        for (uint32_t t = 0; t < nTilesOfDraw; t++) {
            uint32_t* p = rows + (size_t)t * BIN_COUNT + b;
            const uint32_t c = __ldcg(p);
            *p = run;
            run += c;
        }
        tot[q] = run;
    }
    uint32_t sum = 0;
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) sum += tot[q];
    uint32_t base = block_exclusive_scan_256(sum, sWarp, nullptr);
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) {
        const uint32_t b = threadIdx.x * BINS_PER_THREAD + q;
        if (base) for (uint32_t t = 0; t < nTilesOfDraw; t++) rows[(size_t)t * BIN_COUNT + b] += base;
        base += tot[q];
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
k_scatter(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ largeList,
          const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ hist, uint32_t* __restrict__ sorted) {
    // sCnt[(round, warp)][bin]: how many of that warp's lanes hold the bin in that round, then (after the
    // scan over (round, warp), which is arrival order) the in-tile rank of the group's first lane.
    __shared__ uint16_t sCnt[SORT_ROUNDS * 8][BIN_COUNT];
    __shared__ uint32_t sBase[BIN_COUNT];
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t row = blockIdx.x;
    const TriTile T = tiles[__ldg(&largeList[row])];
    const uint32_t tb = draws[T.dFirst].tbase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&sCnt[0][0]);
        for (uint32_t i = threadIdx.x; i < SORT_ROUNDS * 8 * BIN_COUNT / 2; i += SORT_THREADS) z[i] = 0;
#pragma unroll
        for (int q = 0; q < BINS_PER_THREAD; q++) sBase[threadIdx.x + q * SORT_THREADS] = hist[(size_t)row * BIN_COUNT + threadIdx.x + q * SORT_THREADS];
    }
    __syncthreads();
    uint32_t key[SORT_ROUNDS], rank[SORT_ROUNDS];
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = r * SORT_THREADS + threadIdx.x;
        key[r] = (i < T.count) ? __ldg(&tkeys[T.start + i]) : KEY_CULLED;
        const unsigned peers = __match_any_sync(full, key[r]);
        rank[r] = __popc(peers & ((1u << lane) - 1u));
        if (key[r] != KEY_CULLED && rank[r] == 0) sCnt[r * 8 + warp][key[r]] = (uint16_t)__popc(peers);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BINS_PER_THREAD; q++) {
        const uint32_t b = threadIdx.x + q * SORT_THREADS;
        uint32_t run = 0;
#pragma unroll 8
        for (int e = 0; e < SORT_ROUNDS * 8; e++) { const uint32_t c = sCnt[e][b]; sCnt[e][b] = (uint16_t)run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        if (key[r] != KEY_CULLED) {
            const uint32_t i = r * SORT_THREADS + threadIdx.x;
            sorted[tb + sBase[key[r]] + sCnt[r * 8 + warp][key[r]] + rank[r]] = T.start + i;
        }
    }
}

// One block.  Large tile j of a draw holds survivors [j * TILE_TRIS, ...) of the draw's sorted run; packed
// tiles wrote their own counts.
__global__ void __launch_bounds__(1024)
k_tile_scan(const TriTile* __restrict__ tiles, uint32_t nTiles, const DrawStat* __restrict__ stats,
            uint32_t* __restrict__ tileSurv, uint32_t* __restrict__ tileOut, uint32_t* __restrict__ dCount) {
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sCarry;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nTiles; base += 1024) {
        const uint32_t t = base + threadIdx.x;
        uint32_t c = 0;
        if (t < nTiles) {
            const TriTile T = tiles[t];
            if (T.info & TILE_LARGE) {
                const uint32_t surv = stats[T.dFirst].survivors, off = (T.info & ~TILE_LARGE) * TILE_TRIS;
                c = surv > off ? min(surv - off, TILE_TRIS) : 0u;
                tileSurv[t] = c;
            } else {
                c = tileSurv[t];
            }
        }
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(full, inc, o); if (lane >= o) inc += n; }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sWarp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(full, wi, o); if (lane >= o) wi += n; }
            sWarp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = sCarry;
        const uint32_t excl = carry + sWarp[warp] + inc - c;
        if (t < nTiles) tileOut[t] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) sCarry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) { tileOut[nTiles] = sCarry; *dCount = sCarry; }
}

void launch_bin_hist(cudaStream_t s, const DrawRec* draws, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                     DrawStat* stats, uint32_t* tdepth, uint32_t* hist) {
    if (!nLarge) return;
    k_bin_hist<<<nLarge, SORT_THREADS, 0, s>>>(draws, tiles, largeList, stats, tdepth, hist);
}

void launch_scatter(cudaStream_t s, const DrawRec* draws, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                    const uint32_t* tkeys, const uint32_t* hist, uint32_t* sorted) {
    if (!nLarge) return;
    k_scatter<<<nLarge, SORT_THREADS, 0, s>>>(draws, tiles, largeList, tkeys, hist, sorted);
}

void launch_tile_scan(cudaStream_t s, const TriTile* tiles, uint32_t nTiles, const DrawStat* stats, uint32_t* tileSurv,
                      uint32_t* tileOut, uint32_t* dCount) {
    k_tile_scan<<<1, 1024, 0, s>>>(tiles, nTiles, stats, tileSurv, tileOut, dCount);
}

}  // namespace t3d
