// t3d_scan.cuh -- where every tile's triangles start in the output lists: the exclusive scan of the per-tile survivor
// counts (shared by k_tile_scan / k_bin_hist in t3d_sort.cu and k_cull_sort in t3d_kernels.cu).
#pragma once
#include "t3d_device.cuh"

namespace t3d {

// One block.  Large tile j of a draw holds survivors [j * TILE_TRIS, ...) of the draw's sorted run; packed
// tiles wrote their own counts.  Thread t owns the contiguous run of tiles [t * per, (t + 1) * per): one pass
// to count, a block scan of the 1024 run totals, one pass to write the prefixes.
__device__ __forceinline__ uint32_t tile_survivors(const TriTile* __restrict__ tiles, const DrawStat* __restrict__ stats,
                                                   const uint32_t* __restrict__ tileSurv, uint32_t t) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(tiles + t));       // start, count, dFirst, info
    if (a.w & TILE_LARGE) {
        const uint32_t surv = __ldcg(&stats[a.z].survivors),      // (L2: in k_cull_sort other blocks of the same launch wrote it)
        off = (a.w & ~TILE_LARGE) * TILE_TRIS;
        return surv > off ? min(surv - off, TILE_TRIS) : 0u;
    }
    return tileSurv[t];
}

// The tile scan works in chunks of SCAN_CHUNK tiles: tileOut[t] is the exclusive prefix of the survivor counts
// inside t's chunk and chunkSum[c] the chunk's total, so the chunks are independent blocks and the emit adds
// the (few) earlier chunk totals itself.  One block of NT threads per chunk: thread t owns a contiguous run
// of tiles -- one pass to count, a block scan of the run totals, one pass to write the prefixes.
template <int NT>
__device__ __forceinline__ void tile_scan_chunk(const TriTile* __restrict__ tiles, uint32_t nTiles, uint32_t chunk,
                                                const DrawStat* __restrict__ stats, uint32_t* __restrict__ tileSurv,
                                                uint32_t* __restrict__ tileOut, uint32_t* __restrict__ chunkSum, uint32_t* sWarp /*[32]*/) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    const uint32_t cBegin = chunk * SCAN_CHUNK, cEnd = min(cBegin + SCAN_CHUNK, nTiles);
    const uint32_t per = (cEnd - cBegin + NT - 1) / NT;
    const uint32_t t0 = min(cBegin + threadIdx.x * per, cEnd), t1 = min(t0 + per, cEnd);
    uint32_t sum = 0;
    for (uint32_t t = t0; t < t1; t++) sum += tile_survivors(tiles, stats, tileSurv, t);
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(full, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < NW ? sWarp[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(full, wi, o); if (lane >= o) wi += n; }
        if (lane < NW) sWarp[lane] = wi - w;
        if (lane == 31) chunkSum[chunk] = wi;
    }
    __syncthreads();
    uint32_t run = sWarp[warp] + inc - sum;
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t c = tile_survivors(tiles, stats, tileSurv, t);
        tileOut[t] = run;
        tileSurv[t] = c;
        run += c;
    }
}


}  // namespace t3d
