// t3d_sort.cu -- stable LSD radix sort on quantised depth keys (sm_100a).
//
// The reference's sortingTriangleBucket (sortingTriangles.go:38-125) is a counting sort on a 9-bit
// bin with arrival order preserved inside a bin, run once per MeshPart.  Here all draws of a frame are
// sorted together: key = (draw ordinal << 9) | binKey, value = triangle slot.  Triangles enter in
// (draw, triangle index) order, the sort is stable and pass 0 drops culled keys while keeping order
// (order-preserving compaction), so inside every (draw, bin) group the reference's arrival order holds.
//
// One pass = k_hist (per-tile digit histogram) -> k_scan (exclusive scan over tiles per digit) ->
// k_scatter (stable in-tile ranking with warp match + scatter).  Integer work only.
#include "t3d_internal.h"

namespace t3d {

static const int SORT_THREADS = 256;
static const int SORT_ROUNDS = SORT_TILE / SORT_THREADS;
static const int RADIX = 256;

SortPlan sort_plan(uint32_t maxItems, uint32_t nDraws) {
    SortPlan p;
    p.maxItems = maxItems;
    p.nTiles = (maxItems + SORT_TILE - 1) / SORT_TILE;
    uint32_t drawBits = 0;
    while ((1ull << drawBits) < (uint64_t)nDraws) drawBits++;
    p.keyBits = BIN_BITS + drawBits;
    p.nPasses = (p.keyBits + 7) / 8;
    return p;
}

size_t sort_scratch_bytes(const SortPlan& p) {
    // hist[RADIX][nTiles] + binTotal[RADIX]
    return ((size_t)RADIX * (p.nTiles ? p.nTiles : 1) + RADIX) * sizeof(uint32_t);
}

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

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
k_hist(const uint32_t* __restrict__ keys, uint32_t nHost, const uint32_t* __restrict__ dCount, uint32_t shift,
       uint32_t nTiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sHist[RADIX];
    const uint32_t n = FIRST ? nHost : *dCount;
    const uint32_t tile = blockIdx.x;
    sHist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * SORT_TILE;
    if (base < n) {
#pragma unroll 4
        for (int r = 0; r < SORT_ROUNDS; r++) {
            const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
            if (i < n) {
                const uint32_t k = __ldg(&keys[i]);
                if (!FIRST || k != KEY_CULLED) atomicAdd(&sHist[(k >> shift) & (RADIX - 1)], 1u);
            }
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nTiles + tile] = sHist[threadIdx.x];
}

// One block per digit: exclusive scan of hist[digit][0..nTiles) in place; row total to binTotal[digit].
__global__ void __launch_bounds__(SORT_THREADS)
k_scan(uint32_t* __restrict__ hist, uint32_t nTiles, uint32_t* __restrict__ binTotal) {
    __shared__ uint32_t sWarp[8];
    uint32_t* row = hist + (size_t)blockIdx.x * nTiles;
    const uint32_t per = (nTiles + SORT_THREADS - 1) / SORT_THREADS;
    const uint32_t lo = threadIdx.x * per;
    const uint32_t hi = min(lo + per, nTiles);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += row[i];
    uint32_t total;
    uint32_t run = block_exclusive_scan_256(sum, sWarp, &total);
    for (uint32_t i = lo; i < hi; i++) { uint32_t c = row[i]; row[i] = run; run += c; }
    if (threadIdx.x == 0) binTotal[blockIdx.x] = total;
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
k_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint32_t nHost,
          uint32_t* __restrict__ dCount, uint32_t shift, uint32_t nTiles, const uint32_t* __restrict__ hist,
          const uint32_t* __restrict__ binTotal, uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut) {
    __shared__ uint32_t sBase[RADIX];
    __shared__ uint32_t sCnt[8][RADIX];
    __shared__ uint32_t sOff[8][RADIX];
    __shared__ uint32_t sWarp[8];
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // digit base = exclusive scan of the per-digit totals + this tile's offset inside the digit
    uint32_t total;
    const uint32_t bt = binTotal[threadIdx.x];
    const uint32_t binBase = block_exclusive_scan_256(bt, sWarp, &total);
    // n must be read before block 0 overwrites *dCount (FIRST pass defines it, later passes keep it)
    const uint32_t n = FIRST ? nHost : *dCount;
    if (FIRST && tile == 0 && threadIdx.x == 0) *dCount = total;
    sBase[threadIdx.x] = binBase + hist[(size_t)threadIdx.x * nTiles + tile];
#pragma unroll
    for (int w = 0; w < 8; w++) sCnt[w][threadIdx.x] = 0;
    __syncthreads();

    const uint32_t base = tile * SORT_TILE;
    if (base >= n) return;

    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint32_t i = base + r * SORT_THREADS + threadIdx.x;
        uint32_t key = KEY_CULLED, val = 0;
        bool valid = false;
        if (i < n) {
            key = __ldg(&keysIn[i]);
            valid = !FIRST || key != KEY_CULLED;
            val = FIRST ? i : __ldg(&valsIn[i]);
        }
        const uint32_t digit = valid ? ((key >> shift) & (RADIX - 1)) : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(full, digit);
        const uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rankInWarp == 0) sCnt[warp][digit] = __popc(peers);
        __syncthreads();
        {   // thread == digit: offsets of each warp's group, advance the running base, clear the counts
            uint32_t run = sBase[threadIdx.x];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t c = sCnt[w][threadIdx.x];
                sOff[w][threadIdx.x] = run;
                run += c;
                sCnt[w][threadIdx.x] = 0;
            }
            sBase[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = sOff[warp][digit] + rankInWarp;
            keysOut[pos] = key;
            valsOut[pos] = val;
        }
        // sOff is rewritten only after the next round's first barrier; sCnt was cleared above.
    }
}

int sort_pairs(cudaStream_t s, const SortPlan& p, const uint32_t* keysIn, uint32_t* keysA, uint32_t* valsA,
               uint32_t* keysB, uint32_t* valsB, void* scratch, uint32_t* dCount,
               const uint32_t** outKeys, const uint32_t** outVals) {
    int launches = 0;
    uint32_t* hist = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* binTotal = hist + (size_t)RADIX * (p.nTiles ? p.nTiles : 1);
    if (p.maxItems == 0) {
        cudaMemsetAsync(dCount, 0, sizeof(uint32_t), s);
        *outKeys = keysA; *outVals = valsA;
        return 0;
    }
    const uint32_t* curK = keysIn;
    const uint32_t* curV = nullptr;
    uint32_t* dstK = keysA; uint32_t* dstV = valsA;
    for (uint32_t pass = 0; pass < p.nPasses; pass++) {
        const uint32_t shift = 8 * pass;
        if (pass == 0) {
            k_hist<true><<<p.nTiles, SORT_THREADS, 0, s>>>(curK, p.maxItems, dCount, shift, p.nTiles, hist);
            k_scan<<<RADIX, SORT_THREADS, 0, s>>>(hist, p.nTiles, binTotal);
            k_scatter<true><<<p.nTiles, SORT_THREADS, 0, s>>>(curK, curV, p.maxItems, dCount, shift, p.nTiles, hist, binTotal, dstK, dstV);
        } else {
            k_hist<false><<<p.nTiles, SORT_THREADS, 0, s>>>(curK, p.maxItems, dCount, shift, p.nTiles, hist);
            k_scan<<<RADIX, SORT_THREADS, 0, s>>>(hist, p.nTiles, binTotal);
            k_scatter<false><<<p.nTiles, SORT_THREADS, 0, s>>>(curK, curV, p.maxItems, dCount, shift, p.nTiles, hist, binTotal, dstK, dstV);
        }
        launches += 3;
        curK = dstK; curV = dstV;
        if (dstK == keysA) { dstK = keysB; dstV = valsB; } else { dstK = keysA; dstV = valsA; }
    }
    *outKeys = curK; *outVals = curV;
    return launches;
}

}  // namespace t3d
