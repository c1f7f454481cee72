// t3d_internal.h -- prototypes shared by the translation units of libt3dgeo.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "t3d_device.cuh"

struct t3d_light;

namespace t3d {

// t3d_kernels.cu
void launch_prep_draws(cudaStream_t s, DrawRec* draws, uint32_t nDraws, const FrameConst& fc,
                       const t3d_light* lightsIn, LightU* lightsU, DrawStat* stats);
void launch_vertex(cudaStream_t s, const DrawRec* draws, const uint32_t* vbase, const uint32_t* tileFirst,
                   uint32_t totalGroups, const FrameConst& fc, const LightU* lightsU, const float* bones,
                   const VolumeU* vols, float4* clip, float4* light, float4* wpos, float* unbill, float4* tnrm);
void launch_cull(cudaStream_t s, const DrawRec* draws, const uint32_t* tbase, const uint32_t* tileFirst,
                 uint32_t totalTris, const FrameConst& fc, const float4* clip, const float4* wpos,
                 uint32_t* tdepth, uint32_t* tdraw, DrawStat* stats);
void launch_keys(cudaStream_t s, const DrawRec* draws, const DrawStat* stats, const uint32_t* tdepth,
                 const uint32_t* tdraw, uint32_t totalTris, uint32_t* keys);
void launch_emit(cudaStream_t s, const DrawRec* draws, const uint32_t* sortedKeys, const uint32_t* sortedTris,
                 const uint32_t* nSurvivors, uint32_t maxTris, const FrameConst& fc, const float4* clip,
                 const float4* light, const float* unbill, const float4* tnrm, float4* colorList,
                 float4* depthList, float4* normalList, uint32_t* triIds);

// t3d_sort.cu -- stable LSD radix sort of (key, triangle slot) pairs that also drops culled keys.
struct SortPlan {
    uint32_t maxItems;      // worst-case item count (all triangles survive)
    uint32_t nTiles;        // tiles of SORT_TILE items
    uint32_t nPasses;       // 8-bit digits needed for keyBits
    uint32_t keyBits;
};
static const uint32_t SORT_TILE = 4096;
SortPlan sort_plan(uint32_t maxItems, uint32_t nDraws);
size_t   sort_scratch_bytes(const SortPlan& p);      // histogram scratch
// keysIn[maxItems] may contain KEY_CULLED; the surviving pairs come out stably sorted in
// (*outKeys, *outVals), each pointing at one of the two ping-pong buffers.  *dCount (device) receives
// the survivor count.  Returns the number of kernels launched.
int sort_pairs(cudaStream_t s, const SortPlan& p, const uint32_t* keysIn, uint32_t* keysA, uint32_t* valsA,
               uint32_t* keysB, uint32_t* valsB, void* scratch, uint32_t* dCount,
               const uint32_t** outKeys, const uint32_t** outVals);

}  // namespace t3d
