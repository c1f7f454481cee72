// t3d_internal.h -- prototypes shared by the translation units of libt3dgeo.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "t3d_device.cuh"

namespace t3d {

// Launches with (pdl = true) or without the programmatic-dependent-launch attribute (t3d_device.cuh).
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// t3d_kernels.cu
void launch_prep_draws(cudaStream_t s, bool pdl, const void* hostSlim, const void* hostTmpl, const void* hostRot, const void* hostRest,
                       const t3d_light* hostLights, void* devDraws, void* devRest, uint32_t nDraws, size_t restBytes, uint32_t nLightsIn,
                       const FrameConst& fc, LightU* lightsU, DrawStat* stats, const uint32_t (*win)[4],
                       const void* hostBones, float* devBones, uint32_t nBones, bool bonesPacked, bool slimPacked,
                       uint32_t* devVbase, uint32_t* devTbase, const void* hostRuns, uint32_t runPayload16, const uint32_t* hostBlockRun, uint32_t tmplSpanMax,
                       const PoseBatch* poses, uint32_t nPoseBatches, uint32_t blockBase, uint32_t nBlocksThis, bool firstChunk, uint32_t* tileCtr);
#ifndef T3D_PREP_POSE_BATCHES
#define T3D_PREP_POSE_BATCHES 4
#endif
static const uint32_t PREP_POSE_BATCHES = T3D_PREP_POSE_BATCHES;       // t3d_submit_poses calls that can ride in k_prep_draws' grid; PREP_POSE_SMEM: their shared-memory budget
static const size_t PREP_POSE_SMEM = 40 * 1024;
static const int PREP_WINDOW_BLOCKS = 16, PREP_WINDOW_DRAWS = 64;     // k_prep_draws: blocks whose template windows travel as kernel parameters
// tileLists / nTilesOf are indexed by vertex-tile class: 0 generic, 1 static fast, 2 skinned fast (t3d_kernels.cu: K1_*)
void launch_vertex(cudaStream_t s, bool pdl, bool exactLighting, const DrawRec* draws, const uint32_t* vbase,
                   const uint4* const tileLists[3], const uint32_t nTilesOf[3], uint32_t* tileCtr /* [4] zeroed by k_prep_draws */,
                   uint32_t totalV, const FrameConst& fc, const LightU* lightsU, const float* bones,
                   const VolumeU* vols, float4* clip, float2* clipXY, float4* light, float4* wpos, float* unbill, float4* tnrm);
void launch_bake_valence(cudaStream_t s, const DrawRec* draws, const uint32_t* tbase, uint32_t nDraws, uint32_t totalT, uint32_t* valence);
int debug_vertex_trace(unsigned long long* out, uint32_t nBlocks);      // -DT3D_VERTEX_TRACE builds only (else -1)
static const uint32_t VERTEX_FAST_MAX_LIGHTS = 16;      // k_vertex<true> stages the light list: longer lists go to the generic kernel
void launch_cull_large(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                       const FrameConst& fc, const float4* clip, const float4* wpos, uint32_t* tdepth, DrawStat* stats, bool subdiv);
#ifndef T3D_FUSED_MINBLOCKS
#define T3D_FUSED_MINBLOCKS 4
#endif
static const uint32_t FUSED_RESIDENT_BLOCKS = T3D_FUSED_MINBLOCKS;     // k_cull_sort blocks resident per SM (its launch bounds)
void launch_cull_sort(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* fusedList, uint32_t nFused, uint32_t* ctr /* [2], zeroed by k_prep_draws */,
                      const FrameConst& fc, const float4* clip, const float4* wpos, DrawStat* stats, uint32_t* hist, uint32_t* sorted, bool subdiv,
                      uint32_t scanTiles, uint32_t* tileSurv, uint32_t* tileOut, uint32_t* chunkSum);
void launch_tile_packed(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* tbase, const TriTile* tiles,
                        const uint32_t* packedList, uint32_t nPacked, const FrameConst& fc, const float4* clip,
                        const float4* wpos, uint32_t* sorted, uint32_t* tileSurv, DrawStat* stats, bool subdiv);
void launch_draw_medium(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* mediumList, uint32_t nMedium,
                        const FrameConst& fc, const float4* clip, const float4* wpos, uint32_t* sorted, DrawStat* stats, bool subdiv);
void launch_emit(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, uint32_t nTiles, const uint32_t* sorted,
                 const uint32_t* tileSurv, const uint32_t* tileOut, const uint32_t* chunkSum, const FrameConst& fc, const float4* clip, const float2* clipXY,
                 const float4* light, const float* unbill, const float4* tnrm, float4* colorList,
                 float4* depthList, float4* normalList, uint32_t* triIds, bool compactDepth, float* custom0List,
                 uint32_t deliverIndexedTotalV /* compact emit only: 0, or the arena size when launch_emit_indexed runs too */, uint32_t deliverChunks);
void launch_emit_indexed(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* vbase, uint32_t nDraws, uint32_t totalV, const TriTile* tiles, uint32_t nTiles,
                         const uint32_t* sorted, const uint32_t* tileSurv, const uint32_t* tileOut, const uint32_t* chunkSum, uint32_t nChunks, const FrameConst& fc,
                         const float4* clip, const float2* clipXY, const float4* light, const float* unbill,
                         float4* rec32, float* dscal, float* cust, uint32_t* cornerSlot, uint32_t* triIds);
// delivery of sharded frames (t3d_gather_*): mailbox = 32 words per rank in the issuing rank's memory: [0] serial, [1] vertex count
// and [2] 1 when the frame arrived in the indexed form (written by the shard), [16] acknowledged serial, [17] block counter (written by the issuing rank)
static const uint32_t GATHER_MAIL_WORDS = 32;
void launch_gather_wait_ack(cudaStream_t s, bool pdl, const uint32_t* ackWord, uint32_t serial, uint32_t* errWord);
void launch_gather_post(cudaStream_t s, bool pdl, uint32_t* mail, const uint32_t* chunkSum, uint32_t nChunks, uint32_t serial,
                        const void* stats, void* statsOut, uint32_t nDraws, uint32_t totalV /* 0: the frame was not offered the indexed form */);
void launch_gather_expand(cudaStream_t s, bool pdl, uint32_t nShards, const void* const* rec32, const float* const* dscal, const float* const* cust, const uint32_t* const* cornerSlot,
                          t3d_vertex* const* color, t3d_vertex* const* depth, uint32_t* const* mail, uint32_t serial, bool renderDepth, bool pcuv, uint32_t* errWord);
static const uint32_t VERTEX_TILE = 256;        // vertex slots per k_vertex block (tileFirst granularity)

// t3d_sort.cu -- segmented 512-bin counting sort for draws that span several tiles
void launch_bin_hist(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge, uint32_t nTiles,
                     DrawStat* stats, uint32_t* tdepth, uint32_t* hist, uint32_t* tileSurv, uint32_t* tileOut, uint32_t* chunkSum);
void launch_hist_scan_huge(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* hugeList, uint32_t nHuge, uint32_t* hist);
void launch_scatter(cudaStream_t s, bool pdl, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                    const uint32_t* tkeys, const uint32_t* hist, uint32_t* sorted);
void launch_tile_scan(cudaStream_t s, bool pdl, const TriTile* tiles, uint32_t nTiles, const DrawStat* stats, uint32_t* tileSurv,
                      uint32_t* tileOut, uint32_t* chunkSum);

// t3d_bake.cu -- Model.BakeAO (same-model pass)
void launch_ao(cudaStream_t s, const float* pos, uint32_t vstride, const uint32_t* idx, const float* ctr, uint32_t tstride, uint32_t nTris, uint32_t nVerts,
               const float* triNormal, const float* triSpan, const uint8_t* triTarget, const uint8_t* vertTarget, float occlusionAngle,
               const float* occlusionColor4, uint8_t* marks, float4* colors);

// t3d_pose.cu -- bone poses evaluated on the device
cudaError_t launch_pose_bones(cudaStream_t s, const PoseBatch& pb, const PoseBlend* blend = nullptr);
cudaError_t launch_pose_chain(cudaStream_t s, const PoseBatch& pb, const PoseChain& ch);
size_t pose_batch_smem(uint32_t nBones);        // shared memory a block needs for a skeleton of nBones

}  // namespace t3d
