// t3d_pose.cu -- bone poses evaluated on the device (the step before the geometry stage).
//
// Replaces, for armatures driven by an AnimationPlayer without relative motion (blends: k_pose_bones<true>, animation.go:852-907):
//   AnimationTrack.ValueAsVector / ValueAsQuaternion      animation.go:63-164
//   Quaternion.Slerp / ToMatrix4                          quaternion.go:26-59, 132-178
//   AnimationPlayer.forceUpdate's SetLocal* calls         animation.go:961-980
//   Node.Transform of every bone + boneInfluence          node.go:572-594
// so that a frame ships one 80-byte t3d_pose per armature instead of one matrix per bone.
//
// One thread per (armature, bone): sample the bone's three tracks at the armature's playhead, build the bone's
// local matrix Scale x Rotation x Translate, then walk the hierarchy level by level in shared memory
// (world = local x parent's world) and write inverseBind x world into the frame's bone arena.  Everything is
// float32, left to right, without multiply-add contraction, as the reference; Slerp's sine and arctangent are Go's
// float64 math.Sin / math.Atan2 (pure Go on amd64), restated operation by operation (t3d_pose.cuh).
// Batches whose shared memory fits ride as extra blocks of k_prep_draws (t3d_kernels.cu) instead of this launch.
#include "t3d_pose.cuh"
#include "t3d_internal.h"
#include "../../include/t3dgeo.h"

namespace t3d {

template <int MODE>
__global__ void __launch_bounds__(POSE_THREADS)
k_pose_bones(const __grid_constant__ PoseBatch pb, const __grid_constant__ PoseBlend bl) {
    extern __shared__ float poseSmem[];
    pose_block<MODE>(pb, bl, PoseChain(), blockIdx.x, poseSmem);
}
// t3d_submit_pose_chains: blends interrupted by further Play() calls
__global__ void __launch_bounds__(POSE_THREADS)
k_pose_chain(const __grid_constant__ PoseBatch pb, const __grid_constant__ PoseChain ch) {
    extern __shared__ float poseSmem[];
    pose_block<POSE_CHAIN>(pb, PoseBlend(), ch, blockIdx.x, poseSmem);
}

size_t pose_batch_smem(uint32_t nBones) { return pose_smem_bytes(nBones); }

cudaError_t launch_pose_bones(cudaStream_t s, const PoseBatch& pb, const PoseBlend* blend) {
    if (!pb.nPoses || !pb.nBones) return cudaSuccess;
    const uint32_t ipb = pose_armatures_per_block(pb.nBones);
    PoseBlend bl; memset(&bl, 0, sizeof bl);
    if (blend) {        // players in the middle of a blend (t3d_submit_poses_blended)
        bl = *blend;
        cudaFuncSetAttribute(k_pose_bones<POSE_BLEND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pose_smem_bytes(POSE_MAX_BONES));
        return launch_kernel(k_pose_bones<POSE_BLEND>, (pb.nPoses + ipb - 1) / ipb, POSE_THREADS, pose_smem_bytes(pb.nBones), s, false, pb, bl);
    }
    cudaFuncSetAttribute(k_pose_bones<POSE_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pose_smem_bytes(POSE_MAX_BONES));
    return launch_kernel(k_pose_bones<POSE_PLAIN>, (pb.nPoses + ipb - 1) / ipb, POSE_THREADS, pose_smem_bytes(pb.nBones), s, false, pb, bl);
}

cudaError_t launch_pose_chain(cudaStream_t s, const PoseBatch& pb, const PoseChain& ch) {
    if (!pb.nPoses || !pb.nBones) return cudaSuccess;
    const uint32_t ipb = pose_armatures_per_block(pb.nBones);
    cudaFuncSetAttribute(k_pose_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pose_chain_smem_bytes(POSE_MAX_BONES));
    return launch_kernel(k_pose_chain, (pb.nPoses + ipb - 1) / ipb, POSE_THREADS, pose_chain_smem_bytes(pb.nBones), s, false, pb, ch);
}

}  // namespace t3d
