// t3d_kernels.cu -- sm_100a kernels of the Tetra3D geometry stage.
//
//   k_prep_draws  four threads per draw     pulls the frame's uniforms over PCIe; per-draw uniforms of
//                                           Model.ProcessVertices (model.go:616-708) and light.beginModel
//                                           (light.go:175-189, 326-330)
//   k_vertex<M>   one vertex per thread     fused skin + transform + light (model.go:559-596, 757-822;
//                                           light.go:116-120, 191-258, 332-360, 583-663); static-fast, skinned-fast
//                                           and generic instantiations over host-built tile lists
//   k_cull_large  one tile per block        near/far, side, backface cull + depth key + per-draw min/max
//                                           (model.go:855-969) for draws of more than 8192 triangles
//   k_draw_medium one draw per block        the same cull fused with the 512-bin sort (sortingTriangles.go:38-125)
//                                           for draws of up to 8192 triangles and draws alone in their tile
//   k_tile_packed one tile per block        the same for tiles packed with several small draws
//   k_emit        one thread per corner     colour/depth/normal vertex lists (camera.go:1646-1820), written with
//                                           TMA bulk stores
//
// Mesh arrays are SoA planes (x|y|z, a|b|c) so one-element-per-thread loads are fully coalesced.
//
// Built with -fmad=false: no multiply-add contraction anywhere (see t3d_device.cuh).
#include "t3d_device.cuh"
#include "t3d_pose.cuh"
#include "t3d_scan.cuh"
#include "t3d_internal.h"
#include "../../include/t3dgeo.h"
#include <algorithm>
#include <cstring>

namespace t3d {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------

// Largest d in [lo, hi] with base[d] <= slot.  lo/hi come from the per-tile table the host builds,
// so for draws larger than a tile the loop body never runs.
__device__ __forceinline__ uint32_t find_draw(const uint32_t* __restrict__ base, uint32_t lo, uint32_t hi, uint32_t slot) {
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&base[mid]) <= slot) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------
// K0: per-draw uniforms
// ------------------------------------------------------------------------------------------------
// The frame's uniforms sit in one pinned host block (t3d_api.cu packs it: DrawSlim[nDraws], the DrawRec templates,
// the world rotations, then the tables, light lists and bone matrices).  This kernel pulls the block over PCIe
// itself (UVA loads from the mapped host allocation, all of a thread's loads in flight at once) instead of
// waiting for a DMA copy: blocks < nPrepBlocks stage PREP_DRAWS slim records plus the templates and rotations
// they name in shared memory, rebuild the full records there, derive the per-draw uniforms and write the
// finished records to the device; the remaining blocks copy the rest of the block verbatim.
static const int PREP_THREADS = 256;
static_assert(PREP_THREADS == POSE_THREADS, "pose blocks ride in k_prep_draws' grid");
static const int PREP_DRAWS = 64;               // draws per block
static const int PREP_TEMPLATES = 80;           // templates staged per block: the index window [tmin, tmin + 80) of its draws
static const uint32_t PREP_SMEM_LIGHTS = 256;   // light lists up to this many entries are staged in shared memory
static const uint32_t PREP_SORT_LIGHTS = 64;    // default lighting: draws with up to this many lights get their LightU list grouped by type
static const int PREP_PARAM_BLOCKS = PREP_WINDOW_BLOCKS;
static_assert(PREP_DRAWS == PREP_WINDOW_DRAWS, "t3d_internal.h: the host computes the windows per PREP_DRAWS draws");
static const int REC16 = sizeof(DrawRec) / 16;
static const int SLIM16 = sizeof(DrawSlim) / 16;
// shared memory: records | slim records | rotations (run mode parks the runs here first) | templates [tmplCap] | lights
// [lightCap]; the last two are sized per frame, so frames with few templates and lights run more blocks per SM and keep
// more PCIe reads in flight
static const size_t PREP_SMEM_FIXED = (size_t)PREP_DRAWS * sizeof(DrawRec) + (size_t)PREP_DRAWS * sizeof(DrawSlim) + (size_t)PREP_DRAWS * 64;
static const size_t PREP_SMEM_MAX = PREP_SMEM_FIXED + (size_t)PREP_TEMPLATES * sizeof(DrawRec) + (size_t)PREP_SMEM_LIGHTS * sizeof(t3d_light);

struct PrepFrame {
    const uint4* hostSlim;      // DrawSlim[nDraws]          (pinned host, UVA)
    const uint4* hostTmpl;      // DrawRec[nTemplates]
    const uint4* hostRot;       // float[16][nRotations]
    const uint4* hostRest;      // tables, lights, bones ...
    const t3d_light* hostLights;
    uint4* devDraws;            // DrawRec[nDraws]           (device)
    uint4* devRest;
    uint32_t nDraws, rest16, nPrepBlocks, nLightsIn;
    // template / rotation windows {tmin, templates, rmin, rotations} of the first blocks, known to the host: those
    // blocks (all of a small frame) fetch slim records, templates and rotations in ONE round trip over PCIe
    uint32_t win[PREP_PARAM_BLOCKS][4];
    // bone matrices: 48 bytes each on the host when bonesPacked (rows 0-3, columns 0-2: column 3 is exactly 0,0,0,1),
    // else 64; always 64 on the device
    const uint4* hostBones; uint4* devBones; uint32_t nBones, bonesPacked;
    // DrawSlim records: 7 uint4 each, or 6 when every model transform of the frame is affine (the transform travels as
    // rows 0-3 x columns 0-2; column 3 is exactly 0,0,0,1)
    uint32_t slim16;
    uint32_t* devVbase; uint32_t* devTbase;     // [nDraws + 1] arena offsets per draw: written here from the records, not copied
    // run-length mode (t3d_device.cuh: DrawRun): hostSlim then holds 64-byte payloads {transform rows, colour}
    const uint4* hostRuns; const uint32_t* hostBlockRun; uint32_t runMode;
    uint32_t pay16;                 // run mode: uint4 per payload -- 4, or 2 in a frame of nothing but uniformly scaled instances (t3d_submit_instances)
    uint32_t tmplCap, lightCap;     // templates / lights staged per block (<= PREP_TEMPLATES, PREP_SMEM_LIGHTS)
    uint32_t copyEnd;               // this launch's blocks [nPrepBlocks, copyEnd) copy (behind the pose blocks of k_prep_draws<true>)
    uint32_t* tileCtr;              // first launch of the frame: k_vertex's tile counters, zeroed here
    uint32_t blockBase;             // first 64-draw block of this launch: a large staged frame is prepared in chunks, each behind its slice's copy
};
// Bone poses riding along (t3d_pose.cuh): blocks [poseBlock[i], poseBlock[i + 1]) of the grid evaluate batch i.  Only frames
// that pose armatures on the device launch the instantiation that carries this parameter and the pose code: on a frame of
// 50 draws the combined kernel was 4 us slower even when no block took the branch (measured A/B; not explained by the
// parameter bytes or the code length alone).
template <bool POSES> struct PrepPoses { PoseBatch pose[PREP_POSE_BATCHES]; uint32_t poseBlock[PREP_POSE_BATCHES + 1]; uint32_t nPoseBatches; };
template <> struct PrepPoses<false> { uint32_t nPoseBatches; };

#ifndef T3D_PREP_MINBLOCKS
#define T3D_PREP_MINBLOCKS 3
#endif
template <bool POSES>
__global__ void __launch_bounds__(PREP_THREADS, T3D_PREP_MINBLOCKS)
k_prep_draws(const __grid_constant__ PrepFrame pf, const __grid_constant__ FrameConst fc,
             LightU* __restrict__ lightsU, DrawStat* __restrict__ stats, const __grid_constant__ PrepPoses<POSES> pp) {
    extern __shared__ uint4 prepSmem[];
    pdl_launch();
    if (pf.tileCtr && blockIdx.x == 0 && threadIdx.x < 8) pf.tileCtr[threadIdx.x] = 0;      // k_vertex's tile counters (one per class), k_cull_sort's ticket and arrival counters
    // Pose blocks come FIRST in the grid (k_prep_draws<true>): they start at once and their float64 arithmetic overlaps the
    // PCIe round trips of the prep blocks, instead of trailing them as a second wave.
    uint32_t bid = blockIdx.x;
    if constexpr (POSES) {
        const uint32_t nPoseBlocks = pp.poseBlock[pp.nPoseBatches];
        if (bid < nPoseBlocks) {
            uint32_t i = 0;
            while (i + 1 < pp.nPoseBatches && bid >= pp.poseBlock[i + 1]) i++;
            pose_block<POSE_PLAIN>(pp.pose[i], PoseBlend(), PoseChain(), bid - pp.poseBlock[i], reinterpret_cast<float*>(prepSmem));
            return;
        }
        bid -= nPoseBlocks;
    }
    if (bid >= pf.nPrepBlocks) {
        const uint32_t nb = pf.copyEnd - pf.nPrepBlocks;      // copy blocks
        const uint32_t stride = nb * PREP_THREADS;
        const uint4* __restrict__ src = pf.hostRest;
        uint4* __restrict__ dst = pf.devRest;
        uint32_t i = (bid - pf.nPrepBlocks) * PREP_THREADS + threadIdx.x;
        for (; i + 7 * stride < pf.rest16; i += 8 * stride) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = src[i + u * stride];
#pragma unroll
            for (int u = 0; u < 8; u++) dst[i + u * stride] = v[u];
        }
        for (; i < pf.rest16; i += stride) dst[i] = src[i];
        // bone matrices, one per thread: 3 (packed) or 4 uint4 in, 4 out
        for (uint32_t b = (bid - pf.nPrepBlocks) * PREP_THREADS + threadIdx.x; b < pf.nBones; b += stride) {
            uint4* o = pf.devBones + 4 * (size_t)b;
            if (pf.bonesPacked) {
                const uint4* p = pf.hostBones + 3 * (size_t)b;
                const uint4 a = p[0], c = p[1], e = p[2];       // m00 m01 m02 m10 | m11 m12 m20 m21 | m22 m30 m31 m32
                o[0] = make_uint4(a.x, a.y, a.z, 0u);
                o[1] = make_uint4(a.w, c.x, c.y, 0u);
                o[2] = make_uint4(c.z, c.w, e.x, 0u);
                o[3] = make_uint4(e.y, e.z, e.w, 0x3F800000u);
            } else {
                const uint4* p = pf.hostBones + 4 * (size_t)b;
                const uint4 a = p[0], c = p[1], e = p[2], f = p[3];
                o[0] = a; o[1] = c; o[2] = e; o[3] = f;
            }
        }
        return;
    }
    DrawRec* sR = reinterpret_cast<DrawRec*>(prepSmem);
    uint4* sSraw = prepSmem + PREP_DRAWS * REC16;
    const uint32_t s16 = pf.slim16;                      // record stride; the 32-byte header is the same in both layouts
    auto slimOf = [&](uint32_t kk) -> const DrawSlim& { return *reinterpret_cast<const DrawSlim*>(sSraw + kk * s16); };
    uint4* sRot = prepSmem + PREP_DRAWS * (REC16 + SLIM16);
    uint4* sT = sRot + PREP_DRAWS * 4;
    t3d_light* sL = reinterpret_cast<t3d_light*>(sT + pf.tmplCap * REC16);
    bid += pf.blockBase;
    const uint32_t d0 = bid * PREP_DRAWS;
    const uint32_t n = min((uint32_t)PREP_DRAWS, pf.nDraws - d0);
    constexpr int T_SWEEPS = (PREP_TEMPLATES * REC16 + PREP_THREADS - 1) / PREP_THREADS;
    const bool known = !pf.runMode && bid < PREP_PARAM_BLOCKS;      // windows in the kernel parameters
    uint32_t tmin = known ? pf.win[bid][0] : 0u, rmin = known ? pf.win[bid][2] : NO_ROTATION;
    const bool lightsStaged = pf.nLightsIn <= pf.lightCap;
    if (pf.runMode) {
        // Round trip 1: this block's payloads (4 uint4 per draw: three of transform rows, one of colour) go straight into
        // the packed record layout {header 2 | transform 3 | colour 1}; the runs covering the block; the light table.
        const uint32_t pay16 = pf.pay16;
        const uint4* src = pf.hostSlim + (size_t)d0 * pay16;
        const uint32_t r0 = pf.hostBlockRun[bid], r1 = pf.hostBlockRun[bid + 1];
        uint4 pv;
        if (threadIdx.x < n * pay16) pv = src[threadIdx.x];
        const uint32_t* lsrc = reinterpret_cast<const uint32_t*>(pf.hostLights);
        uint32_t* ldst = reinterpret_cast<uint32_t*>(sL);
        const uint32_t nw = lightsStaged ? pf.nLightsIn * (uint32_t)(sizeof(t3d_light) / 4) : 0u;
        for (uint32_t i = threadIdx.x; i < nw; i += PREP_THREADS) ldst[i] = lsrc[i];
        if (threadIdx.x < n * pay16) {      // 64-byte payloads fill slots 2-5; a 32-byte instance payload is {position, scale | colour}: slots 2 and 5
            const uint32_t kk = pay16 == 4 ? threadIdx.x >> 2 : threadIdx.x >> 1, j = pay16 == 4 ? threadIdx.x & 3 : threadIdx.x & 1;
            sSraw[kk * s16 + (pay16 == 4 ? 2 + j : j ? 5 : 2)] = pv;
        }
        // Round trip 2: the runs [r0, r1] (at most one per draw), parked where the rotations will go
        const uint32_t nR = min(r1 - r0 + 1, (uint32_t)PREP_DRAWS);       // a 65th run could only start at the next block's first draw
        if (threadIdx.x < nR * 4) sRot[threadIdx.x] = pf.hostRuns[(size_t)r0 * 4 + threadIdx.x];
        __syncthreads();
        if (threadIdx.x < n) {      // header of draw d from its run
            const uint32_t d = d0 + threadIdx.x;
            const DrawRun* runs = reinterpret_cast<const DrawRun*>(sRot);
            uint32_t r = 0;
            while (r + 1 < nR && runs[r + 1].first <= d) r++;
            const DrawRun& U = runs[r];
            const uint32_t i = d - U.first;
            uint32_t* h = reinterpret_cast<uint32_t*>(sSraw + threadIdx.x * s16);      // DrawSlim header: tmpl vbase tbase boneOffset | lightU rot pad pad
            h[0] = U.tmpl; h[1] = U.vbase + i * U.dV; h[2] = U.tbase + i * U.dT; h[3] = U.boneOffset + i * U.dBone;
            h[4] = U.lightU + i * U.dLightU; h[5] = U.dRot ? U.rot + i : U.rot; h[6] = 0; h[7] = 0;
            if (U.pad[0]) {
                // An instance (t3d_submit_instances): Node.Transform (node.go:577-583) from its position and scale and the run's
                // rotation and parent transform -- Scale . rotation . Translate [. parent], each step the reference's Mult --
                // written where an ordinary draw's 12 transform floats arrive.
                float* f = reinterpret_cast<float*>(sSraw + threadIdx.x * s16 + 2);
                const float px = f[0], py = f[1], pz = f[2], sx = f[3];
                const float sy = pay16 == 4 ? f[4] : sx, sz = pay16 == 4 ? f[5] : sx;
                const float4* rp = reinterpret_cast<const float4*>(pf.hostRot) + (size_t)U.pad[1] * 4;
                Mat4 rot;
#pragma unroll
                for (int r = 0; r < 4; r++) { const float4 v = rp[r]; rot.m[r][0] = v.x; rot.m[r][1] = v.y; rot.m[r][2] = v.z; rot.m[r][3] = v.w; }
                Mat4 M = m4_mult(m4_scale(sx, sy, sz), rot);
                Mat4 T = m4_identity(); T.m[3][0] = px; T.m[3][1] = py; T.m[3][2] = pz;
                M = m4_mult(M, T);
                if (U.pad[2] != NO_ROTATION) {
                    const float4* pp = reinterpret_cast<const float4*>(pf.hostRot) + (size_t)U.pad[2] * 4;
                    Mat4 par;
#pragma unroll
                    for (int r = 0; r < 4; r++) { const float4 v = pp[r]; par.m[r][0] = v.x; par.m[r][1] = v.y; par.m[r][2] = v.z; par.m[r][3] = v.w; }
                    M = m4_mult(M, par);
                }
#pragma unroll
                for (int r = 0; r < 4; r++) { f[3 * r] = M.m[r][0]; f[3 * r + 1] = M.m[r][1]; f[3 * r + 2] = M.m[r][2]; }
                if (U.pad[0] & 0x80000000u) {       // culling bounds follow the run's matrices: Model.onTransformUpdate + Camera.SphereInFrustum
                    const float4* bp = reinterpret_cast<const float4*>(pf.hostRot) + (size_t)((U.pad[2] != NO_ROTATION ? U.pad[2] : U.pad[1]) + 1) * 4;
                    const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2];
                    const float b9[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x};
                    FrustumTerms F;
                    F.camPos = v3(fc.camPos[0], fc.camPos[1], fc.camPos[2]); F.forward = v3(fc.camForward[0], fc.camForward[1], fc.camForward[2]);
                    F.right = v3(fc.camRight[0], fc.camRight[1], fc.camRight[2]); F.up = v3(fc.camUp[0], fc.camUp[1], fc.camUp[2]);
                    F.nearP = fc.nearP; F.farP = fc.farP; F.sphereFactorY = fc.sphereFactorY; F.sphereFactorX = fc.sphereFactorX; F.sphereFactorTang = fc.sphereFactorTang;
                    F.aspect = fc.camW / fc.camH; F.orthoScale = fc.orthoScale; F.perspective = (fc.camFlags & T3D_CAM_PERSPECTIVE) != 0;
                    if (!m4_instance_in_frustum(M, b9, F)) h[6] = 1u;        // picked up below: T3D_DRAW_FRUSTUM_CULLED
                }
            }
        }
        __syncthreads();            // the runs are consumed: the rotation window may overwrite them
    } else
    {   // slim records (7 uint4 each), and for the first blocks the templates and rotations in the same round trip
        const uint4* src = pf.hostSlim + (size_t)d0 * s16;
        uint4* dst = sSraw;
        const uint32_t n16 = n * s16;
        uint4 v[2], tv[T_SWEEPS], rv;
#pragma unroll
        for (int u = 0; u < 2; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < n16) v[u] = src[i]; }
        const uint32_t t16 = known ? min(pf.win[bid][1], pf.tmplCap) * REC16 : 0u, r16 = known ? min(pf.win[bid][3], (uint32_t)PREP_DRAWS) * 4 : 0u;
        const uint4* tsrc = pf.hostTmpl + (size_t)tmin * REC16;
#pragma unroll
        for (int u = 0; u < T_SWEEPS; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < t16) tv[u] = tsrc[i]; }
        if (threadIdx.x < r16) rv = pf.hostRot[(size_t)rmin * 4 + threadIdx.x];
        // the light table rides in the same round trip (its first 512 words; longer tables finish below)
        const uint32_t* lsrc = reinterpret_cast<const uint32_t*>(pf.hostLights);
        uint32_t* ldst = reinterpret_cast<uint32_t*>(sL);
        const uint32_t nw = lightsStaged ? pf.nLightsIn * (uint32_t)(sizeof(t3d_light) / 4) : 0u;
        uint32_t lv[2];
#pragma unroll
        for (int u = 0; u < 2; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < nw) lv[u] = lsrc[i]; }
#pragma unroll
        for (int u = 0; u < 2; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < n16) dst[i] = v[u]; }
#pragma unroll
        for (int u = 0; u < T_SWEEPS; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < t16) sT[i] = tv[u]; }
        if (threadIdx.x < r16) sRot[threadIdx.x] = rv;
#pragma unroll
        for (int u = 0; u < 2; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < nw) ldst[i] = lv[u]; }
        for (uint32_t i = 2 * PREP_THREADS + threadIdx.x; i < nw; i += PREP_THREADS) ldst[i] = lsrc[i];
        static_assert(2 * PREP_THREADS >= PREP_DRAWS * SLIM16, "staging covers a block's slim records in one sweep");
    }
    const t3d_light* lightsIn = lightsStaged ? sL : pf.hostLights;
    __syncthreads();
    static_assert(PREP_THREADS >= PREP_DRAWS * 4, "one sweep stages a block's payloads and runs");
    if (!known) {
        // later blocks find their windows in the slim records: a window of template indices (the host reuses one of
        // its most recent templates when it can, so the window is small) and an ascending run of rotation indices
        uint32_t tmax = 0u, rmax = 0u;
        tmin = 0xFFFFFFFFu;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t t = slimOf(k).tmpl, r = slimOf(k).rot;
            tmin = min(tmin, t); tmax = max(tmax, t);
            if (r != NO_ROTATION) { rmin = min(rmin, r); rmax = max(rmax, r); }
        }
        const uint32_t n16 = min(tmax - tmin + 1, pf.tmplCap) * REC16;
        const uint4* src = pf.hostTmpl + (size_t)tmin * REC16;
        uint4 v[T_SWEEPS];
#pragma unroll
        for (int u = 0; u < T_SWEEPS; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < n16) v[u] = src[i]; }
#pragma unroll
        for (int u = 0; u < T_SWEEPS; u++) { const uint32_t i = u * PREP_THREADS + threadIdx.x; if (i < n16) sT[i] = v[u]; }
        if (rmin != NO_ROTATION) {
            // one rotation per draw at most, but instances share theirs and their runs' matrices sit in between: the window holds
            // PREP_DRAWS rotations, a draw beyond it reads its rotation from the frame block
            const uint32_t r16 = min(rmax - rmin + 1, (uint32_t)PREP_DRAWS) * 4;
            if (threadIdx.x < r16) sRot[threadIdx.x] = pf.hostRot[(size_t)rmin * 4 + threadIdx.x];
        }
        __syncthreads();
    }
    // Four threads per draw (role = warp pair, so the roles run on different schedulers): the serial chain of one
    // draw's uniform derivation is cut into independent pieces, each evaluated exactly as before.
    const uint32_t role = threadIdx.x / PREP_DRAWS, k = threadIdx.x % PREP_DRAWS;
    static_assert(PREP_THREADS == 4 * PREP_DRAWS, "four roles per draw");
    const bool mine = k < n;
    if (mine) {   // full record = template + this draw's placement, assembled by the draw's four threads
        const DrawSlim& S = slimOf(k);
        uint4* dst = reinterpret_cast<uint4*>(&sR[k]);
        const uint32_t ti = S.tmpl - tmin;
        const uint4* src = ti < pf.tmplCap ? sT + ti * REC16 : pf.hostTmpl + (size_t)S.tmpl * REC16;
        const uint4* sl = reinterpret_cast<const uint4*>(&S);
        const uint4* rs = S.rot == NO_ROTATION ? nullptr : S.rot - rmin < (uint32_t)PREP_DRAWS ? sRot + (size_t)(S.rot - rmin) * 4 : pf.hostRot + (size_t)S.rot * 4;
        constexpr int TR16 = offsetof(DrawRec, transform) / 16, WR16 = offsetof(DrawRec, worldRot) / 16, CO16 = offsetof(DrawRec, color) / 16;
        constexpr int STR16 = offsetof(DrawSlim, transform) / 16;
        const bool packed = s16 != (uint32_t)SLIM16;
        for (int i = role; i < REC16; i += 4) {
            uint4 v = src[i];
            if (i >= TR16 && i < TR16 + 4) {
                if (!packed) v = sl[STR16 + (i - TR16)];
                else {      // row r of the transform from the 12 packed floats
                    const int r = i - TR16;
                    const uint32_t* f = reinterpret_cast<const uint32_t*>(sl + STR16) + 3 * r;
                    v = make_uint4(f[0], f[1], f[2], r == 3 ? 0x3F800000u : 0u);
                }
            }
            else if (i == CO16) v = sl[STR16 + (packed ? 3 : 4)];
            else if (rs && i >= WR16 && i < WR16 + 4) v = rs[i - WR16];
            dst[i] = v;
        }
    }
    __syncthreads();
    if (mine && role == 0) {      // the scalar placement fields, and the frame's vbase / tbase tables
        const DrawSlim& S = slimOf(k);
        DrawRec& R = sR[k];
        R.vbase = S.vbase; R.tbase = S.tbase; R.lightU = S.lightU;
        if (S.pad[0] == 1u) R.flags |= T3D_DRAW_FRUSTUM_CULLED;      // an instance outside the frustum (decided above, or by the host for frames not shipped as runs)
        // device-evaluated poses (t3d_submit_poses) sit behind the submitted matrices in the bone arena
        R.boneOffset = (S.boneOffset & POSE_TAG) ? pf.nBones + (S.boneOffset & ~POSE_TAG) : S.boneOffset;
        pf.devVbase[d0 + k] = S.vbase; pf.devTbase[d0 + k] = S.tbase;
        if (d0 + k == pf.nDraws - 1) { pf.devVbase[pf.nDraws] = S.vbase + R.nv; pf.devTbase[pf.nDraws] = S.tbase + R.nt; }
    }
    __syncthreads();
    float* sInv = reinterpret_cast<float*>(sSraw);       // the slim records are consumed: 16 floats per draw {r rows 0-3, invPS}
    const uint32_t invStride = s16 * 4;                  // floats; >= 24
    if (mine) {
    DrawRec& R = sR[k];
    const uint32_t flags = R.flags;
    const bool skinned = flags & T3D_DRAW_SKINNED;
    const Mat4 modelT = m4_load(R.transform);
    const Vec3 camPos = v3(fc.camPos[0], fc.camPos[1], fc.camPos[2]);
    if (role == 0) {
        // model.go:629-630
        Vec3 p, sc; Mat4 r;
        { Mat4 inv = m4_inverted(modelT); m4_decompose(inv, p, sc, r); }
        const Vec3 invPS = v3_mult(p, v3(1.0f / sc.X, 1.0f / sc.Y, 1.0f / sc.Z));
        const Vec3 invCam = v3_add(m4_multvec(r, camPos), invPS);
        R.invCam[0] = invCam.X; R.invCam[1] = invCam.Y; R.invCam[2] = invCam.Z;
        float* o = sInv + invStride * k;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) o[i * 3 + j] = r.m[i][j];       // rows 0-3, columns 0-2: all m4_multvec reads
        o[12] = invPS.X; o[13] = invPS.Y; o[14] = invPS.Z;
        DrawStat st; st.minBits = (flags & T3D_DRAW_FRUSTUM_CULLED) ? DRAW_CULLED_MARK : 0x7F7FFFFFu; st.maxBits = 0u; st.survivors = 0; st.done = 0;   // model.go:710-711
        stats[d0 + k] = st;
    } else if (role == 1) {
        const Mat4 vp = m4_load(fc.vp);
        Mat4 base = modelT;
        if (flags & T3D_DRAW_BILLBOARD) {                               // model.go:634-693
            Vec3 up = v3(0, 1, 0);
            if (flags & T3D_DRAW_BILLBOARD_CAMERA_UP) up = v3(fc.camUp[0], fc.camUp[1], fc.camUp[2]);
            Mat4 lookat = m4_lookat(v3(modelT.m[3][0], modelT.m[3][1], modelT.m[3][2]), camPos, up);
            if (flags & T3D_DRAW_BILLBOARD_LOCK_X) {
                lookat.m[0][0] = 1; lookat.m[0][1] = 0; lookat.m[0][2] = 0; lookat.m[0][3] = 0;
                lookat.m[1][0] = 0; m4_unit_row(lookat, 1);
                lookat.m[2][0] = 0; m4_unit_row(lookat, 2);
            }
            if (flags & T3D_DRAW_BILLBOARD_LOCK_Y) {
                lookat.m[1][0] = 0; lookat.m[1][1] = 1; lookat.m[1][2] = 0; lookat.m[1][3] = 0;
                lookat.m[0][1] = 0; m4_unit_row(lookat, 0);
                lookat.m[2][1] = 0; m4_unit_row(lookat, 2);
            }
            if (flags & T3D_DRAW_BILLBOARD_LOCK_Z) {
                lookat.m[2][0] = 0; lookat.m[2][1] = 0; lookat.m[2][2] = 1; lookat.m[2][3] = 0;
                lookat.m[0][2] = 0; m4_unit_row(lookat, 0);
                lookat.m[1][2] = 0; m4_unit_row(lookat, 1);
            }
            Vec3 bp, bs; Mat4 br;
            m4_decompose(base, bp, bs, br);
            base = m4_mult(m4_mult(br, m4_scale(bs.X, bs.Y, bs.Z)), lookat);
            base.m[3][0] = bp.X; base.m[3][1] = bp.Y; base.m[3][2] = bp.Z; base.m[3][3] = 1;
        }
        const Mat4 mvp = m4_mult(base, vp);                             // model.go:695
        m4_store(mvp, R.mvp);
    } else if (role == 2) {
        if (flags & T3D_DRAW_DEPTH_UNBILLBOARDED) {                     // model.go:700-703
            const Mat4 u = m4_mult(modelT, m4_load(fc.vp));
            R.unbill[0] = u.m[0][2]; R.unbill[1] = u.m[1][2]; R.unbill[2] = u.m[2][2]; R.unbill[3] = u.m[3][2];
        }
        if (fc.camFlags & T3D_CAM_RENDER_NORMALS) {                     // model.go:705-708
            const Mat4 mv = m4_mult(modelT, m4_load(fc.view));
            const Mat4 rot = m4_decompose_rotation(mv);
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 4; rr++) R.nrmRot[c * 4 + rr] = rot.m[rr][c];
        }
    } else {
        // The sun's workingModelRotation (light.go:326-330) is per model: parked in worldRot, in place, when the
        // draw's list holds a directional light.
        const uint32_t nL = (flags & T3D_DRAW_LIT) ? R.lightTotal : 0;      // the lights LightVolume cells use included
        bool haveSun = false;
        for (uint32_t i = 0; i < nL && !haveSun; i++) haveSun = lightsIn[R.lightIn + i].type == T3D_LIGHT_DIRECTIONAL;
        if (haveSun && !skinned) {
            const Mat4 sunRot = m4_transposed(m4_inverted(m4_load(R.worldRot)));
            m4_store(sunRot, R.worldRot);
        }
    }
    }
    __syncthreads();
    if (mine) {
        // Lights: beginModel products + per-draw skip (camera.go:1625-1629, light.go:671-673), the draw's four
        // threads taking every fourth light.
        const DrawRec& R = sR[k];
        const uint32_t flags = R.flags;
        const bool skinned = flags & T3D_DRAW_SKINNED;
        const uint32_t nL = (flags & T3D_DRAW_LIT) ? R.lightTotal : 0;
        const uint32_t nOwn = R.lightCount;         // entries beyond are lights that LightVolume cells use (light.go:652-654)
        // Default lighting: the list goes out grouped by type [point | directional | ambient | skipped] so that k_vertex's fast
        // light loops need no type dispatch (the sums then run in a different order: colours stay within the 1/255 tolerance;
        // t3d_set_exact_lighting keeps submit order).  Every thread of the draw classifies the whole list itself -- two passes
        // over at most PREP_SORT_LIGHTS staged records -- which needs no shared scratch and no barrier.
        const Vec3 modelPos = v3(R.transform[12], R.transform[13], R.transform[14]);
        const float maxSpan = R.maxSpan;
        auto skipped = [&](const t3d_light& L, uint32_t i) -> bool {        // camera.go:1625-1629: only for the list's own lights
            if (!(L.type == T3D_LIGHT_POINT && L.range > 0 && i < nOwn) || fc.bakeValence) return false;     // (BakeLighting lights every vertex with every light)
            const float dist = maxSpan + L.range;
            return v3_dist2(modelPos, v3(L.position[0], L.position[1], L.position[2])) > dist * dist;
        };
        const bool approx = fc.approxLighting != 0;
        bool sortable = approx && lightsStaged && nL == nOwn && nL <= PREP_SORT_LIGHTS;
        uint32_t nP = 0, nS = 0, nA = 0;
        if (sortable) {
            for (uint32_t i = 0; i < nL; i++) {
                const t3d_light& L = lightsIn[R.lightIn + i];
                if (L.type == T3D_LIGHT_VOLUME) { sortable = false; break; }
                if (skipped(L, i)) continue;
                nP += L.type == T3D_LIGHT_POINT; nS += L.type == T3D_LIGHT_DIRECTIONAL; nA += L.type == T3D_LIGHT_AMBIENT;
            }
        }
        if (role == 0) {
            DrawRec& W = sR[k];
            W.nPoint = sortable ? (uint16_t)nP : 0; W.nSun = sortable ? (uint16_t)nS : 0; W.nAmbient = sortable ? (uint16_t)nA : 0;
            W.lightsSorted = sortable ? 1 : 0;
        }
        if (role < nL) {
            const float* o = sInv + invStride * k;
            Mat4 r = m4_identity();
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) r.m[i][j] = o[i * 3 + j];
            const Vec3 invPS = v3(o[12], o[13], o[14]);
            uint32_t cP = 0, cS = 0, cA = 0, cX = 0, seen = 0;      // lights of each group before entry `seen`
            for (uint32_t i = role; i < nL; i += 4) {
                const t3d_light L = lightsIn[R.lightIn + i];
                LightU U;
                U.type = L.type;
                U.color[0] = L.color[0]; U.color[1] = L.color[1]; U.color[2] = L.color[2];
                U.energy = L.energy;
                U.wpos[0] = U.wpos[1] = U.wpos[2] = 0;
                U.range = L.range; U.range2 = L.range * L.range;         // light.go:277-280
                U.fwd[0] = L.forward[0]; U.fwd[1] = L.forward[1]; U.fwd[2] = L.forward[2];
                U.volume = L.volume & 0xFFFu; U.useFirst = L.volume >> 12; U.pad = 0;
                const bool own = i < nOwn;
                if (L.type == T3D_LIGHT_POINT) {
                    const Vec3 lp = v3(L.position[0], L.position[1], L.position[2]);
                    if (skipped(L, i)) U.type = LIGHT_SKIP;
                    const Vec3 wp = skinned ? lp : v3_add(m4_multvec(r, lp), invPS);   // light.go:183-187
                    U.wpos[0] = wp.X; U.wpos[1] = wp.Y; U.wpos[2] = wp.Z;
                    U.fwd[0] = U.fwd[1] = U.fwd[2] = 0;
                    if (approx && own) {        // see LightU: the skip distance (forward[0]: t3d_submit_lights), the falloff's range, colour * energy * 2
                        U.range = L.range > 0 ? U.range2 : L.forward[0];
                        U.range2 = U.range2 > 0 ? U.range2 : __int_as_float(0x7F800000);
                        U.fwd[0] = L.color[0] * L.energy * 2.0f; U.fwd[1] = L.color[1] * L.energy * 2.0f; U.fwd[2] = L.color[2] * L.energy * 2.0f;
                    }
                } else if (L.type == T3D_LIGHT_VOLUME) {
                    if ((flags & T3D_DRAW_HAS_MATERIAL) && R.lvMode == T3D_LV_OFF) U.type = LIGHT_SKIP;   // light.go:671-673
                } else if (L.type == T3D_LIGHT_DIRECTIONAL) {
                    if (approx && own) {
                        // dot(n * workingModelRotation, forward) == dot(n, g) + c0 with g = rotation rows . forward, folded per
                        // draw instead of per vertex (light.go:342-349)
                        const uint8_t lm = (flags & T3D_DRAW_HAS_MATERIAL) ? R.lightingMode : (uint8_t)T3D_LIGHTING_DEFAULT;
                        const float fx = L.forward[0], fy = L.forward[1], fz = L.forward[2];
                        if (lm == T3D_LIGHTING_FIXED_NORMALS) { U.wpos[0] = U.wpos[1] = U.wpos[2] = 0; U.range = fx * fx + fy * fy + fz * fz; }
                        else if (skinned) { U.wpos[0] = fx; U.wpos[1] = fy; U.wpos[2] = fz; U.range = 0; }
                        else {
                            const float* s = R.worldRot;        // the sun's workingModelRotation by now (role 3 above)
                            U.wpos[0] = s[0] * fx + s[1] * fy + s[2] * fz; U.wpos[1] = s[4] * fx + s[5] * fy + s[6] * fz;
                            U.wpos[2] = s[8] * fx + s[9] * fy + s[10] * fz; U.range = s[12] * fx + s[13] * fy + s[14] * fz;
                        }
                    }
                } else if (L.type == T3D_LIGHT_AMBIENT) {
                    if (approx && own) { U.fwd[0] = L.color[0] * L.energy; U.fwd[1] = L.color[1] * L.energy; U.fwd[2] = L.color[2] * L.energy; }
                }
                uint32_t at = i;
                if (sortable) {
                    for (; seen < i; seen++) {          // group sizes among the entries before this one
                        const t3d_light& Lj = lightsIn[R.lightIn + seen];
                        if (skipped(Lj, seen)) cX++;
                        else { cP += Lj.type == T3D_LIGHT_POINT; cS += Lj.type == T3D_LIGHT_DIRECTIONAL; cA += Lj.type == T3D_LIGHT_AMBIENT; }
                    }
                    at = U.type == T3D_LIGHT_POINT ? cP : U.type == T3D_LIGHT_DIRECTIONAL ? nP + cS : U.type == T3D_LIGHT_AMBIENT ? nP + nS + cA : nP + nS + nA + cX;
                }
                lightsU[R.lightU + at] = U;
            }
        }
    }
    __syncthreads();
    {
        uint4* dst = pf.devDraws + (size_t)d0 * REC16;
        const uint32_t n16 = n * REC16;
        for (uint32_t i = threadIdx.x; i < n16; i += PREP_THREADS) dst[i] = prepSmem[i];
    }
}

// ------------------------------------------------------------------------------------------------
// K1: fused skin + transform + light, one vertex per thread
// ------------------------------------------------------------------------------------------------

// VertexLight of a light that a LightVolume cell uses (light.go:116-120, 191-258, 332-360): the arithmetic of shade_vertex's
// light loop, from the LightU record k_prep_draws derived for this draw.  Rare, kept out of line.
__device__ __noinline__ float3 nested_light(const DrawRec* __restrict__ Rp, const LightU* __restrict__ Up, float px, float py, float pz,
                                            float nx, float ny, float nz, bool skinned, float3 acc) {
    const DrawRec& R = *Rp;
    const LightU U = *Up;
    const uint8_t lightingMode = (R.flags & T3D_DRAW_HAS_MATERIAL) ? R.lightingMode : (uint8_t)T3D_LIGHTING_DEFAULT;
    const Vec3 P = v3(px, py, pz), N = v3(nx, ny, nz);
    const float cR = U.color[0], cG = U.color[1], cB = U.color[2], energy = U.energy;
    if (U.type == T3D_LIGHT_POINT) {
        const Vec3 wp = v3(U.wpos[0], U.wpos[1], U.wpos[2]);
        const float range = U.range, range2 = U.range2;
        Vec3 vertNormal = N;
        const float distance = v3_dist2(wp, P);
        if (range > 0) { if (distance > range2) return acc; }
        else { if (1.0f / distance * energy < 0.001f) return acc; }
        const Vec3 lightVec = v3_unit(v3_sub(wp, P));
        if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) vertNormal = lightVec;
        float diffuse = v3_dot(vertNormal, lightVec);
        if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuse = fabsf(diffuse);
        if (diffuse > 0) {
            float diffuseFactor = diffuse * (1.0f / (1.0f + (0.1f * distance))) * 2.0f;
            if (range2 > 0) diffuseFactor *= clampf((range2 - distance) / distance, 0.0f, 1.0f);
            acc.x += cR * diffuseFactor * energy;
            acc.y += cG * diffuseFactor * energy;
            acc.z += cB * diffuseFactor * energy;
        }
    } else if (U.type == T3D_LIGHT_DIRECTIONAL) {
        const Vec3 fwd = v3(U.fwd[0], U.fwd[1], U.fwd[2]);
        Vec3 normal;
        if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) normal = fwd;
        else if (skinned) normal = N;
        else {                                                          // the draw's workingModelRotation, parked in worldRot by k_prep_draws
            const float* s = R.worldRot;
            normal.X = s[0] * N.X + s[4] * N.Y + s[8] * N.Z + s[12];
            normal.Y = s[1] * N.X + s[5] * N.Y + s[9] * N.Z + s[13];
            normal.Z = s[2] * N.X + s[6] * N.Y + s[10] * N.Z + s[14];
        }
        float diffuseFactor = v3_dot(normal, fwd);
        if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuseFactor = fabsf(diffuseFactor);
        if (diffuseFactor <= 0) return acc;
        acc.x += cR * diffuseFactor * energy;
        acc.y += cG * diffuseFactor * energy;
        acc.z += cB * diffuseFactor * energy;
    } else if (U.type == T3D_LIGHT_AMBIENT) {
        acc.x += cR * energy; acc.y += cG * energy; acc.z += cB * energy;
    }
    return acc;
}

// LightVolume.VertexLight (light.go:583-663); rare, kept out of line.  `acc` is the vertex's running light sum: the lights the
// cell uses are added first, then the cell colour.
__device__ __noinline__ float3 light_volume(const DrawRec* __restrict__ Rp, const LightU* __restrict__ Up, const LightU* __restrict__ drawLights,
                                            const VolumeU* __restrict__ vols,
                                            float px, float py, float pz, float nx, float ny, float nz, bool skinned, float3 acc) {
    const DrawRec& R = *Rp;
    const LightU U = *Up;
    const Vec3 apos = v3(px, py, pz), anrm = v3(nx, ny, nz);
    const VolumeU V = vols[U.volume];
    float cr = 0, cg = 0, cb = 0;
    const bool hasMat = R.flags & T3D_DRAW_HAS_MATERIAL;
    const int mode = hasMat ? R.lvMode : 0;
    Mat4 transform = m4_load(R.transform);
    Vec3 cellHalf = v3(V.cellSize[0] / 2, V.cellSize[1] / 2, V.cellSize[2] / 2);
    Vec3 bias = v3(V.bias[0], V.bias[1], V.bias[2]);
    Vec3 lookup;
    if (hasMat && mode == T3D_LV_PER_OBJECT) {
        lookup = v3(transform.m[3][0], transform.m[3][1], transform.m[3][2]);
    } else {
        Vec3 vertPos = v3(0, 0, 0), vertPosWithNormal = v3(0, 0, 0);
        if (skinned) {
            if (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) vertPosWithNormal = v3_add(v3_add(vertPos, v3_mult(anrm, cellHalf)), bias);
            else vertPos = v3_add(apos, bias);
        } else {
            if (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) vertPosWithNormal = v3_add(m4_multvec(transform, v3_add(apos, v3_mult(anrm, cellHalf))), bias);
            else vertPos = v3_add(m4_multvec(transform, apos), bias);
        }
        // Dimensions.Contains on vertPos (mesh.go:82-84) -- zero in WithNormal mode, as in the reference
        if (!(vertPos.X >= V.dimMin[0] && vertPos.Y >= V.dimMin[1] && vertPos.Z >= V.dimMin[2] &&
              vertPos.X <= V.dimMax[0] && vertPos.Y <= V.dimMax[1] && vertPos.Z <= V.dimMax[2])) return acc;
        lookup = (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) ? vertPosWithNormal : vertPos;
    }
    // light.go:717-724, 1267-1272
    int x = (int)roundf(((lookup.X - V.offset[0] - V.dimMin[0] - (V.cellSize[0] / 2)) / V.cellSize[0]));
    int y = (int)roundf(((lookup.Y - V.offset[1] - V.dimMin[1] - (V.cellSize[1] / 2)) / V.cellSize[1]));
    int z = (int)roundf(((lookup.Z - V.offset[2] - V.dimMin[2] - (V.cellSize[2] / 2)) / V.cellSize[2]));
    if (!(x < 0 || y < 0 || z < 0 || x >= V.cellCount[0] || y >= V.cellCount[1] || z >= V.cellCount[2])) {
        const size_t cell = (size_t)x + ((size_t)y * V.cellCount[0]) + ((size_t)z * V.cellCount[0] * V.cellCount[1]);
        const float* c = V.cells + 3 * cell;
        cr = c[0]; cg = c[1]; cb = c[2];
        if (V.cellLightStart && !(hasMat && mode == T3D_LV_PER_OBJECT)) {      // light.go:652-654 (the per-object branch has no nested lights)
            const uint32_t k1 = V.cellLightStart[cell + 1];
            for (uint32_t k = V.cellLightStart[cell]; k < k1; k++)
                acc = nested_light(Rp, drawLights + U.useFirst + V.cellLightIndex[k], px, py, pz, nx, ny, nz, skinned, acc);
        }
    }
    // a vertex outside the volume returned above and adds nothing; inside, x + 0 keeps x (the sums are never -0)
    acc.x += U.color[0] * cr * U.energy; acc.y += U.color[1] * cg * U.energy; acc.z += U.color[2] * cb * U.energy;
    return acc;
}

// Model.skinVertex (model.go:559-596) for one vertex.
struct SkinAcc { float4 k0, k1, k2, k3; bool done; };
// One influence: skip weight 0, take the bone alone on weight 1 (and stop), else Scale then Add (matrix.go:639-661).  `U`
// hands out the draw's bone matrices row by row: from the frame's bone arena, or from the copy the tile's producer staged in
// shared memory.
template <class Uniforms>
__device__ __forceinline__ void skin_influence(SkinAcc& A, const Uniforms& U, float weightPerc, uint32_t joint, bool present) {
    if (A.done || !present || weightPerc == 0.0f) return;
    const float4 r0 = U.bone(joint, 0), r1 = U.bone(joint, 1), r2 = U.bone(joint, 2), r3 = U.bone(joint, 3);
    if (weightPerc == 1.0f) { A.k0 = r0; A.k1 = r1; A.k2 = r2; A.k3 = r3; A.done = true; return; }
    A.k0.x = A.k0.x + r0.x * weightPerc; A.k0.y = A.k0.y + r0.y * weightPerc; A.k0.z = A.k0.z + r0.z * weightPerc; A.k0.w = A.k0.w + r0.w * weightPerc;
    A.k1.x = A.k1.x + r1.x * weightPerc; A.k1.y = A.k1.y + r1.y * weightPerc; A.k1.z = A.k1.z + r1.z * weightPerc; A.k1.w = A.k1.w + r1.w * weightPerc;
    A.k2.x = A.k2.x + r2.x * weightPerc; A.k2.y = A.k2.y + r2.y * weightPerc; A.k2.z = A.k2.z + r2.z * weightPerc; A.k2.w = A.k2.w + r2.w * weightPerc;
    A.k3.x = A.k3.x + r3.x * weightPerc; A.k3.y = A.k3.y + r3.y * weightPerc; A.k3.z = A.k3.z + r3.z * weightPerc; A.k3.w = A.k3.w + r3.w * weightPerc;
}
template <class Uniforms>
__device__ __forceinline__ void skin_one(const Uniforms& U, float4 w, ushort4 j, int n,
                                         Vec3 p, Vec3 nIn, Vec3& pOut, Vec3& nOut) {
    SkinAcc A;
    A.k0 = make_float4(0, 0, 0, 0); A.k1 = A.k0; A.k2 = A.k0; A.k3 = A.k0; A.done = false;
    skin_influence(A, U, w.x, j.x, n > 0);
    skin_influence(A, U, w.y, j.y, n > 1);
    skin_influence(A, U, w.z, j.z, n > 2);
    skin_influence(A, U, w.w, j.w, n > 3);
    const float4 k0 = A.k0, k1 = A.k1, k2 = A.k2, k3 = A.k3;
    pOut.X = k0.x * p.X + k1.x * p.Y + k2.x * p.Z + k3.x;                  // MultVec, matrix.go:619
    pOut.Y = k0.y * p.X + k1.y * p.Y + k2.y * p.Z + k3.y;
    pOut.Z = k0.z * p.X + k1.z * p.Y + k2.z * p.Z + k3.z;
    // skinMatrix[3] = {0,0,0,1} (model.go:587-590)
    nOut.X = k0.x * nIn.X + k1.x * nIn.Y + k2.x * nIn.Z + 0.0f;
    nOut.Y = k0.y * nIn.X + k1.y * nIn.Y + k2.y * nIn.Z + 0.0f;
    nOut.Z = k0.z * nIn.X + k1.z * nIn.Y + k2.z * nIn.Z + 0.0f;
}

// Tunables (see profiles/README.md for the measurements behind the defaults).
#ifndef T3D_K1_MINBLOCKS
#define T3D_K1_MINBLOCKS 3      // resident blocks per SM the generic kernel's register allocation aims for
#endif
static const int VERTEX_BLOCK = 256;                        // vertex slots per tile
// consumer threads of a k_vertex block: the static fast class shades two vertices per thread (slots t and t + 128 of the tile),
// which halves everything a warp pays per tile and per light -- barrier polls, stage bookkeeping, the loads of the uniforms
// and of every light record -- for a kernel that is bound by instruction issue; the other classes keep one vertex per thread
#ifndef T3D_K1_STATIC_CONSUMERS
#define T3D_K1_STATIC_CONSUMERS 128
#endif
template <int MODE> struct VertexShape { static const int CONSUMERS = VERTEX_BLOCK; static const int THREADS = VERTEX_BLOCK + 32; };
template <> struct VertexShape<1 /* K1_STATIC */> { static const int CONSUMERS = T3D_K1_STATIC_CONSUMERS; static const int THREADS = T3D_K1_STATIC_CONSUMERS + 32; };
// tiles in flight per block: the static class has the small stage (7 KB) and the short tiles, so its producer runs furthest ahead
#ifndef T3D_VERTEX_STAGES_STATIC
#define T3D_VERTEX_STAGES_STATIC 3
#endif
#ifndef T3D_VERTEX_STAGES
#define T3D_VERTEX_STAGES 2
#endif
#ifndef T3D_MBAR_SLEEP
#define T3D_MBAR_SLEEP 64       // ns a waiting warp sleeps between polls of a stage barrier (0: poll back to back)
#endif
static const int VERTEX_SMEM_LIGHTS = 16;       // light lists up to this long are staged in shared memory
static const int VERTEX_SMEM_BONES = 64;        // skinned tiles: armatures of up to this many bones are staged in shared memory
static_assert(VERTEX_BLOCK == (int)VERTEX_TILE, "t3d_internal.h: VERTEX_TILE must be the slots one k_vertex block covers");
static_assert(VERTEX_BLOCK % T3D_K1_STATIC_CONSUMERS == 0 && T3D_K1_STATIC_CONSUMERS % 32 == 0, "whole warps, whole passes over a tile");

// Everything one vertex needs from its draw: either the DrawRec / LightU records and the mesh / bone arrays in global
// memory, or the stage the block's producer warp filled in shared memory (mvp rows 0-3, the sun's workingModelRotation rows
// 4-7, the LightU list from 8; weights / joints per tile slot; the draw's bone matrices).
struct UniformsGlobal {
    const float* mvpP; const float* sunP; const float4* lightP;
    const float* wgtP; const uint16_t* jntP; const float* bonesP; uint32_t mv;
    __device__ __forceinline__ float4 mvp(int i) const { return ldg4(mvpP + 4 * i); }
    __device__ __forceinline__ float4 sun(int i) const { return ldg4(sunP + 4 * i); }
    __device__ __forceinline__ float4 light(uint32_t i) const { return __ldg(lightP + i); }
    __device__ __forceinline__ float lightWord(uint32_t i) const { return __ldg(reinterpret_cast<const float*>(lightP + i)); }     // first word of float4 i
    __device__ __forceinline__ float4 weights() const { return ldg4(wgtP + 4 * (size_t)mv); }
    __device__ __forceinline__ ushort4 joints() const { return __ldg(reinterpret_cast<const ushort4*>(jntP + 4 * (size_t)mv)); }
    __device__ __forceinline__ float4 bone(uint32_t joint, int row) const { return __ldg(reinterpret_cast<const float4*>(bonesP + 16 * (size_t)joint) + row); }
};
struct UniformsShared {
    const float4* s; const float4* sWgt; const ushort4* sJnt; const float4* sBones; uint32_t ti;
    __device__ __forceinline__ float4 mvp(int i) const { return s[i]; }
    __device__ __forceinline__ float4 sun(int i) const { return s[4 + i]; }
    __device__ __forceinline__ float4 light(uint32_t i) const { return s[8 + i]; }
    __device__ __forceinline__ float lightWord(uint32_t i) const { return *reinterpret_cast<const float*>(s + 8 + i); }
    __device__ __forceinline__ float4 weights() const { return sWgt[ti]; }
    __device__ __forceinline__ ushort4 joints() const { return sJnt[ti]; }
    __device__ __forceinline__ float4 bone(uint32_t joint, int row) const { return sBones[4 * joint + row]; }
};

// One vertex: shape keys, skinning, clip transform, optional outputs, lighting.  P / N arrive loaded.
// MODE K1_STATIC: the draw is static, has no shape keys, no unbillboarded depth and no volume light, and the frame has
// no vertex snapping and no normal pass -- the rare paths are compiled out, which frees a quarter of the registers.
// MODE K1_SKINNED: the same exclusions for a skinned draw.  MODE K1_GENERIC: everything.
enum { K1_GENERIC = 0, K1_STATIC = 1, K1_SKINNED = 2 };

// Single-instruction reciprocal / reciprocal square root (MUFU.RCP / MUFU.RSQ, <= 2 ulp) for the lighting arithmetic of the
// default build: lit colours only have to match within 1/255 (BASELINE.json north_star), whereas IEEE-exact
// division and square root expand to ~12 instructions each plus a slow-path branch -- five divisions and a root per point
// light (light.go:191-258) were two thirds of k_vertex's instructions.  Nothing that feeds a cull decision, a depth key, a bin
// or a position uses these.
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// PointLight.VertexLight (light.go:191-258) with approximate reciprocals and fused multiply-adds, branch-free; returns the factor
// the light's premultiplied colour (colour * energy * 2, from k_prep_draws) is scaled by (0: no contribution).  The two cut-offs
// are DECISIONS, not continuous quantities (a vertex on the wrong side of `1 / d * energy < 0.001` changes by up to 0.02), so they
// are taken exactly as the reference takes them, on the exact squared distance: `cut` is lightRangeSquared for a ranged light
// and otherwise the largest d that passes the reference's energy test (t3d_submit_lights finds it by bisection over the floats,
// with the reference's two roundings), so one comparison decides either.  `r2eff` is lightRangeSquared or +inf: the range
// falloff (light.go:247-249) of an unranged light saturates to 1.
template <bool PLAIN>     // PLAIN: LightingModeDefault (the normal is the vertex normal, one-sided)
__device__ __forceinline__ float point_light_factor_fast(float wx, float wy, float wz, float cut, float r2eff, Vec3 P, Vec3 N, uint8_t lightingMode) {
    const float dx = wx - P.X, dy = wy - P.Y, dz = wz - P.Z;
    const float d2 = dx * dx + dy * dy + dz * dz;               // -fmad=false: the reference's distance, bit for bit
    const float rl = rsqrt_fast(d2);
    float ndl = __fmaf_rn(N.Z, dz, __fmaf_rn(N.Y, dy, N.X * dx));
    if (!PLAIN && lightingMode == T3D_LIGHTING_FIXED_NORMALS) ndl = d2 * rl;
    float diffuse = ndl * rl;
    if (!PLAIN && lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuse = fabsf(diffuse);
    float f = fmaxf(diffuse, 0.0f) * rcp_fast(__fmaf_rn(0.1f, d2, 1.0f));      // diffuse <= 0 contributes nothing
    f *= __saturatef((r2eff - d2) * (rl * rl));
    // Unit (vector.go:258-266) leaves vectors shorter than 1e-8 as they are: their contribution is below 1e-7
    return (d2 <= cut && d2 >= 1e-16f) ? f : 0.0f;
}

// The light loops of the fast classes with the default lighting arithmetic.  k_prep_draws grouped the draw's LightU list by type
// ([point | directional | ambient], `sorted`) and left pre-folded fields in the records (t3d_device.cuh: LightU), so each group
// is a tight loop without a type dispatch; a list it left in submit order (more than PREP_SORT_LIGHTS entries) is walked entry
// by entry.  Summation order differs from the reference's either way this arithmetic is used: within the 1/255 tolerance.
template <bool PLAIN, class Uniforms>
__device__ __forceinline__ float3 light_loop_fast(const Uniforms& U, uint32_t nL, uint32_t nPoint, uint32_t nSun, uint32_t nAmbient, bool sorted,
                                                  Vec3 P, Vec3 N, uint8_t lightingMode) {
    float lr = 0, lg = 0, lb = 0;
    auto point = [&](uint32_t li) {                                 // light.go:191-258
        const float4 q1 = U.light(4 * li + 1), q2 = U.light(4 * li + 2);
        const float pb = U.lightWord(4 * li + 3);
        const float f = point_light_factor_fast<PLAIN>(q1.x, q1.y, q1.z, q2.x, q2.y, P, N, lightingMode);
        lr = __fmaf_rn(q2.z, f, lr); lg = __fmaf_rn(q2.w, f, lg); lb = __fmaf_rn(pb, f, lb);
    };
    auto sun = [&](uint32_t li) {                                   // light.go:332-360
        const float4 q0 = U.light(4 * li), q1 = U.light(4 * li + 1);
        const float c0 = U.lightWord(4 * li + 2);
        float f = __fmaf_rn(N.Z, q1.z, __fmaf_rn(N.Y, q1.y, __fmaf_rn(N.X, q1.x, c0)));
        if (!PLAIN && lightingMode == T3D_LIGHTING_DOUBLE_SIDED) f = fabsf(f);
        f = fmaxf(f, 0.0f) * q1.w;                                  // diffuseFactor <= 0 contributes nothing
        lr = __fmaf_rn(q0.y, f, lr); lg = __fmaf_rn(q0.z, f, lg); lb = __fmaf_rn(q0.w, f, lb);
    };
    auto ambient = [&](uint32_t li) {                               // light.go:110-120
        const float4 q2 = U.light(4 * li + 2);
        lr += q2.z; lg += q2.w; lb += U.lightWord(4 * li + 3);
    };
    if (sorted) {
        uint32_t li = 0;
#pragma unroll 2
        for (; li < nPoint; li++) point(li);
#pragma unroll 1
        for (const uint32_t e = li + nSun; li < e; li++) sun(li);
#pragma unroll 1
        for (const uint32_t e = li + nAmbient; li < e; li++) ambient(li);
    } else {
#pragma unroll 1
        for (uint32_t li = 0; li < nL; li++) {
            const uint32_t type = __float_as_uint(U.lightWord(4 * li));
            if (type == T3D_LIGHT_POINT) point(li);
            else if (type == T3D_LIGHT_DIRECTIONAL) sun(li);
            else if (type == T3D_LIGHT_AMBIENT) ambient(li);
        }
    }
    return make_float3(lr, lg, lb);
}

// APPROX: lighting with approximate reciprocals (colours within 1/255, the default); !APPROX: the reference's arithmetic
// operation by operation (bit-exact colours; t3d_set_exact_lighting).  Positions, depths and everything the cull and the
// sort read are exact in both.
struct LightGroups { uint32_t nPoint, nSun, nAmbient; bool sorted; };     // DrawRec.nPoint ... lightsSorted
template <int MODE, bool APPROX, class Uniforms>
__device__ __forceinline__ void shade_vertex(const DrawRec& R, const uint32_t flags, const uint32_t nL, const LightGroups G, const uint8_t lightingMode, const Uniforms U, const FrameConst& fc,
                                             const uint32_t slot, const uint32_t mv, Vec3 P, Vec3 N,
                                             const LightU* __restrict__ lightsU, const VolumeU* __restrict__ vols,
                                             float4* __restrict__ clipOut, float2* __restrict__ clipXY, float4* __restrict__ lightOut, float4* __restrict__ wposOut,
                                             float* __restrict__ unbillOut, float4* __restrict__ tnrmOut) {
    constexpr bool FAST = MODE != K1_GENERIC;       // shape keys, snapping, unbillboarded depth, normal pass, volume lights compiled out
    if (flags & T3D_DRAW_FRUSTUM_CULLED) return;    // an instance outside the frustum: nothing of it is read later
    const bool skinned = MODE == K1_SKINNED || (MODE == K1_GENERIC && (flags & T3D_DRAW_SKINNED));
    const bool lit = (flags & T3D_DRAW_LIT) != 0;
    const bool needN = lit || skinned;
    const uint32_t vs = R.vstride;

    // shape keys (mesh.go:796-822)
    for (uint32_t k = 0; !FAST && k < R.nKeys; k++) {
        const float w = __ldg(&R.keyW[k]);
        const float* kp = R.keyPos + (size_t)k * 3 * vs + mv;
        P.X += __ldg(kp) * w; P.Y += __ldg(kp + vs) * w; P.Z += __ldg(kp + 2 * (size_t)vs) * w;
        if (needN) {
            const float* kn = R.keyNrm + (size_t)k * 3 * vs + mv;
            N.X += __ldg(kn) * w; N.Y += __ldg(kn + vs) * w; N.Z += __ldg(kn + 2 * (size_t)vs) * w;
        }
    }

    // skinning (model.go:760-775): positions and normals become world-space
    if (MODE != K1_STATIC && skinned) {
        const float4 w = U.weights();
        const ushort4 j = U.joints();
        const int n = R.ninf ? (int)__ldg(R.ninf + mv) : 4;
        Vec3 po, no;
        skin_one(U, w, j, n, P, N, po, no);
        P = po; N = no;
        wposOut[slot] = make_float4(po.X, po.Y, po.Z, 0.0f);
    }

    // clip = [x y z 1] * M (model.go:772-775 / 795-798)
    {
        float4 m0, m1, m2, m3;
        if (skinned) {
            m0 = make_float4(fc.vp[0], fc.vp[1], fc.vp[2], fc.vp[3]);
            m1 = make_float4(fc.vp[4], fc.vp[5], fc.vp[6], fc.vp[7]);
            m2 = make_float4(fc.vp[8], fc.vp[9], fc.vp[10], fc.vp[11]);
            m3 = make_float4(fc.vp[12], fc.vp[13], fc.vp[14], fc.vp[15]);
        } else {
            m0 = U.mvp(0); m1 = U.mvp(1); m2 = U.mvp(2); m3 = U.mvp(3);
        }
        const float snap = fc.snapping;
        float4 c;
        c.x = m0.x * P.X + m1.x * P.Y + m2.x * P.Z + m3.x;
        c.y = m0.y * P.X + m1.y * P.Y + m2.y * P.Z + m3.y;
        c.z = m0.z * P.X + m1.z * P.Y + m2.z * P.Z + m3.z;
        c.w = m0.w * P.X + m1.w * P.Y + m2.w * P.Z + m3.w;
        if (!FAST && snap > 0) {                                        // model.go:807-812
            c.x = roundf(c.x * snap) / snap;
            c.y = roundf(c.y * snap) / snap;
            c.z = roundf(c.z * snap) / snap;
        }
        // The perspective divide of the cull (model.go:866-875) depends on the vertex alone, so it is done here once
        // per vertex instead of once per triangle corner: {X/w, Y/w, Z, W} with the cull's clamp of a negative w.
        // The emit loop divides by the same w unless w < 0 (camera.go:1655-1675), and then reads X, Y from clipXY.
        float wc = c.w;
        if (wc < 0) wc = 0.000001f;
        clipOut[slot] = make_float4(c.x / wc, c.y / wc, c.z, c.w);
        clipXY[slot] = make_float2(c.x, c.y);
        if (!FAST && !skinned && (flags & T3D_DRAW_DEPTH_UNBILLBOARDED)) {   // model.go:800-803
            const float4 u = ldg4(R.unbill);
            unbillOut[slot] = u.x * P.X + u.y * P.Y + u.z * P.Z + u.w;
        }
        if (!FAST && (fc.camFlags & T3D_CAM_RENDER_NORMALS)) {          // model.go:814-818 (uses vert)
            const float4 c0 = ldg4(R.nrmRot), c1 = ldg4(R.nrmRot + 4), c2 = ldg4(R.nrmRot + 8);
            float4 t;
            t.x = c0.x * P.X + c0.y * P.Y + c0.z * P.Z + c0.w;
            t.y = c1.x * P.X + c1.y * P.Y + c1.z * P.Z + c1.w;
            t.z = c2.x * P.X + c2.y * P.Y + c2.z * P.Z + c2.w;
            t.w = 0.0f;
            tnrmOut[slot] = t;
        }
    }

    if (!lit) return;

    // per-vertex lighting (camera.go:1610-1644): start at (0,0,0,1), add every light of the draw's list
    float lr = 0, lg = 0, lb = 0;
    if constexpr (APPROX && FAST) {
        // (the sun's folded forward assumes the mode k_prep_draws saw, which is this one)
        const float3 acc = lightingMode == T3D_LIGHTING_DEFAULT ? light_loop_fast<true>(U, nL, G.nPoint, G.nSun, G.nAmbient, G.sorted, P, N, lightingMode)
                                                                : light_loop_fast<false>(U, nL, G.nPoint, G.nSun, G.nAmbient, G.sorted, P, N, lightingMode);
        lightOut[slot] = make_float4(acc.x, acc.y, acc.z, 1.0f);
        return;
    }
    // Model.BakeLighting (model.go:1246-1262) walks light by light over the part's triangles and adds the light's contribution
    // to a vertex once per corner that refers to it: `rep` additions of the same value, in sequence, per light
    uint32_t rep = 1;
    if constexpr (!FAST) { if (fc.bakeValence) rep = fc.bakeValence[slot]; }
    for (uint32_t li = 0; li < nL; li++) {
        const float4 q0 = U.light(4 * li), q1 = U.light(4 * li + 1), q2 = U.light(4 * li + 2);
        const uint32_t type = __float_as_uint(q0.x);
        const float cR = q0.y, cG = q0.z, cB = q0.w, energy = q1.w;
        if (APPROX && type == T3D_LIGHT_POINT) {
            const float4 q3 = U.light(4 * li + 3);
            const float f = point_light_factor_fast<false>(q1.x, q1.y, q1.z, q2.x, q2.y, P, N, lightingMode);
            lr = __fmaf_rn(q2.z, f, lr); lg = __fmaf_rn(q2.w, f, lg); lb = __fmaf_rn(q3.x, f, lb);
        } else if (type == T3D_LIGHT_POINT) {                           // light.go:191-258
            const Vec3 wp = v3(q1.x, q1.y, q1.z);
            const float range = q2.x, range2 = q2.y;
            Vec3 vertNormal = N;
            const float distance = v3_dist2(wp, P);
            if (range > 0) { if (distance > range2) continue; }
            else { if (1.0f / distance * energy < 0.001f) continue; }
            const Vec3 lightVec = v3_unit(v3_sub(wp, P));
            if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) vertNormal = lightVec;
            float diffuse = v3_dot(vertNormal, lightVec);
            if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuse = fabsf(diffuse);
            if (diffuse > 0) {
                float diffuseFactor = diffuse * (1.0f / (1.0f + (0.1f * distance))) * 2.0f;
                if (range2 > 0) diffuseFactor *= clampf((range2 - distance) / distance, 0.0f, 1.0f);
                const float ar = cR * diffuseFactor * energy, ag = cG * diffuseFactor * energy, ab = cB * diffuseFactor * energy;
                for (uint32_t r = 0; r < rep; r++) { lr += ar; lg += ag; lb += ab; }
            }
        } else if (type == T3D_LIGHT_DIRECTIONAL) {                     // light.go:332-360
            const float4 q3 = U.light(4 * li + 3);
            const Vec3 fwd = v3(q2.z, q2.w, q3.x);
            Vec3 normal;
            if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) normal = fwd;
            else if (skinned) normal = N;
            else {
                const float4 s0 = U.sun(0), s1 = U.sun(1), s2 = U.sun(2), s3 = U.sun(3);
                if (APPROX) {
                    normal.X = __fmaf_rn(s2.x, N.Z, __fmaf_rn(s1.x, N.Y, __fmaf_rn(s0.x, N.X, s3.x)));
                    normal.Y = __fmaf_rn(s2.y, N.Z, __fmaf_rn(s1.y, N.Y, __fmaf_rn(s0.y, N.X, s3.y)));
                    normal.Z = __fmaf_rn(s2.z, N.Z, __fmaf_rn(s1.z, N.Y, __fmaf_rn(s0.z, N.X, s3.z)));
                } else {
                    normal.X = s0.x * N.X + s1.x * N.Y + s2.x * N.Z + s3.x;
                    normal.Y = s0.y * N.X + s1.y * N.Y + s2.y * N.Z + s3.y;
                    normal.Z = s0.z * N.X + s1.z * N.Y + s2.z * N.Z + s3.z;
                }
            }
            if (APPROX) {
                float f = __fmaf_rn(normal.Z, fwd.Z, __fmaf_rn(normal.Y, fwd.Y, normal.X * fwd.X));
                if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) f = fabsf(f);
                f = fmaxf(f, 0.0f) * energy;        // diffuseFactor <= 0 contributes nothing
                lr = __fmaf_rn(cR, f, lr); lg = __fmaf_rn(cG, f, lg); lb = __fmaf_rn(cB, f, lb);
                continue;
            }
            float diffuseFactor = v3_dot(normal, fwd);
            if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuseFactor = fabsf(diffuseFactor);
            if (diffuseFactor <= 0) continue;
            const float ar = cR * diffuseFactor * energy, ag = cG * diffuseFactor * energy, ab = cB * diffuseFactor * energy;
            for (uint32_t r = 0; r < rep; r++) { lr += ar; lg += ag; lb += ab; }
        } else if (type == T3D_LIGHT_AMBIENT) {                         // light.go:110-120
            const float ar = cR * energy, ag = cG * energy, ab = cB * energy;
            for (uint32_t r = 0; r < rep; r++) { lr += ar; lg += ag; lb += ab; }
        } else if (!FAST && type == T3D_LIGHT_VOLUME) {
            for (uint32_t r = 0; r < rep; r++) {
                const float3 acc = light_volume(&R, lightsU + R.lightU + li, lightsU + R.lightU, vols, P.X, P.Y, P.Z, N.X, N.Y, N.Z, skinned, make_float3(lr, lg, lb));
                lr = acc.x; lg = acc.y; lb = acc.z;
            }
        }
    }
    lightOut[slot] = make_float4(lr, lg, lb, 1.0f);
}

// ---- mbarrier / bulk-copy plumbing of the k_vertex pipeline -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (done) break;
        if (T3D_MBAR_SLEEP) __nanosleep(T3D_MBAR_SLEEP);        // a waiting warp must not eat the issue slots of the warps that shade
    }
}
// global -> shared bulk copy by the copy engine of the SM (TMA, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One stage of the pipeline: what the producer warp fetches for one tile.
struct __align__(16) VStageDesc {
    uint32_t base, end;         // slots [base, end) of the vertex arena
    uint32_t dLo, dHi;          // draw of the first slot; draw of the first slot of the next tile
    uint32_t staged;            // 1: the tile lies inside draw dLo and its inputs are in this stage; 0: consumers load for themselves
    uint32_t mv0;               // mesh vertex of slot `base` (staged tiles)
    uint32_t flags;             // staged tiles: the draw's flags, light count and lighting mode (what every vertex would otherwise
    uint32_t nLmode;            //   fetch from the draw record): lightCount | lightingMode << 16; NO_TILE in `base`: the list is exhausted
    uint32_t groups;            // staged tiles: nPoint | nSun << 16
    uint32_t groups2;           //   nAmbient | lightsSorted << 16
    uint32_t pad[2];
};
static const uint32_t NO_TILE = 0xFFFFFFFFu;
template <bool SKIN> struct __align__(128) VStage {
    float4 U[8 + 4 * VERTEX_SMEM_LIGHTS];       // mvp | sun rotation | LightU list
    float pos[3][VERTEX_BLOCK];
    float nrm[3][VERTEX_BLOCK];
    VStageDesc desc;
};
template <> struct __align__(128) VStage<true> {
    float4 U[8 + 4 * VERTEX_SMEM_LIGHTS];
    float pos[3][VERTEX_BLOCK];
    float nrm[3][VERTEX_BLOCK];
    float4 wgt[VERTEX_BLOCK];
    ushort4 jnt[VERTEX_BLOCK];
    float4 bones[4 * VERTEX_SMEM_BONES];
    VStageDesc desc;
};

// Two vertices of a staged tile of the static fast class under default lighting with a grouped light list (the hot case:
// instanced scenery, terrain): slots `ta` and `tb` of the tile, shaded together so that the draw's mvp and every light record
// are loaded from shared memory once per pair.  The arithmetic per vertex is shade_vertex<K1_STATIC, true>'s, operation for
// operation (positions exact, lighting through light_loop_fast's formulas).  `hasB`: slot tb exists (a draw's last tile is short).
__device__ __forceinline__ void shade2_static(const VStage<false>& S, const VStageDesc& D, const uint32_t ta, const uint32_t tb, const bool hasB,
                                              float4* __restrict__ clipOut, float2* __restrict__ clipXY, float4* __restrict__ lightOut) {
    if (D.flags & T3D_DRAW_FRUSTUM_CULLED) return;
    const bool lit = (D.flags & T3D_DRAW_LIT) != 0;
    const Vec3 Pa = v3(S.pos[0][ta], S.pos[1][ta], S.pos[2][ta]), Pb = v3(S.pos[0][tb], S.pos[1][tb], S.pos[2][tb]);
    {   // clip = [x y z 1] * mvp (model.go:795-798) and the cull's perspective divide (model.go:866-875), as in shade_vertex
        const float4 m0 = S.U[0], m1 = S.U[1], m2 = S.U[2], m3 = S.U[3];
        float4 ca, cb;
        ca.x = m0.x * Pa.X + m1.x * Pa.Y + m2.x * Pa.Z + m3.x; cb.x = m0.x * Pb.X + m1.x * Pb.Y + m2.x * Pb.Z + m3.x;
        ca.y = m0.y * Pa.X + m1.y * Pa.Y + m2.y * Pa.Z + m3.y; cb.y = m0.y * Pb.X + m1.y * Pb.Y + m2.y * Pb.Z + m3.y;
        ca.z = m0.z * Pa.X + m1.z * Pa.Y + m2.z * Pa.Z + m3.z; cb.z = m0.z * Pb.X + m1.z * Pb.Y + m2.z * Pb.Z + m3.z;
        ca.w = m0.w * Pa.X + m1.w * Pa.Y + m2.w * Pa.Z + m3.w; cb.w = m0.w * Pb.X + m1.w * Pb.Y + m2.w * Pb.Z + m3.w;
        float wa = ca.w, wb = cb.w;
        if (wa < 0) wa = 0.000001f;
        if (wb < 0) wb = 0.000001f;
        clipOut[D.base + ta] = make_float4(ca.x / wa, ca.y / wa, ca.z, ca.w);
        clipXY[D.base + ta] = make_float2(ca.x, ca.y);
        if (hasB) {
            clipOut[D.base + tb] = make_float4(cb.x / wb, cb.y / wb, cb.z, cb.w);
            clipXY[D.base + tb] = make_float2(cb.x, cb.y);
        }
    }
    if (!lit) return;
    const Vec3 Na = v3(S.nrm[0][ta], S.nrm[1][ta], S.nrm[2][ta]), Nb = v3(S.nrm[0][tb], S.nrm[1][tb], S.nrm[2][tb]);
    const float4* L = S.U + 8;
    const uint32_t nPoint = D.groups & 0xFFFFu, nSun = D.groups >> 16, nAmbient = D.groups2 & 0xFFFFu;
    float ar = 0, ag = 0, ab = 0, br = 0, bg = 0, bb = 0;
    uint32_t li = 0;
#pragma unroll 2
    for (; li < nPoint; li++) {                                     // light.go:191-258
        const float4 q1 = L[4 * li + 1], q2 = L[4 * li + 2];
        const float pb = L[4 * li + 3].x;
        const float fa = point_light_factor_fast<true>(q1.x, q1.y, q1.z, q2.x, q2.y, Pa, Na, T3D_LIGHTING_DEFAULT);
        const float fb = point_light_factor_fast<true>(q1.x, q1.y, q1.z, q2.x, q2.y, Pb, Nb, T3D_LIGHTING_DEFAULT);
        ar = __fmaf_rn(q2.z, fa, ar); ag = __fmaf_rn(q2.w, fa, ag); ab = __fmaf_rn(pb, fa, ab);
        br = __fmaf_rn(q2.z, fb, br); bg = __fmaf_rn(q2.w, fb, bg); bb = __fmaf_rn(pb, fb, bb);
    }
#pragma unroll 1
    for (const uint32_t e = li + nSun; li < e; li++) {              // light.go:332-360
        const float4 q0 = L[4 * li], q1 = L[4 * li + 1];
        const float c0 = L[4 * li + 2].x;
        const float fa = fmaxf(__fmaf_rn(Na.Z, q1.z, __fmaf_rn(Na.Y, q1.y, __fmaf_rn(Na.X, q1.x, c0))), 0.0f) * q1.w;
        const float fb = fmaxf(__fmaf_rn(Nb.Z, q1.z, __fmaf_rn(Nb.Y, q1.y, __fmaf_rn(Nb.X, q1.x, c0))), 0.0f) * q1.w;
        ar = __fmaf_rn(q0.y, fa, ar); ag = __fmaf_rn(q0.z, fa, ag); ab = __fmaf_rn(q0.w, fa, ab);
        br = __fmaf_rn(q0.y, fb, br); bg = __fmaf_rn(q0.z, fb, bg); bb = __fmaf_rn(q0.w, fb, bb);
    }
#pragma unroll 1
    for (const uint32_t e = li + nAmbient; li < e; li++) {          // light.go:110-120
        const float4 q2 = L[4 * li + 2];
        const float cb = L[4 * li + 3].x;
        ar += q2.z; ag += q2.w; ab += cb; br += q2.z; bg += q2.w; bb += cb;
    }
    lightOut[D.base + ta] = make_float4(ar, ag, ab, 1.0f);
    if (hasB) lightOut[D.base + tb] = make_float4(br, bg, bb, 1.0f);
}

// One-off timeline of k_vertex (build with -DT3D_VERTEX_TRACE; tools/vertex_trace.py prints it): %globaltimer stamps of the
// first 1024 blocks -- [0] block start, [1] past the dependency wait, producer: [2] first entry + record read, [3] first tile's
// copies issued, [4] last tile issued; consumer warp 0: [5] first tile arrived, [6] first tile shaded, [7] last tile shaded,
// [8] tiles this block took.
#ifdef T3D_VERTEX_TRACE
__device__ unsigned long long g_vtrace[1024][16];
__device__ __forceinline__ void vtrace(int i) { if (blockIdx.x < 1024) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g_vtrace[blockIdx.x][i] = t; } }
__device__ __forceinline__ void vtrace_set(int i, unsigned long long v) { if (blockIdx.x < 1024) g_vtrace[blockIdx.x][i] = v; }
#else
__device__ __forceinline__ void vtrace(int) {}
__device__ __forceinline__ void vtrace_set(int, unsigned long long) {}
#endif

// k_vertex: persistent, warp-specialised, TMA-fed.  The host sorts the frame's vertex tiles (256 consecutive arena slots)
// into three lists: tiles whose draws are all of one fast class (see shade_vertex) go to k_vertex<K1_STATIC> or
// k_vertex<K1_SKINNED>, everything else to k_vertex<K1_GENERIC>.  A launch covers its list with ONE wave of blocks (at most
// SMs x resident blocks) that take tiles off the list through an atomic counter until it is exhausted: no tail wave, and an SM
// that falls behind (or starts late) simply takes fewer tiles.
//   producer warp (one elected lane): per tile, reads the entry and the draw's record, and -- when the tile lies inside one
//     draw with a stageable light list -- has the copy engine fetch the tile's position / normal plane runs (3 + 3 x 1 KB),
//     skin weights / joints and the draw's bone matrices, mvp, sun rotation and LightU list into the stage (cp.async.bulk,
//     completion counted on the stage's mbarrier); it runs up to VERTEX_STAGES tiles ahead of the consumers, so the
//     entry -> record -> data dependency chain and the HBM latency are off the shading warps' critical path.
//   consumer warps (8 x 32 threads, one vertex each): wait for the stage, shade from shared memory, release the stage.
// Tiles that straddle draws (small meshes packed back to back) are flagged unstaged: the consumers look their draw up per
// vertex and load for themselves, as before.
#ifndef T3D_K1_FAST_MINBLOCKS
#define T3D_K1_FAST_MINBLOCKS 7
#endif
#ifndef T3D_K1_SKINNED_MINBLOCKS
#define T3D_K1_SKINNED_MINBLOCKS 4
#endif
// `second`: this launch directly follows another k_vertex launch of the frame.  They are independent (disjoint
// tiles) and all only need k_prep_draws, which has completed by the time a later launch's blocks can start (the
// first launch's blocks call pdl_launch() after their own pdl_wait()).  So a later launch starts shading at
// once and waits for its predecessor at its END, which keeps "this kernel complete => every earlier kernel complete".
template <int MODE, bool APPROX>
__global__ void __launch_bounds__(VertexShape<MODE>::THREADS, MODE == K1_STATIC ? T3D_K1_FAST_MINBLOCKS : MODE == K1_SKINNED ? T3D_K1_SKINNED_MINBLOCKS : T3D_K1_MINBLOCKS)
k_vertex(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ vbase,
         const uint4* __restrict__ tileList, uint32_t nTiles, uint32_t* __restrict__ tileCtr, uint32_t totalV, uint32_t second, const __grid_constant__ FrameConst fc,
         const LightU* __restrict__ lightsU, const float* __restrict__ bones, const VolumeU* __restrict__ vols,
         float4* __restrict__ clipOut, float2* __restrict__ clipXY, float4* __restrict__ lightOut, float4* __restrict__ wposOut,
         float* __restrict__ unbillOut, float4* __restrict__ tnrmOut) {
    constexpr bool SKIN = MODE != K1_STATIC;
    constexpr int VERTEX_STAGES = MODE == K1_STATIC ? T3D_VERTEX_STAGES_STATIC : T3D_VERTEX_STAGES;
    constexpr uint32_t CONSUMERS = VertexShape<MODE>::CONSUMERS;
    __shared__ VStage<SKIN> stage[VERTEX_STAGES];
    __shared__ __align__(8) uint64_t full[VERTEX_STAGES], empty[VERTEX_STAGES];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        vtrace(0);
#pragma unroll
        for (int i = 0; i < VERTEX_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!second) pdl_wait();        // everything below reads what k_prep_draws wrote
    pdl_launch();
    if (threadIdx.x == 0) vtrace(1);

    if (warp == CONSUMERS / 32) {
        // ---- producer ----
        if (lane == 0) {
            // tiles are handed out by an atomic counter (zeroed by k_prep_draws): SMs that run ahead take more tiles.  The
            // producer claims a tile and fetches its list entry one iteration before it needs them, so neither the counter's
            // round trip to L2 nor the entry's is on the path of a tile's bulk copies.
            uint32_t t = blockIdx.x;                // the first tile needs no round trip; the counter hands out tiles gridDim.x, ...
            uint4 entry = t < nTiles ? __ldg(&tileList[t]) : make_uint4(0u, 0u, 0u, 0u);
            uint32_t tNext = gridDim.x + atomicAdd(tileCtr, 1u);
            uint4 entryNext = entry; uint32_t tAfter = 0;
            uint32_t sidx = 0, ph = 0;
            for (;; t = tNext, entry = entryNext, tNext = tAfter, ph ^= (sidx + 1 == (uint32_t)VERTEX_STAGES) ? 1u : 0u, sidx = (sidx + 1 == (uint32_t)VERTEX_STAGES) ? 0u : sidx + 1) {
                VStage<SKIN>& S = stage[sidx];
                if (t >= nTiles) {          // tell the consumers that the list is exhausted
                    mbar_wait(&empty[sidx], ph ^ 1u);
                    S.desc.base = NO_TILE;
                    mbar_arrive(&full[sidx]);
                    break;
                }
                // {tile, first draw, first draw of the next tile, end of the first draw's vertices} of the tile after this one
                entryNext = tNext < nTiles ? __ldg(&tileList[tNext]) : make_uint4(0u, 0u, 0u, 0u);
                tAfter = gridDim.x + atomicAdd(tileCtr, 1u);
                const uint32_t base = entry.x * (uint32_t)VERTEX_BLOCK;
                const uint32_t end = min(base + (uint32_t)VERTEX_BLOCK, totalV);
                const uint32_t dLo = entry.y, dHi = entry.z;
                // dHi is the draw of the first slot of the NEXT tile, so dLo == dHi covers this whole tile; otherwise the
                // tile is still uniform when the next draw starts at or after its end
                bool staged = dLo == dHi || entry.w >= end;
                const DrawRec& R = draws[dLo];
                uint32_t flags = 0, nL = 0, mv0 = 0, lmode = T3D_LIGHTING_DEFAULT, groups = 0, groups2 = 0;
                if (staged) {
                    flags = R.flags;
                    nL = (flags & T3D_DRAW_LIT) ? R.lightCount : 0;
                    groups = (uint32_t)R.nPoint | ((uint32_t)R.nSun << 16); groups2 = (uint32_t)R.nAmbient | ((uint32_t)R.lightsSorted << 16);
                    lmode = (flags & T3D_DRAW_HAS_MATERIAL) ? R.lightingMode : (uint8_t)T3D_LIGHTING_DEFAULT;
                    staged = nL <= (uint32_t)VERTEX_SMEM_LIGHTS;
                    if (SKIN && (flags & T3D_DRAW_SKINNED) && R.nBones > (uint32_t)VERTEX_SMEM_BONES) staged = false;
                    mv0 = R.a0 + (base - R.vbase);
                }
                if (t == blockIdx.x) { vtrace_set(15, flags); vtrace(2); }
                mbar_wait(&empty[sidx], ph ^ 1u);       // the consumers are done with the tile that used this stage before
                VStageDesc D;
                D.base = base; D.end = end; D.dLo = dLo; D.dHi = dHi; D.staged = staged ? 1u : 0u; D.mv0 = mv0; D.flags = flags; D.nLmode = nL | (lmode << 16);
                D.groups = groups; D.groups2 = groups2; D.pad[0] = D.pad[1] = 0;
                S.desc = D;
                if (!staged) { mbar_arrive(&full[sidx]); continue; }
                const uint32_t cnt = end - base;            // a multiple of 4: every draw's vertex range is
                const uint32_t vs = R.vstride;
                const bool needN = flags & (T3D_DRAW_LIT | T3D_DRAW_SKINNED);
                const bool skinned = SKIN && (flags & T3D_DRAW_SKINNED);
                uint32_t bytes = 12u * cnt + 128u + 64u * nL;
                if (needN) bytes += 12u * cnt;
                if (skinned) bytes += 24u * cnt + 64u * R.nBones;
                mbar_arrive_expect_tx(&full[sidx], bytes);
                const float* pp = R.pos + mv0;
#pragma unroll
                for (int c = 0; c < 3; c++) bulk_load(S.pos[c], pp + (size_t)c * vs, 4u * cnt, &full[sidx]);
                if (needN) {
                    const float* np = R.nrm + mv0;
#pragma unroll
                    for (int c = 0; c < 3; c++) bulk_load(S.nrm[c], np + (size_t)c * vs, 4u * cnt, &full[sidx]);
                }
                bulk_load(&S.U[0], R.mvp, 64u, &full[sidx]);
                bulk_load(&S.U[4], R.worldRot, 64u, &full[sidx]);
                if (nL) bulk_load(&S.U[8], lightsU + R.lightU, 64u * nL, &full[sidx]);
                if constexpr (SKIN) {
                    if (skinned) {
                        bulk_load(S.wgt, R.wgt + 4 * (size_t)mv0, 16u * cnt, &full[sidx]);
                        bulk_load(S.jnt, R.jnt + 4 * (size_t)mv0, 8u * cnt, &full[sidx]);
                        bulk_load(S.bones, bones + 16 * (size_t)R.boneOffset, 64u * R.nBones, &full[sidx]);
                    }
                }
                if (t == blockIdx.x) vtrace(3);
                vtrace(4);
            }
        }
        return;     // (the consumers keep the block alive; a `second` launch's wait at the end is theirs)
    }

    // ---- consumers ----
    uint32_t sidx = 0, ph = 0;
#ifdef T3D_VERTEX_TRACE
    uint32_t nTaken = 0;
#endif
    for (;; ph ^= (sidx + 1 == (uint32_t)VERTEX_STAGES) ? 1u : 0u, sidx = (sidx + 1 == (uint32_t)VERTEX_STAGES) ? 0u : sidx + 1) {
        VStage<SKIN>& S = stage[sidx];
        mbar_wait(&full[sidx], ph);
        const VStageDesc& D = S.desc;       // (read from shared memory where needed: a register copy of all of it spills)
        if (D.base == NO_TILE) break;
#ifdef T3D_VERTEX_TRACE
        if (threadIdx.x == 0 && nTaken == 0) vtrace(5);
#endif
        bool done = false;
        if constexpr (MODE == K1_STATIC && APPROX && CONSUMERS * 2 == VERTEX_BLOCK) {
            // the hot case, two vertices per thread: a staged tile, default lighting mode, light list grouped by type
            if (D.staged && (D.nLmode >> 16) == T3D_LIGHTING_DEFAULT && ((D.groups2 >> 16) != 0 || !(D.flags & T3D_DRAW_LIT))) {
                const uint32_t cnt = D.end - D.base;
                const bool hasB = threadIdx.x + CONSUMERS < cnt;       // (a missing second vertex repeats the first: no stale shared memory is read)
                if (threadIdx.x < cnt) shade2_static(S, D, threadIdx.x, hasB ? threadIdx.x + CONSUMERS : threadIdx.x, hasB, clipOut, clipXY, lightOut);
                done = true;
            }
        }
        for (uint32_t ti = threadIdx.x; !done && ti < D.end - D.base; ti += CONSUMERS) {
            const uint32_t slot = D.base + ti;
            if (D.staged) {
                const DrawRec& R = draws[D.dLo];
                const uint32_t mv = D.mv0 + ti;
                const Vec3 P = v3(S.pos[0][ti], S.pos[1][ti], S.pos[2][ti]);
                Vec3 N = v3(0, 0, 0);
                if (D.flags & (T3D_DRAW_LIT | T3D_DRAW_SKINNED)) N = v3(S.nrm[0][ti], S.nrm[1][ti], S.nrm[2][ti]);
                UniformsShared U; U.s = S.U; U.ti = ti;
                if constexpr (SKIN) { U.sWgt = S.wgt; U.sJnt = S.jnt; U.sBones = S.bones; }
                else { U.sWgt = nullptr; U.sJnt = nullptr; U.sBones = nullptr; }
                LightGroups G; G.nPoint = D.groups & 0xFFFFu; G.nSun = D.groups >> 16; G.nAmbient = D.groups2 & 0xFFFFu; G.sorted = (D.groups2 >> 16) != 0;
                shade_vertex<MODE, APPROX>(R, D.flags, D.nLmode & 0xFFFFu, G, (uint8_t)(D.nLmode >> 16), U, fc, slot, mv, P, N, lightsU, vols,
                                           clipOut, clipXY, lightOut, wposOut, unbillOut, tnrmOut);
            } else {
                const uint32_t d = find_draw(vbase, D.dLo, D.dHi, slot);
                const DrawRec& R = draws[d];
                const uint32_t flags = R.flags;
                const uint32_t mv = R.a0 + (slot - R.vbase);
                const uint32_t vs = R.vstride;
                const float* pp = R.pos + mv;
                Vec3 P = v3(__ldg(pp), __ldg(pp + vs), __ldg(pp + 2 * (size_t)vs)), N = v3(0, 0, 0);
                if (flags & (T3D_DRAW_LIT | T3D_DRAW_SKINNED)) {
                    const float* np = R.nrm + mv;
                    N = v3(__ldg(np), __ldg(np + vs), __ldg(np + 2 * (size_t)vs));
                }
                UniformsGlobal U;
                U.mvpP = R.mvp; U.sunP = R.worldRot; U.lightP = reinterpret_cast<const float4*>(lightsU + R.lightU);
                U.wgtP = R.wgt; U.jntP = R.jnt; U.bonesP = bones + 16 * (size_t)R.boneOffset; U.mv = mv;
                const uint8_t lightingMode = (flags & T3D_DRAW_HAS_MATERIAL) ? R.lightingMode : (uint8_t)T3D_LIGHTING_DEFAULT;
                LightGroups G; G.nPoint = R.nPoint; G.nSun = R.nSun; G.nAmbient = R.nAmbient; G.sorted = R.lightsSorted != 0;
                shade_vertex<MODE, APPROX>(R, flags, R.lightCount, G, lightingMode, U, fc, slot, mv, P, N, lightsU, vols,
                                           clipOut, clipXY, lightOut, wposOut, unbillOut, tnrmOut);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sidx]);
#ifdef T3D_VERTEX_TRACE
        if (threadIdx.x == 0) { if (nTaken == 0) vtrace(6); vtrace(7); nTaken++; vtrace_set(8, nTaken); }
#endif
    }
    if (second) pdl_wait();
}

// ------------------------------------------------------------------------------------------------
// K2: cull + depth key (model.go:855-969)
// ------------------------------------------------------------------------------------------------

// The keep / drop decision of one triangle (model.go:866-925) from the three per-vertex records k_vertex left:
// {X/w, Y/w, Z, W}, the divide already done with the cull's clamp of a negative w (model.go:866-875).
__device__ __forceinline__ bool cull_keep(const float4 c0, const float4 c1, const float4 c2, bool backface, float nearP, float farP) {
    const float x0 = c0.x, y0 = c0.y, x1 = c1.x, y1 = c1.y, x2 = c2.x, y2 = c2.y;
    bool keep = (c0.z + 1.0f >= nearP && c0.z < farP) || (c1.z + 1.0f >= nearP && c1.z < farP) ||
                (c2.z + 1.0f >= nearP && c2.z < farP);                  // model.go:877-879
    if (keep && ((x0 < -0.5f && x1 < -0.5f && x2 < -0.5f) || (x0 > 0.5f && x1 > 0.5f && x2 > 0.5f) ||
                 (y0 < -0.5f && y1 < -0.5f && y2 < -0.5f) || (y0 > 0.5f && y1 > 0.5f && y2 > 0.5f)))
        keep = false;                                                   // model.go:891-894
    if (keep && backface) {                                             // model.go:902-925
        const float n0x = x0 - x1, n0y = y0 - y1, n1x = x1 - x2, n1y = y1 - y2;
        const float nor = (n0x * n1y) - (n1x * n0y);
        if (nor < 0) keep = false;
    }
    return keep;
}

// ---- auto-subdivision (mesh.go:1877-1959): is triangle entry `mt` the one drawn for its base triangle? ----------------
// closestPointOnTri and its helpers: colliderTriangles.go:177-182, 206-281
__device__ __forceinline__ Vec3 closest_point_on_line(Vec3 point, Vec3 start, Vec3 end) {
    const Vec3 diff = v3_sub(end, start);
    const float dotA = v3_dot(v3_sub(point, start), diff);
    const float dotB = v3_dot(diff, diff);
    float d = dotA / dotB;
    if (d > 1.0f) d = 1.0f; else if (d < 0.0f) d = 0.0f;
    return v3_add(start, v3(diff.X * d, diff.Y * d, diff.Z * d));
}
__device__ __forceinline__ Vec3 closest_point_on_tri(Vec3 point, Vec3 v0, Vec3 v1, Vec3 v2, Vec3 normal) {
    const float distance = v3_dot(normal, v0);
    const float dist = v3_dot(normal, point) - distance;                       // collisionPlane.ClosestPoint
    const Vec3 planePoint = v3_sub(point, v3(normal.X * dist, normal.Y * dist, normal.Z * dist));
    {   // pointInsideTriangle
        const Vec3 ca = v3_sub(v2, v0), ba = v3_sub(v1, v0), pa = v3_sub(planePoint, v0);
        const float dot00 = v3_dot(ca, ca), dot01 = v3_dot(ca, ba), dot02 = v3_dot(ca, pa), dot11 = v3_dot(ba, ba), dot12 = v3_dot(ba, pa);
        const float invDenom = 1.0f / ((dot00 * dot11) - (dot01 * dot01));
        const float u = ((dot11 * dot02) - (dot01 * dot12)) * invDenom;
        const float v = ((dot00 * dot12) - (dot01 * dot02)) * invDenom;
        if ((u >= 0.0f) && (v >= 0.0f) && (u + v < 1.0f)) return planePoint;
    }
    const Vec3 ab = closest_point_on_line(point, v0, v1), bc = closest_point_on_line(point, v1, v2), ca = closest_point_on_line(point, v2, v0);
    Vec3 closest = ab;
    float closestDist = v3_dist2(point, ab);
    const float bcDist = v3_dist2(point, bc), caDist = v3_dist2(point, ca);
    if (bcDist < closestDist) { closest = bc; closestDist = bcDist; }
    if (caDist < closestDist) closest = ca;
    return closest;
}
// Triangle.handleSubdivision for the entry's base triangle, then Triangle.visible of the entry itself (model.go:753-755).
// Skinned draws and frames without levels draw base triangles only.  Rare, kept out of line.
__device__ __noinline__ bool subdiv_visible(const DrawRec* __restrict__ Rp, const FrameConst* __restrict__ fc, uint32_t mt) {
    const DrawRec& R = *Rp;
    const SubdivMesh S = *R.subdiv;
    const int level = S.level[mt];
    const uint32_t bt = S.base[mt];
    const int nLevels = S.nLevels[bt];
    int choice = -1;
    if (!(R.flags & T3D_DRAW_SKINNED) && fc->nSubLevels > 0 && nLevels > 0) {
        const uint32_t ts = R.tstride, vs = R.vstride;
        const uint32_t a = R.idx[bt], b = R.idx[ts + bt], c = R.idx[2 * (size_t)ts + bt];
        const Vec3 v0 = v3(R.pos[a], R.pos[vs + a], R.pos[2 * (size_t)vs + a]);
        const Vec3 v1 = v3(R.pos[b], R.pos[vs + b], R.pos[2 * (size_t)vs + b]);
        const Vec3 v2 = v3(R.pos[c], R.pos[vs + c], R.pos[2 * (size_t)vs + c]);
        const Vec3 nrm = v3(S.normal[3 * (size_t)bt], S.normal[3 * (size_t)bt + 1], S.normal[3 * (size_t)bt + 2]);
        const Vec3 cam = v3(R.invCam[0], R.invCam[1], R.invCam[2]);
        const float dist = v3_dist2(cam, closest_point_on_tri(cam, v0, v1, v2, nrm));
        int active = -1;
        float closestDistance = FLT_MAX;
        for (uint32_t i = 0; i < fc->nSubLevels; i++) {
            const float d2 = fc->subLevels[i][0];
            if (dist < d2 && d2 < closestDistance) { closestDistance = d2; active = (int)i; }
        }
        if (active >= 0) {
            float ms = S.span[bt];
            int subIndex = -1;
            const float maxSize = fc->subLevels[active][1];
            while (ms > maxSize && subIndex < 200) { ms /= 2.0f; subIndex++; }      // (an infinite MaxSpan would spin forever in the reference)
            if (subIndex >= nLevels) subIndex = nLevels - 1;
            choice = subIndex;
        }
    }
    return level == choice;
}

// Depth of a skinned triangle: distance from the camera to the mean of its skinned corners (model.go:840-864, 941-946).
__device__ __forceinline__ float skinned_depth(const float4 p0, const float4 p1, const float4 p2, const float* camPos) {
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    sx += p0.x; sy += p0.y; sz += p0.z;
    sx += p1.x; sy += p1.y; sz += p1.z;
    sx += p2.x; sy += p2.y; sz += p2.z;
    const float dx = camPos[0] - (sx / 3.0f), dy = camPos[1] - (sy / 3.0f), dz = camPos[2] - (sz / 3.0f);
    return dx * dx + dy * dy + dz * dz;
}

static const int TILE_THREADS = 256;
static const int TILE_ROUNDS = TILE_TRIS / TILE_THREADS;        // triangles per thread

// Large draws: one exclusive tile per block, TILE_ROUNDS triangles per thread, all loads of a thread issued
// before the first use.  Depth bits (or KEY_CULLED) go to tdepth; min / max / survivors are reduced per
// block and merged into DrawStat with three atomics.
#ifndef T3D_CULL_MINBLOCKS
#define T3D_CULL_MINBLOCKS 5
#endif
template <bool SUBDIV>      // SUBDIV: the frame holds auto-subdivided meshes (mesh.go:1877-1959); the other instantiation carries none of it
__global__ void __launch_bounds__(TILE_THREADS, T3D_CULL_MINBLOCKS)
k_cull_large(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ largeList,
             const __grid_constant__ FrameConst fc, const float4* __restrict__ clip, const float4* __restrict__ wpos,
             uint32_t* __restrict__ tdepth, DrawStat* __restrict__ stats) {
    __shared__ uint32_t sMin[TILE_THREADS / 32], sMax[TILE_THREADS / 32], sCnt[TILE_THREADS / 32];
    const TriTile T = tiles[__ldg(&largeList[blockIdx.x])];
    const DrawRec& R = draws[T.dFirst];
    const uint32_t flags = R.flags;
    const bool skinned = flags & T3D_DRAW_SKINNED;
    const bool backface = (flags & T3D_DRAW_HAS_MATERIAL) && (flags & T3D_DRAW_BACKFACE_CULLING);
    const uint32_t ts = R.tstride;
    const uint32_t mt0 = R.t0 + (T.start - R.tbase);
    const uint32_t vb = R.vbase - R.a0;
    const uint32_t* ip = R.idx + mt0;
    const float* cp = R.ctr + mt0;
    const float icx = R.invCam[0], icy = R.invCam[1], icz = R.invCam[2];

    uint32_t ia[TILE_ROUNDS], ib[TILE_ROUNDS], ic[TILE_ROUNDS];
    float cx[TILE_ROUNDS], cy[TILE_ROUNDS], cz[TILE_ROUNDS];
#pragma unroll
    for (int r = 0; r < TILE_ROUNDS; r++) {
        const uint32_t i = r * TILE_THREADS + threadIdx.x;
        ia[r] = ib[r] = ic[r] = 0; cx[r] = cy[r] = cz[r] = 0.0f;
        if (i < T.count) {
            ia[r] = __ldg(ip + i); ib[r] = __ldg(ip + ts + i); ic[r] = __ldg(ip + 2 * (size_t)ts + i);
            if (!skinned) { cx[r] = __ldg(cp + i); cy[r] = __ldg(cp + ts + i); cz[r] = __ldg(cp + 2 * (size_t)ts + i); }
        }
    }
    pdl_wait();             // the index / centre loads above only touch the mesh and k_prep_draws' output; clip is k_vertex's
    pdl_launch();
    uint32_t mn = 0xFFFFFFFFu, mx = 0u, cnt = 0u;
#pragma unroll
    for (int r = 0; r < TILE_ROUNDS; r++) {
        const uint32_t i = r * TILE_THREADS + threadIdx.x;
        if (i < T.count) {
            const float4 c0 = __ldg(&clip[vb + ia[r]]), c1 = __ldg(&clip[vb + ib[r]]), c2 = __ldg(&clip[vb + ic[r]]);
            bool keep = !(flags & T3D_DRAW_FRUSTUM_CULLED) && cull_keep(c0, c1, c2, backface, fc.nearP, fc.farP);
            if (SUBDIV && keep && R.subdiv) keep = subdiv_visible(&R, &fc, mt0 + i);       // model.go:753-755
            float depth = 0.0f;
            if (keep) {                                                 // model.go:933-955
                if (skinned) depth = skinned_depth(__ldg(&wpos[vb + ia[r]]), __ldg(&wpos[vb + ib[r]]), __ldg(&wpos[vb + ic[r]]), fc.camPos);
                else { const float dx = icx - cx[r], dy = icy - cy[r], dz = icz - cz[r]; depth = dx * dx + dy * dy + dz * dz; }
                if (isnan(depth) || isinf(depth)) keep = false;        // model.go:957-960
            }
            const uint32_t bits = __float_as_uint(depth);
            tdepth[T.start + i] = keep ? bits : KEY_CULLED;
            if (keep) { mn = min(mn, bits); mx = max(mx, bits); cnt++; }
        }
    }
    const unsigned full = 0xFFFFFFFFu;
    mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx); cnt = __reduce_add_sync(full, cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sMin[warp] = mn; sMax[warp] = mx; sCnt[warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < TILE_THREADS / 32 ? sMin[lane] : 0xFFFFFFFFu;
        mx = lane < TILE_THREADS / 32 ? sMax[lane] : 0u;
        cnt = lane < TILE_THREADS / 32 ? sCnt[lane] : 0u;
        mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx); cnt = __reduce_add_sync(full, cnt);
        if (lane == 0 && cnt) {
            atomicMin(&stats[T.dFirst].minBits, mn);
            atomicMax(&stats[T.dFirst].maxBits, mx);
            atomicAdd(&stats[T.dFirst].survivors, cnt);
        }
    }
}

// ---- k_cull_sort: cull + depth key + the segmented 512-bin sort of the large draws in ONE launch -------------------------
// A draw of 9 .. HUGE_TILES tiles needs two draw-wide facts before its triangles can be placed: its depth range (the bin
// expression reads min / max over the survivors, sortingTriangles.go:38-64) and the per-tile bin counts of all its tiles.  The
// three-kernel path (k_cull_large -> k_bin_hist -> k_scatter, t3d_sort.cu) gets them from kernel boundaries and carries the
// depth keys through global memory in between.  Here the blocks are persistent and take tiles in ticket order from an atomic
// counter; a block that has culled its tile publishes min / max / count and arrives on the draw's counter, waits until all
// tiles of the draw have arrived, bins its survivors (still in registers), publishes its 512 counts, arrives again, waits
// again, then sums the rows of the draw's tiles itself and scatters.  No depth key and no bin rank ever leaves the SM.
//   Why the waits cannot deadlock: a block only waits for tiles of its own draw; those hold tickets within HUGE_TILES of its
//   own; tickets are taken in order by running blocks, and neither arrival needs anything from a later ticket.  So the running
//   block with the lowest ticket always finds its draw's earlier tiles arrived, and the later ones either running -- they
//   arrive without waiting -- or not yet taken, in which case a block that is not waiting exists or is about to start (the
//   grid is sized to be resident at once; blocks still queued behind the previous kernel's tail start as that drains).
// DrawStat.done counts both arrivals: nTilesOfDraw after the cull, 2 x nTilesOfDraw after the counts.
// One extra block computes where every tile's triangles start in the lists (t3d_sort.cu: tile_scan_chunk) once every tile of
// this launch has passed its cull -- the other cull kernels ran before this one.
__device__ __forceinline__ void spin_until(const uint32_t* word, uint32_t target) {
    uint32_t v;
    for (uint32_t polls = 0;; polls++) {
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
        if (v >= target) break;
        __nanosleep(32);
        // (a frame's tiles arrive within microseconds; a wait of seconds means a broken invariant, and then a wrong frame that the
        // survivor-count check of t3d_frame_results reports beats a GPU that never comes back)
        if (polls > (1u << 26)) break;
    }
}
template <bool SUBDIV>
__global__ void __launch_bounds__(TILE_THREADS, T3D_FUSED_MINBLOCKS)
k_cull_sort(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ fusedList, uint32_t nFused,
            uint32_t* __restrict__ ctr /* [0] next ticket, [1] tiles past their cull */, const __grid_constant__ FrameConst fc,
            const float4* __restrict__ clip, const float4* __restrict__ wpos, DrawStat* __restrict__ stats,
            uint32_t* __restrict__ hist, uint32_t* __restrict__ sorted,
            uint32_t scanTiles /* 0: no scan block */, uint32_t* __restrict__ tileSurv, uint32_t* __restrict__ tileOut, uint32_t* __restrict__ chunkSum) {
    __shared__ __align__(16) uint16_t sCnt[8][BIN_COUNT];
    __shared__ uint32_t sBase[BIN_COUNT];
    __shared__ uint32_t sMin[8], sMax[8], sNum[8], sWarp[32];
    __shared__ uint32_t sTicket, sLo, sHi;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ROUNDS = TILE_ROUNDS;
    const uint32_t nWorkers = gridDim.x - (scanTiles ? 1u : 0u);
    if (blockIdx.x == nWorkers) {       // the scan block
        pdl_wait(); pdl_launch();
        if (threadIdx.x == 0) { spin_until(&ctr[1], nFused); __threadfence(); }
        __syncthreads();
        tile_scan_chunk<TILE_THREADS>(tiles, scanTiles, 0, stats, tileSurv, tileOut, chunkSum, sWarp);      // scanTiles <= SCAN_CHUNK
        return;
    }
    uint32_t t = blockIdx.x;
    bool first = true;
    while (t < nFused) {
        const TriTile T = tiles[__ldg(&fusedList[t])];
        const uint32_t d = T.dFirst, nT = T.nTilesOfDraw, j = T.info & ~TILE_LARGE;
        const DrawRec& R = draws[d];
        const uint32_t flags = R.flags;
        const bool skinned = flags & T3D_DRAW_SKINNED;
        const bool backface = (flags & T3D_DRAW_HAS_MATERIAL) && (flags & T3D_DRAW_BACKFACE_CULLING);
        const uint32_t ts = R.tstride;
        const uint32_t mt0 = R.t0 + (T.start - R.tbase);
        const uint32_t vb = R.vbase - R.a0;
        const uint32_t* ip = R.idx + mt0;
        const float* cp = R.ctr + mt0;
        const float icx = R.invCam[0], icy = R.invCam[1], icz = R.invCam[2];
        // ---- cull + depth (model.go:855-969).  Warp w owns slots [w * 128, w * 128 + 128) in rounds of 32: arrival order is
        //      (warp, round, lane) == arena order, which is what the stable ranks below rest on.
        uint32_t ia[ROUNDS], ib[ROUNDS], ic[ROUNDS];
        float cx[ROUNDS], cy[ROUNDS], cz[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            const uint32_t i = warp * (ROUNDS * 32) + r * 32 + lane;
            ia[r] = ib[r] = ic[r] = 0; cx[r] = cy[r] = cz[r] = 0.0f;
            if (i < T.count) {
                ia[r] = __ldg(ip + i); ib[r] = __ldg(ip + ts + i); ic[r] = __ldg(ip + 2 * (size_t)ts + i);
                if (!skinned) { cx[r] = __ldg(cp + i); cy[r] = __ldg(cp + ts + i); cz[r] = __ldg(cp + 2 * (size_t)ts + i); }
            }
        }
        {   // clear the per-warp bin counts for this tile
            uint4* z = reinterpret_cast<uint4*>(&sCnt[0][0]);
            z[threadIdx.x] = make_uint4(0, 0, 0, 0); z[threadIdx.x + TILE_THREADS] = make_uint4(0, 0, 0, 0);
        }
        if (first) { pdl_wait(); pdl_launch(); first = false; }       // clip / wpos are k_vertex's; everything above is mesh data and k_prep_draws' records
        uint32_t key[ROUNDS];
        uint32_t mn = 0xFFFFFFFFu, mx = 0u, cnt = 0u;
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            const uint32_t i = warp * (ROUNDS * 32) + r * 32 + lane;
            key[r] = KEY_CULLED;
            if (i < T.count) {
                const float4 c0 = __ldg(&clip[vb + ia[r]]), c1 = __ldg(&clip[vb + ib[r]]), c2 = __ldg(&clip[vb + ic[r]]);
                bool keep = !(flags & T3D_DRAW_FRUSTUM_CULLED) && cull_keep(c0, c1, c2, backface, fc.nearP, fc.farP);
                if (SUBDIV && keep && R.subdiv) keep = subdiv_visible(&R, &fc, mt0 + i);       // model.go:753-755
                float depth = 0.0f;
                if (keep) {                                                 // model.go:933-955
                    if (skinned) depth = skinned_depth(__ldg(&wpos[vb + ia[r]]), __ldg(&wpos[vb + ib[r]]), __ldg(&wpos[vb + ic[r]]), fc.camPos);
                    else { const float dx = icx - cx[r], dy = icy - cy[r], dz = icz - cz[r]; depth = dx * dx + dy * dy + dz * dz; }
                    if (isnan(depth) || isinf(depth)) keep = false;        // model.go:957-960
                }
                if (keep) { key[r] = __float_as_uint(depth); mn = min(mn, key[r]); mx = max(mx, key[r]); cnt++; }
            }
        }
        mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx); cnt = __reduce_add_sync(full, cnt);
        if (lane == 0) { sMin[warp] = mn; sMax[warp] = mx; sNum[warp] = cnt; }
        __syncthreads();
        if (warp == 0) {
            mn = lane < 8 ? sMin[lane] : 0xFFFFFFFFu; mx = lane < 8 ? sMax[lane] : 0u; cnt = lane < 8 ? sNum[lane] : 0u;
            mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx); cnt = __reduce_add_sync(full, cnt);
            if (lane == 0) {
                if (cnt) { atomicMin(&stats[d].minBits, mn); atomicMax(&stats[d].maxBits, mx); atomicAdd(&stats[d].survivors, cnt); }
                __threadfence();
                atomicAdd(&stats[d].done, 1u);                  // first arrival: this tile's part of min / max / survivors is in
                atomicAdd(&ctr[1], 1u);
                spin_until(&stats[d].done, nT);
                __threadfence();
                sLo = __ldcg(&stats[d].minBits); sHi = __ldcg(&stats[d].maxBits);
            }
        }
        __syncthreads();
        // ---- bin (sortingTriangles.go:38-64) and stable rank inside the tile: per-warp counts, round by round
        const uint32_t minBits = sLo, maxBits = sHi;
        uint32_t rank[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            if (key[r] != KEY_CULLED) key[r] = bin_key(key[r], minBits, maxBits, T.sortMode);
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
        uint2* rows = reinterpret_cast<uint2*>(hist + (size_t)T.row0 * BIN_COUNT);
        {   // exclusive prefix over the 8 warps (two adjacent bins per thread in one 32-bit word; counts <= 1024); the totals are
            // this tile's row of the draw's [tiles x 512] counts
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t* word = reinterpret_cast<uint32_t*>(&sCnt[w][0]) + threadIdx.x;
                const uint32_t c = *word;
                *word = run;
                run += c;
            }
            rows[(size_t)j * (BIN_COUNT / 2) + threadIdx.x] = make_uint2(run & 0xFFFFu, run >> 16);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&stats[d].done, 1u); spin_until(&stats[d].done, 2u * nT); __threadfence();
            // The next ticket is taken only now, when nothing of this tile waits any more: a block that held a second, unprocessed
            // ticket while it waits could be waiting for itself.  Its round trip hides behind the row sums and the scatter.
            sTicket = nWorkers + atomicAdd(&ctr[0], 1u);
        }
        __syncthreads();
        {   // where this tile's share of every bin starts: bins in visiting order, tiles in arena order inside a bin.  Thread x
            // owns bins 2x, 2x + 1; it sums them over the draw's tiles (one 8-byte load per tile row, coalesced over the block).
            uint2 tot = make_uint2(0u, 0u), mine = make_uint2(0u, 0u);
            for (uint32_t t0 = 0; t0 < nT; t0 += 8) {
                uint2 c[8];
#pragma unroll
                for (int u = 0; u < 8; u++) c[u] = (t0 + u < nT) ? __ldcg(&rows[(size_t)(t0 + u) * (BIN_COUNT / 2) + threadIdx.x]) : make_uint2(0u, 0u);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (t0 + u == j) mine = tot;
                    tot.x += c[u].x; tot.y += c[u].y;
                }
            }
            const uint32_t base = block_exclusive_scan_256(tot.x + tot.y, sWarp, nullptr);
            sBase[2 * threadIdx.x] = T.tbase + base + mine.x;
            sBase[2 * threadIdx.x + 1] = T.tbase + base + tot.x + mine.y;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            if (key[r] != KEY_CULLED)
                sorted[sBase[key[r]] + sCnt[warp][key[r]] + rank[r]] = T.start + warp * (ROUNDS * 32) + r * 32 + lane;
        }
        t = sTicket;
        __syncthreads();            // the shared arrays are reused by the next tile
    }
    if (first) { pdl_wait(); pdl_launch(); }
}

// One stable counting-sort pass over the block's (at most TILE_TRIS) keys in shared memory, on the NB-valued
// digit (key >> SHIFT) & (NB - 1).  Warp w owns slots [w * 128, w * 128 + 128) and walks them in rounds of 32,
// so arrival order is (warp, round, lane) == slot order.  Entries at or beyond n, and entries equal to
// PK_INVALID, do not take part; the others land in out[0 .. valid) in (digit, arrival) order.
static const uint32_t PK_INVALID = 0xFFFFFFFFu;
template <int SHIFT, int NB>
__device__ __forceinline__ void smem_radix_pass(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                                uint16_t* __restrict__ sCnt /*[8][NB]*/, uint32_t* __restrict__ sBucket /*[NB]*/,
                                                uint32_t* __restrict__ sWarp /*[8]*/) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WORDS = NB / 2;                       // 32-bit words per warp row (two 16-bit counts each)
    constexpr int WPT = WORDS / TILE_THREADS;           // words per thread in the bucket scan
    static_assert(WPT >= 1 && WORDS % TILE_THREADS == 0, "bucket scan layout");
    {
        uint4* z = reinterpret_cast<uint4*>(sCnt);
        for (int i = threadIdx.x; i < 8 * NB * 2 / 16; i += TILE_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    uint32_t key[TILE_ROUNDS], rank[TILE_ROUNDS];
#pragma unroll
    for (int r = 0; r < TILE_ROUNDS; r++) {
        const uint32_t i = warp * (TILE_ROUNDS * 32) + r * 32 + lane;
        key[r] = (i < n) ? in[i] : PK_INVALID;
        const bool valid = key[r] != PK_INVALID;
        const uint32_t digit = (key[r] >> SHIFT) & (NB - 1);
        const unsigned peers = __match_any_sync(full, valid ? digit : 0xFFFFFFFFu);
        const uint32_t inGroup = __popc(peers & ((1u << lane) - 1u));
        uint32_t before = 0;
        if (valid && inGroup == 0) {
            before = sCnt[warp * NB + digit];
            sCnt[warp * NB + digit] = (uint16_t)(before + __popc(peers));
        }
        rank[r] = __shfl_sync(full, before, __ffs(peers) - 1) + inGroup;
        __syncwarp();
    }
    __syncthreads();
    {   // per bucket: exclusive prefix over the warps; then bucket bases = exclusive scan of the bucket totals
        uint32_t* words = reinterpret_cast<uint32_t*>(sCnt);
        uint32_t tot[WPT];
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < WPT; q++) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t* p = words + w * WORDS + threadIdx.x * WPT + q;
                const uint32_t c = *p;
                *p = run;
                run += c;                                   // two 16-bit lanes, each at most TILE_TRIS: no carry
            }
            tot[q] = run;
            sum += (run & 0xFFFFu) + (run >> 16);
        }
        uint32_t base = block_exclusive_scan_256(sum, sWarp, nullptr);
#pragma unroll
        for (int q = 0; q < WPT; q++) {
            const uint32_t b = 2 * (threadIdx.x * WPT + q);
            sBucket[b] = base; base += tot[q] & 0xFFFFu;
            sBucket[b + 1] = base; base += tot[q] >> 16;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TILE_ROUNDS; r++) {
        if (key[r] != PK_INVALID) {
            const uint32_t digit = (key[r] >> SHIFT) & (NB - 1);
            out[sBucket[digit] + sCnt[warp * NB + digit] + rank[r]] = key[r];
        }
    }
    __syncthreads();
}

// Packed tiles: every draw of the tile lies wholly inside it, so cull, per-draw min / max, binning and the
// stable (draw, bin) ordering all happen in shared memory.  Survivors carry the composite
//   draw-in-tile (10 bits) << 20 | bin rank (10) << 10 | slot-in-tile (10)
// through two stable counting-sort passes (bin rank, then draw; the second is skipped for a one-draw tile) and
// come out in (draw, bin, arrival) order -- the order sortingTriangleBucket.ForEach visits them in.
template <bool SUBDIV>
__global__ void __launch_bounds__(TILE_THREADS)
k_tile_packed(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ tbase, const TriTile* __restrict__ tiles,
              const uint32_t* __restrict__ packedList, const __grid_constant__ FrameConst fc,
              const float4* __restrict__ clip, const float4* __restrict__ wpos,
              uint32_t* __restrict__ sorted, uint32_t* __restrict__ tileSurv, DrawStat* __restrict__ stats) {
    // The per-draw tables of the cull phase and the counters of the sort phase share storage.
    __shared__ __align__(16) uint8_t sRaw[16 * 1024 + 16];
    __shared__ uint32_t sKeyA[TILE_TRIS];           // depth bits, then composite keys
    __shared__ uint32_t sKeyB[TILE_TRIS];
    __shared__ uint32_t sBucket[TILE_TRIS];
    __shared__ uint16_t sSeg[TILE_TRIS];
    __shared__ uint32_t sWarp[8];
    uint32_t* sBase = reinterpret_cast<uint32_t*>(sRaw);            // [TILE_TRIS + 1] tbase of the tile's draws
    uint32_t* sMin = sBucket;                                       // per-draw stats live in sBucket / sKeyB until the sort starts
    uint32_t* sMax = sKeyB;
    uint32_t* sCntSeg = sBase + TILE_TRIS + 4;                      // [TILE_TRIS] -- needs 8 KB + 16: see static_assert
    static_assert(sizeof(sRaw) >= (2 * TILE_TRIS + 4) * sizeof(uint32_t) && sizeof(sRaw) >= 8 * 1024 * sizeof(uint16_t), "shared carve-up");
    uint16_t* sCnt = reinterpret_cast<uint16_t*>(sRaw);             // [8][<=1024] sort phase

    const uint32_t tileId = __ldg(&packedList[blockIdx.x]);
    const TriTile T = tiles[tileId];
    const uint32_t nSeg = T.info;
    const unsigned full = 0xFFFFFFFFu;

    for (uint32_t s = threadIdx.x; s <= nSeg; s += TILE_THREADS) sBase[s] = __ldg(&tbase[T.dFirst + s]);
    for (uint32_t s = threadIdx.x; s < nSeg; s += TILE_THREADS) { sMin[s] = 0x7F7FFFFFu; sMax[s] = 0u; sCntSeg[s] = 0; }
    pdl_wait();
    pdl_launch();
    __syncthreads();

    // ---- cull + depth ----
#pragma unroll 1
    for (int r = 0; r < TILE_ROUNDS; r++) {
        const uint32_t i = r * TILE_THREADS + threadIdx.x;
        const bool active = i < T.count;
        uint32_t seg = 0xFFFFu, bits = KEY_CULLED;
        bool keep = false;
        if (active) {
            const uint32_t slot = T.start + i;
            uint32_t lo = 0, hi = nSeg - 1;                            // largest seg with sBase[seg] <= slot
            while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (sBase[mid] <= slot) lo = mid; else hi = mid - 1; }
            seg = lo;
            const DrawRec& R = draws[T.dFirst + seg];
            const uint32_t flags = R.flags;
            const uint32_t ts = R.tstride;
            const uint32_t mt = R.t0 + (slot - R.tbase);
            const uint32_t vb = R.vbase - R.a0;
            const uint32_t i0 = __ldg(R.idx + mt), i1 = __ldg(R.idx + ts + mt), i2 = __ldg(R.idx + 2 * (size_t)ts + mt);
            const float4 c0 = __ldg(&clip[vb + i0]), c1 = __ldg(&clip[vb + i1]), c2 = __ldg(&clip[vb + i2]);
            keep = !(flags & T3D_DRAW_FRUSTUM_CULLED) && cull_keep(c0, c1, c2, (flags & T3D_DRAW_HAS_MATERIAL) && (flags & T3D_DRAW_BACKFACE_CULLING), fc.nearP, fc.farP);
            if (SUBDIV && keep && R.subdiv) keep = subdiv_visible(&R, &fc, mt);
            if (keep) {
                float depth;
                if (flags & T3D_DRAW_SKINNED) depth = skinned_depth(__ldg(&wpos[vb + i0]), __ldg(&wpos[vb + i1]), __ldg(&wpos[vb + i2]), fc.camPos);
                else {
                    const float dx = R.invCam[0] - __ldg(R.ctr + mt), dy = R.invCam[1] - __ldg(R.ctr + ts + mt), dz = R.invCam[2] - __ldg(R.ctr + 2 * (size_t)ts + mt);
                    depth = dx * dx + dy * dy + dz * dz;
                }
                if (isnan(depth) || isinf(depth)) keep = false;
                else bits = __float_as_uint(depth);
            }
            sSeg[i] = (uint16_t)seg;
        }
        sKeyA[i] = keep ? bits : KEY_CULLED;
        // per-draw min / max / survivors: one set of shared atomics per warp when the warp sits in one draw
        const uint32_t seg0 = __shfl_sync(full, seg, 0);
        const bool uniform = __all_sync(full, (!active) || seg == seg0);
        if (uniform) {
            const unsigned km = __ballot_sync(full, keep);
            if (km) {
                const uint32_t mn = __reduce_min_sync(full, keep ? bits : 0xFFFFFFFFu);
                const uint32_t mx = __reduce_max_sync(full, keep ? bits : 0u);
                if ((threadIdx.x & 31) == 0) { atomicMin(&sMin[seg0], mn); atomicMax(&sMax[seg0], mx); atomicAdd(&sCntSeg[seg0], (uint32_t)__popc(km)); }
            }
        } else if (keep) {
            atomicMin(&sMin[seg], bits); atomicMax(&sMax[seg], bits); atomicAdd(&sCntSeg[seg], 1u);
        }
    }
    __syncthreads();

    // ---- composite keys; per-draw results out (their storage is reused by the sort) ----
    for (uint32_t i = threadIdx.x; i < T.count; i += TILE_THREADS) {
        const uint32_t bits = sKeyA[i], seg = sSeg[i];
        uint32_t key = PK_INVALID;
        if (bits != KEY_CULLED) key = (seg << 20) | (bin_key(bits, sMin[seg], sMax[seg], draws[T.dFirst + seg].sortMode) << 10) | i;
        sKeyA[i] = key;
    }
    for (uint32_t s = threadIdx.x; s < nSeg; s += TILE_THREADS) {
        DrawStat st; st.minBits = sMin[s]; st.maxBits = sMax[s]; st.survivors = sCntSeg[s]; st.done = 0;
        if (draws[T.dFirst + s].flags & T3D_DRAW_FRUSTUM_CULLED) st.minBits = DRAW_CULLED_MARK;     // (k_prep_draws' mark for the host survives)
        stats[T.dFirst + s] = st;
    }
    __syncthreads();

    // ---- stable order: by bin rank, then (tiles of several draws) by draw ----
    uint32_t nSurv = 0;
    for (uint32_t s = threadIdx.x; s < nSeg; s += TILE_THREADS) nSurv += sCntSeg[s];
    nSurv = __reduce_add_sync(full, nSurv);
    __syncthreads();                                    // sCntSeg is about to be overwritten by the pass counters
    if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = nSurv;
    __syncthreads();
    nSurv = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) nSurv += sWarp[w];
    __syncthreads();
    smem_radix_pass<10, 512>(sKeyA, sKeyB, T.count, sCnt, sBucket, sWarp);
    const uint32_t* res = sKeyB;
    if (nSeg > 1) {
        smem_radix_pass<20, 1024>(sKeyB, sKeyA, nSurv, sCnt, sBucket, sWarp);
        res = sKeyA;
    }

    // ---- survivors out: value = draw-in-tile << 20 | bin rank << 10 | slot-in-tile ----
    for (uint32_t i = threadIdx.x; i < nSurv; i += TILE_THREADS) sorted[T.start + i] = res[i];
    if (threadIdx.x == 0) tileSurv[tileId] = nSurv;
}

// Medium draws (more than one tile, at most MEDIUM_TRIS triangles): one block owns the whole draw, so the
// depth keys never leave shared memory -- cull every triangle, reduce min / max in the block, bin, count the
// 512 bins, then walk the draw's tiles in arena order ranking each survivor inside its bin (the same
// warp-contiguous ranking as smem_radix_pass) and store sorted[tbase + bin start + rank] = arena slot.
// No depth array, histogram rows or atomics in global memory.  The draw's tiles stay in the tile table as
// "large" tiles for the tile scan and the emit.
template <bool SUBDIV>
__global__ void __launch_bounds__(TILE_THREADS)
k_draw_medium(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ mediumList,
              const __grid_constant__ FrameConst fc, const float4* __restrict__ clip, const float4* __restrict__ wpos,
              uint32_t* __restrict__ sorted, DrawStat* __restrict__ stats) {
    __shared__ uint32_t sKey[MEDIUM_TRIS];              // depth bits, then bin ranks (KEY_CULLED = dropped)
    __shared__ uint32_t sHist[BIN_COUNT];               // bin counts, then the running start of each bin
    __shared__ __align__(16) uint16_t sCnt[8 * BIN_COUNT];
    __shared__ uint32_t sRed[3][TILE_THREADS / 32];
    __shared__ uint32_t sWarp[8];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const TriTile T = tiles[__ldg(&mediumList[blockIdx.x])];        // the draw's first tile
    const uint32_t d = T.dFirst;
    const DrawRec& R = draws[d];
    const uint32_t flags = R.flags;
    const bool skinned = flags & T3D_DRAW_SKINNED;
    const bool backface = (flags & T3D_DRAW_HAS_MATERIAL) && (flags & T3D_DRAW_BACKFACE_CULLING);
    const uint32_t ts = R.tstride, nt = R.nt, tb = R.tbase;
    const uint32_t vb = R.vbase - R.a0;
    const uint32_t* ip = R.idx + R.t0;
    const float* cp = R.ctr + R.t0;
    const float icx = R.invCam[0], icy = R.invCam[1], icz = R.invCam[2];
    sHist[threadIdx.x] = 0; sHist[threadIdx.x + TILE_THREADS] = 0;
    pdl_wait();
    pdl_launch();

    // ---- cull + depth ----
    uint32_t mn = 0xFFFFFFFFu, mx = 0u, cnt = 0u;
    for (uint32_t i = threadIdx.x; i < nt; i += TILE_THREADS) {
        const uint32_t i0 = __ldg(ip + i), i1 = __ldg(ip + ts + i), i2 = __ldg(ip + 2 * (size_t)ts + i);
        const float4 c0 = __ldg(&clip[vb + i0]), c1 = __ldg(&clip[vb + i1]), c2 = __ldg(&clip[vb + i2]);
        bool keep = !(R.flags & T3D_DRAW_FRUSTUM_CULLED) && cull_keep(c0, c1, c2, backface, fc.nearP, fc.farP);
        if (SUBDIV && keep && R.subdiv) keep = subdiv_visible(&R, &fc, R.t0 + i);
        float depth = 0.0f;
        if (keep) {                                                     // model.go:933-955
            if (skinned) depth = skinned_depth(__ldg(&wpos[vb + i0]), __ldg(&wpos[vb + i1]), __ldg(&wpos[vb + i2]), fc.camPos);
            else { const float dx = icx - __ldg(cp + i), dy = icy - __ldg(cp + ts + i), dz = icz - __ldg(cp + 2 * (size_t)ts + i); depth = dx * dx + dy * dy + dz * dz; }
            if (isnan(depth) || isinf(depth)) keep = false;            // model.go:957-960
        }
        const uint32_t bits = __float_as_uint(depth);
        sKey[i] = keep ? bits : KEY_CULLED;
        if (keep) { mn = min(mn, bits); mx = max(mx, bits); cnt++; }
    }
    mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx); cnt = __reduce_add_sync(full, cnt);
    if (lane == 0) { sRed[0][warp] = mn; sRed[1][warp] = mx; sRed[2][warp] = cnt; }
    __syncthreads();
    mn = 0xFFFFFFFFu; mx = 0u; cnt = 0u;
#pragma unroll
    for (int w = 0; w < TILE_THREADS / 32; w++) { mn = min(mn, sRed[0][w]); mx = max(mx, sRed[1][w]); cnt += sRed[2][w]; }
    if (threadIdx.x == 0) {
        DrawStat st; st.minBits = cnt ? mn : 0x7F7FFFFFu; st.maxBits = cnt ? mx : 0u; st.survivors = cnt; st.done = 0;
        if (R.flags & T3D_DRAW_FRUSTUM_CULLED) st.minBits = DRAW_CULLED_MARK;       // (k_prep_draws' mark for the host survives)
        stats[d] = st;
    }
    if (cnt == 0) return;
    mn = min(mn, 0x7F7FFFFFu);

    // ---- bin rank + counts ----
    const uint32_t sortMode = R.sortMode;
    for (uint32_t i = threadIdx.x; i < nt; i += TILE_THREADS) {
        const uint32_t bits = sKey[i];
        if (bits != KEY_CULLED) {
            const uint32_t key = bin_key(bits, mn, mx, sortMode);
            sKey[i] = key;
            atomicAdd(&sHist[key], 1u);
        }
    }
    __syncthreads();
    {   // counts -> start of every bin in the draw's sorted run (thread owns bins 2t, 2t+1)
        const uint32_t c0 = sHist[2 * threadIdx.x], c1 = sHist[2 * threadIdx.x + 1];
        const uint32_t base = block_exclusive_scan_256(c0 + c1, sWarp, nullptr);
        sHist[2 * threadIdx.x] = base; sHist[2 * threadIdx.x + 1] = base + c0;
    }
    __syncthreads();

    // ---- stable rank inside the bin, tile by tile in arena order ----
    for (uint32_t t0 = 0; t0 < nt; t0 += TILE_TRIS) {
        {
            uint4* z = reinterpret_cast<uint4*>(sCnt);
            z[threadIdx.x] = make_uint4(0, 0, 0, 0); z[threadIdx.x + TILE_THREADS] = make_uint4(0, 0, 0, 0);
            static_assert(sizeof(sCnt) == 2 * TILE_THREADS * sizeof(uint4), "two 128-bit stores per thread clear the counts");
        }
        __syncthreads();
        uint32_t key[TILE_ROUNDS], rank[TILE_ROUNDS];
#pragma unroll
        for (int r = 0; r < TILE_ROUNDS; r++) {
            const uint32_t i = t0 + warp * (TILE_ROUNDS * 32) + r * 32 + lane;
            key[r] = (i < nt) ? sKey[i] : KEY_CULLED;
            const unsigned peers = __match_any_sync(full, key[r]);
            const uint32_t inGroup = __popc(peers & ((1u << lane) - 1u));
            uint32_t before = 0;
            if (key[r] != KEY_CULLED && inGroup == 0) {
                before = sCnt[warp * BIN_COUNT + key[r]];
                sCnt[warp * BIN_COUNT + key[r]] = (uint16_t)(before + __popc(peers));
            }
            rank[r] = __shfl_sync(full, before, __ffs(peers) - 1) + inGroup;
            __syncwarp();
        }
        __syncthreads();
        uint32_t tot;
        {   // exclusive prefix over the warps (two bins per thread, packed), tile totals kept for the bin starts
            uint32_t* words = reinterpret_cast<uint32_t*>(sCnt);
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t* p = words + w * (BIN_COUNT / 2) + threadIdx.x;
                const uint32_t c = *p;
                *p = run;
                run += c;
            }
            tot = run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TILE_ROUNDS; r++) {
            if (key[r] != KEY_CULLED)
                sorted[tb + sHist[key[r]] + sCnt[warp * BIN_COUNT + key[r]] + rank[r]] = tb + t0 + warp * (TILE_ROUNDS * 32) + r * 32 + lane;
        }
        __syncthreads();
        sHist[2 * threadIdx.x] += tot & 0xFFFFu; sHist[2 * threadIdx.x + 1] += tot >> 16;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: emit (camera.go:1646-1820)
// ------------------------------------------------------------------------------------------------
// A block takes EMIT_TRIS consecutive sorted triangles of one tile: the i-th surviving triangle of a tile
// is sorted[tile.start + i] and its place in the lists is tileOut[tile] + i.  Phase A (one thread per
// triangle) resolves draw, mesh triangle and the three corner indices into shared memory; phase B (one
// thread per corner, three rounds) builds the 48-byte ebiten.Vertex records of the colour / depth /
// normal lists in shared memory, and one elected thread hands each finished 12 KB run to the TMA
// (cp.async.bulk shared -> global), so every list is written as whole contiguous lines.
static const int EMIT_TRIS = 256;
static const int EMIT_PER_TILE = TILE_TRIS / EMIT_TRIS;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One corner of the emit loop (camera.go:1646-1814): everything the colour / depth records of a vertex hold, from the draw's
// record and the per-vertex arenas.  It depends on (draw, vertex) only -- which is what lets a shard ship it once per arena
// vertex (k_vemit) instead of once per emitted corner.
__device__ __forceinline__ void emit_record(const DrawRec& R, const FrameConst& fc, const uint32_t slot, const uint32_t vi, const bool persp, const bool pcuv,
                                            const float4* __restrict__ clip, const float2* __restrict__ clipXY, const float4* __restrict__ light,
                                            const float* __restrict__ unbill, float4& a, float4& col, float4& dcol, float& custom0) {
    const float4 Tc = __ldg(&clip[slot]);
    float w = Tc.w;                                             // camera.go:1655-1666
    if (!persp) w = 1.0f;
    const bool behind = w < 0;
    if (behind) w = 0.001f;
    // X/w and Y/w are k_vertex's unless the emit loop's w differs from the cull's (w < 0, orthographic);
    // Y / -w == -(Y / w) exactly
    float qx = Tc.x, qy = -Tc.y;
    if (behind || !persp) {
        const float2 xy = __ldg(&clipXY[slot]);
        qx = xy.x / w; qy = xy.y / -w;
    }
    const float dx = qx * fc.camW + fc.halfW;                   // camera.go:1674-1675
    const float dy = qy * fc.camH + fc.halfH;

    const float2 uv = __ldg(reinterpret_cast<const float2*>(R.uv) + vi);
    float uvU, uvV; custom0 = 0.0f;
    if (pcuv) {                                                 // camera.go:1689-1705
        uvU = (uv.x / w) * R.srcW;
        uvV = ((1.0f - uv.y) / w) * R.srcH;
        custom0 = 1.0f / w;
    } else {
        uvU = uv.x * R.srcW;
        uvV = (1.0f - uv.y) * R.srcH;
    }

    col = ldg4(R.color);                                        // camera.go:1726-1736
    if (R.col) {
        const float4 vc = ldg4(R.col + 4 * (size_t)vi);
        col.x = vc.x * col.x; col.y = vc.y * col.y; col.z = vc.z * col.z; col.w = vc.w * col.w;
    }
    const uint32_t flags = R.flags;
    if (flags & T3D_DRAW_LIT) {                                 // camera.go:1738-1742
        const float4 L = __ldg(&light[slot]);
        col.x *= L.x; col.y *= L.y; col.z *= L.z;
    }

    dcol = make_float4(0, 0, 0, 0);
    if (fc.camFlags & T3D_CAM_RENDER_DEPTH) {                   // camera.go:1744-1782
        float vertexDepth = Tc.z;
        if ((flags & T3D_DRAW_HAS_MATERIAL) && (flags & T3D_DRAW_DEPTH_UNBILLBOARDED)) vertexDepth = __ldg(&unbill[slot]);
        float depth = (vertexDepth + fc.depthMargin) / fc.camSpread;
        if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
        dcol = make_float4(depth, depth, depth, 1.0f);
    } else if (fc.worldPresent && fc.fogOn) {                   // camera.go:1784-1814
        float depth = (Tc.z + fc.depthMargin) / fc.camSpread;
        if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
        depth = fc.fogRange[0] + ((fc.fogRange[1] - fc.fogRange[0]) * 1.0f - depth);
        if (fc.fogMode == T3D_FOG_ADD) {
            col.x += fc.fogColor[0] * depth; col.y += fc.fogColor[1] * depth; col.z += fc.fogColor[2] * depth;
        } else if (fc.fogMode == T3D_FOG_SUB) {
            col.x *= fc.fogColor[0] * depth; col.y *= fc.fogColor[1] * depth; col.z *= fc.fogColor[2] * depth;
        }
    }
    a = make_float4(dx, dy, uvU, uvV);
}

// ---- indexed delivery of a shard's frame ------------------------------------------------------------------------------
// The compact form (k_emit<.., COMPACT>) ships 36 bytes per EMITTED corner.  A corner's record depends on its (draw, vertex)
// only, and a mesh vertex is a corner of several triangles, so a shard can instead ship the record once per ARENA vertex plus a
// 4-byte arena slot per emitted corner: 36 V + 4 C bytes instead of 36 C (1 M-triangle frame: V = 550 550, C = 1 374 564:
// 25.3 MB instead of 49.5 MB over NVLink).  It only pays when 36 V + 4 C < 36 C, which depends on how many triangles survive,
// so every kernel of the delivery takes the decision itself from the arena size and the survivor total the tile scan left.
__device__ __forceinline__ bool delivery_indexed(uint32_t totalV, const uint32_t* __restrict__ chunkSum, uint32_t nChunks) {
    uint64_t tris = 0;
    for (uint32_t c = 0; c < nChunks; c++) tris += __ldg(&chunkSum[c]);
    const uint64_t corners = 3 * tris;
    return 36ull * totalV + 4ull * corners < 36ull * corners;
}

// One record per arena vertex, straight into the issuing rank's inbox (coalesced 32-byte stores over NVLink).
__global__ void __launch_bounds__(256)
k_vemit(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ vbase, uint32_t nDraws, uint32_t totalV, const uint32_t* __restrict__ chunkSum, uint32_t nChunks,
        const __grid_constant__ FrameConst fc, const float4* __restrict__ clip, const float2* __restrict__ clipXY, const float4* __restrict__ light,
        const float* __restrict__ unbill, float4* __restrict__ rec32, float* __restrict__ dscal, float* __restrict__ cust) {
    pdl_wait();
    pdl_launch();
    if (!delivery_indexed(totalV, chunkSum, nChunks)) return;
    const uint32_t slot = blockIdx.x * 256u + threadIdx.x;
    if (slot >= totalV) return;
    const uint32_t d = find_draw(vbase, 0, nDraws - 1, slot);
    const DrawRec& R = draws[d];
    if (slot >= R.vbase + R.nv) return;        // (cannot happen: the arena is dense)
    const bool persp = fc.camFlags & T3D_CAM_PERSPECTIVE, pcuv = fc.camFlags & T3D_CAM_PERSP_CORRECT_UV;
    float4 a, col, dcol; float custom0;
    emit_record(R, fc, slot, R.a0 + (slot - R.vbase), persp, pcuv, clip, clipXY, light, unbill, a, col, dcol, custom0);
    rec32[2 * (size_t)slot] = a; rec32[2 * (size_t)slot + 1] = col;
    dscal[slot] = dcol.x;
    if (pcuv) cust[slot] = custom0;
}

// The arena slot of every emitted corner, in list order (what phase A of k_emit resolves), into the inbox.
__global__ void __launch_bounds__(EMIT_TRIS)
k_emit_index(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ tileSurv, const uint32_t* __restrict__ tileOut, const uint32_t* __restrict__ chunkSum, uint32_t nChunks, uint32_t totalV,
             uint32_t* __restrict__ cornerSlot, uint32_t* __restrict__ triIds) {
    pdl_wait();
    pdl_launch();
    if (!delivery_indexed(totalV, chunkSum, nChunks)) return;
    const uint32_t tileId = blockIdx.x / EMIT_PER_TILE;
    const uint32_t k0 = (blockIdx.x % EMIT_PER_TILE) * EMIT_TRIS;
    const uint32_t nSurv = __ldg(&tileSurv[tileId]);
    if (k0 >= nSurv || threadIdx.x >= min((uint32_t)EMIT_TRIS, nSurv - k0)) return;
    const TriTile T = tiles[tileId];
    uint32_t outTri = __ldg(&tileOut[tileId]) + k0 + threadIdx.x;
    for (uint32_t c = 0; c < tileId / SCAN_CHUNK; c++) outTri += __ldg(&chunkSum[c]);
    const uint32_t val = __ldg(&sorted[T.start + k0 + threadIdx.x]);
    uint32_t d, tslot;
    if (T.info & TILE_LARGE) { d = T.dFirst; tslot = val; }
    else { d = T.dFirst + (val >> 20); tslot = T.start + (val & 1023u); }
    const DrawRec& R = draws[d];
    const uint32_t ts = R.tstride, mt = R.t0 + (tslot - R.tbase), vofs = R.vbase - R.a0;
    cornerSlot[3 * (size_t)outTri] = vofs + __ldg(R.idx + mt);
    cornerSlot[3 * (size_t)outTri + 1] = vofs + __ldg(R.idx + ts + mt);
    cornerSlot[3 * (size_t)outTri + 2] = vofs + __ldg(R.idx + 2 * (size_t)ts + mt);
    triIds[outTri] = mt;
}

// COMPACT (sharded frames whose lists live on another GPU): of the 96 bytes a vertex has in the two lists only 36 carry
// information -- {DstX, DstY, SrcX, SrcY | R, G, B, A}, the depth value d of the depth list's colour (d, d, d, 1), and
// Custom0 when texture mapping is perspective-corrected (Custom1-3 are zero, camera.go:1689-1705, 1760-1782).  So the emit
// stores 32-byte records into colorList (reinterpreted), d into depthList (reinterpreted as float[]) and Custom0 into
// custom0List, and the owner of the lists rebuilds both lists with k_gather_expand.
template <bool NORMALS, bool COMPACT>
__global__ void __launch_bounds__(EMIT_TRIS)
k_emit(const DrawRec* __restrict__ draws, const TriTile* __restrict__ tiles, const uint32_t* __restrict__ sorted,
       const uint32_t* __restrict__ tileSurv, const uint32_t* __restrict__ tileOut, const uint32_t* __restrict__ chunkSum,
       const __grid_constant__ FrameConst fc, const float4* __restrict__ clip, const float2* __restrict__ clipXY, const float4* __restrict__ light,
       const float* __restrict__ unbill, const float4* __restrict__ tnrm,
       float4* __restrict__ colorList, float4* __restrict__ depthList, float4* __restrict__ normalList,
       uint32_t* __restrict__ triIds, float* __restrict__ custom0List, uint32_t deliverIndexedTotalV, uint32_t deliverChunks) {
    extern __shared__ uint4 emitSmem[];
    float4* sCol = reinterpret_cast<float4*>(emitSmem);             // [EMIT_TRIS][3]
    float4* sDep = sCol + 3 * EMIT_TRIS;
    float4* sNrm = sDep + 3 * EMIT_TRIS;                            // NORMALS only
    uint32_t* sSlot = reinterpret_cast<uint32_t*>(sDep + (NORMALS ? 6 : 3) * EMIT_TRIS);   // [3 * EMIT_TRIS] arena slot of each corner
    uint32_t* sVi = sSlot + 3 * EMIT_TRIS;                          // mesh vertex index of each corner
    uint32_t* sDraw = sVi + 3 * EMIT_TRIS;                          // [EMIT_TRIS]

    pdl_wait();
    pdl_launch();
    if (COMPACT && deliverIndexedTotalV && delivery_indexed(deliverIndexedTotalV, chunkSum, deliverChunks)) return;      // k_vemit + k_emit_index deliver this frame
    const uint32_t tileId = blockIdx.x / EMIT_PER_TILE;
    const uint32_t k0 = (blockIdx.x % EMIT_PER_TILE) * EMIT_TRIS;
    const uint32_t nSurv = __ldg(&tileSurv[tileId]);
    if (k0 >= nSurv) return;
    const uint32_t nTri = min((uint32_t)EMIT_TRIS, nSurv - k0);
    const TriTile T = tiles[tileId];
    uint32_t outTri0 = __ldg(&tileOut[tileId]) + k0;
    for (uint32_t c = 0; c < tileId / SCAN_CHUNK; c++) outTri0 += __ldg(&chunkSum[c]);     // earlier scan chunks (t3d_sort.cu)

    if (threadIdx.x < nTri) {
        const uint32_t val = __ldg(&sorted[T.start + k0 + threadIdx.x]);
        uint32_t d, tslot;
        if (T.info & TILE_LARGE) { d = T.dFirst; tslot = val; }
        else { d = T.dFirst + (val >> 20); tslot = T.start + (val & 1023u); }
        const DrawRec& R = draws[d];
        const uint32_t ts = R.tstride;
        const uint32_t mt = R.t0 + (tslot - R.tbase);
        const uint32_t vofs = R.vbase - R.a0;
        const uint32_t i0 = __ldg(R.idx + mt), i1 = __ldg(R.idx + ts + mt), i2 = __ldg(R.idx + 2 * (size_t)ts + mt);
        sVi[3 * threadIdx.x] = i0; sVi[3 * threadIdx.x + 1] = i1; sVi[3 * threadIdx.x + 2] = i2;
        sSlot[3 * threadIdx.x] = vofs + i0; sSlot[3 * threadIdx.x + 1] = vofs + i1; sSlot[3 * threadIdx.x + 2] = vofs + i2;
        sDraw[threadIdx.x] = d;
        triIds[outTri0 + threadIdx.x] = mt;
    }
    __syncthreads();

    const uint32_t nVert = 3 * nTri;
    const bool persp = fc.camFlags & T3D_CAM_PERSPECTIVE;
    const bool pcuv = fc.camFlags & T3D_CAM_PERSP_CORRECT_UV;
#pragma unroll 1
    for (uint32_t v0 = 0; v0 < nVert; v0 += EMIT_TRIS) {
        if (v0) {                                                       // the TMA has read the previous round's records
            if (threadIdx.x == 0) bulk_wait_read();
            __syncthreads();
        }
        const uint32_t v = v0 + threadIdx.x;
        if (v < nVert) {
            const uint32_t kk = v / 3u;
            const uint32_t vi = sVi[v], slot = sSlot[v];
            const DrawRec& R = draws[sDraw[kk]];
            float4 a, col, dcol; float custom0;
            emit_record(R, fc, slot, vi, persp, pcuv, clip, clipXY, light, unbill, a, col, dcol, custom0);
            const float4 cust = make_float4(custom0, 0.0f, 0.0f, 0.0f);
            if (COMPACT) {
                float4* C = sCol + 2 * threadIdx.x;
                C[0] = a; C[1] = col;
                reinterpret_cast<float*>(depthList)[3 * (size_t)outTri0 + v] = dcol.x;
                if (pcuv) custom0List[3 * (size_t)outTri0 + v] = custom0;
            } else {
                float4* C = sCol + 3 * threadIdx.x;
                C[0] = a; C[1] = col; C[2] = cust;
                float4* D = sDep + 3 * threadIdx.x;
                D[0] = a; D[1] = dcol; D[2] = cust;
            }
            if (NORMALS) {                                              // camera.go:1710-1722
                const float4 tn = __ldg(&tnrm[slot]);
                float4* Nn = sNrm + 3 * threadIdx.x;
                Nn[0] = a;
                Nn[1] = make_float4(tn.x * 0.5f + 0.5f, tn.y * 0.5f + 0.5f, tn.z * 0.5f + 0.5f, 0.0f);
                Nn[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        fence_async_smem();                                             // generic-proxy writes -> visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t n = min((uint32_t)EMIT_TRIS, nVert - v0);
            const size_t o = 3 * (size_t)outTri0 + v0;                  // first output vertex of this round
            if (COMPACT) bulk_store(colorList + 2 * o, sCol, n * 32u);
            else { bulk_store(colorList + 3 * o, sCol, n * 48u); bulk_store(depthList + 3 * o, sDep, n * 48u); }
            if (NORMALS) bulk_store(normalList + 3 * o, sNrm, n * 48u);
            bulk_commit();
        }
    }
    if (threadIdx.x == 0) bulk_wait_read();                             // shared memory must outlive the TMA reads
}

// Model.BakeLighting frames: corners per arena vertex (see FrameConst.bakeValence).  One thread per arena triangle.
__global__ void __launch_bounds__(256)
k_bake_valence(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ tbase, uint32_t nDraws, uint32_t totalT, uint32_t* __restrict__ valence) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= totalT) return;
    const DrawRec& R = draws[find_draw(tbase, 0, nDraws - 1, t)];
    const uint32_t mt = R.t0 + (t - R.tbase), ts = R.tstride;
    const uint32_t a = R.idx[mt], b = R.idx[ts + mt], c = R.idx[2 * (size_t)ts + mt];
    atomicAdd(&valence[R.vbase + (a - R.a0)], 1u);
    atomicAdd(&valence[R.vbase + (b - R.a0)], 1u);
    atomicAdd(&valence[R.vbase + (c - R.a0)], 1u);
}
void launch_bake_valence(cudaStream_t s, const DrawRec* draws, const uint32_t* tbase, uint32_t nDraws, uint32_t totalT, uint32_t* valence) {
    if (!totalT) return;
    launch_kernel(k_bake_valence, (totalT + 255) / 256, 256, 0, s, false, draws, tbase, nDraws, totalT, valence);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
int debug_vertex_trace(unsigned long long* out, uint32_t nBlocks) {
#ifdef T3D_VERTEX_TRACE
    return cudaMemcpyFromSymbol(out, g_vtrace, sizeof(unsigned long long) * 16 * std::min<uint32_t>(nBlocks, 1024u)) == cudaSuccess ? 0 : -1;
#else
    (void)out; (void)nBlocks; return -1;
#endif
}

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

void launch_prep_draws(cudaStream_t s, bool pdl, const void* hostSlim, const void* hostTmpl, const void* hostRot, const void* hostRest,
                       const t3d_light* hostLights, void* devDraws, void* devRest, uint32_t nDraws, size_t restBytes, uint32_t nLightsIn,
                       const FrameConst& fc, LightU* lightsU, DrawStat* stats, const uint32_t (*win)[4],
                       const void* hostBones, float* devBones, uint32_t nBones, bool bonesPacked, bool slimPacked,
                       uint32_t* devVbase, uint32_t* devTbase, const void* hostRuns, uint32_t runPayload16, const uint32_t* hostBlockRun, uint32_t tmplSpanMax,
                       const PoseBatch* poses, uint32_t nPoseBatches, uint32_t blockBase, uint32_t nBlocksThis, bool firstChunk, uint32_t* tileCtr) {
    if (!nDraws) return;
    cudaFuncSetAttribute(k_prep_draws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PREP_SMEM_MAX);     // per device; cheap
    cudaFuncSetAttribute(k_prep_draws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PREP_SMEM_MAX);
    PrepFrame pf;
    pf.hostSlim = reinterpret_cast<const uint4*>(hostSlim); pf.hostTmpl = reinterpret_cast<const uint4*>(hostTmpl);
    pf.hostRot = reinterpret_cast<const uint4*>(hostRot); pf.hostRest = reinterpret_cast<const uint4*>(hostRest);
    pf.hostLights = hostLights;
    pf.devDraws = reinterpret_cast<uint4*>(devDraws); pf.devRest = reinterpret_cast<uint4*>(devRest);
    pf.nDraws = nDraws; pf.rest16 = firstChunk ? (uint32_t)(restBytes / 16) : 0u; pf.nPrepBlocks = nBlocksThis; pf.nLightsIn = nLightsIn;
    pf.blockBase = blockBase; pf.tileCtr = firstChunk ? tileCtr : nullptr;
    if (!firstChunk) { nBones = 0; nPoseBatches = 0; }       // the copy and pose blocks ride with the first chunk only
    memcpy(pf.win, win, sizeof pf.win);
    pf.hostBones = reinterpret_cast<const uint4*>(hostBones); pf.devBones = reinterpret_cast<uint4*>(devBones);
    pf.nBones = nBones; pf.bonesPacked = bonesPacked ? 1u : 0u;
    pf.slim16 = slimPacked ? SLIM16 - 1 : SLIM16;
    pf.devVbase = devVbase; pf.devTbase = devTbase;
    pf.hostRuns = reinterpret_cast<const uint4*>(hostRuns); pf.hostBlockRun = hostBlockRun; pf.runMode = hostRuns ? 1u : 0u;
    pf.pay16 = hostRuns ? runPayload16 : 0u;
    pf.tmplCap = std::min<uint32_t>(std::max(tmplSpanMax, 1u), PREP_TEMPLATES);
    pf.lightCap = nLightsIn <= PREP_SMEM_LIGHTS ? nLightsIn : 0u;
#ifdef T3D_PREP_SMEM_STATIC
    size_t smem = PREP_SMEM_MAX;
#else
    size_t smem = PREP_SMEM_FIXED + (size_t)pf.tmplCap * sizeof(DrawRec) + (size_t)pf.lightCap * sizeof(t3d_light);
#endif
    const uint32_t nCopy = std::max(pf.rest16 ? std::min<uint32_t>(cdiv(pf.rest16, PREP_THREADS * 8), 1184u) : 0u,
                                    nBones ? std::min<uint32_t>(cdiv(nBones, PREP_THREADS * 2), 1184u) : 0u);
    // the first kernel of a frame is never a dependent launch: it overwrites what the previous frame's kernels read
    (void)pdl;
    // pose batches ride as extra blocks at the START of the grid (the caller only passes batches whose shared memory fits);
    // the kernel's other blocks then count from the first block behind them
    uint32_t nBlocks = pf.nPrepBlocks + nCopy;
    pf.copyEnd = nBlocks;
    if (nPoseBatches) {
        PrepPoses<true> pp;
        pp.nPoseBatches = std::min<uint32_t>(nPoseBatches, PREP_POSE_BATCHES);
        uint32_t nPose = 0;
        for (uint32_t i = 0; i < pp.nPoseBatches; i++) {
            pp.pose[i] = poses[i];
            pp.poseBlock[i] = nPose;
            nPose += cdiv(poses[i].nPoses, pose_armatures_per_block(poses[i].nBones));
            smem = std::max(smem, pose_smem_bytes(poses[i].nBones));
        }
        pp.poseBlock[pp.nPoseBatches] = nPose;
        launch_kernel(k_prep_draws<true>, nBlocks + nPose, PREP_THREADS, smem, s, false, pf, fc, lightsU, stats, pp);
    } else {
        PrepPoses<false> pp; pp.nPoseBatches = 0;
        launch_kernel(k_prep_draws<false>, nBlocks, PREP_THREADS, smem, s, false, pf, fc, lightsU, stats, pp);
    }
}

// One wave of persistent blocks per class list: min(tiles, SMs x resident blocks) blocks pulling tiles off tileCtr[MODE].
static int g_smCount = 0;
template <int MODE, bool APPROX>
static void launch_vertex_class(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* vbase, const uint4* tileList, uint32_t nTiles, uint32_t* tileCtr,
                                uint32_t totalV, uint32_t second, const FrameConst& fc, const LightU* lightsU, const float* bones,
                                const VolumeU* vols, float4* clip, float2* clipXY, float4* light, float4* wpos, float* unbill, float4* tnrm) {
    if (!g_smCount) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_smCount, cudaDevAttrMultiProcessorCount, dev); if (g_smCount <= 0) g_smCount = 148; }
    constexpr uint32_t resident = MODE == K1_STATIC ? T3D_K1_FAST_MINBLOCKS : MODE == K1_SKINNED ? T3D_K1_SKINNED_MINBLOCKS : T3D_K1_MINBLOCKS;
    const uint32_t grid = std::min(nTiles, (uint32_t)g_smCount * resident);     // every block resident at once
    launch_kernel(k_vertex<MODE, APPROX>, grid, VertexShape<MODE>::THREADS, 0, s, pdl, draws, vbase, tileList, nTiles, tileCtr + MODE, totalV, second, fc,
                  lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
}

template <bool APPROX>
static void launch_vertex_t(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* vbase,
                   const uint4* const tileLists[3], const uint32_t nTilesOf[3], uint32_t* tileCtr,
                   uint32_t totalV, const FrameConst& fc, const LightU* lightsU, const float* bones,
                   const VolumeU* vols, float4* clip, float2* clipXY, float4* light, float4* wpos, float* unbill, float4* tnrm) {
    // the launches write disjoint slots; every one after the first overlaps its predecessor and waits for it at its end
    // (`second`; needs the dependent-launch chain: with stage events between kernels they all wait at the top)
    bool first = true;
    if (nTilesOf[K1_STATIC]) {
        launch_vertex_class<K1_STATIC, APPROX>(s, pdl, draws, vbase, tileLists[K1_STATIC], nTilesOf[K1_STATIC], tileCtr, totalV, 0u, fc,
                                               lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
        first = false;
    }
    if (nTilesOf[K1_SKINNED]) {
        launch_vertex_class<K1_SKINNED, APPROX>(s, pdl, draws, vbase, tileLists[K1_SKINNED], nTilesOf[K1_SKINNED], tileCtr, totalV,
                                                (!first && pdl) ? 1u : 0u, fc, lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
        first = false;
    }
    if (nTilesOf[K1_GENERIC])
        launch_vertex_class<K1_GENERIC, APPROX>(s, pdl, draws, vbase, tileLists[K1_GENERIC], nTilesOf[K1_GENERIC], tileCtr, totalV,
                                                (!first && pdl) ? 1u : 0u, fc, lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
}

void launch_vertex(cudaStream_t s, bool pdl, bool exactLighting, const DrawRec* draws, const uint32_t* vbase,
                   const uint4* const tileLists[3], const uint32_t nTilesOf[3], uint32_t* tileCtr,
                   uint32_t totalV, const FrameConst& fc, const LightU* lightsU, const float* bones,
                   const VolumeU* vols, float4* clip, float2* clipXY, float4* light, float4* wpos, float* unbill, float4* tnrm) {
    if (exactLighting) launch_vertex_t<false>(s, pdl, draws, vbase, tileLists, nTilesOf, tileCtr, totalV, fc, lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
    else launch_vertex_t<true>(s, pdl, draws, vbase, tileLists, nTilesOf, tileCtr, totalV, fc, lightsU, bones, vols, clip, clipXY, light, wpos, unbill, tnrm);
}

void launch_cull_large(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* largeList, uint32_t nLarge,
                       const FrameConst& fc, const float4* clip, const float4* wpos, uint32_t* tdepth, DrawStat* stats, bool subdiv) {
    if (!nLarge) return;
    if (subdiv) launch_kernel(k_cull_large<true>, nLarge, TILE_THREADS, 0, s, pdl, draws, tiles, largeList, fc, clip, wpos, tdepth, stats);
    else launch_kernel(k_cull_large<false>, nLarge, TILE_THREADS, 0, s, pdl, draws, tiles, largeList, fc, clip, wpos, tdepth, stats);
}

// One wave of persistent blocks (plus the scan block when scanTiles != 0); ctr = {next ticket, tiles past their cull}, zeroed by k_prep_draws.
void launch_cull_sort(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* fusedList, uint32_t nFused, uint32_t* ctr,
                      const FrameConst& fc, const float4* clip, const float4* wpos, DrawStat* stats, uint32_t* hist, uint32_t* sorted, bool subdiv,
                      uint32_t scanTiles, uint32_t* tileSurv, uint32_t* tileOut, uint32_t* chunkSum) {
    if (!nFused) return;
    if (!g_smCount) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_smCount, cudaDevAttrMultiProcessorCount, dev); if (g_smCount <= 0) g_smCount = 148; }
    const uint32_t grid = std::min(nFused, (uint32_t)g_smCount * (uint32_t)T3D_FUSED_MINBLOCKS) + (scanTiles ? 1u : 0u);
    if (subdiv) launch_kernel(k_cull_sort<true>, grid, TILE_THREADS, 0, s, pdl, draws, tiles, fusedList, nFused, ctr, fc, clip, wpos, stats, hist, sorted, scanTiles, tileSurv, tileOut, chunkSum);
    else launch_kernel(k_cull_sort<false>, grid, TILE_THREADS, 0, s, pdl, draws, tiles, fusedList, nFused, ctr, fc, clip, wpos, stats, hist, sorted, scanTiles, tileSurv, tileOut, chunkSum);
}

void launch_tile_packed(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* tbase, const TriTile* tiles,
                        const uint32_t* packedList, uint32_t nPacked, const FrameConst& fc, const float4* clip,
                        const float4* wpos, uint32_t* sorted, uint32_t* tileSurv, DrawStat* stats, bool subdiv) {
    if (!nPacked) return;
    if (subdiv) launch_kernel(k_tile_packed<true>, nPacked, TILE_THREADS, 0, s, pdl, draws, tbase, tiles, packedList, fc, clip, wpos, sorted, tileSurv, stats);
    else launch_kernel(k_tile_packed<false>, nPacked, TILE_THREADS, 0, s, pdl, draws, tbase, tiles, packedList, fc, clip, wpos, sorted, tileSurv, stats);
}

void launch_draw_medium(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, const uint32_t* mediumList, uint32_t nMedium,
                        const FrameConst& fc, const float4* clip, const float4* wpos, uint32_t* sorted, DrawStat* stats, bool subdiv) {
    if (!nMedium) return;
    if (subdiv) launch_kernel(k_draw_medium<true>, nMedium, TILE_THREADS, 0, s, pdl, draws, tiles, mediumList, fc, clip, wpos, sorted, stats);
    else launch_kernel(k_draw_medium<false>, nMedium, TILE_THREADS, 0, s, pdl, draws, tiles, mediumList, fc, clip, wpos, sorted, stats);
}

void launch_emit(cudaStream_t s, bool pdl, const DrawRec* draws, const TriTile* tiles, uint32_t nTiles, const uint32_t* sorted,
                 const uint32_t* tileSurv, const uint32_t* tileOut, const uint32_t* chunkSum, const FrameConst& fc, const float4* clip, const float2* clipXY,
                 const float4* light, const float* unbill, const float4* tnrm, float4* colorList,
                 float4* depthList, float4* normalList, uint32_t* triIds, bool compactDepth, float* custom0List,
                 uint32_t deliverIndexedTotalV, uint32_t deliverChunks) {
    if (!nTiles) return;
    const size_t idx = (size_t)(3 + 3 + 1) * EMIT_TRIS * sizeof(uint32_t);
    if (compactDepth)
        launch_kernel(k_emit<false, true>, nTiles * EMIT_PER_TILE, EMIT_TRIS, 6 * EMIT_TRIS * sizeof(float4) + idx, s, pdl,
                      draws, tiles, sorted, tileSurv, tileOut, chunkSum, fc, clip, clipXY, light, unbill, tnrm, colorList, depthList, normalList, triIds, custom0List,
                      deliverIndexedTotalV, deliverChunks);
    else if (normalList)
        launch_kernel(k_emit<true, false>, nTiles * EMIT_PER_TILE, EMIT_TRIS, 9 * EMIT_TRIS * sizeof(float4) + idx, s, pdl,
                      draws, tiles, sorted, tileSurv, tileOut, chunkSum, fc, clip, clipXY, light, unbill, tnrm, colorList, depthList, normalList, triIds, custom0List, 0u, 0u);
    else
        launch_kernel(k_emit<false, false>, nTiles * EMIT_PER_TILE, EMIT_TRIS, 6 * EMIT_TRIS * sizeof(float4) + idx, s, pdl,
                      draws, tiles, sorted, tileSurv, tileOut, chunkSum, fc, clip, clipXY, light, unbill, tnrm, colorList, depthList, normalList, triIds, custom0List, 0u, 0u);
}

// The indexed form of a shard's delivery (both kernels return at once when the frame is better off with the compact form).
void launch_emit_indexed(cudaStream_t s, bool pdl, const DrawRec* draws, const uint32_t* vbase, uint32_t nDraws, uint32_t totalV, const TriTile* tiles, uint32_t nTiles,
                         const uint32_t* sorted, const uint32_t* tileSurv, const uint32_t* tileOut, const uint32_t* chunkSum, uint32_t nChunks, const FrameConst& fc,
                         const float4* clip, const float2* clipXY, const float4* light, const float* unbill,
                         float4* rec32, float* dscal, float* cust, uint32_t* cornerSlot, uint32_t* triIds) {
    if (!nTiles || !totalV) return;
    launch_kernel(k_vemit, (totalV + 255) / 256, 256, 0, s, pdl, draws, vbase, nDraws, totalV, chunkSum, nChunks, fc, clip, clipXY, light, unbill, rec32, dscal, cust);
    launch_kernel(k_emit_index, nTiles * EMIT_PER_TILE, EMIT_TRIS, 0, s, pdl, draws, tiles, sorted, tileSurv, tileOut, chunkSum, nChunks, totalV, cornerSlot, triIds);
}

// ---- delivery of sharded frames to the rank that issues the draws -----------------------------------------
// Mailbox words live in the issuing rank's memory and are written / polled across NVLink: a peer's stores land in the owner's
// L2, volatile loads bypass L1.  Serial numbers wrap: compare as signed differences.
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint64_t global_timer_ns() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
static const uint64_t GATHER_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;       // a peer that never answers must not hang the GPU

// Spins until *word has reached `serial`; returns false on timeout.
__device__ __forceinline__ bool wait_serial(const uint32_t* word, uint32_t serial) {
    const uint64_t t0 = global_timer_ns();
    while ((int32_t)(ld_volatile_u32(word) - serial) < 0) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > GATHER_TIMEOUT_NS) return false;
    }
    return true;
}

// A shard's emit may overwrite its inbox on the issuing rank only after that rank has expanded the previous frame.
// (The delivery kernels are links of the frame's dependent-launch chain like every other kernel: wait, then release.)
__global__ void k_gather_wait_ack(const uint32_t* __restrict__ ackWord, uint32_t serial, uint32_t* __restrict__ errWord) {
    pdl_wait();
    pdl_launch();
    if (!wait_serial(ackWord, serial)) *errWord = 1u;
}

// After the shard's emit (kernel boundary: its stores have been performed): the per-draw statistics, the frame's vertex count
// and its serial go to the issuing rank -- statistics and count first, then a system-wide fence, then the serial the owner polls.
__global__ void __launch_bounds__(256)
k_gather_post(uint32_t* __restrict__ mail, const uint32_t* __restrict__ chunkSum, uint32_t nChunks, uint32_t serial,
              const uint4* __restrict__ stats, uint4* __restrict__ statsOut, uint32_t nDraws, uint32_t totalV) {
    pdl_wait();         // the emit's stores into the inbox have been performed once it has completed
    for (uint32_t i = threadIdx.x; i < nDraws; i += 256) statsOut[i] = stats[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t n = 0;
        for (uint32_t c = 0; c < nChunks; c++) n += chunkSum[c];
        mail[1] = 3u * n;
        mail[2] = (totalV && delivery_indexed(totalV, chunkSum, nChunks)) ? 1u : 0u;       // the form this frame arrived in
        __threadfence_system();
        mail[0] = serial;
    }
}

// The issuing rank: waits for the shards' frames and rebuilds their parts of colorVertexList / depthVertexList (tetra3d.go:21-23)
// from the compact or indexed form, and acknowledges (the last block of a shard's group to finish).  ONE launch covers every
// shard: the grid is cut into one group of blocks per shard, each group waits on its own shard's mailbox, so a shard that has
// arrived is rebuilt while the others are still in flight, and the groups' load latencies overlap (as one launch per shard
// the rebuilds ran one after the other: 0.36 ms for seven 1.37 M-vertex shards).  One thread per 16-byte piece
// of the 48-byte output records, four pieces in flight per thread, so that a warp's store covers 512 contiguous bytes of a list.
struct GatherShard {
    const float4* rec32; const float* dscal; const float* cust; const uint32_t* cornerSlot;
    float4* color; float4* depth;
    uint32_t* mail;     // {serial, nVerts, indexed} | +16: ack | +17: done counter
};
static const uint32_t GATHER_MAX_SHARDS = 15;
struct GatherShards { GatherShard s[GATHER_MAX_SHARDS]; };
__global__ void __launch_bounds__(256, 4)
k_gather_expand(const __grid_constant__ GatherShards shards, uint32_t blocksPerShard, uint32_t serial, int renderDepth, int pcuv, uint32_t* __restrict__ errWord) {
    __shared__ uint32_t sN, sIndexed;
    pdl_launch();       // (nothing follows in the frame; the wait for the own emit comes last: the rebuild touches other ranks' regions only)
    const GatherShard& S = shards.s[blockIdx.x / blocksPerShard];
    const uint32_t blk = blockIdx.x % blocksPerShard;
    if (threadIdx.x == 0) {
        if (!wait_serial(S.mail, serial)) { *errWord = 1u; sN = 0; sIndexed = 0; }
        else { sN = ld_volatile_u32(S.mail + 1); sIndexed = ld_volatile_u32(S.mail + 2); }
    }
    __syncthreads();
    const uint64_t n3 = 3ull * sN;
    const bool indexed = sIndexed != 0;
    const uint64_t stride = (uint64_t)blocksPerShard * 256;
    constexpr int U = 4;
    for (uint64_t i0 = (uint64_t)blk * 256 + threadIdx.x; i0 < n3; i0 += U * stride) {
        uint32_t v[U], part[U];
#pragma unroll
        for (int u = 0; u < U; u++) {       // indexed frames: the record of an output vertex is the one of the arena vertex its corner refers to
            const uint64_t i = i0 + u * stride;
            const uint32_t ov = (uint32_t)(i / 3);
            part[u] = (uint32_t)(i - 3ull * ov);
            v[u] = (i < n3 && indexed) ? __ldcg(S.cornerSlot + ov) : ov;
        }
        float4 c[U], d[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (i0 + u * stride >= n3) continue;
            if (part[u] == 0) c[u] = d[u] = __ldcg(S.rec32 + 2 * (size_t)v[u]);                   // DstX DstY SrcX SrcY
            else if (part[u] == 1) {
                c[u] = __ldcg(S.rec32 + 2 * (size_t)v[u] + 1);                                    // R G B A
                const float z = __ldcg(S.dscal + v[u]);
                d[u] = renderDepth ? make_float4(z, z, z, 1.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            } else c[u] = d[u] = make_float4(pcuv ? __ldcg(S.cust + v[u]) : 0.0f, 0.0f, 0.0f, 0.0f);    // Custom0..3
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = i0 + u * stride;
            if (i < n3) { __stcs(S.color + i, c[u]); __stcs(S.depth + i, d[u]); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(S.mail + 17, 1u) == blocksPerShard - 1) {
            S.mail[17] = 0;
            __threadfence_system();
            S.mail[16] = serial;          // the shard may overwrite its inbox now
        }
    }
    pdl_wait();         // this kernel complete => the frame's own emit complete (what the frame's last event stands for)
}

void launch_gather_wait_ack(cudaStream_t s, bool pdl, const uint32_t* ackWord, uint32_t serial, uint32_t* errWord) {
    launch_kernel(k_gather_wait_ack, 1, 1, 0, s, pdl, ackWord, serial, errWord);
}
void launch_gather_post(cudaStream_t s, bool pdl, uint32_t* mail, const uint32_t* chunkSum, uint32_t nChunks, uint32_t serial,
                        const void* stats, void* statsOut, uint32_t nDraws, uint32_t totalV) {
    launch_kernel(k_gather_post, 1, 256, 0, s, pdl, mail, chunkSum, nChunks, serial, reinterpret_cast<const uint4*>(stats), reinterpret_cast<uint4*>(statsOut), nDraws, totalV);
}
void launch_gather_expand(cudaStream_t s, bool pdl, uint32_t nShards, const void* const* rec32, const float* const* dscal, const float* const* cust, const uint32_t* const* cornerSlot,
                          t3d_vertex* const* color, t3d_vertex* const* depth, uint32_t* const* mail, uint32_t serial, bool renderDepth, bool pcuv, uint32_t* errWord) {
    if (!g_smCount) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_smCount, cudaDevAttrMultiProcessorCount, dev); if (g_smCount <= 0) g_smCount = 148; }
    for (uint32_t first = 0; first < nShards; first += GATHER_MAX_SHARDS) {
        const uint32_t n = std::min(nShards - first, GATHER_MAX_SHARDS);
        GatherShards G; memset(&G, 0, sizeof G);
        for (uint32_t i = 0; i < n; i++) {
            GatherShard& S = G.s[i];
            S.rec32 = reinterpret_cast<const float4*>(rec32[first + i]); S.dscal = dscal[first + i]; S.cust = cust[first + i]; S.cornerSlot = cornerSlot[first + i];
            S.color = reinterpret_cast<float4*>(color[first + i]); S.depth = reinterpret_cast<float4*>(depth[first + i]); S.mail = mail[first + i];
        }
        // every group must be resident at once (a group whose shard is late must not keep an earlier group's blocks from starting)
        const uint32_t per = std::max(1u, (uint32_t)g_smCount * 4u / n);       // (4 blocks per SM: the kernel's launch bounds)
        launch_kernel(k_gather_expand, n * per, 256, 0, s, pdl, G, per, serial, renderDepth ? 1 : 0, pcuv ? 1 : 0, errWord);
    }
}

}  // namespace t3d
