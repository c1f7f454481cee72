// t3d_device.cuh -- device-side records and float32 math for the geometry stage.
//
// Everything that feeds a cull decision, a depth key or a bin index is computed in float32 with
// separate multiply and add (the translation units are built with -fmad=false and IEEE div/sqrt),
// in the operand order of the reference, so the draw lists come out bit-identical to the Go path:
// matrix_amd64.s:4-81 uses MULPS+ADDPS, Go/amd64 (GOAMD64=v1) never fuses.
#pragma once

#include <cuda_runtime.h>
#include "t3d_math.h"
#include "../../include/t3dgeo.h"
#include <stddef.h>

namespace t3d {

// ---- records ----------------------------------------------------------------------------------

// Frame constants (camera.go:1329-1341, 1603; world fog camera.go:1784-1812).
struct FrameConst {
    float vp[16];
    float view[16];
    float camPos[3];
    float camUp[3];
    float nearP, farP;
    float camW, camH, halfW, halfH;
    float depthMargin;      // (far - near) * DepthMargin          camera.go:1339
    float camSpread;        // far - near + 2*margin               camera.go:1341
    float snapping;
    uint32_t camFlags;      // T3D_CAM_*
    int32_t worldPresent, fogOn, fogMode;
    float fogColor[3];
    float fogRange[2];
    // scene.library.autosubdivisionLevels (library.go:206-209): {DistanceSquared, MaximumTriangleSize}
    uint32_t nSubLevels;
    float subLevels[8][2];
    // 1: default lighting (approximate reciprocals, colours within 1/255): k_prep_draws folds the per-light constants the fast
    // light loops read and groups every draw's list by light type; 0: t3d_set_exact_lighting (lists stay in submit order)
    uint32_t approxLighting;
    // Model.BakeLighting frames (t3d_frame_set_bake_lighting): for every arena vertex, how many triangle corners of its draw
    // refer to it -- the reference's bake adds a light's contribution once per corner (model.go:1246-1262) -- else nullptr
    const uint32_t* bakeValence;
    // Camera.SphereInFrustum's terms (camera.go:933-983), for instances submitted with culling bounds (t3d_submit_instances)
    float camForward[3], camRight[3];
    float sphereFactorY, sphereFactorX, sphereFactorTang, orthoScale;
};

// Resolved light as seen by one draw: light.beginRender/beginModel products (light.go:110-112,
// 175-189, 322-330) plus the per-draw range skip of camera.go:1625-1629.
//
// Default (approximate) lighting re-purposes fields a light type does not otherwise read, for the list's own lights:
//   point        range  = the squared distance beyond which the light is skipped: lightRangeSquared when range > 0, else the
//                         largest d with !(1 / d * energy < 0.001) (light.go:205-210; found exactly by the host, t3d_submit_lights)
//                range2 = lightRangeSquared when > 0, else +inf (the range falloff then multiplies by 1)
//                fwd    = colour * energy * 2
//   directional  wpos   = workingModelRotation rows . forward, range = its constant term (light.go:342-349 folded per draw)
//   ambient      fwd    = colour * energy
struct LightU {
    uint32_t type;          // T3D_LIGHT_*, or LIGHT_SKIP
    float color[3];
    float wpos[3];          // PointLight.workingPosition
    float energy;
    float range, range2;    // lightRange, lightRangeSquared
    float fwd[3];           // DirectionalLight.workingForward
    uint32_t volume;        // index into the context's volume table
    uint32_t useFirst;      // volume lights: where the LightU records of the lights its cells use start, relative to the draw's first
    uint32_t pad;
};
static const uint32_t LIGHT_SKIP = 0xFFu;
// DrawRec.flags, set by k_prep_draws: an instance whose culling sphere is outside the frustum (t3d_submit_instances with bounds).
// Its arena ranges stay (the host laid them out) but nothing of it survives the cull; DrawStat.minBits = DRAW_CULLED_MARK tells the host.
static const uint32_t T3D_DRAW_FRUSTUM_CULLED = 1u << 31;
static const uint32_t DRAW_CULLED_MARK = 0xFFFFFFFFu;

// LightVolume on the device.
struct VolumeU {
    float dimMin[3], dimMax[3];
    float cellSize[3];
    int32_t cellCount[3];
    float offset[3];
    float bias[3];
    const float* cells;
    const uint32_t* cellLightStart;     // LightVolumeCell.useLights as CSR over the cells, or nullptr
    const uint32_t* cellLightIndex;     // indices into the volume's use-light run of the draw's LightU list
};

// Auto-subdivision data of a mesh (t3d_mesh_desc.tri_*), one record per mesh in device memory.
struct SubdivMesh {
    const uint32_t* base;       // per triangle entry: its base triangle
    const int8_t*   level;      // -1 base, else level
    const uint8_t*  nLevels;    // base triangles: number of pre-built levels
    const float*    span;       // Triangle.MaxSpan
    const float*    normal;     // packed float3 Triangle.Normal
};

// One render(rp): host-resolved inputs first, then the fields k_prep_draws derives.
struct __align__(16) DrawRec {
    // ---- resolved by the host at submit ----
    const float*    pos;        // SoA planes x|y|z, plane stride = vstride floats (a multiple of 4 vertices)
    const float*    nrm;        // same layout
    const float*    uv;         // float2
    const float*    col;        // float4 or nullptr
    const float*    wgt;        // float4 or nullptr
    const uint16_t* jnt;        // ushort4 or nullptr
    const uint8_t*  ninf;       // or nullptr
    const uint32_t* idx;        // SoA planes a|b|c (Triangle.VertexIndexA/B/C), plane stride = tstride
    const float*    ctr;        // SoA planes x|y|z (Triangle.Center), plane stride = tstride
    const float*    keyPos;     // k keys x 3 planes x vstride
    const float*    keyNrm;
    const float*    keyW;       // k
    const SubdivMesh* subdiv;   // auto-subdivision data of the mesh, or nullptr
    // written by k_prep_draws under default lighting: the draw's LightU list is grouped [point | directional | ambient | skipped]
    // (lightsSorted = 1) with these counts; 0: the list is in submit order and holds lightCount typed entries
    uint16_t        nPoint, nSun, nAmbient, lightsSorted;
    uint32_t a0;                // first mesh vertex of the aligned range (multiple of 4)
    uint32_t nv;                // aligned vertex count (multiple of 4)
    uint32_t vbase;             // first slot in the per-frame vertex arenas (multiple of 4)
    uint32_t t0;                // first mesh triangle
    uint32_t nt;                // triangle count
    uint32_t tbase;             // first slot in the per-frame triangle arenas
    uint32_t flags;             // T3D_DRAW_*
    uint32_t nKeys;
    uint32_t vstride;           // padded mesh vertex count: plane stride of pos / nrm / shape keys
    uint32_t boneOffset;
    uint32_t lightIn;           // first t3d_light of the draw's list
    uint32_t lightCount;
    uint32_t lightU;            // first LightU slot of this draw
    uint8_t  sortMode, lightingMode, lvMode, pad0;
    uint32_t tstride;           // padded mesh triangle count: plane stride of idx / ctr
    uint32_t lightTotal;        // lightCount + the lights the list's LightVolume cells use (their LightU records follow the list's own)
    float transform[16];
    float worldRot[16];
    float color[4];
    float srcW, srcH;
    float maxSpan;
    uint32_t nBones;            // skinned draws: bones the mesh's joints reach (max joint + 1); k_vertex stages that many matrices
    // ---- derived on the device (model.go:616-708) ----
    float mvp[16];              // base * vp (static); skinned draws use FrameConst.vp
    float unbill[4];            // column 2 of modelTransform * vp       model.go:702, 801
    float nrmRot[12];           // mvJustRForNormals columns 0..2, rows 0..3   model.go:707, 815-817
    float invCam[3];            // invertedCamPos                        model.go:629-630
    float pad2;
};

// What crosses PCIe per draw.  Consecutive draws that differ only in placement (instances of one MeshPart:
// particles, crowds, instanced scenery) share one DrawRec *template*; k_prep_draws rebuilds the full record on
// the device from the template and this 112-byte record (a quarter of the bytes of a DrawRec).
struct __align__(16) DrawSlim {
    uint32_t tmpl;              // template index (one of the 16 most recent templates at submit time, or a new one)
    uint32_t vbase, tbase;      // arena slots
    uint32_t boneOffset;
    uint32_t lightU;            // first LightU slot of this draw
    uint32_t rot;               // index of the draw's world rotation in the rotation array, or NO_ROTATION
    uint32_t pad[2];
    float transform[16];
    float color[4];
};
static const uint32_t NO_ROTATION = 0xFFFFFFFFu;

// Run-length form of the placement records, used for frames dominated by long runs of one template (particles,
// instanced scenery): a run is a sequence of consecutive draws whose DrawSlim headers progress regularly, so the
// header travels once per run and each draw ships only its transform rows and colour (64 bytes).
struct __align__(16) DrawRun {
    uint32_t first, count;      // draws [first, first + count)
    uint32_t tmpl;
    uint32_t vbase, tbase, boneOffset, lightU, rot;     // of the run's first draw
    uint32_t dV, dT, dBone, dLightU, dRot;              // per-draw increments (dRot is 0 or 1; rot stays NO_ROTATION when 0)
    uint32_t pad[3];
};
static_assert(sizeof(DrawRun) == 64, "DrawRun is staged as four uint4");
static const uint32_t RUN_MODE_MIN_AVG = 8;     // average run length from which a frame is shipped in run-length form
static_assert(sizeof(DrawSlim) == 112, "DrawSlim is staged as seven uint4");

static_assert(offsetof(DrawRec, transform) % 16 == 0 && offsetof(DrawRec, worldRot) % 16 == 0 &&
              offsetof(DrawRec, color) % 16 == 0 && offsetof(DrawRec, mvp) % 16 == 0 &&
              offsetof(DrawRec, unbill) % 16 == 0 && offsetof(DrawRec, nrmRot) % 16 == 0 &&
              sizeof(DrawRec) % 16 == 0, "DrawRec arrays are read with 128-bit loads");
static_assert(sizeof(LightU) == 64, "LightU is read as four float4");

// Per-draw reduction state written by the cull kernels.
struct DrawStat {
    uint32_t minBits;           // float bits of minDepth (depth >= 0, so uint order == float order)
    uint32_t maxBits;
    uint32_t survivors;
    uint32_t done;              // tiles of this draw whose bin histogram is complete (k_bin_hist)
};

// Triangle tiles: the unit of work of the cull / sort / emit kernels.  A tile is a run of at most
// TILE_TRIS consecutive slots of the frame's triangle arena that never cuts a small draw: a draw with
// more than TILE_TRIS triangles owns ceil(nt / TILE_TRIS) exclusive tiles ("large"), smaller draws are
// packed whole, several to a tile ("packed").  Draws of up to MEDIUM_TRIS triangles keep their large tiles
// for the tile scan and the emit but are culled and sorted by a single block (k_draw_medium).
struct __align__(16) TriTile {
    uint32_t start;             // first triangle slot
    uint32_t count;             // slots in the tile (<= TILE_TRIS)
    uint32_t dFirst;            // first draw
    uint32_t info;              // packed: number of draws (<= TILE_TRIS); large: TILE_LARGE | tile index in the draw
    // large tiles only: where the draw's bin counts live (k_bin_hist / k_scatter)
    uint32_t row0;              // first histogram row of the draw: rows row0 .. row0+nTilesOfDraw-1 are per-tile
                                // offsets, row row0+nTilesOfDraw holds the draw's bin bases
    uint32_t nTilesOfDraw;
    uint32_t sortMode;          // T3D_SORT_* of the draw
    uint32_t tbase;             // first arena slot of the draw
};
static const uint32_t TILE_TRIS = 1024;
static const uint32_t HUGE_TILES = 64;          // draws of more tiles than this get their histogram scan from k_hist_scan_huge
static const uint32_t SCAN_CHUNK = 4096;        // tiles per block of the tile scan (t3d_sort.cu)
static const uint32_t MEDIUM_TRIS = 8192;       // draws of (TILE_TRIS, MEDIUM_TRIS] triangles are culled and sorted by one block
static const uint32_t TILE_LARGE = 0x80000000u;

static const uint32_t KEY_CULLED = 0xFFFFFFFFu;
static const int BIN_BITS = 9;      // 512 bins, sortingTriangles.go:166-168
static const int BIN_COUNT = 512;

#ifdef __CUDACC__
// Programmatic dependent launch (PDL).  Every kernel of a frame after the first is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its blocks may start while the previous kernel's last
// blocks are still running.  pdl_wait() blocks until the previous kernel has completed and its writes are
// visible (a no-op for a plain launch); pdl_launch() lets the next kernel's blocks start once every block of
// this grid has called it.  Every block calls pdl_launch() only AFTER pdl_wait(), so when kernel N+1's blocks
// start, kernel N-1 and everything before it has completed: code placed before pdl_wait() may read what
// kernels two or more launches back wrote (the frame tables, k_prep_draws' records) and static mesh data, but
// nothing the immediately preceding kernel writes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// 512-bin key of one surviving triangle (sortingTriangles.go:38-64).  BackToFront visits bins 511..0
// (sortingTriangles.go:101), so the key is the visiting rank of the bin: ascending key == draw order.
__device__ __forceinline__ uint32_t bin_key(uint32_t depthBits, uint32_t minBits, uint32_t maxBits, uint32_t sortMode) {
    const float minRange = __uint_as_float(minBits);
    const float maxRange = __uint_as_float(maxBits);
    float rangeDiff = maxRange - minRange;
    if (rangeDiff == 0) rangeDiff += 0.001f;
    uint32_t bin = 0;
    if (sortMode != T3D_SORT_NONE) {
        const float dd = (__uint_as_float(depthBits) - minRange) / rangeDiff * (float)BIN_COUNT;
        const float tt = clampf(dd, 0.0f, (float)(BIN_COUNT - 1));
        bin = (uint32_t)(int)tt;
    }
    return (sortMode == T3D_SORT_BACK_TO_FRONT) ? (BIN_COUNT - 1 - bin) : bin;
}
#endif

// ---- bone poses (t3d_pose.cu) -------------------------------------------------------------------------------
// One t3d_submit_poses call: n armatures of one skeleton playing one animation.  Skeleton and animation tables are
// resident device memory; the pose records are read from the pinned frame block.
static const uint32_t POSE_MAX_BONES = 1024;        // bones per skeleton (the hierarchy walk keeps an armature's matrices in shared memory)
static const uint32_t POSE_TAG = 0x80000000u;       // bone offsets of device-evaluated poses: index among the posed bones | tag
struct PoseBatch {
    uint32_t nBones, maxDepth, animBones, nPoses;
    const int32_t* parent; const uint16_t* depth;       // [nBones]
    const float* invBind; const float* restPos; const float* restScale; const float* restRot;
    const t3d_track* tracks; const float* keyTimes; const float4* keyValues;   // animBones == 0: rest pose
    const t3d_pose* poses;                              // [nPoses] pinned host memory (UVA)
    float* out;                                         // first matrix of pose 0 in the frame's bone arena
};
// t3d_submit_poses_blended: the animation the player blends from (stand-alone k_pose_bones<true> only, so that the pose blocks
// riding in k_prep_draws keep their size)
struct PoseBlend {
    uint32_t fromBones;                                 // 0: the previous state animated nothing
    const t3d_track* tracks; const float* keyTimes; const float4* keyValues;
};
// t3d_submit_pose_chains: the animations played since the last finished blend (stand-alone k_pose_chain only); the batch's pose
// records are t3d_pose_chain (24 floats)
struct PoseChainLevel { uint32_t bones; const t3d_track* tracks; const float* keyTimes; const float4* keyValues; };
struct PoseChain { uint32_t nLevels; PoseChainLevel level[T3D_POSE_CHAIN_MAX]; };

}  // namespace t3d
