// t3d_api.cu -- the C ABI of include/t3dgeo.h: context, mesh residency, per-frame arenas, pipeline.
//
// Per frame: the host resolves every submitted draw into a placement record (DrawSlim) plus, when anything else
// changes, a DrawRec template; cuts the vertex and triangle arenas into tiles (t3d_device.cuh: TriTile) and sorts
// them into per-kernel work lists; packs everything into one pinned, mapped block; then enqueues, chained with
// programmatic dependent launches on one stream,
//   k_prep_draws (pulls the block over PCIe, rebuilds DrawRec[D], per-draw uniforms)
//   -> k_vertex<STATIC | SKINNED | GENERIC>
//   -> k_cull_large | k_tile_packed | k_draw_medium          (by draw size; the last two also sort)
//   -> k_bin_hist (+ tile scan) [-> k_hist_scan_huge] -> k_scatter    (draws of more than 8192 triangles)
//   -> k_emit.
// Results stay on the device until t3d_frame_results (pinned D2H) or are handed out as device
// pointers by t3d_frame_results_device; t3d_set_output_target redirects the emit to caller-provided memory.
#include "../../include/t3dgeo.h"
#include "t3d_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace t3d;

namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        size_t want = std::max(bytes, cap + cap / 2);
        want = (want + 255) & ~(size_t)255;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// growable pinned host array that keeps its contents
template <class T> struct PinVec {
    T* p = nullptr; size_t n = 0, cap = 0;
    bool mapped = false;        // also readable by kernels through UVA (cudaHostGetDevicePointer)
    cudaError_t reserve(size_t want) {
        if (want <= cap) return cudaSuccess;
        size_t nc = std::max(want, cap * 2 + 16);
        T* q = nullptr;
        cudaError_t e = cudaHostAlloc((void**)&q, nc * sizeof(T), mapped ? (cudaHostAllocMapped | cudaHostAllocPortable) : cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (p) { memcpy(q, p, n * sizeof(T)); cudaFreeHost(p); }
        p = q; cap = nc;
        return cudaSuccess;
    }
    cudaError_t resize(size_t want) { cudaError_t e = reserve(want); if (e == cudaSuccess) n = want; return e; }
    cudaError_t push(const T& v) { cudaError_t e = reserve(n + 1); if (e == cudaSuccess) p[n++] = v; return e; }
    void clear() { n = 0; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = cap = 0; }
};

struct PartInfo { uint32_t triStart, triEnd, a0, nv; };

// t3d_skeleton_upload / t3d_animation_upload: one device allocation each
struct Skeleton {
    uint32_t nBones = 0, maxDepth = 0;
    void* block = nullptr;
    int32_t* parent = nullptr; uint16_t* depth = nullptr; float *invBind = nullptr, *restPos = nullptr, *restScale = nullptr, *restRot = nullptr;
};
struct Animation {
    uint32_t nBones = 0, nKeys = 0;
    void* block = nullptr;
    t3d_track* tracks = nullptr; float* keyTimes = nullptr; float4* keyValues = nullptr;
};
struct PoseCall { uint64_t skeleton, animation; uint32_t firstPose, nPoses, firstBone; uint64_t fromAnimation; bool blended;
                  uint32_t nLevels; uint64_t chain[T3D_POSE_CHAIN_MAX]; };      // nLevels != 0: t3d_submit_pose_chains (firstPose indexes hPoseChains)      // one t3d_submit_poses[_blended] call of the frame


struct Mesh {
    uint32_t generation = 0;
    uint32_t nVerts = 0, nVertsPad = 0, nTris = 0, nTrisPad = 0, nKeys = 0;
    uint32_t maxJoint = 0;
    float maxSpan = 0;
    float *pos = nullptr, *nrm = nullptr, *uv = nullptr, *col = nullptr, *wgt = nullptr, *ctr = nullptr;
    float *keyPos = nullptr, *keyNrm = nullptr, *keyW = nullptr;
    uint16_t* jnt = nullptr; uint8_t* ninf = nullptr; uint32_t* idx = nullptr;
    void* subdivBlock = nullptr; SubdivMesh* subdiv = nullptr;      // auto-subdivision arrays + the record DrawRec.subdiv points at
    std::vector<PartInfo> parts;
    void release() {
        void* ptrs[] = {pos, nrm, uv, col, wgt, ctr, keyPos, keyNrm, keyW, jnt, ninf, idx, subdivBlock};
        for (void* q : ptrs) if (q) cudaFree(q);
        *this = Mesh();
    }
};

struct LightList { uint32_t off, count, total; };     // total: count + the lights the list's LightVolume cells use

}  // namespace

struct t3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::unordered_map<uint64_t, Mesh> meshes;
    std::unordered_map<uint32_t, uint32_t> volumeIndex;     // id -> slot in volumes
    std::vector<VolumeU> volumes;
    std::vector<float*> volumeCells;
    // LightVolumeCell.useLights: per volume slot the CSR (one device block: start[cells + 1] | index[]), how many lights the
    // cells refer to, and this frame's lights (t3d_volume_use_lights)
    std::vector<uint32_t*> volumeCsr; std::vector<uint32_t> volumeUseCount; std::vector<std::vector<t3d_light>> volumeUse;
    DevBuf dVolumes; bool volumesDirty = true;
    std::unordered_map<uint64_t, Skeleton> skeletons;
    std::unordered_map<uint64_t, Animation> animations;
    PinVec<t3d_pose_chain> hPoseChains;     // records of t3d_submit_pose_chains: read by k_pose_chain straight from pinned memory
    PinVec<t3d_pose> hPoses; std::vector<PoseCall> poseCalls; uint32_t posedBones = 0;     // this frame's t3d_submit_poses calls
    size_t framePosesOff = 0;

    // frame staging (pinned)
    bool inFrame = false, executed = false, haveResults = false;
    bool stageTimings = false, stagesTimed = false;
    bool bakeFrame = false;             // t3d_frame_set_bake_lighting: this frame stops behind k_vertex and lights like Model.BakeLighting
    DevBuf dValence;
    bool exactLighting = false;         // t3d_set_exact_lighting: the reference's lighting arithmetic operation by operation (bit-exact colours)
    bool forceFullUpload = false;       // t3d_debug_force_full_upload: never pack affine matrices (exercises the 16-float paths)
    FrameConst fc;
    // what the host itself needs to remember about a submitted draw (the device rebuilds the full DrawRec from hSlim + hTmpl)
    struct HostDraw { uint32_t a0, nv, vbase, nt, flags; uint8_t sortMode; uint8_t vclass; /* k_vertex instantiation it is eligible for: 0 generic, 1 static fast, 2 skinned fast */ };
    std::vector<HostDraw> hDraws;
    PinVec<DrawSlim> hSlim;
    std::vector<DrawRun> hRuns;          // run-length form of hSlim's headers, rebuilt every frame
    std::vector<uint32_t> hBlockRun;
    PinVec<DrawRec> hTmpl;
    PinVec<float> hRot;
    PinVec<uint32_t> hVbase, hTbase, hTileV, hLarge, hPacked, hMedium, hHuge;
    PinVec<uint32_t> hFused;        // tiles of the draws of 9 .. HUGE_TILES tiles: culled and sorted by k_cull_sort (hLarge keeps the tiles of larger draws)
    uint32_t* pFused = nullptr;
    // k_cull_sort wins when all its tiles are resident at once (one wave: 300 K-triangle frame of 12-tile draws 26.4 -> 21.0 us for
    // cull + sort); with more tiles than resident blocks, draws that straddle the wave boundary hold their first-wave blocks
    // until second-wave blocks find a slot, and the three-kernel path is as fast or faster (1 M triangles: 41.4 vs 38.5 us staged,
    // 102.1 vs 103.4 us chained).  fuseSort: 0 never, 1 when the tiles fit one wave (default), 2 always (T3D_FUSE_SORT, tests).
    int fuseSort = 1; int smCount = 148; bool frameFused = false;
    bool deliverIndexed = true;        // sharded delivery may take the indexed form (T3D_DELIVER_INDEXED=0: always the compact form; A/B measurements)
    PinVec<uint4> hVClass[3];       // vertex tiles by class (0 generic, 1 static fast, 2 skinned fast): {tile, first draw, first draw of the next tile, end of the first draw's vertices}
    PinVec<TriTile> hTiles;
    PinVec<t3d_light> hLights;
    PinVec<float> hBones;
    std::vector<LightList> lightLists;
    std::vector<uint32_t> drawTotalTris;
    std::vector<uint8_t> drawContinue;
    // t3d_submit_instances: what the instances of one call share.  rot / parent: where the two matrices sit in hRot
    struct TrsGroup { uint32_t rot, parent; Mat4 rotation, parentT; bool hasParent, hasBounds; float bounds[9]; };
    std::vector<TrsGroup> trsGroups;
    uint32_t nTrsDraws = 0; bool trsUniform = true;       // draws submitted as instances; all of them with x == y == z scale
    uint32_t vAcc = 0, tAcc = 0, luAcc = 0, nHistRows = 0;
    bool anyLit = false, anySkinned = false, anyUnbill = false, anySubdiv = false, bonesAffine = true, transformsAffine = true;

    // device arenas
    DevBuf dFrame, dLightsU, dStats;            // dFrame: all per-frame uniforms, packed, uploaded with one copy
    PinVec<uint8_t> hFrame;
    void* hFrameDev = nullptr; void* pDevRest = nullptr;
    size_t frameHeadBytes = 0, frameRestBytes = 0, frameTmplOff = 0, frameRotOff = 0, frameLightsOff = 0, frameBonesOff = 0;
    size_t frameRunsOff = 0, frameBlockRunOff = 0; bool frameRunMode = false;
    uint32_t blkTmin = 0, blkTmax = 0, tmplSpanMax = 1;     // widest template-index window of any 64-draw block (sizes k_prep_draws' staging)
    uint32_t frameBoneCount = 0;
    DrawRec* pDraws = nullptr; uint32_t *pVbase = nullptr, *pTbase = nullptr, *pLarge = nullptr, *pPacked = nullptr, *pMedium = nullptr, *pHuge = nullptr; uint4* pVClass[3] = {nullptr, nullptr, nullptr};
    TriTile* pTiles = nullptr;
    t3d_light* pLightsIn = nullptr; float* pBones = nullptr;
    DevBuf dStage;      // device copy of the pinned frame block, for frames large enough that a DMA copy beats k_prep_draws' own PCIe reads
    size_t dmaThreshold = 1u << 20; bool frameStaged = false; size_t frameBlockBytes = 0, frameSlimBytes = 0;
    uint32_t prepChunkBlocks = 0;       // 0: adaptive (see t3d_frame_execute); T3D_PREP_CHUNK_BLOCKS overrides
    cudaStream_t copyStream = nullptr; cudaEvent_t evFork = nullptr; std::vector<cudaEvent_t> evChunk;     // chunked staging of very large frames
    DevBuf dClip, dClipXY, dLight, dWpos, dUnbill, dTnrm;     // dClip: {X/w, Y/w, Z, W} per vertex; dClipXY: {X, Y}
    DevBuf dTdepth, dSorted, dHist, dTileSurv, dTileOut, dCount;
    DevBuf dTileCtr;        // k_vertex's tile counters (one per vertex-tile class), zeroed by k_prep_draws every frame
    DevBuf dColor, dDepth, dNormal, dTriIds;
    // Two frames in flight (t3d_set_async_download): frame k's lists travel to the host on dlStream while frame k + 1 is built and
    // executed, so frames alternate between two sets of device lists and of pinned host buffers (set 1 below; set 0 are the
    // ordinary ones).  A set's device lists are not written again before its last copy has left them (a stream wait on
    // evCopyDone), its host buffers not before that copy has arrived (t3d_frame_results of frame k + 2 waits for it).
    bool asyncDownload = false;
    bool skipUnusedDepth = false;       // t3d_set_skip_unused_depth_list: no D2H of the depth list when Camera.RenderDepth is off
    uint32_t nextSet = 0, frameSet = 0, lastCopySet = 0;
    bool copyPending[2] = {false, false};
    DevBuf dColorB, dDepthB, dNormalB, dTriIdsB;
    cudaStream_t dlStream = nullptr; cudaEvent_t evFrameDone = nullptr, evCopyDone[2] = {nullptr, nullptr};
    uint32_t* oTriIds = nullptr;
    // external output target (t3d_set_output_target): the emit kernel writes the lists there -- another GPU's
    // memory over NVLink, or an interop buffer of the rasteriser -- instead of into dColor / dDepth / dNormal
    t3d_vertex *xColor = nullptr, *xDepth = nullptr, *xNormal = nullptr;
    // delivery of sharded frames to the rank that issues the draws (t3d_gather_*)
    struct Gather {
        bool active = false, allocated = false;
        uint32_t world = 0, rank = 0, dst = 0, drawCap = 0, serial = 0;
        uint64_t cap = 0;                       // vertices per rank region
        uint8_t *color = nullptr, *depth = nullptr, *box = nullptr;     // the issuing rank's allocations (mapped over NVLink on the others)
        size_t statsOff = 0, mailOff = 0, boxBytes = 0;
        // per rank: 32-byte records | depth values | Custom0 values | arena slot per emitted corner (indexed frames), `cap` entries each
        uint8_t* inbox(uint32_t r) const { return box + (size_t)r * cap * 44; }
        uint8_t* stats(uint32_t r) const { return box + statsOff + (size_t)r * drawCap * sizeof(DrawStat); }
        uint32_t* mail(uint32_t r) const { return reinterpret_cast<uint32_t*>(box + mailOff) + (size_t)r * GATHER_MAIL_WORDS; }
        uint32_t* err() const { return reinterpret_cast<uint32_t*>(box + mailOff) + (size_t)world * GATHER_MAIL_WORDS; }
        void layout() {
            statsOff = ((size_t)world * cap * 44 + 255) & ~(size_t)255;
            mailOff = (statsOff + (size_t)world * drawCap * sizeof(DrawStat) + 255) & ~(size_t)255;
            boxBytes = mailOff + ((size_t)world + 1) * GATHER_MAIL_WORDS * 4;
        }
    } gather;
    DevBuf dGatherErr;          // shards: the timeout flag of k_gather_wait_ack (the issuing rank keeps its own in the mailbox block)
    PinVec<uint32_t> hMail; PinVec<DrawStat> hGatherStats;
    t3d_vertex *oColor = nullptr, *oDepth = nullptr, *oNormal = nullptr;     // where the current frame's lists are
    uint64_t xCapacity = 0;
    std::vector<void*> sharedAllocs, sharedOpened;

    // results (pinned)
    PinVec<DrawStat> hStats;
    PinVec<uint32_t> hCount;
    PinVec<t3d_draw_result> hResults, hResultsB;      // B: frames of the second set while two frames are in flight (their results outlive the next frame's)
    PinVec<t3d_vertex> hColor, hDepth, hNormal, hColorB, hDepthB, hNormalB;
    PinVec<uint32_t> hTriIds, hTriIdsB;
    PinVec<uint16_t> hIndex;
    uint64_t nVerticesOut = 0;

    cudaEvent_t ev[10] = {};
    uint64_t h2dBytes = 0, d2hBytes = 0;
    int launches = 0;
    uint64_t vertexKernelBytes = 0;
    float downloadMs = 0;
};

static thread_local std::string g_createErr;

static int fail(t3d_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_createErr = msg;
    return code;
}
static int cuda_fail(t3d_ctx* c, cudaError_t e, const char* what) {
    return fail(c, e == cudaErrorMemoryAllocation ? T3D_E_NOMEM : T3D_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
template <class T>
static cudaError_t upload_padded(T** dst, const T* src, size_t nElems, size_t nPadded, cudaStream_t s) {
    *dst = nullptr;
    cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(nPadded, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (nPadded > nElems) { e = cudaMemsetAsync(*dst + nElems, 0, (nPadded - nElems) * sizeof(T), s); if (e != cudaSuccess) return e; }
    if (nElems) e = cudaMemcpyAsync(*dst, src, nElems * sizeof(T), cudaMemcpyHostToDevice, s);
    return e;
}

// AoS [n][3] on the host -> SoA planes [3][nPad] on the device (zero padded).
template <class T>
static cudaError_t upload_planes3(T** dst, const T* src, size_t n, size_t nPad, std::vector<T>& tmp) {
    *dst = nullptr;
    nPad = std::max<size_t>(nPad, 4);
    tmp.assign(3 * nPad, T(0));
    for (size_t i = 0; i < n; i++) { tmp[i] = src[3 * i]; tmp[nPad + i] = src[3 * i + 1]; tmp[2 * nPad + i] = src[3 * i + 2]; }
    cudaError_t e = cudaMalloc((void**)dst, 3 * nPad * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, tmp.data(), 3 * nPad * sizeof(T), cudaMemcpyHostToDevice);
}

// PointLight.VertexLight skips an unranged light when `1 / d * energy < 0.001` (light.go:205-210; d the squared distance).  The
// left side -- two float32 roundings -- never increases with d, so the vertices that pass are exactly those with d <= T for one
// float T: found here by bisection over the positive floats with the reference's own expression, it lets the default lighting
// take the decision exactly with one comparison and no division (k_prep_draws hands it on in LightU.range).
static float energy_cutoff(float energy) {
    if (energy != energy) return INFINITY;          // NaN never compares below: every vertex passes
    if (!(energy > 0)) return 0.0f;                 // the product is <= 0 (or NaN only where 1 / d overflows): nothing at a distance passes
    auto passes = [&](uint32_t bits) { float d; memcpy(&d, &bits, 4); volatile float q = 1.0f / d; volatile float v = q * energy; return !(v < 0.001f); };
    uint32_t lo = 1u, hi = 0x7F800000u;             // passes(lo) (1 / d overflows to +inf), !passes(+inf) (1 / d is 0)
    while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (passes(mid)) lo = mid; else hi = mid; }
    float T; memcpy(&T, &lo, 4);
    return T;
}

#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call); } while (0)

extern "C" {

int t3d_abi_version(void) { return T3D_ABI_VERSION; }

const char* t3d_last_error(t3d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

int t3d_create(t3d_ctx** out, int device_id) {
    if (!out) return fail(nullptr, T3D_E_INVALID, "t3d_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, T3D_E_CUDA, std::string("t3d_create: no CUDA device (") + cudaGetErrorString(e) + "); libt3dgeo has no CPU fallback");
    if (device_id < 0 || device_id >= n) return fail(nullptr, T3D_E_INVALID, "t3d_create: device_id out of range");
    e = cudaSetDevice(device_id);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    t3d_ctx* ctx = new t3d_ctx();
    ctx->device = device_id;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    if (const char* fs = getenv("T3D_FUSE_SORT")) ctx->fuseSort = atoi(fs);
    if (const char* di = getenv("T3D_DELIVER_INDEXED")) ctx->deliverIndexed = di[0] != '0';
    cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device_id);
    if (ctx->smCount <= 0) ctx->smCount = 148;
    if (const char* th = getenv("T3D_DMA_THRESHOLD")) ctx->dmaThreshold = (size_t)strtoull(th, nullptr, 10);      // diagnostics: bytes; 0 = always copy, huge = never
    if (const char* cb = getenv("T3D_PREP_CHUNK_BLOCKS")) ctx->prepChunkBlocks = std::max(1u, (uint32_t)strtoul(cb, nullptr, 10));    // diagnostics / tests
    *out = ctx;
    return T3D_OK;
}

void t3d_destroy(t3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->meshes) kv.second.release();
    for (float* c : ctx->volumeCells) if (c) cudaFree(c);
    for (uint32_t* c : ctx->volumeCsr) if (c) cudaFree(c);
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); cudaEventDestroy(ctx->evFork); }
    if (ctx->dlStream) {
        cudaStreamSynchronize(ctx->dlStream); cudaStreamDestroy(ctx->dlStream);
        cudaEventDestroy(ctx->evFrameDone); cudaEventDestroy(ctx->evCopyDone[0]); cudaEventDestroy(ctx->evCopyDone[1]);
    }
    ctx->dColorB.release(); ctx->dDepthB.release(); ctx->dNormalB.release(); ctx->dTriIdsB.release();
    ctx->hColorB.release(); ctx->hDepthB.release(); ctx->hNormalB.release(); ctx->hTriIdsB.release();
    for (cudaEvent_t e : ctx->evChunk) cudaEventDestroy(e);
    for (auto& kv : ctx->skeletons) if (kv.second.block) cudaFree(kv.second.block);
    for (auto& kv : ctx->animations) if (kv.second.block) cudaFree(kv.second.block);
    ctx->hPoses.release(); ctx->hPoseChains.release();
    DevBuf* bufs[] = {&ctx->dVolumes, &ctx->dFrame, &ctx->dStage,
                      &ctx->dLightsU, &ctx->dStats, &ctx->dClip, &ctx->dClipXY, &ctx->dLight, &ctx->dWpos, &ctx->dUnbill,
                      &ctx->dTnrm, &ctx->dTdepth, &ctx->dSorted, &ctx->dHist, &ctx->dTileSurv, &ctx->dTileOut,
                      &ctx->dCount, &ctx->dTileCtr, &ctx->dGatherErr, &ctx->dColor, &ctx->dDepth, &ctx->dNormal, &ctx->dTriIds};
    for (DevBuf* b : bufs) b->release();
    ctx->hSlim.release(); ctx->hTmpl.release(); ctx->hRot.release();
    ctx->hFrame.release(); ctx->hVbase.release(); ctx->hTbase.release(); ctx->hTileV.release(); ctx->hLarge.release(); ctx->hPacked.release(); ctx->hMedium.release(); for (auto& v : ctx->hVClass) v.release(); ctx->hHuge.release(); ctx->hTiles.release(); ctx->hFused.release();
    ctx->hLights.release(); ctx->hBones.release(); ctx->hStats.release(); ctx->hCount.release(); ctx->hResults.release(); ctx->hResultsB.release();
    ctx->hMail.release(); ctx->hGatherStats.release(); ctx->hColor.release(); ctx->hDepth.release(); ctx->hNormal.release(); ctx->hTriIds.release(); ctx->hIndex.release();
    for (void* q : ctx->sharedOpened) cudaIpcCloseMemHandle(q);
    for (void* q : ctx->sharedAllocs) cudaFree(q);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---- residency -------------------------------------------------------------------------------------

int t3d_mesh_upload(t3d_ctx* ctx, uint64_t mesh_id, uint32_t generation, const t3d_mesh_desc* d) {
    if (!ctx || !d) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: NULL argument");
    CU(cudaSetDevice(ctx->device));
    if (!d->positions || !d->normals || !d->uvs || (d->n_tris && (!d->indices || !d->centers)) || (d->n_parts && !d->parts))
        return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: positions, normals, uvs, indices, centers and parts are required");
    if ((d->weights == nullptr) != (d->bones == nullptr))
        return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: weights and bones must be given together");
    if (d->n_shape_keys && (!d->key_positions || !d->key_normals))
        return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: shape key arrays missing");
    auto it = ctx->meshes.find(mesh_id);
    if (it != ctx->meshes.end() && it->second.generation == generation) return T3D_OK;      // already resident
    for (size_t t = 0; t < 3 * (size_t)d->n_tris; t++)
        if (d->indices[t] >= d->n_verts) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: triangle index out of range");
    if (ctx->inFrame && !ctx->executed && !ctx->hDraws.empty())
        return fail(ctx, T3D_E_STATE, "t3d_mesh_upload: draws of the open frame hold pointers into resident meshes (upload before t3d_frame_begin or after t3d_frame_execute)");
    // The replacement is built completely first; the resident mesh is only released once nothing can fail any more, and
    // every early return below frees what `m` has allocated so far.
    Mesh m;
    struct MeshGuard { Mesh* m; ~MeshGuard() { if (m) m->release(); } } guard{&m};
    m.generation = generation;
    m.nVerts = d->n_verts; m.nVertsPad = std::max((d->n_verts + 3u) & ~3u, 4u); m.nTris = d->n_tris; m.nKeys = d->n_shape_keys;
    m.nTrisPad = std::max((d->n_tris + 3u) & ~3u, 4u);
    m.maxSpan = d->max_span;
    const size_t nv = m.nVerts, np = m.nVertsPad;
    cudaStream_t s = ctx->stream;
    std::vector<float> ftmp; std::vector<uint32_t> utmp;
    CU(upload_planes3(&m.pos, d->positions, nv, np, ftmp));
    CU(upload_planes3(&m.nrm, d->normals, nv, np, ftmp));
    CU(upload_padded(&m.uv, d->uvs, nv * 2, np * 2, s));
    if (d->colors) CU(upload_padded(&m.col, d->colors, nv * 4, np * 4, s));
    if (d->weights) {
        CU(upload_padded(&m.wgt, d->weights, nv * 4, np * 4, s));
        CU(upload_padded(&m.jnt, d->bones, nv * 4, np * 4, s));
        if (d->n_influences) {
            for (size_t i = 0; i < nv; i++) if (d->n_influences[i] > 4) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: more than 4 influences per vertex");
            CU(upload_padded(&m.ninf, d->n_influences, nv, np, s));
        }
        for (size_t i = 0; i < nv * 4; i++) m.maxJoint = std::max<uint32_t>(m.maxJoint, d->bones[i]);
    }
    CU(upload_planes3(&m.idx, d->indices, (size_t)m.nTris, (size_t)m.nTrisPad, utmp));
    CU(upload_planes3(&m.ctr, d->centers, (size_t)m.nTris, (size_t)m.nTrisPad, ftmp));
    if (m.nKeys) {
        // key k: three planes of np floats at keyPos + k * 3 * np
        std::vector<float> kp((size_t)m.nKeys * 3 * np, 0.0f), kn((size_t)m.nKeys * 3 * np, 0.0f);
        for (uint32_t k = 0; k < m.nKeys; k++)
            for (size_t i = 0; i < nv; i++)
                for (int c = 0; c < 3; c++) {
                    kp[((size_t)k * 3 + c) * np + i] = d->key_positions[((size_t)k * nv + i) * 3 + c];
                    kn[((size_t)k * 3 + c) * np + i] = d->key_normals[((size_t)k * nv + i) * 3 + c];
                }
        CU(cudaMalloc((void**)&m.keyPos, kp.size() * sizeof(float)));
        CU(cudaMalloc((void**)&m.keyNrm, kn.size() * sizeof(float)));
        CU(cudaMalloc((void**)&m.keyW, (size_t)m.nKeys * sizeof(float)));
        CU(cudaMemcpy(m.keyPos, kp.data(), kp.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(m.keyNrm, kn.data(), kn.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemsetAsync(m.keyW, 0, (size_t)m.nKeys * sizeof(float), s));
    }
    if (d->tri_sub_base || d->tri_sub_level || d->tri_n_levels || d->tri_max_span || d->tri_normal) {
        if (!(d->tri_sub_base && d->tri_sub_level && d->tri_n_levels && d->tri_max_span && d->tri_normal)) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: auto-subdivision needs all five tri_* arrays");
        if (d->weights) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_mesh_upload: auto-subdivision of a skinned mesh (model.go:726 skips it)");
        const size_t t = d->n_tris;
        for (size_t i = 0; i < t; i++)
            if (d->tri_sub_base[i] >= t || d->tri_sub_level[i] < -1 || d->tri_sub_base[i] > i) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: tri_sub_base / tri_sub_level out of range");
        // one block: [SubdivMesh | base 4t | span 4t | normal 12t | level t | nLevels t]
        const size_t off0 = 64, bytes = off0 + t * (4 + 4 + 12 + 1 + 1);
        std::vector<uint8_t> h(bytes, 0);
        memcpy(h.data() + off0, d->tri_sub_base, 4 * t);
        memcpy(h.data() + off0 + 4 * t, d->tri_max_span, 4 * t);
        memcpy(h.data() + off0 + 8 * t, d->tri_normal, 12 * t);
        memcpy(h.data() + off0 + 20 * t, d->tri_sub_level, t);
        memcpy(h.data() + off0 + 21 * t, d->tri_n_levels, t);
        CU(cudaMalloc(&m.subdivBlock, bytes));
        uint8_t* b = static_cast<uint8_t*>(m.subdivBlock);
        SubdivMesh sm;
        sm.base = reinterpret_cast<const uint32_t*>(b + off0); sm.span = reinterpret_cast<const float*>(b + off0 + 4 * t);
        sm.normal = reinterpret_cast<const float*>(b + off0 + 8 * t); sm.level = reinterpret_cast<const int8_t*>(b + off0 + 20 * t);
        sm.nLevels = b + off0 + 21 * t;
        memcpy(h.data(), &sm, sizeof sm);
        static_assert(sizeof(SubdivMesh) <= 64, "SubdivMesh header");
        CU(cudaMemcpy(m.subdivBlock, h.data(), bytes, cudaMemcpyHostToDevice));
        m.subdiv = reinterpret_cast<SubdivMesh*>(m.subdivBlock);
    }
    m.parts.resize(d->n_parts);
    for (uint32_t p = 0; p < d->n_parts; p++) {
        const t3d_part_range& pr = d->parts[p];
        if (pr.tri_start > pr.tri_end || pr.tri_end > d->n_tris) return fail(ctx, T3D_E_INVALID, "t3d_mesh_upload: part triangle range out of bounds");
        PartInfo pi; pi.triStart = pr.tri_start; pi.triEnd = pr.tri_end; pi.a0 = 0; pi.nv = 0;
        if (pr.tri_end > pr.tri_start) {
            uint32_t lo = 0xFFFFFFFFu, hi = 0;
            for (size_t t = (size_t)pr.tri_start * 3; t < (size_t)pr.tri_end * 3; t++) { lo = std::min(lo, d->indices[t]); hi = std::max(hi, d->indices[t]); }
            pi.a0 = lo & ~3u;
            pi.nv = ((hi + 4u) & ~3u) - pi.a0;
        }
        m.parts[p] = pi;
    }
    // the source arrays are borrowed only for this call; the sync also retires every kernel that reads the old mesh
    CU(cudaStreamSynchronize(s));
    guard.m = nullptr;
    if (it != ctx->meshes.end()) { it->second.release(); it->second = std::move(m); }
    else ctx->meshes[mesh_id] = std::move(m);
    return T3D_OK;
}

int t3d_mesh_set_shape_key_weights(t3d_ctx* ctx, uint64_t mesh_id, const float* weights, uint32_t k) {
    if (!ctx || !weights) return fail(ctx, T3D_E_INVALID, "t3d_mesh_set_shape_key_weights: NULL argument");
    CU(cudaSetDevice(ctx->device));
    auto it = ctx->meshes.find(mesh_id);
    if (it == ctx->meshes.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_mesh_set_shape_key_weights: unknown mesh");
    if (it->second.nKeys != k) return fail(ctx, T3D_E_INVALID, "t3d_mesh_set_shape_key_weights: key count mismatch");
    if (k == 0) return T3D_OK;
    CU(cudaMemcpyAsync(it->second.keyW, weights, k * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return T3D_OK;
}

int t3d_mesh_release(t3d_ctx* ctx, uint64_t mesh_id) {
    if (!ctx) return T3D_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    auto it = ctx->meshes.find(mesh_id);
    if (it == ctx->meshes.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_mesh_release: unknown mesh");
    if (ctx->inFrame && !ctx->executed && !ctx->hDraws.empty())
        return fail(ctx, T3D_E_STATE, "t3d_mesh_release: draws of the open frame hold pointers into resident meshes");
    CU(cudaStreamSynchronize(ctx->stream));
    it->second.release();
    ctx->meshes.erase(it);
    return T3D_OK;
}

int t3d_volume_upload(t3d_ctx* ctx, uint32_t volume_id, const t3d_volume_desc* d) {
    if (!ctx || !d || !d->cell_colors) return fail(ctx, T3D_E_INVALID, "t3d_volume_upload: NULL argument");
    CU(cudaSetDevice(ctx->device));
    const size_t cells = (size_t)d->cell_count[0] * d->cell_count[1] * d->cell_count[2];
    if (cells == 0) return fail(ctx, T3D_E_INVALID, "t3d_volume_upload: empty grid");
    uint32_t slot;
    auto it = ctx->volumeIndex.find(volume_id);
    CU(cudaStreamSynchronize(ctx->stream));
    if (it == ctx->volumeIndex.end()) {
        slot = (uint32_t)ctx->volumes.size();
        ctx->volumes.push_back(VolumeU());
        ctx->volumeCells.push_back(nullptr);
        ctx->volumeCsr.push_back(nullptr); ctx->volumeUseCount.push_back(0); ctx->volumeUse.emplace_back();
        ctx->volumeIndex[volume_id] = slot;
    } else {
        slot = it->second;
        if (ctx->volumeCells[slot]) { cudaFree(ctx->volumeCells[slot]); ctx->volumeCells[slot] = nullptr; }
        if (ctx->volumeCsr[slot]) { cudaFree(ctx->volumeCsr[slot]); ctx->volumeCsr[slot] = nullptr; }
        ctx->volumeUseCount[slot] = 0;
    }
    float* dc = nullptr;
    CU(cudaMalloc((void**)&dc, cells * 3 * sizeof(float)));
    CU(cudaMemcpy(dc, d->cell_colors, cells * 3 * sizeof(float), cudaMemcpyHostToDevice));
    VolumeU V;
    for (int i = 0; i < 3; i++) {
        V.dimMin[i] = d->dim_min[i]; V.dimMax[i] = d->dim_max[i]; V.cellSize[i] = d->cell_size[i];
        V.cellCount[i] = (int32_t)d->cell_count[i]; V.offset[i] = d->offset[i]; V.bias[i] = d->light_bias[i];
    }
    V.cells = dc;
    V.cellLightStart = V.cellLightIndex = nullptr;
    if (d->cell_light_start) {
        const uint32_t nIdx = d->cell_light_start[cells];
        if (nIdx && !d->cell_light_index) return fail(ctx, T3D_E_INVALID, "t3d_volume_upload: cell_light_index is NULL");
        uint32_t useCount = 0;
        for (size_t c = 0; c < cells; c++) if (d->cell_light_start[c] > d->cell_light_start[c + 1]) return fail(ctx, T3D_E_INVALID, "t3d_volume_upload: cell_light_start must be non-decreasing");
        for (uint32_t k = 0; k < nIdx; k++) useCount = std::max(useCount, d->cell_light_index[k] + 1);
        if (nIdx) {
            uint32_t* csr = nullptr;
            CU(cudaMalloc((void**)&csr, 4 * (cells + 1 + nIdx)));
            CU(cudaMemcpy(csr, d->cell_light_start, 4 * (cells + 1), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(csr + cells + 1, d->cell_light_index, 4ull * nIdx, cudaMemcpyHostToDevice));
            ctx->volumeCsr[slot] = csr;
            ctx->volumeUseCount[slot] = useCount;
            V.cellLightStart = csr; V.cellLightIndex = csr + cells + 1;
        }
    }
    ctx->volumes[slot] = V;
    ctx->volumeCells[slot] = dc;
    ctx->volumesDirty = true;
    return T3D_OK;
}

int t3d_volume_release(t3d_ctx* ctx, uint32_t volume_id) {
    if (!ctx) return T3D_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    auto it = ctx->volumeIndex.find(volume_id);
    if (it == ctx->volumeIndex.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_volume_release: unknown volume");
    CU(cudaStreamSynchronize(ctx->stream));
    uint32_t slot = it->second;
    if (ctx->volumeCells[slot]) { cudaFree(ctx->volumeCells[slot]); ctx->volumeCells[slot] = nullptr; }
    if (ctx->volumeCsr[slot]) { cudaFree(ctx->volumeCsr[slot]); ctx->volumeCsr[slot] = nullptr; }
    ctx->volumeUseCount[slot] = 0;
    ctx->volumes[slot].cells = nullptr; ctx->volumes[slot].cellLightStart = ctx->volumes[slot].cellLightIndex = nullptr;
    ctx->volumeIndex.erase(it);
    ctx->volumesDirty = true;
    return T3D_OK;
}

// ---- per frame -------------------------------------------------------------------------------------

int t3d_frame_begin(t3d_ctx* ctx, const t3d_camera* cam, const t3d_world* world) {
    if (!ctx || !cam) return fail(ctx, T3D_E_INVALID, "t3d_frame_begin: NULL argument");
    CU(cudaSetDevice(ctx->device));
    if (cam->width <= 0 || cam->height <= 0) return fail(ctx, T3D_E_INVALID, "t3d_frame_begin: camera size must be positive");
    // the previous frame's outputs die here
    CU(cudaStreamSynchronize(ctx->stream));
    FrameConst& fc = ctx->fc;
    memset(&fc, 0, sizeof fc);
    memcpy(fc.vp, cam->vp, 64);
    memcpy(fc.view, cam->view, 64);
    memcpy(fc.camPos, cam->position, 12);
    memcpy(fc.camUp, cam->up, 12);
    fc.nearP = cam->near_plane; fc.farP = cam->far_plane;
    fc.camW = (float)cam->width; fc.camH = (float)cam->height;
    fc.halfW = (float)cam->width / 2; fc.halfH = (float)cam->height / 2;          // camera.go:1603
    fc.depthMargin = (cam->far_plane - cam->near_plane) * cam->depth_margin;       // camera.go:1339
    fc.camSpread = cam->far_plane - cam->near_plane + (fc.depthMargin * 2);        // camera.go:1341
    fc.snapping = cam->vertex_snapping;
    fc.camFlags = cam->flags;
    memcpy(fc.camForward, cam->forward, 12); memcpy(fc.camRight, cam->right, 12);
    fc.sphereFactorY = cam->sphere_factor_y; fc.sphereFactorX = cam->sphere_factor_x; fc.sphereFactorTang = cam->sphere_factor_tang;
    fc.orthoScale = cam->ortho_scale;
    if (world) {
        fc.worldPresent = world->present; fc.fogOn = world->fog_on; fc.fogMode = world->fog_mode;
        memcpy(fc.fogColor, world->fog_color, 12);
        memcpy(fc.fogRange, world->fog_range, 8);
    }
    ctx->hSlim.clear(); ctx->hTmpl.clear(); ctx->hRot.clear();
    ctx->hDraws.clear(); ctx->hVbase.clear(); ctx->hTbase.clear(); ctx->hLights.clear(); ctx->hBones.clear();
    ctx->lightLists.clear();
    ctx->lightLists.push_back(LightList{0, 0, 0});     // list 0: empty
    ctx->drawTotalTris.clear(); ctx->drawContinue.clear();
    ctx->vAcc = ctx->tAcc = ctx->luAcc = 0;
    ctx->anyLit = ctx->anySkinned = ctx->anyUnbill = ctx->anySubdiv = false;
    ctx->bonesAffine = ctx->transformsAffine = !ctx->forceFullUpload;
    ctx->tmplSpanMax = 1;
    ctx->hPoses.clear(); ctx->hPoseChains.clear(); ctx->poseCalls.clear(); ctx->posedBones = 0;
    ctx->trsGroups.clear(); ctx->nTrsDraws = 0; ctx->trsUniform = true;
    for (auto& u : ctx->volumeUse) u.clear();
    ctx->inFrame = true; ctx->executed = false; ctx->haveResults = false;
    ctx->nVerticesOut = 0;
    ctx->bakeFrame = false;
    return T3D_OK;
}

int t3d_bake_ao(t3d_ctx* ctx, uint64_t mesh_id, const float* tri_normals3, const float* tri_max_span, const uint8_t* target_parts,
                const uint8_t* target_vertices, float occlusion_angle, const float occlusion_color[4], float* colors4) {
    if (!ctx || !tri_normals3 || !tri_max_span || !occlusion_color || !colors4) return fail(ctx, T3D_E_INVALID, "t3d_bake_ao: NULL argument");
    auto it = ctx->meshes.find(mesh_id);
    if (it == ctx->meshes.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_bake_ao: unknown mesh (t3d_mesh_upload first)");
    const Mesh& m = it->second;
    if (m.subdiv) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_bake_ao: mesh with auto-subdivision data (its triangle arrays hold the sub-triangles too)");
    if (!m.nTris || !m.nVerts) return T3D_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::vector<uint8_t> triTarget;
    if (target_parts) {         // TargetMeshParts (model.go:1054-1066): a triangle takes part when its MeshPart is listed
        triTarget.assign(m.nTris, 0);
        for (size_t p = 0; p < m.parts.size(); p++)
            if (target_parts[p]) for (uint32_t t = m.parts[p].triStart; t < m.parts[p].triEnd && t < m.nTris; t++) triTarget[t] = 1;
    }
    DevBuf dN, dS, dTT, dVT, dM, dC;
    struct Free { DevBuf* b[6]; ~Free() { for (DevBuf* x : b) x->release(); } } guard{{&dN, &dS, &dTT, &dVT, &dM, &dC}};
    CU(dN.ensure(12ull * m.nTris)); CU(dS.ensure(4ull * m.nTris)); CU(dM.ensure(m.nTris)); CU(dC.ensure(16ull * m.nVerts));
    CU(cudaMemcpyAsync(dN.p, tri_normals3, 12ull * m.nTris, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dS.p, tri_max_span, 4ull * m.nTris, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dC.p, colors4, 16ull * m.nVerts, cudaMemcpyHostToDevice, s));
    if (target_parts) { CU(dTT.ensure(m.nTris)); CU(cudaMemcpyAsync(dTT.p, triTarget.data(), m.nTris, cudaMemcpyHostToDevice, s)); }
    if (target_vertices) { CU(dVT.ensure(m.nVerts)); CU(cudaMemcpyAsync(dVT.p, target_vertices, m.nVerts, cudaMemcpyHostToDevice, s)); }
    launch_ao(s, m.pos, m.nVertsPad, m.idx, m.ctr, m.nTrisPad, m.nTris, m.nVerts, dN.as<float>(), dS.as<float>(),
              target_parts ? dTT.as<uint8_t>() : nullptr, target_vertices ? dVT.as<uint8_t>() : nullptr, occlusion_angle, occlusion_color,
              dM.as<uint8_t>(), dC.as<float4>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(colors4, dC.p, 16ull * m.nVerts, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return T3D_OK;
}

int t3d_frame_abort(t3d_ctx* ctx) {
    if (!ctx) return T3D_E_INVALID;
    if (ctx->inFrame && !ctx->executed) ctx->inFrame = false;      // what was submitted is dropped at the next t3d_frame_begin
    return T3D_OK;
}

int t3d_frame_set_bake_lighting(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_frame_set_bake_lighting: not between frame_begin and frame_execute");
    ctx->bakeFrame = enabled != 0;
    return T3D_OK;
}

int t3d_frame_set_autosubdivision(t3d_ctx* ctx, const float* levels2, uint32_t n) {
    if (!ctx || (n && !levels2)) return fail(ctx, T3D_E_INVALID, "t3d_frame_set_autosubdivision: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_frame_set_autosubdivision: not between frame_begin and frame_execute");
    if (n > 8) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_frame_set_autosubdivision: more than 8 levels");
    for (uint32_t i = 0; i < n; i++)
        if (!(levels2[2 * i + 1] > 0.0f)) return fail(ctx, T3D_E_INVALID, "t3d_frame_set_autosubdivision: MaximumTriangleSize must be positive");
    ctx->fc.nSubLevels = n;
    for (uint32_t i = 0; i < n; i++) { ctx->fc.subLevels[i][0] = levels2[2 * i]; ctx->fc.subLevels[i][1] = levels2[2 * i + 1]; }
    return T3D_OK;
}

int t3d_submit_lights(t3d_ctx* ctx, const t3d_light* lights, uint32_t n, uint32_t* list_id) {
    if (!ctx || !list_id || (n && !lights)) return fail(ctx, T3D_E_INVALID, "t3d_submit_lights: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_submit_lights: not between frame_begin and frame_execute");
    const uint32_t off = (uint32_t)ctx->hLights.n;
    CU(ctx->hLights.reserve(off + n));
    uint32_t total = n;
    for (uint32_t i = 0; i < n; i++) {
        t3d_light L = lights[i];
        if (L.type > T3D_LIGHT_VOLUME) return fail(ctx, T3D_E_INVALID, "t3d_submit_lights: unknown light type");
        if (L.type == T3D_LIGHT_VOLUME) {
            auto it = ctx->volumeIndex.find(L.volume);
            if (it == ctx->volumeIndex.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_submit_lights: unknown light volume");
            const uint32_t slot = it->second;
            // the lights the volume's cells use follow the list's own lights; the volume entry remembers where (k_prep_draws
            // derives a LightU for every one of them, k_vertex finds them through LightU.useFirst)
            if (slot >= (1u << 12) || total >= (1u << 20)) return fail(ctx, T3D_E_INVALID, "t3d_submit_lights: too many volumes or lights");
            L.volume = slot;
            if (ctx->volumeUseCount[slot]) {
                if (ctx->volumeUse[slot].size() < ctx->volumeUseCount[slot])
                    return fail(ctx, T3D_E_STATE, "t3d_submit_lights: the volume's cells use lights: call t3d_volume_use_lights first");
                L.volume = slot | (total << 12);
                total += ctx->volumeUseCount[slot];
            }
        }
        if (L.type == T3D_LIGHT_POINT) L.forward[0] = energy_cutoff(L.energy);      // (a point light has no forward)
        ctx->hLights.p[off + i] = L;
    }
    ctx->hLights.n = off + n;          // reserve keeps the first n elements when it has to move the array
    CU(ctx->hLights.reserve(off + total));
    uint32_t at = off + n;
    for (uint32_t i = 0; i < n; i++) {
        const t3d_light& L = ctx->hLights.p[off + i];
        if (L.type != T3D_LIGHT_VOLUME || !(L.volume >> 12)) continue;
        const uint32_t slot = L.volume & 0xFFFu;
        for (uint32_t k = 0; k < ctx->volumeUseCount[slot]; k++) ctx->hLights.p[at++] = ctx->volumeUse[slot][k];
    }
    ctx->hLights.n = off + total;
    *list_id = (uint32_t)ctx->lightLists.size();
    ctx->lightLists.push_back(LightList{off, n, total});
    return T3D_OK;
}

int t3d_volume_use_lights(t3d_ctx* ctx, uint32_t volume_id, const t3d_light* lights, uint32_t n) {
    if (!ctx || (n && !lights)) return fail(ctx, T3D_E_INVALID, "t3d_volume_use_lights: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_volume_use_lights: not between frame_begin and frame_execute");
    auto it = ctx->volumeIndex.find(volume_id);
    if (it == ctx->volumeIndex.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_volume_use_lights: unknown light volume");
    for (uint32_t i = 0; i < n; i++)
        if (lights[i].type >= T3D_LIGHT_VOLUME) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_volume_use_lights: ambient, point and directional lights only");
    if (n < ctx->volumeUseCount[it->second]) return fail(ctx, T3D_E_INVALID, "t3d_volume_use_lights: fewer lights than the volume's cells refer to");
    ctx->volumeUse[it->second].assign(lights, lights + n);
    return T3D_OK;
}

int t3d_skeleton_upload(t3d_ctx* ctx, uint64_t skeleton_id, const t3d_skeleton_desc* d) {
    if (!ctx || !d || !d->parent || !d->inverse_bind || !d->rest_position || !d->rest_scale || !d->rest_rotation)
        return fail(ctx, T3D_E_INVALID, "t3d_skeleton_upload: NULL argument");
    if (d->n_bones == 0) return fail(ctx, T3D_E_INVALID, "t3d_skeleton_upload: no bones");
    if (d->n_bones > POSE_MAX_BONES) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_skeleton_upload: more than 1024 bones (submit the matrices with t3d_submit_bones)");
    const uint32_t n = d->n_bones;
    std::vector<uint16_t> depth(n);
    uint32_t maxDepth = 0;
    for (uint32_t b = 0; b < n; b++) {
        const int32_t par = d->parent[b];
        if (par >= (int32_t)b) return fail(ctx, T3D_E_INVALID, "t3d_skeleton_upload: a parent must precede its children");
        depth[b] = par < 0 ? 0 : (uint16_t)(depth[par] + 1);
        maxDepth = std::max<uint32_t>(maxDepth, depth[b]);
    }
    if (ctx->inFrame && !ctx->executed && !ctx->poseCalls.empty()) return fail(ctx, T3D_E_STATE, "t3d_skeleton_upload: poses of the open frame are pending");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    Skeleton& S = ctx->skeletons[skeleton_id];
    if (S.block) { cudaFree(S.block); S.block = nullptr; }
    // [invBind 64n | restRot 64n | restPos 12n | restScale 12n | parent 4n | depth 2n]
    const size_t bytes = (size_t)n * (64 + 64 + 12 + 12 + 4 + 2);
    std::vector<uint8_t> h(bytes);
    uint8_t* q = h.data();
    memcpy(q, d->inverse_bind, 64ull * n); q += 64ull * n;
    memcpy(q, d->rest_rotation, 64ull * n); q += 64ull * n;
    memcpy(q, d->rest_position, 12ull * n); q += 12ull * n;
    memcpy(q, d->rest_scale, 12ull * n); q += 12ull * n;
    memcpy(q, d->parent, 4ull * n); q += 4ull * n;
    memcpy(q, depth.data(), 2ull * n);
    CU(cudaMalloc(&S.block, bytes));
    CU(cudaMemcpy(S.block, h.data(), bytes, cudaMemcpyHostToDevice));
    uint8_t* b = static_cast<uint8_t*>(S.block);
    S.invBind = reinterpret_cast<float*>(b); b += 64ull * n;
    S.restRot = reinterpret_cast<float*>(b); b += 64ull * n;
    S.restPos = reinterpret_cast<float*>(b); b += 12ull * n;
    S.restScale = reinterpret_cast<float*>(b); b += 12ull * n;
    S.parent = reinterpret_cast<int32_t*>(b); b += 4ull * n;
    S.depth = reinterpret_cast<uint16_t*>(b);
    S.nBones = n; S.maxDepth = maxDepth;
    return T3D_OK;
}

int t3d_skeleton_release(t3d_ctx* ctx, uint64_t skeleton_id) {
    if (!ctx) return T3D_E_INVALID;
    auto it = ctx->skeletons.find(skeleton_id);
    if (it == ctx->skeletons.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_skeleton_release: unknown skeleton");
    if (ctx->inFrame && !ctx->executed && !ctx->poseCalls.empty()) return fail(ctx, T3D_E_STATE, "t3d_skeleton_release: poses of the open frame are pending");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (it->second.block) cudaFree(it->second.block);
    ctx->skeletons.erase(it);
    return T3D_OK;
}

int t3d_animation_upload(t3d_ctx* ctx, uint64_t animation_id, const t3d_animation_desc* d) {
    if (!ctx || !d || !d->tracks || (d->n_keys && (!d->key_times || !d->key_values)))
        return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: NULL argument");
    if (animation_id == 0) return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: id 0 is the rest pose");
    if (d->n_bones == 0 || d->n_bones > POSE_MAX_BONES) return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: bone count out of range");
    for (uint32_t t = 0; t < 3 * d->n_bones; t++) {
        const t3d_track& tr = d->tracks[t];
        if ((uint64_t)tr.first_key + tr.n_keys > d->n_keys) return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: track outside the keyframe arrays");
        if (tr.interpolation > T3D_INTERPOLATION_CUBIC) return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: unknown interpolation");
        for (uint32_t k = 1; k < tr.n_keys; k++)
            if (!(d->key_times[tr.first_key + k] >= d->key_times[tr.first_key + k - 1]))
                return fail(ctx, T3D_E_INVALID, "t3d_animation_upload: keyframe times must be non-decreasing inside a track");
    }
    if (ctx->inFrame && !ctx->executed && !ctx->poseCalls.empty()) return fail(ctx, T3D_E_STATE, "t3d_animation_upload: poses of the open frame are pending");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    Animation& A = ctx->animations[animation_id];
    if (A.block) { cudaFree(A.block); A.block = nullptr; }
    // [keyValues 16k | tracks 36n | keyTimes 4k]
    const size_t nk = d->n_keys, nt = 3ull * d->n_bones;
    const size_t bytes = 16 * std::max<size_t>(nk, 1) + sizeof(t3d_track) * nt + 4 * std::max<size_t>(nk, 1);
    std::vector<uint8_t> h(bytes, 0);
    memcpy(h.data(), d->key_values, 16 * nk);
    memcpy(h.data() + 16 * std::max<size_t>(nk, 1), d->tracks, sizeof(t3d_track) * nt);
    memcpy(h.data() + 16 * std::max<size_t>(nk, 1) + sizeof(t3d_track) * nt, d->key_times, 4 * nk);
    CU(cudaMalloc(&A.block, bytes));
    CU(cudaMemcpy(A.block, h.data(), bytes, cudaMemcpyHostToDevice));
    uint8_t* b = static_cast<uint8_t*>(A.block);
    A.keyValues = reinterpret_cast<float4*>(b);
    A.tracks = reinterpret_cast<t3d_track*>(b + 16 * std::max<size_t>(nk, 1));
    A.keyTimes = reinterpret_cast<float*>(b + 16 * std::max<size_t>(nk, 1) + sizeof(t3d_track) * nt);
    A.nBones = d->n_bones; A.nKeys = d->n_keys;
    return T3D_OK;
}

int t3d_animation_release(t3d_ctx* ctx, uint64_t animation_id) {
    if (!ctx) return T3D_E_INVALID;
    auto it = ctx->animations.find(animation_id);
    if (it == ctx->animations.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_animation_release: unknown animation");
    if (ctx->inFrame && !ctx->executed && !ctx->poseCalls.empty()) return fail(ctx, T3D_E_STATE, "t3d_animation_release: poses of the open frame are pending");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (it->second.block) cudaFree(it->second.block);
    ctx->animations.erase(it);
    return T3D_OK;
}

static int submit_poses(t3d_ctx* ctx, const char* who, uint64_t skeleton_id, bool blended, uint64_t from_animation_id, uint64_t animation_id,
                        const t3d_pose* poses, uint32_t n, uint32_t* offset) {
    char msg[160];
    auto bad = [&](int code, const char* what) { snprintf(msg, sizeof msg, "%s: %s", who, what); return fail(ctx, code, msg); };
    if (!ctx || !offset || (n && !poses)) return bad(T3D_E_INVALID, "NULL argument");
    if (!ctx->inFrame || ctx->executed) return bad(T3D_E_STATE, "not between frame_begin and frame_execute");
    auto sk = ctx->skeletons.find(skeleton_id);
    if (sk == ctx->skeletons.end()) return bad(T3D_E_NOTFOUND, "unknown skeleton");
    for (uint64_t id : {animation_id, blended ? from_animation_id : 0}) {
        if (!id) continue;
        auto an = ctx->animations.find(id);
        if (an == ctx->animations.end()) return bad(T3D_E_NOTFOUND, "unknown animation");
        if (an->second.nBones > sk->second.nBones) return bad(T3D_E_INVALID, "the animation has more bones than the skeleton");
    }
    if ((uint64_t)ctx->posedBones + (uint64_t)n * sk->second.nBones >= POSE_TAG) return bad(T3D_E_INVALID, "too many posed bones in one frame");
    PoseCall c; c.skeleton = skeleton_id; c.animation = animation_id; c.firstPose = (uint32_t)ctx->hPoses.n; c.nPoses = n; c.firstBone = ctx->posedBones;
    c.fromAnimation = blended ? from_animation_id : 0; c.blended = blended; c.nLevels = 0;
    CU(ctx->hPoses.reserve(ctx->hPoses.n + n));
    memcpy(ctx->hPoses.p + ctx->hPoses.n, poses, sizeof(t3d_pose) * (size_t)n);
    ctx->hPoses.n += n;
    ctx->poseCalls.push_back(c);
    *offset = POSE_TAG | ctx->posedBones;
    ctx->posedBones += n * sk->second.nBones;
    return T3D_OK;
}

int t3d_submit_poses(t3d_ctx* ctx, uint64_t skeleton_id, uint64_t animation_id, const t3d_pose* poses, uint32_t n, uint32_t* offset) {
    return submit_poses(ctx, "t3d_submit_poses", skeleton_id, false, 0, animation_id, poses, n, offset);
}

int t3d_submit_poses_blended(t3d_ctx* ctx, uint64_t skeleton_id, uint64_t from_animation_id, uint64_t animation_id, const t3d_pose* poses,
                             uint32_t n, uint32_t* offset) {
    return submit_poses(ctx, "t3d_submit_poses_blended", skeleton_id, true, from_animation_id, animation_id, poses, n, offset);
}

int t3d_submit_pose_chains(t3d_ctx* ctx, uint64_t skeleton_id, const uint64_t* animation_ids, uint32_t n_levels, const t3d_pose_chain* poses,
                           uint32_t n, uint32_t* offset) {
    if (!ctx || !offset || !animation_ids || (n && !poses)) return fail(ctx, T3D_E_INVALID, "t3d_submit_pose_chains: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_submit_pose_chains: not between frame_begin and frame_execute");
    if (n_levels < 1 || n_levels > T3D_POSE_CHAIN_MAX) return fail(ctx, T3D_E_INVALID, "t3d_submit_pose_chains: n_levels must be 1 .. T3D_POSE_CHAIN_MAX");
    auto sk = ctx->skeletons.find(skeleton_id);
    if (sk == ctx->skeletons.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_submit_pose_chains: unknown skeleton");
    for (uint32_t k = 0; k < n_levels; k++) {
        auto an = ctx->animations.find(animation_ids[k]);
        if (an == ctx->animations.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_submit_pose_chains: unknown animation");
        if (an->second.nBones > sk->second.nBones) return fail(ctx, T3D_E_INVALID, "t3d_submit_pose_chains: the animation has more bones than the skeleton");
    }
    if ((uint64_t)ctx->posedBones + (uint64_t)n * sk->second.nBones >= POSE_TAG) return fail(ctx, T3D_E_INVALID, "t3d_submit_pose_chains: too many posed bones in one frame");
    PoseCall c; memset(&c, 0, sizeof c);
    c.skeleton = skeleton_id; c.firstPose = (uint32_t)ctx->hPoseChains.n; c.nPoses = n; c.firstBone = ctx->posedBones; c.nLevels = n_levels;
    for (uint32_t k = 0; k < n_levels; k++) c.chain[k] = animation_ids[k];
    ctx->hPoseChains.mapped = true;
    CU(ctx->hPoseChains.reserve(ctx->hPoseChains.n + n));
    memcpy(ctx->hPoseChains.p + ctx->hPoseChains.n, poses, sizeof(t3d_pose_chain) * (size_t)n);
    ctx->hPoseChains.n += n;
    ctx->poseCalls.push_back(c);
    *offset = POSE_TAG | ctx->posedBones;
    ctx->posedBones += n * sk->second.nBones;
    return T3D_OK;
}

int t3d_submit_bones(t3d_ctx* ctx, const float* matrices16, uint32_t n, uint32_t* offset) {
    if (!ctx || !offset || (n && !matrices16)) return fail(ctx, T3D_E_INVALID, "t3d_submit_bones: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_submit_bones: not between frame_begin and frame_execute");
    const size_t off = ctx->hBones.n;
    CU(ctx->hBones.reserve(off + (size_t)n * 16));
    memcpy(ctx->hBones.p + off, matrices16, (size_t)n * 64);
    ctx->hBones.n = off + (size_t)n * 16;
    // boneInfluence = inverseBind x Transform (node.go:592-594) is affine in every ordinary scene: column 3 is exactly
    // (0, 0, 0, 1), and then only 48 of the 64 bytes cross PCIe (k_prep_draws puts the column back)
    for (uint32_t i = 0; i < n && ctx->bonesAffine; i++) {
        const float* M = matrices16 + 16 * (size_t)i;
        ctx->bonesAffine = M[3] == 0.0f && M[7] == 0.0f && M[11] == 0.0f && M[15] == 1.0f &&
                           !std::signbit(M[3]) && !std::signbit(M[7]) && !std::signbit(M[11]);
    }
    *offset = (uint32_t)(off / 16);
    return T3D_OK;
}

// Node.Transform of an instance (node.go:577-583): Scale . rotation . Translate [. parent], each step the reference's Mult.
// Host and device (k_prep_draws) run this same sequence of t3d_math.h operations.
static Mat4 instance_transform(const float* pos3, const float* scale3, const Mat4& rotation, const Mat4* parent) {
    Mat4 M = m4_mult(m4_scale(scale3[0], scale3[1], scale3[2]), rotation);
    Mat4 T = m4_identity(); T.m[3][0] = pos3[0]; T.m[3][1] = pos3[1]; T.m[3][2] = pos3[2];
    M = m4_mult(M, T);
    if (parent) M = m4_mult(M, *parent);
    return M;
}
static bool affine16(const float* M) {
    return M[3] == 0.0f && M[7] == 0.0f && M[11] == 0.0f && M[15] == 1.0f && !std::signbit(M[3]) && !std::signbit(M[7]) && !std::signbit(M[11]);
}

// draws[0..n) -- or, with `inst`, n instances of draws[0] belonging to ctx->trsGroups[group] (t3d_submit_instances)
static int submit_range(t3d_ctx* ctx, const t3d_draw* draws, uint32_t n, const t3d_instance* inst, uint32_t group);

int t3d_submit_draws(t3d_ctx* ctx, const t3d_draw* draws, uint32_t n) {
    if (!ctx || (n && !draws)) return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_submit_draws: not between frame_begin and frame_execute");
    return submit_range(ctx, draws, n, nullptr, 0);
}

int t3d_submit_instances(t3d_ctx* ctx, const t3d_draw* proto, const float* rotation16, const float* parent16, const float* frustum_bounds9,
                         const t3d_instance* instances, uint32_t n) {
    if (!ctx || !proto || (n && !instances)) return fail(ctx, T3D_E_INVALID, "t3d_submit_instances: NULL argument");
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_submit_instances: not between frame_begin and frame_execute");
    if (proto->flags & T3D_DRAW_SKINNED) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_submit_instances: skinned models carry their own bones: submit them with t3d_submit_draws");
    if ((rotation16 && !affine16(rotation16)) || (parent16 && !affine16(parent16)))
        return fail(ctx, T3D_E_UNSUPPORTED, "t3d_submit_instances: rotation and parent transform must be affine (column 3 = 0, 0, 0, 1)");
    if (!n) return T3D_OK;
    t3d_ctx::TrsGroup G;
    G.rotation = rotation16 ? m4_load(rotation16) : m4_identity();
    G.hasParent = parent16 != nullptr;
    G.parentT = parent16 ? m4_load(parent16) : m4_identity();
    CU(ctx->hRot.reserve(ctx->hRot.n + 48));
    G.rot = (uint32_t)(ctx->hRot.n / 16);
    m4_store(G.rotation, ctx->hRot.p + ctx->hRot.n); ctx->hRot.n += 16;
    G.parent = NO_ROTATION;
    if (G.hasParent) { G.parent = (uint32_t)(ctx->hRot.n / 16); m4_store(G.parentT, ctx->hRot.p + ctx->hRot.n); ctx->hRot.n += 16; }
    G.hasBounds = frustum_bounds9 != nullptr;
    if (G.hasBounds) {          // the culling bounds ride behind the group's matrices (k_prep_draws finds them there)
        memcpy(G.bounds, frustum_bounds9, 36);
        memset(ctx->hRot.p + ctx->hRot.n, 0, 64); memcpy(ctx->hRot.p + ctx->hRot.n, frustum_bounds9, 36); ctx->hRot.n += 16;
    }
    ctx->trsGroups.push_back(G);
    return submit_range(ctx, proto, n, instances, (uint32_t)ctx->trsGroups.size() - 1);
}

static int submit_range(t3d_ctx* ctx, const t3d_draw* draws, uint32_t n, const t3d_instance* inst, uint32_t group) {
    if (ctx->hDraws.size() + n >= 0xFFFFFFF0ull) return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: too many draws in one frame");
    ctx->hDraws.reserve(ctx->hDraws.size() + n);
    CU(ctx->hVbase.reserve(ctx->hDraws.size() + n + 1));
    CU(ctx->hTbase.reserve(ctx->hDraws.size() + n + 1));
    CU(ctx->hSlim.reserve(ctx->hSlim.n + n));
    ctx->drawTotalTris.reserve(ctx->drawTotalTris.size() + n); ctx->drawContinue.reserve(ctx->drawContinue.size() + n);
    const Mesh* lastMesh = nullptr; uint64_t lastMeshId = 0;
    for (uint32_t i = 0; i < n; i++) {
        const t3d_draw& in = inst ? draws[0] : draws[i];
        if (!lastMesh || in.mesh_id != lastMeshId) {        // runs of instances share a mesh: one lookup per run
            auto it = ctx->meshes.find(in.mesh_id);
            if (it == ctx->meshes.end()) return fail(ctx, T3D_E_NOTFOUND, "t3d_submit_draws: unknown mesh");
            lastMesh = &it->second; lastMeshId = in.mesh_id;
        }
        const Mesh& m = *lastMesh;
        if (in.part >= m.parts.size()) return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: part index out of range");
        if (in.light_list >= ctx->lightLists.size()) return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: unknown light list");
        if (in.sort_mode > T3D_SORT_NONE || in.lighting_mode > T3D_LIGHTING_DOUBLE_SIDED || in.light_volume_mode > T3D_LV_OFF)
            return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: enum out of range");
        const PartInfo& pi = m.parts[in.part];
        uint32_t flags = in.flags;
        if (inst && i) flags |= T3D_DRAW_CONTINUE_FLUSH;
        if (flags & T3D_DRAW_SKINNED) {
            if (!m.wgt) return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: skinned draw of a mesh without weights");
            const bool posed = in.bone_offset & POSE_TAG;
            if (in.bone_count <= m.maxJoint || (size_t)(in.bone_offset & ~POSE_TAG) + in.bone_count > (posed ? (size_t)ctx->posedBones : ctx->hBones.n / 16))
                return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: bone span does not cover the mesh's joints");
            flags &= ~(uint32_t)T3D_DRAW_DEPTH_UNBILLBOARDED;   // the reference never writes it for skinned models (model.go:800)
        }
        if ((uint64_t)ctx->vAcc + pi.nv > 0xFFFFFFF0ull || (uint64_t)ctx->tAcc + (pi.triEnd - pi.triStart) > 0x55555550ull)
            return fail(ctx, T3D_E_INVALID, "t3d_submit_draws: frame exceeds 32-bit arena indexing");
        DrawRec R;
        memset(&R, 0, sizeof R);
        R.pos = m.pos; R.nrm = m.nrm; R.uv = m.uv; R.col = m.col; R.wgt = m.wgt; R.jnt = m.jnt; R.ninf = m.ninf;
        R.idx = m.idx; R.ctr = m.ctr; R.keyPos = m.keyPos; R.keyNrm = m.keyNrm; R.keyW = m.keyW; R.subdiv = m.subdiv;
        R.a0 = pi.a0; R.nv = pi.nv; R.vbase = ctx->vAcc;
        R.t0 = pi.triStart; R.nt = pi.triEnd - pi.triStart; R.tbase = ctx->tAcc;
        R.flags = flags; R.nKeys = m.nKeys; R.vstride = m.nVertsPad; R.tstride = m.nTrisPad;
        R.boneOffset = in.bone_offset;
        const LightList& ll = ctx->lightLists[in.light_list];
        R.lightIn = ll.off; R.lightCount = ll.count; R.lightTotal = ll.total; R.lightU = ctx->luAcc;
        R.sortMode = (flags & T3D_DRAW_HAS_MATERIAL) ? in.sort_mode : (uint8_t)T3D_SORT_BACK_TO_FRONT;   // camera.go:1540-1543
        R.lightingMode = in.lighting_mode; R.lvMode = in.light_volume_mode;
        if (!inst) { memcpy(R.transform, in.transform, 64); memcpy(R.color, in.color, 16); }
        else for (int c = 0; c < 4; c++) R.color[c] = inst[i].color[c] * in.color[c];     // mpColor = model.Color x Material.Color (camera.go:1597-1601)
        memcpy(R.worldRot, in.world_rotation, 64);
        R.srcW = in.src_w; R.srcH = in.src_h; R.maxSpan = m.maxSpan;
        R.nBones = (flags & T3D_DRAW_SKINNED) ? m.maxJoint + 1 : 0;
        if (flags & T3D_DRAW_LIT) { ctx->luAcc += ll.total; ctx->anyLit = true; }
        if (flags & T3D_DRAW_SKINNED) ctx->anySkinned = true;
        if (flags & T3D_DRAW_DEPTH_UNBILLBOARDED) ctx->anyUnbill = true;
        if (m.subdiv) ctx->anySubdiv = true;
        ctx->hVbase.p[ctx->hDraws.size()] = ctx->vAcc;
        ctx->hTbase.p[ctx->hDraws.size()] = ctx->tAcc;
        ctx->hVbase.n = ctx->hTbase.n = ctx->hDraws.size() + 1;        // (a later call's reserve keeps the first n entries when it has to move the array)
        ctx->vAcc += pi.nv; ctx->tAcc += R.nt;
        {   // what goes over PCIe: a 112-byte placement record, plus a template whenever anything else changes
            // reuse one of the most recent templates when everything but the placement matches (instances of one
            // MeshPart; models cycling through a handful of LightGroups or MeshParts)
            uint32_t tmplIndex = 0xFFFFFFFFu;
            const size_t nT = ctx->hTmpl.n;
            for (size_t back = 1; back <= 16 && back <= nT; back++) {
                const DrawRec& T = ctx->hTmpl.p[nT - back];
                if (T.pos == R.pos && T.idx == R.idx && T.a0 == R.a0 && T.nv == R.nv && T.t0 == R.t0 && T.nt == R.nt && T.flags == R.flags &&
                    T.lightIn == R.lightIn && T.lightCount == R.lightCount && T.sortMode == R.sortMode && T.lightingMode == R.lightingMode &&
                    T.lvMode == R.lvMode && T.srcW == R.srcW && T.srcH == R.srcH && T.col == R.col && T.uv == R.uv && T.nrm == R.nrm) {
                    tmplIndex = (uint32_t)(nT - back);
                    break;
                }
            }
            if (tmplIndex == 0xFFFFFFFFu) {
                DrawRec T = R;      // every mesh-, material- and light-list-derived field; the per-draw ones are patched on the device
                T.vbase = T.tbase = T.boneOffset = T.lightU = 0;
                memset(T.transform, 0, sizeof T.transform); memset(T.worldRot, 0, sizeof T.worldRot); memset(T.color, 0, sizeof T.color);
                CU(ctx->hTmpl.push(T));
                tmplIndex = (uint32_t)ctx->hTmpl.n - 1;
            }
            DrawSlim S;
            memset(&S, 0, sizeof S);
            S.tmpl = tmplIndex; S.vbase = R.vbase; S.tbase = R.tbase; S.boneOffset = R.boneOffset; S.lightU = R.lightU;
            memcpy(S.transform, R.transform, 64); memcpy(S.color, R.color, 16);
            if (inst) {         // position and scale instead of a matrix: expanded on the device, or at execute time when the frame is not shipped as runs
                memcpy(S.transform, inst[i].position, 12); memcpy(S.transform + 3, inst[i].scale, 12);
                S.pad[0] = (group + 1) | (ctx->trsGroups[group].hasBounds ? 0x80000000u : 0u);
                ctx->nTrsDraws++;
                if (!(inst[i].scale[0] == inst[i].scale[1] && inst[i].scale[1] == inst[i].scale[2])) ctx->trsUniform = false;
            } else if (ctx->transformsAffine) {       // Node.Transform() of an ordinary scene graph: column 3 is exactly (0, 0, 0, 1)
                ctx->transformsAffine = affine16(R.transform);
            }
            // the world rotation only feeds a directional light's workingModelRotation of an unskinned, lit draw (light.go:326-330)
            S.rot = NO_ROTATION;
            if ((flags & T3D_DRAW_LIT) && !(flags & T3D_DRAW_SKINNED) && ll.total) {
                bool sun = false;
                for (uint32_t k = 0; k < ll.total && !sun; k++) sun = ctx->hLights.p[ll.off + k].type == T3D_LIGHT_DIRECTIONAL;
                if (sun && inst && i) S.rot = ctx->hSlim.p[ctx->hSlim.n - 1].rot;      // instances share their world rotation: shipped once
                else if (sun) {
                    S.rot = (uint32_t)(ctx->hRot.n / 16);
                    CU(ctx->hRot.reserve(ctx->hRot.n + 16));
                    memcpy(ctx->hRot.p + ctx->hRot.n, R.worldRot, 64);
                    ctx->hRot.n += 16;
                }
            }
            if (ctx->hSlim.n % PREP_WINDOW_DRAWS == 0) ctx->blkTmin = ctx->blkTmax = tmplIndex;
            ctx->blkTmin = std::min(ctx->blkTmin, tmplIndex); ctx->blkTmax = std::max(ctx->blkTmax, tmplIndex);
            ctx->tmplSpanMax = std::max(ctx->tmplSpanMax, ctx->blkTmax - ctx->blkTmin + 1);
            CU(ctx->hSlim.push(S));
        }
        // k_vertex<K1_STATIC / K1_SKINNED> (t3d_kernels.cu: shade_vertex): no shape keys, no unbillboarded depth (ignored for
        // skinned draws anyway), a light list the kernel can stage and that holds no volume light
        bool fast = m.nKeys == 0 && ((flags & T3D_DRAW_SKINNED) || !(flags & T3D_DRAW_DEPTH_UNBILLBOARDED));
        if (fast && (flags & T3D_DRAW_LIT)) {
            fast = ll.count <= VERTEX_FAST_MAX_LIGHTS;
            for (uint32_t k = 0; k < ll.count && fast; k++) fast = ctx->hLights.p[ll.off + k].type != T3D_LIGHT_VOLUME;
        }
        ctx->hDraws.push_back(t3d_ctx::HostDraw{R.a0, R.nv, R.vbase, R.nt, R.flags, R.sortMode, (uint8_t)(!fast ? 0 : (flags & T3D_DRAW_SKINNED) ? 2 : 1)});
        ctx->drawTotalTris.push_back(R.nt);
        ctx->drawContinue.push_back((flags & T3D_DRAW_CONTINUE_FLUSH) ? 1 : 0);
    }
    return T3D_OK;
}

// tileFirst[b] = largest d with base[d] <= b*tile (b = 0..nTiles)
static void build_tiles(const uint32_t* base, uint32_t nDraws, uint32_t nTiles, uint32_t tile, uint32_t* out) {
    uint32_t d = 0;
    for (uint32_t b = 0; b <= nTiles; b++) {
        const uint64_t slot = (uint64_t)b * tile;
        while (d + 1 < nDraws && base[d + 1] <= slot) d++;
        out[b] = d;
    }
}

// Cuts the triangle arena into tiles (t3d_device.cuh).  Draws with more than TILE_TRIS triangles get
// exclusive tiles and go through k_cull_large + the segmented counting sort; smaller draws are packed whole
// into tiles of at most TILE_TRIS triangles and TILE_TRIS draws and are culled and sorted by k_tile_packed.
static cudaError_t build_tri_tiles(t3d_ctx* ctx, uint32_t nDraws) {
    ctx->hTiles.clear(); ctx->hLarge.clear(); ctx->hPacked.clear(); ctx->hMedium.clear(); ctx->hHuge.clear(); ctx->hFused.clear();
    const uint32_t* tb = ctx->hTbase.p;
    {   // would k_cull_sort's tiles be resident all at once?
        uint64_t cand = 0;
        for (uint32_t d = 0; d < nDraws; d++) {
            const uint32_t nt = tb[d + 1] - tb[d];
            if (nt > MEDIUM_TRIS && (nt + TILE_TRIS - 1) / TILE_TRIS <= HUGE_TILES) cand += (nt + TILE_TRIS - 1) / TILE_TRIS;
        }
        ctx->frameFused = ctx->fuseSort == 2 || (ctx->fuseSort == 1 && cand <= (uint64_t)ctx->smCount * FUSED_RESIDENT_BLOCKS);
    }
    TriTile cur; memset(&cur, 0, sizeof cur);
    cudaError_t e = cudaSuccess;
    ctx->nHistRows = 0;
    auto close = [&]() -> cudaError_t {
        if (cur.info == 0) return cudaSuccess;
        if (cur.count) {        // a tile of empty draws has nothing to do
            cudaError_t e2;
            if (cur.info == 1) {
                // a draw alone in its tile needs none of the packed kernel's per-slot draw lookup: k_draw_medium takes it
                // (half the time per triangle), and the tile scan / emit see a one-tile "large" draw
                cur.info = TILE_LARGE; cur.row0 = 0; cur.nTilesOfDraw = 1; cur.sortMode = ctx->hDraws[cur.dFirst].sortMode; cur.tbase = cur.start;
                e2 = ctx->hMedium.push((uint32_t)ctx->hTiles.n);
            } else {
                e2 = ctx->hPacked.push((uint32_t)ctx->hTiles.n);
            }
            if (e2 != cudaSuccess) return e2;
            e2 = ctx->hTiles.push(cur);
            if (e2 != cudaSuccess) return e2;
        }
        cur.count = 0; cur.info = 0;
        return cudaSuccess;
    };
    for (uint32_t d = 0; d < nDraws; d++) {
        const uint32_t nt = tb[d + 1] - tb[d];
        if (nt > TILE_TRIS) {
            if ((e = close()) != cudaSuccess) return e;
            const uint32_t n = (nt + TILE_TRIS - 1) / TILE_TRIS;
            const bool medium = nt <= MEDIUM_TRIS;      // one block culls and sorts the whole draw (k_draw_medium)
            if (medium && (e = ctx->hMedium.push((uint32_t)ctx->hTiles.n)) != cudaSuccess) return e;
            if (n > HUGE_TILES && (e = ctx->hHuge.push((uint32_t)ctx->hTiles.n)) != cudaSuccess) return e;
            for (uint32_t j = 0; j < n; j++) {
                TriTile t; t.start = tb[d] + j * TILE_TRIS; t.count = std::min(TILE_TRIS, nt - j * TILE_TRIS); t.dFirst = d; t.info = TILE_LARGE | j;
                t.row0 = ctx->nHistRows; t.nTilesOfDraw = n; t.sortMode = ctx->hDraws[d].sortMode; t.tbase = tb[d];
                if (!medium && (e = ((n > HUGE_TILES || !ctx->frameFused) ? ctx->hLarge : ctx->hFused).push((uint32_t)ctx->hTiles.n)) != cudaSuccess) return e;
                if ((e = ctx->hTiles.push(t)) != cudaSuccess) return e;
            }
            if (!medium) ctx->nHistRows += n + 1;
            continue;
        }
        if (cur.info && (cur.count + nt > TILE_TRIS || cur.info >= TILE_TRIS)) { if ((e = close()) != cudaSuccess) return e; }
        if (cur.info == 0) { cur.start = tb[d]; cur.dFirst = d; }
        cur.count += nt; cur.info++;
    }
    return close();
}

// Sharded delivery, enqueued behind the frame's emit.  A shard ships its per-draw statistics and reports the frame's serial
// and vertex count to the issuing rank's mailbox; the issuing rank rebuilds every shard's part of the lists as it arrives.
static int gather_enqueue(t3d_ctx* ctx, uint32_t nDraws, uint32_t nChunks, bool pdl) {
    t3d_ctx::Gather& G = ctx->gather;
    cudaStream_t s = ctx->stream;
    if (G.rank != G.dst) {
        launch_gather_post(s, pdl, G.mail(G.rank), ctx->dCount.as<uint32_t>(), nChunks, G.serial, ctx->dStats.p, G.stats(G.rank), nDraws,
                           (nDraws && ctx->deliverIndexed && ctx->vAcc <= G.cap) ? ctx->vAcc : 0u);
        ctx->launches += 1;
    } else {
        const bool renderDepth = ctx->fc.camFlags & T3D_CAM_RENDER_DEPTH, pcuv = ctx->fc.camFlags & T3D_CAM_PERSP_CORRECT_UV;
        std::vector<const void*> rec; std::vector<const float*> dsc, cus; std::vector<const uint32_t*> slot;
        std::vector<t3d_vertex*> col, dep; std::vector<uint32_t*> mail;
        for (uint32_t r = 0; r < G.world; r++) {
            if (r == G.dst) continue;
            const uint8_t* in = G.inbox(r);
            rec.push_back(in); dsc.push_back(reinterpret_cast<const float*>(in + G.cap * 32)); cus.push_back(reinterpret_cast<const float*>(in + G.cap * 36));
            slot.push_back(reinterpret_cast<const uint32_t*>(in + G.cap * 40));
            col.push_back(reinterpret_cast<t3d_vertex*>(G.color) + (size_t)r * G.cap); dep.push_back(reinterpret_cast<t3d_vertex*>(G.depth) + (size_t)r * G.cap);
            mail.push_back(G.mail(r));
        }
        if (!rec.empty()) {
            launch_gather_expand(s, pdl, (uint32_t)rec.size(), rec.data(), dsc.data(), cus.data(), slot.data(), col.data(), dep.data(), mail.data(), G.serial, renderDepth, pcuv, G.err());
            ctx->launches += ((int)rec.size() + 14) / 15;
        }
    }
    CU(cudaGetLastError());
    return T3D_OK;
}

int t3d_frame_execute(t3d_ctx* ctx) {
    if (!ctx) return T3D_E_INVALID;
    if (!ctx->inFrame || ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_frame_execute: no open frame");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t nDraws = (uint32_t)ctx->hDraws.size();
    const uint32_t totalV = ctx->vAcc, totalT = ctx->tAcc;
    ctx->launches = 0;
    ctx->executed = true;
    ctx->stagesTimed = ctx->stageTimings;
    ctx->fc.approxLighting = (ctx->exactLighting || ctx->bakeFrame) ? 0u : 1u;
    ctx->fc.bakeValence = nullptr;
    if (ctx->bakeFrame) {
        if (ctx->anySubdiv) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_frame_execute: BakeLighting of an auto-subdivided mesh (its sub-triangles are lit from stale positions in the reference, model.go:751-755)");
        if (ctx->gather.active) return fail(ctx, T3D_E_STATE, "t3d_frame_execute: a bake frame cannot be delivered (t3d_gather_detach first)");
    }
    if (ctx->gather.active) {
        t3d_ctx::Gather& G = ctx->gather;
        G.serial++;
        if (3ull * totalT > G.cap || nDraws > G.drawCap)
            return fail(ctx, T3D_E_NOMEM, "t3d_frame_execute: the frame exceeds the capacity given to t3d_gather_attach (3 vertices per submitted triangle, one entry per draw)");
    }
    if (nDraws == 0 || totalT == 0) {
        for (int i = 0; i < 6; i++) CU(cudaEventRecord(ctx->ev[i], s));
        CU(cudaEventRecord(ctx->ev[9], s));
        if (ctx->gather.active) { const int rc = gather_enqueue(ctx, 0, 0, false); if (rc != T3D_OK) return rc; }     // an empty shard still reports in
        CU(cudaEventRecord(ctx->ev[6], s));
        ctx->h2dBytes = 0;
        return T3D_OK;
    }
    const uint32_t nTilesV = (totalV + VERTEX_TILE - 1) / VERTEX_TILE;
    ctx->hVbase.p[nDraws] = totalV; ctx->hVbase.n = nDraws + 1;
    ctx->hTbase.p[nDraws] = totalT; ctx->hTbase.n = nDraws + 1;
    CU(ctx->hTileV.resize(nTilesV + 1));
    build_tiles(ctx->hVbase.p, nDraws, nTilesV, VERTEX_TILE, ctx->hTileV.p);
    {   // vertex tiles whose draws all belong to one fast class -> that k_vertex instantiation, the rest -> the generic one
        const bool frameFast = !(ctx->fc.camFlags & T3D_CAM_RENDER_NORMALS) && !(ctx->fc.snapping > 0) && !ctx->bakeFrame;
        for (auto& v : ctx->hVClass) { v.clear(); CU(v.reserve(nTilesV)); }
        for (uint32_t b = 0; b < nTilesV; b++) {
            const uint64_t end = std::min<uint64_t>((uint64_t)(b + 1) * VERTEX_TILE, totalV);
            int cls = -1;       // -1: no draw seen yet
            for (uint32_t d = ctx->hTileV.p[b]; cls != 0 && d < nDraws && ctx->hVbase.p[d] < end; d++) {
                if (ctx->hVbase.p[d + 1] == ctx->hVbase.p[d]) continue;        // a draw without vertices
                const int c = frameFast ? ctx->hDraws[d].vclass : 0;
                cls = (cls == -1 || cls == c) ? c : 0;
            }
            auto& list = ctx->hVClass[cls < 0 ? 0 : cls];
            const uint32_t dLo = ctx->hTileV.p[b];
            list.p[list.n++] = make_uint4(b, dLo, ctx->hTileV.p[b + 1], ctx->hVbase.p[dLo + 1]);
        }
    }
    CU(build_tri_tiles(ctx, nDraws));
    const uint32_t nTiles = (uint32_t)ctx->hTiles.n, nLarge = (uint32_t)ctx->hLarge.n, nPacked = (uint32_t)ctx->hPacked.n;
    const uint32_t nMedium = (uint32_t)ctx->hMedium.n, nFused = (uint32_t)ctx->hFused.n;

    const bool normals = ctx->fc.camFlags & T3D_CAM_RENDER_NORMALS;

    // arenas
    CU(ctx->dLightsU.ensure(sizeof(LightU) * std::max<size_t>(ctx->luAcc, 1)));
    CU(ctx->dStats.ensure(sizeof(DrawStat) * nDraws));
    CU(ctx->dTileCtr.ensure(64));
    if (ctx->bakeFrame) { CU(ctx->dValence.ensure(4ull * std::max(totalV, 4u))); ctx->fc.bakeValence = ctx->dValence.as<uint32_t>(); }
    CU(ctx->dClip.ensure(16ull * std::max(totalV, 4u)));
    CU(ctx->dClipXY.ensure(8ull * std::max(totalV, 4u)));
    CU(ctx->dLight.ensure(16ull * std::max(ctx->anyLit ? totalV : 4u, 4u)));
    CU(ctx->dWpos.ensure(16ull * std::max(ctx->anySkinned ? totalV : 4u, 4u)));
    CU(ctx->dUnbill.ensure(4ull * std::max(ctx->anyUnbill ? totalV : 4u, 4u)));
    CU(ctx->dTnrm.ensure(16ull * std::max(normals ? totalV : 4u, 4u)));
    CU(ctx->dTdepth.ensure(4ull * totalT));
    CU(ctx->dSorted.ensure(4ull * totalT));
    CU(ctx->dHist.ensure(4ull * BIN_COUNT * std::max(ctx->nHistRows, 1u)));
    CU(ctx->dTileSurv.ensure(4ull * std::max(nTiles, 1u)));
    CU(ctx->dTileOut.ensure(4ull * (nTiles + 1)));
    CU(ctx->dCount.ensure(4ull * std::max<uint32_t>((nTiles + SCAN_CHUNK - 1) / SCAN_CHUNK, 64u)));     // per-chunk survivor totals of the tile scan
    if (ctx->gather.active && ctx->gather.rank == ctx->gather.dst) {        // the issuing rank emits into its own region of the shared lists
        t3d_ctx::Gather& G = ctx->gather;
        ctx->xColor = reinterpret_cast<t3d_vertex*>(G.color) + (size_t)G.rank * G.cap;
        ctx->xDepth = reinterpret_cast<t3d_vertex*>(G.depth) + (size_t)G.rank * G.cap;
        ctx->xNormal = nullptr; ctx->xCapacity = G.cap;
    }
    const bool external = ctx->xColor != nullptr;
    if (external) {
        if (3ull * totalT > ctx->xCapacity) return fail(ctx, T3D_E_NOMEM, "t3d_frame_execute: the output target is smaller than the frame's worst case (3 vertices per submitted triangle)");
        if (normals && !ctx->xNormal) return fail(ctx, T3D_E_INVALID, "t3d_frame_execute: RenderNormals needs a normal list in the output target");
    }
    // which set of lists this frame writes: alternating while two frames are in flight, set 0 otherwise
    const uint32_t set = (ctx->asyncDownload && !external && !ctx->gather.active) ? ctx->nextSet : 0u;
    ctx->frameSet = set;
    DevBuf &bColor = set ? ctx->dColorB : ctx->dColor, &bDepth = set ? ctx->dDepthB : ctx->dDepth, &bNormal = set ? ctx->dNormalB : ctx->dNormal,
           &bTriIds = set ? ctx->dTriIdsB : ctx->dTriIds;
    if (!external) {
        CU(bColor.ensure(144ull * totalT));
        CU(bDepth.ensure(144ull * totalT));
        CU(bNormal.ensure(normals ? 144ull * totalT : 16));
    }
    ctx->oColor = external ? ctx->xColor : bColor.as<t3d_vertex>();
    ctx->oDepth = external ? ctx->xDepth : bDepth.as<t3d_vertex>();
    ctx->oNormal = external ? ctx->xNormal : bNormal.as<t3d_vertex>();
    CU(bTriIds.ensure(4ull * totalT));
    ctx->oTriIds = bTriIds.as<uint32_t>();
    if (ctx->copyPending[set]) CU(cudaStreamWaitEvent(s, ctx->evCopyDone[set], 0));     // the set's previous lists are still leaving: the kernels queue up behind that copy
    if (ctx->volumesDirty) {
        CU(ctx->dVolumes.ensure(sizeof(VolumeU) * std::max<size_t>(ctx->volumes.size(), 1)));
        if (!ctx->volumes.empty())
            CU(cudaMemcpyAsync(ctx->dVolumes.p, ctx->volumes.data(), sizeof(VolumeU) * ctx->volumes.size(), cudaMemcpyHostToDevice, s));
        ctx->volumesDirty = false;
    }

    // uniforms: pack every per-frame array into one pinned, mapped block that k_prep_draws reads over PCIe
    {
        // host block: [DrawSlim[D] | templates | rotations | tables ... ]; device block: [DrawRec[D] | tables ...]
        struct Seg { const void* src; size_t bytes; size_t off; };
        // Runs of consecutive draws whose headers progress regularly (instances of one MeshPart).  When they are long on
        // average, the frame ships run records + 64-byte payloads instead of one 96-byte DrawSlim per draw.
        ctx->hRuns.clear(); ctx->hBlockRun.clear();
        bool runMode = false;
        if (ctx->transformsAffine && nDraws >= 1024) {
            const DrawSlim* S = ctx->hSlim.p;
            for (uint32_t d = 0; d < nDraws; d++) {
                bool extend = false;
                if (!ctx->hRuns.empty()) {
                    DrawRun& U = ctx->hRuns.back();
                    const uint32_t i = d - U.first;
                    if (S[d].tmpl == U.tmpl && ((S[d].rot != NO_ROTATION) == (U.rot != NO_ROTATION)) && S[d].pad[0] == U.pad[0]) {
                        if (U.count == 1) {      // the second draw of a run fixes the increments
                            U.dV = S[d].vbase - U.vbase; U.dT = S[d].tbase - U.tbase; U.dBone = S[d].boneOffset - U.boneOffset;
                            U.dLightU = S[d].lightU - U.lightU; U.dRot = U.rot != NO_ROTATION ? S[d].rot - U.rot : 0;
                            extend = U.dRot <= 1;       // (0: instances sharing one world rotation)
                        } else {
                            extend = S[d].vbase == U.vbase + i * U.dV && S[d].tbase == U.tbase + i * U.dT && S[d].boneOffset == U.boneOffset + i * U.dBone &&
                                     S[d].lightU == U.lightU + i * U.dLightU && (U.rot == NO_ROTATION || S[d].rot == U.rot + i * U.dRot);
                        }
                    }
                    if (extend) U.count++;
                    else if (U.count == 1) { U.dV = U.dT = U.dBone = U.dLightU = U.dRot = 0; }
                }
                if (!extend) {
                    DrawRun U; memset(&U, 0, sizeof U);
                    U.first = d; U.count = 1; U.tmpl = S[d].tmpl; U.vbase = S[d].vbase; U.tbase = S[d].tbase; U.boneOffset = S[d].boneOffset;
                    U.lightU = S[d].lightU; U.rot = S[d].rot;
                    U.pad[0] = S[d].pad[0];         // t3d_submit_instances group + 1: k_prep_draws builds the run's transforms from position and scale
                    if (U.pad[0]) { const t3d_ctx::TrsGroup& G = ctx->trsGroups[(U.pad[0] & 0x7FFFFFFFu) - 1]; U.pad[1] = G.rot; U.pad[2] = G.parent; }
                    ctx->hRuns.push_back(U);
                }
            }
            runMode = (uint64_t)ctx->hRuns.size() * RUN_MODE_MIN_AVG <= nDraws;
            if (runMode) {      // hBlockRun[b] = run of draw b * 64 (k_prep_draws: 64 draws per block), one extra entry for the last draw
                const uint32_t nBlocks = (nDraws + PREP_WINDOW_DRAWS - 1) / PREP_WINDOW_DRAWS;
                ctx->hBlockRun.resize(nBlocks + 1);
                uint32_t r = 0;
                for (uint32_t b = 0; b <= nBlocks; b++) {
                    const uint32_t d = std::min(b * (uint32_t)PREP_WINDOW_DRAWS, nDraws - 1);
                    while (r + 1 < ctx->hRuns.size() && ctx->hRuns[r + 1].first <= d) r++;
                    ctx->hBlockRun[b] = r;
                }
            }
        }
        ctx->frameRunMode = runMode;
        // Instances (t3d_submit_instances) travel as position + scale + colour when the frame is shipped as runs: 32 bytes each if
        // the frame holds nothing else and every scale is uniform, else in the 64-byte payload slot of an ordinary draw.  A
        // frame that is not shipped as runs gets their transforms from the host: the same sequence of operations (instance_transform).
        const bool trsOnly = runMode && ctx->nTrsDraws == nDraws && ctx->trsUniform;
        if (!runMode && ctx->nTrsDraws) {
            for (uint32_t d = 0; d < nDraws; d++) {
                DrawSlim& S = ctx->hSlim.p[d];
                if (!S.pad[0]) continue;
                const t3d_ctx::TrsGroup& G = ctx->trsGroups[(S.pad[0] & 0x7FFFFFFFu) - 1];
                float ps[6]; memcpy(ps, S.transform, 24);
                const Mat4 M = instance_transform(ps, ps + 3, G.rotation, G.hasParent ? &G.parentT : nullptr);
                m4_store(M, S.transform);
                S.pad[0] = 0;
                if (G.hasBounds) {      // the culling sphere test, as k_prep_draws runs it for frames shipped as runs; 1 = outside
                    const FrameConst& fc = ctx->fc;
                    FrustumTerms F;
                    F.camPos = v3(fc.camPos[0], fc.camPos[1], fc.camPos[2]); F.forward = v3(fc.camForward[0], fc.camForward[1], fc.camForward[2]);
                    F.right = v3(fc.camRight[0], fc.camRight[1], fc.camRight[2]); F.up = v3(fc.camUp[0], fc.camUp[1], fc.camUp[2]);
                    F.nearP = fc.nearP; F.farP = fc.farP; F.sphereFactorY = fc.sphereFactorY; F.sphereFactorX = fc.sphereFactorX; F.sphereFactorTang = fc.sphereFactorTang;
                    F.aspect = fc.camW / fc.camH; F.orthoScale = fc.orthoScale; F.perspective = (fc.camFlags & T3D_CAM_PERSPECTIVE) != 0;
                    if (!m4_instance_in_frustum(M, G.bounds, F)) S.pad[0] = 1;
                }
            }
            ctx->nTrsDraws = 0;
        }
        const size_t slimBytes = runMode ? (trsOnly ? 32 : 64) : ctx->transformsAffine ? sizeof(DrawSlim) - 16 : sizeof(DrawSlim);    // affine transforms travel as 12 floats
        Seg head[6] = {{nullptr /* packed below */, slimBytes * (size_t)nDraws, 0}, {ctx->hTmpl.p, sizeof(DrawRec) * ctx->hTmpl.n, 0},
                       {ctx->hRot.p, 4ull * ctx->hRot.n, 0},
                       {ctx->hRuns.data(), runMode ? sizeof(DrawRun) * ctx->hRuns.size() : 0, 0},
                       {ctx->hBlockRun.data(), runMode ? 4ull * ctx->hBlockRun.size() : 0, 0},
                       {ctx->hPoses.p, sizeof(t3d_pose) * ctx->hPoses.n, 0}};       // read by k_pose_bones
        Seg segs[14] = {
            {nullptr, 0, 0}, {nullptr, 0, 0} /* vbase / tbase: k_prep_draws writes the device tables from the placement records */,
            {nullptr, 0, 0} /* the vertex-tile table stays on the host: its entries travel inside the class lists */, {ctx->hTiles.p, sizeof(TriTile) * (size_t)nTiles, 0},
            {ctx->hLarge.p, 4ull * nLarge, 0}, {ctx->hPacked.p, 4ull * nPacked, 0},
            {ctx->hLights.p, sizeof(t3d_light) * ctx->hLights.n, 0}, {nullptr, 0, 0} /* bones travel separately, below */,
            {ctx->hMedium.p, 4ull * nMedium, 0}, {ctx->hVClass[0].p, 16ull * ctx->hVClass[0].n, 0}, {ctx->hVClass[1].p, 16ull * ctx->hVClass[1].n, 0},
            {ctx->hHuge.p, 4ull * ctx->hHuge.n, 0}, {ctx->hVClass[2].p, 16ull * ctx->hVClass[2].n, 0}, {ctx->hFused.p, 4ull * nFused, 0}};
        size_t headBytes = 0, restBytes = 0;
        for (Seg& g : head) { g.off = headBytes; headBytes += (g.bytes + 255) & ~(size_t)255; }
        for (Seg& g : segs) { g.off = restBytes; restBytes += (g.bytes + 255) & ~(size_t)255; }
        const size_t devDrawBytes = (sizeof(DrawRec) * (size_t)nDraws + 255) & ~(size_t)255;
        // bone matrices: 12 floats each on the host side when all are affine, always 16 on the device
        const size_t nBones = ctx->hBones.n / 16;
        const size_t hostBoneBytes = ((ctx->bonesAffine ? 48 : 64) * nBones + 255) & ~(size_t)255;
        const size_t devBoneBytes = (64 * (nBones + ctx->posedBones) + 255) & ~(size_t)255;      // submitted matrices, then the device-evaluated poses
        const size_t devTableBytes = (4ull * (nDraws + 1) + 255) & ~(size_t)255;        // vbase, tbase
        ctx->hFrame.mapped = true;
        CU(ctx->hFrame.resize(headBytes + restBytes + hostBoneBytes));
        CU(ctx->dFrame.ensure(devDrawBytes + restBytes + devBoneBytes + 2 * devTableBytes));
        for (Seg& g : head) if (g.bytes && g.src) memcpy(ctx->hFrame.p + g.off, g.src, g.bytes);
        for (Seg& g : segs) if (g.bytes) memcpy(ctx->hFrame.p + headBytes + g.off, g.src, g.bytes);
        if (!ctx->transformsAffine) memcpy(ctx->hFrame.p, ctx->hSlim.p, sizeof(DrawSlim) * (size_t)nDraws);
        else if (trsOnly) for (uint32_t d = 0; d < nDraws; d++) {       // {position, scale | colour}
            const DrawSlim& S = ctx->hSlim.p[d];
            uint8_t* o = ctx->hFrame.p + 32 * (size_t)d;
            memcpy(o, S.transform, 16); memcpy(o + 16, S.color, 16);
        }
        else for (uint32_t d = 0; d < nDraws; d++) {
            const DrawSlim& S = ctx->hSlim.p[d];
            uint8_t* o = ctx->hFrame.p + slimBytes * (size_t)d - (runMode ? 32 : 0);      // run mode: no header, payload only
            if (!runMode) memcpy(o, &S, 32);                            // header
            float* t = reinterpret_cast<float*>(o + 32);
            if (runMode && S.pad[0]) { memcpy(t, S.transform, 24); memset(t + 6, 0, 24); }        // an instance: {position, scale x | scale y, z, - , - | - }
            else for (int r = 0; r < 4; r++) { t[3 * r] = S.transform[4 * r]; t[3 * r + 1] = S.transform[4 * r + 1]; t[3 * r + 2] = S.transform[4 * r + 2]; }
            memcpy(o + 32 + 48, S.color, 16);
        }
        {
            float* hb = reinterpret_cast<float*>(ctx->hFrame.p + headBytes + restBytes);
            if (!ctx->bonesAffine) memcpy(hb, ctx->hBones.p, 64 * nBones);
            else for (size_t i = 0; i < nBones; i++) {
                const float* M = ctx->hBones.p + 16 * i;
                float* o = hb + 12 * i;
                o[0] = M[0]; o[1] = M[1]; o[2] = M[2]; o[3] = M[4]; o[4] = M[5]; o[5] = M[6];
                o[6] = M[8]; o[7] = M[9]; o[8] = M[10]; o[9] = M[12]; o[10] = M[13]; o[11] = M[14];
            }
        }
        ctx->frameBonesOff = headBytes + restBytes; ctx->frameBoneCount = (uint32_t)nBones;
        // no DMA copy: k_prep_draws reads the pinned block through UVA and writes the device copy
        CU(cudaHostGetDevicePointer(&ctx->hFrameDev, ctx->hFrame.p, 0));
        // ... unless the block is large: zero-copy reads from the SMs top out at ~40 GB/s, the copy engine reaches ~55, so from
        // about a megabyte on the block is copied into device memory first and k_prep_draws reads that copy instead
        ctx->frameBlockBytes = headBytes + restBytes + hostBoneBytes; ctx->frameSlimBytes = slimBytes;
        ctx->frameStaged = ctx->frameBlockBytes >= ctx->dmaThreshold;
        if (ctx->frameStaged) CU(ctx->dStage.ensure(ctx->frameBlockBytes));
        ctx->frameHeadBytes = headBytes; ctx->frameRestBytes = restBytes; ctx->frameTmplOff = head[1].off; ctx->frameRotOff = head[2].off;
        ctx->frameRunsOff = head[3].off; ctx->frameBlockRunOff = head[4].off; ctx->framePosesOff = head[5].off;
        ctx->frameLightsOff = headBytes + segs[6].off;
        uint8_t* base = ctx->dFrame.as<uint8_t>();
        uint8_t* rest = base + devDrawBytes;
        ctx->pDevRest = rest;
        ctx->pDraws = reinterpret_cast<DrawRec*>(base);
        ctx->pVbase = reinterpret_cast<uint32_t*>(rest + restBytes + devBoneBytes);
        ctx->pTbase = reinterpret_cast<uint32_t*>(rest + restBytes + devBoneBytes + devTableBytes);
        ctx->pTiles = reinterpret_cast<TriTile*>(rest + segs[3].off);
        ctx->pLarge = reinterpret_cast<uint32_t*>(rest + segs[4].off);
        ctx->pPacked = reinterpret_cast<uint32_t*>(rest + segs[5].off);
        ctx->pLightsIn = reinterpret_cast<t3d_light*>(rest + segs[6].off);
        ctx->pBones = reinterpret_cast<float*>(rest + restBytes);
        ctx->pMedium = reinterpret_cast<uint32_t*>(rest + segs[8].off);
        ctx->pVClass[0] = reinterpret_cast<uint4*>(rest + segs[9].off);
        ctx->pVClass[1] = reinterpret_cast<uint4*>(rest + segs[10].off);
        ctx->pHuge = reinterpret_cast<uint32_t*>(rest + segs[11].off);
        ctx->pVClass[2] = reinterpret_cast<uint4*>(rest + segs[12].off);
        ctx->pFused = reinterpret_cast<uint32_t*>(rest + segs[13].off);
        ctx->h2dBytes = 0;
        for (Seg& g : head) ctx->h2dBytes += g.bytes;
        for (Seg& g : segs) ctx->h2dBytes += g.bytes;
        ctx->h2dBytes += (ctx->bonesAffine ? 48 : 64) * nBones;
    }
    // Stage timings need an event between kernels, which also serialises them; the default path records only
    // the frame's first and last event and chains the kernels with programmatic dependent launches.
    const bool st = ctx->stageTimings;
    const bool pdl = !st;
    auto mark = [&](int i) -> cudaError_t { return st ? cudaEventRecord(ctx->ev[i], s) : cudaSuccess; };
    CU(cudaEventRecord(ctx->ev[0], s));
    CU(mark(1));

    DrawStat* stats = ctx->dStats.as<DrawStat>();
    // K0 (fetches the uniforms over PCIe itself) + K1
    {
        // template / rotation index windows of the first blocks of k_prep_draws (64 draws each)
        uint32_t win[PREP_WINDOW_BLOCKS][4];
        for (int b = 0; b < PREP_WINDOW_BLOCKS; b++) {
            uint32_t tmin = 0xFFFFFFFFu, tmax = 0, rmin = NO_ROTATION, rmax = 0;
            const uint32_t lo = (uint32_t)b * PREP_WINDOW_DRAWS, hi = std::min(lo + PREP_WINDOW_DRAWS, nDraws);
            for (uint32_t d = lo; d < hi; d++) {
                const DrawSlim& S = ctx->hSlim.p[d];
                tmin = std::min(tmin, S.tmpl); tmax = std::max(tmax, S.tmpl);
                if (S.rot != NO_ROTATION) { rmin = std::min(rmin, S.rot); rmax = std::max(rmax, S.rot); }
            }
            win[b][0] = lo < hi ? tmin : 0; win[b][1] = lo < hi ? tmax - tmin + 1 : 0;
            win[b][2] = rmin; win[b][3] = rmin != NO_ROTATION ? rmax - rmin + 1 : 0;
        }
        const uint8_t* hf = static_cast<const uint8_t*>(ctx->hFrameDev);
        // Very large staged frames are prepared in chunks of PREP_CHUNK_BLOCKS blocks: the placement records are copied slice by
        // slice on a second stream and every slice's k_prep_draws launch waits only for its own slice, so the copy engine runs
        // back to back while the kernels work on what has arrived (2 M particles: 128 MB of records).
        const uint32_t nPrepBlocks = (nDraws + PREP_WINDOW_DRAWS - 1) / PREP_WINDOW_DRAWS;
        // slice size: an eighth of the frame, between 1024 and 4096 blocks (64 K .. 256 K draws) -- small enough to overlap, large
        // enough that the per-slice launch and event cost stays invisible (measured on 200 K and 2 M particle frames)
        const uint32_t PREP_CHUNK_BLOCKS = ctx->prepChunkBlocks ? ctx->prepChunkBlocks : std::min(4096u, std::max(1024u, nPrepBlocks / 8));
        const bool chunked = ctx->frameStaged && nPrepBlocks >= 3 * PREP_CHUNK_BLOCKS;
        if (ctx->frameStaged) {
            hf = ctx->dStage.as<uint8_t>();
            if (!chunked) CU(cudaMemcpyAsync(ctx->dStage.p, ctx->hFrame.p, ctx->frameBlockBytes, cudaMemcpyHostToDevice, s));
            else {      // everything but the placement records (templates, rotations, runs, tables, lights, bones) goes first
                const size_t off1 = ctx->frameTmplOff;
                CU(cudaMemcpyAsync(ctx->dStage.as<uint8_t>() + off1, ctx->hFrame.p + off1, ctx->frameBlockBytes - off1, cudaMemcpyHostToDevice, s));
                if (!ctx->copyStream) { CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming)); }
                CU(cudaEventRecord(ctx->evFork, s));
                CU(cudaStreamWaitEvent(ctx->copyStream, ctx->evFork, 0));
            }
        }
        // bone poses evaluated on the device: one launch per t3d_submit_poses call, into the arena behind the submitted matrices
        PoseBatch fused[PREP_POSE_BATCHES]; uint32_t nFused = 0;
        for (const PoseCall& c : ctx->poseCalls) {
            const Skeleton& S = ctx->skeletons.at(c.skeleton);
            PoseBatch pb; memset(&pb, 0, sizeof pb);
            pb.nBones = S.nBones; pb.maxDepth = S.maxDepth; pb.nPoses = c.nPoses;
            pb.parent = S.parent; pb.depth = S.depth; pb.invBind = S.invBind; pb.restPos = S.restPos; pb.restScale = S.restScale; pb.restRot = S.restRot;
            if (c.animation) {
                const Animation& A = ctx->animations.at(c.animation);
                pb.animBones = A.nBones; pb.tracks = A.tracks; pb.keyTimes = A.keyTimes; pb.keyValues = A.keyValues;
            }
            pb.poses = reinterpret_cast<const t3d_pose*>(hf + ctx->framePosesOff) + c.firstPose;
            pb.out = ctx->pBones + 16 * ((size_t)ctx->frameBoneCount + c.firstBone);
            // small skeletons (every ordinary character) ride as extra blocks of k_prep_draws: no launch of their own, and
            // their PCIe round trip and double-precision Slerp overlap the draw records' fetch
            if (c.nLevels) {      // blends interrupted by further Play() calls: k_pose_chain reads its 96-byte records from pinned memory
                PoseChain ch; memset(&ch, 0, sizeof ch);
                ch.nLevels = c.nLevels;
                for (uint32_t k = 0; k < c.nLevels; k++) {
                    const Animation& A = ctx->animations.at(c.chain[k]);
                    ch.level[k].bones = A.nBones; ch.level[k].tracks = A.tracks; ch.level[k].keyTimes = A.keyTimes; ch.level[k].keyValues = A.keyValues;
                }
                pb.poses = reinterpret_cast<const t3d_pose*>(ctx->hPoseChains.p + c.firstPose);
                CU(launch_pose_chain(s, pb, ch)); ctx->launches += 1;
            }
            else if (c.blended) {      // players in the middle of a blend: k_pose_bones<true>, a launch of its own (transitions are short and rare)
                PoseBlend bl; memset(&bl, 0, sizeof bl);
                if (c.fromAnimation) {
                    const Animation& F = ctx->animations.at(c.fromAnimation);
                    bl.fromBones = F.nBones; bl.tracks = F.tracks; bl.keyTimes = F.keyTimes; bl.keyValues = F.keyValues;
                }
                CU(launch_pose_bones(s, pb, &bl)); ctx->launches += 1;
            }
            else if (nFused < PREP_POSE_BATCHES && pose_batch_smem(pb.nBones) <= PREP_POSE_SMEM) fused[nFused++] = pb;
            else { CU(launch_pose_bones(s, pb)); ctx->launches += 1; }
        }
        for (uint32_t b0 = 0, k = 0; b0 < nPrepBlocks; k++) {
            const uint32_t b1 = chunked ? std::min(b0 + PREP_CHUNK_BLOCKS, nPrepBlocks) : nPrepBlocks;
            if (chunked) {
                const size_t lo = (size_t)b0 * PREP_WINDOW_DRAWS * ctx->frameSlimBytes;
                const size_t hi = (size_t)std::min<uint64_t>((uint64_t)b1 * PREP_WINDOW_DRAWS, nDraws) * ctx->frameSlimBytes;
                CU(cudaMemcpyAsync(ctx->dStage.as<uint8_t>() + lo, ctx->hFrame.p + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copyStream));
                if (ctx->evChunk.size() <= k) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->evChunk.push_back(e); }
                CU(cudaEventRecord(ctx->evChunk[k], ctx->copyStream));
                CU(cudaStreamWaitEvent(s, ctx->evChunk[k], 0));
            }
            launch_prep_draws(s, pdl, hf, hf + ctx->frameTmplOff, hf + ctx->frameRotOff, hf + ctx->frameHeadBytes,
                              reinterpret_cast<const t3d_light*>(hf + ctx->frameLightsOff), ctx->pDraws, ctx->pDevRest, nDraws, ctx->frameRestBytes,
                              (uint32_t)ctx->hLights.n, ctx->fc, ctx->dLightsU.as<LightU>(), stats, win,
                              hf + ctx->frameBonesOff, ctx->pBones, ctx->frameBoneCount, ctx->bonesAffine, ctx->transformsAffine,
                              ctx->pVbase, ctx->pTbase, ctx->frameRunMode ? hf + ctx->frameRunsOff : nullptr, (uint32_t)(ctx->frameSlimBytes / 16),
                              reinterpret_cast<const uint32_t*>(hf + ctx->frameBlockRunOff), ctx->tmplSpanMax, fused, nFused, b0, b1 - b0, b0 == 0, ctx->dTileCtr.as<uint32_t>());
            if (b0) ctx->launches += 1;
            b0 = b1;
        }
    }
    CU(mark(2));
    const uint32_t nVClass[3] = {(uint32_t)ctx->hVClass[0].n, (uint32_t)ctx->hVClass[1].n, (uint32_t)ctx->hVClass[2].n};
    if (ctx->bakeFrame) {       // corners per vertex, from the records k_prep_draws just wrote
        CU(cudaMemsetAsync(ctx->dValence.p, 0, 4ull * totalV, s));
        launch_bake_valence(s, ctx->pDraws, ctx->pTbase, nDraws, totalT, ctx->dValence.as<uint32_t>());
        ctx->launches += 1;
    }
    launch_vertex(s, pdl && !ctx->bakeFrame, ctx->exactLighting || ctx->bakeFrame, ctx->pDraws, ctx->pVbase, ctx->pVClass, nVClass, ctx->dTileCtr.as<uint32_t>(), totalV, ctx->fc,
                  ctx->dLightsU.as<LightU>(), ctx->pBones, ctx->dVolumes.as<VolumeU>(), ctx->dClip.as<float4>(), ctx->dClipXY.as<float2>(),
                  ctx->dLight.as<float4>(), ctx->dWpos.as<float4>(), ctx->dUnbill.as<float>(), ctx->dTnrm.as<float4>());
    ctx->launches += 1 + (nVClass[0] ? 1 : 0) + (nVClass[1] ? 1 : 0) + (nVClass[2] ? 1 : 0);
    CU(mark(3));
    if (ctx->bakeFrame) {       // a bake frame ends here: its product is the per-vertex light sums (t3d_frame_read_vertex_lights)
        for (int i = 4; i <= 5; i++) CU(mark(i));
        CU(cudaEventRecord(ctx->ev[9], s)); CU(cudaEventRecord(ctx->ev[6], s));
        CU(cudaGetLastError());
        ctx->vertexKernelBytes = 0;
        return T3D_OK;
    }
    // K2: cull (small and medium draws are also sorted here; so are the large ones up to HUGE_TILES tiles, by k_cull_sort,
    // whose extra block also computes where every tile starts in the lists when the frame's tiles fit one scan chunk)
    launch_tile_packed(s, pdl, ctx->pDraws, ctx->pTbase, ctx->pTiles, ctx->pPacked, nPacked, ctx->fc, ctx->dClip.as<float4>(),
                       ctx->dWpos.as<float4>(), ctx->dSorted.as<uint32_t>(), ctx->dTileSurv.as<uint32_t>(), stats, ctx->anySubdiv);
    launch_draw_medium(s, pdl, ctx->pDraws, ctx->pTiles, ctx->pMedium, nMedium, ctx->fc, ctx->dClip.as<float4>(), ctx->dWpos.as<float4>(),
                       ctx->dSorted.as<uint32_t>(), stats, ctx->anySubdiv);
    launch_cull_large(s, pdl, ctx->pDraws, ctx->pTiles, ctx->pLarge, nLarge, ctx->fc, ctx->dClip.as<float4>(), ctx->dWpos.as<float4>(),
                      ctx->dTdepth.as<uint32_t>(), stats, ctx->anySubdiv);
    const bool oneChunk = nTiles <= SCAN_CHUNK;
    const bool scanInFused = nFused && oneChunk;
    launch_cull_sort(s, pdl, ctx->pDraws, ctx->pTiles, ctx->pFused, nFused, ctx->dTileCtr.as<uint32_t>() + 4, ctx->fc, ctx->dClip.as<float4>(), ctx->dWpos.as<float4>(),
                     stats, ctx->dHist.as<uint32_t>(), ctx->dSorted.as<uint32_t>(), ctx->anySubdiv, scanInFused ? nTiles : 0u,
                     ctx->dTileSurv.as<uint32_t>(), ctx->dTileOut.as<uint32_t>(), ctx->dCount.as<uint32_t>());
    ctx->launches += (nLarge ? 1 : 0) + (nPacked ? 1 : 0) + (nMedium ? 1 : 0) + (nFused ? 1 : 0);
    CU(mark(4));
    // K3: where every tile starts in the lists (needs only the survivor counts), then the segmented 512-bin
    // counting sort of the draws of more than HUGE_TILES tiles
    const bool scanRides = !scanInFused && nLarge && oneChunk;     // one chunk: the scan rides along as k_bin_hist's extra block
    if (!scanInFused && !scanRides) {
        launch_tile_scan(s, pdl, ctx->pTiles, nTiles, stats, ctx->dTileSurv.as<uint32_t>(), ctx->dTileOut.as<uint32_t>(), ctx->dCount.as<uint32_t>());
        ctx->launches += 1;
    }
    if (nLarge) {
        launch_bin_hist(s, pdl, ctx->pTiles, ctx->pLarge, nLarge, scanRides ? nTiles : 0xFFFFFFFFu, stats, ctx->dTdepth.as<uint32_t>(), ctx->dHist.as<uint32_t>(),
                        ctx->dTileSurv.as<uint32_t>(), ctx->dTileOut.as<uint32_t>(), ctx->dCount.as<uint32_t>());
        launch_hist_scan_huge(s, pdl, ctx->pTiles, ctx->pHuge, (uint32_t)ctx->hHuge.n, ctx->dHist.as<uint32_t>());
        launch_scatter(s, pdl, ctx->pTiles, ctx->pLarge, nLarge, ctx->dTdepth.as<uint32_t>(), ctx->dHist.as<uint32_t>(), ctx->dSorted.as<uint32_t>());
        ctx->launches += 2 + (ctx->hHuge.n ? 1 : 0);
    }
    CU(mark(5));
    // K4
    t3d_ctx::Gather& G = ctx->gather;
    const bool shard = G.active && G.rank != G.dst;        // this rank's lists live on the issuing rank: compact emit into its inbox there
    if (G.active && normals) return fail(ctx, T3D_E_UNSUPPORTED, "t3d_frame_execute: sharded delivery with RenderNormals");
    const uint32_t nScanChunks = (nTiles + SCAN_CHUNK - 1) / SCAN_CHUNK;
    const bool offerIndexed = shard && ctx->deliverIndexed && totalV <= G.cap;
    if (shard) {
        CU(ctx->dGatherErr.ensure(64));
        launch_gather_wait_ack(s, pdl, G.mail(G.rank) + 16, G.serial - 1, ctx->dGatherErr.as<uint32_t>());      // the previous frame has been expanded
        uint8_t* in = G.inbox(G.rank);
        launch_emit(s, pdl, ctx->pDraws, ctx->pTiles, nTiles, ctx->dSorted.as<uint32_t>(), ctx->dTileSurv.as<uint32_t>(), ctx->dTileOut.as<uint32_t>(),
                    ctx->dCount.as<uint32_t>(), ctx->fc, ctx->dClip.as<float4>(), ctx->dClipXY.as<float2>(), ctx->dLight.as<float4>(), ctx->dUnbill.as<float>(), ctx->dTnrm.as<float4>(),
                    reinterpret_cast<float4*>(in), reinterpret_cast<float4*>(in + G.cap * 32), nullptr, ctx->oTriIds, true,
                    reinterpret_cast<float*>(in + G.cap * 36), offerIndexed ? totalV : 0u, nScanChunks);
        // the indexed form (one record per arena vertex + a slot per emitted corner); each of the three kernels decides on the
        // device, from the survivor total, which form this frame takes, and the other form's kernels return at once
        if (offerIndexed) {
            launch_emit_indexed(s, pdl, ctx->pDraws, ctx->pVbase, nDraws, totalV, ctx->pTiles, nTiles, ctx->dSorted.as<uint32_t>(), ctx->dTileSurv.as<uint32_t>(),
                                ctx->dTileOut.as<uint32_t>(), ctx->dCount.as<uint32_t>(), nScanChunks, ctx->fc, ctx->dClip.as<float4>(), ctx->dClipXY.as<float2>(),
                                ctx->dLight.as<float4>(), ctx->dUnbill.as<float>(), reinterpret_cast<float4*>(in), reinterpret_cast<float*>(in + G.cap * 32),
                                reinterpret_cast<float*>(in + G.cap * 36), reinterpret_cast<uint32_t*>(in + G.cap * 40), ctx->oTriIds);
            ctx->launches += 2;
        }
        ctx->launches += 1;
    } else {
        launch_emit(s, pdl, ctx->pDraws, ctx->pTiles, nTiles, ctx->dSorted.as<uint32_t>(), ctx->dTileSurv.as<uint32_t>(), ctx->dTileOut.as<uint32_t>(),
                    ctx->dCount.as<uint32_t>(), ctx->fc, ctx->dClip.as<float4>(), ctx->dClipXY.as<float2>(), ctx->dLight.as<float4>(), ctx->dUnbill.as<float>(), ctx->dTnrm.as<float4>(),
                    reinterpret_cast<float4*>(ctx->oColor), reinterpret_cast<float4*>(ctx->oDepth),
                    normals ? reinterpret_cast<float4*>(ctx->oNormal) : nullptr, ctx->oTriIds, false, nullptr, 0u, 0u);
    }
    ctx->launches += nTiles ? 1 : 0;
    CU(cudaEventRecord(ctx->ev[9], s));
    if (G.active) { const int rc = gather_enqueue(ctx, nDraws, ((uint32_t)ctx->hTiles.n + SCAN_CHUNK - 1) / SCAN_CHUNK, pdl); if (rc != T3D_OK) return rc; }
    CU(cudaEventRecord(ctx->ev[6], s));
    CU(cudaGetLastError());

    // algorithmic bytes of the fused vertex kernel (SURVEY.md 8d / DESIGN.md): static unlit 28 B/vertex
    // (pos 12 -> clip 16), static lit 52 (+ normal 12, light rgb 12), skinned lit 88 (+ weights 16, joints 8,
    // world position 12), skinned unlit 76 (normals are still skinned).
    uint64_t bytes = 0;
    for (uint32_t i = 0; i < nDraws; i++) {
        const t3d_ctx::HostDraw& R = ctx->hDraws[i];
        const bool lit = R.flags & T3D_DRAW_LIT, sk = R.flags & T3D_DRAW_SKINNED;
        uint64_t per = 12 + 16;
        if (lit) per += 12 + 12;
        if (sk) per += (lit ? 0 : 12) + 16 + 8 + 12;
        bytes += per * R.nv;
    }
    ctx->vertexKernelBytes = bytes;
    return T3D_OK;
}

// Per-draw statistics and scan totals stored straight into pinned host memory by the SMs.  With two frames in flight the copy
// engine is busy with the previous frame's lists for milliseconds, and a small cudaMemcpyAsync of this frame's statistics would
// queue behind them -- t3d_frame_results would then return only after the previous frame's lists had arrived, and the next copy
// could not follow the previous one without a gap.
__global__ void __launch_bounds__(256)
k_publish_stats(const uint4* __restrict__ stats, uint32_t nDraws, const uint32_t* __restrict__ count, uint32_t nChunks,
                uint4* __restrict__ hostStats, uint32_t* __restrict__ hostCount) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < nDraws) hostStats[i] = stats[i];
    if (i < nChunks) hostCount[i] = count[i];
}

static int finish_frame(t3d_ctx* ctx, bool download, t3d_frame_output* out, t3d_timings* tm) {
    if (!ctx || !out) return fail(ctx, T3D_E_INVALID, "t3d_frame_results: NULL argument");
    if (!ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_frame_results: frame not executed");
    if (ctx->bakeFrame) return fail(ctx, T3D_E_STATE, "t3d_frame_results: a bake frame (t3d_frame_set_bake_lighting) emits no lists: read its vertex lights with t3d_frame_read_vertex_lights");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t nDraws = (uint32_t)ctx->hDraws.size();
    const uint32_t totalT = ctx->tAcc;
    const bool normals = ctx->fc.camFlags & T3D_CAM_RENDER_NORMALS;
    memset(out, 0, sizeof *out);
    out->n_draws = nDraws;
    PinVec<t3d_draw_result>& hRes = (download && ctx->frameSet) ? ctx->hResultsB : ctx->hResults;
    CU(hRes.resize(nDraws));
    ctx->downloadMs = 0;
    ctx->d2hBytes = 0;
    uint64_t nVerts = 0;
    if (nDraws && totalT) {
        ctx->hStats.mapped = true; ctx->hCount.mapped = true;       // k_publish_stats stores into them
        CU(ctx->hStats.resize(nDraws));
        const uint32_t nChunks = ((uint32_t)ctx->hTiles.n + SCAN_CHUNK - 1) / SCAN_CHUNK;
        CU(ctx->hCount.resize(std::max(nChunks, 1u)));
        if (ctx->asyncDownload && (ctx->copyPending[0] || ctx->copyPending[1])) {
            static_assert(sizeof(DrawStat) == sizeof(uint4), "k_publish_stats moves a DrawStat as one 16-byte word");
            const uint32_t nc = std::max(nChunks, 1u), n = std::max(nDraws, nc);
            k_publish_stats<<<(n + 255) / 256, 256, 0, s>>>(ctx->dStats.as<uint4>(), nDraws, ctx->dCount.as<uint32_t>(), nc,
                                                          reinterpret_cast<uint4*>(ctx->hStats.p), ctx->hCount.p);
            CU(cudaGetLastError());
            ctx->launches += 1;
        } else {
            CU(cudaMemcpyAsync(ctx->hStats.p, ctx->dStats.p, sizeof(DrawStat) * nDraws, cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(ctx->hCount.p, ctx->dCount.p, 4ull * std::max(nChunks, 1u), cudaMemcpyDeviceToHost, s));
        }
        CU(cudaStreamSynchronize(s));
        ctx->d2hBytes += sizeof(DrawStat) * (uint64_t)nDraws + 4;
        uint64_t acc = 0;
        for (uint32_t i = 0; i < nDraws; i++) {
            const DrawStat& st = ctx->hStats.p[i];
            t3d_draw_result& r = hRes.p[i];
            r.first_vertex = 3 * acc;
            r.vertex_count = 3 * st.survivors;
            r.total_tris = st.minBits == DRAW_CULLED_MARK ? 0u : ctx->drawTotalTris[i];      // an instance outside the frustum was never render()ed
            if (st.survivors) { memcpy(&r.min_depth, &st.minBits, 4); memcpy(&r.max_depth, &st.maxBits, 4); }
            else { r.min_depth = FLT_MAX; r.max_depth = -FLT_MAX; }
            acc += st.survivors;
        }
        uint64_t scanned = 0;
        for (uint32_t c = 0; c < nChunks; c++) scanned += ctx->hCount.p[c];
        if (acc != scanned) return fail(ctx, T3D_E_CUDA, "internal: survivor count mismatch between cull and tile scan");
        nVerts = 3 * acc;
    } else {
        CU(cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < nDraws; i++) {
            t3d_draw_result& r = hRes.p[i];
            r.first_vertex = 0; r.vertex_count = 0; r.total_tris = ctx->drawTotalTris[i]; r.min_depth = FLT_MAX; r.max_depth = -FLT_MAX;
        }
    }
    ctx->nVerticesOut = nVerts;
    out->n_vertices = nVerts;
    out->draws = hRes.p;

    // identity index table (camera.go:1816), sized for the largest flush
    uint64_t maxFlush = 0, cur = 0;
    for (uint32_t i = 0; i < nDraws; i++) {
        if (!ctx->drawContinue[i]) cur = 0;
        cur += hRes.p[i].vertex_count;
        maxFlush = std::max(maxFlush, cur);
    }
    if (ctx->hIndex.n < maxFlush || ctx->hIndex.n == 0) {
        size_t old = ctx->hIndex.n;
        size_t want = std::max<size_t>(std::max<uint64_t>(maxFlush, 65536), old);
        CU(ctx->hIndex.resize(want));
        for (size_t i = old; i < want; i++) ctx->hIndex.p[i] = (uint16_t)i;
    }
    out->index_list = ctx->hIndex.p;

    // a shard's frame left no lists here: they were delivered in compact form to the issuing rank
    const bool compactDepth = ctx->gather.active && ctx->gather.rank != ctx->gather.dst;
    if (download && compactDepth)
        return fail(ctx, T3D_E_STATE, "t3d_frame_results: this rank's lists were delivered to the issuing rank (t3d_gather_attach); use t3d_frame_results_device here and t3d_gather_results there");
    if (download) {
        const uint32_t set = ctx->frameSet;
        const bool async = ctx->asyncDownload && !ctx->gather.active && !ctx->xColor;
        PinVec<t3d_vertex> &hC = set ? ctx->hColorB : ctx->hColor, &hD = set ? ctx->hDepthB : ctx->hDepth, &hN = set ? ctx->hNormalB : ctx->hNormal;
        PinVec<uint32_t>& hT = set ? ctx->hTriIdsB : ctx->hTriIds;
        // flush() draws the depth list only when Camera.RenderDepth is on (camera.go:1880-1934); without it the list holds positions
        // and stale colours that nothing reads
        const bool wantDepth = !ctx->skipUnusedDepth || (ctx->fc.camFlags & T3D_CAM_RENDER_DEPTH);
        // this set's host buffers are about to be resized and overwritten: its previous copy (two frames ago) has to have arrived
        if (ctx->copyPending[set]) { CU(cudaEventSynchronize(ctx->evCopyDone[set])); ctx->copyPending[set] = false; }
        if (nVerts) {
            cudaEvent_t e0 = ctx->ev[7], e1 = ctx->ev[8];
            CU(hC.resize(nVerts)); if (wantDepth) CU(hD.resize(nVerts)); CU(hT.resize(nVerts / 3));
            if (normals) CU(hN.resize(nVerts));
            cudaStream_t cs = s;
            if (async) {        // the copies run on a stream of their own, behind the frame; the caller goes on to the next frame
                if (!ctx->dlStream) {
                    CU(cudaStreamCreateWithFlags(&ctx->dlStream, cudaStreamNonBlocking));
                    CU(cudaEventCreateWithFlags(&ctx->evFrameDone, cudaEventDisableTiming));
                    CU(cudaEventCreateWithFlags(&ctx->evCopyDone[0], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&ctx->evCopyDone[1], cudaEventDisableTiming));
                }
                cs = ctx->dlStream;
                CU(cudaEventRecord(ctx->evFrameDone, s));
                CU(cudaStreamWaitEvent(cs, ctx->evFrameDone, 0));
            } else CU(cudaEventRecord(e0, s));
            CU(cudaMemcpyAsync(hC.p, ctx->oColor, nVerts * sizeof(t3d_vertex), cudaMemcpyDeviceToHost, cs));
            if (wantDepth) CU(cudaMemcpyAsync(hD.p, ctx->oDepth, nVerts * sizeof(t3d_vertex), cudaMemcpyDeviceToHost, cs));
            if (normals) CU(cudaMemcpyAsync(hN.p, ctx->oNormal, nVerts * sizeof(t3d_vertex), cudaMemcpyDeviceToHost, cs));
            CU(cudaMemcpyAsync(hT.p, ctx->oTriIds, (nVerts / 3) * 4, cudaMemcpyDeviceToHost, cs));
            if (async) {
                CU(cudaEventRecord(ctx->evCopyDone[set], cs));
                ctx->copyPending[set] = true;
            } else {
                CU(cudaEventRecord(e1, s));
                CU(cudaStreamSynchronize(s));
                cudaEventElapsedTime(&ctx->downloadMs, e0, e1);
            }
            ctx->d2hBytes += nVerts * sizeof(t3d_vertex) * ((normals ? 2 : 1) + (wantDepth ? 1 : 0)) + (nVerts / 3) * 4;
        }
        if (async) { ctx->lastCopySet = set; ctx->nextSet = set ^ 1u; }
        out->color_list = hC.p; out->depth_list = wantDepth ? hD.p : nullptr;
        out->normal_list = normals ? hN.p : nullptr;
        out->tri_ids = hT.p;
    } else {
        out->color_list = compactDepth ? nullptr : ctx->oColor; out->depth_list = compactDepth ? nullptr : ctx->oDepth;
        out->normal_list = normals ? ctx->oNormal : nullptr;
        out->tri_ids = ctx->oTriIds;
    }
    if (tm) {
        memset(tm, 0, sizeof *tm);
        if (ctx->stagesTimed || !(nDraws && totalT)) {
            cudaEventElapsedTime(&tm->upload_ms, ctx->ev[0], ctx->ev[1]);
            cudaEventElapsedTime(&tm->prep_ms, ctx->ev[1], ctx->ev[2]);
            cudaEventElapsedTime(&tm->vertex_ms, ctx->ev[2], ctx->ev[3]);
            cudaEventElapsedTime(&tm->cull_ms, ctx->ev[3], ctx->ev[4]);
            cudaEventElapsedTime(&tm->sort_ms, ctx->ev[4], ctx->ev[5]);
            cudaEventElapsedTime(&tm->emit_ms, ctx->ev[5], ctx->ev[9]);
        }
        cudaEventElapsedTime(&tm->total_ms, ctx->ev[0], ctx->ev[6]);
        cudaEventElapsedTime(&tm->gather_ms, ctx->ev[9], ctx->ev[6]);
        tm->download_ms = ctx->downloadMs;
    }
    ctx->haveResults = true;
    return T3D_OK;
}

int t3d_frame_results(t3d_ctx* ctx, t3d_frame_output* out, t3d_timings* timings) { return finish_frame(ctx, true, out, timings); }
int t3d_frame_results_device(t3d_ctx* ctx, t3d_frame_output* out, t3d_timings* timings) { return finish_frame(ctx, false, out, timings); }

int t3d_set_skip_unused_depth_list(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    ctx->skipUnusedDepth = enabled != 0;
    return T3D_OK;
}

int t3d_set_async_download(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    if (!enabled && ctx->asyncDownload) {       // drain: nothing may still be in flight when the ordinary set is used again
        CU(cudaSetDevice(ctx->device));
        for (int k = 0; k < 2; k++) if (ctx->copyPending[k]) { CU(cudaEventSynchronize(ctx->evCopyDone[k])); ctx->copyPending[k] = false; }
        ctx->nextSet = 0;
    }
    ctx->asyncDownload = enabled != 0;
    return T3D_OK;
}

int t3d_frame_download_wait(t3d_ctx* ctx, uint32_t frames_back) {
    if (!ctx) return T3D_E_INVALID;
    if (frames_back > 1) return fail(ctx, T3D_E_INVALID, "t3d_frame_download_wait: only the last two frames' lists exist");
    const uint32_t set = frames_back ? ctx->lastCopySet ^ 1u : ctx->lastCopySet;
    if (ctx->copyPending[set]) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventSynchronize(ctx->evCopyDone[set]));
        ctx->copyPending[set] = false;
    }
    return T3D_OK;
}

// ---- introspection ---------------------------------------------------------------------------------

int t3d_debug_read_draw_vertices(t3d_ctx* ctx, uint32_t draw, uint32_t vtx_start, uint32_t vtx_end, float* clip4, float* light4) {
    if (!ctx) return T3D_E_INVALID;
    if (!ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_debug_read_draw_vertices: frame not executed");
    if (draw >= ctx->hDraws.size()) return fail(ctx, T3D_E_INVALID, "t3d_debug_read_draw_vertices: draw out of range");
    CU(cudaSetDevice(ctx->device));
    const t3d_ctx::HostDraw& R = ctx->hDraws[draw];
    if (vtx_start > vtx_end) return fail(ctx, T3D_E_INVALID, "t3d_debug_read_draw_vertices: bad range");
    const size_t n = vtx_end - vtx_start;
    if (n == 0) return T3D_OK;
    if (vtx_start < R.a0 || vtx_end > R.a0 + R.nv) return fail(ctx, T3D_E_INVALID, "t3d_debug_read_draw_vertices: range outside the part's vertices");
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t slot = R.vbase + (vtx_start - R.a0);
    if (clip4) {        // the arenas hold {X/w, Y/w, Z, W} and {X, Y}: reassemble globalVertexTransforms' {X, Y, Z, W}
        std::vector<float> xy(2 * n);
        CU(cudaMemcpy(clip4, ctx->dClip.as<float4>() + slot, n * 16, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(xy.data(), ctx->dClipXY.as<float2>() + slot, n * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) { clip4[4 * i] = xy[2 * i]; clip4[4 * i + 1] = xy[2 * i + 1]; }
    }
    if (light4) {
        if (R.flags & T3D_DRAW_LIT) CU(cudaMemcpy(light4, ctx->dLight.as<float4>() + slot, n * 16, cudaMemcpyDeviceToHost));
        else memset(light4, 0, n * 16);
    }
    return T3D_OK;
}

int t3d_frame_read_vertex_lights(t3d_ctx* ctx, uint32_t draw, uint32_t vtx_start, uint32_t vtx_end, float* light4) {
    if (!ctx || !light4) return fail(ctx, T3D_E_INVALID, "t3d_frame_read_vertex_lights: NULL argument");
    if (!ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_frame_read_vertex_lights: frame not executed");
    if (draw >= ctx->hDraws.size() || vtx_start > vtx_end) return fail(ctx, T3D_E_INVALID, "t3d_frame_read_vertex_lights: draw or range out of bounds");
    const t3d_ctx::HostDraw& R = ctx->hDraws[draw];
    // vertices of the range that no triangle of the part refers to were never lit: zero
    for (size_t i = 0; i < (size_t)(vtx_end - vtx_start); i++) { light4[4 * i] = light4[4 * i + 1] = light4[4 * i + 2] = 0.0f; light4[4 * i + 3] = 1.0f; }
    const uint32_t lo = std::max(vtx_start, R.a0), hi = std::min(vtx_end, R.a0 + R.nv);
    if (lo >= hi || !(R.flags & T3D_DRAW_LIT)) return T3D_OK;
    return t3d_debug_read_draw_vertices(ctx, draw, lo, hi, nullptr, light4 + 4 * (size_t)(lo - vtx_start));
}

int t3d_set_output_target(t3d_ctx* ctx, t3d_vertex* color, t3d_vertex* depth, t3d_vertex* normal, uint64_t vertex_capacity) {
    if (!ctx) return T3D_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));       // a frame still in flight keeps writing to the old target
    if ((color == nullptr) != (depth == nullptr)) return fail(ctx, T3D_E_INVALID, "t3d_set_output_target: colour and depth lists go together");
    if (color && (((uintptr_t)color | (uintptr_t)depth | (uintptr_t)normal) & 15)) return fail(ctx, T3D_E_INVALID, "t3d_set_output_target: lists must be 16-byte aligned");
    ctx->xColor = color; ctx->xDepth = depth; ctx->xNormal = color ? normal : nullptr;
    ctx->xCapacity = color ? vertex_capacity : 0;
    return T3D_OK;
}

int t3d_shared_alloc(t3d_ctx* ctx, uint64_t bytes, void** device_ptr, t3d_shared_handle* handle) {
    if (!ctx || !device_ptr || !handle || bytes == 0) return fail(ctx, T3D_E_INVALID, "t3d_shared_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(t3d_shared_handle), "handle size");
    CU(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    memset(handle, 0, sizeof *handle);
    memcpy(handle->bytes, &h, sizeof h);
    ctx->sharedAllocs.push_back(p);
    *device_ptr = p;
    return T3D_OK;
}

int t3d_shared_open(t3d_ctx* ctx, const t3d_shared_handle* handle, void** device_ptr) {
    if (!ctx || !device_ptr || !handle) return fail(ctx, T3D_E_INVALID, "t3d_shared_open: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->sharedOpened.push_back(p);
    *device_ptr = p;
    return T3D_OK;
}

int t3d_gather_alloc(t3d_ctx* ctx, uint32_t world, uint64_t vertex_capacity, uint32_t draw_capacity, t3d_shared_handle handles[3]) {
    if (!ctx || !handles || world < 2 || vertex_capacity == 0) return fail(ctx, T3D_E_INVALID, "t3d_gather_alloc: bad argument");
    if (ctx->gather.allocated || ctx->gather.active) return fail(ctx, T3D_E_STATE, "t3d_gather_alloc: already allocated");
    t3d_ctx::Gather& G = ctx->gather;
    G.world = world; G.cap = (vertex_capacity + 255) & ~(uint64_t)255; G.drawCap = std::max(draw_capacity, 1u);
    G.layout();
    void *c = nullptr, *d = nullptr, *b = nullptr;
    int rc = t3d_shared_alloc(ctx, (uint64_t)world * G.cap * sizeof(t3d_vertex), &c, &handles[0]);
    if (rc == T3D_OK) rc = t3d_shared_alloc(ctx, (uint64_t)world * G.cap * sizeof(t3d_vertex), &d, &handles[1]);
    if (rc == T3D_OK) rc = t3d_shared_alloc(ctx, G.boxBytes, &b, &handles[2]);
    if (rc != T3D_OK) return rc;
    CU(cudaMemset(static_cast<uint8_t*>(b) + G.mailOff, 0, G.boxBytes - G.mailOff));
    G.color = static_cast<uint8_t*>(c); G.depth = static_cast<uint8_t*>(d); G.box = static_cast<uint8_t*>(b);
    G.allocated = true;
    return T3D_OK;
}

int t3d_gather_attach(t3d_ctx* ctx, uint32_t world, uint32_t rank, uint32_t dst, uint64_t vertex_capacity, uint32_t draw_capacity, const t3d_shared_handle handles[3]) {
    if (!ctx || world < 2 || rank >= world || dst >= world) return fail(ctx, T3D_E_INVALID, "t3d_gather_attach: bad argument");
    if (ctx->inFrame && !ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_gather_attach: a frame is open");
    t3d_ctx::Gather& G = ctx->gather;
    if (G.active) return fail(ctx, T3D_E_STATE, "t3d_gather_attach: already attached");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (rank == dst) {
        if (!G.allocated || G.world != world || G.cap != ((vertex_capacity + 255) & ~(uint64_t)255) || G.drawCap != std::max(draw_capacity, 1u))
            return fail(ctx, T3D_E_STATE, "t3d_gather_attach: the issuing rank attaches what it allocated with t3d_gather_alloc");
        // serials restart at 1 with every attach (all ranks detach and attach together, with a barrier of the caller's in between)
        CU(cudaMemset(G.box + G.mailOff, 0, G.boxBytes - G.mailOff));
    } else {
        if (!handles) return fail(ctx, T3D_E_INVALID, "t3d_gather_attach: handles is NULL");
        G.world = world; G.cap = (vertex_capacity + 255) & ~(uint64_t)255; G.drawCap = std::max(draw_capacity, 1u);
        G.layout();
        void *c = nullptr, *d = nullptr, *b = nullptr;
        int rc = t3d_shared_open(ctx, &handles[0], &c);
        if (rc == T3D_OK) rc = t3d_shared_open(ctx, &handles[1], &d);
        if (rc == T3D_OK) rc = t3d_shared_open(ctx, &handles[2], &b);
        if (rc != T3D_OK) return rc;
        G.color = static_cast<uint8_t*>(c); G.depth = static_cast<uint8_t*>(d); G.box = static_cast<uint8_t*>(b);
        CU(ctx->dGatherErr.ensure(64));
        CU(cudaMemset(ctx->dGatherErr.p, 0, 64));
    }
    G.rank = rank; G.dst = dst; G.serial = 0; G.active = true;
    return T3D_OK;
}

int t3d_gather_detach(t3d_ctx* ctx) {
    if (!ctx) return T3D_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gather.active && ctx->gather.rank == ctx->gather.dst) { ctx->xColor = ctx->xDepth = ctx->xNormal = nullptr; ctx->xCapacity = 0; }
    ctx->gather.active = false;      // the allocations stay mapped until t3d_destroy
    return T3D_OK;
}

int t3d_gather_results(t3d_ctx* ctx, t3d_gather_output* out) {
    if (!ctx || !out) return fail(ctx, T3D_E_INVALID, "t3d_gather_results: NULL argument");
    t3d_ctx::Gather& G = ctx->gather;
    if (!G.active || G.rank != G.dst) return fail(ctx, T3D_E_STATE, "t3d_gather_results: only on the issuing rank of an attached gather");
    if (!ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_gather_results: frame not executed");
    CU(cudaSetDevice(ctx->device));
    const size_t words = ((size_t)G.world + 1) * GATHER_MAIL_WORDS;
    CU(ctx->hMail.resize(words));
    CU(cudaMemcpyAsync(ctx->hMail.p, G.box + G.mailOff, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->hMail.p[(size_t)G.world * GATHER_MAIL_WORDS]) return fail(ctx, T3D_E_CUDA, "t3d_gather_results: timed out waiting for a shard's frame");
    memset(out, 0, sizeof *out);
    out->world = G.world; out->vertex_capacity = G.cap;
    out->color_list = reinterpret_cast<const t3d_vertex*>(G.color); out->depth_list = reinterpret_cast<const t3d_vertex*>(G.depth);
    out->n_vertices = ctx->hMail.p;        // stride GATHER_MAIL_WORDS; the issuing rank's own entry is patched below
    ctx->hMail.p[(size_t)G.dst * GATHER_MAIL_WORDS + 1] = (uint32_t)ctx->nVerticesOut;
    out->n_vertices_stride = GATHER_MAIL_WORDS; out->n_vertices += 1;
    return T3D_OK;
}

int t3d_gather_rank_stats(t3d_ctx* ctx, uint32_t rank, uint32_t n_draws, t3d_draw_stat* out) {
    if (!ctx || (n_draws && !out)) return fail(ctx, T3D_E_INVALID, "t3d_gather_rank_stats: NULL argument");
    t3d_ctx::Gather& G = ctx->gather;
    if (!G.active || G.rank != G.dst || rank >= G.world || n_draws > G.drawCap) return fail(ctx, T3D_E_STATE, "t3d_gather_rank_stats: only on the issuing rank, for a rank and draw count inside the attached capacity");
    static_assert(sizeof(t3d_draw_stat) == sizeof(DrawStat), "t3d_draw_stat mirrors DrawStat");
    CU(cudaSetDevice(ctx->device));
    if (!n_draws) return T3D_OK;
    const void* src = rank == G.dst ? ctx->dStats.p : (const void*)G.stats(rank);
    CU(cudaMemcpyAsync(out, src, sizeof(DrawStat) * (size_t)n_draws, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return T3D_OK;
}

int t3d_debug_read_bones(t3d_ctx* ctx, uint32_t offset, uint32_t n, float* matrices16) {
    if (!ctx || (n && !matrices16)) return fail(ctx, T3D_E_INVALID, "t3d_debug_read_bones: NULL argument");
    if (!ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_debug_read_bones: frame not executed");
    const size_t first = (offset & POSE_TAG) ? (size_t)ctx->frameBoneCount + (offset & ~POSE_TAG) : offset;
    if (first + n > (size_t)ctx->frameBoneCount + ctx->posedBones) return fail(ctx, T3D_E_INVALID, "t3d_debug_read_bones: range outside the bone arena");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n) CU(cudaMemcpy(matrices16, ctx->pBones + 16 * first, 64ull * n, cudaMemcpyDeviceToHost));
    return T3D_OK;
}

int t3d_debug_force_full_upload(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    ctx->forceFullUpload = enabled != 0;
    return T3D_OK;
}

int t3d_set_exact_lighting(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    if (ctx->inFrame && !ctx->executed) return fail(ctx, T3D_E_STATE, "t3d_set_exact_lighting: a frame is open");
    ctx->exactLighting = enabled != 0;
    return T3D_OK;
}

int t3d_set_stage_timings(t3d_ctx* ctx, int enabled) {
    if (!ctx) return T3D_E_INVALID;
    ctx->stageTimings = enabled != 0;
    return T3D_OK;
}

int t3d_debug_last_launch_count(t3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
float t3d_debug_energy_cutoff(float energy) { return energy_cutoff(energy); }
int t3d_debug_set_fused_sort(t3d_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return T3D_E_INVALID;
    ctx->fuseSort = mode;
    return T3D_OK;
}
int t3d_debug_vertex_trace(t3d_ctx* ctx, unsigned long long* out, uint32_t n_blocks) {
    if (!ctx || !out) return T3D_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return T3D_E_CUDA;
    return debug_vertex_trace(out, n_blocks) == 0 ? T3D_OK : fail(ctx, T3D_E_UNSUPPORTED, "t3d_debug_vertex_trace: the library was not built with -DT3D_VERTEX_TRACE");
}
void* t3d_debug_stream(t3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t t3d_debug_vertex_kernel_bytes(t3d_ctx* ctx) { return ctx ? ctx->vertexKernelBytes : 0; }
int t3d_debug_frame_bytes(t3d_ctx* ctx, uint64_t* h2d, uint64_t* d2h) {
    if (!ctx) return T3D_E_INVALID;
    if (h2d) *h2d = ctx->h2dBytes;
    if (d2h) *d2h = ctx->d2hBytes;
    return T3D_OK;
}

}  // extern "C"
