// tetra3d.hpp -- host-side mirror of the Tetra3D API for the geometry stage (C++17).
//
// The reference is a Go library; no Go toolchain exists in this image, so the host side above the C ABI
// (include/t3dgeo.h) is written in C++ with the reference's own type and method names: Scene, Node,
// Model, Mesh, MeshPart, Material, Camera, AmbientLight / PointLight / DirectionalLight / LightVolume,
// LightGroup, World.  It keeps what stays on the host in the reference too -- the scene graph and cached
// transforms (node.go:572-608), the Camera.Render prologue that builds and orders the render pairs
// (camera.go:1269-1505, 2050-2109) and frustum culling (camera.go:933-989) -- and hands every render(rp)
// to libt3dgeo.  There is no CPU implementation of ProcessVertices / lighting / emit in here.
#pragma once

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../csrc/t3d_math.h"
#include "../../include/t3dgeo.h"

namespace tetra3d {

using Vector3 = t3d::Vec3;
using Matrix4 = t3d::Mat4;

struct Color4 { float R = 1, G = 1, B = 1, A = 1; };

// ---- matrix.go builders ---------------------------------------------------------------------------
Matrix4 NewMatrix4();
Matrix4 NewMatrix4Translate(float x, float y, float z);
Matrix4 NewMatrix4Scale(float x, float y, float z);
Matrix4 NewMatrix4Rotate(float x, float y, float z, float angle);                       // matrix.go:100
Matrix4 NewProjectionMatrix4Perspective(float fovy, float near_, float far_, float w, float h);   // matrix.go:588
Matrix4 NewProjectionMatrix4Orthographic(float near_, float far_, float r, float l, float t, float b);
bool    Matrix4Equals(const Matrix4& a, const Matrix4& b);                              // matrix.go:511

// ---- mesh.go ----------------------------------------------------------------------------------------
struct Dimensions {
    Vector3 Min{0, 0, 0}, Max{0, 0, 0};
    float MaxSpan() const;                  // mesh.go:41
    Vector3 Center() const;                 // mesh.go:46
};

enum { TriangleSortModeBackToFront = 0, TriangleSortModeFrontToBack, TriangleSortModeNone };
enum { TransparencyModeAuto = 0, TransparencyModeOpaque, TransparencyModeAlphaClip, TransparencyModeTransparent };
enum { BillboardYUpGlobalY = 0, BillboardYUpCameraY };
enum { LightingModeDefault = 0, LightingModeFixedNormals, LightingModeDoubleSided };
enum { DepthModeDefault = 0, DepthModeUnbillboarded };
enum { LightVolumeShadingModePerVertexWithoutNormal = 0, LightVolumeShadingModePerVertexWithNormal,
       LightVolumeShadingModePerObject, LightVolumeShadingModeOff };
enum { FogAdd = 0, FogSub, FogOverwrite };

// material.go:66-131 (fields the geometry stage reads)
struct Material {
    std::string name;
    Color4 Color;
    bool BackfaceCulling = true;
    int  TriangleSortMode = TriangleSortModeBackToFront;
    bool Shadeless = false;
    int  LightVolumeShadingMode = 0;
    bool BillboardEnabled = false, BillboardLockX = false, BillboardLockY = false, BillboardLockZ = false;
    int  BillboardUpDirection = BillboardYUpGlobalY;
    bool Visible = true;
    int  TransparencyMode = TransparencyModeAuto;
    int  BillboardedDepthMode = DepthModeDefault;
    int  LightingMode = LightingModeDefault;
    int  RenderOrder = 0;
    bool BlendSourceOver = true;            // mat.Blend == ebiten.BlendSourceOver
    bool UseTexture = false;
    int  TextureWidth = 0, TextureHeight = 0;   // Texture.Bounds(); 0 = no texture
    // host callbacks the device path cannot run (material.go:129): presence makes Render fail loudly
    bool HasCustomDepthFunction = false;
};
std::shared_ptr<Material> NewMaterial(const std::string& name);   // material.go:150-170 defaults

struct Mesh;
// mesh.go:2183-2190; TriangleEnd is inclusive as in the reference
struct MeshPart {
    Mesh* mesh = nullptr;
    std::shared_ptr<Material> material;
    int VertexIndexStart = 0, VertexIndexEnd = 0, TriangleStart = 0, TriangleEnd = -1;
    int TriangleCount() const { return TriangleEnd - TriangleStart + 1; }
    bool isVisible() const { return material ? material->Visible : true; }   // mesh.go:2402
};

// mesh.go:260-302
struct Mesh {
    uint64_t id = 0;
    uint32_t generation = 1;          // bump (MarkDirty) after mutating vertex arrays
    std::string Name;
    std::vector<float> VertexPositions, VertexNormals, VertexUVs;   // 3n, 3n, 2n
    std::vector<float> VertexColors;                                // 4n: the active channel
    int VertexActiveColorChannel = -1;
    std::vector<float> VertexWeights;  std::vector<uint16_t> VertexBones;  std::vector<uint8_t> VertexInfluences;  // 4n,4n,n
    std::vector<uint32_t> TriangleIndices;                          // 3t
    std::vector<float> TriangleCenters;                             // 3t  (Triangle.Center)
    // Auto-subdivision (mesh.go:447-485, 1758-1861): TriangleIndices / TriangleCenters hold every base triangle followed by
    // its pre-built sub-triangles level by level (the order forEachTri(true, ...) visits them in); per entry the base
    // triangle, the level (-1: base), and for base triangles the number of levels, MaxSpan and Normal.
    bool autoSubdivide = false;
    std::vector<uint32_t> TriangleSubBase; std::vector<int8_t> TriangleSubLevel; std::vector<uint8_t> TriangleLevelCount;
    std::vector<float> TriangleMaxSpan, TriangleNormals;
    std::vector<std::shared_ptr<MeshPart>> MeshParts;
    std::vector<float> ShapeKeyPositions, ShapeKeyNormals, ShapeKeyWeights;   // k*n*3, k*n*3, k
    uint32_t shapeKeyWeightGeneration = 0;
    Dimensions dimensions;
    int VertexCount() const { return (int)(VertexPositions.size() / 3); }
    int TriangleTotal() const { return (int)(TriangleIndices.size() / 3); }
    std::shared_ptr<MeshPart> AddMeshPart(std::shared_ptr<Material> mat, int vtxStart, int vtxEnd, int triStart, int triEndInclusive);
    void UpdateDimensions();          // mesh.go:719
    void RecalculateCenters();        // Triangle.RecalculateCenter, mesh.go:1972
    void AutoNormal();                // mesh.go:755
    void MarkDirty() { generation++; }
    void SetShapeKeyWeights(const std::vector<float>& w);
};
std::shared_ptr<Mesh> NewMesh(const std::string& name);
std::shared_ptr<Mesh> NewCubeMesh(float w, float h, float d);       // mesh.go:1251

// ---- node.go ----------------------------------------------------------------------------------------
enum class NodeType { Node, Model, Camera, AmbientLight, PointLight, DirectionalLight, LightVolume };

struct Node : std::enable_shared_from_this<Node> {
    std::string name;
    Vector3 position{0, 0, 0}, scale{1, 1, 1};
    Matrix4 rotation;
    Node* parent = nullptr;
    std::vector<std::shared_ptr<Node>> children;
    bool visible = true;
    Matrix4 cachedTransform;
    bool isTransformDirty = true;

    explicit Node(const std::string& n);
    virtual ~Node() = default;
    virtual NodeType Type() const { return NodeType::Node; }
    virtual void onTransformUpdate() {}

    Matrix4 Transform();                    // node.go:572
    void dirtyTransform();                  // node.go:620
    Vector3 WorldPosition();                // node.go:737
    Matrix4 WorldRotation();                // node.go:916
    void SetLocalPosition(float x, float y, float z);
    void SetLocalScale(float x, float y, float z);
    void SetLocalRotation(const Matrix4& r);
    void AddChildren(const std::vector<std::shared_ptr<Node>>& kids);
    void SetVisible(bool v, bool recursive);
    bool IsVisible() const { return visible; }
    // depth-first pre-order, like rootNode.ForEachChild(true, ...) in RenderNodes (camera.go:1200)
    void ForEachChild(const std::function<void(Node*)>& f);
};

// ---- light.go / lightGroup.go ---------------------------------------------------------------------
struct ILight : Node {
    Color4 color; float energy = 1;
    ILight(const std::string& n, float r, float g, float b, float e) : Node(n) { color = Color4{r, g, b, 1}; energy = e; }
};
struct AmbientLight : ILight { using ILight::ILight; NodeType Type() const override { return NodeType::AmbientLight; } };
struct PointLight : ILight {
    using ILight::ILight;
    float lightRange = 0;
    void SetRange(float r) { lightRange = r; }
    NodeType Type() const override { return NodeType::PointLight; }
};
struct DirectionalLight : ILight { using ILight::ILight; NodeType Type() const override { return NodeType::DirectionalLight; } };
struct LightVolume : ILight {
    LightVolume(const std::string& n) : ILight(n, 1, 1, 1, 1) {}
    Dimensions dimensions;
    float cellSize[3] = {1, 1, 1};
    int cellCount[3] = {0, 0, 0};
    Vector3 offset{0, 0, 0}, lightBias{0, 0, 0};
    std::vector<float> cellColors;      // 3 * cells
    uint32_t volumeId = 0; uint32_t generation = 1, uploadedGeneration = 0;
    // LightVolumeCell.useLights (light.go:1381, 1508-1521), keyed by the cell's index x + y*cx + z*cx*cy, in SetUseLights
    // order; useLightsRenderSet (light.go:392): every light some cell has used, first-added order
    std::map<size_t, std::vector<std::shared_ptr<ILight>>> cellUseLights;
    std::vector<std::shared_ptr<ILight>> useLightsRenderSet;
    uint64_t useLightsFrame = 0;        // GeometryDevice frame the render set was last submitted for
    void CellSetUseLights(int x, int y, int z, const std::vector<std::shared_ptr<ILight>>& lights);   // LightVolumeCell.SetUseLights
    void CellClearUseLights(int x, int y, int z);                                                     // LightVolumeCell.ClearUseLights
    NodeType Type() const override { return NodeType::LightVolume; }
};
struct LightGroup {
    std::vector<std::shared_ptr<ILight>> Lights;
    bool Active = true;
};

// ---- animation.go -------------------------------------------------------------------------------------
enum { TrackTypePosition = 0, TrackTypeScale = 1, TrackTypeRotation = 2 };           // animation.go:10-14 (strings there)
enum { InterpolationLinear = 0, InterpolationConstant = 1, InterpolationCubic = 2 }; // animation.go:17-21
enum { AnimationLoopModeLoop = 0, AnimationLoopModePingPong = 1, AnimationLoopModeOneshot = 2 };
struct Keyframe { float Time; float Data[4]; };                                       // animation.go:36-39
struct AnimationTrack {                                                              // animation.go:49-54
    int Type = TrackTypePosition, Interpolation = InterpolationLinear;
    std::vector<Keyframe> Keyframes;
    void AddKeyframe(float time, float x, float y, float z, float w) { Keyframes.push_back(Keyframe{time, {x, y, z, w}}); }
};
struct AnimationChannel {                                                            // animation.go:168-173
    std::string Name;
    std::shared_ptr<AnimationTrack> Tracks[3];
    std::shared_ptr<AnimationTrack> AddTrack(int trackType);
};
struct Animation {                                                                   // animation.go:204-222
    std::string name;
    uint64_t id;
    std::vector<std::shared_ptr<AnimationChannel>> channels;    // Go: map by channel name
    float length = 0;
    int loopMode = AnimationLoopModeLoop;
    bool relativeMotion = false;
    uint32_t generation = 1;                                    // bump (MarkDirty) after editing keyframes: the device copy is rebuilt
    explicit Animation(const std::string& n);
    std::shared_ptr<AnimationChannel> AddChannel(const std::string& name);
    void MarkDirty() { generation++; }
};

// The joint list of a skin (gltf.go:1452-1511) as one shareable object: which bone hangs off which, the bones'
// inverse bind matrices and the local transforms their Nodes hold when no track overrides them.  Every armature
// cloned from one GLTF skin shares it.
struct Skeleton {
    uint64_t id;
    std::vector<std::string> boneNames;
    std::vector<int32_t> parent;                 // parent bone or -1; parents first
    std::vector<float> inverseBind, restPosition, restScale, restRotation;   // 16, 3, 3, 16 floats per bone
    uint32_t generation = 1, uploadedGeneration = 0;
    Skeleton();
    int BoneCount() const { return (int)parent.size(); }
    // bone nodes in joint order; parents must come before children.  Bones whose parent is not in the list hang off the pose's root.
    static std::shared_ptr<Skeleton> FromJoints(const std::vector<std::shared_ptr<Node>>& joints, const std::vector<float>& inverseBind16);
};

// AnimationPlayer (animation.go:437-480) for armatures whose bone poses are evaluated on the device: it keeps the
// playhead logic of updateValues (animation.go:724-806) on the host and records the playhead the tracks were last
// sampled at; the sampling itself, the bones' Transform() and boneInfluence run in k_pose_bones.
//
// Blending (BlendTime > 0, animation.go:548-554, 852-907): the reference blends from prevAnimatedProperties -- the copy Play()
// took of currentProperties, i.e. of the targets of the last update -- into the playing animation's samples.  Those targets
// are a function of a short history: the animations played since the last blend ended, each with the playhead and blend
// fraction of its last update.  The player keeps that history (a chain of BlendLevels) instead of sampled values, and
// Camera::Render ships it: t3d_submit_poses for a chain of one, t3d_submit_poses_blended for two (a blend between two
// animations), t3d_submit_pose_chains for blends that were interrupted by further Play() calls (up to T3D_POSE_CHAIN_MAX levels;
// longer ones are refused).  Relative motion, markers and sequence callbacks need the host path and are refused.
struct BlendLevel { std::shared_ptr<Animation> anim; float playhead; float bp; };
struct AnimationPlayer {
    std::shared_ptr<Node> RootNode;
    std::shared_ptr<Animation> animation;
    float playhead = 0, prevPlayhead = 0, PlaySpeed = 1, BlendTime = 0;
    bool playing = false, finished = false, justLooped = false, sampled = false;
    int loopCount = 0;
    float sampledPlayhead = 0;                  // the playhead updateValues evaluated the tracks at (animation.go:686-712)
    // time.Now() in milliseconds (animation.go:553, 854); replaceable so that tests can step the clock
    std::function<int64_t()> NowMs;
    // ---- blend bookkeeping ----
    bool blendStarted = false;                  // !blendStart.IsZero()
    int64_t blendStartMs = 0;
    std::vector<BlendLevel> prevChain;          // what prevAnimatedProperties holds (empty: nothing)
    std::vector<BlendLevel> curChain;           // what currentProperties holds: filled only by updates with BlendTime > 0 (animation.go:909-937)
    bool currentHas = false;                    //   ... and cleared by assignChannels at the first update after a Play() (animation.go:624)
    std::vector<BlendLevel> sampledChain;       // what the last forceUpdate applied, for Camera::Render: the last level is the playing animation
    explicit AnimationPlayer(std::shared_ptr<Node> root) : RootNode(std::move(root)) {}
    void Play(std::shared_ptr<Animation> anim);         // animation.go:517-566
    void Update(float dt);                              // animation.go:810-824
    void SetPlayhead(float time) { playhead = time; forceUpdate(0); }    // animation.go:828-831
    float Playhead() const { return playhead; }
    bool Finished() const { return finished; }
    bool IsPlaying() const { return playing; }
    void Pause() { playing = false; }
    void Resume() { playing = true; }
    void forceUpdate(float dt);                         // animation.go:837 -> updateValues 674-806
    bool Blending() const { return sampledChain.size() >= 2; }
    bool BlendStartIsBlend() const { return sampledChain.size() >= 3; }       // the blend started from an unfinished blend (animation.go:905-908)
    float BlendFraction() const { return sampledChain.empty() ? 0.0f : sampledChain.back().bp; }
    int64_t now() const;
};

// ---- model.go -----------------------------------------------------------------------------------------
struct Model : Node {
    std::shared_ptr<Mesh> mesh;
    bool FrustumCulling = true;
    Color4 Color;
    bool Shadeless = false;
    bool skinned = false;
    std::vector<float> boneInfluences;      // 16 per skin joint: Node.boneInfluence (node.go:592-594)
    std::shared_ptr<LightGroup> lightGroup;
    std::map<int, std::vector<std::shared_ptr<Model>>> DynamicBatchModels;   // keyed by MeshPart index (Go: map[*MeshPart])
    Model* DynamicBatchOwner = nullptr;
    // host callbacks the device path cannot run (model.go:47, 53)
    bool HasVertexTransformFunction = false, HasVertexClipFunction = false;
    // frustumCullingSphere (model.go:142-182)
    Vector3 spherePosition{0, 0, 0}; float sphereRadius = 0;

    Model(const std::string& n, std::shared_ptr<Mesh> m);
    NodeType Type() const override { return NodeType::Model; }
    void onTransformUpdate() override;
    bool isTransparent(const MeshPart& part) const;     // model.go:1274
    bool DynamicBatcher() const { return !DynamicBatchModels.empty(); }
    void DynamicBatchAdd(int meshPartIndex, const std::vector<std::shared_ptr<Model>>& models);   // model.go:217
    void SetBoneInfluences(const std::vector<float>& m) { boneInfluences = m; skinned = !m.empty(); }
    // model.go:1203-1268 on the device (t3d_frame_set_bake_lighting): the baked VertexColors[targetChannel] of the model's mesh,
    // 4 floats per vertex; also stored in mesh->VertexColors when targetChannel is the mesh's active channel.  `scene` stands in
    // for model.Scene() (its World.AmbientLight joins the lights).
    // model.go:1040-1124 on the device (t3d_bake_ao): the same-model pass.  targetParts: MeshPart indices (empty: all);
    // targetVertices: vertex indices (empty: all).  Returns VertexColors[targetChannel] after the bake, 4 floats per vertex.
    std::vector<float> BakeAO(struct GeometryDevice& device, int targetChannel, float occlusionAngle, const Color4& occlusionColor,
                              const std::vector<int>& targetParts, const std::vector<int>& targetVertices);
    std::vector<float> BakeLighting(int targetChannel, struct Camera& camera, struct Scene& scene, const std::vector<std::shared_ptr<struct ILight>>& lights);
    // Bones posed on the device: the skin's skeleton, the node its root bones are children of (Model.SkinRoot, model.go:33)
    // and the player animating that armature (nullptr: rest pose).
    std::shared_ptr<Skeleton> skeleton;
    std::shared_ptr<Node> SkinRoot;
    std::shared_ptr<AnimationPlayer> posePlayer;
    void SetSkeleton(std::shared_ptr<Skeleton> sk, std::shared_ptr<Node> skinRoot, std::shared_ptr<AnimationPlayer> player) {
        skeleton = std::move(sk); SkinRoot = std::move(skinRoot); posePlayer = std::move(player); skinned = skeleton != nullptr; boneInfluences.clear();
    }
};

// ---- world.go / scene.go ----------------------------------------------------------------------------
struct World {
    Color4 FogColor{0, 0, 0, 0};
    int FogMode = FogAdd;
    bool FogOn = true;
    float FogRange[2] = {0, 1};
    bool LightingOn = true;
    std::shared_ptr<AmbientLight> ambientLight = std::make_shared<AmbientLight>("ambient light", 1, 1, 1, 0);   // world.go:57
};
struct Scene {
    std::string Name;
    std::shared_ptr<Node> Root;
    std::shared_ptr<World> world;
    std::vector<float> autosubdivisionLevels;      // scene.library.autosubdivisionLevels: {DistanceSquared, MaximumTriangleSize} pairs
    explicit Scene(const std::string& n);
};

// ---- the device ---------------------------------------------------------------------------------------
// Owns the t3d_ctx of one GPU; meshes are uploaded lazily and re-uploaded when their generation changes.
struct GeometryDevice {
    t3d_ctx* ctx = nullptr;
    int deviceId = 0;
    uint32_t nextVolumeId = 1;
    std::map<uint64_t, uint64_t> residentMeshes;   // mesh id -> generation key last uploaded
    explicit GeometryDevice(int device);
    ~GeometryDevice();
    void ensureMesh(Mesh& m);
    void ensureVolume(LightVolume& v);
    uint64_t frameSerial = 0;       // bumped by Camera::Render before t3d_frame_begin
    void ensureSkeleton(Skeleton& s);
    // animation tracks reordered to the skeleton's bones (channels find their bone by name, animation.go:618-672): device id
    uint64_t ensureAnimation(Animation& a, Skeleton& s);
    std::map<std::pair<uint64_t, uint64_t>, std::pair<uint64_t, uint64_t>> residentAnimations;   // (animation, skeleton) -> (device id, generation key)
    uint64_t nextAnimationId = 1;
};

// ---- camera.go ------------------------------------------------------------------------------------------
struct DebugInfo {   // camera.go:41-64
    int drawnParts = 0, totalParts = 0, batchedParts = 0, drawnTris = 0, totalTris = 0, lightCount = 0, activeLightCount = 0;
};

struct RenderRecord { Model* model; int part; int flush; uint64_t firstVertex; uint32_t vertexCount; uint32_t totalTris; float minDepth, maxDepth; int drawIndex; /* index among the submitted t3d_draws, -1 = ProcessVertices early-out */ };
struct FlushRecord { Model* model; int part; uint64_t firstVertex; uint32_t vertexCount; };

struct Camera : Node {
    bool RenderDepth = true, RenderNormals = false, PerspectiveCorrectedTextureMapping = false;
    int MaxLightCount = 0;                              // camera.go:223: 0 = every light
    std::shared_ptr<Node> MaxLightCountFocusPoint;      // camera.go:226: nullptr = the camera
    float VertexSnapping = 0, DepthMargin = 0.04f;
    float near_ = 0.1f, far_ = 100, fieldOfView = 60, orthoScale = 20;
    bool perspective = true;
    int width, height;
    Vector3 cameraForward{0, 0, 1}, cameraRight{1, 0, 0}, cameraUp{0, 1, 0};
    float sphereFactorTang = 0, sphereFactorY = 0, sphereFactorX = 0;
    DebugInfo debugInfo;
    std::shared_ptr<GeometryDevice> device;

    // sharding across ranks (object-wise, whole flushes): rank r keeps flush f when owner(f) == r
    int shardRank = 0, shardCount = 1;

    Camera(const std::string& n, int w, int h);     // camera.go:324
    NodeType Type() const override { return NodeType::Camera; }
    float AspectRatio() const { return (float)width / (float)height; }
    Matrix4 ViewMatrix();                           // camera.go:638
    Matrix4 Projection();                           // camera.go:651
    bool SphereInFrustum(Vector3 position, float radius);   // camera.go:933
    bool ModelInFrustum(Model& m);                  // camera.go:986
    void Clear();                                   // camera.go:1002 (DebugInfo reset + cached camera axes)
    void RenderScene(Scene& scene);                 // camera.go:1095
    std::map<std::string, uint32_t> lastBoneOffsets;   // skinned model name -> t3d_draw.bone_offset of the last Render (tests read the bones back)
    void RenderNodes(Scene& scene, Node& root);     // camera.go:1105
    void Render(Scene& scene, const std::vector<ILight*>& lights, const std::vector<Model*>& models);   // camera.go:1269
    // when false the lists stay on the device (t3d_frame_results_device)
    bool downloadResults = true;
    // Two frames in flight (t3d_set_async_download): Render returns once the frame's kernels are done and the lists' copy to the
    // host has been STARTED; the game loop goes on to the next frame and calls WaitDownload(1) before it hands the previous frame's
    // lists to DrawTriangles (camera.go:1829-2048), or WaitDownload(0) for the frame just rendered.
    bool AsyncDownload = false;
    // t3d_set_skip_unused_depth_list: with RenderDepth off the depth list is not copied to the host (flush never draws it then)
    bool SkipUnusedDepthList = false;
    void WaitDownload(int framesBack = 0);
    // runs of Models that differ only in position, scale and colour are submitted with t3d_submit_instances (the device
    // evaluates their Node.Transform); false: every render(rp) is a t3d_draw carrying its matrix
    bool UseInstancing = true;
    // with UseInstancing: DynamicBatch children that are instances of one another also get their ModelInFrustum test
    // (camera.go:2088-2093) on the device; false: the host tests every child, as the reference does
    bool DeviceFrustumCulling = true;

    // results of the last Render (library-owned memory, valid until the next Render)
    t3d_frame_output output{};
    t3d_frame_output prevOutput{};      // the Render before (AsyncDownload: its lists stay valid until the next Render's results)
    t3d_timings timings{};
    std::vector<RenderRecord> renders;
    std::vector<FlushRecord> flushes;
    std::vector<int> flushOwnerRank;        // for every flush of the frame (all ranks), the rank that ran it
    std::vector<std::vector<int>> shardDrawFlush;   // per rank: the flush of every draw that rank submitted, in its submit order (what the
                                                    // issuing rank needs to slice a shard's delivered lists by flush: t3d_gather_rank_stats)
    std::vector<int> localFlushIds;         // global flush ids of this rank's flushes, in order
    int launchCount = 0;
};

// Object-wise sharding: whole flushes, largest first onto the least-loaded rank (deterministic; every rank computes
// the same assignment from the same scene).  Returns owner rank per flush.
std::vector<int> AssignFlushesToRanks(const std::vector<uint64_t>& flushTriangles, int rankCount);

struct T3DError : std::runtime_error { int code; T3DError(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

}  // namespace tetra3d
