// tetra3d.cpp -- host-side mirror of the Tetra3D API (see tetra3d.hpp).  Build with -ffp-contract=off.
#include <functional>
#include <stdexcept>
#include "tetra3d.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <tuple>
#include <unordered_map>

namespace tetra3d {

using namespace t3d;

static uint64_t g_nextMeshId = 1;

// ---- math32 wrappers (math32/math32.go): float64 libm rounded to float32 ----------------------------
static inline float m32_tan(float x) { return (float)std::tan((double)x); }
static inline float m32_cos(float x) { return (float)std::cos((double)x); }
static inline float m32_sin(float x) { return (float)std::sin((double)x); }
static inline float m32_atan(float x) { return (float)std::atan((double)x); }
static const float M32_PI = (float)3.14159265358979323846;

Matrix4 NewMatrix4() { return m4_identity(); }
Matrix4 NewMatrix4Translate(float x, float y, float z) { Matrix4 r = m4_identity(); r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z; return r; }
Matrix4 NewMatrix4Scale(float x, float y, float z) { return m4_scale(x, y, z); }
Matrix4 NewMatrix4Rotate(float x, float y, float z, float angle) {       // matrix.go:100-128
    if (x == 0 && y == 0 && z == 0) y = 1;
    Matrix4 mat = m4_identity();
    Vector3 v = v3_unit(v3(x, y, z));
    float s = m32_sin(angle), c = m32_cos(angle), m = 1 - c;
    mat.m[0][0] = m * v.X * v.X + c;
    mat.m[0][1] = m * v.X * v.Y + v.Z * s;
    mat.m[0][2] = m * v.Z * v.X - v.Y * s;
    mat.m[1][0] = m * v.X * v.Y - v.Z * s;
    mat.m[1][1] = m * v.Y * v.Y + c;
    mat.m[1][2] = m * v.Y * v.Z + v.X * s;
    mat.m[2][0] = m * v.Z * v.X + v.Y * s;
    mat.m[2][1] = m * v.Y * v.Z - v.X * s;
    mat.m[2][2] = m * v.Z * v.Z + c;
    return mat;
}
Matrix4 NewProjectionMatrix4Perspective(float fovy, float near_, float far_, float viewWidth, float viewHeight) {   // matrix.go:588-604
    float aspect = viewWidth / viewHeight;
    float t = m32_tan(fovy * M32_PI / 360);
    float b = -t;
    float r = t * aspect;
    float l = -r;
    Matrix4 m; std::memset(&m, 0, sizeof m);
    m.m[0][0] = (2 * near_) / (r - l); m.m[0][2] = (r + l) / (r - l);
    m.m[1][1] = (2 * near_) / (t - b); m.m[1][2] = (t + b) / (t - b);
    m.m[2][2] = -((far_ + near_) / (far_ - near_)); m.m[2][3] = -((2 * far_ * near_) / (far_ - near_));
    m.m[3][2] = -1;
    return m;
}
Matrix4 NewProjectionMatrix4Orthographic(float near_, float far_, float right, float left, float top, float bottom) {   // matrix.go:609-616
    Matrix4 m; std::memset(&m, 0, sizeof m);
    m.m[0][0] = 2 / (right - left);
    m.m[1][1] = 2 / (top - bottom);
    m.m[2][2] = -2 / (far_ - near_);
    m.m[3][3] = 1;
    return m;
}
bool Matrix4Equals(const Matrix4& a, const Matrix4& b) {                  // matrix.go:511-522
    const float eps = 0.0001f;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (std::fabs(a.m[i][j] - b.m[i][j]) > eps) return false;
    return true;
}
static bool isIdentity(const Matrix4& a) { return Matrix4Equals(a, m4_identity()); }

// ---- Dimensions / Mesh ------------------------------------------------------------------------------
float Dimensions::MaxSpan() const { return v3_magnitude(v3_sub(Max, Min)); }
Vector3 Dimensions::Center() const { return v3((Max.X + Min.X) / 2, (Max.Y + Min.Y) / 2, (Max.Z + Min.Z) / 2); }

std::shared_ptr<Material> NewMaterial(const std::string& name) { auto m = std::make_shared<Material>(); m->name = name; return m; }

std::shared_ptr<Mesh> NewMesh(const std::string& name) { auto m = std::make_shared<Mesh>(); m->Name = name; m->id = g_nextMeshId++; return m; }

std::shared_ptr<MeshPart> Mesh::AddMeshPart(std::shared_ptr<Material> mat, int vs, int ve, int ts, int te) {
    auto p = std::make_shared<MeshPart>();
    p->mesh = this; p->material = std::move(mat);
    p->VertexIndexStart = vs; p->VertexIndexEnd = ve; p->TriangleStart = ts; p->TriangleEnd = te;
    MeshParts.push_back(p);
    return p;
}
void Mesh::UpdateDimensions() {                                           // mesh.go:719-751
    dimensions.Min = v3(FLT_MAX, FLT_MAX, FLT_MAX);
    dimensions.Max = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int i = 0; i < VertexCount(); i++) {
        float x = VertexPositions[3 * i], y = VertexPositions[3 * i + 1], z = VertexPositions[3 * i + 2];
        if (dimensions.Min.X > x) dimensions.Min.X = x;
        if (dimensions.Min.Y > y) dimensions.Min.Y = y;
        if (dimensions.Min.Z > z) dimensions.Min.Z = z;
        if (dimensions.Max.X < x) dimensions.Max.X = x;
        if (dimensions.Max.Y < y) dimensions.Max.Y = y;
        if (dimensions.Max.Z < z) dimensions.Max.Z = z;
    }
}
void Mesh::RecalculateCenters() {                                         // mesh.go:1980-1982
    const int nt = TriangleTotal();
    TriangleCenters.resize((size_t)nt * 3);
    for (int t = 0; t < nt; t++) {
        const float* a = &VertexPositions[3 * (size_t)TriangleIndices[3 * t]];
        const float* b = &VertexPositions[3 * (size_t)TriangleIndices[3 * t + 1]];
        const float* c = &VertexPositions[3 * (size_t)TriangleIndices[3 * t + 2]];
        TriangleCenters[3 * t + 0] = (a[0] + b[0] + c[0]) / 3;
        TriangleCenters[3 * t + 1] = (a[1] + b[1] + c[1]) / 3;
        TriangleCenters[3 * t + 2] = (a[2] + b[2] + c[2]) / 3;
    }
}
// calculateNormal (vector.go) + Mesh.AutoNormal (mesh.go:755-764)
void Mesh::AutoNormal() {
    const int nt = TriangleTotal();
    for (int t = 0; t < nt; t++) {
        uint32_t ia = TriangleIndices[3 * t], ib = TriangleIndices[3 * t + 1], ic = TriangleIndices[3 * t + 2];
        Vector3 p1 = v3(VertexPositions[3 * ia], VertexPositions[3 * ia + 1], VertexPositions[3 * ia + 2]);
        Vector3 p2 = v3(VertexPositions[3 * ib], VertexPositions[3 * ib + 1], VertexPositions[3 * ib + 2]);
        Vector3 p3 = v3(VertexPositions[3 * ic], VertexPositions[3 * ic + 1], VertexPositions[3 * ic + 2]);
        Vector3 n = v3_unit(v3_cross(v3_sub(p2, p1), v3_sub(p3, p2)));
        for (uint32_t v : {ia, ib, ic}) { VertexNormals[3 * v] = n.X; VertexNormals[3 * v + 1] = n.Y; VertexNormals[3 * v + 2] = n.Z; }
    }
}
void Mesh::SetShapeKeyWeights(const std::vector<float>& w) {
    if (w.size() != ShapeKeyWeights.size()) throw T3DError(T3D_E_INVALID, "SetShapeKeyWeights: key count mismatch");
    ShapeKeyWeights = w; shapeKeyWeightGeneration++;
}

// mesh.go:1251-1342: vertex and index tables of the unit cube
std::shared_ptr<Mesh> NewCubeMesh(float w, float h, float d) {
    w /= 2; h /= 2; d /= 2;
    auto mesh = NewMesh("Cube");
    const float V[24][5] = {
        {-w, h, -d, 0, 0}, {w, h, d, 1, 1}, {w, h, -d, 1, 0}, {-w, h, d, 0, 1},           // top
        {w, -h, -d, 1, 0}, {w, -h, d, 1, 1}, {-w, -h, -d, 0, 0}, {-w, -h, d, 0, 1},       // bottom
        {-w, h, d, 0, 0}, {w, -h, d, 1, 1}, {w, h, d, 1, 0}, {-w, -h, d, 0, 1},           // front
        {w, h, -d, 1, 0}, {w, -h, -d, 1, 1}, {-w, h, -d, 0, 0}, {-w, -h, -d, 0, 1},       // back
        {w, h, -d, 1, 0}, {w, h, d, 1, 1}, {w, -h, -d, 0, 0}, {w, -h, d, 0, 1},           // right
        {-w, -h, -d, 0, 0}, {-w, h, d, 1, 1}, {-w, h, -d, 1, 0}, {-w, -h, d, 0, 1},       // left
    };
    const uint32_t I[36] = {0, 1, 2, 3, 1, 0, 4, 5, 6, 6, 5, 7, 8, 9, 10, 11, 9, 8,
                            12, 13, 14, 14, 13, 15, 16, 17, 18, 18, 17, 19, 20, 21, 22, 23, 21, 20};
    for (auto& v : V) {
        mesh->VertexPositions.insert(mesh->VertexPositions.end(), {v[0], v[1], v[2]});
        mesh->VertexNormals.insert(mesh->VertexNormals.end(), {0.f, 0.f, 0.f});
        mesh->VertexUVs.insert(mesh->VertexUVs.end(), {v[3], v[4]});
    }
    mesh->TriangleIndices.assign(I, I + 36);
    mesh->AddMeshPart(NewMaterial("Cube"), 0, 24, 0, 11);
    mesh->RecalculateCenters();
    mesh->UpdateDimensions();
    mesh->AutoNormal();
    return mesh;
}

// ---- Node ---------------------------------------------------------------------------------------------
Node::Node(const std::string& n) : name(n) { rotation = m4_identity(); cachedTransform = m4_identity(); }

Matrix4 Node::Transform() {                                               // node.go:572-608
    if (!isTransformDirty) return cachedTransform;
    Matrix4 transform = NewMatrix4Scale(scale.X, scale.Y, scale.Z);
    transform = m4_mult(transform, rotation);
    transform = m4_mult(transform, NewMatrix4Translate(position.X, position.Y, position.Z));
    if (parent) transform = m4_mult(transform, parent->Transform());
    cachedTransform = transform;
    isTransformDirty = false;
    onTransformUpdate();
    for (auto& c : children) c->Transform();
    return transform;
}
void Node::dirtyTransform() { for (auto& c : children) c->dirtyTransform(); isTransformDirty = true; }
Vector3 Node::WorldPosition() { Matrix4 t = Transform(); return v3(t.m[3][0], t.m[3][1], t.m[3][2]); }
Matrix4 Node::WorldRotation() {                                           // node.go:916-938
    bool need = false;
    for (Node* p = parent; p; p = p->parent) if (!isIdentity(p->rotation)) { need = true; break; }
    if (need) return m4_decompose_rotation(Transform());
    return rotation;
}
void Node::SetLocalPosition(float x, float y, float z) { position = v3(x, y, z); dirtyTransform(); }
void Node::SetLocalScale(float x, float y, float z) { scale = v3(x, y, z); dirtyTransform(); }
void Node::SetLocalRotation(const Matrix4& r) { rotation = r; dirtyTransform(); }
void Node::AddChildren(const std::vector<std::shared_ptr<Node>>& kids) {
    for (auto& k : kids) { k->parent = this; children.push_back(k); k->dirtyTransform(); }
}
void Node::SetVisible(bool v, bool recursive) { if (recursive) for (auto& c : children) c->SetVisible(v, true); visible = v; }
void Node::ForEachChild(const std::function<void(Node*)>& f) { for (auto& c : children) { f(c.get()); c->ForEachChild(f); } }

// ---- Model --------------------------------------------------------------------------------------------
Model::Model(const std::string& n, std::shared_ptr<Mesh> m) : Node(n), mesh(std::move(m)) {
    if (mesh) sphereRadius = mesh->dimensions.MaxSpan() / 2;            // model.go:83-87
}
void Model::onTransformUpdate() {                                         // model.go:142-182
    if (!mesh) return;
    Vector3 pos, sc; Matrix4 rot;
    m4_decompose(cachedTransform, pos, sc, rot);
    Vector3 center = v3(0, 0, 0);
    if (!skinned) center = mesh->dimensions.Center();
    center = m4_multvec(rot, center);
    pos.X += center.X * sc.X;
    pos.Y += center.Y * sc.Y;
    pos.Z += center.Z * sc.Z;
    spherePosition = pos;
    Dimensions dim = mesh->dimensions;
    dim.Min.X *= sc.X; dim.Min.Y *= sc.Y; dim.Min.Z *= sc.Z;
    dim.Max.X *= sc.X; dim.Max.Y *= sc.Y; dim.Max.Z *= sc.Z;
    sphereRadius = dim.MaxSpan() / 2;
}
bool Model::isTransparent(const MeshPart& part) const {                   // model.go:1274-1282
    const Material* mat = part.material.get();
    if (mat) {
        bool matT = mat->TransparencyMode == TransparencyModeTransparent || !mat->BlendSourceOver ||
                    (mat->TransparencyMode == TransparencyModeAuto && mat->Color.A < 0.999f);
        bool modelT = mat->TransparencyMode != TransparencyModeOpaque && Color.A < 0.999f;
        return matT || modelT;
    }
    return Color.A < 0.999f;
}
void Model::DynamicBatchAdd(int meshPartIndex, const std::vector<std::shared_ptr<Model>>& models) {   // model.go:217-247
    for (auto& other : models) {
        if (other.get() == this) continue;
        auto& slice = DynamicBatchModels[meshPartIndex];
        if (std::find(slice.begin(), slice.end(), other) != slice.end()) continue;
        slice.push_back(other);
        other->DynamicBatchOwner = this;
    }
}

Scene::Scene(const std::string& n) : Name(n), Root(std::make_shared<Node>("Root")), world(std::make_shared<World>()) {}

// ---- GeometryDevice -----------------------------------------------------------------------------------
static void check(t3d_ctx* ctx, int rc, const char* what) {
    if (rc != T3D_OK) throw T3DError(rc, std::string(what) + ": " + t3d_last_error(ctx));
}

GeometryDevice::GeometryDevice(int device) : deviceId(device) {
    int rc = t3d_create(&ctx, device);
    if (rc != T3D_OK) throw T3DError(rc, std::string("t3d_create: ") + t3d_last_error(nullptr));
}
GeometryDevice::~GeometryDevice() { if (ctx) t3d_destroy(ctx); }

void GeometryDevice::ensureMesh(Mesh& m) {
    const uint64_t genKey = ((uint64_t)m.generation << 32) | m.shapeKeyWeightGeneration;
    auto res = residentMeshes.find(m.id);
    if (res != residentMeshes.end() && res->second == genKey) return;
    std::vector<t3d_part_range> parts(m.MeshParts.size());
    for (size_t i = 0; i < parts.size(); i++) {
        const MeshPart& p = *m.MeshParts[i];
        parts[i].vtx_start = (uint32_t)p.VertexIndexStart; parts[i].vtx_end = (uint32_t)p.VertexIndexEnd;
        parts[i].tri_start = (uint32_t)p.TriangleStart; parts[i].tri_end = (uint32_t)(p.TriangleEnd + 1);
    }
    t3d_mesh_desc d; std::memset(&d, 0, sizeof d);
    d.n_verts = (uint32_t)m.VertexCount();
    d.positions = m.VertexPositions.data(); d.normals = m.VertexNormals.data(); d.uvs = m.VertexUVs.data();
    d.colors = (m.VertexActiveColorChannel >= 0 && !m.VertexColors.empty()) ? m.VertexColors.data() : nullptr;
    if (!m.VertexWeights.empty()) {
        d.weights = m.VertexWeights.data(); d.bones = m.VertexBones.data();
        d.n_influences = m.VertexInfluences.empty() ? nullptr : m.VertexInfluences.data();
    }
    d.n_tris = (uint32_t)m.TriangleTotal();
    d.indices = m.TriangleIndices.data(); d.centers = m.TriangleCenters.data();
    d.n_parts = (uint32_t)parts.size(); d.parts = parts.data();
    d.max_span = m.dimensions.MaxSpan();
    d.n_shape_keys = (uint32_t)m.ShapeKeyWeights.size();
    d.key_positions = m.ShapeKeyPositions.data(); d.key_normals = m.ShapeKeyNormals.data();
    if (m.autoSubdivide) {
        const size_t t = (size_t)m.TriangleTotal();
        if (m.TriangleSubBase.size() != t || m.TriangleSubLevel.size() != t || m.TriangleLevelCount.size() != t || m.TriangleMaxSpan.size() != t || m.TriangleNormals.size() != 3 * t)
            throw T3DError(T3D_E_INVALID, "Mesh: auto-subdivision arrays do not match the triangle count");
        d.tri_sub_base = m.TriangleSubBase.data(); d.tri_sub_level = m.TriangleSubLevel.data(); d.tri_n_levels = m.TriangleLevelCount.data();
        d.tri_max_span = m.TriangleMaxSpan.data(); d.tri_normal = m.TriangleNormals.data();
    }
    // t3d_mesh_upload is a no-op when (id, generation) is already resident
    check(ctx, t3d_mesh_upload(ctx, m.id, m.generation, &d), "t3d_mesh_upload");
    if (d.n_shape_keys) check(ctx, t3d_mesh_set_shape_key_weights(ctx, m.id, m.ShapeKeyWeights.data(), d.n_shape_keys), "t3d_mesh_set_shape_key_weights");
    residentMeshes[m.id] = genKey;
}

void LightVolume::CellSetUseLights(int x, int y, int z, const std::vector<std::shared_ptr<ILight>>& lights) {   // light.go:1508-1514
    if (x < 0 || y < 0 || z < 0 || x >= cellCount[0] || y >= cellCount[1] || z >= cellCount[2]) throw T3DError(T3D_E_INVALID, "light volume cell out of range");
    auto& cell = cellUseLights[(size_t)x + (size_t)y * cellCount[0] + (size_t)z * cellCount[0] * cellCount[1]];
    for (auto& l : lights) {
        if (l->Type() == NodeType::LightVolume) throw T3DError(T3D_E_UNSUPPORTED, "a LightVolume used by a LightVolume cell is not evaluated on the device");
        if (std::find(useLightsRenderSet.begin(), useLightsRenderSet.end(), l) == useLightsRenderSet.end()) useLightsRenderSet.push_back(l);
        cell.push_back(l);
    }
    generation++;
}
void LightVolume::CellClearUseLights(int x, int y, int z) {                                                      // light.go:1516-1518
    cellUseLights.erase((size_t)x + (size_t)y * cellCount[0] + (size_t)z * cellCount[0] * cellCount[1]);
    generation++;
}

namespace { t3d_light toLight(GeometryDevice& dev, ILight* l); }

void GeometryDevice::ensureVolume(LightVolume& v) {
    if (v.volumeId == 0) v.volumeId = nextVolumeId++;
    if (!v.useLightsRenderSet.empty() && v.useLightsFrame != frameSerial) {
        // LightVolume.beginRender / beginModel over the render set (light.go:565-580): this frame's world-space records
        std::vector<t3d_light> recs;
        for (auto& l : v.useLightsRenderSet) recs.push_back(toLight(*this, l.get()));
        if (v.uploadedGeneration == v.generation)       // else: submitted below, after the upload
            check(ctx, t3d_volume_use_lights(ctx, v.volumeId, recs.data(), (uint32_t)recs.size()), "t3d_volume_use_lights");
    }
    if (v.uploadedGeneration == v.generation) { v.useLightsFrame = frameSerial; return; }
    t3d_volume_desc d; std::memset(&d, 0, sizeof d);
    d.dim_min[0] = v.dimensions.Min.X; d.dim_min[1] = v.dimensions.Min.Y; d.dim_min[2] = v.dimensions.Min.Z;
    d.dim_max[0] = v.dimensions.Max.X; d.dim_max[1] = v.dimensions.Max.Y; d.dim_max[2] = v.dimensions.Max.Z;
    for (int i = 0; i < 3; i++) { d.cell_size[i] = v.cellSize[i]; d.cell_count[i] = (uint32_t)v.cellCount[i]; }
    d.offset[0] = v.offset.X; d.offset[1] = v.offset.Y; d.offset[2] = v.offset.Z;
    d.light_bias[0] = v.lightBias.X; d.light_bias[1] = v.lightBias.Y; d.light_bias[2] = v.lightBias.Z;
    d.cell_colors = v.cellColors.data();
    std::vector<uint32_t> start, index;
    if (!v.cellUseLights.empty()) {
        const size_t cells = (size_t)v.cellCount[0] * v.cellCount[1] * v.cellCount[2];
        start.assign(cells + 1, 0);
        for (size_t c = 0; c < cells; c++) {
            auto it = v.cellUseLights.find(c);
            if (it != v.cellUseLights.end())
                for (auto& l : it->second)
                    index.push_back((uint32_t)(std::find(v.useLightsRenderSet.begin(), v.useLightsRenderSet.end(), l) - v.useLightsRenderSet.begin()));
            start[c + 1] = (uint32_t)index.size();
        }
        if (index.empty()) index.push_back(0);
        d.cell_light_start = start.data(); d.cell_light_index = index.data();
    }
    check(ctx, t3d_volume_upload(ctx, v.volumeId, &d), "t3d_volume_upload");
    v.uploadedGeneration = v.generation;
    if (!v.useLightsRenderSet.empty()) {
        std::vector<t3d_light> recs;
        for (auto& l : v.useLightsRenderSet) recs.push_back(toLight(*this, l.get()));
        check(ctx, t3d_volume_use_lights(ctx, v.volumeId, recs.data(), (uint32_t)recs.size()), "t3d_volume_use_lights");
    }
    v.useLightsFrame = frameSerial;
}

void GeometryDevice::ensureSkeleton(Skeleton& s) {
    if (s.uploadedGeneration == s.generation) return;
    t3d_skeleton_desc d; std::memset(&d, 0, sizeof d);
    d.n_bones = (uint32_t)s.BoneCount();
    d.parent = s.parent.data(); d.inverse_bind = s.inverseBind.data();
    d.rest_position = s.restPosition.data(); d.rest_scale = s.restScale.data(); d.rest_rotation = s.restRotation.data();
    check(ctx, t3d_skeleton_upload(ctx, s.id, &d), "t3d_skeleton_upload");
    s.uploadedGeneration = s.generation;
}

uint64_t GeometryDevice::ensureAnimation(Animation& a, Skeleton& s) {
    const auto key = std::make_pair(a.id, s.id);
    const uint64_t genKey = ((uint64_t)a.generation << 32) | s.generation;
    auto it = residentAnimations.find(key);
    if (it != residentAnimations.end() && it->second.second == genKey) return it->second.first;
    // every bone gets the tracks of the channel carrying its name (assignChannels, animation.go:618-672)
    const size_t nb = (size_t)s.BoneCount();
    std::vector<t3d_track> tracks(3 * nb);
    std::vector<float> times, values;
    for (size_t b = 0; b < nb; b++) {
        const AnimationChannel* ch = nullptr;
        for (auto& c : a.channels) if (c->Name == s.boneNames[b]) { ch = c.get(); break; }
        for (int t = 0; t < 3; t++) {
            t3d_track& T = tracks[3 * b + t];
            T.first_key = (uint32_t)times.size(); T.n_keys = 0; T.interpolation = T3D_INTERPOLATION_LINEAR;
            const AnimationTrack* tr = ch ? ch->Tracks[t].get() : nullptr;
            if (!tr) continue;
            T.n_keys = (uint32_t)tr->Keyframes.size(); T.interpolation = (uint32_t)tr->Interpolation;
            for (const Keyframe& k : tr->Keyframes) { times.push_back(k.Time); values.insert(values.end(), k.Data, k.Data + 4); }
        }
    }
    t3d_animation_desc d; std::memset(&d, 0, sizeof d);
    d.n_bones = (uint32_t)nb; d.tracks = tracks.data(); d.n_keys = (uint32_t)times.size(); d.key_times = times.data(); d.key_values = values.data();
    const uint64_t devId = it != residentAnimations.end() ? it->second.first : nextAnimationId++;
    check(ctx, t3d_animation_upload(ctx, devId, &d), "t3d_animation_upload");
    residentAnimations[key] = {devId, genKey};
    return devId;
}

// ---- animation.go -------------------------------------------------------------------------------------
static uint64_t g_nextAnimationId = 1, g_nextSkeletonId = 1;
Animation::Animation(const std::string& n) : name(n), id(g_nextAnimationId++) {}
std::shared_ptr<AnimationChannel> Animation::AddChannel(const std::string& chName) {      // animation.go:238-242
    auto c = std::make_shared<AnimationChannel>();
    c->Name = chName;
    for (auto& old : channels) if (old->Name == chName) { old = c; return c; }            // a map in Go: same name replaces
    channels.push_back(c);
    return c;
}
std::shared_ptr<AnimationTrack> AnimationChannel::AddTrack(int trackType) {                // animation.go:188-192
    if (trackType < 0 || trackType > 2) throw T3DError(T3D_E_INVALID, "AddTrack: unknown track type");
    Tracks[trackType] = std::make_shared<AnimationTrack>();
    Tracks[trackType]->Type = trackType;
    return Tracks[trackType];
}

Skeleton::Skeleton() : id(g_nextSkeletonId++) {}
std::shared_ptr<Skeleton> Skeleton::FromJoints(const std::vector<std::shared_ptr<Node>>& joints, const std::vector<float>& inverseBind16) {
    if (inverseBind16.size() != joints.size() * 16) throw T3DError(T3D_E_INVALID, "Skeleton: 16 inverse-bind floats per joint");
    auto sk = std::make_shared<Skeleton>();
    sk->inverseBind = inverseBind16;
    for (size_t j = 0; j < joints.size(); j++) {
        Node& n = *joints[j];
        int32_t par = -1;
        for (size_t k = 0; k < joints.size(); k++) if (joints[k].get() == n.parent) { par = (int32_t)k; break; }
        if (par >= (int32_t)j) throw T3DError(T3D_E_INVALID, "Skeleton: a parent bone must precede its children in the joint list");
        sk->parent.push_back(par);
        sk->boneNames.push_back(n.name);
        sk->restPosition.insert(sk->restPosition.end(), {n.position.X, n.position.Y, n.position.Z});
        sk->restScale.insert(sk->restScale.end(), {n.scale.X, n.scale.Y, n.scale.Z});
        float r[16]; m4_store(n.rotation, r);
        sk->restRotation.insert(sk->restRotation.end(), r, r + 16);
    }
    return sk;
}

int64_t AnimationPlayer::now() const {
    if (NowMs) return NowMs();
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void AnimationPlayer::Play(std::shared_ptr<Animation> anim) {                               // animation.go:517-566
    if (!anim) return;
    if (animation == anim && playing) return;
    animation = std::move(anim);
    playing = true;
    loopCount = 0;
    playhead = PlaySpeed > 0 ? 0.0f : animation->length;
    if (BlendTime > 0) {                    // animation.go:548-554: prevAnimatedProperties = copy of currentProperties
        if (currentHas) prevChain = curChain; else prevChain.clear();
        blendStarted = true;
        blendStartMs = now();
    }
    // assignChannels (animation.go:618-672) runs at the next updateValues because ChannelsUpdated is false now, and clears currentProperties
    currentHas = false; curChain.clear();
}

void AnimationPlayer::Update(float dt) {                                                    // animation.go:810-824
    finished = false;
    if (!playing && blendStarted) { blendStarted = false; prevChain.clear(); }
    forceUpdate(dt);
}

void AnimationPlayer::forceUpdate(float dt) {                                               // animation.go:837 -> updateValues 674-806
    if (!animation) return;
    sampledPlayhead = playhead;             // the tracks are sampled here (animation.go:686-712): on the device, at the next Render
    sampled = true;
    // the blend this frame's targets go through (animation.go:852-907); `prevExists` per node is the device's business
    sampledChain.clear();
    if (blendStarted && !prevChain.empty()) {
        float bp = (float)(now() - blendStartMs) / (BlendTime * 1000.0f);
        if (bp > 1) bp = 1;
        sampledChain = prevChain;
        sampledChain.push_back(BlendLevel{animation, sampledPlayhead, bp});
        if (bp == 1) { blendStarted = false; prevChain.clear(); }                           // animation.go:899-902
    } else {
        sampledChain.push_back(BlendLevel{animation, sampledPlayhead, 0.0f});
    }
    if (BlendTime > 0) { curChain = sampledChain; currentHas = true; }                     // animation.go:909-937
    playhead += dt * PlaySpeed;
    float ph = playhead;
    prevPlayhead = ph;
    justLooped = false;
    const float length = animation->length;
    if (animation->loopMode == AnimationLoopModeLoop && (ph >= length || ph < 0)) {
        if (ph >= length) playhead -= length;
        if (ph < 0) playhead += length;
        justLooped = true;
        loopCount++;
        finished = true;
    } else if (animation->loopMode == AnimationLoopModePingPong && (ph >= length || ph < 0)) {
        if (ph >= length) playhead -= playhead - length;
        if (ph < 0) { playhead = -playhead; loopCount++; }
        finished = true;
        PlaySpeed *= -1;
    } else if (animation->loopMode == AnimationLoopModeOneshot && ((ph >= length && PlaySpeed > 0) || (ph <= 0 && PlaySpeed < 0))) {
        finished = true;
        playing = false;
    }
}

std::vector<int> AssignFlushesToRanks(const std::vector<uint64_t>& flushTris, int rankCount) {
    const int n = (int)flushTris.size();
    std::vector<int> owner(n, 0);
    if (rankCount <= 1) return owner;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return flushTris[a] > flushTris[b]; });
    std::vector<uint64_t> load(rankCount, 0);
    for (int f : order) {
        int best = 0;
        for (int r = 1; r < rankCount; r++) if (load[r] < load[best]) best = r;
        owner[f] = best; load[best] += flushTris[f];
    }
    return owner;
}

// ---- Camera -------------------------------------------------------------------------------------------
Camera::Camera(const std::string& n, int w, int h) : Node(n), width(w), height(h) {}

void Camera::WaitDownload(int framesBack) {
    if (!device) return;
    check(device->ctx, t3d_frame_download_wait(device->ctx, (uint32_t)framesBack), "t3d_frame_download_wait");
}

Matrix4 Camera::ViewMatrix() {                                            // camera.go:638-648
    Vector3 camPos = v3_invert(WorldPosition());
    Matrix4 transform = NewMatrix4Translate(camPos.X, camPos.Y, camPos.Z);
    return m4_mult(transform, m4_transposed(WorldRotation()));
}
Matrix4 Camera::Projection() {                                            // camera.go:651-679
    float angle = fieldOfView * 3.1415f / 360;                            // the reference's literal (camera.go:660)
    sphereFactorTang = m32_tan(angle);
    sphereFactorY = 1.0f / m32_cos(angle);
    sphereFactorX = 1.0f / m32_cos(m32_atan(sphereFactorTang * AspectRatio()));
    if (perspective) return NewProjectionMatrix4Perspective(fieldOfView, near_, far_, (float)width, (float)height);
    float asr = (float)height / (float)width;
    return NewProjectionMatrix4Orthographic(near_, far_, 1 * orthoScale, -1 * orthoScale, asr * orthoScale, -asr * orthoScale);
}
bool Camera::SphereInFrustum(Vector3 position, float radius) {            // camera.go:933-983
    Vector3 diff = v3_sub(position, WorldPosition());
    float pcZ = v3_dot(diff, v3_invert(cameraForward));
    if (pcZ > far_ + radius || pcZ < near_ - radius) return false;
    if (perspective) {
        float d = sphereFactorY * radius;
        pcZ *= sphereFactorTang;
        float pcY = v3_dot(diff, cameraUp);
        if (pcY > pcZ + d || pcY < -pcZ - d) return false;
        pcZ *= AspectRatio();
        d = sphereFactorX * radius;
        float pcX = v3_dot(diff, cameraRight);
        if (pcX > pcZ + d || pcX < -pcZ - d) return false;
    } else {
        float w = orthoScale;
        float h = w / AspectRatio();
        float pcY = v3_dot(diff, cameraUp);
        if (-h / 2 - radius > pcY || pcY > h / 2 + radius) return false;
        float pcX = v3_dot(diff, cameraRight);
        if (-w / 2 - radius > pcX || pcX > w / 2 + radius) return false;
    }
    return true;
}
bool Camera::ModelInFrustum(Model& m) {                                   // camera.go:986-989
    m.Transform();                  // make sure the transform (and the culling sphere with it) is up to date
    return SphereInFrustum(m.spherePosition, m.sphereRadius);
}
void Camera::Clear() {                                                    // camera.go:1018-1090
    debugInfo = DebugInfo();
    Matrix4 rot = WorldRotation();
    cameraForward = v3_unit(v3(rot.m[2][0], rot.m[2][1], rot.m[2][2]));
    cameraRight = v3_unit(v3(rot.m[0][0], rot.m[0][1], rot.m[0][2]));
    cameraUp = v3_unit(v3(rot.m[1][0], rot.m[1][1], rot.m[1][2]));
}
void Camera::RenderScene(Scene& scene) { RenderNodes(scene, *scene.Root); }   // camera.go:1095

void Camera::RenderNodes(Scene& scene, Node& root) {                      // camera.go:1105-1216
    std::vector<Model*> meshes; std::vector<ILight*> lights;
    root.ForEachChild([&](Node* n) {
        switch (n->Type()) {
        case NodeType::Model: { Model* m = static_cast<Model*>(n); if (!m->DynamicBatchOwner) meshes.push_back(m); break; }
        case NodeType::AmbientLight: case NodeType::PointLight: case NodeType::DirectionalLight: case NodeType::LightVolume:
            lights.push_back(static_cast<ILight*>(n)); break;
        default: break;
        }
    });
    Render(scene, lights, meshes);
}

namespace {
struct renderPair { Model* model; int part; float depth; int order; };   // camera.go:1249-1254

// sort.SliceStable as Go's package sort implements it (zsortfunc.go: stable_func, insertionSort_func, symMerge_func,
// rotate_func, swapRange_func): insertion sort on blocks of 20, then bottom-up SymMerge.  The exact sequence of
// less / swap calls matters when the comparator is inconsistent (Camera.MaxLightCount, camera.go:1484-1495).
template <class Less, class Swap> void goInsertionSort(int a, int b, Less& less, Swap& swap) {
    for (int i = a + 1; i < b; i++)
        for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
}
template <class Swap> void goSwapRange(int a, int b, int n, Swap& swap) { for (int i = 0; i < n; i++) swap(a + i, b + i); }
template <class Swap> void goRotate(int a, int m, int b, Swap& swap) {
    int i = m - a, j = b - m;
    while (i != j) {
        if (i > j) { goSwapRange(m - i, m, j, swap); i -= j; }
        else { goSwapRange(m - i, m + j - i, i, swap); j -= i; }
    }
    goSwapRange(m - i, m, i, swap);
}
template <class Less, class Swap> void goSymMerge(int a, int m, int b, Less& less, Swap& swap) {
    if (m - a == 1) {
        int i = m, j = b;
        while (i < j) { const int h = (int)(((unsigned)i + (unsigned)j) >> 1); if (less(h, a)) i = h + 1; else j = h; }
        for (int k = a; k < i - 1; k++) swap(k, k + 1);
        return;
    }
    if (b - m == 1) {
        int i = a, j = m;
        while (i < j) { const int h = (int)(((unsigned)i + (unsigned)j) >> 1); if (!less(m, h)) i = h + 1; else j = h; }
        for (int k = m; k > i; k--) swap(k, k - 1);
        return;
    }
    const int mid = (int)(((unsigned)a + (unsigned)b) >> 1);
    const int n = mid + m;
    int start, r;
    if (m > mid) { start = n - b; r = mid; } else { start = a; r = m; }
    const int p = n - 1;
    while (start < r) {
        const int c = (int)(((unsigned)start + (unsigned)r) >> 1);
        if (!less(p - c, c)) start = c + 1; else r = c;
    }
    const int end = n - start;
    if (start < m && m < end) goRotate(start, m, end, swap);
    if (a < start && start < mid) goSymMerge(a, start, mid, less, swap);
    if (mid < end && end < b) goSymMerge(mid, end, b, less, swap);
}
template <class Less, class Swap> void goStableSort(int n, Less less, Swap swap) {
    int blockSize = 20;
    int a = 0, b = blockSize;
    while (b <= n) { goInsertionSort(a, b, less, swap); a = b; b += blockSize; }
    goInsertionSort(a, n, less, swap);
    while (blockSize < n) {
        a = 0; b = 2 * blockSize;
        while (b <= n) { goSymMerge(a, a + blockSize, b, less, swap); a = b; b += 2 * blockSize; }
        const int m = a + blockSize;
        if (m < n) goSymMerge(a, m, n, less, swap);
        blockSize *= 2;
    }
}

template <class T, class Less> void stableSort(std::vector<T>& v, Less less) { std::stable_sort(v.begin(), v.end(), less); }

t3d_light toLight(GeometryDevice& dev, ILight* l) {
    t3d_light L; std::memset(&L, 0, sizeof L);
    L.color[0] = l->color.R; L.color[1] = l->color.G; L.color[2] = l->color.B; L.energy = l->energy;
    switch (l->Type()) {
    case NodeType::AmbientLight: L.type = T3D_LIGHT_AMBIENT; break;
    case NodeType::PointLight: {
        L.type = T3D_LIGHT_POINT;
        Vector3 p = l->WorldPosition();
        L.position[0] = p.X; L.position[1] = p.Y; L.position[2] = p.Z;
        L.range = static_cast<PointLight*>(l)->lightRange;
        break;
    }
    case NodeType::DirectionalLight: {
        L.type = T3D_LIGHT_DIRECTIONAL;
        Matrix4 r = l->WorldRotation();                                   // light.go:322-324
        Vector3 f = v3_unit(v3(r.m[2][0], r.m[2][1], r.m[2][2]));
        L.forward[0] = f.X; L.forward[1] = f.Y; L.forward[2] = f.Z;
        break;
    }
    case NodeType::LightVolume: {
        LightVolume* v = static_cast<LightVolume*>(l);
        dev.ensureVolume(*v);
        L.type = T3D_LIGHT_VOLUME; L.volume = v->volumeId;
        break;
    }
    default: throw T3DError(T3D_E_INVALID, "node is not a light");
    }
    return L;
}
}  // namespace

void Camera::Render(Scene& scene, const std::vector<ILight*>& lights, const std::vector<Model*>& models) {   // camera.go:1269-2114
    if (!device) throw T3DError(T3D_E_STATE, "Camera has no GeometryDevice: the geometry stage has no CPU path");
    static const bool prof = getenv("T3D_HOST_PROFILE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!prof) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[host] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    t3d_ctx* ctx = device->ctx;
    World* world = scene.world.get();

    // sceneLights (camera.go:1275-1303)
    std::vector<ILight*> sceneLights;
    if (world) {
        debugInfo.lightCount++;
        if (world->LightingOn && world->ambientLight->IsVisible()) sceneLights.push_back(world->ambientLight.get());
    }
    for (ILight* l : lights) {
        debugInfo.lightCount++;
        if ((!world || world->LightingOn) && l->IsVisible()) sceneLights.push_back(l);
    }
    if (MaxLightCount > 0) {                                               // camera.go:1474-1497
        const Vector3 pos = MaxLightCountFocusPoint ? MaxLightCountFocusPoint->WorldPosition() : WorldPosition();
        // The reference's comparator is not a strict weak ordering (every ambient light sorts before every other one,
        // itself included), so std::stable_sort may not be used: goStableSort replays Go's sort.SliceStable.
        goStableSort((int)sceneLights.size(),
                     [&](int i, int j) {
                         const NodeType ti = sceneLights[i]->Type(), tj = sceneLights[j]->Type();
                         if (ti == NodeType::AmbientLight) return true;
                         else if (tj == NodeType::AmbientLight) return false;
                         else if (ti == NodeType::DirectionalLight) return true;
                         else if (tj == NodeType::DirectionalLight) return false;
                         return v3_dist2(pos, sceneLights[i]->WorldPosition()) < v3_dist2(pos, sceneLights[j]->WorldPosition());
                     },
                     [&](int i, int j) { std::swap(sceneLights[i], sceneLights[j]); });
        if ((size_t)MaxLightCount < sceneLights.size()) sceneLights.resize((size_t)MaxLightCount);
    }
    for (ILight* l : sceneLights) if (world && world->LightingOn && l->IsVisible()) debugInfo.activeLightCount++;   // camera.go:1499-1505

    Matrix4 proj = Projection();
    Matrix4 view = ViewMatrix();
    Matrix4 vpMatrix = m4_mult(view, proj);                                // camera.go:1329
    Vector3 cameraPos = WorldPosition();

    std::vector<renderPair> solids, transparents;
    for (Model* model : models) {                                          // camera.go:1343-1457
        if (!model->visible || model->DynamicBatchOwner) continue;
        if (!model->DynamicBatcher()) {
            if (model->mesh) debugInfo.totalParts += (int)model->mesh->MeshParts.size();
            if (model->FrustumCulling && !ModelInFrustum(*model)) continue;
            if (!model->mesh) continue;
            for (size_t pi = 0; pi < model->mesh->MeshParts.size(); pi++) {
                MeshPart& mp = *model->mesh->MeshParts[pi];
                if (!mp.isVisible()) continue;
                renderPair rp{model, (int)pi, v3_dist2(cameraPos, model->WorldPosition()), 0};
                if (mp.material) rp.order = mp.material->RenderOrder;
                if (model->isTransparent(mp)) transparents.push_back(rp); else solids.push_back(rp);
            }
        } else {
            bool transparent = false;
            // Go ranges over a map here (random order); the contract is ascending MeshPart index.
            for (auto& kv : model->DynamicBatchModels) {
                if (!model->mesh || kv.first >= (int)model->mesh->MeshParts.size()) continue;
                MeshPart& meshPart = *model->mesh->MeshParts[kv.first];
                if (!meshPart.isVisible()) continue;
                for (auto& child : kv.second) {
                    if (!child->visible) continue;
                    if (!transparent)
                        for (auto& mp : child->mesh->MeshParts)
                            if (mp->isVisible() && child->isTransparent(*mp)) { transparent = true; break; }
                }
                renderPair rp{model, kv.first, 0, 0};
                if (transparent || !RenderDepth) rp.depth = v3_dist2(cameraPos, model->WorldPosition());
                if (meshPart.material) rp.order = meshPart.material->RenderOrder;
                if (transparent) transparents.push_back(rp); else solids.push_back(rp);
                debugInfo.totalParts += (int)kv.second.size();
            }
        }
    }
    if (!RenderDepth)                                                      // camera.go:1460-1469
        // not a strict weak ordering once RenderOrder values differ (less(i, j) and less(j, i) can both hold), so the outcome
        // is whatever Go's sort.SliceStable leaves: goStableSort replays it (std::stable_sort would be undefined behaviour)
        goStableSort((int)solids.size(),
                     [&](int i, int j) { if (solids[i].order < solids[j].order) return true; return solids[i].depth > solids[j].depth; },
                     [&](int i, int j) { std::swap(solids[i], solids[j]); });
    stableSort(transparents, [](const renderPair& i, const renderPair& j) {   // camera.go:2050-2055
        if (i.order == j.order) return i.depth > j.depth;
        return i.order < j.order;
    });

    lap("pairs + sorts");
    // ---- walk the passes and collect every render(rp) call, grouped by flush (camera.go:2057-2109) ----
    // deferFrustum: the child's ModelInFrustum test (camera.go:2088-2093) is left to the device (t3d_submit_instances with bounds)
    struct PendingDraw { Model* model; int part; int flush; bool first; bool deferFrustum; bool uniformRun; };
    std::vector<PendingDraw> pending;
    std::vector<std::pair<Model*, int>> flushOwners;
    for (auto* pass : {&solids, &transparents}) {
        for (renderPair& pair : *pass) {
            if (!pair.model->visible) continue;
            const int flushId = (int)flushOwners.size();
            bool first = true;
            if (pair.model->DynamicBatcher()) {
                auto& modelSlice = pair.model->DynamicBatchModels[pair.part];
                // sort.Slice far -> near (camera.go:2078-2080); ties keep insertion order (contract)
                // (the distance of every child is evaluated once, not once per comparison: same keys, same order)
                {
                    std::vector<std::pair<float, uint32_t>> keyed(modelSlice.size());
                    for (size_t k = 0; k < modelSlice.size(); k++) keyed[k] = {v3_dist2(cameraPos, modelSlice[k]->WorldPosition()), (uint32_t)k};
                    bool radix = keyed.size() >= 4096;
                    for (size_t k = 0; radix && k < keyed.size(); k++) radix = keyed[k].first >= 0.0f;      // (also false for NaN)
                    if (radix) {
                        // large batches: the same order from a stable LSD radix sort -- squared distances are non-negative floats, whose
                        // bit patterns order like the values; descending = ascending on the complement; equal keys keep their order
                        std::vector<std::pair<float, uint32_t>> tmp(keyed.size());
                        auto key = [](const std::pair<float, uint32_t>& e) { uint32_t u; std::memcpy(&u, &e.first, 4); return ~u; };
                        for (int shift = 0; shift < 32; shift += 11) {
                            uint32_t count[2049] = {0};
                            for (auto& e : keyed) count[((key(e) >> shift) & 2047u) + 1]++;
                            for (int b = 0; b < 2048; b++) count[b + 1] += count[b];
                            for (auto& e : keyed) tmp[count[(key(e) >> shift) & 2047u]++] = e;
                            keyed.swap(tmp);
                        }
                    } else
                    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<float, uint32_t>& a, const std::pair<float, uint32_t>& b) { return a.first > b.first; });
                    std::vector<std::shared_ptr<Model>> sorted(modelSlice.size());
                    for (size_t k = 0; k < keyed.size(); k++) sorted[k] = std::move(modelSlice[keyed[k].second]);
                    modelSlice.swap(sorted);
                }
                // A batch whose children are instances of one another (one mesh of one part, one material, one parent, one
                // rotation: particles) is submitted as instances, and then the device also runs the children's frustum test.
                bool uniform = UseInstancing && modelSlice.size() >= 8;
                for (size_t k = 0; uniform && k < modelSlice.size(); k++) {
                    Model& a = *modelSlice[0]; Model& b = *modelSlice[k];
                    uniform = b.mesh && b.mesh == a.mesh && b.mesh->MeshParts.size() == 1 && !b.skinned && b.parent == a.parent &&
                              std::memcmp(&b.rotation, &a.rotation, sizeof(Matrix4)) == 0 && b.lightGroup == a.lightGroup && b.Shadeless == a.Shadeless;
                }
                for (auto& merged : modelSlice) {
                    if (!merged->visible) continue;
                    const bool defer = uniform && DeviceFrustumCulling && merged->FrustumCulling;
                    if (merged->FrustumCulling && !defer && !ModelInFrustum(*merged)) continue;
                    if (!merged->mesh) continue;
                    for (size_t cp = 0; cp < merged->mesh->MeshParts.size(); cp++) { pending.push_back({merged.get(), (int)cp, flushId, first, defer, uniform}); first = false; }
                }
            } else {
                if (!pair.model->mesh) continue;
                pending.push_back({pair.model, pair.part, flushId, true, false, false});
            }
            flushOwners.push_back({pair.model, pair.part});
        }
    }

    lap("pending (batch order)");
    // ---- object-wise sharding: whole flushes, largest first onto the least-loaded rank ----
    const int nFlush = (int)flushOwners.size();
    {
        std::vector<uint64_t> flushTris(nFlush, 0);
        for (auto& pd : pending) flushTris[pd.flush] += (uint64_t)pd.model->mesh->MeshParts[pd.part]->TriangleCount();
        flushOwnerRank = AssignFlushesToRanks(flushTris, shardCount);
        shardDrawFlush.assign((size_t)shardCount, {});
        if (shardCount > 1)
            for (auto& pd : pending) {      // the same skip rule as below (ProcessVertices early-out, model.go:604)
                MeshPart& mp = *pd.model->mesh->MeshParts[pd.part];
                const bool skip = (pd.model->Color.A == 0 && pd.model->isTransparent(mp)) || !pd.model->visible;
                if (!skip) shardDrawFlush[(size_t)flushOwnerRank[pd.flush]].push_back(pd.flush);
            }
    }

    // ---- frame ----
    t3d_camera cam; std::memset(&cam, 0, sizeof cam);
    m4_store(vpMatrix, cam.vp); m4_store(view, cam.view);
    cam.position[0] = cameraPos.X; cam.position[1] = cameraPos.Y; cam.position[2] = cameraPos.Z;
    cam.up[0] = cameraUp.X; cam.up[1] = cameraUp.Y; cam.up[2] = cameraUp.Z;
    cam.near_plane = near_; cam.far_plane = far_; cam.width = width; cam.height = height;
    cam.depth_margin = DepthMargin; cam.vertex_snapping = VertexSnapping;
    cam.flags = (perspective ? T3D_CAM_PERSPECTIVE : 0) | (RenderDepth ? T3D_CAM_RENDER_DEPTH : 0) |
                (RenderNormals ? T3D_CAM_RENDER_NORMALS : 0) | (PerspectiveCorrectedTextureMapping ? T3D_CAM_PERSP_CORRECT_UV : 0);
    // what SphereInFrustum reads (camera.go:933-983): the device runs it for instances submitted with culling bounds
    cam.forward[0] = cameraForward.X; cam.forward[1] = cameraForward.Y; cam.forward[2] = cameraForward.Z;
    cam.right[0] = cameraRight.X; cam.right[1] = cameraRight.Y; cam.right[2] = cameraRight.Z;
    cam.sphere_factor_y = sphereFactorY; cam.sphere_factor_x = sphereFactorX; cam.sphere_factor_tang = sphereFactorTang;
    cam.ortho_scale = orthoScale;
    t3d_world tw; std::memset(&tw, 0, sizeof tw);
    if (world) {
        tw.present = 1; tw.fog_on = world->FogOn; tw.fog_mode = world->FogMode;
        tw.fog_color[0] = world->FogColor.R; tw.fog_color[1] = world->FogColor.G; tw.fog_color[2] = world->FogColor.B;
        tw.fog_range[0] = world->FogRange[0]; tw.fog_range[1] = world->FogRange[1];
    }
    device->frameSerial++;
    check(ctx, t3d_set_async_download(ctx, AsyncDownload && downloadResults), "t3d_set_async_download");
    check(ctx, t3d_set_skip_unused_depth_list(ctx, SkipUnusedDepthList), "t3d_set_skip_unused_depth_list");
    check(ctx, t3d_frame_begin(ctx, &cam, &tw), "t3d_frame_begin");
    // anything that throws before the frame is executed (an unsupported draw, a failed upload) must not leave the frame open
    struct FrameGuard { t3d_ctx* c; bool armed; ~FrameGuard() { if (armed) t3d_frame_abort(c); } } frameGuard{ctx, true};
    if (!scene.autosubdivisionLevels.empty())      // camera.go:1547-1553
        check(ctx, t3d_frame_set_autosubdivision(ctx, scene.autosubdivisionLevels.data(), (uint32_t)(scene.autosubdivisionLevels.size() / 2)), "t3d_frame_set_autosubdivision");

    // light lists: the scene list once, each active LightGroup once (camera.go:1576-1581)
    uint32_t sceneList = 0; bool sceneListMade = false;
    std::unordered_map<LightGroup*, uint32_t> groupLists;
    auto makeList = [&](const std::vector<ILight*>& ls) {
        std::vector<t3d_light> recs; recs.reserve(ls.size());
        for (ILight* l : ls) recs.push_back(toLight(*device, l));
        uint32_t id = 0;
        check(ctx, t3d_submit_lights(ctx, recs.data(), (uint32_t)recs.size(), &id), "t3d_submit_lights");
        return id;
    };
    std::unordered_map<Model*, uint32_t> boneOffsets;
    {   // armatures posed on the device: one call per (skeleton, chain of animations), in first-use order
        struct Group { Skeleton* sk; std::vector<uint64_t> anims; std::vector<t3d_pose_chain> poses; std::vector<Model*> models; };
        std::vector<Group> groups;
        std::map<std::pair<uint64_t, std::vector<uint64_t>>, size_t> groupOf;
        for (auto& pd : pending) {
            if (flushOwnerRank[pd.flush] != shardRank) continue;
            Model* model = pd.model;
            if (!model->skinned || !model->skeleton || boneOffsets.count(model)) continue;
            Skeleton& sk = *model->skeleton;
            device->ensureSkeleton(sk);
            std::vector<uint64_t> anims;           // empty: rest pose
            t3d_pose_chain P; std::memset(&P, 0, sizeof P);
            if (AnimationPlayer* ap = model->posePlayer.get()) {
                if (ap->animation && ap->sampled) {
                    if (ap->animation->relativeMotion)
                        throw T3DError(T3D_E_UNSUPPORTED, "AnimationPlayer relative motion (animation.go:962-984) is not evaluated on the device");
                    if (ap->sampledChain.size() > T3D_POSE_CHAIN_MAX)
                        throw T3DError(T3D_E_UNSUPPORTED, "AnimationPlayer: a blend interrupted by more than T3D_POSE_CHAIN_MAX - 1 further Play() calls (animation.go:905-908) is not evaluated on the device");
                    for (size_t k = 0; k < ap->sampledChain.size(); k++) {
                        const BlendLevel& L = ap->sampledChain[k];
                        anims.push_back(device->ensureAnimation(*L.anim, sk));
                        P.playhead[k] = L.playhead; P.blend[k] = L.bp;
                    }
                }
            }
            m4_store(model->SkinRoot ? model->SkinRoot->Transform() : m4_identity(), P.root);
            auto key = std::make_pair(sk.id, anims);
            auto it = groupOf.find(key);
            if (it == groupOf.end()) { it = groupOf.emplace(key, groups.size()).first; groups.push_back(Group{&sk, anims, {}, {}}); }
            Group& g = groups[it->second];
            g.poses.push_back(P); g.models.push_back(model);
            boneOffsets.emplace(model, 0u);        // filled in below
        }
        for (Group& g : groups) {
            uint32_t off = 0;
            const uint32_t n = (uint32_t)g.poses.size();
            if (g.anims.size() <= 2) {          // plain poses and blends between two animations: the 80-byte record
                std::vector<t3d_pose> poses(n);
                for (uint32_t i = 0; i < n; i++) {
                    std::memset(&poses[i], 0, sizeof(t3d_pose));
                    std::memcpy(poses[i].root, g.poses[i].root, 64);
                    if (g.anims.size() == 2) { poses[i].from_playhead = g.poses[i].playhead[0]; poses[i].playhead = g.poses[i].playhead[1]; poses[i].blend = g.poses[i].blend[1]; }
                    else poses[i].playhead = g.poses[i].playhead[0];
                }
                if (g.anims.size() == 2) check(ctx, t3d_submit_poses_blended(ctx, g.sk->id, g.anims[0], g.anims[1], poses.data(), n, &off), "t3d_submit_poses_blended");
                else check(ctx, t3d_submit_poses(ctx, g.sk->id, g.anims.empty() ? 0 : g.anims[0], poses.data(), n, &off), "t3d_submit_poses");
            } else {
                check(ctx, t3d_submit_pose_chains(ctx, g.sk->id, g.anims.data(), (uint32_t)g.anims.size(), g.poses.data(), n, &off), "t3d_submit_pose_chains");
            }
            for (size_t i = 0; i < g.models.size(); i++) boneOffsets[g.models[i]] = off + (uint32_t)(i * g.sk->BoneCount());
        }
    }

    renders.clear(); flushes.clear(); localFlushIds.clear();
    std::vector<t3d_draw> draws; draws.reserve(pending.size());
    // per draw: may it travel as an instance (t3d_submit_instances)?  Its Node-local placement, the Material colour (the
    // instance carries Model.Color), and what a run of instances has to share beyond the t3d_draw fields
    struct InstInfo { bool eligible, frustum; Node* parent; const Matrix4* rotation; float matColor[4]; t3d_instance inst; float bounds[9];
                      size_t extraFirst, extraCount; };
    std::vector<InstInfo> instInfo; instInfo.reserve(pending.size());
    // Children of a uniform DynamicBatch (uniformRun: instances of the batch's first child by construction -- one mesh of one
    // part, one material, parent, rotation, light group) do not get a t3d_draw of their own: only their t3d_instance is kept, behind
    // the draw of the child that opened the run (InstInfo.extraFirst / extraCount).  A 2 M-particle frame builds one 200-byte
    // record per batch instead of one per particle.
    std::vector<t3d_instance> extraInsts;
    int runHeadDraw = -1, runHeadFlush = -1; bool runHeadDefer = false;
    uint32_t drawCounter = 0;
    std::vector<const PendingDraw*> submitted;
    submitted.reserve(pending.size()); renders.reserve(pending.size());
    const Mesh* lastEnsured = nullptr;
    for (auto& pd : pending) {
        if (flushOwnerRank[pd.flush] != shardRank) continue;
        Model* model = pd.model;
        Mesh& mesh = *model->mesh;
        MeshPart& meshPart = *mesh.MeshParts[pd.part];
        Material* mat = meshPart.material.get();
        if (model->HasVertexTransformFunction || model->HasVertexClipFunction || (mat && mat->HasCustomDepthFunction))
            throw T3DError(T3D_E_UNSUPPORTED, "host callbacks (VertexTransformFunction / VertexClipFunction / CustomDepthFunction) cannot run on the device");
        if (model->DynamicBatchOwner) debugInfo.batchedParts++;            // camera.go:1534-1536
        // ProcessVertices early-out (model.go:604): nothing is transformed or emitted
        const bool skip = (model->Color.A == 0 && model->isTransparent(meshPart)) || !model->visible;
        if (pd.uniformRun && !skip && runHeadDraw >= 0 && runHeadFlush == pd.flush && runHeadDefer == pd.deferFrustum && !pd.first) {
            // one more instance of the run its batch's first child opened: everything but position, scale and colour is the head's
            InstInfo& head = instInfo[(size_t)runHeadDraw];
            t3d_instance in;
            in.position[0] = model->position.X; in.position[1] = model->position.Y; in.position[2] = model->position.Z;
            in.scale[0] = model->scale.X; in.scale[1] = model->scale.Y; in.scale[2] = model->scale.Z;
            in.color[0] = model->Color.R; in.color[1] = model->Color.G; in.color[2] = model->Color.B; in.color[3] = model->Color.A;
            extraInsts.push_back(in);
            head.extraCount++;
            debugInfo.totalTris += meshPart.TriangleCount();               // camera.go:1532
            submitted.push_back(&pd);
            renders.push_back({model, pd.part, pd.flush, 0, 0, 0, 0, 0, (int)drawCounter++});
            continue;
        }

        if (&mesh != lastEnsured) { device->ensureMesh(mesh); lastEnsured = &mesh; }     // runs of instances share a mesh
        bool lighting = false;                                             // camera.go:1521-1528
        if (world) lighting = world->LightingOn && !(mat && mat->Shadeless) && !model->Shadeless;

        t3d_draw d; std::memset(&d, 0, sizeof d);
        d.mesh_id = mesh.id; d.part = (uint32_t)pd.part;
        uint32_t f = 0;
        if (mat) {
            f |= T3D_DRAW_HAS_MATERIAL;
            if (mat->BackfaceCulling) f |= T3D_DRAW_BACKFACE_CULLING;
            if (mat->BillboardEnabled) f |= T3D_DRAW_BILLBOARD;
            if (mat->BillboardLockX) f |= T3D_DRAW_BILLBOARD_LOCK_X;
            if (mat->BillboardLockY) f |= T3D_DRAW_BILLBOARD_LOCK_Y;
            if (mat->BillboardLockZ) f |= T3D_DRAW_BILLBOARD_LOCK_Z;
            if (mat->BillboardUpDirection == BillboardYUpCameraY) f |= T3D_DRAW_BILLBOARD_CAMERA_UP;
            if (mat->BillboardedDepthMode == DepthModeUnbillboarded) f |= T3D_DRAW_DEPTH_UNBILLBOARDED;
            d.sort_mode = (uint8_t)mat->TriangleSortMode; d.lighting_mode = (uint8_t)mat->LightingMode;
            d.light_volume_mode = (uint8_t)mat->LightVolumeShadingMode;
        }
        if (lighting) f |= T3D_DRAW_LIT;
        if (model->skinned) f |= T3D_DRAW_SKINNED;
        if (!pd.first) f |= T3D_DRAW_CONTINUE_FLUSH;
        d.flags = f;
        m4_store(model->Transform(), d.transform);
        if (lighting) m4_store(model->WorldRotation(), d.world_rotation);    // only a directional light's beginModel reads it (light.go:326-330)
        Color4 mp = model->Color;                                          // camera.go:1597-1601
        if (mat) { mp.R *= mat->Color.R; mp.G *= mat->Color.G; mp.B *= mat->Color.B; mp.A *= mat->Color.A; }
        d.color[0] = mp.R; d.color[1] = mp.G; d.color[2] = mp.B; d.color[3] = mp.A;
        if (mat && mat->UseTexture && mat->TextureWidth > 0) { d.src_w = (float)mat->TextureWidth; d.src_h = (float)mat->TextureHeight; }   // camera.go:1560-1566
        if (lighting) {
            if (model->lightGroup && model->lightGroup->Active) {
                auto it = groupLists.find(model->lightGroup.get());
                if (it == groupLists.end()) {
                    std::vector<ILight*> ls; for (auto& l : model->lightGroup->Lights) ls.push_back(l.get());
                    it = groupLists.emplace(model->lightGroup.get(), makeList(ls)).first;
                }
                d.light_list = it->second;
            } else {
                if (!sceneListMade) { sceneList = makeList(sceneLights); sceneListMade = true; }
                d.light_list = sceneList;
            }
        }
        if (model->skinned && model->skeleton) {
            d.bone_offset = boneOffsets.at(model); d.bone_count = (uint32_t)model->skeleton->BoneCount();
        } else if (model->skinned) {
            auto it = boneOffsets.find(model);
            if (it == boneOffsets.end()) {
                uint32_t off = 0;
                check(ctx, t3d_submit_bones(ctx, model->boneInfluences.data(), (uint32_t)(model->boneInfluences.size() / 16), &off), "t3d_submit_bones");
                it = boneOffsets.emplace(model, off).first;
            }
            d.bone_offset = it->second; d.bone_count = (uint32_t)(model->boneInfluences.size() / 16);
        }
        debugInfo.totalTris += meshPart.TriangleCount();                   // camera.go:1532
        if (skip) {
            // keep the flush structure, submit nothing for this render(): record an empty render
            renders.push_back({model, pd.part, pd.flush, 0, 0, (uint32_t)meshPart.TriangleCount(), FLT_MAX, -FLT_MAX, -1});
            submitted.push_back(nullptr);
            continue;
        }
        {
            InstInfo ii; std::memset(&ii, 0, sizeof ii);
            ii.eligible = UseInstancing && !model->skinned;
            ii.frustum = pd.deferFrustum;
            if (ii.frustum) {       // mesh.Dimensions {Center, Min, Max}: what Model.onTransformUpdate builds the culling sphere from
                const Dimensions& dm = mesh.dimensions; const Vector3 c = dm.Center();
                const float b9[9] = {c.X, c.Y, c.Z, dm.Min.X, dm.Min.Y, dm.Min.Z, dm.Max.X, dm.Max.Y, dm.Max.Z};
                std::memcpy(ii.bounds, b9, sizeof b9);
            }
            ii.parent = model->parent; ii.rotation = &model->rotation;
            ii.matColor[0] = mat ? mat->Color.R : 1.0f; ii.matColor[1] = mat ? mat->Color.G : 1.0f;
            ii.matColor[2] = mat ? mat->Color.B : 1.0f; ii.matColor[3] = mat ? mat->Color.A : 1.0f;
            ii.inst.position[0] = model->position.X; ii.inst.position[1] = model->position.Y; ii.inst.position[2] = model->position.Z;
            ii.inst.scale[0] = model->scale.X; ii.inst.scale[1] = model->scale.Y; ii.inst.scale[2] = model->scale.Z;
            ii.inst.color[0] = model->Color.R; ii.inst.color[1] = model->Color.G; ii.inst.color[2] = model->Color.B; ii.inst.color[3] = model->Color.A;
            ii.extraFirst = extraInsts.size(); ii.extraCount = 0;
            instInfo.push_back(ii);
        }
        draws.push_back(d);
        // a deferred child opens (or, after an early-out sibling, re-opens) a run that the following deferred children of its flush join;
        // extras always hang off the LAST materialised draw, so the submission order stays the render order
        if (pd.uniformRun) { runHeadDraw = (int)draws.size() - 1; runHeadFlush = pd.flush; runHeadDefer = pd.deferFrustum; } else runHeadDraw = -1;
        submitted.push_back(&pd);
        renders.push_back({model, pd.part, pd.flush, 0, 0, 0, 0, 0, (int)drawCounter++});
    }
    lastBoneOffsets.clear();
    for (auto& kv : boneOffsets) lastBoneOffsets[kv.first->name] = kv.second;
    lap("build t3d_draw records");
    {
        // Runs of Models that share everything but position, scale and colour (DynamicBatch children: particles; instanced
        // scenery) go down as instances -- the device then evaluates their Node.Transform (node.go:577-583) -- everything else
        // as draws, in render order.
        auto sameRun = [&](size_t a, size_t b) {        // may draw b continue the instance run that draw a is in?
            const t3d_draw& x = draws[a]; const t3d_draw& y = draws[b];
            const InstInfo& p = instInfo[a]; const InstInfo& q = instInfo[b];
            return q.eligible && (y.flags & T3D_DRAW_CONTINUE_FLUSH) && x.mesh_id == y.mesh_id && x.part == y.part &&
                   (x.flags | T3D_DRAW_CONTINUE_FLUSH) == y.flags && std::memcmp(x.world_rotation, y.world_rotation, 64) == 0 &&
                   x.src_w == y.src_w && x.src_h == y.src_h && x.sort_mode == y.sort_mode && x.lighting_mode == y.lighting_mode &&
                   x.light_volume_mode == y.light_volume_mode && x.light_list == y.light_list && p.parent == q.parent && p.frustum == q.frustum &&
                   std::memcmp(p.rotation, q.rotation, sizeof(Matrix4)) == 0 && std::memcmp(p.matColor, q.matColor, 16) == 0;
        };
        const size_t MIN_RUN = 8;
        std::vector<t3d_instance> insts;
        size_t plainFrom = 0;
        for (size_t i = 0; i < draws.size();) {
            size_t j = i + 1;
            if (instInfo[i].eligible) while (j < draws.size() && sameRun(j - 1, j)) j++;
            bool extras = false;
            for (size_t k = i; k < j; k++) extras = extras || instInfo[k].extraCount != 0;
            // (a child whose frustum test was deferred, and one that carries further children of its batch, must go down as instances, however short the run)
            if (j - i >= MIN_RUN || instInfo[i].frustum || extras) {
                if (i > plainFrom) check(ctx, t3d_submit_draws(ctx, draws.data() + plainFrom, (uint32_t)(i - plainFrom)), "t3d_submit_draws");
                t3d_draw proto = draws[i];
                std::memcpy(proto.color, instInfo[i].matColor, 16);
                insts.clear();
                for (size_t k = i; k < j; k++) {
                    insts.push_back(instInfo[k].inst);
                    insts.insert(insts.end(), extraInsts.begin() + (ptrdiff_t)instInfo[k].extraFirst, extraInsts.begin() + (ptrdiff_t)(instInfo[k].extraFirst + instInfo[k].extraCount));
                }
                Matrix4 par = m4_identity();
                if (instInfo[i].parent) par = instInfo[i].parent->Transform();
                float rot16[16], par16[16];
                m4_store(*instInfo[i].rotation, rot16); m4_store(par, par16);
                check(ctx, t3d_submit_instances(ctx, &proto, rot16, instInfo[i].parent ? par16 : nullptr, instInfo[i].frustum ? instInfo[i].bounds : nullptr,
                                                insts.data(), (uint32_t)insts.size()), "t3d_submit_instances");
                plainFrom = j;
            }
            i = j;
        }
        if (draws.size() > plainFrom) check(ctx, t3d_submit_draws(ctx, draws.data() + plainFrom, (uint32_t)(draws.size() - plainFrom)), "t3d_submit_draws");
    }
    lap("t3d_submit_draws");
    check(ctx, t3d_frame_execute(ctx), "t3d_frame_execute");
    frameGuard.armed = false;
    lap("t3d_frame_execute");
    prevOutput = output;
    if (downloadResults) check(ctx, t3d_frame_results(ctx, &output, &timings), "t3d_frame_results");
    else check(ctx, t3d_frame_results_device(ctx, &output, &timings), "t3d_frame_results_device");
    launchCount = t3d_debug_last_launch_count(ctx);
    lap("t3d_frame_results");

    // per-render and per-flush records (what flush() would slice: camera.go:1829-2048)
    size_t di = 0;
    uint64_t lastEnd = 0;
    std::vector<char> culledOnDevice(renders.size(), 0);
    for (size_t i = 0; i < renders.size(); i++) {
        RenderRecord& r = renders[i];
        if (submitted[i]) {
            const t3d_draw_result& dr = output.draws[di++];
            // an instance the device found outside the frustum: the reference never calls render() for it (camera.go:2088-2093)
            if (submitted[i]->deferFrustum && dr.total_tris == 0 && dr.vertex_count == 0) culledOnDevice[i] = 1;
            r.firstVertex = dr.first_vertex; r.vertexCount = dr.vertex_count; r.totalTris = dr.total_tris;
            r.minDepth = dr.min_depth; r.maxDepth = dr.max_depth;
            lastEnd = dr.first_vertex + dr.vertex_count;
        } else {
            r.firstVertex = lastEnd;
        }
    }
    {
        size_t ri = 0;
        for (int fId = 0; fId < nFlush; fId++) {
            if (flushOwnerRank[fId] != shardRank) continue;
            FlushRecord fr{flushOwners[fId].first, flushOwners[fId].second, 0, 0};
            bool any = false;
            while (ri < renders.size() && renders[ri].flush == fId) {
                if (!any) { fr.firstVertex = renders[ri].firstVertex; any = true; }
                fr.vertexCount += renders[ri].vertexCount;
                ri++;
            }
            if (!any) fr.firstVertex = lastEnd;
            if (fr.vertexCount) { debugInfo.drawnTris += (int)(fr.vertexCount / 3); debugInfo.drawnParts++; }   // camera.go:2041-2044
            flushes.push_back(fr);
            localFlushIds.push_back(fId);
        }
    }
    {   // instances the device culled were never render()ed in the reference: no record, no DebugInfo counts (camera.go:1532-1536)
        size_t w = 0;
        for (size_t i = 0; i < renders.size(); i++) {
            if (culledOnDevice[i]) {
                Model* m = submitted[i]->model;
                debugInfo.totalTris -= m->mesh->MeshParts[submitted[i]->part]->TriangleCount();
                if (m->DynamicBatchOwner) debugInfo.batchedParts--;
                continue;
            }
            if (w != i) renders[w] = renders[i];
            w++;
        }
        renders.resize(w);
    }
    lap("render/flush records");
}

// Triangle.Normal and Triangle.MaxSpan of every triangle (mesh.go:2018-2037, 2171-2179), load-time products of the reference.
void TriangleNormalsAndSpans(const Mesh& mesh, std::vector<float>& normals3, std::vector<float>& spans) {
    const int nt = mesh.TriangleTotal();
    normals3.resize((size_t)3 * nt); spans.resize((size_t)nt);
    auto P = [&](uint32_t v) { return v3(mesh.VertexPositions[3 * v], mesh.VertexPositions[3 * v + 1], mesh.VertexPositions[3 * v + 2]); };
    for (int t = 0; t < nt; t++) {
        const uint32_t a = mesh.TriangleIndices[3 * t], b = mesh.TriangleIndices[3 * t + 1], c = mesh.TriangleIndices[3 * t + 2];
        Vector3 n = v3_unit(v3_cross(v3_sub(P(b), P(a)), v3_sub(P(c), P(b))));        // calculateNormal
        const float eps = 1e-4f;
        if (!(std::fabs(n.X) > eps || std::fabs(n.Y) > eps || std::fabs(n.Z) > eps))  // IsZero: fall back to the first vertex normal
            n = v3(mesh.VertexNormals[3 * a], mesh.VertexNormals[3 * a + 1], mesh.VertexNormals[3 * a + 2]);
        normals3[3 * t] = n.X; normals3[3 * t + 1] = n.Y; normals3[3 * t + 2] = n.Z;
        spans[t] = std::max(std::max(v3_magnitude(v3_sub(P(a), P(b))), v3_magnitude(v3_sub(P(b), P(c)))), v3_magnitude(v3_sub(P(c), P(a))));
    }
}

std::vector<float> Model::BakeAO(GeometryDevice& dev, int targetChannel, float occlusionAngle, const Color4& oc,
                                 const std::vector<int>& targetParts, const std::vector<int>& targetVertices) {
    std::vector<float> out;
    if (!mesh || targetChannel < 0) return out;                            // model.go:1042-1044
    const int n = mesh->VertexCount();
    if (mesh->VertexActiveColorChannel == targetChannel && (int)mesh->VertexColors.size() == 4 * n) out = mesh->VertexColors;
    else out.assign((size_t)4 * n, 1.0f);                                  // ensureEnoughVertexColorChannels
    std::vector<float> normals, spans;
    TriangleNormalsAndSpans(*mesh, normals, spans);
    std::vector<uint8_t> parts, verts;
    if (!targetParts.empty()) { parts.assign(mesh->MeshParts.size(), 0); for (int p : targetParts) if (p >= 0 && p < (int)parts.size()) parts[p] = 1; }
    if (!targetVertices.empty()) { verts.assign((size_t)n, 0); for (int v : targetVertices) if (v >= 0 && v < n) verts[v] = 1; }
    dev.ensureMesh(*mesh);
    const float c4[4] = {oc.R, oc.G, oc.B, oc.A};
    check(dev.ctx, t3d_bake_ao(dev.ctx, mesh->id, normals.data(), spans.data(), parts.empty() ? nullptr : parts.data(),
                               verts.empty() ? nullptr : verts.data(), occlusionAngle, c4, out.data()), "t3d_bake_ao");
    if (mesh->VertexActiveColorChannel == targetChannel) { mesh->VertexColors = out; mesh->MarkDirty(); }
    return out;
}

// ---- Model.BakeLighting (model.go:1203-1268) -------------------------------------------------------------------------
std::vector<float> Model::BakeLighting(int targetChannel, Camera& camera, Scene& scene, const std::vector<std::shared_ptr<ILight>>& lights) {
    std::vector<float> out;
    if (!mesh || targetChannel < 0) return out;                            // model.go:1205-1207
    if (!camera.device) throw T3DError(T3D_E_STATE, "Camera has no GeometryDevice: the geometry stage has no CPU path");
    GeometryDevice& dev = *camera.device;
    t3d_ctx* ctx = dev.ctx;
    const int n = mesh->VertexCount();
    // ensureEnoughVertexColorChannels (mesh.go:543-558): a channel that did not exist starts at 1, 1, 1, 1
    if (mesh->VertexActiveColorChannel == targetChannel && (int)mesh->VertexColors.size() == 4 * n) out = mesh->VertexColors;
    else out.assign((size_t)4 * n, 1.0f);

    // lightsInput = lights + World.AmbientLight (model.go:1212-1218); the invisible ones take no part (model.go:1227, 1253)
    std::vector<ILight*> ls;
    for (auto& l : lights) if (l && l->IsVisible()) ls.push_back(l.get());
    if (scene.world && scene.world->ambientLight->IsVisible()) ls.push_back(scene.world->ambientLight.get());

    camera.Clear();     // the camera axes ProcessVertices reads (camera.go:1085-1088)
    t3d_camera cam; std::memset(&cam, 0, sizeof cam);
    const Matrix4 view = camera.ViewMatrix();
    m4_store(m4_mult(view, camera.Projection()), cam.vp); m4_store(view, cam.view);
    const Vector3 cp = camera.WorldPosition();
    cam.position[0] = cp.X; cam.position[1] = cp.Y; cam.position[2] = cp.Z;
    cam.up[0] = camera.cameraUp.X; cam.up[1] = camera.cameraUp.Y; cam.up[2] = camera.cameraUp.Z;
    cam.near_plane = camera.near_; cam.far_plane = camera.far_; cam.width = camera.width; cam.height = camera.height;
    cam.depth_margin = camera.DepthMargin; cam.vertex_snapping = camera.VertexSnapping;
    cam.flags = (camera.perspective ? T3D_CAM_PERSPECTIVE : 0) | (camera.RenderDepth ? T3D_CAM_RENDER_DEPTH : 0);
    t3d_world tw; std::memset(&tw, 0, sizeof tw);
    tw.present = scene.world ? 1 : 0;
    dev.frameSerial++;
    check(ctx, t3d_frame_begin(ctx, &cam, &tw), "t3d_frame_begin");
    struct FrameGuard { t3d_ctx* c; ~FrameGuard() { t3d_frame_abort(c); } } frameGuard{ctx};       // (a no-op once the frame has been executed)
    check(ctx, t3d_frame_set_bake_lighting(ctx, 1), "t3d_frame_set_bake_lighting");
    std::vector<t3d_light> recs;
    for (ILight* l : ls) recs.push_back(toLight(dev, l));
    uint32_t list = 0;
    check(ctx, t3d_submit_lights(ctx, recs.data(), (uint32_t)recs.size(), &list), "t3d_submit_lights");
    dev.ensureMesh(*mesh);
    uint32_t boneOff = 0;
    if (skinned) check(ctx, t3d_submit_bones(ctx, boneInfluences.data(), (uint32_t)(boneInfluences.size() / 16), &boneOff), "t3d_submit_bones");
    std::vector<t3d_draw> draws; std::vector<int> drawPart;
    for (size_t pi = 0; pi < mesh->MeshParts.size(); pi++) {
        MeshPart& mp = *mesh->MeshParts[pi];
        Material* mat = mp.material.get();
        if (mat && mat->Shadeless) {                                       // model.go:1241-1248
            for (int v = mp.VertexIndexStart; v < mp.VertexIndexEnd; v++) { out[4 * v] = 1; out[4 * v + 1] = 1; out[4 * v + 2] = 1; }
            continue;
        }
        t3d_draw d; std::memset(&d, 0, sizeof d);
        d.mesh_id = mesh->id; d.part = (uint32_t)pi;
        uint32_t f = T3D_DRAW_LIT;
        if (mat) { f |= T3D_DRAW_HAS_MATERIAL; d.lighting_mode = (uint8_t)mat->LightingMode; d.light_volume_mode = (uint8_t)mat->LightVolumeShadingMode; d.sort_mode = (uint8_t)mat->TriangleSortMode; }
        if (skinned) { f |= T3D_DRAW_SKINNED; d.bone_offset = boneOff; d.bone_count = (uint32_t)(boneInfluences.size() / 16); }
        d.flags = f;
        m4_store(Transform(), d.transform);
        m4_store(WorldRotation(), d.world_rotation);
        d.color[0] = d.color[1] = d.color[2] = d.color[3] = 1.0f;
        d.light_list = list;
        draws.push_back(d); drawPart.push_back((int)pi);
    }
    if (!draws.empty()) check(ctx, t3d_submit_draws(ctx, draws.data(), (uint32_t)draws.size()), "t3d_submit_draws");
    check(ctx, t3d_frame_execute(ctx), "t3d_frame_execute");
    std::vector<float> l4;
    for (size_t i = 0; i < draws.size(); i++) {
        MeshPart& mp = *mesh->MeshParts[drawPart[i]];
        const int v0 = mp.VertexIndexStart, v1 = mp.VertexIndexEnd;       // ForEachVertexIndex (mesh.go:2261-2267)
        if (v1 <= v0) continue;
        l4.resize((size_t)4 * (v1 - v0));
        check(ctx, t3d_frame_read_vertex_lights(ctx, (uint32_t)i, (uint32_t)v0, (uint32_t)v1, l4.data()), "t3d_frame_read_vertex_lights");
        for (int v = v0; v < v1; v++) { out[4 * v] = l4[4 * (v - v0)]; out[4 * v + 1] = l4[4 * (v - v0) + 1]; out[4 * v + 2] = l4[4 * (v - v0) + 2]; }
    }
    if (mesh->VertexActiveColorChannel == targetChannel) { mesh->VertexColors = out; mesh->MarkDirty(); }
    return out;
}

}  // namespace tetra3d
