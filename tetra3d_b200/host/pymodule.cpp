// pymodule.cpp -- pybind11 binding of the host mirror (tetra3d.hpp) as tetra3d_b200._host.
// Names follow the Go API (SetLocalPosition, AddChildren, RenderScene, ...) so tests and benches read
// like the reference's examples.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#include <pybind11/functional.h>

#include <cstring>
#include "tetra3d.hpp"

namespace py = pybind11;
using namespace tetra3d;

using farr = py::array_t<float, py::array::c_style | py::array::forcecast>;

static Matrix4 mat_from(const farr& a) {
    if (a.size() != 16) throw std::invalid_argument("Matrix4 needs 16 floats");
    Matrix4 m; std::memcpy(&m, a.data(), 64); return m;
}
static farr mat_to(const Matrix4& m) {
    farr a({4, 4}); std::memcpy(a.mutable_data(), &m, 64); return a;
}
static farr vec_to(const Vector3& v) { farr a(3); a.mutable_data()[0] = v.X; a.mutable_data()[1] = v.Y; a.mutable_data()[2] = v.Z; return a; }
static Color4 col_from(const std::vector<float>& c) {
    if (c.size() != 4) throw std::invalid_argument("Color4 needs 4 floats");
    return Color4{c[0], c[1], c[2], c[3]};
}
static std::vector<float> col_to(const Color4& c) { return {c.R, c.G, c.B, c.A}; }
template <class T> static std::vector<T> vec_from(const py::array_t<T, py::array::c_style | py::array::forcecast>& a) {
    return std::vector<T>(a.data(), a.data() + a.size());
}

// numpy view over library-owned vertex lists (copied: the memory dies at the next Render)
static py::array vertex_list(const t3d_vertex* p, uint64_t n, bool onDevice) {
    if (onDevice) throw std::runtime_error("lists are device-resident (downloadResults=False); use DevicePointers()");
    if (!p) n = 0;          // a list the frame did not produce (normals off; depth skipped with RenderDepth off)
    farr a({(py::ssize_t)n, (py::ssize_t)12});
    if (n && p) std::memcpy(a.mutable_data(), p, n * sizeof(t3d_vertex));
    return a;
}

namespace tetra3d { void TriangleNormalsAndSpans(const Mesh& mesh, std::vector<float>& normals3, std::vector<float>& spans); }
PYBIND11_MODULE(_host, m) {
    m.doc() = "Host-side mirror of the Tetra3D API driving libt3dgeo (B200 geometry stage)";

    py::register_exception<T3DError>(m, "T3DError");

    m.def("NewMatrix4", [] { return mat_to(NewMatrix4()); });
    m.def("NewMatrix4Translate", [](float x, float y, float z) { return mat_to(NewMatrix4Translate(x, y, z)); });
    m.def("NewMatrix4Scale", [](float x, float y, float z) { return mat_to(NewMatrix4Scale(x, y, z)); });
    m.def("NewMatrix4Rotate", [](float x, float y, float z, float a) { return mat_to(NewMatrix4Rotate(x, y, z, a)); });
    m.def("Matrix4Mult", [](const farr& a, const farr& b) { return mat_to(t3d::m4_mult(mat_from(a), mat_from(b))); });
    m.def("Matrix4Inverted", [](const farr& a) { return mat_to(t3d::m4_inverted(mat_from(a))); });
    m.def("Matrix4Equals", [](const farr& a, const farr& b) { return Matrix4Equals(mat_from(a), mat_from(b)); });

    for (auto kv : std::initializer_list<std::pair<const char*, int>>{
             {"TriangleSortModeBackToFront", TriangleSortModeBackToFront}, {"TriangleSortModeFrontToBack", TriangleSortModeFrontToBack},
             {"TriangleSortModeNone", TriangleSortModeNone}, {"TransparencyModeAuto", TransparencyModeAuto},
             {"TransparencyModeOpaque", TransparencyModeOpaque}, {"TransparencyModeAlphaClip", TransparencyModeAlphaClip},
             {"TransparencyModeTransparent", TransparencyModeTransparent}, {"BillboardYUpGlobalY", BillboardYUpGlobalY},
             {"BillboardYUpCameraY", BillboardYUpCameraY}, {"LightingModeDefault", LightingModeDefault},
             {"LightingModeFixedNormals", LightingModeFixedNormals}, {"LightingModeDoubleSided", LightingModeDoubleSided},
             {"DepthModeDefault", DepthModeDefault}, {"DepthModeUnbillboarded", DepthModeUnbillboarded},
             {"LightVolumeShadingModePerVertexWithoutNormal", LightVolumeShadingModePerVertexWithoutNormal},
             {"LightVolumeShadingModePerVertexWithNormal", LightVolumeShadingModePerVertexWithNormal},
             {"LightVolumeShadingModePerObject", LightVolumeShadingModePerObject}, {"LightVolumeShadingModeOff", LightVolumeShadingModeOff},
             {"FogAdd", FogAdd}, {"FogSub", FogSub}, {"FogOverwrite", FogOverwrite}})
        m.attr(kv.first) = kv.second;

    py::class_<Material, std::shared_ptr<Material>>(m, "Material")
        .def(py::init([](const std::string& n) { return NewMaterial(n); }), py::arg("name") = "material")
        .def_property("Color", [](Material& s) { return col_to(s.Color); }, [](Material& s, const std::vector<float>& c) { s.Color = col_from(c); })
        .def_readwrite("BackfaceCulling", &Material::BackfaceCulling)
        .def_readwrite("TriangleSortMode", &Material::TriangleSortMode)
        .def_readwrite("Shadeless", &Material::Shadeless)
        .def_readwrite("LightVolumeShadingMode", &Material::LightVolumeShadingMode)
        .def_readwrite("BillboardEnabled", &Material::BillboardEnabled)
        .def_readwrite("BillboardLockX", &Material::BillboardLockX)
        .def_readwrite("BillboardLockY", &Material::BillboardLockY)
        .def_readwrite("BillboardLockZ", &Material::BillboardLockZ)
        .def_readwrite("BillboardUpDirection", &Material::BillboardUpDirection)
        .def_readwrite("Visible", &Material::Visible)
        .def_readwrite("TransparencyMode", &Material::TransparencyMode)
        .def_readwrite("BillboardedDepthMode", &Material::BillboardedDepthMode)
        .def_readwrite("LightingMode", &Material::LightingMode)
        .def_readwrite("RenderOrder", &Material::RenderOrder)
        .def_readwrite("BlendSourceOver", &Material::BlendSourceOver)
        .def_readwrite("UseTexture", &Material::UseTexture)
        .def_readwrite("TextureWidth", &Material::TextureWidth)
        .def_readwrite("TextureHeight", &Material::TextureHeight)
        .def_readwrite("HasCustomDepthFunction", &Material::HasCustomDepthFunction);

    py::class_<MeshPart, std::shared_ptr<MeshPart>>(m, "MeshPart")
        .def_readwrite("Material", &MeshPart::material)
        .def_readonly("VertexIndexStart", &MeshPart::VertexIndexStart)
        .def_readonly("VertexIndexEnd", &MeshPart::VertexIndexEnd)
        .def_readonly("TriangleStart", &MeshPart::TriangleStart)
        .def_readonly("TriangleEnd", &MeshPart::TriangleEnd)
        .def("TriangleCount", &MeshPart::TriangleCount);

    py::class_<Mesh, std::shared_ptr<Mesh>>(m, "Mesh")
        .def(py::init([](const std::string& n) { return NewMesh(n); }), py::arg("name") = "mesh")
        .def_readonly("ID", &Mesh::id)
        .def("TriangleNormalsAndSpans", [](Mesh& s) {       // Triangle.Normal / Triangle.MaxSpan (mesh.go:2018-2037): load-time products
            std::vector<float> n, sp; TriangleNormalsAndSpans(s, n, sp);
            py::array_t<float> a({(py::ssize_t)sp.size(), (py::ssize_t)3}), b((py::ssize_t)sp.size());
            std::memcpy(a.mutable_data(), n.data(), n.size() * sizeof(float)); std::memcpy(b.mutable_data(), sp.data(), sp.size() * sizeof(float));
            return py::make_tuple(a, b);
        })
        .def_readwrite("Name", &Mesh::Name)
        .def_readwrite("VertexActiveColorChannel", &Mesh::VertexActiveColorChannel)
        .def_readonly("MeshParts", &Mesh::MeshParts)
        // Mesh.AddVertices (mesh.go:678) in bulk: SoA arrays
        .def("SetVertices", [](Mesh& s, const farr& pos, const farr& nrm, const farr& uv) {
            if (pos.size() % 3 || nrm.size() != pos.size() || uv.size() * 3 != pos.size() * 2) throw std::invalid_argument("SetVertices: array sizes disagree");
            s.VertexPositions = vec_from<float>(pos); s.VertexNormals = vec_from<float>(nrm); s.VertexUVs = vec_from<float>(uv);
            s.MarkDirty();
        })
        .def("SetVertexColors", [](Mesh& s, const farr& col, int channel) {
            if ((size_t)col.size() != s.VertexPositions.size() / 3 * 4) throw std::invalid_argument("SetVertexColors: need 4 floats per vertex");
            s.VertexColors = vec_from<float>(col); s.VertexActiveColorChannel = channel; s.MarkDirty();
        }, py::arg("colors"), py::arg("channel") = 0)
        .def("SetSkin", [](Mesh& s, const farr& w, const py::array_t<uint16_t, py::array::c_style | py::array::forcecast>& b,
                           const py::array_t<uint8_t, py::array::c_style | py::array::forcecast>& n) {
            size_t nv = s.VertexPositions.size() / 3;
            if ((size_t)w.size() != nv * 4 || (size_t)b.size() != nv * 4 || (size_t)n.size() != nv) throw std::invalid_argument("SetSkin: need 4 weights, 4 bones, 1 count per vertex");
            s.VertexWeights = vec_from<float>(w); s.VertexBones = vec_from<uint16_t>(b); s.VertexInfluences = vec_from<uint8_t>(n); s.MarkDirty();
        })
        .def("SetShapeKeys", [](Mesh& s, const farr& kp, const farr& kn, const farr& kw) {
            size_t nv = s.VertexPositions.size() / 3;
            if ((size_t)kp.size() != (size_t)kw.size() * nv * 3 || kn.size() != kp.size()) throw std::invalid_argument("SetShapeKeys: need k*n*3 positions and normals");
            s.ShapeKeyPositions = vec_from<float>(kp); s.ShapeKeyNormals = vec_from<float>(kn); s.ShapeKeyWeights = vec_from<float>(kw); s.MarkDirty();
        })
        .def("SetShapeKeyWeights", [](Mesh& s, const farr& kw) { s.SetShapeKeyWeights(vec_from<float>(kw)); })
        // MeshPart.AddTriangles (mesh.go:2274) in bulk: absolute vertex indices
        .def("SetTriangles", [](Mesh& s, const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& idx) {
            if (idx.size() % 3) throw std::invalid_argument("SetTriangles: need 3 indices per triangle");
            s.TriangleIndices = vec_from<uint32_t>(idx);
            for (uint32_t i : s.TriangleIndices) if (i >= s.VertexPositions.size() / 3) throw std::invalid_argument("SetTriangles: index out of range");
            s.RecalculateCenters(); s.MarkDirty();
        })
        // the result of Mesh.SetAutoSubdivide(true) (mesh.go:447-485): the triangle arrays are already in the expanded order
        .def("SetAutoSubdivision", [](Mesh& s, const farr& centers, const py::array_t<uint32_t, py::array::c_style | py::array::forcecast>& base,
                                      const py::array_t<int8_t, py::array::c_style | py::array::forcecast>& level,
                                      const py::array_t<uint8_t, py::array::c_style | py::array::forcecast>& nLevels, const farr& span, const farr& normals) {
            const size_t t = (size_t)s.TriangleTotal();
            if ((size_t)centers.size() != 3 * t || (size_t)base.size() != t || (size_t)level.size() != t || (size_t)nLevels.size() != t || (size_t)span.size() != t || (size_t)normals.size() != 3 * t)
                throw std::invalid_argument("SetAutoSubdivision: arrays do not match the triangle count");
            s.TriangleCenters = vec_from<float>(centers); s.TriangleSubBase = vec_from<uint32_t>(base); s.TriangleSubLevel = vec_from<int8_t>(level);
            s.TriangleLevelCount = vec_from<uint8_t>(nLevels); s.TriangleMaxSpan = vec_from<float>(span); s.TriangleNormals = vec_from<float>(normals);
            s.autoSubdivide = true; s.MarkDirty();
        })
        .def("AutoSubdivide", [](Mesh& s) { return s.autoSubdivide; })
        .def("AddMeshPart", &Mesh::AddMeshPart, py::arg("material"), py::arg("vertex_start"), py::arg("vertex_end"), py::arg("triangle_start"), py::arg("triangle_end"))
        .def("UpdateDimensions", &Mesh::UpdateDimensions)
        .def("AutoNormal", [](Mesh& s) { s.AutoNormal(); s.MarkDirty(); })
        .def("MarkDirty", &Mesh::MarkDirty)
        .def("VertexCount", &Mesh::VertexCount)
        .def("TriangleCount", &Mesh::TriangleTotal)
        .def("Positions", [](Mesh& s) { farr a({(py::ssize_t)s.VertexCount(), (py::ssize_t)3}); std::memcpy(a.mutable_data(), s.VertexPositions.data(), s.VertexPositions.size() * 4); return a; })
        .def("Normals", [](Mesh& s) { farr a({(py::ssize_t)s.VertexCount(), (py::ssize_t)3}); std::memcpy(a.mutable_data(), s.VertexNormals.data(), s.VertexNormals.size() * 4); return a; })
        .def("UVs", [](Mesh& s) { farr a({(py::ssize_t)s.VertexCount(), (py::ssize_t)2}); std::memcpy(a.mutable_data(), s.VertexUVs.data(), s.VertexUVs.size() * 4); return a; })
        .def("Indices", [](Mesh& s) { py::array_t<uint32_t> a({(py::ssize_t)s.TriangleTotal(), (py::ssize_t)3}); std::memcpy(a.mutable_data(), s.TriangleIndices.data(), s.TriangleIndices.size() * 4); return a; })
        .def("Centers", [](Mesh& s) { farr a({(py::ssize_t)s.TriangleTotal(), (py::ssize_t)3}); std::memcpy(a.mutable_data(), s.TriangleCenters.data(), s.TriangleCenters.size() * 4); return a; })
        .def("MaxSpan", [](Mesh& s) { return s.dimensions.MaxSpan(); });
    m.def("NewMesh", &NewMesh);
    m.def("NewCubeMesh", &NewCubeMesh);
    m.def("NewMaterial", &NewMaterial);

    py::class_<Node, std::shared_ptr<Node>>(m, "Node")
        .def(py::init<const std::string&>(), py::arg("name") = "node")
        .def("Name", [](Node& n) { return n.name; })
        .def("SetLocalPosition", &Node::SetLocalPosition)
        .def("SetLocalScale", &Node::SetLocalScale)
        .def("SetLocalRotation", [](Node& n, const farr& r) { n.SetLocalRotation(mat_from(r)); })
        .def("LocalPosition", [](Node& n) { return vec_to(n.position); })
        .def("AddChildren", [](Node& n, py::args kids) { std::vector<std::shared_ptr<Node>> v; for (auto k : kids) v.push_back(k.cast<std::shared_ptr<Node>>()); n.AddChildren(v); })
        .def("Children", [](Node& n) { return n.children; })
        .def("Transform", [](Node& n) { return mat_to(n.Transform()); })
        .def("WorldPosition", [](Node& n) { return vec_to(n.WorldPosition()); })
        .def("WorldRotation", [](Node& n) { return mat_to(n.WorldRotation()); })
        .def("SetVisible", &Node::SetVisible, py::arg("visible"), py::arg("recursive") = false)
        .def("IsVisible", &Node::IsVisible);

    py::class_<ILight, Node, std::shared_ptr<ILight>>(m, "ILight")
        .def("Color", [](ILight& l) { return col_to(l.color); })
        .def("SetColor", [](ILight& l, const std::vector<float>& c) { l.color = col_from(c); })
        .def("Energy", [](ILight& l) { return l.energy; })
        .def("SetEnergy", [](ILight& l, float e) { l.energy = e; });
    py::class_<AmbientLight, ILight, std::shared_ptr<AmbientLight>>(m, "AmbientLight")
        .def(py::init<const std::string&, float, float, float, float>());
    py::class_<PointLight, ILight, std::shared_ptr<PointLight>>(m, "PointLight")
        .def(py::init<const std::string&, float, float, float, float>())
        .def("SetRange", &PointLight::SetRange)
        .def("Range", [](PointLight& p) { return p.lightRange; });
    py::class_<DirectionalLight, ILight, std::shared_ptr<DirectionalLight>>(m, "DirectionalLight")
        .def(py::init<const std::string&, float, float, float, float>());
    py::class_<LightVolume, ILight, std::shared_ptr<LightVolume>>(m, "LightVolume")
        .def(py::init<const std::string&>())
        .def("SetGrid", [](LightVolume& v, const farr& dmin, const farr& dmax, const farr& cellSize, const std::vector<int>& counts, const farr& colors) {
            if (dmin.size() != 3 || dmax.size() != 3 || cellSize.size() != 3 || counts.size() != 3) throw std::invalid_argument("SetGrid: 3-vectors expected");
            if ((size_t)colors.size() != (size_t)counts[0] * counts[1] * counts[2] * 3) throw std::invalid_argument("SetGrid: need 3 floats per cell");
            v.dimensions.Min = t3d::v3(dmin.data()[0], dmin.data()[1], dmin.data()[2]);
            v.dimensions.Max = t3d::v3(dmax.data()[0], dmax.data()[1], dmax.data()[2]);
            for (int i = 0; i < 3; i++) { v.cellSize[i] = cellSize.data()[i]; v.cellCount[i] = counts[i]; }
            v.cellColors = vec_from<float>(colors); v.generation++;
        })
        .def("LightVolumeSetOffset", [](LightVolume& v, float x, float y, float z) { v.offset = t3d::v3(x, y, z); v.generation++; })
        .def("LightVolumeSetLightBias", [](LightVolume& v, float x, float y, float z) { v.lightBias = t3d::v3(x, y, z); v.generation++; })
        // LightVolumeCell.SetUseLights / ClearUseLights of the cell at (x, y, z) (light.go:1508-1518)
        .def("CellSetUseLights", [](LightVolume& v, int x, int y, int z, py::args lights) {
            std::vector<std::shared_ptr<ILight>> ls; for (auto l : lights) ls.push_back(l.cast<std::shared_ptr<ILight>>());
            v.CellSetUseLights(x, y, z, ls);
        })
        .def("CellClearUseLights", &LightVolume::CellClearUseLights);
    m.def("NewAmbientLight", [](const std::string& n, float r, float g, float b, float e) { return std::make_shared<AmbientLight>(n, r, g, b, e); });
    m.def("NewPointLight", [](const std::string& n, float r, float g, float b, float e) { return std::make_shared<PointLight>(n, r, g, b, e); });
    m.def("NewDirectionalLight", [](const std::string& n, float r, float g, float b, float e) { return std::make_shared<DirectionalLight>(n, r, g, b, e); });

    py::class_<LightGroup, std::shared_ptr<LightGroup>>(m, "LightGroup")
        .def(py::init([](py::args lights) { auto g = std::make_shared<LightGroup>(); for (auto l : lights) g->Lights.push_back(l.cast<std::shared_ptr<ILight>>()); return g; }))
        .def("Add", [](LightGroup& g, py::args lights) { for (auto l : lights) g.Lights.push_back(l.cast<std::shared_ptr<ILight>>()); })
        .def_readwrite("Active", &LightGroup::Active);
    m.def("NewLightGroup", [](py::args lights) { auto g = std::make_shared<LightGroup>(); for (auto l : lights) g->Lights.push_back(l.cast<std::shared_ptr<ILight>>()); return g; });

    // ---- animation.go ----
    py::class_<AnimationTrack, std::shared_ptr<AnimationTrack>>(m, "AnimationTrack")
        .def_readwrite("Interpolation", &AnimationTrack::Interpolation)
        .def("AddKeyframe", [](AnimationTrack& t, float time, const std::vector<float>& v) {
            if (v.size() != 3 && v.size() != 4) throw std::invalid_argument("keyframe data: a Vector3 or a Quaternion");
            t.AddKeyframe(time, v[0], v[1], v[2], v.size() == 4 ? v[3] : 0.0f);
        })
        .def("KeyframeCount", [](AnimationTrack& t) { return t.Keyframes.size(); });
    py::class_<AnimationChannel, std::shared_ptr<AnimationChannel>>(m, "AnimationChannel")
        .def_readonly("Name", &AnimationChannel::Name)
        .def("AddTrack", &AnimationChannel::AddTrack);
    py::class_<Animation, std::shared_ptr<Animation>>(m, "Animation")
        .def(py::init<const std::string&>())
        .def("AddChannel", &Animation::AddChannel)
        .def("Name", [](Animation& a) { return a.name; })
        .def("Length", [](Animation& a) { return a.length; })
        .def("SetLength", [](Animation& a, float l) { a.length = l; })
        .def("LoopMode", [](Animation& a) { return a.loopMode; })
        .def("SetLoopMode", [](Animation& a, int m) { a.loopMode = m; })
        .def("SetRelativeMotion", [](Animation& a, bool r) { a.relativeMotion = r; })
        .def("MarkDirty", &Animation::MarkDirty);
    m.def("NewAnimation", [](const std::string& n) { return std::make_shared<Animation>(n); });
    m.attr("TrackTypePosition") = (int)TrackTypePosition; m.attr("TrackTypeScale") = (int)TrackTypeScale; m.attr("TrackTypeRotation") = (int)TrackTypeRotation;
    m.attr("InterpolationLinear") = (int)InterpolationLinear; m.attr("InterpolationConstant") = (int)InterpolationConstant; m.attr("InterpolationCubic") = (int)InterpolationCubic;
    m.attr("AnimationLoopModeLoop") = (int)AnimationLoopModeLoop; m.attr("AnimationLoopModePingPong") = (int)AnimationLoopModePingPong; m.attr("AnimationLoopModeOneshot") = (int)AnimationLoopModeOneshot;
    py::class_<Skeleton, std::shared_ptr<Skeleton>>(m, "Skeleton")
        .def("BoneCount", &Skeleton::BoneCount)
        .def_readonly("boneNames", &Skeleton::boneNames);
    m.def("NewSkeletonFromJoints", [](const std::vector<std::shared_ptr<Node>>& joints, const farr& inverseBind) {
        return Skeleton::FromJoints(joints, vec_from<float>(inverseBind));
    });
    py::class_<AnimationPlayer, std::shared_ptr<AnimationPlayer>>(m, "AnimationPlayer")
        .def(py::init<std::shared_ptr<Node>>())
        .def_readwrite("PlaySpeed", &AnimationPlayer::PlaySpeed)
        .def_readwrite("BlendTime", &AnimationPlayer::BlendTime)
        .def("Play", &AnimationPlayer::Play)
        .def("Update", &AnimationPlayer::Update)
        .def("SetPlayhead", &AnimationPlayer::SetPlayhead)
        .def("Playhead", &AnimationPlayer::Playhead)
        .def("Finished", &AnimationPlayer::Finished)
        .def("IsPlaying", &AnimationPlayer::IsPlaying)
        .def("Pause", &AnimationPlayer::Pause)
        .def("Resume", &AnimationPlayer::Resume)
        .def("LoopCount", [](AnimationPlayer& a) { return a.loopCount; })
        .def("SampledPlayhead", [](AnimationPlayer& a) { return a.sampledPlayhead; })
        .def("Blending", &AnimationPlayer::Blending)
        .def("BlendStartIsBlend", &AnimationPlayer::BlendStartIsBlend)
        .def("BlendFraction", &AnimationPlayer::BlendFraction)
        .def("BlendFromPlayhead", [](AnimationPlayer& a) { return a.sampledChain.size() >= 2 ? a.sampledChain[a.sampledChain.size() - 2].playhead : 0.0f; })
        .def("BlendFrom", [](AnimationPlayer& a) { return a.sampledChain.size() >= 2 ? a.sampledChain[a.sampledChain.size() - 2].anim : std::shared_ptr<Animation>(); })
        // what the last update applied: [(Animation, playhead, blend fraction)], in the order the animations were played
        .def("SampledChain", [](AnimationPlayer& a) {
            py::list out;
            for (auto& L : a.sampledChain) out.append(py::make_tuple(L.anim, L.playhead, L.bp));
            return out;
        })
        .def("SetClock", [](AnimationPlayer& a, py::object fn) {      // time.Now() in milliseconds, for deterministic tests (None: the steady clock)
            if (fn.is_none()) a.NowMs = nullptr;
            else a.NowMs = [fn]() { py::gil_scoped_acquire g; return fn().cast<int64_t>(); };
        });
    m.def("NewAnimationPlayer", [](std::shared_ptr<Node> n) { return std::make_shared<AnimationPlayer>(n); });

    py::class_<Model, Node, std::shared_ptr<Model>>(m, "Model")
        .def(py::init<const std::string&, std::shared_ptr<Mesh>>())
        .def("Mesh", [](Model& s) { return s.mesh; })
        .def_readwrite("FrustumCulling", &Model::FrustumCulling)
        .def_property("Color", [](Model& s) { return col_to(s.Color); }, [](Model& s, const std::vector<float>& c) { s.Color = col_from(c); })
        .def_readwrite("Shadeless", &Model::Shadeless)
        .def_readwrite("LightGroup", &Model::lightGroup)
        .def_readwrite("HasVertexTransformFunction", &Model::HasVertexTransformFunction)
        .def_readwrite("HasVertexClipFunction", &Model::HasVertexClipFunction)
        .def("Skinned", [](Model& s) { return s.skinned; })
        // the per-joint Node.boneInfluence matrices (node.go:592-594), n x 16
        .def("SetBoneInfluences", [](Model& s, const farr& mats) { if (mats.size() % 16) throw std::invalid_argument("need 16 floats per bone"); s.SetBoneInfluences(vec_from<float>(mats)); })
        // bones posed on the device: skeleton, the node its root bones hang from (Model.SkinRoot), the armature's player (or None)
        .def("SetSkeleton", &Model::SetSkeleton, py::arg("skeleton"), py::arg("skin_root"), py::arg("player") = nullptr)
        .def("BakeAO", [](Model& s, GeometryDevice& dev, int targetChannel, float occlusionAngle, std::vector<float> occlusionColor,
                          std::vector<int> targetParts, std::vector<int> targetVertices) {
            if (occlusionColor.size() != 4) throw std::invalid_argument("BakeAO: occlusionColor must have 4 components");
            std::vector<float> out = s.BakeAO(dev, targetChannel, occlusionAngle, Color4{occlusionColor[0], occlusionColor[1], occlusionColor[2], occlusionColor[3]}, targetParts, targetVertices);
            py::array_t<float> a({(py::ssize_t)(out.size() / 4), (py::ssize_t)4});
            std::memcpy(a.mutable_data(), out.data(), out.size() * sizeof(float));
            return a;
        }, py::arg("device"), py::arg("targetChannel") = 0, py::arg("occlusionAngle") = 1.0471976f, py::arg("occlusionColor") = std::vector<float>{0.4f, 0.4f, 0.4f, 1.0f},
           py::arg("targetParts") = std::vector<int>{}, py::arg("targetVertices") = std::vector<int>{})
        .def("BakeLighting", [](Model& s, int targetChannel, Camera& cam, Scene& scene, py::list lights) {
            std::vector<std::shared_ptr<ILight>> ls;
            for (auto l : lights) ls.push_back(l.cast<std::shared_ptr<ILight>>());
            std::vector<float> out = s.BakeLighting(targetChannel, cam, scene, ls);
            py::array_t<float> a({(py::ssize_t)(out.size() / 4), (py::ssize_t)4});
            std::memcpy(a.mutable_data(), out.data(), out.size() * sizeof(float));
            return a;
        }, py::arg("targetChannel"), py::arg("camera"), py::arg("scene"), py::arg("lights"))
        .def("DynamicBatchAdd", [](Model& s, int part, py::args models) { std::vector<std::shared_ptr<Model>> v; for (auto k : models) v.push_back(k.cast<std::shared_ptr<Model>>()); s.DynamicBatchAdd(part, v); })
        .def("DynamicBatcher", &Model::DynamicBatcher);
    m.def("NewModel", [](const std::string& n, std::shared_ptr<Mesh> mesh) { return std::make_shared<Model>(n, mesh); });

    py::class_<World, std::shared_ptr<World>>(m, "World")
        .def(py::init<>())
        .def_property("FogColor", [](World& w) { return col_to(w.FogColor); }, [](World& w, const std::vector<float>& c) { w.FogColor = col_from(c); })
        .def_readwrite("FogMode", &World::FogMode)
        .def_readwrite("FogOn", &World::FogOn)
        .def_property("FogRange", [](World& w) { return std::vector<float>{w.FogRange[0], w.FogRange[1]}; },
                      [](World& w, const std::vector<float>& r) { w.FogRange[0] = r.at(0); w.FogRange[1] = r.at(1); })
        .def_readwrite("LightingOn", &World::LightingOn)
        .def_readwrite("AmbientLight", &World::ambientLight);

    py::class_<Scene, std::shared_ptr<Scene>>(m, "Scene")
        .def(py::init<const std::string&>(), py::arg("name") = "scene")
        .def_readwrite("Root", &Scene::Root)
        .def_readwrite("World", &Scene::world)
        .def("SetAutoSubdivisionLevels", [](Scene& s, const farr& levels) { if (levels.size() % 2) throw std::invalid_argument("levels: (distance_squared, max_size) pairs"); s.autosubdivisionLevels = vec_from<float>(levels); });
    m.def("NewScene", [](const std::string& n) { return std::make_shared<Scene>(n); });

    py::class_<GeometryDevice, std::shared_ptr<GeometryDevice>>(m, "GeometryDevice")
        .def(py::init<int>(), py::arg("device") = 0)
        .def_readonly("DeviceID", &GeometryDevice::deviceId)
        .def("Stream", [](GeometryDevice& d) { return (uintptr_t)t3d_debug_stream(d.ctx); })
        .def("SetFusedSort", [](GeometryDevice& d, int mode) { if (t3d_debug_set_fused_sort(d.ctx, mode) != T3D_OK) throw std::invalid_argument("SetFusedSort: 0, 1 or 2"); })
        .def("Handle", [](GeometryDevice& d) { return (uintptr_t)d.ctx; })      // the t3d_ctx*, for ctypes callers of the C ABI
        .def("SharedAlloc", [](GeometryDevice& d, uint64_t bytes) {
            void* p = nullptr; t3d_shared_handle h;
            const int rc = t3d_shared_alloc(d.ctx, bytes, &p, &h);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            return py::make_tuple((uintptr_t)p, py::bytes(reinterpret_cast<const char*>(h.bytes), sizeof h.bytes));
        })
        .def("SharedOpen", [](GeometryDevice& d, py::bytes handle) {
            std::string b = handle; t3d_shared_handle h;
            if (b.size() != sizeof h.bytes) throw std::invalid_argument("SharedOpen: handle must be 64 bytes");
            memcpy(h.bytes, b.data(), sizeof h.bytes);
            void* p = nullptr;
            const int rc = t3d_shared_open(d.ctx, &h, &p);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            return (uintptr_t)p;
        })
        .def("SetOutputTarget", [](GeometryDevice& d, uintptr_t color, uintptr_t depth, uintptr_t normal, uint64_t capacity) {
            const int rc = t3d_set_output_target(d.ctx, reinterpret_cast<t3d_vertex*>(color), reinterpret_cast<t3d_vertex*>(depth),
                                                 reinterpret_cast<t3d_vertex*>(normal), capacity);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
        })
        .def("GatherAlloc", [](GeometryDevice& d, uint32_t world, uint64_t vertexCapacity, uint32_t drawCapacity) {
            t3d_shared_handle h[3];
            const int rc = t3d_gather_alloc(d.ctx, world, vertexCapacity, drawCapacity, h);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            py::list out;
            for (auto& x : h) out.append(py::bytes(reinterpret_cast<const char*>(x.bytes), sizeof x.bytes));
            return out;
        })
        .def("GatherAttach", [](GeometryDevice& d, uint32_t world, uint32_t rank, uint32_t dst, uint64_t vertexCapacity, uint32_t drawCapacity, py::object handles) {
            t3d_shared_handle h[3]; bool have = !handles.is_none();
            if (have) {
                py::list l = handles;
                if (l.size() != 3) throw std::invalid_argument("GatherAttach: three handles");
                for (int i = 0; i < 3; i++) { std::string b = py::bytes(l[i]); if (b.size() != sizeof h[i].bytes) throw std::invalid_argument("GatherAttach: handles are 64 bytes"); memcpy(h[i].bytes, b.data(), sizeof h[i].bytes); }
            }
            const int rc = t3d_gather_attach(d.ctx, world, rank, dst, vertexCapacity, drawCapacity, have ? h : nullptr);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
        })
        .def("GatherDetach", [](GeometryDevice& d) { const int rc = t3d_gather_detach(d.ctx); if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx)); })
        .def("GatherResults", [](GeometryDevice& d) {
            t3d_gather_output o;
            const int rc = t3d_gather_results(d.ctx, &o);
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            py::dict r; r["color"] = (uintptr_t)o.color_list; r["depth"] = (uintptr_t)o.depth_list; r["capacity"] = o.vertex_capacity;
            py::list n; for (uint32_t k = 0; k < o.world; k++) n.append(o.n_vertices[(size_t)k * o.n_vertices_stride]);
            r["n_vertices"] = n;
            return r;
        })
        .def("GatherRankStats", [](GeometryDevice& d, uint32_t rank, uint32_t nDraws) {
            py::array_t<uint32_t> out({(py::ssize_t)nDraws, (py::ssize_t)4});
            const int rc = t3d_gather_rank_stats(d.ctx, rank, nDraws, reinterpret_cast<t3d_draw_stat*>(out.mutable_data()));
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            return out;
        })
        .def("ReadBones", [](GeometryDevice& d, uint32_t offset, uint32_t n) {
            py::array_t<float> out({(py::ssize_t)n, (py::ssize_t)16});
            int rc = t3d_debug_read_bones(d.ctx, offset, n, out.mutable_data());
            if (rc != T3D_OK) throw std::runtime_error(t3d_last_error(d.ctx));
            return out;
        })
        .def("ForceFullUpload", [](GeometryDevice& d, bool on) { t3d_debug_force_full_upload(d.ctx, on ? 1 : 0); })
        .def("SetStageTimings", [](GeometryDevice& d, bool on) { t3d_set_stage_timings(d.ctx, on ? 1 : 0); })
        .def("SetExactLighting", [](GeometryDevice& d, bool on) { const int rc = t3d_set_exact_lighting(d.ctx, on ? 1 : 0); if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx)); })
        .def("VertexKernelBytes", [](GeometryDevice& d) { return t3d_debug_vertex_kernel_bytes(d.ctx); })
        .def("FrameBytes", [](GeometryDevice& d) { uint64_t a = 0, b = 0; t3d_debug_frame_bytes(d.ctx, &a, &b); return py::make_tuple(a, b); })
        .def("ReadDrawVertices", [](GeometryDevice& d, uint32_t draw, uint32_t v0, uint32_t v1) {
            farr clip({(py::ssize_t)(v1 - v0), (py::ssize_t)4}), light({(py::ssize_t)(v1 - v0), (py::ssize_t)4});
            int rc = t3d_debug_read_draw_vertices(d.ctx, draw, v0, v1, clip.mutable_data(), light.mutable_data());
            if (rc != T3D_OK) throw T3DError(rc, t3d_last_error(d.ctx));
            return py::make_tuple(clip, light);
        });

    py::class_<Camera, Node, std::shared_ptr<Camera>>(m, "Camera")
        .def(py::init<const std::string&, int, int>())
        .def_readwrite("RenderDepth", &Camera::RenderDepth)
        .def_readwrite("RenderNormals", &Camera::RenderNormals)
        .def_readwrite("PerspectiveCorrectedTextureMapping", &Camera::PerspectiveCorrectedTextureMapping)
        .def_readwrite("VertexSnapping", &Camera::VertexSnapping)
        .def_readwrite("DepthMargin", &Camera::DepthMargin)
        .def_readwrite("Device", &Camera::device)
        .def_readwrite("DownloadResults", &Camera::downloadResults)
        .def_readwrite("AsyncDownload", &Camera::AsyncDownload)
        .def_readwrite("SkipUnusedDepthList", &Camera::SkipUnusedDepthList)
        .def("WaitDownload", [](Camera& c, int framesBack) { py::gil_scoped_release rel; c.WaitDownload(framesBack); }, py::arg("frames_back") = 0)
        .def_readwrite("UseInstancing", &Camera::UseInstancing)
        .def_readwrite("DeviceFrustumCulling", &Camera::DeviceFrustumCulling)
        .def_readwrite("MaxLightCount", &Camera::MaxLightCount)
        .def_readwrite("MaxLightCountFocusPoint", &Camera::MaxLightCountFocusPoint)
        .def("SetFieldOfView", [](Camera& c, float f) { c.fieldOfView = f; })
        .def("SetNear", [](Camera& c, float v) { c.near_ = v; })
        .def("SetFar", [](Camera& c, float v) { c.far_ = v; })
        .def("SetPerspective", [](Camera& c, bool p) { c.perspective = p; })
        .def("SetOrthoScale", [](Camera& c, float s) { c.orthoScale = s; })
        .def("SetShard", [](Camera& c, int rank, int count) { if (count < 1 || rank < 0 || rank >= count) throw std::invalid_argument("bad shard"); c.shardRank = rank; c.shardCount = count; })
        .def("Width", [](Camera& c) { return c.width; })
        .def("Height", [](Camera& c) { return c.height; })
        .def("ViewMatrix", [](Camera& c) { return mat_to(c.ViewMatrix()); })
        .def("Projection", [](Camera& c) { return mat_to(c.Projection()); })
        .def("Clear", &Camera::Clear)
        .def("RenderScene", [](Camera& c, Scene& s) { py::gil_scoped_release rel; c.RenderScene(s); })
        .def("RenderNodes", [](Camera& c, Scene& s, Node& root) { py::gil_scoped_release rel; c.RenderNodes(s, root); })
        // frames_back = 1: the lists of the Render before the last one (AsyncDownload keeps them valid for one more frame)
        .def("ColorVertexList", [](Camera& c, int fb) { c.WaitDownload(fb); const t3d_frame_output& o = fb ? c.prevOutput : c.output; return vertex_list(o.color_list, o.n_vertices, !c.downloadResults); }, py::arg("frames_back") = 0)
        .def("DepthVertexList", [](Camera& c, int fb) { c.WaitDownload(fb); const t3d_frame_output& o = fb ? c.prevOutput : c.output; return vertex_list(o.depth_list, o.n_vertices, !c.downloadResults); }, py::arg("frames_back") = 0)
        .def("NormalVertexList", [](Camera& c, int fb) { c.WaitDownload(fb); const t3d_frame_output& o = fb ? c.prevOutput : c.output; return vertex_list(o.normal_list, o.normal_list ? o.n_vertices : 0, !c.downloadResults); }, py::arg("frames_back") = 0)
        .def("IndexList", [](Camera& c, uint64_t n) {
            py::array_t<uint16_t> a((py::ssize_t)n);
            for (uint64_t i = 0; i < n; i++) a.mutable_data()[i] = c.output.index_list ? c.output.index_list[i] : (uint16_t)i;
            return a;
        })
        .def("TriangleIDs", [](Camera& c) {
            if (!c.downloadResults) throw std::runtime_error("lists are device-resident");
            c.WaitDownload(0);
            py::array_t<uint32_t> a((py::ssize_t)(c.output.n_vertices / 3));
            if (c.output.n_vertices) std::memcpy(a.mutable_data(), c.output.tri_ids, (c.output.n_vertices / 3) * 4);
            return a;
        })
        .def("VertexCount", [](Camera& c) { return c.output.n_vertices; })
        .def("DevicePointers", [](Camera& c) {
            py::dict d;
            d["color"] = (uintptr_t)c.output.color_list; d["depth"] = (uintptr_t)c.output.depth_list;
            d["normal"] = (uintptr_t)c.output.normal_list; d["tri_ids"] = (uintptr_t)c.output.tri_ids;
            d["n_vertices"] = c.output.n_vertices; d["on_device"] = !c.downloadResults;
            return d;
        })
        .def("BoneOffsets", [](Camera& c) { return c.lastBoneOffsets; })
        .def("Renders", [](Camera& c) {
            py::list out;
            for (auto& r : c.renders) {
                py::dict d; d["model"] = r.model->name; d["part"] = r.part; d["flush"] = r.flush; d["first_vertex"] = r.firstVertex;
                d["vertex_count"] = r.vertexCount; d["total_tris"] = r.totalTris; d["min_depth"] = r.minDepth; d["max_depth"] = r.maxDepth; d["draw_index"] = r.drawIndex;
                out.append(d);
            }
            return out;
        })
        .def("Flushes", [](Camera& c) {
            py::list out;
            for (size_t i = 0; i < c.flushes.size(); i++) {
                auto& f = c.flushes[i];
                py::dict d; d["model"] = f.model->name; d["part"] = f.part; d["first_vertex"] = f.firstVertex; d["vertex_count"] = f.vertexCount;
                d["global_flush"] = c.localFlushIds[i];
                out.append(d);
            }
            return out;
        })
        .def("FlushOwnerRanks", [](Camera& c) { return c.flushOwnerRank; })
        .def("ShardDrawFlushes", [](Camera& c) { return c.shardDrawFlush; })
        .def("DebugInfo", [](Camera& c) {
            py::dict d; d["drawnParts"] = c.debugInfo.drawnParts; d["totalParts"] = c.debugInfo.totalParts; d["batchedParts"] = c.debugInfo.batchedParts;
            d["drawnTris"] = c.debugInfo.drawnTris; d["totalTris"] = c.debugInfo.totalTris; d["lightCount"] = c.debugInfo.lightCount;
            d["activeLightCount"] = c.debugInfo.activeLightCount;
            return d;
        })
        .def("Timings", [](Camera& c) {
            py::dict d; d["upload_ms"] = c.timings.upload_ms; d["prep_ms"] = c.timings.prep_ms; d["vertex_ms"] = c.timings.vertex_ms; d["cull_ms"] = c.timings.cull_ms;
            d["sort_ms"] = c.timings.sort_ms; d["emit_ms"] = c.timings.emit_ms; d["download_ms"] = c.timings.download_ms; d["total_ms"] = c.timings.total_ms; d["gather_ms"] = c.timings.gather_ms;
            return d;
        })
        .def("LaunchCount", [](Camera& c) { return c.launchCount; });
    m.def("AssignFlushesToRanks", &AssignFlushesToRanks, py::arg("flush_triangles"), py::arg("rank_count"));
    m.def("NewCamera", [](const std::string& n, int w, int h) { return std::make_shared<Camera>(n, w, h); });
}
