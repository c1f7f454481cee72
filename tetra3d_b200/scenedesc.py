"""Neutral, pure-data description of a Tetra3D scene (numpy arrays + dataclasses).

One description can be materialised twice: as host-mirror objects driving the GPU path
(`tetra3d_b200.build.build_scene`) and, in tests / bench baselines only, as the flat structs the CPU
oracle reads (`oracle.binding`).  Field names follow the reference (material.go:66-131,
model.go:20-61, mesh.go:260-302, camera.go:259-332, world.go:25-42, light.go).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F32 = np.float32


def identity4() -> np.ndarray:
    return np.eye(4, dtype=F32)


@dataclass
class MaterialDesc:
    color: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    backface_culling: bool = True
    triangle_sort_mode: int = 0
    shadeless: bool = False
    lighting_mode: int = 0
    light_volume_shading_mode: int = 0
    billboard_enabled: bool = False
    billboard_lock_x: bool = False
    billboard_lock_y: bool = False
    billboard_lock_z: bool = False
    billboard_up_direction: int = 0
    billboarded_depth_mode: int = 0
    transparency_mode: int = 0
    blend_source_over: bool = True
    visible: bool = True
    render_order: int = 0
    use_texture: bool = False
    tex_w: int = 0
    tex_h: int = 0


@dataclass
class PartDesc:
    vertex_start: int
    vertex_end: int
    tri_start: int
    tri_end: int                       # inclusive, as MeshPart.TriangleEnd (mesh.go:2189)
    material: Optional[MaterialDesc] = field(default_factory=MaterialDesc)


@dataclass
class MeshDesc:
    positions: np.ndarray              # (n, 3) float32
    normals: np.ndarray                # (n, 3)
    uvs: np.ndarray                    # (n, 2)
    indices: np.ndarray                # (t, 3) uint32
    parts: List[PartDesc]
    colors: Optional[np.ndarray] = None          # (n, 4): the active colour channel
    active_color_channel: int = -1
    weights: Optional[np.ndarray] = None         # (n, 4)
    bones: Optional[np.ndarray] = None           # (n, 4) uint16
    n_influences: Optional[np.ndarray] = None    # (n,) uint8
    key_positions: Optional[np.ndarray] = None   # (k, n, 3)
    key_normals: Optional[np.ndarray] = None     # (k, n, 3)
    key_weights: Optional[np.ndarray] = None     # (k,)
    name: str = "mesh"
    # Auto-subdivision (Mesh.SetAutoSubdivide, mesh.go:447-485): `indices` then hold every base triangle followed by its
    # pre-built sub-triangles level by level -- the order MeshPart.forEachTri(true, ...) visits them in (mesh.go:2238-2256)
    auto_subdivide: bool = False
    tri_centers: Optional[np.ndarray] = None     # (t, 3) Triangle.Center of every entry (sub-triangles use their own formula)
    tri_sub_base: Optional[np.ndarray] = None    # (t,) int32: index of the base triangle (a base triangle names itself)
    tri_sub_level: Optional[np.ndarray] = None   # (t,) int8: -1 base triangle, else index into the base's subdivisionLevels
    tri_n_levels: Optional[np.ndarray] = None    # (t,) uint8: len(subdivisionLevels) of a base triangle (0: never subdivided)
    tri_max_span: Optional[np.ndarray] = None    # (t,) Triangle.MaxSpan
    tri_normal: Optional[np.ndarray] = None      # (t, 3) Triangle.Normal (read for base triangles only)

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=F32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(self.normals, dtype=F32).reshape(-1, 3)
        self.uvs = np.ascontiguousarray(self.uvs, dtype=F32).reshape(-1, 2)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32).reshape(-1, 3)
        if self.colors is not None:
            self.colors = np.ascontiguousarray(self.colors, dtype=F32).reshape(-1, 4)
        if self.weights is not None:
            self.weights = np.ascontiguousarray(self.weights, dtype=F32).reshape(-1, 4)
            self.bones = np.ascontiguousarray(self.bones, dtype=np.uint16).reshape(-1, 4)
            if self.n_influences is None:
                self.n_influences = np.full(len(self.positions), 4, dtype=np.uint8)
            self.n_influences = np.ascontiguousarray(self.n_influences, dtype=np.uint8)
        if self.key_weights is not None:
            self.key_positions = np.ascontiguousarray(self.key_positions, dtype=F32)
            self.key_normals = np.ascontiguousarray(self.key_normals, dtype=F32)
            self.key_weights = np.ascontiguousarray(self.key_weights, dtype=F32)

    @property
    def n_verts(self) -> int:
        return len(self.positions)

    @property
    def n_tris(self) -> int:
        return len(self.indices)


@dataclass(eq=False)
class NodeDesc:
    position: Sequence[float] = (0.0, 0.0, 0.0)
    scale: Sequence[float] = (1.0, 1.0, 1.0)
    rotation: np.ndarray = field(default_factory=identity4)     # row-major Matrix4
    parent: Optional["NodeDesc"] = None
    name: str = "node"


@dataclass
class VolumeDesc:
    dim_min: Sequence[float]
    dim_max: Sequence[float]
    cell_size: Sequence[float]
    cell_count: Sequence[int]
    colors: np.ndarray                 # (cx*cy*cz, 3), index x + y*cx + z*cx*cy
    offset: Sequence[float] = (0.0, 0.0, 0.0)
    light_bias: Sequence[float] = (0.0, 0.0, 0.0)
    cell_lights: Optional[dict] = None  # (x, y, z) -> [LightDesc, ...]: LightVolumeCell.SetUseLights (light.go:1508-1514)


LIGHT_AMBIENT, LIGHT_POINT, LIGHT_DIRECTIONAL, LIGHT_VOLUME = 0, 1, 2, 3


@dataclass(eq=False)
class LightDesc:
    type: int
    node: NodeDesc = field(default_factory=NodeDesc)
    color: Sequence[float] = (1.0, 1.0, 1.0)
    energy: float = 1.0
    visible: bool = True
    range: float = 0.0
    volume: Optional[VolumeDesc] = None
    in_scene: bool = True              # under scene.Root (a LightGroup may hold lights that are not)
    name: str = "light"


@dataclass(eq=False)
class LightGroupDesc:
    lights: List[LightDesc]
    active: bool = True


@dataclass(eq=False)
class SkeletonDesc:
    """The joint list of a skin (gltf.go:1452-1511): parents first; a parent of -1 hangs off the armature node."""
    names: List[str]
    parent: np.ndarray              # (n,) int32
    inverse_bind: np.ndarray        # (n, 16)
    rest_position: np.ndarray       # (n, 3)
    rest_scale: np.ndarray          # (n, 3)
    rest_rotation: np.ndarray       # (n, 16)

    def __post_init__(self):
        self.parent = np.ascontiguousarray(self.parent, np.int32).reshape(-1)
        n = len(self.parent)
        self.inverse_bind = np.ascontiguousarray(self.inverse_bind, np.float32).reshape(n, 16)
        self.rest_position = np.ascontiguousarray(self.rest_position, np.float32).reshape(n, 3)
        self.rest_scale = np.ascontiguousarray(self.rest_scale, np.float32).reshape(n, 3)
        self.rest_rotation = np.ascontiguousarray(self.rest_rotation, np.float32).reshape(n, 16)


TRACK_POSITION, TRACK_SCALE, TRACK_ROTATION = 0, 1, 2
INTERPOLATION_LINEAR, INTERPOLATION_CONSTANT, INTERPOLATION_CUBIC = 0, 1, 2


@dataclass(eq=False)
class TrackDesc:
    times: np.ndarray               # (k,)
    values: np.ndarray              # (k, 4): x y z _ or a quaternion x y z w
    interpolation: int = INTERPOLATION_LINEAR

    def __post_init__(self):
        self.times = np.ascontiguousarray(self.times, np.float32).reshape(-1)
        v = np.asarray(self.values, np.float32).reshape(len(self.times), -1)
        if v.shape[1] == 3:
            v = np.concatenate([v, np.zeros((len(v), 1), np.float32)], axis=1)
        self.values = np.ascontiguousarray(v, np.float32)


@dataclass(eq=False)
class AnimationDesc:
    """animation.go:204-222: channel name (= bone name) -> {track type -> TrackDesc}."""
    name: str
    length: float
    channels: dict
    loop_mode: int = 0


@dataclass(eq=False)
class PoseDesc:
    """An armature whose bones are posed from an animation at `playhead` (None: rest pose); `root` is the armature node
    the skeleton's root bones are children of (Model.SkinRoot, model.go:33)."""
    skeleton: SkeletonDesc
    animation: Optional[AnimationDesc] = None
    playhead: float = 0.0
    root: NodeDesc = field(default_factory=NodeDesc)
    # The player is in the middle of a blend (AnimationPlayer.BlendTime > 0, animation.go:548-554, 852-907): blend_ms
    # milliseconds after Play(animation), coming from `blend_from` last sampled at from_playhead (None: nothing was animated).
    blend_ms: Optional[int] = None
    blend_time: float = 1.0
    blend_from: Optional[AnimationDesc] = None
    from_playhead: float = 0.0

    def blend_fraction(self) -> float:
        """bp of animation.go:854-857: float32(milliseconds) / (BlendTime * 1000), clamped to 1."""
        bp = np.float32(int(self.blend_ms)) / (np.float32(self.blend_time) * np.float32(1000.0))
        return float(min(bp, np.float32(1.0)))


@dataclass(eq=False)
class ModelDesc:
    mesh: MeshDesc
    node: NodeDesc = field(default_factory=NodeDesc)
    color: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    shadeless: bool = False
    visible: bool = True
    frustum_culling: bool = True
    bone_influences: Optional[np.ndarray] = None     # (n_bones, 16): makes the model skinned
    pose: Optional[PoseDesc] = None                  # instead of bone_influences: bones posed from a skeleton + animation
    light_group: Optional[LightGroupDesc] = None
    batch_owner: Optional["ModelDesc"] = None        # Model.DynamicBatchOwner
    batch_owner_part: int = 0
    name: str = "model"


def auto_subdivide(mesh: MeshDesc, levels, max_count: int = 3) -> MeshDesc:
    """Mesh.SetAutoSubdivide(true) (mesh.go:447-485) + Triangle.performSubdivision (mesh.go:1758-1861) at load time: every
    triangle whose MaxSpan exceeds the smallest MaximumTriangleSize of `levels` [(distance_squared, max_size), ...] gets up to
    `max_count` pre-built levels of 4-way splits; the new vertices (edge midpoints: Lerp 0.5 of position, normal, uv, colour)
    are appended to the vertex arrays.  Returns a new MeshDesc in the expanded layout described on MeshDesc."""
    assert mesh.weights is None and mesh.key_weights is None, "auto-subdivision: static meshes only"
    P = [np.asarray(v, F32) for v in mesh.positions]; N = [np.asarray(v, F32) for v in mesh.normals]; U = [np.asarray(v, F32) for v in mesh.uvs]
    Cc = [np.asarray(v, F32) for v in mesh.colors] if mesh.colors is not None else None
    half = F32(0.5)
    lerp = lambda a, b: (a + ((b - a) * half)).astype(F32)                  # Vector3.Lerp, vector.go:559-565
    third = F32(1.0 / 3.0)
    smallest = min(F32(ms) for _, ms in levels)

    def span(a, b, c):                                                      # Triangle.RecalculateCenter, mesh.go:2018-2022
        mag = lambda v: F32(np.sqrt(np.float64(F32(F32(v[0] * v[0]) + F32(v[1] * v[1])) + F32(v[2] * v[2]))))
        return max(mag(P[a] - P[b]), mag(P[b] - P[c]), mag(P[c] - P[a]))

    def normal(a, b, c):                                                    # calculateNormal, mesh.go:2171-2181 (any unit normal will do: it is data)
        n = np.cross((P[b] - P[a]).astype(np.float64), (P[c] - P[b]).astype(np.float64))
        l = np.linalg.norm(n)
        return (n / l).astype(F32) if l > 0 else N[a]

    out_idx, out_ctr, out_base, out_lvl, out_nl, out_span, out_nrm = [], [], [], [], [], [], []
    parts = []
    for part in mesh.parts:
        t_start = len(out_idx)
        for ti in range(part.tri_start, part.tri_end + 1):
            a, b, c = (int(x) for x in mesh.indices[ti])
            base_slot = len(out_idx)
            ms = span(a, b, c)
            ctr = ((P[a] + P[b] + P[c]) / F32(3)).astype(F32)                # RecalculateCenter, mesh.go:1980-1982
            subs = [[] for _ in range(max_count)]

            def subdivide(level, ia, ib, ic, tspan):                        # performSubdivision
                if level == max_count or tspan <= smallest:
                    return
                si = len(P)
                P.extend([lerp(P[ia], P[ib]), lerp(P[ib], P[ic]), lerp(P[ic], P[ia])])
                N.extend([lerp(N[ia], N[ib]), lerp(N[ib], N[ic]), lerp(N[ic], N[ia])])
                U.extend([lerp(U[ia], U[ib]), lerp(U[ib], U[ic]), lerp(U[ic], U[ia])])
                if Cc is not None:
                    Cc.extend([lerp(Cc[ia], Cc[ib]), lerp(Cc[ib], Cc[ic]), lerp(Cc[ic], Cc[ia])])
                kids = [(ia, si, si + 2), (si, ib, si + 1), (si + 1, ic, si + 2), (si, si + 1, si + 2)]
                hs = F32(tspan / F32(2))
                for k in kids:
                    kc = (((P[k[0]] + P[k[1]]).astype(F32) + P[k[2]]).astype(F32) * third).astype(F32)
                    subs[level].append((k, kc, hs))
                for k in kids:
                    subdivide(level + 1, k[0], k[1], k[2], hs)

            n_levels = 0
            if ms > smallest:
                n_levels = max_count
                subdivide(0, a, b, c, ms)
            out_idx.append((a, b, c)); out_ctr.append(ctr); out_base.append(base_slot); out_lvl.append(-1)
            out_nl.append(n_levels); out_span.append(ms); out_nrm.append(normal(a, b, c))
            for lv in range(n_levels):
                for k, kc, hs in subs[lv]:
                    out_idx.append(k); out_ctr.append(kc); out_base.append(base_slot); out_lvl.append(lv)
                    out_nl.append(0); out_span.append(hs); out_nrm.append(np.zeros(3, F32))
        parts.append(PartDesc(part.vertex_start, part.vertex_end, t_start, len(out_idx) - 1, part.material))
    return MeshDesc(np.asarray(P, F32), np.asarray(N, F32), np.asarray(U, F32), np.asarray(out_idx, np.uint32), parts,
                    colors=np.asarray(Cc, F32) if Cc is not None else None, active_color_channel=mesh.active_color_channel,
                    name=mesh.name + "_subdiv", auto_subdivide=True, tri_centers=np.asarray(out_ctr, F32),
                    tri_sub_base=np.asarray(out_base, np.int32), tri_sub_level=np.asarray(out_lvl, np.int8),
                    tri_n_levels=np.asarray(out_nl, np.uint8), tri_max_span=np.asarray(out_span, F32), tri_normal=np.asarray(out_nrm, F32))


@dataclass
class CameraDesc:
    node: NodeDesc = field(default_factory=lambda: NodeDesc(position=(0.0, 0.0, 10.0)))
    width: int = 640
    height: int = 360
    field_of_view: float = 60.0
    near: float = 0.1
    far: float = 100.0
    perspective: bool = True
    ortho_scale: float = 20.0
    depth_margin: float = 0.04
    vertex_snapping: float = 0.0
    render_depth: bool = True
    render_normals: bool = False
    perspective_corrected_texture_mapping: bool = False
    max_light_count: int = 0                        # Camera.MaxLightCount (camera.go:1474); 0 = unlimited
    max_light_focus: Optional["NodeDesc"] = None    # Camera.MaxLightCountFocusPoint; None = the camera itself


@dataclass
class WorldDesc:
    present: bool = True
    lighting_on: bool = True
    fog_on: bool = True
    fog_mode: int = 0
    fog_color: Sequence[float] = (0.0, 0.0, 0.0, 0.0)
    fog_range: Sequence[float] = (0.0, 1.0)
    ambient_color: Sequence[float] = (1.0, 1.0, 1.0)
    ambient_energy: float = 0.0
    ambient_visible: bool = True


@dataclass
class SceneDesc:
    models: List[ModelDesc]            # in the depth-first order RenderNodes visits them
    lights: List[LightDesc] = field(default_factory=list)
    camera: CameraDesc = field(default_factory=CameraDesc)
    world: WorldDesc = field(default_factory=WorldDesc)
    name: str = "scene"
    autosubdivision_levels: Sequence = ()      # scene.library.autosubdivisionLevels (library.go:206-209): [(distance_squared, max_triangle_size), ...]

    def total_triangles(self) -> int:
        return sum(m.mesh.n_tris for m in self.models if m.visible)

    def total_vertices(self) -> int:
        return sum(m.mesh.n_verts for m in self.models if m.visible)
