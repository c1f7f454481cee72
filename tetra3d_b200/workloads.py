"""Synthetic scenes for the BASELINE.json configs and small parity cases.

Everything is seeded numpy (default_rng / PCG64); the same SceneDesc feeds the GPU path and, in tests
and baselines, the CPU oracle.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

from .scenedesc import (INTERPOLATION_CONSTANT, LIGHT_AMBIENT, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_VOLUME, TRACK_POSITION, TRACK_ROTATION,
                        TRACK_SCALE, AnimationDesc, CameraDesc, LightDesc, LightGroupDesc, MaterialDesc, MeshDesc, ModelDesc, NodeDesc, PartDesc,
                        PoseDesc, SceneDesc, SkeletonDesc, TrackDesc, VolumeDesc, WorldDesc)

F32 = np.float32


# ---- small math helpers (float32 numpy; these only build INPUT data) ----------------------------------
def rotation_matrix(axis, angle: float) -> np.ndarray:
    """Row-major Matrix4 in the layout of NewMatrix4Rotate (matrix.go:100-128)."""
    x, y, z = [float(a) for a in axis]
    n = math.sqrt(x * x + y * y + z * z)
    x, y, z = x / n, y / n, z / n
    s, c = math.sin(angle), math.cos(angle)
    m = 1 - c
    r = np.eye(4, dtype=F32)
    r[0, 0] = m * x * x + c; r[0, 1] = m * x * y + z * s; r[0, 2] = m * z * x - y * s
    r[1, 0] = m * x * y - z * s; r[1, 1] = m * y * y + c; r[1, 2] = m * y * z + x * s
    r[2, 0] = m * z * x + y * s; r[2, 1] = m * y * z - x * s; r[2, 2] = m * z * z + c
    return r


def triangle_normals(pos: np.ndarray, idx: np.ndarray) -> np.ndarray:
    a, b, c = pos[idx[:, 0]], pos[idx[:, 1]], pos[idx[:, 2]]
    n = np.cross(b - a, c - b).astype(F32)
    l = np.linalg.norm(n, axis=1, keepdims=True)
    l[l < 1e-8] = 1
    return (n / l).astype(F32)


def flat_normals(pos: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Mesh.AutoNormal semantics (mesh.go:755-764): each vertex takes the normal of the last triangle using it."""
    tn = triangle_normals(pos, idx)
    nrm = np.zeros_like(pos)
    for k in range(3):
        nrm[idx[:, k]] = tn
    return nrm


def smooth_normals(pos: np.ndarray, idx: np.ndarray) -> np.ndarray:
    tn = triangle_normals(pos, idx)
    nrm = np.zeros_like(pos)
    for k in range(3):
        np.add.at(nrm, idx[:, k], tn)
    l = np.linalg.norm(nrm, axis=1, keepdims=True)
    l[l < 1e-8] = 1
    return (nrm / l).astype(F32)


# ---- meshes --------------------------------------------------------------------------------------------
_CUBE_V = np.array([
    [-1, 1, -1, 0, 0], [1, 1, 1, 1, 1], [1, 1, -1, 1, 0], [-1, 1, 1, 0, 1],
    [1, -1, -1, 1, 0], [1, -1, 1, 1, 1], [-1, -1, -1, 0, 0], [-1, -1, 1, 0, 1],
    [-1, 1, 1, 0, 0], [1, -1, 1, 1, 1], [1, 1, 1, 1, 0], [-1, -1, 1, 0, 1],
    [1, 1, -1, 1, 0], [1, -1, -1, 1, 1], [-1, 1, -1, 0, 0], [-1, -1, -1, 0, 1],
    [1, 1, -1, 1, 0], [1, 1, 1, 1, 1], [1, -1, -1, 0, 0], [1, -1, 1, 0, 1],
    [-1, -1, -1, 0, 0], [-1, 1, 1, 1, 1], [-1, 1, -1, 1, 0], [-1, -1, 1, 0, 1]], dtype=F32)
_CUBE_I = np.array([0, 1, 2, 3, 1, 0, 4, 5, 6, 6, 5, 7, 8, 9, 10, 11, 9, 8, 12, 13, 14, 14, 13, 15,
                    16, 17, 18, 18, 17, 19, 20, 21, 22, 23, 21, 20], dtype=np.uint32).reshape(-1, 3)


def cube_mesh(w=2.0, h=2.0, d=2.0, copies: int = 1, material: Optional[MaterialDesc] = None) -> MeshDesc:
    """NewCubeMesh (mesh.go:1251-1342); `copies` stacks co-located cubes into one MeshPart like the
    StaticMerge loop of examples/stress/main.go:72-74."""
    pos = _CUBE_V[:, :3] * np.array([w / 2, h / 2, d / 2], dtype=F32)
    uv = _CUBE_V[:, 3:5].copy()
    nrm = flat_normals(pos, _CUBE_I)
    if copies > 1:
        pos = np.tile(pos, (copies, 1)); uv = np.tile(uv, (copies, 1)); nrm = np.tile(nrm, (copies, 1))
        idx = (np.tile(_CUBE_I, (copies, 1)).reshape(copies, 12, 3) + (np.arange(copies, dtype=np.uint32) * 24)[:, None, None]).reshape(-1, 3)
    else:
        idx = _CUBE_I.copy()
    mat = material if material is not None else MaterialDesc()
    return MeshDesc(pos, nrm, uv, idx.astype(np.uint32), [PartDesc(0, len(pos), 0, len(idx) - 1, mat)], name="Cube")


def plane_mesh(w=2.0, d=2.0, material: Optional[MaterialDesc] = None, colors: Optional[np.ndarray] = None) -> MeshDesc:
    """A 4-vertex / 2-triangle quad in the XY plane facing +Z (the particle quad of config 4)."""
    pos = np.array([[-w / 2, -d / 2, 0], [w / 2, -d / 2, 0], [w / 2, d / 2, 0], [-w / 2, d / 2, 0]], dtype=F32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=F32)
    idx = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    nrm = np.tile(np.array([[0, 0, 1]], dtype=F32), (4, 1))
    return MeshDesc(pos, nrm, uv, idx, [PartDesc(0, 4, 0, 1, material if material is not None else MaterialDesc())],
                    colors=colors, active_color_channel=0 if colors is not None else -1, name="Plane")


def grid_mesh(nx: int, nz: int, x_range=(-50.0, 50.0), z_range=(-25.0, 25.0), rows_per_part: int = 10, seed: int = 1,
              material: Optional[MaterialDesc] = None) -> MeshDesc:
    """Height-field grid of (nx+1) x (nz+1) vertices split into MeshParts of `rows_per_part` quad rows, each part
    with its own contiguous vertex range (1000 x 500 quads -> 550 550 v / 1 000 000 t)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(x_range[0], x_range[1], nx + 1, dtype=F32)
    n_parts = (nz + rows_per_part - 1) // rows_per_part
    pos_l, uv_l, idx_l, parts = [], [], [], []
    vbase = 0; tbase = 0
    mat = material if material is not None else MaterialDesc()
    zs_all = np.linspace(z_range[0], z_range[1], nz + 1, dtype=F32)
    noise = rng.uniform(-1, 1, size=(nz + 1, nx + 1)).astype(F32)
    for p in range(n_parts):
        r0 = p * rows_per_part; r1 = min(nz, r0 + rows_per_part)
        zs = zs_all[r0:r1 + 1]
        X, Z = np.meshgrid(xs, zs)
        Y = (2 * np.sin(0.3 * X) * np.cos(0.3 * Z) + 0.25 * noise[r0:r1 + 1]).astype(F32)
        pos = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(F32)
        uv = np.stack([(X - x_range[0]) / (x_range[1] - x_range[0]), (Z - z_range[0]) / (z_range[1] - z_range[0])], axis=-1).reshape(-1, 2).astype(F32)
        rows = r1 - r0
        i = np.arange(nx, dtype=np.uint32)[None, :] + (np.arange(rows, dtype=np.uint32) * (nx + 1))[:, None]
        a = i; b = i + 1; c = i + (nx + 1); d = c + 1
        # counter-clockwise seen from +Y
        tris = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], axis=2).reshape(-1, 3)
        idx_l.append(tris + vbase)
        pos_l.append(pos); uv_l.append(uv)
        parts.append(PartDesc(vbase, vbase + len(pos), tbase, tbase + len(tris) - 1, mat))
        vbase += len(pos); tbase += len(tris)
    pos = np.concatenate(pos_l); uv = np.concatenate(uv_l); idx = np.concatenate(idx_l).astype(np.uint32)
    nrm = smooth_normals(pos, idx)
    return MeshDesc(pos, nrm, uv, idx, parts, name="Grid")


def icosphere_mesh(subdivisions: int = 4, radius: float = 1.0, material: Optional[MaterialDesc] = None) -> MeshDesc:
    """Icosphere with welded vertices; subdivisions=4 gives 5120 triangles (config 5 uses 5120-triangle spheres)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdivisions):
        cache = {}
        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m)); cache[key] = len(verts) - 1
            return cache[key]
        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    pos = (np.array(verts) * radius).astype(F32)
    idx = np.array(faces, dtype=np.uint32)
    nrm = (pos / np.linalg.norm(pos, axis=1, keepdims=True)).astype(F32)
    uv = np.stack([0.5 + np.arctan2(nrm[:, 2], nrm[:, 0]) / (2 * np.pi), 0.5 - np.arcsin(np.clip(nrm[:, 1], -1, 1)) / np.pi], axis=1).astype(F32)
    return MeshDesc(pos, nrm, uv, idx, [PartDesc(0, len(pos), 0, len(idx) - 1, material if material is not None else MaterialDesc())], name="Icosphere")


def tube_character_mesh(rings: int = 17, sides: int = 24, n_bones: int = 8, height: float = 2.0, radius: float = 0.25,
                        material: Optional[MaterialDesc] = None) -> Tuple[MeshDesc, np.ndarray]:
    """The synthetic skinned character of config 3: a tube with a duplicated seam column (rings x (sides+1) vertices),
    weighted to its 2 nearest bones of an n_bones chain along +Y.  Returns (mesh, bone rest positions y)."""
    ys = np.linspace(0, height, rings, dtype=F32)
    ang = np.linspace(0, 2 * np.pi, sides + 1, dtype=F32)
    Y, A = np.meshgrid(ys, ang, indexing="ij")
    pos = np.stack([radius * np.cos(A), Y, radius * np.sin(A)], -1).reshape(-1, 3).astype(F32)
    nrm = np.stack([np.cos(A), np.zeros_like(A), np.sin(A)], -1).reshape(-1, 3).astype(F32)
    uv = np.stack([A / (2 * np.pi), Y / height], -1).reshape(-1, 2).astype(F32)
    i = np.arange(sides, dtype=np.uint32)[None, :] + (np.arange(rings - 1, dtype=np.uint32) * (sides + 1))[:, None]
    a = i; b = i + 1; c = i + (sides + 1); d = c + 1
    idx = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], axis=2).reshape(-1, 3).astype(np.uint32)
    bone_y = np.linspace(0, height, n_bones, dtype=F32)
    dist = np.abs(pos[:, 1:2] - bone_y[None, :])
    order = np.argsort(dist, axis=1)[:, :2]
    d2 = np.take_along_axis(dist, order, axis=1)
    span = bone_y[1] - bone_y[0]
    w0 = np.clip(1 - d2[:, 0] / span, 0, 1).astype(F32)
    weights = np.zeros((len(pos), 4), F32); bones = np.zeros((len(pos), 4), np.uint16)
    weights[:, 0] = w0; weights[:, 1] = (1 - w0).astype(F32)
    bones[:, :2] = order.astype(np.uint16)
    ninf = np.full(len(pos), 2, np.uint8)
    mesh = MeshDesc(pos, nrm, uv, idx, [PartDesc(0, len(pos), 0, len(idx) - 1, material if material is not None else MaterialDesc())],
                    weights=weights, bones=bones, n_influences=ninf, name="TubeCharacter")
    return mesh, bone_y


def chain_bone_influences(bone_y: np.ndarray, base_pos, phase: float, t: float = 0.0) -> np.ndarray:
    """boneInfluence = inverseBind * boneWorld (node.go:592-594) for a bending chain: bone b is rotated about Z by
    theta_b = 0.4*sin(t + phase + 0.5 b) around its rest position, the chain translated to base_pos."""
    out = np.zeros((len(bone_y), 16), F32)
    parent = np.eye(4, dtype=np.float64)
    parent[3, :3] = base_pos
    prev_y = 0.0
    for b, y in enumerate(bone_y):
        theta = 0.4 * math.sin(t + phase + 0.5 * b)
        local = rotation_matrix((0, 0, 1), theta).astype(np.float64)
        local[3, 1] = float(y) - prev_y
        world = local @ parent
        inv_bind = np.eye(4, dtype=np.float64); inv_bind[3, 1] = -float(y)
        out[b] = (inv_bind @ world).astype(F32).reshape(16)
        parent = world; prev_y = float(y)
    return out


def chain_skeleton(bone_y: np.ndarray) -> "SkeletonDesc":
    """The bone chain of tube_character_mesh as a skin's joint list: bone b is a child of bone b - 1, sits at (0, bone_y[b], 0)
    in the bind pose, inverseBind = Translate(0, -bone_y[b], 0)."""
    n = len(bone_y)
    inv = np.tile(np.eye(4, dtype=F32).reshape(1, 16), (n, 1))
    inv[:, 13] = -bone_y
    pos = np.zeros((n, 3), F32)
    pos[:, 1] = np.diff(np.concatenate([[0.0], bone_y.astype(np.float64)])).astype(F32)
    return SkeletonDesc(names=["bone%d" % b for b in range(n)], parent=np.arange(-1, n - 1, dtype=np.int32), inverse_bind=inv,
                        rest_position=pos, rest_scale=np.ones((n, 3), F32), rest_rotation=np.tile(np.eye(4, dtype=F32).reshape(1, 16), (n, 1)))


def chain_animation(skeleton: "SkeletonDesc", length: float = 2.0, n_keys: int = 9, amplitude: float = 0.4, seed: int = 11) -> "AnimationDesc":
    """A swaying loop for chain_skeleton: every bone has a rotation track (quaternions about Z, slerped), the root bone also a
    position track (bobbing), one bone a stepped scale track and one bone no channel at all (it keeps its rest transform)."""
    rng = np.random.default_rng(seed)
    times = np.linspace(0.0, length, n_keys).astype(F32)
    channels = {}
    n = len(skeleton.names)
    for b, name in enumerate(skeleton.names):
        if b == n - 2 and n > 3:
            continue
        ang = amplitude * np.sin(2 * np.pi * times.astype(np.float64) / length + 0.5 * b)
        axis_tilt = 0.15 * rng.standard_normal(3)
        axis = np.array([0.0, 0.0, 1.0]) + axis_tilt
        axis /= np.linalg.norm(axis)
        q = np.concatenate([axis[None, :] * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1).astype(F32)
        tracks = {TRACK_ROTATION: TrackDesc(times, q)}
        if b == 0:
            p = np.zeros((n_keys, 3), F32)
            p[:, 1] = (0.1 * np.sin(4 * np.pi * times.astype(np.float64) / length)).astype(F32)
            tracks[TRACK_POSITION] = TrackDesc(times, p)
        if b == 2:
            sc = np.ones((3, 3), F32); sc[1] = (1.1, 0.95, 1.1)
            tracks[TRACK_SCALE] = TrackDesc(np.array([0.0, length / 2, length], F32), sc, interpolation=INTERPOLATION_CONSTANT)
        channels[name] = tracks
    return AnimationDesc("sway", length, channels)


def skinned_crowd_animated_scene(n: int = 10000, seed: int = 3) -> SceneDesc:
    """configs[2] with the bones posed on the device: the same characters, one shared skeleton and animation, every character at
    its own playhead -- a frame ships one 80-byte pose per character instead of 8 matrices."""
    rng = np.random.default_rng(seed)
    mesh, bone_y = tube_character_mesh()
    sk = chain_skeleton(bone_y)
    anim = chain_animation(sk)
    side = int(math.ceil(math.sqrt(n)))
    playheads = rng.uniform(0, anim.length, n).astype(F32)
    groups = []
    for g in range(8):
        ls = []
        for k in range(2):
            p = rng.uniform((-side, 1.0, -side), (side, 4.0, side))
            c = rng.uniform(0.3, 1.0, 3)
            ls.append(LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in p)), color=tuple(float(x) for x in c), energy=3.0,
                                range=12.0 + side, in_scene=False, name="g%dp%d" % (g, k)))
        groups.append(LightGroupDesc(ls))
    models = []
    for i in range(n):
        gx, gz = i % side, i // side
        root = NodeDesc(position=(2.0 * (gx - side / 2), 0.0, 2.0 * (gz - side / 2)))
        models.append(ModelDesc(mesh, NodeDesc(), pose=PoseDesc(sk, anim, float(playheads[i]), root), light_group=groups[i % 8],
                                frustum_culling=False, name="char%d" % i))
    cam = CameraDesc(node=NodeDesc(position=(0.0, 12.0, float(side) + 20.0), rotation=rotation_matrix((1, 0, 0), math.radians(-25))), far=400.0)
    return SceneDesc(models, [], cam, WorldDesc(ambient_energy=0.2), name="skinned_crowd_animated")


# ---- lights ----------------------------------------------------------------------------------------------
def standard_lights(seed: int = 2, box_min=(-50, 2, -25), box_max=(50, 6, 25), n_point: int = 4, point_range: float = 0.0,
                    energy: float = 2.0) -> List[LightDesc]:
    rng = np.random.default_rng(seed)
    lights = []
    for i in range(n_point):
        p = rng.uniform(box_min, box_max).astype(F32)
        c = rng.uniform(0.3, 1.0, 3).astype(F32)
        lights.append(LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in p)), color=tuple(float(x) for x in c),
                                energy=energy, range=point_range, name="point%d" % i))
    sun_rot = (rotation_matrix((1, 0, 0), math.radians(45)).astype(np.float64) @ rotation_matrix((0, 1, 0), math.radians(30)).astype(np.float64)).astype(F32)
    lights.append(LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=sun_rot), color=(1.0, 0.95, 0.9), energy=1.0, name="sun"))
    return lights


# ---- BASELINE.json configs ------------------------------------------------------------------------------------
def stress_scene(lit: bool = False) -> SceneDesc:
    """configs[0]: examples/stress/main.go:46-103 -- 25 models sharing a 1000-cube merged mesh (600 000 v / 300 000 t).
    lit=False is the faithful shadeless scene; lit=True is variant 1b of SURVEY.md 8d."""
    mat = MaterialDesc(shadeless=not lit, use_texture=True, tex_w=16, tex_h=16)
    mesh = cube_mesh(2, 2, 2, copies=1000, material=mat)
    models = [ModelDesc(mesh, NodeDesc(position=(4.0 * x, 0.0, 0.0)), name="merged%d" % x) for x in range(25)]
    if lit:
        cam = CameraDesc(node=NodeDesc(position=(48.0, 0.0, 120.0)), far=250.0)
        lights = standard_lights(seed=2, box_min=(0, 2, -5), box_max=(96, 6, 5))
    else:
        cam = CameraDesc()
        lights = []
    return SceneDesc(models, lights, cam, WorldDesc(), name="stress")


def million_triangle_scene(nx: int = 1000, nz: int = 500, rows_per_part: int = 10) -> SceneDesc:
    """configs[1]: 1M-triangle static mesh in 50 MeshParts, 4 point lights + 1 directional, backface cull + depth sort."""
    mesh = grid_mesh(nx, nz, rows_per_part=rows_per_part, seed=1, material=MaterialDesc(backface_culling=True, triangle_sort_mode=0))
    pitch = rotation_matrix((1, 0, 0), math.radians(-20))
    cam = CameraDesc(node=NodeDesc(position=(0.0, 15.0, 40.0), rotation=pitch), far=200.0)
    sx = 50.0 * nx / 1000.0; sz = 25.0 * nz / 500.0
    return SceneDesc([ModelDesc(mesh, NodeDesc(), name="terrain")], standard_lights(seed=2, box_min=(-sx, 2, -sz), box_max=(sx, 6, sz)), cam,
                     WorldDesc(), name="million_triangles")


def skinned_crowd_scene(n: int = 10000, t: float = 0.0, seed: int = 3) -> SceneDesc:
    """configs[2]: n skinned characters (425 v / 768 t / 8 bones) on a grid, 8 LightGroups of 2 ranged point lights."""
    rng = np.random.default_rng(seed)
    mesh, bone_y = tube_character_mesh()
    side = int(math.ceil(math.sqrt(n)))
    phases = rng.uniform(0, 2 * np.pi, n)
    groups = []
    for g in range(8):
        ls = []
        for k in range(2):
            p = rng.uniform((-side, 1.0, -side), (side, 4.0, side))
            c = rng.uniform(0.3, 1.0, 3)
            ls.append(LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in p)), color=tuple(float(x) for x in c), energy=3.0,
                                range=12.0 + side, in_scene=False, name="g%dp%d" % (g, k)))
        groups.append(LightGroupDesc(ls))
    models = []
    for i in range(n):
        gx, gz = i % side, i // side
        base = (2.0 * (gx - side / 2), 0.0, 2.0 * (gz - side / 2))
        bi = chain_bone_influences(bone_y, base, float(phases[i]), t)
        models.append(ModelDesc(mesh, NodeDesc(), bone_influences=bi, light_group=groups[i % 8], frustum_culling=False, name="char%d" % i))
    cam = CameraDesc(node=NodeDesc(position=(0.0, 12.0, float(side) + 20.0), rotation=rotation_matrix((1, 0, 0), math.radians(-25))), far=400.0)
    return SceneDesc(models, [], cam, WorldDesc(ambient_energy=0.2), name="skinned_crowd")


def particles_scene(n: int = 2000000, seed: int = 4, batch_cap: int = 10922) -> SceneDesc:
    """configs[3]: n billboard particle quads with per-vertex colour, transparent pass, dynamic batches of <= batch_cap."""
    rng = np.random.default_rng(seed)
    colors = rng.uniform(0.1, 1.0, (4, 4)).astype(F32)
    mat = MaterialDesc(billboard_enabled=True, transparency_mode=3, shadeless=True, backface_culling=False)
    mesh = plane_mesh(2, 2, material=mat, colors=colors)
    owner_mesh = plane_mesh(2, 2, material=MaterialDesc(transparency_mode=3, shadeless=True, backface_culling=False))
    rng2 = np.random.default_rng(seed + 1)
    pos = rng2.uniform((-30, -30, -60), (30, 30, 0), (n, 3)).astype(F32)
    scl = rng2.uniform(0.05, 0.4, n).astype(F32)
    models = []
    n_owner = (n + batch_cap - 1) // batch_cap
    owners = [ModelDesc(owner_mesh, NodeDesc(position=(0.0, 0.0, -30.0)), frustum_culling=False, name="batch%d" % b) for b in range(n_owner)]
    models.extend(owners)
    for i in range(n):
        c = 0.5 + 0.5 * math.sin(0.001 * i)
        models.append(ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in pos[i]), scale=(float(scl[i]),) * 3), color=(1.0, c, 1.0 - c, 0.8),
                                frustum_culling=False, batch_owner=owners[i // batch_cap], batch_owner_part=0, name="p%d" % i))
    return SceneDesc(models, [], CameraDesc(far=150.0), WorldDesc(), name="particles")


def instanced_scene(n: int = 19532, seed: int = 6, subdivisions: int = 4, box=(400.0, 60.0, 400.0)) -> SceneDesc:
    """configs[4]: n instances of a 5120-triangle icosphere (19 532 -> 100.0 M triangles), 4 point lights + sun."""
    rng = np.random.default_rng(seed)
    mesh = icosphere_mesh(subdivisions)
    models = []
    for i in range(n):
        p = rng.uniform((-box[0] / 2, -box[1] / 2, -box[2] / 2), (box[0] / 2, box[1] / 2, box[2] / 2))
        axis = rng.normal(size=3); ang = rng.uniform(0, 2 * np.pi); s = float(rng.uniform(0.5, 1.5))
        models.append(ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in p), scale=(s, s, s), rotation=rotation_matrix(axis, ang)), name="s%d" % i))
    cam = CameraDesc(node=NodeDesc(position=(0.0, 40.0, box[2] / 2 + 60.0), rotation=rotation_matrix((1, 0, 0), math.radians(-15))), far=800.0)
    lights = standard_lights(seed=7, box_min=(-box[0] / 2, 10, -box[2] / 2), box_max=(box[0] / 2, 40, box[2] / 2), energy=40.0)
    return SceneDesc(models, lights, cam, WorldDesc(), name="instanced")


WORKLOADS = {
    "stress": lambda: stress_scene(lit=False),
    "stress_lit": lambda: stress_scene(lit=True),
    "million_triangles": million_triangle_scene,
    "skinned_crowd": skinned_crowd_scene,
    "particles": particles_scene,
    "instanced": instanced_scene,
    # reduced variants for exploration (not BASELINE configs): one GPU's share of configs[4] at 8 GPUs, and a
    # tenth of configs[3]
    "million_one_part": lambda: million_triangle_scene(rows_per_part=500),      # the same mesh as ONE 1 M-triangle MeshPart
    "instanced_shard": lambda: instanced_scene(n=2442),
    "particles_200k": lambda: particles_scene(n=200000),
    "skinned_crowd_1k": lambda: skinned_crowd_scene(n=1000),
    "skinned_crowd_animated": skinned_crowd_animated_scene,
    "skinned_crowd_animated_1k": lambda: skinned_crowd_animated_scene(n=1000),
}
