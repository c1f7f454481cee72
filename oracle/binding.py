"""ctypes binding of the CPU oracle (oracle/libt3doracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (tetra3d_b200) never does.  See oracle/t3d_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from tetra3d_b200.scenedesc import (LIGHT_VOLUME, AnimationDesc, MaterialDesc, MeshDesc, NodeDesc, SceneDesc, SkeletonDesc)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libt3doracle.so")

f32p = C.POINTER(C.c_float)
u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


class orc_node(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("scale", C.c_float * 3), ("rotation", C.c_float * 16), ("parent", C.c_int32)]


class orc_material(C.Structure):
    _fields_ = [("present", C.c_int32), ("color", C.c_float * 4), ("backface_culling", C.c_int32),
                ("triangle_sort_mode", C.c_int32), ("shadeless", C.c_int32), ("lighting_mode", C.c_int32),
                ("light_volume_shading_mode", C.c_int32), ("billboard_enabled", C.c_int32),
                ("billboard_lock_x", C.c_int32), ("billboard_lock_y", C.c_int32), ("billboard_lock_z", C.c_int32),
                ("billboard_up_direction", C.c_int32), ("billboarded_depth_mode", C.c_int32),
                ("transparency_mode", C.c_int32), ("blend_source_over", C.c_int32), ("visible", C.c_int32),
                ("render_order", C.c_int32), ("use_texture", C.c_int32), ("tex_w", C.c_float), ("tex_h", C.c_float)]


class orc_part(C.Structure):
    _fields_ = [("vertex_index_start", C.c_int32), ("vertex_index_end", C.c_int32), ("triangle_start", C.c_int32),
                ("triangle_end", C.c_int32), ("material", orc_material)]


class orc_mesh(C.Structure):
    _fields_ = [("n_verts", C.c_int32), ("positions", f32p), ("normals", f32p), ("uvs", f32p), ("colors", f32p),
                ("active_color_channel", C.c_int32), ("weights", f32p), ("bones", u16p), ("n_influences", u8p),
                ("n_tris", C.c_int32), ("indices", u32p), ("n_parts", C.c_int32), ("parts", C.POINTER(orc_part)),
                ("n_shape_keys", C.c_int32), ("key_positions", f32p), ("key_normals", f32p), ("key_weights", f32p),
                ("auto_subdivide", C.c_int32), ("tri_centers", f32p), ("tri_sub_base", i32p), ("tri_sub_level", C.POINTER(C.c_int8)),
                ("tri_n_levels", u8p), ("tri_max_span", f32p), ("tri_normal", f32p)]


class orc_model(C.Structure):
    _fields_ = [("node", C.c_int32), ("mesh", C.c_int32), ("color", C.c_float * 4), ("shadeless", C.c_int32),
                ("visible", C.c_int32), ("frustum_culling", C.c_int32), ("skinned", C.c_int32), ("n_bones", C.c_int32),
                ("bone_influences", f32p), ("light_group", C.c_int32), ("batch_owner", C.c_int32),
                ("batch_owner_part", C.c_int32)]


class orc_light(C.Structure):
    _fields_ = [("type", C.c_int32), ("node", C.c_int32), ("color", C.c_float * 3), ("energy", C.c_float),
                ("visible", C.c_int32), ("range", C.c_float), ("volume", C.c_int32)]


class orc_light_group(C.Structure):
    _fields_ = [("n_lights", C.c_int32), ("lights", i32p), ("active", C.c_int32)]


class orc_light_volume(C.Structure):
    _fields_ = [("dim_min", C.c_float * 3), ("dim_max", C.c_float * 3), ("cell_size", C.c_float * 3),
                ("cell_count", C.c_int32 * 3), ("offset", C.c_float * 3), ("light_bias", C.c_float * 3),
                ("cell_colors", f32p), ("cell_light_start", i32p), ("cell_light_index", i32p)]


class orc_camera(C.Structure):
    _fields_ = [("node", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("field_of_view", C.c_float),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("perspective", C.c_int32),
                ("ortho_scale", C.c_float), ("depth_margin", C.c_float), ("vertex_snapping", C.c_float),
                ("render_depth", C.c_int32), ("render_normals", C.c_int32),
                ("perspective_corrected_texture_mapping", C.c_int32), ("max_light_count", C.c_int32),
                ("max_light_focus_node", C.c_int32)]


class orc_world(C.Structure):
    _fields_ = [("present", C.c_int32), ("lighting_on", C.c_int32), ("fog_on", C.c_int32), ("fog_mode", C.c_int32),
                ("fog_color", C.c_float * 4), ("fog_range", C.c_float * 2), ("ambient_color", C.c_float * 3),
                ("ambient_energy", C.c_float), ("ambient_visible", C.c_int32)]


class orc_scene(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("nodes", C.POINTER(orc_node)), ("n_meshes", C.c_int32), ("meshes", C.POINTER(orc_mesh)),
                ("n_models", C.c_int32), ("models", C.POINTER(orc_model)), ("n_lights", C.c_int32), ("lights", C.POINTER(orc_light)),
                ("n_scene_lights", C.c_int32), ("scene_lights", i32p), ("n_light_groups", C.c_int32),
                ("light_groups", C.POINTER(orc_light_group)), ("n_volumes", C.c_int32), ("volumes", C.POINTER(orc_light_volume)),
                ("n_subdivision_levels", C.c_int32), ("subdivision_levels", f32p),
                ("camera", orc_camera), ("world", orc_world)]


class orc_render_rec(C.Structure):
    _fields_ = [("model", C.c_int32), ("part", C.c_int32), ("flush", C.c_int32), ("first_vertex", C.c_int64),
                ("vertex_count", C.c_int32), ("total_tris", C.c_int32), ("min_depth", C.c_float), ("max_depth", C.c_float)]


class orc_flush_rec(C.Structure):
    _fields_ = [("model", C.c_int32), ("part", C.c_int32), ("first_vertex", C.c_int64), ("vertex_count", C.c_int32)]


class orc_result(C.Structure):
    _fields_ = [("n_renders", C.c_int32), ("renders", C.POINTER(orc_render_rec)), ("n_flushes", C.c_int32),
                ("flushes", C.POINTER(orc_flush_rec)), ("n_vertices", C.c_int64), ("color_list", f32p), ("depth_list", f32p),
                ("normal_list", f32p), ("index_list", u16p), ("tri_ids", u32p), ("n_captured", C.c_int64),
                ("cap_clip", f32p), ("cap_light", f32p), ("total_parts", C.c_int32), ("drawn_parts", C.c_int32),
                ("batched_parts", C.c_int32), ("total_tris", C.c_int32), ("drawn_tris", C.c_int32),
                ("light_count", C.c_int32), ("active_light_count", C.c_int32), ("t_total", C.c_double),
                ("t_process", C.c_double), ("t_light", C.c_double), ("t_emit", C.c_double)]


class orc_track(C.Structure):
    _fields_ = [("first_key", C.c_int32), ("n_keys", C.c_int32), ("interpolation", C.c_int32)]


class orc_skeleton(C.Structure):
    _fields_ = [("n_bones", C.c_int32), ("parent", i32p), ("inverse_bind", f32p), ("rest_position", f32p),
                ("rest_scale", f32p), ("rest_rotation", f32p)]


class orc_animation(C.Structure):
    _fields_ = [("n_bones", C.c_int32), ("tracks", C.POINTER(orc_track)), ("key_times", f32p), ("key_values", f32p)]


_lib = None


def build(force: bool = False) -> str:
    """Compiles oracle/libt3doracle.so with the committed Makefile (gcc -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "t3d_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_render_scene.argtypes = [C.POINTER(orc_scene), C.c_int, C.POINTER(C.POINTER(orc_result))]
        L.orc_render_scene.restype = C.c_int
        L.orc_free_result.argtypes = [C.POINTER(orc_result)]
        L.orc_bake_ao.argtypes = [C.POINTER(orc_mesh), f32p, f32p, u8p, u8p, C.c_float, f32p, f32p]
        L.orc_bake_ao.restype = C.c_int
        L.orc_go_acos.argtypes = [C.c_double]; L.orc_go_acos.restype = C.c_double
        L.orc_bake_lighting.argtypes = [C.POINTER(orc_scene), C.c_int32, i32p, C.c_int32, f32p]
        L.orc_bake_lighting.restype = C.c_int
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_mat4_mult",):
            getattr(L, name).argtypes = [f32p, f32p, f32p]
        for name in ("orc_mat4_inverted", "orc_mat4_transposed"):
            getattr(L, name).argtypes = [f32p, f32p]
        L.orc_mat4_identity.argtypes = [f32p]
        L.orc_mat4_rotate.argtypes = [C.c_float] * 4 + [f32p]
        L.orc_mat4_translate.argtypes = [C.c_float] * 3 + [f32p]
        L.orc_mat4_scale.argtypes = [C.c_float] * 3 + [f32p]
        L.orc_mat4_equals.argtypes = [f32p, f32p]
        L.orc_mat4_equals.restype = C.c_int
        L.orc_mat4_decompose.argtypes = [f32p, f32p, f32p, f32p]
        L.orc_mat4_lookat.argtypes = [f32p, f32p, f32p, f32p]
        L.orc_mat4_perspective.argtypes = [C.c_float] * 5 + [f32p]
        L.orc_mat4_orthographic.argtypes = [C.c_float] * 6 + [f32p]
        L.orc_triangle_centers.argtypes = [f32p, u32p, C.c_int32, f32p]
        L.orc_bin_index.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_bin_index.restype = C.c_int
        L.orc_node_transform.argtypes = [C.POINTER(orc_node), C.c_int32, f32p]
        L.orc_pose_bones.argtypes = [C.POINTER(orc_skeleton), C.POINTER(orc_animation), C.c_float, f32p, f32p]
        L.orc_pose_bones_blended.argtypes = [C.POINTER(orc_skeleton), C.POINTER(orc_animation), C.c_float, C.POINTER(orc_animation), C.c_float,
                                             C.c_float, f32p, f32p]
        _anp = C.POINTER(C.POINTER(orc_animation))
        L.orc_pose_bones_chain.argtypes = [C.POINTER(orc_skeleton), C.c_int, _anp, f32p, f32p, f32p, f32p]
        L.orc_pose_locals_chain.argtypes = [C.POINTER(orc_skeleton), C.c_int, _anp, f32p, f32p, f32p, f32p, f32p, C.POINTER(C.c_int32)]
        L.orc_sample_bone.argtypes = [C.POINTER(orc_animation), C.c_int, C.c_float, C.POINTER(C.c_int32), f32p, f32p, f32p]
        L.orc_quat_slerp.argtypes = [f32p, f32p, C.c_float, f32p]
        L.orc_quat_to_matrix.argtypes = [f32p, f32p]
        L.orc_go_sin.argtypes = [C.c_double]; L.orc_go_sin.restype = C.c_double
        L.orc_go_atan2.argtypes = [C.c_double, C.c_double]; L.orc_go_atan2.restype = C.c_double
        _lib = L
    return _lib


def _fp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(f32p) if a is not None else None


# ---- math helpers for known-answer tests -------------------------------------------------------------
def _m(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).reshape(16)


def mat4_mult(a, b) -> np.ndarray:
    out = np.zeros(16, np.float32); A = _m(a); B = _m(b)
    lib().orc_mat4_mult(_fp(A), _fp(B), _fp(out)); return out.reshape(4, 4)


def mat4_inverted(a) -> np.ndarray:
    out = np.zeros(16, np.float32); A = _m(a)
    lib().orc_mat4_inverted(_fp(A), _fp(out)); return out.reshape(4, 4)


def mat4_transposed(a) -> np.ndarray:
    out = np.zeros(16, np.float32); A = _m(a)
    lib().orc_mat4_transposed(_fp(A), _fp(out)); return out.reshape(4, 4)


def mat4_identity() -> np.ndarray:
    out = np.zeros(16, np.float32); lib().orc_mat4_identity(_fp(out)); return out.reshape(4, 4)


def mat4_rotate(x, y, z, angle) -> np.ndarray:
    out = np.zeros(16, np.float32); lib().orc_mat4_rotate(x, y, z, angle, _fp(out)); return out.reshape(4, 4)


def mat4_translate(x, y, z) -> np.ndarray:
    out = np.zeros(16, np.float32); lib().orc_mat4_translate(x, y, z, _fp(out)); return out.reshape(4, 4)


def mat4_scale(x, y, z) -> np.ndarray:
    out = np.zeros(16, np.float32); lib().orc_mat4_scale(x, y, z, _fp(out)); return out.reshape(4, 4)


def mat4_equals(a, b) -> bool:
    A = _m(a); B = _m(b)
    return bool(lib().orc_mat4_equals(_fp(A), _fp(B)))


def mat4_decompose(a):
    A = _m(a); p = np.zeros(3, np.float32); s = np.zeros(3, np.float32); r = np.zeros(16, np.float32)
    lib().orc_mat4_decompose(_fp(A), _fp(p), _fp(s), _fp(r)); return p, s, r.reshape(4, 4)


def mat4_lookat(frm, to, up) -> np.ndarray:
    f = np.asarray(frm, np.float32); t = np.asarray(to, np.float32); u = np.asarray(up, np.float32); out = np.zeros(16, np.float32)
    lib().orc_mat4_lookat(_fp(f), _fp(t), _fp(u), _fp(out)); return out.reshape(4, 4)


def mat4_perspective(fovy, near, far, w, h) -> np.ndarray:
    out = np.zeros(16, np.float32); lib().orc_mat4_perspective(fovy, near, far, w, h, _fp(out)); return out.reshape(4, 4)


def triangle_centers(positions: np.ndarray, indices: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(positions, np.float32); i = np.ascontiguousarray(indices, np.uint32)
    out = np.zeros((len(i.reshape(-1, 3)), 3), np.float32)
    lib().orc_triangle_centers(_fp(p), i.ctypes.data_as(u32p), len(out), _fp(out)); return out


# ---- bone poses (animation.go, quaternion.go, node.go:572-594) -------------------------------------------
def go_sin(x: float) -> float:
    return float(lib().orc_go_sin(float(x)))


def go_atan2(y: float, x: float) -> float:
    return float(lib().orc_go_atan2(float(y), float(x)))


def quat_slerp(a, b, percent: float) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32); out = np.zeros(4, np.float32)
    lib().orc_quat_slerp(_fp(a), _fp(b), float(percent), _fp(out))
    return out


def quat_to_matrix(q) -> np.ndarray:
    q = np.ascontiguousarray(q, np.float32); out = np.zeros(16, np.float32)
    lib().orc_quat_to_matrix(_fp(q), _fp(out))
    return out.reshape(4, 4)


def node_world_transform(nd: NodeDesc) -> np.ndarray:
    """Node.Transform() (node.go:572-608) of a NodeDesc with its chain of parents."""
    chain = []
    while nd is not None:
        chain.append(nd); nd = nd.parent
    arr = (orc_node * len(chain))()
    for i, n in enumerate(chain):
        arr[i].position[:] = [float(x) for x in n.position]; arr[i].scale[:] = [float(x) for x in n.scale]
        arr[i].rotation[:] = [float(x) for x in np.asarray(n.rotation, np.float32).reshape(16)]
        arr[i].parent = i + 1 if i + 1 < len(chain) else -1
    out = np.zeros(16, np.float32)
    lib().orc_node_transform(arr, 0, _fp(out))
    return out


def flatten_animation(anim: AnimationDesc, sk: SkeletonDesc):
    """Per bone of `sk` the three tracks of the channel carrying the bone's name (assignChannels, animation.go:618-672)."""
    n = len(sk.names)
    tracks = np.zeros((n, 3, 3), np.int32)
    times, values = [], []
    k = 0
    for b, name in enumerate(sk.names):
        ch = anim.channels.get(name, {})
        for t in range(3):
            td = ch.get(t)
            tracks[b, t, 0] = k
            if td is None:
                continue
            tracks[b, t, 1] = len(td.times); tracks[b, t, 2] = int(td.interpolation)
            times.append(td.times); values.append(td.values); k += len(td.times)
    kt = np.concatenate(times).astype(np.float32) if times else np.zeros(1, np.float32)
    kv = np.concatenate(values).astype(np.float32).reshape(-1, 4) if values else np.zeros((1, 4), np.float32)
    return tracks, np.ascontiguousarray(kt), np.ascontiguousarray(kv)


def _c_skeleton(sk: SkeletonDesc):
    cs = orc_skeleton()
    cs.n_bones = len(sk.parent); cs.parent = sk.parent.ctypes.data_as(i32p); cs.inverse_bind = _fp(sk.inverse_bind)
    cs.rest_position = _fp(sk.rest_position); cs.rest_scale = _fp(sk.rest_scale); cs.rest_rotation = _fp(sk.rest_rotation)
    return cs


def _c_animation(anim: Optional[AnimationDesc], sk: SkeletonDesc):
    """(orc_animation or None, the arrays it points into)."""
    if anim is None:
        return None, None
    tracks, kt, kv = flatten_animation(anim, sk)
    ct = (orc_track * (3 * len(sk.parent)))()
    for i, (f, n, ip) in enumerate(tracks.reshape(-1, 3)):
        ct[i].first_key = int(f); ct[i].n_keys = int(n); ct[i].interpolation = int(ip)
    ca = orc_animation(); ca.n_bones = len(sk.parent); ca.tracks = ct; ca.key_times = _fp(kt); ca.key_values = _fp(kv)
    return ca, (ct, kt, kv)


def pose_bones(sk: SkeletonDesc, anim: Optional[AnimationDesc], playhead: float, root16) -> np.ndarray:
    """boneInfluence of every bone of an armature posed at `playhead`: (n_bones, 16)."""
    L = lib()
    cs = _c_skeleton(sk)
    ca, _keep = _c_animation(anim, sk)
    root = np.ascontiguousarray(root16, np.float32).reshape(16)
    out = np.zeros((len(sk.parent), 16), np.float32)
    L.orc_pose_bones(C.byref(cs), C.byref(ca) if ca is not None else None, float(playhead), _fp(root), _fp(out))
    return out


def pose_bones_blended(sk: SkeletonDesc, anim_from: Optional[AnimationDesc], from_playhead: float, anim: Optional[AnimationDesc],
                       playhead: float, blend: float, root16) -> np.ndarray:
    """The same while the AnimationPlayer blends from `anim_from` (frozen at from_playhead) into `anim` (animation.go:852-907)."""
    L = lib()
    cs = _c_skeleton(sk)
    cf, _keep_f = _c_animation(anim_from, sk)
    ca, _keep_a = _c_animation(anim, sk)
    root = np.ascontiguousarray(root16, np.float32).reshape(16)
    out = np.zeros((len(sk.parent), 16), np.float32)
    L.orc_pose_bones_blended(C.byref(cs), C.byref(cf) if cf is not None else None, float(from_playhead),
                             C.byref(ca) if ca is not None else None, float(playhead), float(blend), _fp(root), _fp(out))
    return out


def _c_chain(sk: SkeletonDesc, anims, playheads, blends):
    keep, ptrs = [], (C.POINTER(orc_animation) * len(anims))()
    for i, a in enumerate(anims):
        ca, k = _c_animation(a, sk)
        keep.append((ca, k))
        ptrs[i] = C.pointer(ca) if ca is not None else None
    ph = np.ascontiguousarray(playheads, np.float32); bp = np.ascontiguousarray(blends, np.float32)
    assert len(ph) == len(anims) == len(bp)
    return ptrs, ph, bp, keep


def pose_bones_chain(sk: SkeletonDesc, anims, playheads, blends, root16) -> np.ndarray:
    """boneInfluence of every bone for a player whose blends were interrupted by further Play() calls: `anims` in the order they
    were played (the last one playing), playheads[i] / blends[i] the values of level i's last update (blends[0] unused)."""
    cs = _c_skeleton(sk)
    ptrs, ph, bp, _keep = _c_chain(sk, anims, playheads, blends)
    root = np.ascontiguousarray(root16, np.float32).reshape(16)
    out = np.zeros((len(sk.parent), 16), np.float32)
    lib().orc_pose_bones_chain(C.byref(cs), len(anims), ptrs, _fp(ph), _fp(bp), _fp(root), _fp(out))
    return out


def pose_locals_chain(sk: SkeletonDesc, anims, playheads, blends):
    """The bones' local position (n, 3), scale (n, 3) and rotation (n, 16) after the same fold, and which of the three a level of
    the chain set, (n, 3) bool (the others are the skeleton's rest values)."""
    cs = _c_skeleton(sk)
    ptrs, ph, bp, _keep = _c_chain(sk, anims, playheads, blends)
    n = len(sk.parent)
    pos, scl, rot = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 16), np.float32)
    was_set = np.zeros((n, 3), np.int32)
    lib().orc_pose_locals_chain(C.byref(cs), len(anims), ptrs, _fp(ph), _fp(bp), _fp(pos), _fp(scl), _fp(rot), was_set.ctypes.data_as(C.POINTER(C.c_int32)))
    return pos, scl, rot, was_set.astype(bool)


def sample_bone(sk: SkeletonDesc, anim: AnimationDesc, bone: int, time: float):
    """(flags[3], position[3], scale[3], quaternion[4]) of one bone's channel sampled at `time` (animation.go:674-722)."""
    ca, _keep = _c_animation(anim, sk)
    flags = np.zeros(3, np.int32); pos = np.zeros(3, np.float32); scl = np.zeros(3, np.float32); q = np.zeros(4, np.float32)
    lib().orc_sample_bone(C.byref(ca), int(bone), float(time), flags.ctypes.data_as(C.POINTER(C.c_int32)), _fp(pos), _fp(scl), _fp(q))
    return flags.astype(bool), pos, scl, q


def bin_index(depth, mn, mx, sort_mode) -> int:
    return int(lib().orc_bin_index(np.float32(depth), np.float32(mn), np.float32(mx), int(sort_mode)))


# ---- scene marshalling ---------------------------------------------------------------------------------
def _fill_material(dst: orc_material, md: Optional[MaterialDesc]):
    if md is None:
        dst.present = 0
        return
    dst.present = 1
    dst.color[:] = [float(x) for x in md.color]
    dst.backface_culling = int(md.backface_culling)
    dst.triangle_sort_mode = int(md.triangle_sort_mode)
    dst.shadeless = int(md.shadeless)
    dst.lighting_mode = int(md.lighting_mode)
    dst.light_volume_shading_mode = int(md.light_volume_shading_mode)
    dst.billboard_enabled = int(md.billboard_enabled)
    dst.billboard_lock_x = int(md.billboard_lock_x)
    dst.billboard_lock_y = int(md.billboard_lock_y)
    dst.billboard_lock_z = int(md.billboard_lock_z)
    dst.billboard_up_direction = int(md.billboard_up_direction)
    dst.billboarded_depth_mode = int(md.billboarded_depth_mode)
    dst.transparency_mode = int(md.transparency_mode)
    dst.blend_source_over = int(md.blend_source_over)
    dst.visible = int(md.visible)
    dst.render_order = int(md.render_order)
    dst.use_texture = int(md.use_texture and md.tex_w > 0)
    dst.tex_w = float(md.tex_w)
    dst.tex_h = float(md.tex_h)


@dataclass
class OracleResult:
    color_list: np.ndarray      # (n, 12) float32
    depth_list: np.ndarray
    normal_list: np.ndarray
    index_list: np.ndarray      # (n,) uint16
    tri_ids: np.ndarray         # (n/3,) uint32
    renders: List[dict]
    flushes: List[dict]
    cap_clip: Optional[np.ndarray]
    cap_light: Optional[np.ndarray]
    debug: Dict[str, int]
    times: Dict[str, float]

    @property
    def n_vertices(self) -> int:
        return len(self.color_list)


class _Marshalled:
    """Keeps every numpy array / ctypes array referenced by the orc_scene alive."""

    def __init__(self, desc: SceneDesc):
        self.keep = []
        nodes: List[NodeDesc] = []
        node_index: Dict[int, int] = {}

        def node_of(nd: NodeDesc) -> int:
            if id(nd) in node_index:
                return node_index[id(nd)]
            parent = node_of(nd.parent) if nd.parent is not None else -1
            node_index[id(nd)] = len(nodes)
            nodes.append((nd, parent))
            return node_index[id(nd)]

        meshes: List[MeshDesc] = []
        mesh_index: Dict[int, int] = {}
        for m in desc.models:
            if id(m.mesh) not in mesh_index:
                mesh_index[id(m.mesh)] = len(meshes)
                meshes.append(m.mesh)

        c_meshes = (orc_mesh * max(len(meshes), 1))()
        for i, md in enumerate(meshes):
            cm = c_meshes[i]
            cm.n_verts = md.n_verts
            cm.positions = _fp(md.positions); cm.normals = _fp(md.normals); cm.uvs = _fp(md.uvs)
            cm.colors = _fp(md.colors) if (md.colors is not None and md.active_color_channel >= 0) else None
            cm.active_color_channel = int(md.active_color_channel)
            if md.weights is not None:
                cm.weights = _fp(md.weights)
                cm.bones = md.bones.ctypes.data_as(u16p)
                cm.n_influences = md.n_influences.ctypes.data_as(u8p)
            cm.n_tris = md.n_tris
            cm.indices = md.indices.ctypes.data_as(u32p)
            parts = (orc_part * max(len(md.parts), 1))()
            for j, p in enumerate(md.parts):
                parts[j].vertex_index_start = int(p.vertex_start); parts[j].vertex_index_end = int(p.vertex_end)
                parts[j].triangle_start = int(p.tri_start); parts[j].triangle_end = int(p.tri_end)
                _fill_material(parts[j].material, p.material)
            self.keep.append(parts)
            cm.n_parts = len(md.parts); cm.parts = parts
            if md.auto_subdivide:
                cm.auto_subdivide = 1
                cm.tri_centers = _fp(md.tri_centers); cm.tri_sub_base = md.tri_sub_base.ctypes.data_as(i32p)
                cm.tri_sub_level = md.tri_sub_level.ctypes.data_as(C.POINTER(C.c_int8)); cm.tri_n_levels = md.tri_n_levels.ctypes.data_as(u8p)
                cm.tri_max_span = _fp(md.tri_max_span); cm.tri_normal = _fp(md.tri_normal)
            if md.key_weights is not None:
                cm.n_shape_keys = len(md.key_weights)
                cm.key_positions = _fp(md.key_positions); cm.key_normals = _fp(md.key_normals); cm.key_weights = _fp(md.key_weights)
        self.keep.append(c_meshes)

        # lights: scene lights first (in order), then group-only lights
        all_lights = list(desc.lights)
        for m in desc.models:
            if m.light_group is not None:
                for l in m.light_group.lights:
                    if not any(l is x for x in all_lights):
                        all_lights.append(l)
        for l in list(all_lights):          # lights that only LightVolume cells use (LightVolumeCell.SetUseLights)
            if l.type == LIGHT_VOLUME and l.volume is not None and l.volume.cell_lights:
                for key in sorted(l.volume.cell_lights):
                    for nl in l.volume.cell_lights[key]:
                        if not any(nl is x for x in all_lights):
                            all_lights.append(nl)
        light_index = {id(l): i for i, l in enumerate(all_lights)}
        volumes = []
        c_lights = (orc_light * max(len(all_lights), 1))()
        for i, ld in enumerate(all_lights):
            cl = c_lights[i]
            cl.type = int(ld.type); cl.node = node_of(ld.node)
            cl.color[:] = [float(x) for x in ld.color]
            cl.energy = float(ld.energy); cl.visible = int(ld.visible); cl.range = float(ld.range)
            cl.volume = -1
            if ld.type == LIGHT_VOLUME:
                cl.volume = len(volumes); volumes.append(ld.volume)
        self.keep.append(c_lights)
        scene_light_ids = np.asarray([light_index[id(l)] for l in desc.lights if l.in_scene], dtype=np.int32)
        self.keep.append(scene_light_ids)

        c_volumes = (orc_light_volume * max(len(volumes), 1))()
        for i, v in enumerate(volumes):
            cv = c_volumes[i]
            cv.dim_min[:] = [float(x) for x in v.dim_min]; cv.dim_max[:] = [float(x) for x in v.dim_max]
            cv.cell_size[:] = [float(x) for x in v.cell_size]; cv.cell_count[:] = [int(x) for x in v.cell_count]
            cv.offset[:] = [float(x) for x in v.offset]; cv.light_bias[:] = [float(x) for x in v.light_bias]
            cols = np.ascontiguousarray(v.colors, np.float32); self.keep.append(cols)
            cv.cell_colors = _fp(cols)
            if v.cell_lights:
                cx, cy, cz = (int(x) for x in v.cell_count)
                start = np.zeros(cx * cy * cz + 1, np.int32); index = []
                per_cell = {int(x) + int(y) * cx + int(z) * cx * cy: ls for (x, y, z), ls in v.cell_lights.items()}
                for c in range(cx * cy * cz):
                    index.extend(light_index[id(nl)] for nl in per_cell.get(c, []))
                    start[c + 1] = len(index)
                index = np.asarray(index if index else [0], np.int32)
                self.keep.extend([start, index])
                cv.cell_light_start = start.ctypes.data_as(i32p); cv.cell_light_index = index.ctypes.data_as(i32p)
        self.keep.append(c_volumes)

        groups = []
        group_index: Dict[int, int] = {}
        for m in desc.models:
            if m.light_group is not None and id(m.light_group) not in group_index:
                group_index[id(m.light_group)] = len(groups); groups.append(m.light_group)
        c_groups = (orc_light_group * max(len(groups), 1))()
        for i, g in enumerate(groups):
            ids = np.asarray([light_index[id(l)] for l in g.lights], dtype=np.int32); self.keep.append(ids)
            c_groups[i].n_lights = len(ids); c_groups[i].lights = ids.ctypes.data_as(i32p); c_groups[i].active = int(g.active)
        self.keep.append(c_groups)

        model_index = {id(m): i for i, m in enumerate(desc.models)}
        c_models = (orc_model * max(len(desc.models), 1))()
        for i, md in enumerate(desc.models):
            cm = c_models[i]
            cm.node = node_of(md.node); cm.mesh = mesh_index[id(md.mesh)]
            cm.color[:] = [float(x) for x in md.color]
            cm.shadeless = int(md.shadeless); cm.visible = int(md.visible); cm.frustum_culling = int(md.frustum_culling)
            bi = md.bone_influences
            if md.pose is not None:     # bones posed from a skeleton + animation: the oracle's own restatement of that step
                if md.pose.blend_ms is not None:
                    bi = pose_bones_blended(md.pose.skeleton, md.pose.blend_from, md.pose.from_playhead, md.pose.animation, md.pose.playhead,
                                            md.pose.blend_fraction(), node_world_transform(md.pose.root))
                else:
                    bi = pose_bones(md.pose.skeleton, md.pose.animation, md.pose.playhead, node_world_transform(md.pose.root))
            if bi is not None:
                b = np.ascontiguousarray(bi, np.float32).reshape(-1, 16); self.keep.append(b)
                cm.skinned = 1; cm.n_bones = len(b); cm.bone_influences = _fp(b)
            cm.light_group = group_index[id(md.light_group)] if md.light_group is not None else -1
            cm.batch_owner = model_index[id(md.batch_owner)] if md.batch_owner is not None else -1
            cm.batch_owner_part = int(md.batch_owner_part)
        self.keep.append(c_models)

        cam_node = node_of(desc.camera.node)
        focus_node = node_of(desc.camera.max_light_focus) if desc.camera.max_light_focus is not None else -1

        c_nodes = (orc_node * max(len(nodes), 1))()
        for i, (nd, parent) in enumerate(nodes):
            c_nodes[i].position[:] = [float(x) for x in nd.position]
            c_nodes[i].scale[:] = [float(x) for x in nd.scale]
            c_nodes[i].rotation[:] = [float(x) for x in np.asarray(nd.rotation, np.float32).reshape(16)]
            c_nodes[i].parent = parent
        self.keep.append(c_nodes)

        sc = orc_scene()
        sc.n_nodes = len(nodes); sc.nodes = c_nodes
        sc.n_meshes = len(meshes); sc.meshes = c_meshes
        sc.n_models = len(desc.models); sc.models = c_models
        sc.n_lights = len(all_lights); sc.lights = c_lights
        sc.n_scene_lights = len(scene_light_ids); sc.scene_lights = scene_light_ids.ctypes.data_as(i32p)
        sc.n_light_groups = len(groups); sc.light_groups = c_groups
        sc.n_volumes = len(volumes); sc.volumes = c_volumes
        lv = np.ascontiguousarray(np.asarray(list(desc.autosubdivision_levels), np.float32).reshape(-1, 2)); self.keep.append(lv)
        sc.n_subdivision_levels = len(lv); sc.subdivision_levels = _fp(lv) if len(lv) else None
        cd = desc.camera
        sc.camera.node = cam_node; sc.camera.width = int(cd.width); sc.camera.height = int(cd.height)
        sc.camera.field_of_view = float(cd.field_of_view); sc.camera.near_plane = float(cd.near); sc.camera.far_plane = float(cd.far)
        sc.camera.perspective = int(cd.perspective); sc.camera.ortho_scale = float(cd.ortho_scale)
        sc.camera.depth_margin = float(cd.depth_margin); sc.camera.vertex_snapping = float(cd.vertex_snapping)
        sc.camera.render_depth = int(cd.render_depth); sc.camera.render_normals = int(cd.render_normals)
        sc.camera.perspective_corrected_texture_mapping = int(cd.perspective_corrected_texture_mapping)
        sc.camera.max_light_count = int(cd.max_light_count)
        sc.camera.max_light_focus_node = focus_node
        w = desc.world
        sc.world.present = int(w.present); sc.world.lighting_on = int(w.lighting_on); sc.world.fog_on = int(w.fog_on)
        sc.world.fog_mode = int(w.fog_mode); sc.world.fog_color[:] = [float(x) for x in w.fog_color]
        sc.world.fog_range[:] = [float(x) for x in w.fog_range]; sc.world.ambient_color[:] = [float(x) for x in w.ambient_color]
        sc.world.ambient_energy = float(w.ambient_energy); sc.world.ambient_visible = int(w.ambient_visible)
        self.scene = sc
        self.desc = desc
        self.light_index = light_index
        self.model_index = model_index


def marshal(desc: SceneDesc) -> _Marshalled:
    return _Marshalled(desc)


def bake_ao(desc_or_marshalled, model, tri_normals, tri_max_span, occlusion_angle, occlusion_color, colors4: Optional[np.ndarray] = None,
            target_parts=None, target_vertices=None) -> np.ndarray:
    """Model.BakeAO's same-model pass (model.go:1040-1124) on `model`'s mesh: VertexColors[TargetChannel] after the bake,
    [n_verts][4].  tri_normals / tri_max_span: Triangle.Normal / MaxSpan; target_parts / target_vertices: index lists (None: all)."""
    m = desc_or_marshalled if isinstance(desc_or_marshalled, _Marshalled) else _Marshalled(desc_or_marshalled)
    mesh_i = m.scene.models[m.model_index[id(model)]].mesh
    n = model.mesh.n_verts
    out = np.ones((n, 4), np.float32) if colors4 is None else np.ascontiguousarray(colors4, np.float32).reshape(n, 4).copy()
    tn = np.ascontiguousarray(tri_normals, np.float32); ts = np.ascontiguousarray(tri_max_span, np.float32)
    tp = tv = None
    if target_parts is not None:
        tp = np.zeros(len(model.mesh.parts), np.uint8); tp[list(target_parts)] = 1
    if target_vertices is not None:
        tv = np.zeros(n, np.uint8); tv[list(target_vertices)] = 1
    oc = np.asarray(occlusion_color, np.float32)
    L = lib()
    rc = L.orc_bake_ao(C.byref(m.scene.meshes[mesh_i]), _fp(tn), _fp(ts), tp.ctypes.data_as(u8p) if tp is not None else None,
                       tv.ctypes.data_as(u8p) if tv is not None else None, float(occlusion_angle), _fp(oc), _fp(out))
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.orc_last_error().decode())
    return out


def go_acos(x: float) -> float:
    return float(lib().orc_go_acos(float(x)))


def bake_lighting(desc_or_marshalled, model, lights, rgb: Optional[np.ndarray] = None) -> np.ndarray:
    """Model.BakeLighting (model.go:1203-1268) of `model` (a ModelDesc of the scene) with `lights` (LightDescs of the scene, in
    order; the world's ambient light is added by the oracle): the target channel's RGB, [n_verts][3].  `rgb`: the channel before
    the call (default: 1, 1, 1 -- a channel that did not exist, mesh.go:543-558)."""
    m = desc_or_marshalled if isinstance(desc_or_marshalled, _Marshalled) else _Marshalled(desc_or_marshalled)
    n = model.mesh.n_verts
    out = np.ones((n, 3), np.float32) if rgb is None else np.ascontiguousarray(rgb, np.float32).reshape(n, 3).copy()
    ids = np.asarray([m.light_index[id(l)] for l in lights], dtype=np.int32)
    L = lib()
    rc = L.orc_bake_lighting(C.byref(m.scene), m.model_index[id(model)], ids.ctypes.data_as(i32p) if len(ids) else None, len(ids), _fp(out))
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.orc_last_error().decode())
    return out


def render_scene(desc_or_marshalled, capture: bool = False, copy: bool = True) -> OracleResult:
    """Runs the oracle's Camera.RenderScene on the scene; returns numpy copies of its lists."""
    m = desc_or_marshalled if isinstance(desc_or_marshalled, _Marshalled) else _Marshalled(desc_or_marshalled)
    L = lib()
    out = C.POINTER(orc_result)()
    rc = L.orc_render_scene(C.byref(m.scene), int(capture), C.byref(out))
    if rc != 0:
        raise RuntimeError("oracle: %s" % L.orc_last_error().decode())
    r = out.contents
    n = int(r.n_vertices)

    def arr(ptr, shape, dtype):
        count = int(np.prod(shape))
        if count == 0:
            return np.zeros(shape, dtype)
        a = np.ctypeslib.as_array(ptr, shape=(count,)).view(dtype).reshape(shape)
        return a.copy() if copy else a

    res = OracleResult(
        color_list=arr(r.color_list, (n, 12), np.float32), depth_list=arr(r.depth_list, (n, 12), np.float32),
        normal_list=arr(r.normal_list, (n, 12), np.float32), index_list=arr(r.index_list, (n,), np.uint16),
        tri_ids=arr(r.tri_ids, (n // 3,), np.uint32),
        renders=[dict(model=r.renders[i].model, part=r.renders[i].part, flush=r.renders[i].flush,
                      first_vertex=r.renders[i].first_vertex, vertex_count=r.renders[i].vertex_count,
                      total_tris=r.renders[i].total_tris, min_depth=r.renders[i].min_depth, max_depth=r.renders[i].max_depth)
                 for i in range(r.n_renders)],
        flushes=[dict(model=r.flushes[i].model, part=r.flushes[i].part, first_vertex=r.flushes[i].first_vertex,
                      vertex_count=r.flushes[i].vertex_count) for i in range(r.n_flushes)],
        cap_clip=arr(r.cap_clip, (int(r.n_captured), 4), np.float32) if capture else None,
        cap_light=arr(r.cap_light, (int(r.n_captured), 4), np.float32) if capture else None,
        debug=dict(totalParts=r.total_parts, drawnParts=r.drawn_parts, batchedParts=r.batched_parts, totalTris=r.total_tris,
                   drawnTris=r.drawn_tris, lightCount=r.light_count, activeLightCount=r.active_light_count),
        times=dict(total=r.t_total, process=r.t_process, light=r.t_light, emit=r.t_emit))
    L.orc_free_result(out)
    return res
