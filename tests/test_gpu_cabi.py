"""-m gpu: the C ABI of include/t3dgeo.h driven directly with ctypes, the way a cgo shim would:
a minimal frame end to end, and the error behaviour (negative code + t3d_last_error text, nothing aborts)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T3D_OK, T3D_E_INVALID, T3D_E_CUDA, T3D_E_UNSUPPORTED, T3D_E_NOMEM, T3D_E_STATE, T3D_E_NOTFOUND = 0, -1, -2, -3, -4, -5, -6
f32p = C.POINTER(C.c_float)


class PartRange(C.Structure):
    _fields_ = [("vtx_start", C.c_uint32), ("vtx_end", C.c_uint32), ("tri_start", C.c_uint32), ("tri_end", C.c_uint32)]


class MeshDesc(C.Structure):
    _fields_ = [("n_verts", C.c_uint32), ("positions", f32p), ("normals", f32p), ("uvs", f32p), ("colors", f32p), ("weights", f32p),
                ("bones", C.POINTER(C.c_uint16)), ("n_influences", C.POINTER(C.c_uint8)), ("n_tris", C.c_uint32),
                ("indices", C.POINTER(C.c_uint32)), ("centers", f32p), ("n_parts", C.c_uint32), ("parts", C.POINTER(PartRange)),
                ("max_span", C.c_float), ("n_shape_keys", C.c_uint32), ("key_positions", f32p), ("key_normals", f32p),
                ("tri_sub_base", C.POINTER(C.c_uint32)), ("tri_sub_level", C.POINTER(C.c_int8)), ("tri_n_levels", C.POINTER(C.c_uint8)),
                ("tri_max_span", f32p), ("tri_normal", f32p)]


class Camera(C.Structure):
    _fields_ = [("vp", C.c_float * 16), ("view", C.c_float * 16), ("position", C.c_float * 3), ("up", C.c_float * 3),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("depth_margin", C.c_float), ("vertex_snapping", C.c_float), ("flags", C.c_uint32),
                ("forward", C.c_float * 3), ("right", C.c_float * 3), ("sphere_factor_y", C.c_float), ("sphere_factor_x", C.c_float),
                ("sphere_factor_tang", C.c_float), ("ortho_scale", C.c_float)]


class World(C.Structure):
    _fields_ = [("present", C.c_int32), ("fog_on", C.c_int32), ("fog_mode", C.c_int32), ("fog_color", C.c_float * 3), ("fog_range", C.c_float * 2)]


class Draw(C.Structure):
    _fields_ = [("mesh_id", C.c_uint64), ("part", C.c_uint32), ("flags", C.c_uint32), ("transform", C.c_float * 16),
                ("world_rotation", C.c_float * 16), ("color", C.c_float * 4), ("src_w", C.c_float), ("src_h", C.c_float),
                ("sort_mode", C.c_uint8), ("lighting_mode", C.c_uint8), ("light_volume_mode", C.c_uint8), ("reserved0", C.c_uint8),
                ("bone_offset", C.c_uint32), ("bone_count", C.c_uint32), ("light_list", C.c_uint32), ("reserved1", C.c_uint32)]


class DrawResult(C.Structure):
    _fields_ = [("first_vertex", C.c_uint64), ("vertex_count", C.c_uint32), ("total_tris", C.c_uint32), ("min_depth", C.c_float), ("max_depth", C.c_float)]


class FrameOutput(C.Structure):
    _fields_ = [("n_vertices", C.c_uint64), ("n_draws", C.c_uint32), ("color_list", f32p), ("depth_list", f32p), ("normal_list", f32p),
                ("index_list", C.POINTER(C.c_uint16)), ("tri_ids", C.POINTER(C.c_uint32)), ("draws", C.POINTER(DrawResult))]


@pytest.fixture(scope="module")
def abi():
    import tetra3d_b200
    lib = C.CDLL(tetra3d_b200.LIB_PATH)
    lib.t3d_last_error.restype = C.c_char_p
    lib.t3d_last_error.argtypes = [C.c_void_p]
    for name in ("t3d_mesh_upload", "t3d_frame_begin", "t3d_submit_draws", "t3d_frame_execute", "t3d_frame_results", "t3d_mesh_release",
                 "t3d_submit_lights", "t3d_submit_bones", "t3d_create", "t3d_set_output_target"):
        getattr(lib, name).restype = C.c_int
    ctx = C.c_void_p()
    assert lib.t3d_create(C.byref(ctx), 0) == T3D_OK, lib.t3d_last_error(None)
    yield lib, ctx
    lib.t3d_destroy.argtypes = [C.c_void_p]
    lib.t3d_destroy(ctx)


def _triangle_mesh():
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (3, 1))
    uv = np.array([[0, 0], [1, 0], [0.5, 1]], np.float32)
    idx = np.array([0, 1, 2], np.uint32)
    ctr = pos.mean(axis=0).astype(np.float32)
    parts = (PartRange * 1)(PartRange(0, 3, 0, 1))
    d = MeshDesc()
    d.n_verts = 3; d.positions = pos.ctypes.data_as(f32p); d.normals = nrm.ctypes.data_as(f32p); d.uvs = uv.ctypes.data_as(f32p)
    d.n_tris = 1; d.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32)); d.centers = ctr.ctypes.data_as(f32p)
    d.n_parts = 1; d.parts = parts; d.max_span = 2.0
    return d, (pos, nrm, uv, idx, ctr, parts)


def _camera():
    cam = Camera()
    # view = translate(0, 0, -5); a projection whose row-vector form keeps the triangle in front of the camera
    n, f = 0.1, 100.0
    t = n * np.tan(np.radians(60.0) / 2); r = t * (640 / 360)
    proj = np.array([[n / r, 0, 0, 0], [0, n / t, 0, 0], [0, 0, -(f + n) / (f - n), -2 * f * n / (f - n)], [0, 0, -1, 0]], np.float32)
    view = np.eye(4, dtype=np.float32); view[3, 2] = -5.0
    vp = (view @ proj.T).astype(np.float32)          # row vectors: v * view * projT (matrix.go:588-604 is applied row-wise)
    cam.vp[:] = vp.reshape(-1).tolist(); cam.view[:] = view.reshape(-1).tolist()
    cam.position[:] = [0, 0, 5]; cam.up[:] = [0, 1, 0]
    cam.near_plane = n; cam.far_plane = f; cam.width = 640; cam.height = 360; cam.depth_margin = 0.04
    cam.flags = 1 | 2       # perspective, render depth
    return cam


def _draw(mesh_id=7, part=0):
    d = Draw()
    d.mesh_id = mesh_id; d.part = part; d.flags = 0
    d.transform[:] = np.eye(4, dtype=np.float32).reshape(-1).tolist()
    d.world_rotation[:] = np.eye(4, dtype=np.float32).reshape(-1).tolist()
    d.color[:] = [1, 0.5, 0.25, 1]
    return d


def test_minimal_frame_through_the_c_abi(abi):
    lib, ctx = abi
    md, keep = _triangle_mesh()
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(7), C.c_uint32(1), C.byref(md)) == T3D_OK, lib.t3d_last_error(ctx)
    cam, world = _camera(), World()
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    dr = _draw()
    assert lib.t3d_submit_draws(ctx, C.byref(dr), C.c_uint32(1)) == T3D_OK, lib.t3d_last_error(ctx)
    assert lib.t3d_frame_execute(ctx) == T3D_OK, lib.t3d_last_error(ctx)
    out = FrameOutput()
    assert lib.t3d_frame_results(ctx, C.byref(out), None) == T3D_OK, lib.t3d_last_error(ctx)
    assert out.n_draws == 1 and out.n_vertices == 3
    r = out.draws[0]
    assert (r.first_vertex, r.vertex_count, r.total_tris) == (0, 3, 1)
    assert abs(r.min_depth - 25.111113) < 1e-3 and r.min_depth == r.max_depth       # |cam - centre|^2 = 25 + 1/9
    v = np.ctypeslib.as_array(out.color_list, shape=(3, 12))
    assert np.allclose(v[:, 4:8], [[1, 0.5, 0.25, 1]] * 3)                         # unlit: model colour
    assert np.all((v[:, 0] > 0) & (v[:, 0] < 640) & (v[:, 1] > 0) & (v[:, 1] < 360))   # on screen
    assert [out.index_list[i] for i in range(3)] == [0, 1, 2] and out.tri_ids[0] == 0
    # re-upload with the same generation is a no-op; release works; the mesh is gone afterwards
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(7), C.c_uint32(1), C.byref(md)) == T3D_OK


def test_error_codes_and_messages(abi):
    lib, ctx = abi
    cam, world = _camera(), World()
    dr = _draw()
    # call order
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    assert lib.t3d_frame_execute(ctx) == T3D_OK                                  # an empty frame is fine
    assert lib.t3d_frame_execute(ctx) == T3D_E_STATE and b"no open frame" in lib.t3d_last_error(ctx)
    assert lib.t3d_submit_draws(ctx, C.byref(dr), C.c_uint32(1)) == T3D_E_STATE
    out = FrameOutput()
    assert lib.t3d_frame_results(ctx, C.byref(out), None) == T3D_OK and out.n_vertices == 0
    # arguments
    assert lib.t3d_frame_begin(ctx, None, None) == T3D_E_INVALID
    bad = _camera(); bad.width = 0
    assert lib.t3d_frame_begin(ctx, C.byref(bad), C.byref(world)) == T3D_E_INVALID and b"camera size" in lib.t3d_last_error(ctx)
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    unknown = _draw(mesh_id=12345)
    assert lib.t3d_submit_draws(ctx, C.byref(unknown), C.c_uint32(1)) == T3D_E_NOTFOUND and b"unknown mesh" in lib.t3d_last_error(ctx)
    md, keep = _triangle_mesh()
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(8), C.c_uint32(1), C.byref(md)) == T3D_OK
    badpart = _draw(mesh_id=8, part=3)
    assert lib.t3d_submit_draws(ctx, C.byref(badpart), C.c_uint32(1)) == T3D_E_INVALID and b"part index" in lib.t3d_last_error(ctx)
    badenum = _draw(mesh_id=8); badenum.sort_mode = 9
    assert lib.t3d_submit_draws(ctx, C.byref(badenum), C.c_uint32(1)) == T3D_E_INVALID
    skinned = _draw(mesh_id=8); skinned.flags = 1 << 3
    assert lib.t3d_submit_draws(ctx, C.byref(skinned), C.c_uint32(1)) == T3D_E_INVALID and b"without weights" in lib.t3d_last_error(ctx)
    nolist = _draw(mesh_id=8); nolist.light_list = 77
    assert lib.t3d_submit_draws(ctx, C.byref(nolist), C.c_uint32(1)) == T3D_E_INVALID
    # the frame is still usable after rejected submits
    ok = _draw(mesh_id=8)
    assert lib.t3d_submit_draws(ctx, C.byref(ok), C.c_uint32(1)) == T3D_OK
    assert lib.t3d_frame_execute(ctx) == T3D_OK and lib.t3d_frame_results(ctx, C.byref(out), None) == T3D_OK and out.n_vertices == 3
    # mesh upload validation
    md2, keep2 = _triangle_mesh()
    keep2[3][2] = 5                                                             # index out of range
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(9), C.c_uint32(1), C.byref(md2)) == T3D_E_INVALID and b"out of range" in lib.t3d_last_error(ctx)
    assert lib.t3d_mesh_release(ctx, C.c_uint64(4242)) == T3D_E_NOTFOUND
    assert lib.t3d_mesh_release(ctx, C.c_uint64(8)) == T3D_OK


class Light(C.Structure):
    _fields_ = [("type", C.c_uint32), ("color", C.c_float * 3), ("energy", C.c_float), ("position", C.c_float * 3), ("range", C.c_float),
                ("forward", C.c_float * 3), ("volume", C.c_uint32)]


class VolumeDesc(C.Structure):
    _fields_ = [("dim_min", C.c_float * 3), ("dim_max", C.c_float * 3), ("cell_size", C.c_float * 3), ("cell_count", C.c_uint32 * 3),
                ("offset", C.c_float * 3), ("light_bias", C.c_float * 3), ("cell_colors", f32p),
                ("cell_light_start", C.POINTER(C.c_uint32)), ("cell_light_index", C.POINTER(C.c_uint32))]


def test_volume_cell_lights_and_subdivision_call_rules(abi):
    """t3d_volume_use_lights / t3d_frame_set_autosubdivision: call order and argument checks (light.go:565-580, 1508-1514;
    camera.go:1547-1553)."""
    lib, ctx = abi
    for name in ("t3d_volume_upload", "t3d_volume_use_lights", "t3d_volume_release", "t3d_frame_set_autosubdivision"):
        getattr(lib, name).restype = C.c_int
    cam, world = _camera(), World()
    colors = np.ones(2 * 3, np.float32)
    start = np.array([0, 1, 3], np.uint32); index = np.array([0, 1, 0], np.uint32)
    vd = VolumeDesc()
    vd.dim_min[:] = (-1, -1, -1); vd.dim_max[:] = (1, 1, 1); vd.cell_size[:] = (1, 2, 2); vd.cell_count[:] = (2, 1, 1)
    vd.cell_colors = colors.ctypes.data_as(f32p)
    vd.cell_light_start = start.ctypes.data_as(C.POINTER(C.c_uint32)); vd.cell_light_index = index.ctypes.data_as(C.POINTER(C.c_uint32))
    assert lib.t3d_volume_upload(ctx, C.c_uint32(5), C.byref(vd)) == T3D_OK
    bad_start = np.array([0, 2, 1], np.uint32)
    vd2 = VolumeDesc(); C.memmove(C.byref(vd2), C.byref(vd), C.sizeof(vd)); vd2.cell_light_start = bad_start.ctypes.data_as(C.POINTER(C.c_uint32))
    assert lib.t3d_volume_upload(ctx, C.c_uint32(6), C.byref(vd2)) == T3D_E_INVALID and b"non-decreasing" in lib.t3d_last_error(ctx)
    point = Light(); point.type = 1; point.energy = 1.0
    volume = Light(); volume.type = 3; volume.volume = 5; volume.energy = 1.0
    lights2 = (Light * 2)(point, point)
    lid = C.c_uint32()
    levels = (C.c_float * 4)(100.0, 1.0, 25.0, 0.5)
    # outside a frame
    assert lib.t3d_volume_use_lights(ctx, C.c_uint32(5), lights2, C.c_uint32(2)) == T3D_E_STATE
    assert lib.t3d_frame_set_autosubdivision(ctx, levels, C.c_uint32(2)) == T3D_E_STATE
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    # a list holding the volume before its cell lights were passed
    assert lib.t3d_submit_lights(ctx, C.byref(volume), C.c_uint32(1), C.byref(lid)) == T3D_E_STATE and b"t3d_volume_use_lights" in lib.t3d_last_error(ctx)
    assert lib.t3d_volume_use_lights(ctx, C.c_uint32(99), lights2, C.c_uint32(2)) == T3D_E_NOTFOUND
    assert lib.t3d_volume_use_lights(ctx, C.c_uint32(5), lights2, C.c_uint32(1)) == T3D_E_INVALID and b"fewer lights" in lib.t3d_last_error(ctx)
    nested = (Light * 2)(point, volume)
    assert lib.t3d_volume_use_lights(ctx, C.c_uint32(5), nested, C.c_uint32(2)) == T3D_E_UNSUPPORTED
    assert lib.t3d_volume_use_lights(ctx, C.c_uint32(5), lights2, C.c_uint32(2)) == T3D_OK
    assert lib.t3d_submit_lights(ctx, C.byref(volume), C.c_uint32(1), C.byref(lid)) == T3D_OK and lid.value >= 1
    # subdivision levels
    assert lib.t3d_frame_set_autosubdivision(ctx, levels, C.c_uint32(2)) == T3D_OK
    assert lib.t3d_frame_set_autosubdivision(ctx, levels, C.c_uint32(9)) == T3D_E_UNSUPPORTED
    zero = (C.c_float * 2)(100.0, 0.0)
    assert lib.t3d_frame_set_autosubdivision(ctx, zero, C.c_uint32(1)) == T3D_E_INVALID and b"positive" in lib.t3d_last_error(ctx)
    assert lib.t3d_frame_execute(ctx) == T3D_OK
    assert lib.t3d_volume_release(ctx, C.c_uint32(5)) == T3D_OK
    # mesh upload: the five subdivision arrays come together
    md, keep = _triangle_mesh()
    base = np.zeros(1, np.uint32)
    md.tri_sub_base = base.ctypes.data_as(C.POINTER(C.c_uint32))
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(21), C.c_uint32(1), C.byref(md)) == T3D_E_INVALID and b"all five" in lib.t3d_last_error(ctx)


# ---- t3d_submit_instances: the device builds Node.Transform (node.go:577-583) -------------------------------------------
class Instance(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("scale", C.c_float * 3), ("color", C.c_float * 4)]


def _go_mult(a, b):
    """Matrix4.Mult (matrix_amd64.s:4-81 / matrix_none.go:6-32): row i = sum_k a[i][k] * b[k], k ascending, float32, multiply then add."""
    out = np.zeros((4, 4), np.float32)
    for i in range(4):
        acc = np.float32(a[i, 0]) * b[0]
        for k in range(1, 4):
            acc = (acc + np.float32(a[i, k]) * b[k]).astype(np.float32)
        out[i] = acc
    return out


def _node_transform(pos, scale, rot, parent):
    m = _go_mult(np.diag(np.array([scale[0], scale[1], scale[2], 1], np.float32)), rot)
    t = np.eye(4, dtype=np.float32); t[3, :3] = pos
    m = _go_mult(m, t)
    return _go_mult(m, parent) if parent is not None else m


def _rot_y(a):
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], np.float32)


def _frame(lib, ctx, submit):
    cam, world = _camera(), World()
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    submit()
    assert lib.t3d_frame_execute(ctx) == T3D_OK, lib.t3d_last_error(ctx)
    out = FrameOutput()
    assert lib.t3d_frame_results(ctx, C.byref(out), None) == T3D_OK, lib.t3d_last_error(ctx)
    n = int(out.n_vertices)
    col = np.ctypeslib.as_array(out.color_list, shape=(max(n, 1), 12))[:n].copy()
    dep = np.ctypeslib.as_array(out.depth_list, shape=(max(n, 1), 12))[:n].copy()
    res = [(out.draws[i].first_vertex, out.draws[i].vertex_count, out.draws[i].min_depth, out.draws[i].max_depth) for i in range(out.n_draws)]
    return col, dep, res


@pytest.mark.parametrize("n,uniform,with_parent,with_rotation,mixed", [
    (40, True, False, False, False),        # small frame: the host expands the instances (same operations)
    (3000, True, False, False, False),      # a frame of nothing but uniformly scaled instances: 32-byte payloads
    (3000, False, True, True, False),       # non-uniform scale, shared rotation, parent: 64-byte payloads
    (2500, True, True, False, True),        # instances between ordinary draws
])
def test_instances_equal_draws_bit_for_bit(abi, n, uniform, with_parent, with_rotation, mixed):
    lib, ctx = abi
    lib.t3d_submit_instances.restype = C.c_int
    md, keep = _triangle_mesh()
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(7), C.c_uint32(1), C.byref(md)) == T3D_OK, lib.t3d_last_error(ctx)
    rng = np.random.default_rng(n)
    pos = rng.uniform((-6, -3, -20), (6, 3, 2), (n, 3)).astype(np.float32)
    scl = rng.uniform(0.2, 1.5, (n, 3)).astype(np.float32)
    scl[::7, 0] *= -1          # mirrored instances: products with -0
    if uniform:
        scl[:] = scl[:, :1]
    colr = rng.uniform(0.1, 1.0, (n, 4)).astype(np.float32)
    rot = _rot_y(0.7) if with_rotation else np.eye(4, dtype=np.float32)
    parent = _node_transform((0.5, -0.25, 1.0), (1.0, 2.0, 1.0), _rot_y(-0.3), None) if with_parent else None
    mat_color = np.array([0.9, 0.8, 0.7, 1.0], np.float32)
    extra = _draw(); extra.transform[12] = 1.5

    def as_instances():
        if mixed:
            assert lib.t3d_submit_draws(ctx, C.byref(extra), C.c_uint32(1)) == T3D_OK
        proto = _draw(); proto.color[:] = mat_color.tolist()
        inst = (Instance * n)()
        for i in range(n):
            inst[i].position[:] = pos[i].tolist(); inst[i].scale[:] = scl[i].tolist(); inst[i].color[:] = colr[i].tolist()
        r16 = (C.c_float * 16)(*rot.reshape(-1).tolist())
        p16 = (C.c_float * 16)(*parent.reshape(-1).tolist()) if parent is not None else None
        assert lib.t3d_submit_instances(ctx, C.byref(proto), r16 if with_rotation else None, p16, None, inst, C.c_uint32(n)) == T3D_OK, lib.t3d_last_error(ctx)
        if mixed:
            assert lib.t3d_submit_draws(ctx, C.byref(extra), C.c_uint32(1)) == T3D_OK

    def as_draws():
        if mixed:
            assert lib.t3d_submit_draws(ctx, C.byref(extra), C.c_uint32(1)) == T3D_OK
        arr = (Draw * n)()
        for i in range(n):
            d = _draw()
            d.transform[:] = _node_transform(pos[i], scl[i], rot, parent).reshape(-1).tolist()
            d.color[:] = (colr[i] * mat_color).astype(np.float32).tolist()
            d.flags = 0 if i == 0 else (1 << 10)        # T3D_DRAW_CONTINUE_FLUSH
            arr[i] = d
        assert lib.t3d_submit_draws(ctx, arr, C.c_uint32(n)) == T3D_OK, lib.t3d_last_error(ctx)
        if mixed:
            assert lib.t3d_submit_draws(ctx, C.byref(extra), C.c_uint32(1)) == T3D_OK

    ci, di, ri = _frame(lib, ctx, as_instances)
    cd, dd, rd = _frame(lib, ctx, as_draws)
    assert len(ci) == len(cd) and len(ci) > 0
    assert np.array_equal(ci.view(np.uint32), cd.view(np.uint32)) and np.array_equal(di.view(np.uint32), dd.view(np.uint32))
    assert ri == rd


def test_instances_reject_what_they_cannot_express(abi):
    lib, ctx = abi
    lib.t3d_submit_instances.restype = C.c_int
    cam, world = _camera(), World()
    assert lib.t3d_frame_begin(ctx, C.byref(cam), C.byref(world)) == T3D_OK
    proto = _draw(); inst = (Instance * 1)()
    bad = np.eye(4, dtype=np.float32); bad[0, 3] = 0.5
    b16 = (C.c_float * 16)(*bad.reshape(-1).tolist())
    assert lib.t3d_submit_instances(ctx, C.byref(proto), b16, None, None, inst, C.c_uint32(1)) == T3D_E_UNSUPPORTED and b"affine" in lib.t3d_last_error(ctx)
    proto.flags = 1 << 3     # T3D_DRAW_SKINNED
    assert lib.t3d_submit_instances(ctx, C.byref(proto), None, None, None, inst, C.c_uint32(1)) == T3D_E_UNSUPPORTED
    assert lib.t3d_frame_execute(ctx) == T3D_OK


def test_draws_submitted_one_call_each_equal_one_call(abi):
    """A shim may call t3d_submit_draws once per render(rp): the frame must not depend on how the calls are cut."""
    lib, ctx = abi
    md, keep = _triangle_mesh()
    assert lib.t3d_mesh_upload(ctx, C.c_uint64(7), C.c_uint32(1), C.byref(md)) == T3D_OK, lib.t3d_last_error(ctx)
    n = 2200
    rng = np.random.default_rng(5)
    arr = (Draw * n)()
    for i in range(n):
        d = _draw()
        d.transform[12], d.transform[13], d.transform[14] = [float(x) for x in rng.uniform((-5, -3, -15), (5, 3, 1))]
        arr[i] = d

    def one_call():
        assert lib.t3d_submit_draws(ctx, arr, C.c_uint32(n)) == T3D_OK, lib.t3d_last_error(ctx)

    def many_calls():
        for i in range(n):
            assert lib.t3d_submit_draws(ctx, C.byref(arr[i]), C.c_uint32(1)) == T3D_OK, lib.t3d_last_error(ctx)

    c1, d1, r1 = _frame(lib, ctx, one_call)
    c2, d2, r2 = _frame(lib, ctx, many_calls)
    assert len(c1) > 0 and np.array_equal(c1.view(np.uint32), c2.view(np.uint32)) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32)) and r1 == r2
