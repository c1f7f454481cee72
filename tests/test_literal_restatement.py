"""CPU: the oracle against a second restatement of the hot path that was written separately from it and keeps the Go source's own
structure (tests/literal_go.py: ProcessVertices' per-triangle loop over global scratch arrays, the 512-bin bucket as lists, light
objects with beginRender / beginModel / VertexLight, the lighting and emit loops of camera.render).  Every render(rp) call the
oracle reports must come out of the literal code with the same triangles in the same order and the same colour and depth records,
bit for bit -- on the synthetic parity scenes and on the scenes built from the reference's own GLB assets.  A misreading of the Go
source would have to be made twice, in two differently shaped programs, to pass."""
import gzip
import os
import pickle
import warnings

import numpy as np
import pytest

import literal_go as LG
import scenes
from oracle import binding as O
from tetra3d_b200.scenedesc import CameraDesc, MaterialDesc, WorldDesc

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assets")


def covered(desc):
    """Scenes the literal code restates: static and skinned models, shape keys, billboards, dynamic-batch children, every light
    type, auto-subdivision, MaxLightCount."""
    for m in desc.models:
        skinned = m.bone_influences is not None or m.pose is not None
        for p in m.mesh.parts:      # (the reference reads a never-written global for unbillboarded depth on skinned models, model.go:800)
            if skinned and p.material is not None and p.material.billboarded_depth_mode:
                return False
    return True


def same_bits(a, b):
    """Bit for bit, any NaN equal to any NaN (payload bits are not part of the contract)."""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def check(desc):
    ref = O.render_scene(desc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)      # divisions by a zero w and the like are part of the data (Go does not trap them either)
        lit = LG.render_scene(desc)
    n_checked = 0
    for r in ref.renders:
        ids, col, dep = lit[(r["model"], r["part"])]
        a, n = r["first_vertex"], r["vertex_count"]
        assert n == len(col) == len(dep) == 3 * len(ids), (desc.name, r, len(ids))
        assert list(ref.tri_ids[a // 3:(a + n) // 3]) == ids, (desc.name, r["model"], r["part"], "triangle set / order")
        assert same_bits(ref.color_list[a:a + n], col), (desc.name, r["model"], r["part"], "colour list")
        assert same_bits(ref.depth_list[a:a + n], dep), (desc.name, r["model"], r["part"], "depth list")
        n_checked += n
    # which render pairs there are (visibility, the frustum test of every model) and the order they are rendered in
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        assert LG.render_order(desc) == [(r["model"], r["part"]) for r in ref.renders], (desc.name, "render pairs / order")
    return n_checked


SYNTHETIC = {
    "soup_lit": lambda: scenes.soup_scene(1),
    "soup_unlit_colors": lambda: scenes.soup_scene(2, lit=False, colors=True),
    "soup_ranged": lambda: scenes.soup_scene(3, ranged=True),
    "soup_parent": lambda: scenes.soup_scene(4, parent=True),
    "soup_ortho": lambda: scenes.soup_scene(5, cam=CameraDesc(perspective=False, ortho_scale=16.0)),
    "soup_snap_pctm": lambda: scenes.soup_scene(6, cam=CameraDesc(vertex_snapping=8.0, perspective_corrected_texture_mapping=True)),
    "soup_fog_add": lambda: scenes.soup_scene(7, cam=CameraDesc(render_depth=False), world=WorldDesc(fog_mode=0, fog_color=(0.3, 0.2, 0.1, 1), fog_range=(0.1, 0.9))),
    "soup_fog_sub": lambda: scenes.soup_scene(8, cam=CameraDesc(render_depth=False), world=WorldDesc(fog_mode=1, fog_color=(0.9, 0.8, 0.7, 1))),
    "soup_fixed_normals": lambda: scenes.soup_scene(9, material=MaterialDesc(lighting_mode=1, backface_culling=False)),
    "soup_double_sided": lambda: scenes.soup_scene(10, material=MaterialDesc(lighting_mode=2, backface_culling=False, triangle_sort_mode=1)),
    "soup_sort_none": lambda: scenes.soup_scene(13, material=MaterialDesc(triangle_sort_mode=2)),
    "soup_lighting_off": lambda: scenes.soup_scene(11, world=WorldDesc(lighting_on=False)),
    "soup_no_world": lambda: scenes.soup_scene(12, world=WorldDesc(present=False)),
    "transparent": lambda: scenes.transparent_scene(1),
    "billboard": lambda: scenes.billboard_scene(1),
    "billboard_locks": lambda: scenes.billboard_scene(2, locks=(True, False, True), camera_up=True),
    "billboard_lock_y_unbill": lambda: scenes.billboard_scene(3, locks=(False, True, False), unbillboarded=True),
    "batch_transparent": lambda: scenes.batch_scene(1),
    "batch_solid": lambda: scenes.batch_scene(2, transparent=False),
    "batch_frustum": lambda: scenes.batch_frustum_scene(3, n_children=60),
    "volume_mode0": lambda: scenes.volume_scene(1, 0),
    "volume_mode1": lambda: scenes.volume_scene(2, 1),
    "volume_mode2": lambda: scenes.volume_scene(3, 2),
    "volume_off": lambda: scenes.volume_scene(4, 3),
    "volume_cell_lights": lambda: scenes.volume_cell_lights_scene(5, 0),
    "volume_cell_lights_normal": lambda: scenes.volume_cell_lights_scene(6, 1, skinned=True),
    "volume_cell_lights_object": lambda: scenes.volume_cell_lights_scene(7, 2),
    "max_light_count": lambda: scenes.max_light_count_scene(51),
    "max_light_count_focus": lambda: scenes.max_light_count_scene(52, n_lights=9, max_count=4, focus=True),
    "max_light_count_many": lambda: scenes.max_light_count_scene(53, n_lights=47, max_count=6),
    "render_order_45_solids": lambda: scenes.render_order_scene(0, n=45),
    "subdivision": lambda: scenes.subdivision_scene(0),
    "subdivision_no_levels": lambda: scenes.subdivision_scene(1, with_levels=False),
    "skinned": lambda: scenes.skinned_scene(1, n=4, rings=6, sides=7),
    "shape_keys": lambda: scenes.shape_key_scene(1),
    "posed": lambda: scenes.posed_scene(0, n_models=4),
    "posed_blending": lambda: scenes.posed_scene(41, n_models=4, blends=[250, None, 999]),
    "render_order": lambda: scenes.render_order_scene(2),
}


@pytest.mark.parametrize("name", sorted(SYNTHETIC))
def test_oracle_equals_the_literal_restatement(name):
    desc = SYNTHETIC[name]()
    assert covered(desc), name
    assert check(desc) > 0


def _asset_names():
    return sorted(f[:-7] for f in os.listdir(ASSETS) if f.endswith(".pkl.gz"))


@pytest.mark.parametrize("name", _asset_names())
def test_oracle_equals_the_literal_restatement_on_reference_assets(name):
    with gzip.open(os.path.join(ASSETS, name + ".pkl.gz"), "rb") as f:
        desc = pickle.load(f)
    if not covered(desc):
        pytest.skip("uses a feature the literal code does not restate")
    if desc.total_triangles() > 6000:
        pytest.skip("too large for a scalar Python loop in the CPU suite")
    check(desc)


def test_fuzz_scenes_the_literal_code_covers():
    n = 0
    for seed in range(120):
        desc = scenes.fuzz_scene(seed)
        if covered(desc) and desc.total_triangles() <= 6000:
            check(desc); n += 1
    assert n >= 4, n


def test_matrix_primitives_of_both_restatements_agree():
    # Matrix4.Mult (matrix_amd64.s order), Inverted (derived from the cofactor pattern, not typed off), Node.Transform, projections
    from tetra3d_b200.scenedesc import NodeDesc
    rng = np.random.default_rng(0)
    with np.errstate(all="ignore"):
        for i in range(200):
            a, b = rng.standard_normal((4, 4)).astype(np.float32), rng.standard_normal((4, 4)).astype(np.float32)
            if i % 3 == 0:
                a[:, 3] = [0, 0, 0, 1]
            assert np.array_equal(LG.mat4_mult(a, b).view(np.uint32), np.asarray(O.mat4_mult(a, b), np.float32).reshape(4, 4).view(np.uint32))
            assert np.array_equal(LG.mat4_inverted(a).view(np.uint32), np.asarray(O.mat4_inverted(a), np.float32).reshape(4, 4).view(np.uint32))
    for i in range(20):
        parent = NodeDesc(position=tuple(rng.uniform(-3, 3, 3)), scale=tuple(rng.uniform(0.5, 2, 3)), rotation=scenes.random_rotation(rng))
        node = NodeDesc(position=tuple(rng.uniform(-3, 3, 3)), scale=tuple(rng.uniform(0.5, 2, 3)), rotation=scenes.random_rotation(rng), parent=parent if i % 2 else None)
        assert np.array_equal(LG.node_transform(node).view(np.uint32), O.node_world_transform(node).reshape(4, 4).view(np.uint32))
    for fov, near, far, w, h in [(60.0, 0.1, 100.0, 640, 360), (75.0, 0.05, 250.0, 1920, 1080), (30.0, 1.0, 20.0, 100, 100)]:
        assert np.array_equal(LG.perspective(fov, near, far, w, h).view(np.uint32), np.asarray(O.mat4_perspective(fov, near, far, w, h), np.float32).reshape(4, 4).view(np.uint32))


def test_both_restatements_of_go_stable_sort_agree_on_arbitrary_relations():
    # sort.SliceStable is handed comparators that are not orderings (camera.go:1460-1469, 1484-1495), so its outcome is defined by
    # the algorithm alone; the oracle's C restatement and literal_go's Python one must make the same sequence of decisions
    import ctypes
    lib = O.lib()
    lib.orc_test_go_stable_sort_table.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int)]
    lib.orc_test_go_stable_sort_table.restype = None
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(0, 130))
        table = (rng.random((max(n, 1), max(n, 1))) < rng.uniform(0.1, 0.9)).astype(np.uint8)
        perm = np.zeros(max(n, 1), np.int32)
        lib.orc_test_go_stable_sort_table(n, table.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        assert list(perm[:n]) == LG.slice_stable(list(range(n)), lambda x, y: bool(table[x][y])), trial


def test_bakes_equal_the_literal_restatement():
    # Model.BakeLighting (model.go:1203-1268: every light once per corner, nothing culled) and Model.BakeAO's same-model pass
    # (model.go:1040-1124), restated in literal_go.py
    import tetra3d_b200 as t3d
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for desc in (scenes.soup_scene(1), scenes.soup_scene(3, ranged=True), scenes.soup_scene(4, parent=True), scenes.skinned_scene(1, n=3, rings=6, sides=7),
                     scenes.volume_scene(1, 0), scenes.volume_scene(2, 1), scenes.volume_scene(3, 2)):
            for md in desc.models[:2]:
                a = O.bake_lighting(desc, md, list(desc.lights))
                b = LG.bake_lighting(desc, md, list(desc.lights))
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (desc.name, md.name)
            md = desc.models[0]
            sub = list(desc.lights)[:1]
            assert np.array_equal(O.bake_lighting(desc, md, sub).view(np.uint32), LG.bake_lighting(desc, md, sub).view(np.uint32))
        for desc, angle, parts, verts in ((scenes.soup_scene(4, n_models=1, n_tris=250), 0.5, None, None), (scenes.soup_scene(6, n_models=1, n_tris=200), 1.2, [0], None),
                                          (scenes.skinned_scene(1, n=1, rings=5, sides=6), 0.8, None, [0, 1, 2, 5, 7, 11])):
            md = desc.models[0]
            scene, cam, objs = t3d.build_scene(desc, None)
            normals, spans = objs[id(md.mesh)].TriangleNormalsAndSpans()
            oc = [0.1, 0.2, 0.3, 0.9]
            a = O.bake_ao(desc, md, normals, spans, angle, oc, target_parts=parts, target_vertices=verts)
            b = LG.bake_ao(md.mesh, np.asarray(normals, np.float32).reshape(-1, 3), np.asarray(spans, np.float32), angle, oc, target_parts=parts, target_vertices=verts)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (desc.name, angle)
            assert not np.array_equal(a, np.ones_like(a))



@pytest.mark.gpu
@pytest.mark.parametrize("name", ["soup_lit", "soup_ranged", "billboard_locks", "skinned", "volume_cell_lights", "subdivision", "max_light_count"])
def test_gpu_lists_equal_the_literal_restatement(device, name):
    # the CUDA path against the second restatement directly (exact-lighting mode: colours bit for bit too)
    import tetra3d_b200 as t3d
    desc = SYNTHETIC[name]()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        lit = LG.render_scene(desc)
    scene, cam, objs = t3d.build_scene(desc, device)
    device.SetExactLighting(True)
    try:
        cam.Clear(); cam.RenderScene(scene)
    finally:
        device.SetExactLighting(False)
    col, dep, tri = cam.ColorVertexList(), cam.DepthVertexList(), cam.TriangleIDs()
    n_checked = 0
    ref_renders = O.render_scene(desc).renders              # (only for the model index of every render(rp) call: same calls, same order)
    renders = cam.Renders()
    assert len(renders) == len(ref_renders)
    for r, rr in zip(renders, ref_renders):
        assert r["model"] == desc.models[rr["model"]].name and r["part"] == rr["part"]
        ids, lcol, ldep = lit[(rr["model"], rr["part"])]
        a, n = r["first_vertex"], r["vertex_count"]
        assert n == len(lcol)
        assert list(tri[a // 3:(a + n) // 3]) == ids
        assert np.array_equal(col[a:a + n].view(np.uint32), lcol.view(np.uint32)) and np.array_equal(dep[a:a + n].view(np.uint32), ldep.view(np.uint32))
        n_checked += n
    assert n_checked > 0


def test_edge_cases_equal_the_literal_restatement():
    # the corners where a restatement goes wrong first: vertices behind the camera (both w clamps), zero and mirrored scales
    # (divisions by zero in Decompose), a point light sitting on a vertex (1 / 0, Unit of a zero vector), the 0.001 energy cut-off, a
    # range that just misses, NaN and Inf positions, a scale of 1e18, one triangle (rangeDiff == 0), nothing surviving, a camera under
    # rotated parents with an orthographic projection, snapping and the fog mode that does nothing, alpha 0 and invisible materials
    from tetra3d_b200 import workloads as W
    from tetra3d_b200.scenedesc import LIGHT_DIRECTIONAL, LIGHT_POINT, LightDesc, ModelDesc, NodeDesc, SceneDesc
    rng = np.random.default_rng(0)
    lights = lambda: scenes.lights_mixed(np.random.default_rng(1), ranged=False)
    soup = lambda seed, **kw: scenes.random_soup(120, seed, **kw)
    cases = [SceneDesc([ModelDesc(soup(1, extent=12.0), NodeDesc(position=(0, 0, 9.5)), frustum_culling=False, name="m")], lights(), CameraDesc(),
                       WorldDesc(ambient_energy=0.2), name="straddle"),
             SceneDesc([ModelDesc(soup(2), NodeDesc(scale=(1.0, 0.0, 1.0)), frustum_culling=False, name="flat"),
                        ModelDesc(soup(3), NodeDesc(position=(2, 0, 0), scale=(-1.0, 1.0, 1.5)), name="mirror")], lights(), CameraDesc(), WorldDesc(ambient_energy=0.2), name="scales")]
    m = soup(4)
    L = [LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in m.positions[0])), color=(1, 0.5, 0.25), energy=2.0, name="onvertex"),
         LightDesc(LIGHT_POINT, NodeDesc(position=(30, 0, 0)), energy=0.0005, name="tiny"),
         LightDesc(LIGHT_POINT, NodeDesc(position=(0, 3, 0)), energy=3.0, range=2.5, name="ranged"),
         LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=W.rotation_matrix((1, 0, 0), 0.7)), energy=0.0, name="darksun")]
    cases.append(SceneDesc([ModelDesc(m, NodeDesc(), name="m")], L, CameraDesc(), WorldDesc(ambient_energy=0.0), name="lights"))
    m = soup(5); m.positions[3] = [np.nan, 0, 0]; m.positions[10] = [np.inf, 1, 1]
    cases.append(SceneDesc([ModelDesc(m, NodeDesc(), frustum_culling=False, name="nan")], lights(), CameraDesc(), WorldDesc(ambient_energy=0.3), name="nan"))
    cases.append(SceneDesc([ModelDesc(soup(6), NodeDesc(scale=(1e18, 1e18, 1e18)), frustum_culling=False, name="huge")], lights(), CameraDesc(far=1e30),
                           WorldDesc(ambient_energy=0.3), name="huge"))
    cases.append(SceneDesc([ModelDesc(scenes.single_triangle(z=0.0), NodeDesc(), name="t")], lights(), CameraDesc(), WorldDesc(ambient_energy=0.5), name="one"))
    cases.append(SceneDesc([ModelDesc(scenes.single_triangle(z=0.0, flip=True), NodeDesc(), name="t")], lights(), CameraDesc(), WorldDesc(), name="culled"))
    par = NodeDesc(position=(1, 2, 3), rotation=scenes.random_rotation(rng), scale=(1.0, 2.0, 0.5))
    cam = CameraDesc(node=NodeDesc(position=(0, 0, 10), rotation=scenes.random_rotation(rng), parent=NodeDesc(rotation=scenes.random_rotation(rng))),
                     perspective=False, ortho_scale=30.0, vertex_snapping=4.0, render_depth=False)
    cases.append(SceneDesc([ModelDesc(soup(7), NodeDesc(position=(0, 1, 0), parent=par), frustum_culling=False, name="p")], lights(), cam,
                           WorldDesc(fog_mode=2, fog_color=(0.2, 0.3, 0.4, 1.0), ambient_energy=0.4), name="ortho"))
    cases.append(SceneDesc([ModelDesc(scenes.random_soup(60, 8, material=MaterialDesc(transparency_mode=3)), NodeDesc(), color=(1, 1, 1, 0.0), name="a0"),
                            ModelDesc(scenes.random_soup(60, 9, n_parts=2, material=MaterialDesc(visible=False)), NodeDesc(position=(1, 0, 0)), name="inv")],
                           lights(), CameraDesc(), WorldDesc(ambient_energy=0.4), name="early"))
    emitted = sum(check(c) for c in cases)
    assert emitted > 500

