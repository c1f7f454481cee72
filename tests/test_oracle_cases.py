"""Hand-derived micro-cases and properties of the oracle's render (the reference has no
golden vectors for this path, so these pin the restatement to what the cited source lines say)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import scenes
from oracle import binding as O
from tetra3d_b200 import workloads as W
from tetra3d_b200.scenedesc import (CameraDesc, MaterialDesc, MeshDesc, ModelDesc, NodeDesc, PartDesc, SceneDesc, WorldDesc)

F32 = np.float32


def one_model(mesh, **kw):
    return SceneDesc([ModelDesc(mesh, NodeDesc(), **kw)], [], CameraDesc(), WorldDesc())


def test_single_triangle_facing_camera_is_drawn():
    r = O.render_scene(one_model(scenes.single_triangle()))
    assert r.n_vertices == 3 and list(r.tri_ids) == [0]
    # screen mapping (camera.go:1674-1675): the triangle is centred, the apex above the base (dy smaller)
    xs, ys = r.color_list[:, 0], r.color_list[:, 1]
    assert xs[0] < 320 < xs[1] and abs(xs[2] - 320) < 1e-3 and ys[2] < ys[0]
    assert np.array_equal(r.index_list, [0, 1, 2])                       # camera.go:1816
    # depth colour = clamp((Z + margin) / camSpread) (camera.go:1771), alpha 1
    assert np.all(r.depth_list[:, 7] == 1) and np.all((r.depth_list[:, 4] > 0) & (r.depth_list[:, 4] < 1))


def test_single_triangle_facing_away_is_backface_culled():
    assert O.render_scene(one_model(scenes.single_triangle(flip=True))).n_vertices == 0      # model.go:902-925
    assert O.render_scene(one_model(scenes.single_triangle(flip=True, backface=False))).n_vertices == 3


def test_triangle_straddling_near_plane_is_kept_whole():
    # one corner behind the camera (w < 0 -> 1e-6, model.go:870-872); kept because another corner is in range
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 0, 14]], F32)
    mesh = MeshDesc(pos, np.tile([[0, 0, 1]], (3, 1)), np.zeros((3, 2)), np.array([[0, 1, 2]], np.uint32),
                    [PartDesc(0, 3, 0, 0, MaterialDesc(backface_culling=False))])
    r = O.render_scene(one_model(mesh, frustum_culling=False))
    assert r.n_vertices == 3 and np.isfinite(r.color_list[:, :2]).all()


def test_triangle_beyond_far_plane_is_dropped():
    sc = one_model(scenes.single_triangle(z=-200.0), frustum_culling=False)
    assert O.render_scene(sc).n_vertices == 0                            # model.go:877-886


def test_triangle_left_of_screen_is_dropped():
    m = scenes.single_triangle(); m.positions[:, 0] += 40                # all sx > 0.5 (model.go:891-894)
    assert O.render_scene(one_model(m, frustum_culling=False)).n_vertices == 0


def test_all_equal_depth_keeps_arrival_order():
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]] * 4, F32)
    idx = np.arange(12, dtype=np.uint32).reshape(4, 3)
    mesh = MeshDesc(pos, np.tile([[0, 0, 1]], (12, 1)), np.zeros((12, 2)), idx, [PartDesc(0, 12, 0, 3, MaterialDesc())])
    r = O.render_scene(one_model(mesh))
    assert list(r.tri_ids) == [0, 1, 2, 3]                               # rangeDiff == 0 -> one bin, arrival order


def test_back_to_front_and_front_to_back_are_bin_reversed_not_element_reversed():
    def scene(mode):
        mesh = scenes.random_soup(400, 7, n_parts=1, material=MaterialDesc(backface_culling=False, triangle_sort_mode=mode))
        return one_model(mesh, frustum_culling=False)
    b2f = O.render_scene(scene(0)); f2b = O.render_scene(scene(1)); none = O.render_scene(scene(2))
    assert sorted(b2f.tri_ids) == sorted(f2b.tri_ids) == sorted(none.tri_ids)
    assert list(none.tri_ids) == sorted(none.tri_ids)                    # unsorted = part order (sortingTriangles.go:52)
    mesh = scene(0).models[0].mesh
    cam = np.array([0, 0, 10], F32)
    centers = O.triangle_centers(mesh.positions, mesh.indices)
    d = ((cam - centers) ** 2).sum(1)
    mn, mx = b2f.renders[0]["min_depth"], b2f.renders[0]["max_depth"]
    bins = {t: O.bin_index(d[t], mn, mx, 0) for t in none.tri_ids}
    # BackToFront: bins descending, ascending triangle index inside a bin (sortingTriangles.go:101-111)
    seq = [(-bins[t], t) for t in b2f.tri_ids]
    assert seq == sorted(seq)
    seq = [(bins[t], t) for t in f2b.tri_ids]
    assert seq == sorted(seq)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 300), mode=st.integers(0, 2))
def test_draw_order_is_a_stable_sort_by_bin(seed, n, mode):
    mesh = scenes.random_soup(n, seed, n_parts=1, material=MaterialDesc(backface_culling=bool(seed % 2), triangle_sort_mode=mode))
    r = O.render_scene(SceneDesc([ModelDesc(mesh, NodeDesc(), frustum_culling=False)], [], CameraDesc(), WorldDesc()))
    ids = list(r.tri_ids)
    assert len(set(ids)) == len(ids) and r.n_vertices == 3 * len(ids)
    if not ids:
        return
    centers = O.triangle_centers(mesh.positions, mesh.indices)
    d = ((np.array([0, 0, 10], F32) - centers) ** 2).sum(1, dtype=F32)
    mn, mx = r.renders[0]["min_depth"], r.renders[0]["max_depth"]
    key = lambda t: (-O.bin_index(d[t], mn, mx, mode) if mode == 0 else O.bin_index(d[t], mn, mx, mode), t)
    assert ids == sorted(ids, key=key)


def test_ambient_light_colours_every_vertex():
    mesh = scenes.single_triangle()
    sc = SceneDesc([ModelDesc(mesh, NodeDesc(), color=(0.5, 1.0, 0.25, 1.0))], [], CameraDesc(),
                   WorldDesc(ambient_color=(1.0, 0.5, 0.25), ambient_energy=2.0))
    r = O.render_scene(sc)
    # colour = mpColor * (ambient colour * energy) (camera.go:1597-1601, 1738-1742; light.go:110-120)
    assert np.allclose(r.color_list[:, 4:8], [[0.5 * 2.0, 1.0 * 1.0, 0.25 * 0.5, 1.0]] * 3)


def test_shadeless_material_ignores_lights():
    mesh = scenes.single_triangle(); mesh.parts[0].material.shadeless = True
    r = O.render_scene(SceneDesc([ModelDesc(mesh, NodeDesc())], scenes.lights_mixed(np.random.default_rng(0)), CameraDesc(), WorldDesc()))
    assert np.all(r.color_list[:, 4:8] == 1.0)


def test_point_light_formula():
    """One vertex straight under a point light: diffuse = 1, factor = 1/(1+0.1 d^2) * 2 (light.go:245)."""
    mesh = scenes.single_triangle()
    from tetra3d_b200.scenedesc import LightDesc, LIGHT_POINT
    light = LightDesc(LIGHT_POINT, NodeDesc(position=(0.0, 1.0, 2.0)), color=(1.0, 0.5, 0.25), energy=3.0)
    r = O.render_scene(SceneDesc([ModelDesc(mesh, NodeDesc())], [light], CameraDesc(), WorldDesc()), capture=True)
    d2 = F32(4.0)
    f = F32(1.0) * (F32(1.0) / (F32(1.0) + F32(0.1) * d2)) * F32(2)
    apex = r.cap_light[2]       # vertex (0,1,0), normal +Z, light at distance 2 along +Z
    assert np.allclose(apex[:3], [1.0 * f * 3, 0.5 * f * 3, 0.25 * f * 3], rtol=1e-6) and apex[3] == 1


def test_transparent_pass_follows_solids_and_is_far_to_near():
    sc = scenes.transparent_scene(3)
    r = O.render_scene(sc)
    order = [sc.models[f["model"]] for f in r.flushes]
    from oracle.binding import orc_material  # noqa: F401
    def transparent(m):
        mat = m.mesh.parts[0].material
        if mat is None:
            return m.color[3] < 0.999
        return mat.transparency_mode == 3 or (mat.transparency_mode == 0 and mat.color[3] < 0.999) or (mat.transparency_mode != 1 and m.color[3] < 0.999)
    flags = [transparent(m) for m in order]
    assert flags == sorted(flags)                                        # solids first (camera.go:2057-2059)


def test_invisible_and_fully_transparent_models_emit_nothing():
    sc = scenes.transparent_scene(1)
    r = O.render_scene(sc)
    names = {sc.models[x["model"]].name: x for x in r.renders}
    assert "hidden" not in names and "hiddenmat" not in names
    assert names["ghost"]["vertex_count"] == 0                           # model.go:604


def test_dynamic_batch_children_share_one_flush():
    sc = scenes.batch_scene(1)
    r = O.render_scene(sc)
    owner_flush = [f for f in r.flushes if sc.models[f["model"]].name == "owner"]
    assert len(owner_flush) == 1
    kids = [x for x in r.renders if sc.models[x["model"]].name.startswith("c")]
    assert kids and all(k["flush"] == kids[0]["flush"] for k in kids)
    assert sum(k["vertex_count"] for k in kids) == owner_flush[0]["vertex_count"]
    # far -> near (camera.go:2078-2080)
    cam = np.array(sc.camera.node.position, F32)
    dist = [float(((cam - np.array(sc.models[k["model"]].node.position, F32)) ** 2).sum()) for k in kids]
    assert dist == sorted(dist, reverse=True)
    # the index list keeps counting across the children of one flush (camera.go:1816)
    f = owner_flush[0]
    assert np.array_equal(r.index_list[f["first_vertex"]:f["first_vertex"] + f["vertex_count"]], np.arange(f["vertex_count"], dtype=np.uint16))


def test_frustum_culling_rejects_models_behind_the_camera():
    sc = W.stress_scene(lit=False)
    r = O.render_scene(sc)
    assert r.debug["totalParts"] == 25 and 0 < r.debug["drawnParts"] < 25     # camera.go:1362-1368


def test_oracle_is_deterministic():
    sc = scenes.soup_scene(5)
    a = O.render_scene(sc); b = O.render_scene(sc)
    assert np.array_equal(a.color_list.view(np.uint32), b.color_list.view(np.uint32)) and np.array_equal(a.tri_ids, b.tri_ids)


def test_bake_lighting_adds_once_per_corner():
    """Model.BakeLighting (model.go:1246-1262) walks the part's triangles light by light: a vertex two triangles share receives a
    light's contribution twice.  A quad of two triangles sharing vertices 0 and 2, one ambient light of 0.25 and the world's ambient
    light at 0.5 (0.125, 0.25, 0.5 each): corners 1 and 3 get one helping of each, 0 and 2 two."""
    from tetra3d_b200 import workloads as W
    from tetra3d_b200.scenedesc import CameraDesc, LightDesc, ModelDesc, NodeDesc, SceneDesc, WorldDesc, LIGHT_AMBIENT
    mesh = W.plane_mesh(2, 2)
    amb = LightDesc(LIGHT_AMBIENT, color=(1.0, 0.5, 0.25), energy=0.25)
    desc = SceneDesc([ModelDesc(mesh, NodeDesc(), name="quad")], [amb], CameraDesc(node=NodeDesc(position=(0, 0, 5))),
                     WorldDesc(ambient_color=(0.25, 0.5, 1.0), ambient_energy=0.5), name="bake")
    rgb = O.bake_lighting(desc, desc.models[0], [amb])
    one = np.array([0.25 + 0.125, 0.125 + 0.25, 0.0625 + 0.5], np.float32)
    two = np.array([0.25 + 0.25 + 0.125 + 0.125, 0.125 + 0.125 + 0.25 + 0.25, 0.0625 + 0.0625 + 0.5 + 0.5], np.float32)
    assert np.array_equal(rgb[1], one) and np.array_equal(rgb[3], one)
    assert np.array_equal(rgb[0], two) and np.array_equal(rgb[2], two)
    # an invisible light takes no part; a Shadeless part is painted white (model.go:1241-1248)
    amb.visible = False
    assert np.array_equal(O.bake_lighting(desc, desc.models[0], [amb])[1], np.array([0.125, 0.25, 0.5], np.float32))
    mesh.parts[0].material.shadeless = True
    assert np.array_equal(O.bake_lighting(desc, desc.models[0], [amb]), np.ones((4, 3), np.float32))


def test_go_acos_restatement():
    """math32.Acos is float32(math.Acos(float64(x))); Go's amd64 Acos is pure Go (Pi/2 - Asin, Asin through satan).  The
    restatement agrees with glibc within 4 ulp of Pi/2 (two subtractions from Pi/2 cancel near |x| = 1) over [-1, 1] and on the special cases asin.go lists."""
    import math
    xs = np.concatenate([np.linspace(-1, 1, 4001), [0.7, 0.70000001, -0.7, 1e-300, -1e-300, 0.66]])     # (closer to 1, asin.go's 1 - x*x cancels: Go's own error, not the restatement's)
    for x in xs:
        g, r = O.go_acos(float(x)), math.acos(float(x))
        assert abs(g - r) <= 4 * np.spacing(math.pi / 2), (x, g, r)
    assert O.go_acos(1.0) == 0.0 and O.go_acos(0.0) == math.pi / 2 and math.isnan(O.go_acos(1.5))
