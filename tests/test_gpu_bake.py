"""-m gpu: Model.BakeLighting (model.go:1203-1268) on the device against the oracle's restatement, bit for bit.

The reference's bake walks light by light over the part's triangles and adds a light's contribution to a vertex once per
corner that refers to it, so a vertex of valence k receives k additions per light: the device repeats them (t3d_frame_set_bake_lighting)."""
import numpy as np
import pytest

import scenes
from compare import same_bits
from oracle import binding as O

pytestmark = pytest.mark.gpu


def _bake_both(desc, device, model_index, light_descs):
    import tetra3d_b200 as t3d
    scene, cam, objs = t3d.build_scene(desc, device)
    md = desc.models[model_index]
    model = objs[id(md)]
    gpu = model.BakeLighting(0, cam, scene, [objs[id(l)] for l in light_descs])
    ref = O.bake_lighting(desc, md, light_descs)
    return gpu, ref, md


@pytest.mark.parametrize("seed,ranged", [(1, False), (2, True), (3, False)])
def test_bake_lighting_static_soup(device, seed, ranged):
    desc = scenes.soup_scene(seed, n_models=3, n_tris=500, ranged=ranged)
    for mi in range(len(desc.models)):
        gpu, ref, md = _bake_both(desc, device, mi, list(desc.lights))
        assert gpu.shape == (md.mesh.n_verts, 4)
        assert same_bits(gpu[:, :3], ref), "model %d: baked colours differ (max %g)" % (mi, float(np.nanmax(np.abs(gpu[:, :3] - ref))))
        assert np.all(gpu[:, 3] == 1.0)
        lit = ref[np.isfinite(ref).all(axis=1)]
        assert lit.max() > 0.0          # the scene's lights reach the mesh: not a comparison of zeros


def test_bake_lighting_subset_of_lights_and_parent(device):
    desc = scenes.soup_scene(7, n_models=4, n_tris=300, parent=True)
    lights = list(desc.lights)[::2]
    gpu, ref, md = _bake_both(desc, device, 0, lights)
    assert same_bits(gpu[:, :3], ref)
    gpu0, ref0, _ = _bake_both(desc, device, 0, [])            # only the world's ambient light
    assert same_bits(gpu0[:, :3], ref0)


def test_bake_lighting_skinned(device):
    desc = scenes.skinned_scene(2, n=3)
    gpu, ref, md = _bake_both(desc, device, 1, list(desc.lights))
    assert same_bits(gpu[:, :3], ref)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bake_lighting_light_volume(device, mode):
    desc = scenes.volume_scene(3, mode=mode)
    for mi in range(min(2, len(desc.models))):
        gpu, ref, md = _bake_both(desc, device, mi, list(desc.lights))
        assert same_bits(gpu[:, :3], ref)


def test_bake_frame_emits_no_lists(device):
    import tetra3d_b200 as t3d
    desc = scenes.soup_scene(1, n_models=1, n_tris=100)
    scene, cam, objs = t3d.build_scene(desc, device)
    objs[id(desc.models[0])].BakeLighting(0, cam, scene, [])
    cam.Clear(); cam.RenderScene(scene)                        # an ordinary frame afterwards is unaffected
    from compare import assert_parity
    assert_parity(desc, device)


# ---- Model.BakeAO, same-model pass (model.go:1040-1124) ---------------------------------------------------------------
def _box_room(seed=0):
    """A mesh with creases: a few cubes' worth of faces sharing corner positions but not vertices (cube meshes duplicate
    vertices per face), plus a random soup part that shares nothing."""
    from tetra3d_b200 import workloads as W
    from tetra3d_b200.scenedesc import CameraDesc, ModelDesc, NodeDesc, SceneDesc, WorldDesc
    mesh = W.cube_mesh(2, 2, 2)
    desc = SceneDesc([ModelDesc(mesh, NodeDesc(), name="cube")], [], CameraDesc(node=NodeDesc(position=(0, 0, 8))), WorldDesc(), name="ao%d" % seed)
    return desc


@pytest.mark.parametrize("angle,parts,verts", [(np.radians(60.0), None, None), (np.radians(100.0), None, None), (0.1, [0], None), (np.radians(60.0), None, [0, 1, 2, 5, 7, 11])])
def test_bake_ao_cube(device, angle, parts, verts):
    import tetra3d_b200 as t3d
    desc = _box_room()
    scene, cam, objs = t3d.build_scene(desc, device)
    md = desc.models[0]
    model = objs[id(md)]
    normals, spans = objs[id(md.mesh)].TriangleNormalsAndSpans()
    oc = [0.4, 0.3, 0.2, 1.0]
    gpu = model.BakeAO(device, 0, float(angle), oc, parts or [], verts or [])
    ref = O.bake_ao(desc, md, normals, spans, float(angle), oc, target_parts=parts, target_vertices=verts)
    assert gpu.shape == ref.shape and same_bits(gpu, ref)
    if angle < np.radians(95.0) and verts is None:
        assert np.any(np.all(np.isclose(ref, oc), axis=1))       # the cube's edges are 90-degree creases: corners get occluded


def test_bake_ao_soup_and_crowd_mesh(device):
    import tetra3d_b200 as t3d
    for desc, mi in ((scenes.soup_scene(4, n_models=1, n_tris=700), 0), (scenes.skinned_scene(1, n=2), 0)):
        scene, cam, objs = t3d.build_scene(desc, device)
        md = desc.models[mi]
        normals, spans = objs[id(md.mesh)].TriangleNormalsAndSpans()
        gpu = objs[id(md)].BakeAO(device, 0, 0.5, [0.1, 0.2, 0.3, 0.9], [], [])
        ref = O.bake_ao(desc, md, normals, spans, 0.5, [0.1, 0.2, 0.3, 0.9])
        assert same_bits(gpu, ref)
