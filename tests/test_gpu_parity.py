"""-m gpu: the CUDA path (through the host mirror and the C ABI) against the CPU oracle."""
import numpy as np
import pytest

import scenes
from compare import assert_parity, render_gpu, bits
from oracle import binding as O
from tetra3d_b200 import workloads as W
from tetra3d_b200.scenedesc import CameraDesc, MaterialDesc, ModelDesc, NodeDesc, SceneDesc, WorldDesc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(scenes.ALL_SMALL))
def test_small_scene_parity(device, name):
    assert_parity(scenes.ALL_SMALL[name](), device)


def test_stress_scene_parity(device):
    # configs[0]: examples/stress as shipped (shadeless, frustum culling on)
    assert_parity(W.stress_scene(lit=False), device, check_capture=False)


def test_stress_lit_parity(device):
    assert_parity(W.stress_scene(lit=True), device, check_capture=False)


def test_million_triangles_reduced_parity(device):
    # configs[1] at 1/16 size: the oracle finishes in well under a second
    assert_parity(W.million_triangle_scene(nx=250, nz=125), device)


def test_skinned_crowd_reduced_parity(device):
    assert_parity(W.skinned_crowd_scene(n=64), device, check_capture=False)


def test_particles_reduced_parity(device):
    assert_parity(W.particles_scene(n=600, batch_cap=256), device, check_capture=False)


def test_instanced_reduced_parity(device):
    assert_parity(W.instanced_scene(n=12, subdivisions=2, box=(30.0, 10.0, 30.0)), device)


def test_tile_boundaries_parity(device):
    # draws of 1 .. 8200 triangles around the 1024-slot tile: packed tiles of several draws, exactly full tiles,
    # one-draw tiles, draws spanning 2..9 tiles, all in one frame
    assert_parity(scenes.tile_boundary_scene(31), device)
    assert_parity(scenes.tile_boundary_scene(32, sizes=(1024,) * 3 + (1,) * 40 + (4097, 2, 2, 1030)), device, check_capture=False)


def test_more_than_one_scan_chunk_parity(device):
    """More than 4096 triangle tiles in a frame with large (> 8192-triangle) draws: the tile scan runs as its own
    multi-block kernel and the emit adds the earlier chunks' totals.  480 instances of a 9000-triangle soup."""
    rng = np.random.default_rng(61)
    mesh = scenes.random_soup(9000, 61, n_parts=1, extent=3.0)
    models = [ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-10, 10, 3)), rotation=scenes.random_rotation(rng)),
                        frustum_culling=False, name="big%d" % i) for i in range(480)]
    sc = SceneDesc(models, scenes.lights_mixed(rng), CameraDesc(node=NodeDesc(position=(0.0, 2.0, 30.0))), WorldDesc(), name="chunks")
    assert_parity(sc, device, check_capture=False)


def test_huge_draw_parity(device):
    # one MeshPart of more than 64 tiles (66 000 triangles): its histogram scan runs in k_hist_scan_huge; next to it a
    # large draw below that limit and a packed one
    assert_parity(scenes.tile_boundary_scene(35, sizes=(66000, 9000, 300, 131073)), device, check_capture=False)


def test_mid_size_draws_parity(device):
    # draws of 9 .. 64 tiles (segmented sort, histogram scan by the draw's last block), a last tile of one triangle, the
    # largest such draw (65 536) next to the smallest one whose scan runs in k_hist_scan_huge (65 537)
    assert_parity(scenes.tile_boundary_scene(41, sizes=(8193, 9217, 15360, 16385, 20000, 700, 33000)), device, check_capture=False)
    assert_parity(scenes.tile_boundary_scene(42, sizes=(65536, 65537, 12000, 8192, 40001)), device, check_capture=False)


def test_mid_size_draws_skinned_and_culled_parity(device):
    # a skinned draw above 8192 triangles (depth from the skinned positions) and a large draw with no survivors
    sc = scenes.skinned_scene(7, n=2, rings=60, sides=80)
    assert_parity(sc, device, check_capture=False)
    sc = scenes.tile_boundary_scene(43, sizes=(10000, 9000))
    sc.models[0].node.position = (0.0, 0.0, 500.0)        # behind the camera: everything culled
    assert_parity(sc, device, check_capture=False)


@pytest.mark.parametrize("seed", range(8))
def test_volume_cell_lights_fuzz(device, seed):
    # LightVolumeCell.useLights: random cells with 1-3 lights each, the three shading modes, with and without skinned models
    assert_parity(scenes.volume_cell_lights_scene(40 + seed, mode=seed % 3, skinned=bool(seed % 2)), device)


def test_volume_cell_lights_many(device):
    # 40 lights behind one volume: the expanded light list outgrows its first allocation
    assert_parity(scenes.volume_cell_lights_scene(60, mode=0, n_extra=40), device)


def test_staged_and_chunked_uniform_upload_parity(monkeypatch):
    # frames of a megabyte and more are copied to the device by DMA before k_prep_draws reads them, and very large ones in
    # slices that overlap the kernel: both forced onto small frames here (run-length records, plain records, bones, poses)
    import tetra3d_b200 as t3d
    monkeypatch.setenv("T3D_DMA_THRESHOLD", "0")
    monkeypatch.setenv("T3D_PREP_CHUNK_BLOCKS", "2")
    dev = t3d.GeometryDevice(0)
    assert_parity(W.particles_scene(n=3000, batch_cap=4096), dev, check_capture=False)      # run-length records, 47 blocks -> 24 slices
    assert_parity(scenes.tile_boundary_scene(33, sizes=(1, 2) * 700), dev, check_capture=False)
    assert_parity(W.skinned_crowd_scene(n=500), dev, check_capture=False)
    assert_parity(W.skinned_crowd_animated_scene(n=500), dev, check_capture=False)
    assert_parity(scenes.fuzz_scene(3), dev, check_capture=False)                            # a small frame: staged, one slice


def test_non_affine_bone_matrices_parity(device):
    # bone matrices normally travel as 12 floats (column 3 is exactly 0,0,0,1); one that is not switches the frame to
    # the full 16-float upload
    sc = scenes.skinned_scene(3)
    bi = np.array(sc.models[1].bone_influences, np.float32).reshape(-1, 4, 4).copy()
    bi[2, 1, 3] = 0.015625; bi[0, 3, 3] = 1.0009765625
    sc.models[1].bone_influences = bi
    assert_parity(sc, device)
    sc2 = scenes.skinned_scene(3)
    bi = np.array(sc2.models[0].bone_influences, np.float32).reshape(-1, 4, 4).copy()
    bi[1, 0, 3] = -0.0                     # a negative zero must survive too
    sc2.models[0].bone_influences = bi
    assert_parity(sc2, device)


def test_full_upload_paths_parity(device):
    # the 16-float transform / bone upload paths (taken when a matrix of the frame is not affine), forced on
    device.ForceFullUpload(True)
    try:
        assert_parity(scenes.skinned_scene(5), device)
        assert_parity(scenes.tile_boundary_scene(36), device, check_capture=False)
        assert_parity(W.particles_scene(n=600, batch_cap=256), device, check_capture=False)
    finally:
        device.ForceFullUpload(False)


def test_many_tiny_draws_parity(device):
    # > 1024 draws of one or two triangles: packed tiles are cut by the draw-count limit, not by triangles
    sc = scenes.tile_boundary_scene(33, sizes=(1, 2) * 700)
    assert_parity(sc, device, check_capture=False)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_parity(device, seed):
    assert_parity(scenes.fuzz_scene(seed), device, check_capture=False)


def test_stage_timing_mode_gives_the_same_lists(device):
    """t3d_set_stage_timings(1) serialises the kernels with events instead of chaining them with programmatic
    dependent launches: same lists, and per-stage times that add up to no more than the frame."""
    sc = scenes.tile_boundary_scene(34)
    scene, cam, objs = render_gpu(sc, device)
    col0 = cam.ColorVertexList().copy(); dep0 = cam.DepthVertexList().copy(); tri0 = cam.TriangleIDs().copy()
    t0 = cam.Timings()
    assert t0["total_ms"] > 0 and t0["vertex_ms"] == 0
    device.SetStageTimings(True)
    try:
        cam.Clear(); cam.RenderScene(scene)
        t1 = cam.Timings()
        assert np.array_equal(bits(col0), bits(cam.ColorVertexList())) and np.array_equal(bits(dep0), bits(cam.DepthVertexList()))
        assert np.array_equal(tri0, cam.TriangleIDs())
        stages = t1["prep_ms"] + t1["vertex_ms"] + t1["cull_ms"] + t1["sort_ms"] + t1["emit_ms"]
        assert t1["vertex_ms"] > 0 and t1["emit_ms"] > 0 and stages <= t1["total_ms"] * 1.001
    finally:
        device.SetStageTimings(False)


def test_external_output_target(device):
    """t3d_set_output_target: the emit kernel writes the lists into caller-provided device memory (what the
    multi-GPU path maps from another rank); the host copy made from there is bit-identical."""
    sc = scenes.soup_scene(41)
    scene, cam, objs = render_gpu(sc, device)
    col0 = cam.ColorVertexList().copy(); dep0 = cam.DepthVertexList().copy()
    cap = 3 * sc.total_triangles()
    pc, _ = device.SharedAlloc(cap * 48)
    pd, _ = device.SharedAlloc(cap * 48)
    device.SetOutputTarget(pc, pd, 0, cap)
    try:
        cam.DownloadResults = False
        cam.Clear(); cam.RenderScene(scene)
        assert cam.DevicePointers()["color"] == pc and cam.DevicePointers()["depth"] == pd
        cam.DownloadResults = True
        cam.Clear(); cam.RenderScene(scene)
        assert np.array_equal(bits(col0), bits(cam.ColorVertexList())) and np.array_equal(bits(dep0), bits(cam.DepthVertexList()))
        device.SetOutputTarget(pc, pd, 0, 3)         # too small for the frame: refused, not overrun
        cam.Clear()
        import tetra3d_b200 as t3d
        with pytest.raises(t3d.T3DError):
            cam.RenderScene(scene)
    finally:
        device.SetOutputTarget(0, 0, 0, 0)
    cam.Clear(); cam.RenderScene(scene)
    assert np.array_equal(bits(col0), bits(cam.ColorVertexList()))


def test_everything_culled(device):
    sc = scenes.soup_scene(21)
    sc.camera.node.position = (0.0, 0.0, 500.0)      # everything beyond the far plane
    for m in sc.models:
        m.frustum_culling = False
    cam, ref = assert_parity(sc, device)
    assert cam.VertexCount() == 0 and ref.n_vertices == 0


def test_empty_scene(device):
    from tetra3d_b200.scenedesc import SceneDesc
    cam, ref = assert_parity(SceneDesc([]), device)
    assert cam.VertexCount() == 0


def test_equal_depth_range_zero(device):
    # every surviving triangle at the same depth: rangeDiff == 0 branch (sortingTriangles.go:43-46)
    from tetra3d_b200.scenedesc import MeshDesc, PartDesc, MaterialDesc, ModelDesc, NodeDesc, SceneDesc
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]] * 5, np.float32)
    idx = np.arange(15, dtype=np.uint32).reshape(5, 3)
    mesh = MeshDesc(pos, np.tile([[0, 0, 1]], (15, 1)), np.zeros((15, 2)), idx, [PartDesc(0, 15, 0, 4, MaterialDesc())])
    assert_parity(SceneDesc([ModelDesc(mesh, NodeDesc())]), device)


def test_nan_depth_dropped(device):
    # a NaN vertex makes the depth NaN: the triangle must be dropped (model.go:957-960)
    sc = scenes.soup_scene(22, n_models=1, lit=False)
    sc.models[0].mesh.positions[5] = np.nan
    assert_parity(sc, device, exact=False)


def test_full_size_million_triangles_properties(device):
    """configs[1] at full size: size-independent properties instead of the oracle."""
    sc = W.million_triangle_scene()
    scene, cam, objs = render_gpu(sc, device)
    renders = cam.Renders()
    assert sum(r["total_tris"] for r in renders) == 1_000_000
    tri = cam.TriangleIDs()
    assert len(tri) * 3 == cam.VertexCount()
    mesh = sc.models[0].mesh
    centers = O.triangle_centers(mesh.positions, mesh.indices)
    # recompute every surviving triangle's bin from the reported min/max and check the draw order:
    # bins non-increasing (BackToFront) and triangle index ascending inside a bin
    for r in renders:
        n = r["vertex_count"] // 3
        ids = tri[r["first_vertex"] // 3: r["first_vertex"] // 3 + n].astype(np.int64)
        part = mesh.parts[r["part"]]
        assert ((ids >= part.tri_start) & (ids <= part.tri_end)).all()
        assert len(np.unique(ids)) == n                       # a triangle is emitted at most once
    # idempotence: a second frame gives identical lists
    col1 = cam.ColorVertexList().copy(); tri1 = tri.copy()
    cam.Clear(); cam.RenderScene(scene)
    assert np.array_equal(tri1, cam.TriangleIDs()) and np.array_equal(bits(col1), bits(cam.ColorVertexList()))


def test_full_size_million_triangles_vs_oracle(device):
    # the oracle needs ~0.3 s for this frame, so the headline workload is also checked bit for bit
    assert_parity(W.million_triangle_scene(), device, check_capture=False)


@pytest.mark.parametrize("name", ["skinned_crowd", "skinned_crowd_animated", "instanced_shard", "particles_200k", "stress_lit", "million_one_part"])
def test_full_size_workloads_vs_oracle(device, name):
    """The other BASELINE configs at (or near) full size, bit for bit against the oracle: configs[2] whole (10 000
    skinned characters, 7.68 M triangles), an eighth of configs[4] (the share of one of 8 GPUs: 12.5 M triangles),
    a tenth of configs[3] (200 000 particle draws), configs[0] lit, and configs[1]'s mesh as one 1 M-triangle part."""
    assert_parity(W.WORKLOADS[name](), device, check_capture=False)


def test_full_size_particles_vs_oracle(device):
    """configs[3] at full size: 2 000 000 billboard particle draws (4 M triangles) in 184 dynamic batches, bit for bit against
    the oracle (the scene is shadeless, so one lighting mode covers it; ~3 minutes, most of it building 2 M Python objects)."""
    assert_parity(W.WORKLOADS["particles"](), device, check_capture=False, modes=("fast",))


def test_unsupported_host_callbacks_fail_loudly(device):
    import tetra3d_b200 as t3d
    sc = scenes.soup_scene(23, n_models=1)
    scene, cam, objs = t3d.build_scene(sc, device)
    objs[id(sc.models[0])].HasVertexTransformFunction = True
    cam.Clear()
    with pytest.raises(t3d.T3DError):
        cam.RenderScene(scene)


def test_multi_gpu_sharded_gather_matches_oracle():
    """Runs only on boxes with >= 2 GPUs: tests/run_multigpu_parity.py under torchrun."""
    import os, subprocess, sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "run_multigpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


@pytest.mark.parametrize("mode", [0, 2])
def test_both_sort_paths_of_large_draws(device, mode):
    """Draws of 9 .. 64 tiles are sorted either by the three-kernel segmented sort or by the single persistent kernel (k_cull_sort:
    tiles in ticket order, two draw-wide arrivals); which one a frame gets is a heuristic (t3d_debug_set_fused_sort), and both
    must give the oracle's lists: the 1 M-triangle frame (1000 tiles: more than are resident at once, so tickets wrap), the lit
    stress scene, a frame that mixes them with draws of more than 64 tiles, and all three sort modes."""
    try:
        device.SetFusedSort(mode)
        assert_parity(W.million_triangle_scene(), device, check_capture=False, modes=("fast",))
        assert_parity(W.stress_scene(lit=True), device, check_capture=False, modes=("fast",))
        rng = np.random.default_rng(77)
        models = []
        for k, n in enumerate((9000, 70000, 12000, 30000, 8193, 66000, 20000)):
            mesh = scenes.random_soup(n, 700 + k, n_parts=1, extent=3.0, material=MaterialDesc(triangle_sort_mode=k % 3, backface_culling=bool(k % 2)))
            models.append(ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-6, 6, 3)), rotation=scenes.random_rotation(rng)),
                                    frustum_culling=False, name="l%d" % k))
        sc = SceneDesc(models, scenes.lights_mixed(rng), CameraDesc(node=NodeDesc(position=(0.0, 2.0, 25.0))), WorldDesc(), name="largemix")
        assert_parity(sc, device, check_capture=False, modes=("fast",))
    finally:
        device.SetFusedSort(1)


@pytest.mark.parametrize("n,perspective,parent,billboard", [(60, True, False, True), (60, False, True, False), (3000, True, False, True), (2600, True, True, False), (2200, False, False, True)])
def test_device_frustum_culling_of_batch_children(device, n, perspective, parent, billboard):
    """DynamicBatch children that are instances of one another: their ModelInFrustum test (camera.go:2088-2093: the culling sphere
    of model.go:142-182 against Camera.SphereInFrustum) runs on the device -- in k_prep_draws for frames shipped as runs (>= 1024
    draws), on the host through the same function otherwise.  The oracle tests every child on the CPU: same render() calls, same
    lists; and the host-tested path (DeviceFrustumCulling off) gives the same frame."""
    import tetra3d_b200 as t3d
    desc = scenes.batch_frustum_scene(n, n_children=n, perspective=perspective, parent=parent, billboard=billboard)
    cam, ref = assert_parity(desc, device, check_capture=False, modes=("fast",))
    col1, n1 = cam.ColorVertexList().copy(), len(cam.Renders())      # (the lists are the library's: the next frame overwrites them)
    assert 0 < len(ref.renders) < n, "the scene must have children on both sides of the frustum (%d of %d rendered)" % (len(ref.renders), n)
    scene, cam2, objs = t3d.build_scene(desc, device)
    cam2.DeviceFrustumCulling = False
    cam2.Clear(); cam2.RenderScene(scene)
    assert np.array_equal(bits(col1), bits(cam2.ColorVertexList())) and n1 == len(cam2.Renders())
