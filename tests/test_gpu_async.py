"""-m gpu: two frames in flight (t3d_set_async_download / Camera.AsyncDownload).  The lists of frame k travel to the host while
frame k + 1 is built and executed; what arrives must be exactly what the synchronous path returns, frame after frame, also when
consecutive frames differ in size (the two sets of device lists and host buffers grow independently)."""
import ctypes as C

import numpy as np
import pytest

import scenes
import tetra3d_b200 as t3d
from compare import bits
from oracle import binding as O
from tetra3d_b200 import workloads as W

pytestmark = pytest.mark.gpu


def sync_lists(desc, device):
    scene, cam, _ = t3d.build_scene(desc, device)
    cam.Clear(); cam.RenderScene(scene)
    return cam.ColorVertexList().copy(), cam.DepthVertexList().copy(), cam.TriangleIDs().copy()


def test_lists_of_frames_in_flight_equal_the_synchronous_ones(device):
    descs = [scenes.soup_scene(1), scenes.tile_boundary_scene(2), scenes.soup_scene(3, n_models=2, n_tris=90), scenes.skinned_scene(1),
             scenes.batch_scene(1), scenes.soup_scene(4, n_models=7, n_tris=2500), scenes.posed_scene(0), scenes.soup_scene(1)]
    want = [sync_lists(d, device) for d in descs]
    built = [t3d.build_scene(d, device) for d in descs]
    # one camera per scene, all on one device context: frames alternate between the context's two sets
    prev = None
    for k, (scene, cam, _) in enumerate(built):
        cam.AsyncDownload = True
        cam.Clear(); cam.RenderScene(scene)             # returns with frame k's copy under way
        if prev is not None:                            # frame k - 1: complete by now or after the wait
            pcam, pk = prev
            pcam.WaitDownload(1)
            col, dep = pcam.ColorVertexList(), pcam.DepthVertexList()      # that camera's last Render == frame k - 1
            assert np.array_equal(bits(col), bits(want[pk][0])) and np.array_equal(bits(dep), bits(want[pk][1])), pk
        prev = (cam, k)
    cam, k = prev
    assert np.array_equal(bits(cam.ColorVertexList()), bits(want[k][0])) and np.array_equal(cam.TriangleIDs(), want[k][2])


def test_one_camera_two_frames_in_flight_against_the_oracle(device):
    # the frame loop of bench.py's e2e arm: Render k, then consume k - 1 (frames_back = 1), a moving camera so that every frame differs
    desc = scenes.soup_scene(5, n_models=6, n_tris=1500)
    scene, cam, objs = t3d.build_scene(desc, device)
    cam.AsyncDownload = True
    refs = []
    for k in range(6):
        desc.camera.node.position = (0.3 * k, 0.1 * k, 10.0 - 0.5 * k)
        cam.SetLocalPosition(*desc.camera.node.position)
        refs.append(O.render_scene(desc))
        cam.Clear(); cam.RenderScene(scene)
        assert cam.VertexCount() == refs[k].n_vertices
        if k:
            col = cam.ColorVertexList(frames_back=1)
            assert len(col) == refs[k - 1].n_vertices
            assert np.array_equal(bits(col[:, :4]), bits(refs[k - 1].color_list[:, :4])), k      # positions + texture coordinates, bit for bit
            assert np.abs(col[:, 4:8] - refs[k - 1].color_list[:, 4:8]).max() <= 1.0 / 255
            dep = cam.DepthVertexList(frames_back=1)
            assert np.array_equal(bits(dep), bits(refs[k - 1].depth_list)), k
    col = cam.ColorVertexList()
    assert np.array_equal(bits(col[:, :4]), bits(refs[-1].color_list[:, :4]))
    # switching it off drains what is in flight and returns to the single set
    cam.AsyncDownload = False
    cam.Clear(); cam.RenderScene(scene)
    assert np.array_equal(bits(cam.ColorVertexList()[:, :4]), bits(refs[-1].color_list[:, :4]))


def test_async_download_at_headline_size(device):
    # the 1 M-triangle frame, ten frames back to back with two in flight: every frame's lists equal the synchronous frame's
    desc = W.million_triangle_scene()
    scene, cam, _ = t3d.build_scene(desc, device)
    cam.Clear(); cam.RenderScene(scene)
    want_c, want_d = cam.ColorVertexList().copy(), cam.DepthVertexList().copy()
    cam.AsyncDownload = True
    for k in range(10):
        cam.Clear(); cam.RenderScene(scene)
        if k:
            cam.WaitDownload(1)
    assert np.array_equal(bits(cam.ColorVertexList(frames_back=1)), bits(want_c))
    assert np.array_equal(bits(cam.ColorVertexList()), bits(want_c)) and np.array_equal(bits(cam.DepthVertexList()), bits(want_d))


def test_async_download_api(device):
    lib = C.CDLL(t3d.LIB_PATH)
    ctx = C.c_void_p()
    assert lib.t3d_create(C.byref(ctx), 0) == 0
    assert lib.t3d_set_async_download(ctx, 1) == 0
    assert lib.t3d_frame_download_wait(ctx, C.c_uint32(0)) == 0 and lib.t3d_frame_download_wait(ctx, C.c_uint32(1)) == 0     # nothing in flight
    assert lib.t3d_frame_download_wait(ctx, C.c_uint32(2)) == -1
    assert lib.t3d_set_async_download(ctx, 0) == 0
    assert lib.t3d_set_async_download(None, 1) == -1
    lib.t3d_destroy(ctx)


def test_depth_list_is_not_copied_when_render_depth_is_off(device):
    from tetra3d_b200.scenedesc import CameraDesc, WorldDesc
    desc = scenes.soup_scene(7, cam=CameraDesc(render_depth=False), world=WorldDesc(fog_mode=0, fog_color=(0.3, 0.2, 0.1, 1), fog_range=(0.1, 0.9)))
    scene, cam, _ = t3d.build_scene(desc, device)
    cam.Clear(); cam.RenderScene(scene)
    col, dep = cam.ColorVertexList().copy(), cam.DepthVertexList()
    _, d2h_full = device.FrameBytes()
    assert len(dep) == len(col) > 0
    cam.SkipUnusedDepthList = True
    cam.Clear(); cam.RenderScene(scene)
    _, d2h_skip = device.FrameBytes()
    assert np.array_equal(bits(cam.ColorVertexList()), bits(col)) and len(cam.DepthVertexList()) == 0
    assert d2h_full - d2h_skip == 48 * len(col)
    # a camera that renders depth is not affected by the switch
    desc2 = scenes.soup_scene(7)
    scene2, cam2, _ = t3d.build_scene(desc2, device)
    cam2.SkipUnusedDepthList = True
    cam2.Clear(); cam2.RenderScene(scene2)
    assert len(cam2.DepthVertexList()) == len(cam2.ColorVertexList()) > 0
