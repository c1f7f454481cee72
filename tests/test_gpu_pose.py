"""-m gpu: bone poses evaluated on the device (t3d_submit_poses -> k_pose_bones) against the oracle's restatement of
AnimationTrack sampling, Slerp, Node.Transform and boneInfluence -- bit for bit, matrices and rendered lists."""
import ctypes as C

import numpy as np
import pytest

import scenes
from compare import assert_parity, render_gpu
from oracle import binding as O
from tetra3d_b200 import workloads as W

pytestmark = pytest.mark.gpu


def assert_bones_match(desc, device):
    """Every posed model's matrices in the device's bone arena == orc_pose_bones, bit for bit."""
    scene, cam, objs = render_gpu(desc, device)
    offsets = cam.BoneOffsets()
    checked = 0
    for md in desc.models:
        if md.pose is None:
            continue
        want = O.pose_bones(md.pose.skeleton, md.pose.animation, md.pose.playhead, O.node_world_transform(md.pose.root))
        got = device.ReadBones(offsets[md.name], len(want))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (md.name, np.abs(got - want).max())
        checked += 1
    assert checked


@pytest.mark.parametrize("seed", range(6))
def test_posed_scene_parity(device, seed):
    # two skeletons (a random tree and a chain), two animations each, a rest-pose armature, explicit matrices and a static model
    desc = scenes.posed_scene(seed, n_bones=[5, 8, 2, 13, 5, 31][seed], chain=bool(seed % 2))
    assert_bones_match(desc, device)
    assert_parity(desc, device, check_capture=False)


@pytest.mark.parametrize("seed", range(20, 36))
def test_posed_fuzz(device, seed):
    # many more random skeletons / animations / playheads: every bone matrix must equal the oracle's bit for bit
    rng = np.random.default_rng(seed)
    desc = scenes.posed_scene(seed, n_models=int(rng.integers(2, 12)), n_bones=int(rng.integers(2, 64)), chain=bool(rng.integers(0, 2)))
    assert_bones_match(desc, device)


def test_playheads_on_and_around_keyframes(device):
    # chain_animation has keys every 0.25 s over [0, 2]: on a key, before the first, after the last, between
    desc = scenes.posed_scene(7, n_models=8, n_bones=5, playheads=[0.0, 0.25, 2.0, 2.5, -1.0, 1.0, 0.3333, 1.9999])
    assert_bones_match(desc, device)
    assert_parity(desc, device, check_capture=False)


def test_deep_and_wide_skeletons(device):
    # a 40-bone chain (40 hierarchy levels) and a skeleton of more than 256 bones (a block's threads loop over the bones)
    assert_bones_match(scenes.posed_scene(8, n_models=3, n_bones=40, chain=True), device)
    big = scenes.posed_scene(9, n_models=3, n_bones=300)
    assert_bones_match(big, device)
    assert_parity(big, device, check_capture=False)


def test_animated_crowd_parity(device):
    desc = W.skinned_crowd_animated_scene(n=200)
    assert_bones_match(desc, device)
    assert_parity(desc, device, check_capture=False)


def test_pose_api_errors(device):
    import tetra3d_b200
    lib = C.CDLL(tetra3d_b200.LIB_PATH)
    lib.t3d_last_error.restype = C.c_char_p; lib.t3d_last_error.argtypes = [C.c_void_p]
    ctx = C.c_void_p()
    assert lib.t3d_create(C.byref(ctx), 0) == 0
    f32p = C.POINTER(C.c_float)

    class Skel(C.Structure):
        _fields_ = [("n_bones", C.c_uint32), ("parent", C.POINTER(C.c_int32)), ("inverse_bind", f32p), ("rest_position", f32p),
                    ("rest_scale", f32p), ("rest_rotation", f32p)]

    class Track(C.Structure):
        _fields_ = [("first_key", C.c_uint32), ("n_keys", C.c_uint32), ("interpolation", C.c_uint32)]

    class Anim(C.Structure):
        _fields_ = [("n_bones", C.c_uint32), ("tracks", C.POINTER(Track)), ("n_keys", C.c_uint32), ("key_times", f32p), ("key_values", f32p)]

    n = 2
    eye = np.tile(np.eye(4, dtype=np.float32).reshape(16), n); zeros = np.zeros(3 * n, np.float32); ones = np.ones(3 * n, np.float32)
    fp = lambda a: a.ctypes.data_as(f32p)
    parent = np.array([-1, 0], np.int32)
    sk = Skel(n, parent.ctypes.data_as(C.POINTER(C.c_int32)), fp(eye), fp(zeros), fp(ones), fp(eye))
    assert lib.t3d_skeleton_upload(ctx, C.c_uint64(1), C.byref(sk)) == 0
    bad_parent = np.array([1, -1], np.int32)
    sk2 = Skel(n, bad_parent.ctypes.data_as(C.POINTER(C.c_int32)), fp(eye), fp(zeros), fp(ones), fp(eye))
    assert lib.t3d_skeleton_upload(ctx, C.c_uint64(2), C.byref(sk2)) == -1 and b"precede" in lib.t3d_last_error(ctx)
    tracks = (Track * (3 * n))()
    tracks[0] = Track(0, 3, 0)
    times = np.array([0.0, 2.0, 1.0], np.float32); vals = np.zeros(12, np.float32)
    an = Anim(n, tracks, 3, fp(times), fp(vals))
    assert lib.t3d_animation_upload(ctx, C.c_uint64(5), C.byref(an)) == -1 and b"non-decreasing" in lib.t3d_last_error(ctx)
    times[:] = [0.0, 1.0, 2.0]
    assert lib.t3d_animation_upload(ctx, C.c_uint64(5), C.byref(an)) == 0
    assert lib.t3d_animation_upload(ctx, C.c_uint64(0), C.byref(an)) == -1
    tracks[1] = Track(2, 5, 0)
    assert lib.t3d_animation_upload(ctx, C.c_uint64(6), C.byref(an)) == -1 and b"outside" in lib.t3d_last_error(ctx)
    off = C.c_uint32()
    pose = (C.c_float * 20)()
    assert lib.t3d_submit_poses(ctx, C.c_uint64(1), C.c_uint64(5), pose, 1, C.byref(off)) == -5     # no open frame
    assert lib.t3d_skeleton_release(ctx, C.c_uint64(77)) == -6 and lib.t3d_animation_release(ctx, C.c_uint64(77)) == -6
    assert lib.t3d_skeleton_release(ctx, C.c_uint64(1)) == 0 and lib.t3d_animation_release(ctx, C.c_uint64(5)) == 0
    lib.t3d_destroy(ctx)


def test_blending_and_relative_motion_are_refused(device):
    import tetra3d_b200 as t3d
    desc = scenes.posed_scene(1, n_models=2)
    scene, cam, objs = t3d.build_scene(desc, device)
    player = objs[id(desc.models[0].pose)]
    player.BlendTime = 0.2
    with pytest.raises(Exception, match="blending"):
        cam.Clear(); cam.RenderScene(scene)
