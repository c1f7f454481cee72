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
        p = md.pose
        if p.blend_ms is not None:
            want = O.pose_bones_blended(p.skeleton, p.blend_from, p.from_playhead, p.animation, p.playhead, p.blend_fraction(), O.node_world_transform(p.root))
        else:
            want = O.pose_bones(p.skeleton, p.animation, p.playhead, O.node_world_transform(p.root))
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


@pytest.mark.parametrize("seed", range(8))
def test_blended_poses_parity(device, seed):
    # players in the middle of a blend (animation.go:852-907): at the start (bp = 0), inside, one millisecond before the end, at the
    # end (bp = 1) and past it (clamped); from the skeleton's other animation or from nothing; next to players that do not blend
    blends = [[0, 250, 999, 1000, 1500, None, 333], [500], [1, None, 1000], [125, 875, None, None, 640]][seed % 4]
    desc = scenes.posed_scene(40 + seed, n_models=int([7, 4, 9, 12][seed % 4]), n_bones=[5, 8, 2, 13, 31, 64, 3, 300][seed], chain=bool(seed % 2), blends=blends)
    assert any(md.pose is not None and md.pose.blend_ms is not None for md in desc.models)
    assert_bones_match(desc, device)
    assert_parity(desc, device, check_capture=False)


def test_blended_poses_fuzz(device):
    for seed in range(60, 72):
        rng = np.random.default_rng(seed)
        blends = [None if rng.random() < 0.2 else int(rng.integers(0, 1200)) for _ in range(5)]
        if all(b is None for b in blends):
            blends[0] = 400
        desc = scenes.posed_scene(seed, n_models=int(rng.integers(2, 12)), n_bones=int(rng.integers(2, 64)), chain=bool(rng.integers(0, 2)), blends=blends)
        assert_bones_match(desc, device)


def test_blend_through_the_player_frame_by_frame(device):
    # one armature driven the way a game drives it: walk plays, Play(run) starts a 200 ms blend, frames every 16 ms of a stepped
    # clock -- every frame's bones against the oracle fed with what the reference's player would hold at that point
    import tetra3d_b200 as t3d
    desc = scenes.posed_scene(3, n_models=2, n_bones=6)
    pose = desc.models[0].pose
    other = desc.models[2 if len(desc.models) > 4 else 0].pose
    rng = np.random.default_rng(9)
    run = scenes.random_animation(rng, pose.skeleton, length=1.0)
    scene, cam, objs = t3d.build_scene(desc, device)
    player = objs[id(pose)]
    anim_run = t3d.build_animation(run)
    clock = [1000]
    player.SetClock(lambda: clock[0])
    player.BlendTime = 0.2
    root = O.node_world_transform(pose.root)
    name = desc.models[0].name

    def frame():
        cam.Clear(); cam.RenderScene(scene)
        return device.ReadBones(cam.BoneOffsets()[name], len(pose.skeleton.parent))

    player.Update(0.016)                                    # walk, BlendTime > 0: currentProperties is being tracked now
    walk_ph = player.SampledPlayhead()
    assert not player.Blending()
    got = frame()
    assert np.array_equal(got.view(np.uint32), O.pose_bones(pose.skeleton, pose.animation, walk_ph, root).view(np.uint32))
    player.Play(anim_run)
    seen_end = False
    for step in range(20):
        clock[0] += 16
        player.Update(0.016)
        ph = player.SampledPlayhead()
        got = frame()
        if player.Blending():
            bp = player.BlendFraction()
            assert bp == float(min(np.float32(16 * (step + 1)) / (np.float32(0.2) * np.float32(1000.0)), np.float32(1.0)))
            assert player.BlendFromPlayhead() == walk_ph
            want = O.pose_bones_blended(pose.skeleton, pose.animation, walk_ph, run, ph, bp, root)
            seen_end = seen_end or bp == 1.0
        else:
            assert seen_end                                  # the blend is over: the frame after bp reached 1 is a plain pose again
            want = O.pose_bones(pose.skeleton, run, ph, root)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), step
    assert seen_end and not player.Blending()


def test_blended_pose_api_errors(device):
    import tetra3d_b200
    lib = C.CDLL(tetra3d_b200.LIB_PATH)
    ctx = C.c_void_p()
    assert lib.t3d_create(C.byref(ctx), 0) == 0
    off = C.c_uint32()
    pose = (C.c_float * 20)()
    assert lib.t3d_submit_poses_blended(ctx, C.c_uint64(1), C.c_uint64(2), C.c_uint64(3), pose, 1, C.byref(off)) == -5     # no open frame
    assert lib.t3d_submit_poses_blended(ctx, C.c_uint64(1), C.c_uint64(2), C.c_uint64(3), pose, 1, None) == -1
    lib.t3d_destroy(ctx)


def test_interrupted_blends_through_the_player(device):
    # Play() in the middle of unfinished blends, again and again: every frame's bones against the oracle's fold of the chain the
    # player reports (t3d_submit_pose_chains / k_pose_chain for three and four levels; two levels and one go the usual ways)
    import tetra3d_b200 as t3d
    for seed in range(4):
        desc = scenes.posed_scene(10 + seed, n_models=2, n_bones=[6, 11, 3, 40][seed], chain=bool(seed % 2))
        pose = desc.models[0].pose
        rng = np.random.default_rng(70 + seed)
        anim_descs = [pose.animation] + [scenes.random_animation(rng, pose.skeleton, length=30.0) for _ in range(3)]
        scene, cam, objs = t3d.build_scene(desc, device)
        player = objs[id(pose)]
        host = [player.SampledChain()[-1][0]] + [t3d.build_animation(a) for a in anim_descs[1:]]
        clock = [0]
        player.SetClock(lambda: clock[0])
        player.BlendTime = 0.4
        root = O.node_world_transform(pose.root)
        name = desc.models[0].name
        depths = set()
        cur = 0
        for step in range(36):
            if step % 3 == 2:                       # a new animation every third frame: 48 ms into a 400 ms blend
                cur = (cur + 1 + int(rng.integers(0, 3))) % 4
                player.Play(host[cur])
            clock[0] += 16
            player.Update(0.016)
            chain = player.SampledChain()
            if len(chain) > 4:
                with pytest.raises(Exception, match="T3D_POSE_CHAIN_MAX"):
                    cam.Clear(); cam.RenderScene(scene)
                clock[0] += 1000                    # let that blend run out: one more frame at bp = 1, then a plain pose again
                player.Update(0.016); player.Update(0.016)
                chain = player.SampledChain()
                assert len(chain) == 1
            depths.add(len(chain))
            cam.Clear(); cam.RenderScene(scene)
            got = device.ReadBones(cam.BoneOffsets()[name], len(pose.skeleton.parent))
            anims = [anim_descs[[i for i, h in enumerate(host) if h is a][0]] for a, _, _ in chain]
            want = O.pose_bones_chain(pose.skeleton, anims, [p for _, p, _ in chain], [b for _, _, b in chain], root)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (seed, step, len(chain))
        assert {1, 2, 3, 4} <= depths, depths


def test_pose_chain_api_errors(device):
    import tetra3d_b200
    lib = C.CDLL(tetra3d_b200.LIB_PATH)
    ctx = C.c_void_p()
    assert lib.t3d_create(C.byref(ctx), 0) == 0
    off = C.c_uint32()
    pose = (C.c_float * 24)()
    ids = (C.c_uint64 * 4)(1, 2, 3, 4)
    assert lib.t3d_submit_pose_chains(ctx, C.c_uint64(1), ids, 3, pose, 1, C.byref(off)) == -5      # no open frame
    assert lib.t3d_submit_pose_chains(ctx, C.c_uint64(1), None, 3, pose, 1, C.byref(off)) == -1
    lib.t3d_destroy(ctx)


