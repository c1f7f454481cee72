"""CPU: the oracle's restatement of the bone-pose step (animation.go, quaternion.go, node.go:572-594) and the host
mirror's AnimationPlayer playhead logic (animation.go:724-806).  No GPU."""
import math

import numpy as np
import pytest

import scenes
from oracle import binding as O
from tetra3d_b200 import workloads as W
from tetra3d_b200.scenedesc import (INTERPOLATION_CONSTANT, TRACK_POSITION, TRACK_ROTATION, TRACK_SCALE, AnimationDesc, SkeletonDesc, TrackDesc)

F32 = np.float32
I16 = np.eye(4, dtype=F32).reshape(16)


def ulps(a: float, b: float) -> float:
    return abs(a - b) / max(math.ulp(b), 5e-324)


def test_go_sin_and_atan2_restatements_are_correct_to_one_ulp():
    # Go's math.Sin / math.Atan2 are Cephes-derived and accurate to < 1 ulp; a mistyped coefficient would show up here
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.uniform(0, math.pi, 4000), rng.uniform(-40, 40, 2000), [0.0, math.pi / 4, math.pi / 2, 1e-9, 3.0]]):
        assert ulps(O.go_sin(float(x)), math.sin(float(x))) <= 1.0, x
    for y, x in rng.uniform(-3, 3, (4000, 2)):
        assert ulps(O.go_atan2(float(y), float(x)), math.atan2(float(y), float(x))) <= 1.0, (y, x)
    # special cases of atan2.go
    assert O.go_atan2(0.0, 1.0) == 0.0 and math.copysign(1, O.go_atan2(-0.0, 1.0)) == -1
    assert O.go_atan2(0.0, -1.0) == math.pi and O.go_atan2(-0.0, -1.0) == -math.pi
    assert O.go_atan2(1.0, 0.0) == math.pi / 2 and O.go_atan2(-1.0, 0.0) == -math.pi / 2
    assert O.go_atan2(math.inf, math.inf) == math.pi / 4 and O.go_atan2(math.inf, -math.inf) == 3 * math.pi / 4
    assert math.isnan(O.go_atan2(math.nan, 1.0)) and math.isnan(O.go_sin(math.inf)) and O.go_sin(0.0) == 0.0


def one_bone(tracks):
    sk = SkeletonDesc(["b"], [-1], I16, [[0, 0, 0]], [[1, 1, 1]], I16)
    return sk, AnimationDesc("a", 2.0, {"b": tracks})


def test_track_sampling_follows_value_as_vector():
    # animation.go:63-112: before the first key / after the last key / on a key / between keys / constant interpolation
    times = [0.5, 1.0, 2.0]
    vals = [[1, 0, 0], [3, 2, 0], [3, 2, 8]]
    sk, an = one_bone({TRACK_POSITION: TrackDesc(times, vals)})
    pos = lambda t: O.pose_bones(sk, an, t, I16)[0].reshape(4, 4)[3, :3]
    assert np.array_equal(pos(-1.0), [1, 0, 0]) and np.array_equal(pos(0.5), [1, 0, 0])
    assert np.array_equal(pos(2.0), [3, 2, 8]) and np.array_equal(pos(9.0), [3, 2, 8])
    assert np.array_equal(pos(1.0), [3, 2, 0])
    t = F32(0.75)
    pct = (t - F32(0.5)) / (F32(1.0) - F32(0.5))
    assert np.array_equal(pos(float(t)), [F32(1) + (F32(3) - F32(1)) * pct, F32(0) + (F32(2) - F32(0)) * pct, 0])
    sk, an = one_bone({TRACK_POSITION: TrackDesc(times, vals, interpolation=INTERPOLATION_CONSTANT)})
    assert np.array_equal(O.pose_bones(sk, an, 0.99, I16)[0].reshape(4, 4)[3, :3], [1, 0, 0])
    # repeated key times: "last key before, first key not before" (the linear scan of animation.go:77-86)
    sk, an = one_bone({TRACK_POSITION: TrackDesc([0.0, 1.0, 1.0, 2.0], [[0, 0, 0], [1, 0, 0], [5, 0, 0], [7, 0, 0]])})
    assert O.pose_bones(sk, an, 1.0, I16)[0][12] == 1.0
    assert O.pose_bones(sk, an, 1.5, I16)[0][12] == 6.0
    # a bone without tracks keeps its rest transform; scale tracks scale
    sk, an = one_bone({TRACK_SCALE: TrackDesc([0.0], [[2, 3, 4]])})
    assert np.array_equal(np.diag(O.pose_bones(sk, an, 0.3, I16)[0].reshape(4, 4)), [2, 3, 4, 1])


def test_slerp_branches():
    a = np.array([0, 0, 0, 1], F32)
    b = np.array([0, math.sin(0.6), 0, math.cos(0.6)], F32)          # 1.2 rad about Y
    h = O.quat_slerp(a, b, 0.5)
    assert np.allclose(h, [0, math.sin(0.3), 0, math.cos(0.3)], atol=2e-7)
    assert np.array_equal(O.quat_slerp(a, -b, 0.5), h)               # opposite hemisphere: end is negated first
    assert np.array_equal(O.quat_slerp(a, b, 0.0)[[0, 2]], [0, 0])
    # cosOmega > 0.9999: plain linear blend, no normalisation (quaternion.go:40-43)
    c = np.array([0, 0.001, 0, 1], F32)
    got = O.quat_slerp(a, c, 0.25)
    assert np.array_equal(got, a * (F32(1) - F32(0.25)) + c * F32(0.25))
    # ToMatrix4 of the identity quaternion and of a zero quaternion (which SetLocalRotation ignores, node.go:906-909)
    assert np.array_equal(O.quat_to_matrix([0, 0, 0, 1]), np.eye(4, dtype=F32))
    assert not O.quat_to_matrix([0, 0, 0, 0]).any()
    sk, an = one_bone({TRACK_ROTATION: TrackDesc([0.0], [[0, 0, 0, 0]])})
    assert np.array_equal(O.pose_bones(sk, an, 0.0, I16)[0], I16)


def test_pose_hierarchy_matches_a_float64_evaluation():
    rng = np.random.default_rng(5)
    sk = scenes.random_skeleton(rng, 12)
    root = scenes.random_rotation(rng).astype(np.float64).reshape(4, 4)
    root[3, :3] = (1.0, -2.0, 0.5)
    got = O.pose_bones(sk, None, 0.0, root.astype(F32).reshape(16)).reshape(-1, 4, 4)
    world = []
    for b in range(12):
        local = np.diag(np.append(sk.rest_scale[b].astype(np.float64), 1.0)) @ sk.rest_rotation[b].astype(np.float64).reshape(4, 4)
        tr = np.eye(4); tr[3, :3] = sk.rest_position[b]
        local = local @ tr
        parent = root.astype(F32).astype(np.float64) if sk.parent[b] < 0 else world[sk.parent[b]]
        world.append(local @ parent)
        want = sk.inverse_bind[b].astype(np.float64).reshape(4, 4) @ world[b]
        assert np.allclose(got[b], want, rtol=1e-5, atol=1e-5)


def test_chain_animation_reproduces_the_handwritten_chain_at_rest():
    # chain_skeleton at rest == chain_bone_influences with no bend: translations only
    mesh, bone_y = W.tube_character_mesh(rings=9, sides=10, n_bones=5)
    sk = W.chain_skeleton(bone_y)
    root = I16.copy(); root[12:15] = (2.0, 0.0, -1.0)
    got = O.pose_bones(sk, None, 0.0, root).reshape(-1, 4, 4)
    for b in range(5):
        assert np.allclose(got[b], np.array(np.eye(4), F32) + np.pad(np.array([[2.0, 0.0, -1.0]], F32), ((3, 0), (0, 1))), atol=1e-6)


def test_animation_player_playhead_logic():
    import tetra3d_b200 as t3d
    anim = t3d.NewAnimation("walk"); anim.SetLength(2.0)
    ap = t3d.NewAnimationPlayer(t3d.Node("armature"))
    ap.Play(anim)
    assert ap.IsPlaying() and ap.Playhead() == 0.0
    seen = []
    for _ in range(4):
        ap.Update(0.75)
        seen.append((ap.SampledPlayhead(), ap.Playhead(), ap.Finished(), ap.LoopCount()))
    # tracks are sampled BEFORE the playhead advances (animation.go:686-722); the loop wraps by subtracting the length
    assert seen == [(0.0, 0.75, False, 0), (0.75, 1.5, False, 0), (1.5, 0.25, True, 1), (0.25, 1.0, False, 1)]
    # ping-pong: clamps to the end and reverses (animation.go:773-788)
    anim.SetLoopMode(t3d.AnimationLoopModePingPong)
    ap2 = t3d.NewAnimationPlayer(t3d.Node("armature")); ap2.Play(anim)
    ap2.Update(1.5); ap2.Update(1.5)
    assert ap2.Playhead() == 2.0 and ap2.PlaySpeed == -1.0 and ap2.Finished()
    ap2.Update(0.5)
    assert ap2.Playhead() == 1.5 and ap2.SampledPlayhead() == 2.0
    # one-shot: stops at the end (animation.go:790-802); SetPlayhead samples at once (animation.go:828-831)
    anim.SetLoopMode(t3d.AnimationLoopModeOneshot)
    ap3 = t3d.NewAnimationPlayer(t3d.Node("armature")); ap3.Play(anim)
    ap3.Update(2.5)
    assert ap3.Finished() and not ap3.IsPlaying()
    ap3.SetPlayhead(0.4)
    assert ap3.SampledPlayhead() == pytest.approx(0.4) and ap3.Playhead() == pytest.approx(0.4)
    # negative play speed starts at the end (animation.go:541-545)
    ap4 = t3d.NewAnimationPlayer(t3d.Node("armature")); ap4.PlaySpeed = -1.0; ap4.Play(anim)
    assert ap4.Playhead() == 2.0


def test_posed_scenes_render_in_the_oracle():
    for seed in range(3):
        r = O.render_scene(scenes.posed_scene(seed))
        assert r.n_vertices > 0
