"""The reference's own example assets as parity inputs.

tests/golden/assets/*.pkl.gz are SceneDescs built from /root/reference/examples/{animation,lighting}/assets/*.glb -- and, as
ex_*.pkl.gz, from every other GLB under /root/reference/examples (28 files: shapes, transparency, blendmodes, baking, orthographic,
skeletal2d, worlds, ...) -- by tests/golden/make_asset_fixtures.py through tetra3d_b200.glb (gltf.go's load rules): real skins
with the influences the loader keeps, morph targets, an 18-channel armature animation, KHR_lights_punctual lights, vertex colours,
several materials per mesh, transparent materials, the exporter's world / camera settings.
CPU: the oracle reproduces the committed digests, and (where /root/reference exists) the loader reproduces the fixtures.
GPU: the CUDA path against the oracle on the same inputs, bit for bit."""
import gzip
import json
import os
import pickle

import numpy as np
import pytest

from oracle import binding as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ASSETS = os.path.join(HERE, "assets")
import importlib.util
_spec = importlib.util.spec_from_file_location("make_asset_fixtures", os.path.join(HERE, "make_asset_fixtures.py"))
maf = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(maf)

DIGESTS = json.load(open(os.path.join(ASSETS, "digests.json")))
NAMES = sorted(DIGESTS)


def load(name):
    with gzip.open(os.path.join(ASSETS, name + ".pkl.gz"), "rb") as f:
        return pickle.load(f)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_asset_digests(name):
    assert maf.summarize(O.render_scene(load(name))) == DIGESTS[name]


def test_fixtures_hold_real_skins_and_morph_targets():
    sc = load("animations_t1")
    skinned = [m for m in sc.models if m.pose is not None]
    assert len(skinned) == 1 and skinned[0].pose.skeleton.parent.tolist() == [-1, 0, 1, 2, 3, 4]
    assert len(skinned[0].pose.animation.channels) == 6 and all(len(c) == 3 for c in skinned[0].pose.animation.channels.values())
    morph = [m for m in sc.models if m.mesh.key_weights is not None]
    assert len(morph) == 1 and morph[0].mesh.key_positions.shape == (2, 66, 3) and np.abs(morph[0].mesh.key_positions).max() > 0
    sc = load("lighting_wide")
    sk = [m for m in sc.models if m.pose is not None][0]
    assert set(np.unique(sk.mesh.n_influences)) == {1, 2}          # gltf.go:692-700 keeps only the non-zero weights
    assert sorted(l.type for l in sc.lights) == [1, 1, 1, 2]


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/animation/assets/animations.glb"), reason="the reference tree is not on this box")
def test_loader_reproduces_the_fixtures():
    for name, sc in maf.scenes().items():
        fx = load(name)
        assert len(sc.models) == len(fx.models) and len(sc.lights) == len(fx.lights)
        for a, b in zip(sc.models, fx.models):
            for f in ("positions", "normals", "uvs", "indices", "colors", "weights", "bones", "n_influences", "key_positions", "key_weights"):
                x, y = getattr(a.mesh, f), getattr(b.mesh, f)
                assert (x is None) == (y is None) and (x is None or np.array_equal(x, y)), (name, a.name, f)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_oracle_on_reference_assets(device, name):
    from compare import assert_parity
    assert_parity(load(name), device, check_capture=False)
