"""Extended one-off fuzz on a GPU box: 160 more random frames plus 40 seeds each of the posed, cell-light and subdivision
scenes, bit-exact against the oracle (tests/ keeps a smaller fixed set).  Usage: gpurun -- python tools/fuzz_more.py"""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import tetra3d_b200 as t3d, scenes
from compare import assert_parity
dev = t3d.GeometryDevice(0)
bad = 0
for seed in range(100, 260):
    try:
        assert_parity(scenes.fuzz_scene(seed), dev, check_capture=False)
    except AssertionError as e:
        bad += 1; print("FUZZ FAIL", seed, str(e)[:200])
for seed in range(100, 140):
    try:
        assert_parity(scenes.posed_scene(seed, n_models=int(3 + seed % 7), n_bones=int(2 + seed % 40), chain=bool(seed % 2)), dev, check_capture=False)
        assert_parity(scenes.volume_cell_lights_scene(seed, mode=seed % 3, skinned=bool(seed % 2)), dev)
        assert_parity(scenes.subdivision_scene(seed), dev, check_capture=False)
    except AssertionError as e:
        bad += 1; print("NEW FAIL", seed, str(e)[:200])
print("done, failures:", bad)
