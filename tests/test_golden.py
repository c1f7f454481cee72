"""The oracle against the committed fixtures of tests/golden (regression pin; see make_golden.py)."""
import os

import numpy as np
import pytest

import scenes
from oracle import binding as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys_path_hack = None
import importlib.util
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


@pytest.mark.parametrize("name", make_golden.SCENES)
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(HERE, name + ".npz"))
    r = O.render_scene(scenes.ALL_SMALL[name]())
    assert np.array_equal(r.tri_ids, g["tri_ids"])
    assert [x["vertex_count"] for x in r.renders] == list(g["render_counts"])
    assert [x["vertex_count"] for x in r.flushes] == list(g["flush_counts"])
    s = int(g["stride"])
    assert np.array_equal(r.color_list[::s].view(np.uint32), g["color_sample"].view(np.uint32))
    assert np.array_equal(r.depth_list[::s].view(np.uint32), g["depth_sample"].view(np.uint32))
    assert make_golden.digest(r.color_list) == str(g["color_sha256"])
    assert make_golden.digest(r.depth_list) == str(g["depth_sha256"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", make_golden.SCENES)
def test_gpu_matches_golden(device, name):
    """The CUDA path against the committed fixtures directly (no oracle call)."""
    from compare import render_gpu
    g = np.load(os.path.join(HERE, name + ".npz"))
    scene, cam, objs = render_gpu(scenes.ALL_SMALL[name](), device)
    assert np.array_equal(cam.TriangleIDs(), g["tri_ids"])
    assert [x["vertex_count"] for x in cam.Renders()] == list(g["render_counts"])
    s = int(g["stride"])
    col = cam.ColorVertexList(); dep = cam.DepthVertexList()
    assert np.array_equal(dep[::s].view(np.uint32), g["depth_sample"].view(np.uint32))
    assert np.array_equal(col[::s][:, :4].view(np.uint32), g["color_sample"][:, :4].view(np.uint32))
    assert np.max(np.abs(col[::s][:, 4:8] - g["color_sample"][:, 4:8])) <= 1.0 / 255
