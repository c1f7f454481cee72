"""The C-ABI library loads without a GPU, exports every symbol include/t3dgeo.h declares, and fails loudly
(no CPU fallback) when asked to create a context on a machine without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "t3dgeo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(t3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import tetra3d_b200
    lib = ctypes.CDLL(tetra3d_b200.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libt3dgeo.so does not export %s" % n
    assert lib.t3d_abi_version() == 6


def test_struct_sizes_match_header():
    """ctypes mirrors of the ABI structs a cgo shim would declare: sizes must match the C compiler's."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "t3dgeo.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(t3d_vertex), sizeof(t3d_light), sizeof(t3d_draw), sizeof(t3d_camera), sizeof(t3d_draw_result), sizeof(t3d_part_range), sizeof(t3d_world), sizeof(t3d_pose), sizeof(t3d_track));printf("%zu %zu\\n", sizeof(t3d_instance), sizeof(t3d_pose_chain));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "s")]).split()))
    assert sizes[9] == 40           # t3d_instance: position[3] + scale[3] + color[4]
    assert sizes[10] == 96          # t3d_pose_chain: root[16] + playhead[4] + blend[4]
    assert sizes[0] == 48           # ebiten.Vertex: 12 x float32
    assert sizes[1] == 52 and sizes[5] == 16 and sizes[4] == 24
    assert sizes[2] % 8 == 0
    assert sizes[7] == 80 and sizes[8] == 12      # t3d_pose: root[16] + playhead + from_playhead + blend + 1 reserved float; t3d_track: 3 x uint32


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import tetra3d_b200 as t3d
    with pytest.raises(t3d.T3DError) as e:
        t3d.GeometryDevice(0)
    assert "no CPU fallback" in str(e.value)
    # and a Camera without a device refuses to render
    import scenes
    scene, cam, _ = t3d.build_scene(scenes.soup_scene(1))
    cam.Clear()
    with pytest.raises(t3d.T3DError):
        cam.RenderScene(scene)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tetra3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", "") or f in ("scenedesc.py", "workloads.py", "build.py") and "import oracle" not in text and "from oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text and "t3d_oracle" not in text, f


def test_energy_cutoff_is_the_exact_threshold_of_the_reference_test():
    """The default lighting decides `1 / d * energy < 0.001` (light.go:205-210) with ONE comparison against a distance the host
    finds by bisection (t3d_submit_lights).  That distance must be the exact threshold of the reference's float32 expression:
    the test passes at T and fails at the next float above it -- for any energy, including the degenerate ones."""
    import numpy as np
    import tetra3d_b200
    lib = ctypes.CDLL(tetra3d_b200.LIB_PATH)
    lib.t3d_debug_energy_cutoff.restype = ctypes.c_float
    lib.t3d_debug_energy_cutoff.argtypes = [ctypes.c_float]

    def skipped(d, e):          # the reference's expression, float32 with two roundings
        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            return bool(np.float32(np.float32(1.0) / np.float32(d)) * np.float32(e) < np.float32(0.001))

    rng = np.random.default_rng(3)
    energies = np.concatenate([rng.uniform(0.01, 50, 200), 10.0 ** rng.uniform(-30, 30, 200), [1.0, 2.0, 40.0, 1e-45, 3e38]]).astype(np.float32)
    for e in energies:
        T = np.float32(lib.t3d_debug_energy_cutoff(float(e)))
        assert T > 0 and not skipped(T, e), (e, T)
        up = np.nextafter(T, np.float32(np.inf))
        assert np.isinf(up) or skipped(up, e), (e, T, up)
    assert lib.t3d_debug_energy_cutoff(0.0) == 0.0 and lib.t3d_debug_energy_cutoff(-2.0) == 0.0        # nothing at a distance passes
    assert np.isinf(lib.t3d_debug_energy_cutoff(float("nan")))                                           # NaN never compares below
