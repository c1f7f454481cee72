import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def device():
    """One GeometryDevice (t3d_ctx) for the whole GPU test session; creation fails loudly without a GPU."""
    import tetra3d_b200 as t3d
    return t3d.GeometryDevice(0)


def pytest_sessionfinish(session, exitstatus):
    """Largest colour deviation of the default (approximate-reciprocal) lighting against the oracle over the GPU session:
    printed and, when gpurun_out/ exists, written next to the other run artefacts (tolerance: 1/255 = 3.92e-3)."""
    try:
        import compare
    except Exception:
        return
    d = compare.FAST_COLOR_MAXDIFF
    if not d:
        return
    import json
    worst = max(d.values())
    print("\nfast-lighting colour deviation vs oracle: max %.3g over %d scenes (tolerance 1/255 = %.3g)" % (worst, len(d), 1.0 / 255))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "fast_lighting_color_maxdiff.json"), "w") as f:
            json.dump({"max": worst, "tolerance": 1.0 / 255, "scenes": d}, f, indent=1, sort_keys=True)
