"""tetra3d_b200 -- B200-native geometry stage for Tetra3D.

`tetra3d_b200._host` is the C++ host-side mirror of the reference's Go API (Scene, Node, Model, Mesh,
Material, Camera, lights, LightGroup); it drives `lib/libt3dgeo.so` (hand-written sm_100a kernels behind
the C ABI of include/t3dgeo.h).  There is no CPU implementation of the geometry stage in this package:
importing works without a GPU (so the build can be checked), creating a `GeometryDevice` does not.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "lib", "libt3dgeo.so")

try:
    from . import _host  # noqa: F401
    from ._host import *  # noqa: F401,F403
except ImportError as _e:  # pragma: no cover - build missing
    raise ImportError(
        "tetra3d_b200: the native extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C tetra3d_b200/csrc && make -C tetra3d_b200/host`): %s" % (_e,)) from _e

from . import scenedesc  # noqa: E402,F401
from .build import build_scene  # noqa: E402,F401
