"""tetra3d_b200 -- B200-native geometry stage for Tetra3D.

`tetra3d_b200._host` is the C++ host-side mirror of the reference's Go API (Scene, Node, Model, Mesh,
Material, Camera, lights, LightGroup); it drives `lib/libt3dgeo.so` (hand-written sm_100a kernels behind
the C ABI of include/t3dgeo.h).  There is no CPU implementation of the geometry stage in this package:
importing works without a GPU (so the build can be checked), creating a `GeometryDevice` does not.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
LIB_PATH = _os.path.join(_HERE, "lib", "libt3dgeo.so")

from . import scenedesc  # noqa: E402,F401   (pure Python: scene descriptions shared by tests and benches)


def _load_host():
    """The native extension is loaded on first use, not at import: `tetra3d_b200.scenedesc` / `.workloads` are plain
    data modules that the CPU reference arm uses without touching the product's native code."""
    import importlib
    try:
        return importlib.import_module("._host", __name__)
    except ImportError as _e:  # pragma: no cover - build missing
        raise ImportError(
            "tetra3d_b200: the native extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C tetra3d_b200/csrc && make -C tetra3d_b200/host`): %s" % (_e,)) from _e


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    if name in ("build_scene", "build_animation"):
        from . import build as _build
        globals()[name] = getattr(_build, name)
        return globals()[name]
    if name in ("workloads", "parallel", "build", "glb"):
        import importlib
        return importlib.import_module("." + name, __name__)
    host = _load_host()
    if name == "_host":
        return host
    try:
        v = getattr(host, name)
    except AttributeError:
        raise AttributeError("module 'tetra3d_b200' has no attribute %r" % (name,)) from None
    globals()[name] = v
    return v
