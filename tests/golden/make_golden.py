#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference cannot run in this image (no Go toolchain) and holds no golden vectors for this path, so these
fixtures pin the ORACLE against regressions (parity itself stays "unpinned", see oracle/t3d_oracle.h): triangle
order, per-render counts and a strided sample of the emitted vertices for a few seeded scenes.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import scenes  # noqa: E402
from oracle import binding as O  # noqa: E402

SCENES = ["soup_lit", "soup_ortho", "skinned", "billboard_locks", "batch_transparent", "volume_mode1", "transparent",
          "posed", "volume_cell_lights", "subdivision"]


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def record(name):
    r = O.render_scene(scenes.ALL_SMALL[name]())
    stride = max(1, r.n_vertices // 64)
    return dict(
        tri_ids=r.tri_ids, render_counts=np.array([x["vertex_count"] for x in r.renders], np.int64),
        flush_counts=np.array([x["vertex_count"] for x in r.flushes], np.int64),
        color_sample=r.color_list[::stride].copy(), depth_sample=r.depth_list[::stride].copy(), stride=np.int64(stride),
        color_sha256=np.array(digest(r.color_list)), depth_sha256=np.array(digest(r.depth_list)))


if __name__ == "__main__":
    for n in (sys.argv[1:] or SCENES):
        np.savez_compressed(os.path.join(HERE, n + ".npz"), **record(n))
        print("wrote", n)
