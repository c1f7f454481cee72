#!/usr/bin/env python
"""Prints the timeline of k_vertex on one frame of a workload (needs a libt3dgeo.so built with -DT3D_VERTEX_TRACE first on
LD_LIBRARY_PATH: tools/build_variants.sh trace "-DT3D_VERTEX_TRACE").  usage: tools/vertex_trace.py [workload] [staged|chained]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tetra3d_b200 as t3d  # noqa: E402
from tetra3d_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "million_triangles"
staged = (sys.argv[2] if len(sys.argv) > 2 else "staged") == "staged"
device = t3d.GeometryDevice(0)
scene, cam, _ = t3d.build_scene(W.WORKLOADS[name](), device)
cam.DownloadResults = False
device.SetStageTimings(staged)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.add_(1); torch.cuda.synchronize()
    cam.Clear(); cam.RenderScene(scene)
tm = cam.Timings()
lib = C.CDLL("libt3dgeo.so")       # the copy the extension module already loaded (LD_LIBRARY_PATH decides which build)
n = 1024
buf = (C.c_ulonglong * (16 * n))()
lib.t3d_debug_vertex_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
rc = lib.t3d_debug_vertex_trace(C.c_void_p(device.Handle()), buf, n)
assert rc == 0, rc
a = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.int64)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
print("%s (%s): vertex_ms %.4f total_ms %.4f, %d blocks traced" % (name, "stage events" if staged else "chained", tm["vertex_ms"], tm["total_ms"], len(a)))
labels = ["block start", "past dependency wait", "producer: first record read", "producer: first tile issued", "producer: last tile issued",
          "consumer: first tile arrived", "consumer: first tile shaded", "consumer: last tile shaded"]
for i, lab in enumerate(labels):
    col = a[:, i][a[:, i] > 0] - t0
    if len(col):
        print("  %-30s min %7.2f  median %7.2f  p90 %7.2f  max %7.2f us" % (lab, col.min() / 1e3, np.median(col) / 1e3, np.percentile(col, 90) / 1e3, col.max() / 1e3))
print("  tiles per block: min %d median %d max %d" % (a[:, 8].min(), np.median(a[:, 8]), a[:, 8].max()))
d = (a[:, 6] - a[:, 5])[a[:, 6] > 0]
print("  first tile shading time: median %.2f us, max %.2f us" % (np.median(d) / 1e3, d.max() / 1e3))
