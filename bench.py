#!/usr/bin/env python
"""bench.py -- geometry-stage throughput of the B200 path (and, with --impl reference, of the CPU reference arm).

Metric (BASELINE.json): transformed+lit triangles/s through the full geometry stage (skin, transform, cull, light,
512-bin depth sort, emit), plus geometry ms/frame.  A step is one frame of the workload.

  value  device-resident: meshes already in HBM; per step the frame's uniforms go up, the kernels run, the draw
         lists stay on the device.  Timed with CUDA events on the library's launch stream (t3d_timings), L2 flushed
         between steps.  At N > 1 the scene is N copies of the workload sharded object-wise (weak scaling) and the
         NCCL gather of the draw lists to rank 0 is inside the step.
  e2e    the same frame through the reference-facing API with host buffers: Camera.RenderScene (scene traversal,
         uniform derivation, pinned H2D, kernels, pinned D2H of the colour + depth lists), wall clock.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "geometry-stage transformed+lit triangles/s"
UNIT = "triangles/s"
DEFAULT_WORKLOAD = "million_triangles"      # BASELINE.json configs[1]: the configuration the metric is quoted on


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def base_config(workload, desc, copies):
    """The `config` object both arms print: identical keys and values for the same command line."""
    return {"workload": workload, "copies": copies, "triangles_per_frame": desc.total_triangles(), "vertices": desc.total_vertices(),
            "note": "the scene holds `copies` copies of the workload: one per GPU by default (weak scaling), 1 when ONE scene is sharded over the GPUs (--copies 1, strong scaling); the reference arm renders the same scene on host cores"}


def make_scene(workload, copies=1):
    from tetra3d_b200 import workloads as W
    from tetra3d_b200.scenedesc import ModelDesc, NodeDesc
    desc = W.WORKLOADS[workload]()
    if copies > 1:       # weak scaling: N instances of the workload's models, 250 units apart along Z
        base = list(desc.models)
        for c in range(1, copies):
            for m in base:
                nd = NodeDesc(position=(m.node.position[0], m.node.position[1], m.node.position[2]), scale=m.node.scale,
                              rotation=m.node.rotation, parent=m.node.parent, name=m.node.name)
                desc.models.append(ModelDesc(m.mesh, nd, color=m.color, shadeless=m.shadeless, visible=m.visible,
                                             frustum_culling=m.frustum_culling, bone_influences=m.bone_influences,
                                             light_group=m.light_group, name="%s_copy%d" % (m.name, c)))
    return desc


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; Go toolchain absent) on host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_frames(desc, seconds, min_frames=3, max_frames=50):
    from oracle import binding as O
    m = O.marshal(desc)
    O.render_scene(m, copy=False)         # warm-up (page-faults the lists)
    frames, t_frame, split = 0, 0.0, {"process": 0.0, "light": 0.0, "emit": 0.0}
    t0 = time.perf_counter()
    while frames < min_frames or (time.perf_counter() - t0 < seconds and frames < max_frames):
        r = O.render_scene(m, copy=False)
        t_frame += r.times["total"]
        for k in split:
            split[k] += r.times[k]
        frames += 1
    return frames, t_frame, {k: v / frames * 1e3 for k, v in split.items()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc = make_scene(args.workload, copies=args.copies if args.copies else max(args.gpus, 1))
    tris = desc.total_triangles()
    from oracle import binding as O
    m = O.marshal(desc)
    for _ in range(args.warmup):
        O.render_scene(m, copy=False)
    t = 0.0
    for _ in range(args.steps):
        t += O.render_scene(m, copy=False).times["total"]
    value = tris * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.workload, desc, args.copies if args.copies else max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d full frames of the workload; C restatement of the single-goroutine Go path (no Go toolchain in the image), gcc -O2 -ffp-contract=off, 1 thread" % args.steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
STAGE_KERNELS = {"vertex": "k_vertex", "cull": "k_tile_packed / k_draw_medium / k_cull_large / k_cull_sort (the last also sorts)",
                 "sort": "k_tile_scan + k_bin_hist + k_scatter (draws the cull kernels did not sort themselves)", "emit": "k_emit"}


def algorithmic_bytes(cam, desc, device, n_vertices_out):
    """Per-stage algorithmic bytes of this rank's frame (DESIGN.md 4 / SURVEY.md 8d)."""
    surv = n_vertices_out // 3
    local_tris = sum(r["total_tris"] for r in cam.Renders())
    return {"vertex": device.VertexKernelBytes(), "cull": 72 * local_tris + 8 * surv, "sort": (16 * 1 + 4) * surv,
            "emit": (600 if desc.camera.render_normals else 456) * surv}, surv


def stage_table(stage_sum, steps, alg, peak):
    stages = {}
    for k in ("vertex", "cull", "sort", "emit"):
        ms_k = stage_sum.get(k + "_ms", 0.0) / steps
        gbs = alg[k] / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
        stages[k] = {"ms": ms_k, "algorithmic_bytes": alg[k], "achieved_gbs": gbs, "frac": gbs / peak}
    stages["prep"] = {"ms": stage_sum.get("prep_ms", 0.0) / steps}
    return stages


def d2h_ceiling_gbs(torch, nbytes=256 << 20, reps=5):
    """What a plain pinned cudaMemcpy device -> host reaches on this box (the ceiling of `e2e`, which has to bring the lists back)."""
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    best = 0.0
    for _ in range(reps + 1):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def parity_check(desc, gpu_lists):
    """Outside every timed region: the frame's lists against the CPU oracle on the same scene.  gpu_lists(flush_id) ->
    (colour records, depth records) as float32 (n, 12) arrays.  Positions, texture coordinates, custom fields and the whole depth
    list must be bit-identical in the oracle's order; colours within 1/255 (default lighting)."""
    import hashlib
    import numpy as np
    from oracle import binding as O
    ref = O.render_scene(desc)
    hg, ho = hashlib.sha256(), hashlib.sha256()
    ok, worst, nv = True, 0.0, 0
    geo = (0, 1, 2, 3, 8, 9, 10, 11)
    for fid, f in enumerate(ref.flushes):
        n, first = f["vertex_count"], f["first_vertex"]
        col, dep = gpu_lists(fid)
        if len(col) != n or len(dep) != n:
            ok = False
            continue
        rc, rd = ref.color_list[first:first + n], ref.depth_list[first:first + n]
        ga, oa = np.ascontiguousarray(col[:, geo]).view(np.uint32), np.ascontiguousarray(rc[:, geo]).view(np.uint32)
        gd, od = np.ascontiguousarray(dep).view(np.uint32), np.ascontiguousarray(rd).view(np.uint32)
        hg.update(ga.tobytes()); hg.update(gd.tobytes()); ho.update(oa.tobytes()); ho.update(od.tobytes())
        ok = ok and np.array_equal(ga, oa) and np.array_equal(gd, od)
        if n:
            worst = max(worst, float(np.max(np.abs(col[:, 4:8].astype(np.float64) - rc[:, 4:8].astype(np.float64)))))
        nv += n
    return {"lists_bit_exact": bool(ok and hg.hexdigest() == ho.hexdigest()), "colour_max_abs_diff": worst, "colour_tolerance": 1.0 / 255,
            "colours_within_tolerance": bool(worst <= 1.0 / 255), "vertices": nv, "flushes": len(ref.flushes),
            "sha256_gpu": hg.hexdigest()[:16], "sha256_oracle": ho.hexdigest()[:16],
            "what": "DstX/DstY/SrcX/SrcY/Custom0-3 of the colour list and the whole depth list, flush by flush in the oracle's order, hashed on both sides; colours compared by value"}


def measure_workload(t3d, device, torch, name, peak, steps, flush_l2):
    """Device-resident frame, per-stage roofline fractions and e2e of one more workload (N = 1, after the headline region)."""
    from tetra3d_b200 import workloads as W
    t0 = time.perf_counter()
    desc = W.WORKLOADS[name]()
    scene, cam, objs = t3d.build_scene(desc, device)
    cam.DownloadResults = False
    device.SetStageTimings(False)
    for _ in range(3):
        flush_l2(); cam.Clear(); cam.RenderScene(scene)
    ms = 0.0
    for _ in range(steps):
        flush_l2(); cam.Clear(); cam.RenderScene(scene)
        ms += cam.Timings()["total_ms"]
    device.SetStageTimings(True)
    flush_l2(); cam.Clear(); cam.RenderScene(scene)
    stage_sum = {}
    for _ in range(steps):
        flush_l2(); cam.Clear(); cam.RenderScene(scene)
        for k, v in cam.Timings().items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    device.SetStageTimings(False)
    alg, surv = algorithmic_bytes(cam, desc, device, cam.VertexCount())
    stages = stage_table(stage_sum, steps, alg, peak)
    launches = cam.LaunchCount()
    cam.DownloadResults = True
    cam.Clear(); cam.RenderScene(scene)
    e0 = time.perf_counter()
    for _ in range(max(steps // 2, 2)):
        cam.Clear(); cam.RenderScene(scene)
    e2e_ms = (time.perf_counter() - e0) / max(steps // 2, 2) * 1e3
    h2d, d2h = device.FrameBytes()
    # the same with two frames in flight (the lists of frame k cross PCIe while frame k + 1 is built and executed)
    e2e_async_ms = None
    try:
        cam.AsyncDownload = True
        for k in range(3):      # both sets of device lists and pinned buffers exist after two frames (the second set is allocated on first use)
            cam.Clear(); cam.RenderScene(scene)
            if k:
                cam.WaitDownload(1)
        cam.WaitDownload(0)
        n_async = max(steps // 2, 2)
        e0 = time.perf_counter()
        for _ in range(n_async):
            cam.Clear(); cam.RenderScene(scene)
            cam.WaitDownload(1)
        cam.WaitDownload(0)
        e2e_async_ms = (time.perf_counter() - e0) / n_async * 1e3
    except Exception:
        e2e_async_ms = None
    finally:
        cam.AsyncDownload = False
    tris = desc.total_triangles()
    out = {"triangles_per_frame": tris, "vertices": desc.total_vertices(), "draws": len([r for r in cam.Renders() if r["draw_index"] >= 0]),
           "surviving_triangles": surv, "device_ms": ms / steps, "triangles_per_s": tris / (ms / steps * 1e-3), "launches_per_frame": launches,
           "e2e_ms": e2e_ms, "e2e_triangles_per_s": tris / (e2e_ms * 1e-3), "h2d_bytes": int(h2d), "d2h_bytes": int(d2h),
           "e2e_two_frames_in_flight_ms": e2e_async_ms,
           "stages": {k: ({"ms": v["ms"], "frac": v["frac"]} if "frac" in v else {"ms": v["ms"]}) for k, v in stages.items()},
           "build_s": None}
    del scene, cam, objs
    out["build_s"] = time.perf_counter() - t0
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import tetra3d_b200 as t3d
    from tetra3d_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the geometry stage has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_start = time.perf_counter()

    copies = args.copies if args.copies else world
    desc = make_scene(args.workload, copies=copies)
    tris_total = desc.total_triangles()
    device = t3d.GeometryDevice(local)
    scene, cam, objs = t3d.build_scene(desc, device)
    cam.SetShard(rank, world)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def flush_l2():
        flush_buf.add_(1)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    delivered = [None]

    def device_step():
        """One device-resident frame: the event pair of the library's stream around everything the frame enqueues.  At N > 1
        that includes the delivery to rank 0 (t3d_gather_*): a shard's compact emit over NVLink, statistics and post; on rank 0
        the wait for every shard and the kernels that rebuild their lists.  The ranks are aligned by a barrier before the step;
        nothing crosses the host inside it."""
        flush_l2()
        if world > 1:
            dist.barrier()
        cam.Clear()
        cam.RenderScene(scene)
        tm = cam.Timings()
        return tm["total_ms"], tm

    # ---- device-resident value: kernels chained with programmatic dependent launches, one event pair per frame ----
    cam.DownloadResults = False
    device.SetStageTimings(False)
    if world > 1:
        cam.Clear(); cam.RenderScene(scene)          # sizes the delivery: 3 vertices per triangle and one entry per draw this rank submits
        delivered[0] = parallel.DeliveredLists(device, local, 3 * sum(r["total_tris"] for r in cam.Renders()), len(cam.Renders()), dst=0)
    for _ in range(max(args.warmup, 3)):
        device_step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ms_sum = 0.0
    gather_sum = 0.0
    for _ in range(args.steps):
        ms, tm = device_step()
        ms_sum += ms
        gather_sum += tm.get("gather_ms", 0.0)
    barrier()
    wall1 = time.perf_counter()
    # ---- per-stage durations for the roofline: same frames with a CUDA event between the kernels (which
    #      serialises them, so the stages add up to more than the chained frame above) ----
    device.SetStageTimings(True)
    stage_sum = {}
    for _ in range(3):
        device_step()
    ms_staged = 0.0
    for _ in range(args.steps):
        ms, tm = device_step()
        ms_staged += tm["total_ms"]
        for k, v in tm.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    device.SetStageTimings(False)
    barrier()
    launches_per_step = cam.LaunchCount()
    n_vertices_out = cam.VertexCount()
    alg, surv = algorithmic_bytes(cam, desc, device, n_vertices_out)
    t = torch.tensor([ms_sum], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = torch.tensor([ms_sum / args.steps, gather_sum / args.steps], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"frame_ms": float(a[0]), "of_which_delivery_ms": float(a[1])} for a in allr]
    ms_per_step = float(t.item()) / args.steps
    value = tris_total / (ms_per_step * 1e-3)

    # ---- parity of what was just measured: the delivered lists (N > 1) or this GPU's lists (N = 1) against the oracle ----
    parity = None
    if world > 1:
        device_step()
        lists, table = delivered[0].results(cam)
        if rank == 0:
            def flush_lists(fid):      # one flush's slice at a time: the delivered lists of a 100 M-triangle frame are 13 GB
                a, n = table[fid]
                return (lists[0][48 * a:48 * (a + n)].cpu().numpy().view(np.float32).reshape(-1, 12),
                        lists[1][48 * a:48 * (a + n)].cpu().numpy().view(np.float32).reshape(-1, 12))
            parity = parity_check(desc, flush_lists)
        barrier()

    before_gather = None
    if world > 1:
        # the same frames with every rank's lists staying in its own HBM: the geometry stage before the delivery
        # (SURVEY.md 8e asks for both figures; `value` above is the one after delivery to the issuing rank)
        delivered[0].release()
        for _ in range(3):
            flush_l2(); cam.Clear(); cam.RenderScene(scene)
        barrier()
        loc = 0.0
        for _ in range(args.steps):
            flush_l2(); cam.Clear(); cam.RenderScene(scene)
            loc += cam.Timings()["total_ms"]
        tl = torch.tensor([loc], dtype=torch.float64, device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        loc_ms = float(tl.item()) / args.steps
        before_gather = {"value": tris_total / (loc_ms * 1e-3), "unit": UNIT, "ms_per_step": loc_ms,
                         "note": "every rank's lists stay in its own HBM (no delivery to the issuing rank), max over ranks"}

    # ---- e2e through the public API with host buffers ----
    # (a) one frame at a time: Camera.RenderScene returns with the lists complete in pinned host memory
    cam.DownloadResults = True
    for _ in range(3):
        cam.Clear(); cam.RenderScene(scene)
    barrier()
    e0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        cam.Clear()
        cam.RenderScene(scene)
        h2d, d2h = device.FrameBytes()
    barrier()
    sync_s = time.perf_counter() - e0
    ts = torch.tensor([sync_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    sync_ms = float(ts.item()) / args.steps * 1e3
    # (b) two frames in flight (Camera.AsyncDownload): RenderScene(k) starts the copy of frame k's lists and returns; the loop
    # consumes frame k - 1 (waits until ITS lists are complete in host memory) while that copy runs, as a game loop that hands frame
    # k - 1 to DrawTriangles while frame k renders.  Every step still moves one full frame of lists over PCIe; the last frame's
    # copy is inside the timed region.  This is the figure reported as `e2e.value`; (a) stays next to it.
    e2e_s, frames_in_flight, async_note = sync_s, 1, None
    try:
        cam.AsyncDownload = True
        for k in range(4):
            cam.Clear(); cam.RenderScene(scene)
            if k:
                cam.WaitDownload(1)
        cam.WaitDownload(0)
        barrier()
        e0 = time.perf_counter()
        for k in range(args.steps):
            cam.Clear()
            cam.RenderScene(scene)
            if k:
                cam.WaitDownload(1)
        cam.WaitDownload(0)
        barrier()
        e2e_s, frames_in_flight = time.perf_counter() - e0, 2
        h2d, d2h = device.FrameBytes()
    except Exception as e:      # the synchronous figure stands in
        async_note = "two frames in flight failed (%s): e2e.value is the one-frame-at-a-time figure" % (repr(e)[:160],)
    finally:
        cam.AsyncDownload = False
    clocks = sampler.stop()      # sampled every 20 ms across the timed loops (device-resident + e2e)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = tris_total * args.steps / float(te.item())
    e2e_ms = float(te.item()) / args.steps * 1e3
    cam.Clear(); cam.RenderScene(scene)      # a synchronous frame for the parity check below
    if world == 1:
        flushes = cam.Flushes()
        col, dep = cam.ColorVertexList(), cam.DepthVertexList()
        parity = parity_check(desc, lambda fid: (col[flushes[fid]["first_vertex"]:flushes[fid]["first_vertex"] + flushes[fid]["vertex_count"]],
                                                 dep[flushes[fid]["first_vertex"]:flushes[fid]["first_vertex"] + flushes[fid]["vertex_count"]]))
    # the ceiling of e2e: a plain pinned D2H copy on this box, all ranks at once (they share the host's memory system)
    barrier()
    ceil = d2h_ceiling_gbs(torch)
    tc = torch.tensor([ceil], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    ceil_total = float(tc.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md / SURVEY.md 8d) ----
    peak, peak_src = load_peaks()
    n_draws = len([r for r in cam.Renders() if r["draw_index"] >= 0])
    stages = stage_table(stage_sum, args.steps, alg, peak)
    if stages["vertex"]["ms"] > 0:      # BASELINE.json's metric also asks for vertices/s of the fused vertex kernel (this rank's draws)
        stages["vertex"]["vertices_per_s"] = sum(m.mesh.n_verts for m in desc.models if m.visible) / world / (stages["vertex"]["ms"] * 1e-3)
    stages["frame_with_stage_events"] = {"ms": ms_staged / args.steps}
    if world > 1:
        stages["delivery"] = {"ms": stage_sum.get("gather_ms", 0.0) / args.steps}
    dom = max(("vertex", "cull", "sort", "emit"), key=lambda k: stages[k]["ms"])
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            tj = json.load(open(tp))
            traffic = tj.get(args.workload, {}).get(dom)
            traffic_src = tj.get("_source")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": STAGE_KERNELS[dom],
                "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "stages": stages}

    # ---- CPU baseline (oracle port) on this box's host cores, N = 1 only ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        base_desc = make_scene(args.workload)
        frames, t_cpu, split = cpu_frames(base_desc, args.cpu_seconds)
        cpu = {"value": base_desc.total_triangles() * frames / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d full frames of the same workload (%.1f s); C restatement of the single-goroutine Go path, gcc -O2 -ffp-contract=off, 1 thread of %d host cores; GOMAXPROCS n/a"
                         % (frames, t_cpu, os.cpu_count() or 0),
               "ms_per_frame": t_cpu / frames * 1e3, "split_ms": split}

    # ---- the other configurations (N = 1): driver-visible numbers next to the headline, bounded by a time budget ----
    extra = None
    if world == 1 and not args.no_extra:
        extra = {}
        for name in args.extra_workloads.split(","):
            if not name:
                continue
            if time.perf_counter() - t_start > args.extra_budget_s:
                extra[name] = {"skipped": "time budget of %.0f s for the whole run used up" % args.extra_budget_s}
                continue
            try:
                extra[name] = measure_workload(t3d, device, torch, name, peak, max(args.steps // 4, 3), flush_l2)
            except Exception as e:      # a failure here must not take the headline line with it
                extra[name] = {"error": repr(e)[:200]}

    d2h_gbs = d2h * world / (e2e_ms * 1e-3) / 1e9 if e2e_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if copies == world else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args.workload, desc, copies),
        "details": {"draws_rank0": n_draws, "surviving_triangles_rank0": surv, "l2": "flushed between steps (512 MiB write)",
                    "lighting": "default (MUFU reciprocals, colours within 1/255; t3d_set_exact_lighting restores bit-exact colours)",
                    "parallelism": ("object-wise flush sharding x%d; every shard's last kernels store its frame into rank 0's memory over NVLink (CUDA IPC peer memory) -- 36 bytes per emitted vertex, or one record per arena vertex + a 4-byte slot per emitted vertex when smaller -- and post to a mailbox there; rank 0's frame ends with the kernel that rebuilds the 96-byte colour / depth records (t3d_gather_*); no collective and no host round trip inside the step" % world) if world > 1 else "1 GPU",
                    "timing": "CUDA events on the library's launch stream around everything the frame enqueues (the first kernel fetches the uniforms over PCIe; at N > 1 the delivery to rank 0 is inside), a barrier before every step, max over ranks; roofline.stages from a second loop with an event between kernels",
                    "wall_ms_per_step_incl_flush_and_host": (wall1 - wall0) / args.steps * 1e3,
                    **({"per_rank": per_rank} if per_rank is not None else {})},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                "d2h_gbs_all_ranks": d2h_gbs, "d2h_ceiling_gbs_all_ranks": ceil_total, "frac_of_d2h_ceiling": d2h_gbs / ceil_total if ceil_total else None,
                "ceiling": "plain pinned cudaMemcpy device->host of 256 MiB, best of 5, every rank at the same time, summed",
                "frames_in_flight": frames_in_flight,
                "one_frame_at_a_time": {"value": tris_total / (sync_ms * 1e-3), "ms_per_step": sync_ms,
                                        "frac_of_d2h_ceiling": (d2h * world / (sync_ms * 1e-3) / 1e9) / ceil_total if ceil_total else None},
                **({"note": async_note} if async_note else {}),
                "path": "Camera.RenderScene -> t3d_frame_begin/submit/execute/results (pinned H2D of uniforms, pinned D2H of colour+depth lists); meshes resident; "
                        "two frames in flight: frame k's lists cross PCIe while frame k+1 is built and executed (t3d_set_async_download), the loop waits for frame k-1's lists every step"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "parity_check": parity,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if before_gather is not None:
        line["before_gather"] = before_gather
    if extra is not None:
        line["extra"] = {"workloads": extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--copies", type=int, default=0, help="copies of the workload in the scene (default: one per GPU, weak scaling); 1 at N > 1 shards ONE scene (strong scaling: configs[4] is `--workload instanced --gpus 8 --copies 1`)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configurations (extra.workloads)")
    ap.add_argument("--extra-workloads", default="stress,stress_lit,skinned_crowd,instanced_shard,particles")
    ap.add_argument("--extra-budget-s", type=float, default=200.0, help="a further workload is only started while the run is younger than this")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
