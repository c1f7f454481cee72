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
    desc = make_scene(args.workload)
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
        "config": {"workload": args.workload, "triangles_per_frame": tris, "vertices": desc.total_vertices()},
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
def run_ours(args):
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

    desc = make_scene(args.workload, copies=world)
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

    shared = [None]

    def device_step(timed):
        """One device-resident frame; returns (ms of this rank, stage timings).  At N > 1 every rank's emit kernel
        writes straight into rank 0's lists over NVLink (parallel.SharedLists), so the transfer is inside the
        frame's own event pair; what remains is the all-reduce of the per-flush counts, timed as gather_ms."""
        flush_l2()
        cam.Clear()
        cam.RenderScene(scene)
        tm = cam.Timings()
        ms = tm["total_ms"]
        if world > 1:
            table = [(f["global_flush"], f["vertex_count"]) for f in cam.Flushes()]
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            shared[0].collect(table, len(cam.FlushOwnerRanks()))
            e1.record(); torch.cuda.synchronize()
            tm = dict(tm); tm["gather_ms"] = e0.elapsed_time(e1)
            ms += tm["gather_ms"]
        return ms, tm

    # ---- device-resident value: kernels chained with programmatic dependent launches, one event pair per frame ----
    cam.DownloadResults = False
    device.SetStageTimings(False)
    if world > 1:
        cam.Clear(); cam.RenderScene(scene)          # sizes the shared lists: 3 vertices per triangle this rank draws
        shared[0] = parallel.SharedLists(device, local, 3 * sum(r["total_tris"] for r in cam.Renders()), dst=0)
    for _ in range(max(args.warmup, 3)):
        device_step(False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    ms_sum = 0.0
    gather_sum = 0.0
    for _ in range(args.steps):
        ms, tm = device_step(True)
        ms_sum += ms
        gather_sum += tm.get("gather_ms", 0.0)
    barrier()
    wall1 = time.perf_counter()
    # ---- per-stage durations for the roofline: same frames with a CUDA event between the kernels (which
    #      serialises them, so the stages add up to more than the chained frame above) ----
    device.SetStageTimings(True)
    stage_sum = {}
    for _ in range(3):
        device_step(False)
    ms_staged = 0.0
    for _ in range(args.steps):
        ms, tm = device_step(True)
        ms_staged += tm["total_ms"]
        for k, v in tm.items():
            stage_sum[k] = stage_sum.get(k, 0.0) + v
    stage_sum["gather_ms"] = gather_sum
    device.SetStageTimings(False)
    barrier()
    launches_per_step = cam.LaunchCount()
    n_vertices_out = cam.VertexCount()
    vbytes = device.VertexKernelBytes()
    h2d_dev, _ = device.FrameBytes()
    t = torch.tensor([ms_sum], dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = torch.tensor([(ms_sum - gather_sum) / args.steps, gather_sum / args.steps], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"frame_ms": float(a[0]), "count_exchange_ms": float(a[1])} for a in allr]
    ms_per_step = float(t.item()) / args.steps
    value = tris_total / (ms_per_step * 1e-3)

    # ---- e2e through the public API with host buffers ----
    before_gather = None
    if world > 1:
        # the same frames with every rank's lists staying in its own HBM: the geometry stage before the gather
        # (SURVEY.md 8e asks for both figures; `value` above is the one after delivery to the issuing rank)
        shared[0].release()
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
    e2e_s = time.perf_counter() - e0
    clocks = sampler.stop()      # sampled every 20 ms across both timed loops (device-resident + e2e)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = tris_total * args.steps / float(te.item())
    e2e_ms = float(te.item()) / args.steps * 1e3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: DESIGN.md / SURVEY.md 8d) ----
    peak, peak_src = load_peaks()
    surv = n_vertices_out // 3
    local_tris = sum(r["total_tris"] for r in cam.Renders())
    n_draws = len([r for r in cam.Renders() if r["draw_index"] >= 0])
    passes = 1          # one segmented counting-sort pass on the 9-bit bin (the draw ordinal is already the arena order)
    alg = {
        "vertex": vbytes,
        "cull": 72 * local_tris + 8 * surv,
        "sort": (16 * passes + 4) * surv,
        "emit": (600 if desc.camera.render_normals else 456) * surv,
    }
    stages = {}
    for k in ("vertex", "cull", "sort", "emit"):
        ms_k = stage_sum.get(k + "_ms", 0.0) / args.steps
        gbs = alg[k] / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
        stages[k] = {"ms": ms_k, "algorithmic_bytes": alg[k], "achieved_gbs": gbs, "frac": gbs / peak}
    if stages["vertex"]["ms"] > 0:      # BASELINE.json's metric also asks for vertices/s of the fused vertex kernel (this rank's draws)
        stages["vertex"]["vertices_per_s"] = sum(m.mesh.n_verts for m in desc.models if m.visible) / world / (stages["vertex"]["ms"] * 1e-3)
    stages["prep"] = {"ms": stage_sum.get("prep_ms", 0.0) / args.steps}
    stages["frame_with_stage_events"] = {"ms": ms_staged / args.steps}
    if world > 1:
        stages["gather"] = {"ms": stage_sum.get("gather_ms", 0.0) / args.steps}
    dom = max(("vertex", "cull", "sort", "emit"), key=lambda k: stages[k]["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": {"vertex": "k_vertex", "cull": "k_cull", "sort": "k_tile_scan+k_bin_hist+k_scatter", "emit": "k_emit"}[dom],
                "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic,
                "peak_source": peak_src, "stages": stages}

    # ---- CPU baseline (oracle port) on this box's host cores, N = 1 only ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        base_desc = make_scene(args.workload)
        frames, t_cpu, split = cpu_frames(base_desc, args.cpu_seconds)
        cpu = {"value": base_desc.total_triangles() * frames / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d full frames of the same workload (%.1f s); C restatement of the single-goroutine Go path, gcc -O2 -ffp-contract=off, 1 thread of %d host cores; GOMAXPROCS n/a"
                         % (frames, t_cpu, os.cpu_count() or 0),
               "ms_per_frame": t_cpu / frames * 1e3, "split_ms": split}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "triangles_per_frame": tris_total, "vertices": desc.total_vertices(), "draws_rank0": n_draws,
                   "surviving_triangles_rank0": surv, "l2": "flushed between steps (512 MiB write)",
                   "parallelism": "object-wise flush sharding x%d; every rank's emit kernel stores its colour list and one depth value per vertex into rank 0's memory over NVLink (CUDA IPC peer memory), rank 0 rebuilds the depth records; per-flush counts by NCCL all-reduce" % world if world > 1 else "1 GPU",
                   "timing": "CUDA events on the library's launch stream around the frame's kernels (the first one fetches the uniforms over PCIe), max over ranks; roofline.stages from a second loop with an event between kernels",
                   "wall_ms_per_step_incl_flush_and_host": (wall1 - wall0) / args.steps * 1e3,
                   **({"per_rank": per_rank} if per_rank is not None else {})},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                "path": "Camera.RenderScene -> t3d_frame_begin/submit/execute/results (pinned H2D of uniforms, pinned D2H of colour+depth lists); meshes resident"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if before_gather is not None:
        line["before_gather"] = before_gather
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
