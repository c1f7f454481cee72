#!/usr/bin/env python
"""Run under torchrun (one process per GPU): object-wise sharded render, delivered to rank 0 inside the frame (t3d_gather_*) and,
as a cross-check, by an NCCL gather afterwards; both checked against the oracle's single-process lists on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/run_multigpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import tetra3d_b200 as t3d  # noqa: E402
from tetra3d_b200 import parallel, workloads as W  # noqa: E402
import scenes  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = t3d.GeometryDevice(local)
    ok = True
    scene_list = (("soup_lit", scenes.soup_scene(1, n_models=9)), ("batch", scenes.batch_scene(1)),
                  ("million_reduced", W.million_triangle_scene(nx=250, nz=125)), ("instanced", W.instanced_scene(n=24, subdivisions=2, box=(30.0, 10.0, 30.0))),
                  ("posed", scenes.posed_scene(2, n_models=9, n_bones=6)), ("cell_lights", scenes.volume_cell_lights_scene(5, 0, skinned=True)),
                  ("subdivision", scenes.subdivision_scene(0)), ("animated_crowd", W.skinned_crowd_animated_scene(n=64)))
    # ---- delivery inside the frame: every shard's emit stores its compact records into rank 0's memory over NVLink, rank 0
    #      rebuilds the lists (t3d_gather_*, parallel.DeliveredLists) ----
    extra = (("pcuv_fog", scenes.soup_scene(6, n_models=7, cam=scenes.CameraDesc(perspective_corrected_texture_mapping=True),
                                            material=scenes.MaterialDesc(use_texture=True, tex_w=64, tex_h=32))),
             ("no_depth", scenes.render_order_scene(1)), ("empty_shard", scenes.soup_scene(3, n_models=1)))
    all_scenes = scene_list + extra
    delivered = parallel.DeliveredLists(device, local, vertex_capacity=3 * max(d.total_triangles() for _, d in all_scenes),
                                        draw_capacity=max(sum(len(m.mesh.parts) for m in d.models) for _, d in all_scenes), dst=0)
    for name, desc in all_scenes:
        scene, cam, objs = t3d.build_scene(desc, device)
        cam.SetShard(rank, world)
        cam.DownloadResults = False
        for frame in range(3):      # several frames: the inboxes are reused, the acknowledgements gate the overwrite
            cam.Clear(); cam.RenderScene(scene)
            out, tab = delivered.results(cam)
        if rank == 0:
            from oracle import binding as O
            ref = O.render_scene(desc)
            gcol = out[0].cpu().numpy().view(np.float32).reshape(-1, 12); gdep = out[1].cpu().numpy().view(np.float32).reshape(-1, 12)
            good = len(tab) == len(ref.flushes)
            for fid, f in enumerate(ref.flushes):
                first, cnt = tab[fid]
                good = good and cnt == f["vertex_count"]
                a, b = gcol[first:first + cnt], ref.color_list[f["first_vertex"]:f["first_vertex"] + cnt]
                good = good and np.array_equal(a[:, (0, 1, 2, 3, 8, 9, 10, 11)].view(np.uint32), b[:, (0, 1, 2, 3, 8, 9, 10, 11)].view(np.uint32))
                good = good and (cnt == 0 or float(np.max(np.abs(a[:, 4:8] - b[:, 4:8]))) <= 1.0 / 255)      # default lighting: colours within 1/255
                good = good and np.array_equal(gdep[first:first + cnt].view(np.uint32), ref.depth_list[f["first_vertex"]:f["first_vertex"] + cnt].view(np.uint32))
            print("%-16s world=%d delivered in-frame: flushes=%d -> %s" % (name, world, len(tab), "OK (lists bit-exact vs oracle, colours <= 1/255)" if good else "MISMATCH"))
            ok = ok and good
        dist.barrier()
    delivered.release()
    # ---- NCCL gather path ----
    for name, desc in scene_list:
        scene, cam, objs = t3d.build_scene(desc, device)
        cam.SetShard(rank, world)
        cam.DownloadResults = False
        cam.Clear(); cam.RenderScene(scene)
        col, dep = parallel.device_lists(cam, local)
        table = [(f["global_flush"], f["vertex_count"]) for f in cam.Flushes()]
        out, tab = parallel.gather_lists([col, dep], table, dst=0)
        if rank == 0:
            from oracle import binding as O
            ref = O.render_scene(desc)
            gcol = out[0].cpu().numpy().view(np.float32).reshape(-1, 12); gdep = out[1].cpu().numpy().view(np.float32).reshape(-1, 12)
            good = len(tab) == len(ref.flushes) and len(gcol) == ref.n_vertices
            for fid, f in enumerate(ref.flushes):
                first, cnt = tab[fid]
                good = good and cnt == f["vertex_count"]
                a = gcol[first:first + cnt]; b = ref.color_list[f["first_vertex"]:f["first_vertex"] + cnt]
                good = good and np.array_equal(a[:, (0, 1, 2, 3, 8, 9, 10, 11)].view(np.uint32), b[:, (0, 1, 2, 3, 8, 9, 10, 11)].view(np.uint32))
                good = good and (cnt == 0 or float(np.max(np.abs(a[:, 4:8] - b[:, 4:8]))) <= 1.0 / 255)
                a = gdep[first:first + cnt]; b = ref.depth_list[f["first_vertex"]:f["first_vertex"] + cnt]
                good = good and np.array_equal(a.view(np.uint32), b.view(np.uint32))
            owners = cam.FlushOwnerRanks()
            print("%-16s world=%d flushes=%d vertices=%d per-rank flushes=%s -> %s" % (name, world, len(tab), len(gcol), [owners.count(r) for r in range(world)], "OK (lists bit-exact vs oracle, colours <= 1/255)" if good else "MISMATCH"))
            ok = ok and good
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
