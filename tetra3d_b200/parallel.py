"""Object-wise multi-GPU: every rank renders the flushes Camera.SetShard assigned to it; the emitted lists
end up on the rank that issues the draws (no cross-GPU merge is needed: each MeshPart is sorted in isolation,
model.go:602, 989).  Two ways to get them there:

  DeliveredLists  the library delivers inside the frame (t3d_gather_*): each shard's last kernels store its frame straight
                  into the issuing rank's memory over NVLink -- 36 bytes per emitted vertex, or one such record per arena
                  vertex plus a 4-byte slot per emitted vertex when that is smaller -- and post to a mailbox there; the
                  issuing rank's frame ends with the kernel that rebuilds the lists.  No collective, no host round trip per
                  frame.
  gather_lists    the lists stay local and are collected afterwards with NCCL send / recv (also what the gloo
                  CPU tests exercise).

torch.distributed is plumbing only (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


class DeviceArray:
    """Zero-copy view of library-owned device memory for torch (``torch.as_tensor(DeviceArray(...))``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_lists(cam, device_index: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(colour list, depth list) of the last Render as uint8 CUDA tensors over the library's buffers."""
    p = cam.DevicePointers()
    assert p["on_device"], "Camera.DownloadResults must be False for device-resident gathers"
    n = int(p["n_vertices"]) * 48
    dev = torch.device("cuda", device_index)
    if n == 0:
        e = torch.empty(0, dtype=torch.uint8, device=dev)
        return e, e
    return (torch.as_tensor(DeviceArray(p["color"], n), device=dev), torch.as_tensor(DeviceArray(p["depth"], n), device=dev))


def gather_lists(local: List[torch.Tensor], flush_table: List[Tuple[int, int]], dst: int = 0, group=None,
                 out: Optional[List[torch.Tensor]] = None, n_flushes_total: Optional[int] = None):
    """Gathers variable-size per-rank lists to `dst`.

    local:        this rank's lists (same byte length on one rank, e.g. [colour, depth]), uint8
    flush_table:  this rank's (global_flush_id, vertex_count) in local order
    Returns on dst: (lists, table) where lists[i] is the concatenation over ranks 0..W-1 of list i and table maps
    global_flush_id -> (first_vertex, vertex_count) in the concatenated lists.  Other ranks get (None, None).

    The only metadata exchanged is one small all_reduce: a [W + F] int64 vector holding every rank's byte count
    and every flush's (owner rank + 1, vertex count) packed as owner * 2^40 + count.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nbytes = int(local[0].numel())
    dev = local[0].device
    if n_flushes_total is None:
        m = torch.tensor([max([f for f, _ in flush_table], default=-1) + 1], dtype=torch.int64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        n_flushes_total = int(m.item())
    meta = torch.zeros(world + n_flushes_total, dtype=torch.int64)
    meta[rank] = nbytes
    for fid, cnt in flush_table:
        meta[world + fid] = ((rank + 1) << 40) + int(cnt)
    meta = meta.to(dev)
    dist.all_reduce(meta, op=dist.ReduceOp.SUM, group=group)
    meta = meta.cpu().tolist()
    sizes = meta[:world]
    offsets = [0]
    for sz in sizes:
        offsets.append(offsets[-1] + sz)
    if rank == dst:
        total = offsets[-1]
        if out is None or out[0].numel() < total:
            out = [torch.empty(total, dtype=torch.uint8, device=dev) for _ in local]
        ops = []
        for i, t in enumerate(local):
            if nbytes:
                out[i][offsets[rank]:offsets[rank] + nbytes].copy_(t)
            for r in range(world):
                if r != rank and sizes[r]:
                    ops.append(dist.P2POp(dist.irecv, out[i][offsets[r]:offsets[r] + sizes[r]], r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        table: Dict[int, Tuple[int, int]] = {}
        cursor = [offsets[r] // 48 for r in range(world)]
        for fid in range(n_flushes_total):
            v = meta[world + fid]
            if v == 0:
                continue
            owner, cnt = (v >> 40) - 1, v & ((1 << 40) - 1)
            table[fid] = (cursor[owner], int(cnt))      # a rank's flushes are stored in ascending global id
            cursor[owner] += int(cnt)
        return [o[:total] for o in out], table
    ops = [dist.P2POp(dist.isend, t, dst, group) for t in local if nbytes]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None, None


def delivered_flush_table(owners: List[int], draw_flush: List[List[int]], survivors: List, capacity: int) -> Dict[int, Tuple[int, int]]:
    """Where every flush's vertices are in the delivered lists (t3d_gather_*): global flush id -> (first_vertex, vertex_count).

    owners:      Camera.FlushOwnerRanks(): for every flush of the frame, the rank that ran it
    draw_flush:  Camera.ShardDrawFlushes(): per rank, the flush of every draw it submitted, in submit order
    survivors:   per rank, the surviving triangles of every submitted draw (t3d_gather_rank_stats, column 2)
    capacity:    vertices per rank region: rank r's vertices start at r * capacity, its flushes follow each other in ascending id
    """
    import numpy as np
    table: Dict[int, Tuple[int, int]] = {}
    n_flush = len(owners)
    for r in range(len(draw_flush)):
        dfl = np.asarray(draw_flush[r], dtype=np.int64)
        sv = np.asarray(survivors[r], dtype=np.int64).reshape(-1)
        assert len(dfl) == len(sv), "one survivor count per submitted draw"
        counts = np.bincount(dfl, weights=3 * sv, minlength=n_flush).astype(np.int64) if len(dfl) else np.zeros(n_flush, np.int64)
        cursor = r * int(capacity)
        for f in range(n_flush):
            if owners[f] == r:
                table[f] = (cursor, int(counts[f]))
                cursor += int(counts[f])
    return table


class DeliveredLists:
    """The draw lists of a sharded frame on the rank `dst` that issues the draws, delivered inside the frame by the library
    (include/t3dgeo.h: t3d_gather_*): every shard stores a compact or indexed form of its frame straight into dst's memory over
    NVLink and posts the frame to a mailbox there; dst's frame ends with the kernel that rebuilds the 96-byte colour / depth
    records.  No collective and no host round trip per frame: torch.distributed only carries the three CUDA IPC
    handles at construction.  All ranks render the same number of frames, in lockstep.
    """

    def __init__(self, device, device_index: int, vertex_capacity: int, draw_capacity: int, dst: int = 0, group=None):
        self.group, self.dst, self.device = group, dst, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = torch.device("cuda", device_index)
        cap = torch.tensor([int(vertex_capacity), int(draw_capacity)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self.capacity = (int(cap[0].item()) + 255) // 256 * 256
        self.draw_capacity = max(int(cap[1].item()), 1)
        payload = [None]
        if self.rank == dst:
            payload = [device.GatherAlloc(self.world, self.capacity, self.draw_capacity)]
        dist.broadcast_object_list(payload, src=dst, group=group)
        device.GatherAttach(self.world, self.rank, dst, self.capacity, self.draw_capacity, None if self.rank == dst else payload[0])
        dist.barrier(group=group)
        self.attached = True

    def results(self, cam):
        """On dst, after Camera.RenderScene: ([colour list, depth list] as uint8 CUDA tensors over the whole shared allocation,
        {global flush id: (first_vertex, vertex_count)}).  Waits for every shard's frame.  Other ranks: (None, None)."""
        if self.rank != self.dst:
            return None, None
        res = self.device.GatherResults()
        draw_flush = cam.ShardDrawFlushes()
        surv = [self.device.GatherRankStats(r, len(draw_flush[r]))[:, 2] for r in range(self.world)]
        table = delivered_flush_table(cam.FlushOwnerRanks(), draw_flush, surv, res["capacity"])
        for r in range(self.world):
            n = sum(c for f, (_, c) in table.items() if cam.FlushOwnerRanks()[f] == r)
            assert n == res["n_vertices"][r], ("rank %d delivered %d vertices, its statistics add up to %d" % (r, res["n_vertices"][r], n))
        nbytes = self.world * res["capacity"] * 48
        lists = [torch.as_tensor(DeviceArray(res["color"], nbytes), device=self.dev), torch.as_tensor(DeviceArray(res["depth"], nbytes), device=self.dev)]
        return lists, table

    def release(self):
        if self.attached:
            dist.barrier(group=self.group)
            self.device.GatherDetach()
            dist.barrier(group=self.group)
            self.attached = False
