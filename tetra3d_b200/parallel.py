"""Object-wise multi-GPU: every rank renders the flushes Camera.SetShard assigned to it; the emitted lists
end up on the rank that issues the draws (no cross-GPU merge is needed: each MeshPart is sorted in isolation,
model.go:602, 989).  Two ways to get them there:

  SharedLists   the issuing rank's lists are mapped into every rank (CUDA IPC) and each rank's emit kernel
                writes its records straight into its own region over NVLink -- the transfer is the kernel's
                TMA stores, there is no separate copy; only a small per-flush count vector is all-reduced.
  gather_lists  the lists stay local and are collected afterwards with NCCL send / recv (also what the gloo
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


class SharedLists:
    """Colour / depth lists of the issuing rank `dst`, one fixed-capacity region per rank, mapped into every rank.

    After construction every rank's GeometryDevice emits into region `rank` (t3d_set_output_target).  `collect`
    exchanges the per-flush vertex counts and, on `dst`, returns views of the complete lists plus the table
    global_flush_id -> (first_vertex, vertex_count): region r starts at vertex r * capacity, a rank's flushes
    sit there in ascending global id.
    """

    def __init__(self, device, device_index: int, vertex_capacity: int, dst: int = 0, group=None, compact_depth: bool = True,
                 render_depth: bool = True):
        self.group, self.dst, self.device = group, dst, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = torch.device("cuda", device_index)
        cap = torch.tensor([int(vertex_capacity)], dtype=torch.int64, device=self.dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self.capacity = (int(cap.item()) + 255) // 256 * 256
        nbytes = self.world * self.capacity * 48
        payload = [None]
        if self.rank == dst:
            self.color_ptr, hc = device.SharedAlloc(nbytes)
            self.depth_ptr, hd = device.SharedAlloc(nbytes)
            payload = [(hc, hd)]
        dist.broadcast_object_list(payload, src=dst, group=group)
        if self.rank != dst:
            hc, hd = payload[0]
            self.color_ptr = device.SharedOpen(hc)
            self.depth_ptr = device.SharedOpen(hd)
        off = self.rank * self.capacity * 48
        device.SetOutputTarget(self.color_ptr + off, self.depth_ptr + off, 0, self.capacity)
        # compact depth: the depth list is the colour list with (d, d, d, 1) for a colour, so the other ranks ship one float per
        # vertex instead of the second list (52 instead of 96 bytes per vertex into dst) and dst rebuilds the records
        self.compact = bool(compact_depth) and self.world > 1
        self.render_depth = bool(render_depth)
        if self.compact:
            payload = [None]
            if self.rank == dst:
                self.scalar_ptr, hs = device.SharedAlloc(self.world * self.capacity * 4)
                payload = [hs]
            dist.broadcast_object_list(payload, src=dst, group=group)
            if self.rank != dst:
                self.scalar_ptr = device.SharedOpen(payload[0])
                device.SetOutputCompactDepth(self.scalar_ptr + self.rank * self.capacity * 4)
        self.nbytes = nbytes
        self._meta_host = None
        self._meta_dev = None

    def collect(self, flush_table: List[Tuple[int, int]], n_flushes_total: int):
        """Call after the frame has completed on this rank (Camera.RenderScene with DownloadResults = False waits
        for it).  The all-reduce doubles as the barrier: when it returns on `dst`, every rank's emit has finished."""
        if self._meta_host is None or self._meta_host.numel() != n_flushes_total:
            self._meta_host = torch.zeros(n_flushes_total, dtype=torch.int64).pin_memory()
            self._meta_dev = torch.zeros(n_flushes_total, dtype=torch.int64, device=self.dev)
        mh = self._meta_host.numpy()
        mh[:] = 0
        if flush_table:
            import numpy as np
            ft = np.asarray(flush_table, dtype=np.int64)
            mh[ft[:, 0]] = ((self.rank + 1) << 40) + ft[:, 1]
        self._meta_dev.copy_(self._meta_host, non_blocking=True)
        dist.all_reduce(self._meta_dev, op=dist.ReduceOp.SUM, group=self.group)
        if self.rank != self.dst:
            return None, None
        meta = self._meta_dev.cpu().tolist()
        table: Dict[int, Tuple[int, int]] = {}
        cursor = [r * self.capacity for r in range(self.world)]
        for fid, v in enumerate(meta):
            if v == 0:
                continue
            owner, cnt = (v >> 40) - 1, v & ((1 << 40) - 1)
            table[fid] = (cursor[owner], int(cnt))
            cursor[owner] += int(cnt)
        if self.compact:        # rebuild the other ranks' depth records from their colour records and depth values
            for r in range(self.world):
                n = cursor[r] - r * self.capacity
                if r != self.dst and n:
                    self.device.ExpandDepthList(self.color_ptr + r * self.capacity * 48, self.scalar_ptr + r * self.capacity * 4,
                                                self.depth_ptr + r * self.capacity * 48, n, self.render_depth)
        lists = [torch.as_tensor(DeviceArray(self.color_ptr, self.nbytes), device=self.dev),
                 torch.as_tensor(DeviceArray(self.depth_ptr, self.nbytes), device=self.dev)]
        return lists, table

    def release(self):
        self.device.SetOutputTarget(0, 0, 0, 0)      # also drops the compact-depth redirection
