"""Object-wise multi-GPU: every rank renders the flushes Camera.SetShard assigned to it; the emitted lists
are gathered to the rank that issues the draws (no cross-GPU merge is needed: each MeshPart is sorted in
isolation, model.go:602, 989).  torch.distributed is plumbing only (NCCL on GPUs, gloo in CPU tests)."""
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
                 out: Optional[List[torch.Tensor]] = None):
    """Gathers variable-size per-rank lists to `dst`.

    local:        this rank's lists (same count and byte length on one rank, e.g. [colour, depth]), uint8
    flush_table:  this rank's (global_flush_id, vertex_count) in local order
    Returns on dst: (lists, table) where lists[i] is the concatenation over ranks 0..W-1 of list i and table maps
    global_flush_id -> (first_vertex, vertex_count) in the concatenated lists.  Other ranks get (None, None).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nbytes = int(local[0].numel())
    dev = local[0].device
    # sizes + flush tables (tiny, host-side objects)
    meta = [None] * world
    dist.all_gather_object(meta, (nbytes, list(flush_table)), group=group)
    sizes = [m[0] for m in meta]
    offsets = [0]
    for s in sizes:
        offsets.append(offsets[-1] + s)
    if rank == dst:
        total = offsets[-1]
        if out is None:
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
        for r in range(world):
            v = offsets[r] // 48
            for fid, cnt in meta[r][1]:
                table[int(fid)] = (v, int(cnt))
                v += int(cnt)
        return out, table
    ops = [dist.P2POp(dist.isend, t, dst, group) for t in local if nbytes]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None, None
