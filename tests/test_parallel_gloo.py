"""Multi-GPU host logic on CPU: flush sharding and the gather of variable-size draw lists (gloo, world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tetra3d_b200 as t3d
from tetra3d_b200 import parallel


def test_assignment_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    tris = [int(x) for x in rng.integers(1, 20000, 200)]
    for ranks in (1, 2, 4, 8):
        owner = t3d.AssignFlushesToRanks(tris, ranks)
        assert len(owner) == len(tris) and set(owner) <= set(range(ranks))
        load = [sum(t for t, o in zip(tris, owner) if o == r) for r in range(ranks)]
        assert max(load) - min(load) <= max(tris)            # greedy largest-first bound
        assert owner == t3d.AssignFlushesToRanks(tris, ranks)  # deterministic: every rank computes the same map
    assert t3d.AssignFlushesToRanks([], 4) == []
    assert t3d.AssignFlushesToRanks([5, 5, 5], 8) == [0, 1, 2]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, counts, owner, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every flush f emits counts[f] vertices of 48 bytes whose bytes are all (f % 251)
        mine = [f for f in range(len(counts)) if owner[f] == rank]
        def make(off):
            parts = [torch.full((counts[f] * 48,), (f + off) % 251, dtype=torch.uint8) for f in mine]
            return torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8)
        col, dep = make(0), make(7)
        table = [(f, counts[f]) for f in mine]
        out, tab = parallel.gather_lists([col, dep], table, dst=0)
        if rank == 0:
            ok = set(tab.keys()) == set(range(len(counts)))
            for f, (first, cnt) in tab.items():
                ok = ok and cnt == counts[f]
                ok = ok and bool((out[0][first * 48:(first + cnt) * 48] == f % 251).all())
                ok = ok and bool((out[1][first * 48:(first + cnt) * 48] == (f + 7) % 251).all())
            ok = ok and out[0].numel() == 48 * sum(counts)
            ret.put(bool(ok))
        else:
            assert out is None and tab is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[30, 0, 12, 999, 3, 3, 60, 0], [0, 0], [5]])
def test_gather_lists_world_size_2(counts):
    world = 2
    owner = t3d.AssignFlushesToRanks([max(c, 1) for c in counts], world)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, owner, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_delivered_flush_table():
    """The slicing of delivered lists by flush (parallel.delivered_flush_table) from the owners, every rank's draw -> flush map
    and the survivor counts the shards ship."""
    owners = [0, 1, 0, 1, 1, 0]
    draw_flush = [[0, 0, 2, 5], [1, 3, 3, 3]]           # flush 4 (rank 1) submitted nothing
    surv = [[2, 1, 0, 7], [4, 1, 0, 2]]
    t = parallel.delivered_flush_table(owners, draw_flush, surv, capacity=1000)
    assert t == {0: (0, 9), 2: (9, 0), 5: (9, 21), 1: (1000, 12), 3: (1012, 9), 4: (1021, 0)}
    assert parallel.delivered_flush_table([], [[], []], [[], []], 256) == {}
