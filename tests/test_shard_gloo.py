"""Multi-GPU host logic on CPU: world_size-2 gloo run of the read-range sharding and the ordered gather
(no GPU: the per-rank "result" is the oracle's event count, standing in for the kernel output)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sigtk_b200 import synth  # noqa: E402
from sigtk_b200.shard import gather_in_read_order, shard_ranges  # noqa: E402


def test_shard_ranges_partition_and_balance():
    lens = synth.read_lengths(5000, seed=3)
    for world in (1, 2, 4, 8):
        rg = shard_ranges(lens, world)
        assert rg[0][0] == 0 and rg[-1][1] == len(lens)
        assert all(rg[k][1] == rg[k + 1][0] for k in range(world - 1))
        per = np.array([lens[a:b].sum() for a, b in rg], dtype=np.float64)
        assert per.max() / per.mean() < 1.02
    # degenerate: fewer reads than ranks, empty input
    assert [b - a for a, b in shard_ranges(np.array([10, 20]), 4)].count(0) >= 2
    assert shard_ranges(np.array([], dtype=np.int64), 2) == [(0, 0), (0, 0)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _oracle import Oracle
    orc = Oracle()
    reads = synth.make_reads(12, mean=6000.0, seed=77)
    lens = np.array([len(r[0]) for r in reads])
    lo, hi = shard_ranges(lens, world)[rank]
    mine = [int(len(orc.events(*reads[r])[0])) for r in range(lo, hi)]
    # per-rank totals: max over ranks of "time", sum of units (what bench.py does with NCCL)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(sum(lens[lo:hi]))], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((gather_in_read_order(gathered), float(t.item()), float(tot.item()), int(lens.sum())))
    dist.destroy_process_group()


def test_two_ranks_gloo_sharded_event_counts_match_single_process():
    from _oracle import Oracle
    orc = Oracle()
    reads = synth.make_reads(12, mean=6000.0, seed=77)
    expect = [int(len(orc.events(*rd)[0])) for rd in reads]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax, total, nsamp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == expect
    assert tmax == 2.0 and total == float(nsamp)
