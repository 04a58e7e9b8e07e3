"""N > 1 host logic on CPU: contiguous sharding + output gather with world_size 2 over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from areafunctionsynthesis_b200.sharding import gather_outputs, shard, shard_bounds


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 4096, 65536 + 3):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = np.arange(n_units, dtype=np.float64)
    mine = shard(units, rank, world)
    # stand-in for the per-rank kernel: a pure per-unit function (the real one needs a GPU)
    out = torch.from_numpy(np.stack([mine * 2.0, mine + 0.5], axis=1))
    full = gather_outputs(out, n_units, dist)
    ok = torch.equal(full, torch.from_numpy(np.stack([units * 2.0, units + 0.5], axis=1)))
    # timing convention of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == world
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [9, 4096])
def test_two_rank_gloo_gather(n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_units % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
