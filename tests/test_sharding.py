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


@pytest.mark.gpu
def test_device_group_equals_single_device(shapes):
    """afs_group_*: the C-ABI multi-device layer (one process, one host thread per device, contiguous shards, every
    device writing its rows of the caller's arrays).  Bit-identical to the single-device call on every visible
    device count, ragged shards included; the shard rule is sharding.shard_bounds'."""
    import torch
    import areafunctionsynthesis_b200 as afs
    from areafunctionsynthesis_b200.sharding import shard_bounds
    n_dev = torch.cuda.device_count()
    names = shapes["names"]
    rng = np.random.default_rng(11)
    B = 301  # ragged over 2, 4 and 8 devices
    idx = rng.integers(0, 16, size=(B, 2))
    fr = shapes["frames"][idx].copy()
    fr["velum_area_cm2"] = rng.random((B, 2)) * 0.5
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (B, 2, 1))
    gl[:, :, 0] = rng.uniform(90, 220, size=(B, 2))
    seeds = rng.integers(1, 2**31 - 1, size=B).astype(np.uint32)
    one = afs.Context(0)
    a1, f1, s1 = one.tds_synthesize(fr, gl, [120], 44100.0, 2, seeds=seeds, want_flow=True)
    re1, im1 = one.tl_sweep(fr[:, 0], 22050.0 / 2048, 512)
    one.close()
    g_one = afs.Context(0)
    for devices in (None, [0], list(range(n_dev))[::-1]):
        g = afs.Group(devices)
        G = len(g)
        assert G == (n_dev if devices is None else len(devices))
        assert [g.shard(B, r) for r in range(G)] == shard_bounds(B, G)
        a, f, s = g.tds_synthesize(fr, gl, [120], 44100.0, 2, seeds=seeds, want_flow=True)
        assert np.array_equal(a, a1) and np.array_equal(f, f1) and np.array_equal(s, s1)
        re, im = g.tl_sweep(fr[:, 0], 22050.0 / 2048, 512)
        assert np.array_equal(re, re1) and np.array_equal(im, im1)
        # chunked synthesis over the group: every device keeps the models of its shard between calls
        gs = afs.TdsStream(g, B, 44100.0, 2, seeds=seeds)
        p1, _, _ = gs.synthesize(fr, gl, [0, 120])
        blob = gs.save()                                   # checkpoint of the group stream -> a new one continues bit for bit
        p2, _, _ = gs.synthesize(fr[:, :1], gl[:, :1], [60])
        gs.close()
        gs = afs.TdsStream(g, B, 44100.0, 2)
        gs.load(blob)
        assert gs.samples == 120
        p3, _, _ = gs.synthesize(fr[:, :1], gl[:, :1], [60])
        gs.close()
        assert np.array_equal(p2, p3)
        ss = afs.TdsStream(g_one, B, 44100.0, 2, seeds=seeds)
        q1, _, _ = ss.synthesize(fr, gl, [0, 120])
        q2, _, _ = ss.synthesize(fr[:, :1], gl[:, :1], [60])
        ss.close()
        assert np.array_equal(p1, q1) and np.array_equal(p2, q2) and np.abs(p2).max() > 0
        # fewer units than devices, empty batch, pinned output
        a, _, _ = g.tds_synthesize(fr[:1], gl[:1], [120], 44100.0, 2, seeds=seeds[:1])
        assert np.array_equal(a, a1[:1])
        a, _, _ = g.tds_synthesize(fr[:0], gl[:0], [120], 44100.0, 2)
        assert a.shape == (0, 120)
        pin = afs.pinned_empty((B, 120))
        g.tds_synthesize(fr, gl, [120], 44100.0, 2, seeds=seeds, audio=pin)
        assert np.array_equal(pin, a1)
        afs.pinned_free(pin)
        g.close()
    g_one.close()
    with pytest.raises(afs.AfsError):
        afs.Group([0, 0])
