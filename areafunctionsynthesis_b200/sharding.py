"""Host-side sharding of a batch of independent units (utterances / area functions) over ranks.

The hot path has no exchange step: every rank runs the kernels on a contiguous slice of the
batch; the only collective is the gather of OUTPUTS (NCCL over NVLink on GPUs, gloo in the CPU tests).
One process per GPU, `torch.distributed` for the plumbing.
"""
import numpy as np


def shard_bounds(n_units, world_size):
    """Contiguous ranges [lo, hi) per rank; sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(int(n_units), int(world_size))
    bounds, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard(array, rank, world_size):
    lo, hi = shard_bounds(len(array), world_size)[rank]
    return array[lo:hi]


def gather_outputs(local, n_units, dist, device=None):
    """all_gather of per-unit outputs `local` [n_local, ...] (torch tensor) into [n_units, ...] on every rank.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    world = dist.get_world_size()
    bounds = shard_bounds(n_units, world)
    n_max = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device if device is None else device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=0)


def run_sharded(fn, inputs, rank, world_size):
    """Apply `fn` (a callable on a slice of the leading batch axis of every array in `inputs`) to this rank's shard."""
    return fn(*[shard(np.asarray(x), rank, world_size) for x in inputs])
