"""Multi-GPU sharding of the batched hot path: independent pulses (or response-function frequencies)
are split contiguously over ranks; the only inter-GPU traffic is one all-gather of the per-shard
[cost | grad] block per evaluation (no traffic inside a pulse's time scan)."""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n units for `rank` of `world`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(cost, grad):
    """[cost (Bs) | grad (Bs, nx) row per pulse] as one contiguous float64 block (the all-gather payload)."""
    cost = np.asarray(cost, dtype=np.float64)
    grad = np.asarray(grad, dtype=np.float64)          # (nx, Bs) column per pulse
    return np.concatenate([cost, np.ascontiguousarray(grad.T).reshape(-1)])


def unpack_results(gathered, world, bs, nx):
    """Inverse of pack_results over `world` equal shards: cost (B,), grad (B, nx)."""
    blk = np.asarray(gathered).reshape(world, bs * (1 + nx))
    cost = blk[:, :bs].reshape(-1)
    grad = blk[:, bs:].reshape(world * bs, nx)
    return cost, grad
