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


class _DevMem:
    """__cuda_array_interface__ holder so torch can view library-owned device memory without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """All-gather of equal per-rank float64 blocks through peer memory (one process per GPU of one NVLink node).

    Every rank owns `nbuf` gathered buffers of world*block doubles (allocated by the library, exported over CUDA IPC and
    mapped by every peer). `push(src_ptr, buf)` copies this rank's block into slot `rank` of buffer `buf` on every rank,
    asynchronously to the evaluation stream (copy engines by default), so the gather of evaluation i overlaps the kernels
    of evaluation i+1. `exchange` is the only place that needs a process group (handles are 64-byte blobs).
    Completion across processes is the caller's barrier (in multi-start optimisation: once per optimiser iteration)."""

    def __init__(self, ctx, rank, world, block_doubles, nbuf=2, mode=0, exchange=None):
        self.ctx, self.rank, self.world, self.block, self.mode = ctx, rank, world, int(block_doubles), int(mode)
        self.own, self.peers = [], []
        handles = []
        for _ in range(nbuf):
            p, h = ctx.peer_buffer_create(world * self.block * 8)
            self.own.append(p)
            handles.append(h)
        if exchange is None:
            import torch.distributed as dist

            def exchange(obj):
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
        every = exchange(handles)                       # every[r][b] = handle of rank r's buffer b
        self._opened = []
        for b in range(nbuf):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(self.own[b])
                else:
                    q = ctx.peer_buffer_open(every[r][b])
                    self._opened.append(q)
                    row.append(q)
            self.peers.append(row)

    def push(self, src_ptr, buf, slot=None):
        self.ctx.gather_to_peers(src_ptr, self.block * 8, self.peers[buf], self.rank * self.block * 8,
                                 slot=buf & 1 if slot is None else slot, mode=self.mode)

    def wait(self, slot):
        self.ctx.gather_wait(slot)

    def scatter_targets(self, buf, include_self=True):
        """(device pointers of buffer `buf` on the ranks to write to, byte offset of this rank's slot in each): the arguments of
        Problem.cost_and_grad_batch_dev_scatter.  With include_self the rank's own gathered buffer is one of the destinations, so
        after the call every buffer -- local and remote -- holds this rank's block at the same offset."""
        targets = [q for r, q in enumerate(self.peers[buf]) if include_self or r != self.rank]
        return targets, self.rank * self.block * 8

    def wait_on(self, slot, cuda_stream):
        """`cuda_stream` waits for this rank's pushes of `slot` (then a barrier there = cross-rank completion)."""
        self.ctx.gather_wait_on(slot, cuda_stream)

    def view(self, buf, device):
        """torch view (no copy) of this rank's gathered buffer `buf`: world*block doubles."""
        import torch
        return torch.as_tensor(_DevMem(self.own[buf], self.world * self.block), device=device)

    def close(self, barrier=None):
        """Unmap the peers' buffers, then (after `barrier()`, so that no peer still maps them) free this rank's."""
        for q in self._opened:
            self.ctx.peer_buffer_close(q)
        self._opened = []
        if barrier is not None:
            barrier()
        for p in self.own:
            self.ctx.peer_buffer_destroy(p)
        self.own = []
