"""The N>1 path of bench.py on CPU: world_size=2 over gloo.  Pulses are sharded contiguously over
ranks, each rank evaluates its shard (the C++ port stands in for the GPU kernels here), and one
all-gather of [cost | grad] reproduces the single-process result."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, B, N, out):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robustgrape_b200.sharding import shard_range, pack_results, unpack_results
    from cases import cz_problem
    from oracle import cpu_port
    import bench
    X = bench.make_pulses(N, B)
    lo, hi = shard_range(B, rank, world)
    pp = cpu_port.PortProblem(cz_problem(N, 7.613 * N / 1000))
    cost, grad = pp.cost_and_grad_batch(X[lo:hi].T, (), 1)
    local = torch.from_numpy(pack_results(cost, grad))
    gathered = torch.empty(world * local.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, local)
    c_all, g_all = unpack_results(gathered.numpy(), world, hi - lo, N + 1)
    if rank == 0:
        np.savez(out, cost=c_all, grad=g_all)
    dist.destroy_process_group()


def test_two_rank_gloo_allgather(tmp_path):
    sys.path.insert(0, str(ROOT))
    import bench
    from cases import cz_problem
    from oracle import cpu_port
    B, N = 8, 20
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, 29533, B, N, out), nprocs=2, join=True)
    z = np.load(out)
    pp = cpu_port.PortProblem(cz_problem(N, 7.613 * N / 1000))
    c, g = pp.cost_and_grad_batch(bench.make_pulses(N, B).T, (), 1)
    assert np.array_equal(z["cost"], c)
    assert np.array_equal(z["grad"], g.T)


def test_shard_range_covers_batch():
    from robustgrape_b200.sharding import shard_range
    for B in (1, 7, 8192, 4096 + 3):
        for w in (1, 2, 4, 8):
            r = [shard_range(B, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


class _FakeCtx:
    """Stands in for robustgrape_b200._lib.Context on a CPU box: 'device memory' is a numpy array per buffer in a
    process-wide table, handles are the table keys, and gather_to_peers performs the copies the library would queue."""
    table = {}

    def __init__(self, rank):
        self.rank, self.log, self.n = rank, [], 0

    def peer_buffer_create(self, nbytes):
        key = (self.rank, self.n); self.n += 1
        _FakeCtx.table[key] = np.zeros(nbytes // 8)
        return key, repr(key).encode().ljust(64, b"\0")

    def peer_buffer_open(self, handle):
        key = eval(handle.rstrip(b"\0").decode())
        self.log.append(("open", key))
        return key

    def peer_buffer_close(self, ptr):
        self.log.append(("close", ptr))

    def peer_buffer_destroy(self, ptr):
        self.log.append(("destroy", ptr))

    def gather_to_peers(self, src, nbytes, peers, dst_offset, slot=0, mode=0):
        for p in peers:
            _FakeCtx.table[p][dst_offset // 8: dst_offset // 8 + nbytes // 8] = src
        self.log.append(("gather", slot, mode))

    def gather_wait(self, slot=0):
        self.log.append(("wait", slot))


def test_peer_gather_host_logic():
    """sharding.PeerGather without CUDA: every rank creates its buffers, maps every peer's in rank order, pushes its block
    into slot `rank` of every rank's buffer, and unmaps before it frees (the device side is test_peer_gather_single_process
    and the bit-for-bit check against NCCL inside bench.py)."""
    from robustgrape_b200.sharding import PeerGather
    world, blk = 3, 5
    _FakeCtx.table.clear()
    ctxs = [_FakeCtx(r) for r in range(world)]
    # what all_gather_object would return: creation order is deterministic, so handles can be listed up front
    def handles_of(r, nbuf=2):
        return [repr((r, b)).encode().ljust(64, b"\0") for b in range(nbuf)]
    every = [handles_of(r) for r in range(world)]
    pgs = [PeerGather(ctxs[r], r, world, blk, nbuf=2, mode=0, exchange=lambda h, r=r: (every.__setitem__(r, h) or every))
           for r in range(world)]
    for r in range(world):
        assert [k for op, k in ctxs[r].log if op == "open"] == [(q, b) for b in range(2) for q in range(world) if q != r]
        assert pgs[r].peers[1] == [(q, 1) for q in range(world)]
    for buf in (0, 1):
        for r in range(world):
            pgs[r].push(np.full(blk, 10.0 * buf + r), buf)
        for r in range(world):
            got = _FakeCtx.table[(r, buf)].reshape(world, blk)
            assert np.array_equal(got, np.repeat(10.0 * buf + np.arange(world), blk).reshape(world, blk))
    order = []
    pgs[0].close(barrier=lambda: order.append("barrier"))
    ops = [e[0] for e in ctxs[0].log if e[0] in ("close", "destroy")]
    assert ops == ["close"] * 4 + ["destroy"] * 2 and order == ["barrier"]
