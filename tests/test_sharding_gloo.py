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
