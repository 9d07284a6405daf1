#!/usr/bin/env python
"""End-to-end (host buffers through the C ABI) throughput of rg_cost_and_grad_batch on C4 for several slab settings of the
pipelined host entry point, next to the PCIe ceiling of the same traffic (pinned H2D and D2H copies on two streams).
  python tools/e2e_sweep.py [--batch 8192] [--reps 10]"""
import argparse, ctypes as C, json, os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from robustgrape_b200._lib import Context, Problem  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--ntimes", type=int, default=1000)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
N, B = a.ntimes, a.batch
nx = N + 1
hX = torch.from_numpy(bench.make_pulses(N, B)).pin_memory()
hcost = torch.empty(B, dtype=torch.float64).pin_memory()
hgrad = torch.empty(B * nx, dtype=torch.float64).pin_memory()
# PCIe ceiling: the same bytes, both directions at once
dX = torch.empty(B * nx, dtype=torch.float64, device="cuda"); dG = torch.empty(B * (nx + 1), dtype=torch.float64, device="cuda")
hG = torch.empty(B * (nx + 1), dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for it in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(a.reps):
        with torch.cuda.stream(s1): dX.copy_(hX.view(-1), non_blocking=True)
        with torch.cuda.stream(s2): hG.copy_(dG, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / a.reps
print(json.dumps({"pcie_both_directions_ms": dt * 1e3, "h2d_GBps": B * nx * 8 / dt / 1e9, "d2h_GBps": B * (nx + 1) * 8 / dt / 1e9,
                  "ceiling_evals_per_s": B / dt}))
cp = C.c_void_p
for slabs, smin in ((8, 1024), (4, 2048), (16, 512), (32, 256), (64, 128), (16, 256)):
    os.environ["RG_HOST_SLABS"], os.environ["RG_HOST_SLAB_MIN"] = str(slabs), str(smin)
    ctx = Context(0)
    prob = Problem(bench.make_problem(N, 0), ctx)
    h = prob.handle_for(nx)[0]
    def step():
        ctx.check(ctx.lib.rg_cost_and_grad_batch(h, B, cp(hX.data_ptr()), None, cp(hcost.data_ptr()), cp(hgrad.data_ptr())))
    for _ in range(3): step()
    t = time.perf_counter()
    for _ in range(a.reps): step()
    dt = (time.perf_counter() - t) / a.reps
    print(json.dumps({"host_slabs": slabs, "slab_min": smin, "ms_per_eval_batch": dt * 1e3, "e2e_evals_per_s": B / dt}))
    prob.close(); ctx.close()
