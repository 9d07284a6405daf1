#!/usr/bin/env python
"""Per-kernel timing of one workload on one GPU (development aid; bench.py is the contract benchmark).
  python tools/quick_bench.py [--batch B] [--ntimes N] [--nerr E] [--model M] [--reps R]
Environment switches of the library (RG_WS, RG_CHUNK, ...) apply."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from robustgrape_b200._lib import Context, Problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--ntimes", type=int, default=1000)
ap.add_argument("--nerr", type=int, default=0)
ap.add_argument("--model", default="symmetric_blockaded")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tag", default="")
a = ap.parse_args()

ctx = Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
if a.model == "symmetric_blockaded":
    fp = bench.make_problem(a.ntimes, a.nerr)
else:
    from cases import cz_problem
    fp = cz_problem(a.ntimes, bench.T0, ("amp", "freq")[:a.nerr], a.model)
prob = Problem(fp, ctx)
nx = a.ntimes + 1
X = bench.make_pulses(a.ntimes, a.batch)
dX = torch.from_numpy(X).cuda()
out = torch.empty(a.batch * (nx + 1), dtype=torch.float64, device="cuda")
coeff = [1e-4] * a.nerr


def step():
    prob.cost_and_grad_batch_dev(a.batch, nx, dX.data_ptr(), coeff, out[:a.batch].data_ptr(), out[a.batch:].data_ptr())


for _ in range(3):
    step()
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
ctx.set_timing(True)
ctx.get_timing(reset=True)
for _ in range(a.reps):
    step()
t = ctx.get_timing(reset=True)
ctx.set_timing(False)
print(json.dumps({"tag": a.tag, "batch": a.batch, "ntimes": a.ntimes, "nerr": a.nerr, "model": a.model, "ms_per_step": ms,
                  "evals_per_s": a.batch / (ms * 1e-3), "kernel_ms": {k: v[0] / v[1] for k, v in t.items() if v[1]},
                  "cost0": float(out[0].item())}))
