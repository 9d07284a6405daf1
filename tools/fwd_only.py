#!/usr/bin/env python
"""Time of the fused kernel without / with the backward sweep (C4): rg_fidelity_and_derivatives_batch_dev with and without gradient outputs."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench
from robustgrape_b200._lib import Context, Problem
B, N = 8192, 1000
nx = N + 1
ctx = Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
prob = Problem(bench.make_problem(N, 0), ctx)
h = prob.handle_for(nx)[0]
dX = torch.from_numpy(bench.make_pulses(N, B)).cuda()
F = torch.empty(B, dtype=torch.float64, device="cuda"); G = torch.empty(B * nx, dtype=torch.float64, device="cuda")
vp = C.c_void_p
lib = ctx.lib
lib.rg_fidelity_and_derivatives_batch_dev.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
for name, g in (("forward only (F)", None), ("forward + backward (F, F_dx)", G.data_ptr())):
    def step():
        ctx.check(lib.rg_fidelity_and_derivatives_batch_dev(h, B, vp(dX.data_ptr()), vp(F.data_ptr()), vp(g) if g else None, None, None))
    for _ in range(3): step()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 20, "ms")
