#!/usr/bin/env python
"""Smallest cases that reach every kernel family, for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import robustgrape_b200 as rg  # noqa: E402
from cases import cz_problem, dense_random_problem, random_pulse  # noqa: E402


def run(tag, env, fp, X):
    for k in ("RG_WS", "RG_B2", "RG_GROUP", "RG_DENSE_ALG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    print(tag, "F[0] =", out[0][0], flush=True)


N, B = 33, 3
X = 2 * np.pi * np.random.default_rng(0).random((N + 1, B))
run("fused quaternion (k_fused_q, diagonal algebra)", {}, cz_problem(N, 1.0, ("amp",)), X)
run("fused quaternion (dense algebra)", {"RG_DENSE_ALG": "1"}, cz_problem(N, 1.0, ("amp",)), X)
run("block-2 three-kernel (k_agg_b2, k_scan, k_grad_b2, k_grad_err_b2)", {"RG_B2": "1"}, cz_problem(N, 1.0, ("amp",)), X)
run("block-2 with diagonal terms", {}, cz_problem(N, 1.0, ("amp", "freq"), delta=0.2), X)
run("workspace path (k_steps_t, k_steps_so_t, k_chunk_agg_t, k_grad_t, k_grad_err_t)", {"RG_WS": "1"}, cz_problem(N, 1.0, ("amp",)), X)
run("group kernels (k_steps, k_steps_so, k_scan, k_grad)", {"RG_GROUP": "1"}, cz_problem(N, 1.0, ("amp",)), X)
fpd = dense_random_problem(12, 4, nparam=2, nerr=1, seed=1)
run("dense DMMA path (k_big_*)", {}, fpd, np.random.default_rng(1).uniform(-1, 1, (8, 2)))
fpa = cz_problem(70, 1.0, ("amp", "freq"))
xa = random_pulse(fpa, 1, 3)
print("response", rg.calculate_fidelity_response(fpa, xa, np.linspace(0, 2, 5))[0], flush=True)
print("expectation", rg.calculate_expectation_values(fpa, xa)[-1], flush=True)
print("sanitizer cases done")
