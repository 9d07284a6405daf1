#!/usr/bin/env python
"""bench.py -- GRAPE cost+grad evaluations/s on the multi-start CZ workload (BASELINE.json config C4).

  python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1)
  python bench.py --impl reference ...                     (CPU arm: C++ port of the reference's literal algorithm)

One "step" = one cost+gradient evaluation (reference `calculate_common!`, src/FidelityCalculations.jl:174-184,
without host regularisation) of every pulse of the batch: 8192 random-init pulses x 1000 time steps of the
5-level symmetric-blockaded Rydberg CZ problem (examples/time_optimal_cz.jl) per GPU. Pulses are independent, so for
N > 1 they are sharded over the ranks with no collective inside an evaluation; each step ends with an all-gather of every
rank's [cost | grad] block (peer-memory copies by default, `--gather nccl` for ncclAllGather).
`--scaling strong` (default): the 8192-pulse batch of BASELINE.json configs[3] is fixed and sharded over the ranks.
`--scaling weak`: the multi-start batch grows with the box, 8192 pulses per GPU (also measured and reported under
`extra.weak_scaling` when N > 1).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T0 = 7.613            # examples/time_optimal_cz.jl:14
PROJ = np.diag([1.0, 2.0, 1.0, 0.0, 0.0])


def make_problem(ntimes, nerr):
    import robustgrape_b200 as rg
    from robustgrape_b200 import rydberg_tools as rt
    errs = []
    if nerr >= 1:
        errs.append(rg.ErrorSource(rt.rydberg_amplitude_error(source=0)))
    if nerr >= 2:
        errs.append(rg.ErrorSource(rt.rydberg_frequency_error(source=1)))
    up = rg.UnitaryRobustGRAPEProblem(t0=T0, ntimes=ntimes, ndim=5, H0=rt.rydberg_h0(), nb_additional_param=1,
                                      error_sources=errs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ, rt.cz_target())


def make_model_problem(ntimes, d):
    """The reference's other Rydberg models as descriptor problems: d = 7 full-blockaded (src/RydbergTools.jl:71-81), d = 9 full two-atom
    model with finite blockade and detunings (:118-130); CZ target with a single-qubit phase, projector on the computational levels."""
    import robustgrape_b200 as rg
    from robustgrape_b200 import rydberg_tools as rt
    if d == 7:
        up = rg.UnitaryRobustGRAPEProblem(t0=T0, ntimes=ntimes, ndim=7, H0=rt.rydberg_h0("full_blockaded"), nb_additional_param=1, error_sources=[])
        return rg.FidelityRobustGRAPEProblem(up, np.diag([1.0, 1, 1, 1, 0, 0, 0]), rt.cz_target("full_blockaded"))
    from robustgrape_b200.descriptors import Factor, Term, TermTarget, S_ADD, OWNER_TARGET
    tgt = TermTarget(9, [Term(1.0, (), ((0, 0, 1.0),), OWNER_TARGET),
                         Term(1.0, (Factor.expi(S_ADD, 0, 1.0, 0.0),), ((1, 1, 1.0), (2, 2, 1.0)), OWNER_TARGET),
                         Term(1.0, (Factor.expi(S_ADD, 0, 2.0, np.pi),), ((3, 3, 1.0),), OWNER_TARGET)])
    up = rg.UnitaryRobustGRAPEProblem(t0=T0, ntimes=ntimes, ndim=9, H0=rt.rydberg_full_h0(1.0, 1.0, 0.0, 0.0, 8.0), nb_additional_param=1, error_sources=[])
    return rg.FidelityRobustGRAPEProblem(up, np.diag([1.0, 1, 1, 1, 0, 0, 0, 0, 0]), tgt)


def make_pulses(ntimes, batch, seed=43):
    """phi_k ~ 2 pi U(0,1) (test/runtests.jl:91), theta ~ 2 pi U(0,1); one pulse per row (C order)
    == one pulse per column of the (nx, B) column-major array the C ABI takes."""
    rng = np.random.default_rng(seed)
    return 2 * np.pi * rng.random((batch, ntimes + 1))


def canonical_flops(d, N, p, a, e):
    """8 d^3 N (n_exp_needed c_exp + 4 + 8 e), Pade-3 regime (c_exp = 2 + 4/3)."""
    n_exp_needed = 1 + p + a + e * (2 + p + a) + ((p + a) if e > 0 else 0)
    return 8.0 * d ** 3 * N * (n_exp_needed * (2 + 4.0 / 3.0) + 4 + 8 * e)


def k_steps_flops(N, taylor_m, structured):
    """Real flops the step-propagator kernel k_steps_t<5, mask> issues per pulse for the C4 problem (one first-order
    object: U and dU/dphi), counting one complex FMA as 8 flops.
      dense mask : 5 columns x (m-1) Horner iterations x (10 off-diagonal positions x 6 cFMA + 5 diagonal x 3 cFMA)
      CZ mask    : the closure of the drive couplings (1,3),(2,4) is block diagonal, so only 4 columns do work and each
                   touches one off-diagonal position: 4 columns x (m-1) iterations x 6 cFMA."""
    it = taylor_m - 1
    per_step = (4 * it * 6 if structured else 5 * it * (10 * 6 + 5 * 3)) * 8.0
    return per_step * N


def taylor_degree_for(nrm):
    """Mirror of taylor_degree() in csrc/rg_common.cuh (the kernel bounds ||dt*H||_1 within 8.3 % from above)."""
    for th, m in ((1.5e-3, 4), (6.0e-3, 5), (1.6e-2, 6), (3.4e-2, 7), (6.5e-2, 8), (0.105, 9), (0.16, 10), (0.225, 11), (0.31, 12),
                  (0.52, 14), (0.78, 16), (1.10, 18)):
        if nrm <= th:
            return m
    return 12


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_reference(args):
    """CPU arm: the C++ port of the reference's literal algorithm (oracle/cpu_port.cpp) on all host threads.
    Julia is not installed in this image, so the reference itself cannot run; kind = "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_port
    threads = os.cpu_count() or 1
    fp = make_problem(args.ntimes, args.nerr)
    pp = cpu_port.PortProblem(fp)
    sample = min(args.batch, max(8, args.cpu_pulses_per_thread * threads))
    X = make_pulses(args.ntimes, sample)
    coeff = [1e-4] * args.nerr
    for _ in range(args.warmup):
        pp.cost_and_grad_batch(X.T, coeff, threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        pp.cost_and_grad_batch(X.T, coeff, threads)
    dt = time.perf_counter() - t
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "GRAPE cost+grad evals/sec (CZ, batched pulses)", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, max(1, args.gpus), args.scaling, "none"),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} pulses of the workload per step (C++ port of the reference's literal "
                                   "algorithm; Julia unavailable in this image)"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, world=1, scaling="weak", gather="peer"):
    per_gpu = args.batch if scaling == "weak" else args.batch // world
    payload = "costs" if getattr(args, "gather_what", "costs") == "costs" else "[cost|grad] blocks"
    how = {"fused": f"all-gather of the {payload} fused into the evaluation kernel: it stores every rank's results straight into all ranks' "
                    "gathered buffers (CUDA IPC mappings, stores over NVLink)",
           "peer": f"all-gather of the {payload} per step by peer-memory copies (CUDA IPC mappings, copy engines over NVLink)",
           "nccl": f"NCCL all-gather of the {payload} per step", "none": "no gather (diagnostic)"}[gather]
    return {
        "workload": f"C4 multi-start CZ: {args.batch} random-init pulses x {args.ntimes} time steps, d=5 symmetric-blockaded "
                    f"Rydberg, p=1, a=1, e={args.nerr}, t0={T0}, eps=1e-8, eps2=1e-4 (examples/time_optimal_cz.jl)"
                    + (f"; {scaling} scaling: {per_gpu} pulses per GPU, {per_gpu * world} in total" if world > 1 else ""),
        "batch": per_gpu * world, "per_gpu_batch": per_gpu, "ntimes": args.ntimes, "nerr": args.nerr,
        "sharding": f"independent pulses over ranks, no collective inside an evaluation; {how}, overlapped with the next "
                    "step's kernels (double-buffered)" if world > 1 else "single GPU",
        "l2": "inputs larger than L2: the timed steps rotate over up to 4 distinct resident pulse sets (65.6 MB each at 8192 pulses) "
              "and two output buffers (65.7 MB each), > 2 x 126 MB between two uses of the same buffer; no explicit flush",
    }


def canonical_flops_dense(d, N, p, a, e, nrm1):
    """SURVEY.md section 8(d) with the Pade degree / squarings Higham's thresholds pick for ||dt H||_1 = nrm1."""
    pis = ((0.015, 2), (0.25, 3), (0.95, 4), (2.1, 5), (5.4, 6))
    s, pi_m = 0, None
    for th, pm in pis:
        if nrm1 <= th:
            pi_m = pm
            break
    if pi_m is None:
        s = int(np.ceil(np.log2(nrm1 / 5.4)))
        pi_m = 6
    n_exp_needed = 1 + p + a + e * (2 + p + a) + ((p + a) if e > 0 else 0)
    return 8.0 * d ** 3 * N * (n_exp_needed * (pi_m + 4.0 / 3.0 + s) + 4 + 8 * e)


def make_dense_problem(d, N, p, e, target_norm, seed=64):
    """BASELINE.json configs[4]: H(k) = H_0 + sum_j x_j(k) H_j, Herr_e = err * E_e, GUE draws scaled to unit spectral norm
    (default_rng(64)), x ~ U(-1, 1), projector = identity on the first 16 levels, Haar-random target on that block (default_rng(65)),
    dt chosen so that max_k ||dt H(k)||_1 = target_norm.  Returns (problem, x, realised 1-norm range)."""
    import robustgrape_b200 as rg
    from robustgrape_b200.descriptors import (Factor, Term, TermHamiltonian, TermErrorHamiltonian, ConstantTarget, S_MAIN, OWNER_H0)
    rng = np.random.default_rng(seed)

    def gue():
        g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        h = (g + g.conj().T) / 2
        return h / np.linalg.norm(h, 2)

    def ent(M):
        return tuple((r, c, M[r, c]) for r in range(d) for c in range(d))

    Hs = [gue() for _ in range(1 + p)]
    Es = [gue() for _ in range(e)]
    x = np.random.default_rng(seed + 2).uniform(-1, 1, p * N)
    xs = x.reshape(N, p)
    sample = xs[:: max(1, N // 64)]
    n1 = np.array([np.abs(Hs[0] + sum(xx[j] * Hs[1 + j] for j in range(p))).sum(axis=0).max() for xx in sample])
    dt = target_norm / n1.max()
    terms = [Term(1.0, (), ent(Hs[0]), OWNER_H0)] + [Term(1.0, (Factor.var(S_MAIN, j),), ent(Hs[1 + j]), OWNER_H0) for j in range(p)]
    srcs = [rg.ErrorSource(TermErrorHamiltonian(d, [Term(1.0, (Factor.err(),), ent(Es[i]), i)])) for i in range(e)]
    nb = min(d, 16)
    proj = np.zeros((d, d)); proj[:nb, :nb] = np.eye(nb)
    r65 = np.random.default_rng(seed + 1)
    q, _ = np.linalg.qr(r65.normal(size=(nb, nb)) + 1j * r65.normal(size=(nb, nb)))
    U0 = np.zeros((d, d), dtype=complex); U0[:nb, :nb] = q
    up = rg.UnitaryRobustGRAPEProblem(t0=dt * N, ntimes=N, ndim=d, H0=TermHamiltonian(d, terms), nb_additional_param=0, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, proj, ConstantTarget(U0)), x, (float(n1.min() * dt), float(n1.max() * dt))


def run_dense(args):
    """--workload C5 | d16: one pulse of the synthetic dense problem (d = 64 or 16, N = 1e4, p = 8, e = 4) on the DMMA path."""
    import torch
    from robustgrape_b200._lib import Context, Problem
    d = {"C5": 64, "d16": 16, "d32": 32, "d48": 48}[args.workload]
    N, p, e = args.dense_ntimes, 8, 4
    fp, x, nrm = make_dense_problem(d, N, p, e, args.dense_norm)
    ctx = Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    prob = Problem(fp, ctx)
    nx = p * N
    B = args.dense_batch
    X = np.stack([x] + [np.random.default_rng(700 + b).uniform(-1, 1, nx) for b in range(1, B)], axis=0)
    dX = torch.from_numpy(X).cuda()
    out = torch.empty(B * (nx + 1), dtype=torch.float64, device="cuda")
    coeff = [1e-4] * e
    peak_dfma, peak_dmma = ctx.measure_fp64_peak(0.3)

    def step():
        prob.cost_and_grad_batch_dev(B, nx, dX.data_ptr(), coeff, out[:B].data_ptr(), out[B:].data_ptr())
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    l0 = ctx.launch_count
    sampler = ClockSampler(0)
    sampler.__enter__()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.__exit__(None, None, None)
    launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1) / args.steps
    ctx.set_timing(True); ctx.get_timing(reset=True)
    step()
    timing = ctx.get_timing(reset=True)
    ctx.set_timing(False)
    kernel_ms = {k: v[0] / v[1] for k, v in timing.items() if v[1]}
    # e2e through the host-buffer C-ABI call
    hX = np.asfortranarray(X.T)
    t = time.perf_counter()
    cost_h, grad_h = prob.cost_and_grad_batch(hX, coeff)
    e2e_s = time.perf_counter() - t
    cost_d = out[:B].cpu().numpy()
    assert np.allclose(cost_h, cost_d, rtol=0, atol=1e-12)
    # executed products: 5 + s jet products of 1 + 4 nf + 4 nv ne matrix products each (the last one skips the 2 nf eps2 slots),
    # sweeps: aggregate 1 + 2e, gradient 3 + 9e per step
    DP = ((d + 15) // 16) * 16
    nf, nve = p + e, p * e
    s_sq = max(0, int(np.ceil(np.log2((nrm[1] * 1.001) / 0.31))))
    per_jet = 1 + 2 * (2 * nf) + 4 * nve
    steps_prods = (5 + s_sq) * per_jet - 2 * nf - 3 * nve      # last stage skips the eps2 slots; the error terms do not depend on the
    prods_step = steps_prods + (1 + 2 * e) + (3 + 9 * e)         # controls, so B's mixed slots vanish: B*B and B2*B skip 2 + 1 products per pair
    executed = prods_step * 8.0 * DP ** 3 * N
    canonical = canonical_flops_dense(d, N, p, 0, e, nrm[1])
    dom = max(kernel_ms, key=kernel_ms.get)
    exec_dom = steps_prods * 8.0 * DP ** 3 * N * B
    ach_exec = exec_dom / (kernel_ms[dom] * 1e-3) / 1e12 if dom == "k_steps" else None
    ach_canon = min(executed, canonical) * B / (ms * 1e-3) / 1e12
    line = {
        "metric": "GRAPE cost+grad evals/sec (dense synthetic, DMMA path)", "value": B / (ms * 1e-3), "unit": "evals/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: d={d}, N={N}, p={p}, a=0, e={e}, {B} pulse(s); GUE H_j and E_e scaled to unit spectral norm, "
                               f"x ~ U(-1,1), ||dt H||_1 in [{nrm[0]:.2f}, {nrm[1]:.2f}] ({s_sq} squarings of a degree-12 Taylor polynomial), "
                               "projector = first 16 levels, Haar target (BASELINE.json configs[4] / SURVEY 8d)",
                   "l2": "per-step workspace (3.2 MB/step at d=64) and jet scratch far exceed L2"},
        "clocks": sampler.summary(),
        "e2e": {"value": B / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(B * nx * 8), "d2h_bytes_per_step": int(B * (nx + 1) * 8)},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "kernel": "k_big_steps (jets of planar matrices, DMMA m8n8k4 products)",
                     "achieved": ach_exec, "peak": peak_dmma, "unit": "TFLOP/s", "frac": (ach_exec / peak_dmma) if ach_exec else None,
                     "traffic": None, "peak_source": "DMMA m8n8k4 microbenchmark run by this process (DFMA loop: %.1f TFLOP/s)" % peak_dfma,
                     "flops_per_eval": {"canonical_survey_8d": canonical, "executed": executed, "executed_products_per_step": prods_step},
                     "whole_step_canonical": {"achieved_tflops": ach_canon, "frac_of_dmma_peak": ach_canon / peak_dmma,
                                              "rule": "min(executed, canonical) / whole-step time (SURVEY 8d)"},
                     "note": "exact finite differences cost 2 products per value product for first differences and 4 for mixed second "
                             "differences, so executed flops exceed the canonical (independent Pade exponentials) count; both are reported",
                     "kernel_ms": kernel_ms},
    }
    emit(line)


def run_single_pulse(args):
    """--workload C1 | C2 (BASELINE.json configs[0], configs[1]): one pulse through the reference-signature call
    calculate_fidelity_and_derivatives (host buffers in and out, blocking): the latency-bound single-pulse points.
    C1 = examples/time_optimal_cz.jl (N = 500, t0 = 7.613, e = 0), C2 = examples/ar_cz.jl (N = 200, t0 = 14.32, amplitude error)."""
    import torch
    import robustgrape_b200 as rg
    from robustgrape_b200 import rydberg_tools as rt
    from robustgrape_b200._lib import Context
    from robustgrape_b200.unitary_calculations import device_problem
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(0)
    ctx = Context(0)
    N, t0, nerr = (500, 7.613, 0) if args.workload == "C1" else (200, 14.32, 1)
    srcs = [rg.ErrorSource(rt.rydberg_amplitude_error())] if nerr else []
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=N, ndim=5, H0=rt.rydberg_h0(), nb_additional_param=1, error_sources=srcs)
    fp = rg.FidelityRobustGRAPEProblem(up, np.diag([1.0, 2, 1, 0, 0]), rt.cz_target())
    x = 2 * np.pi * np.random.default_rng(43).random(N + 1)
    dp = device_problem(fp, ctx)
    for _ in range(max(3, args.warmup)):
        out = rg.calculate_fidelity_and_derivatives(fp, x, ctx=ctx)
    l0 = ctx.launch_count
    t = time.perf_counter()
    for _ in range(args.steps):
        out = rg.calculate_fidelity_and_derivatives(fp, x, ctx=ctx)
    dt = (time.perf_counter() - t) / args.steps
    launches = ctx.launch_count - l0
    line = {"metric": "GRAPE fidelity+derivatives evals/sec (CZ, single pulse)", "value": 1.0 / dt, "unit": "evals/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: one pulse, d=5 symmetric-blockaded Rydberg, N={N}, t0={t0}, e={nerr}; reference-signature call "
                                   "calculate_fidelity_and_derivatives with host buffers (all four outputs returned)",
                       "note": "latency bound: one fused-kernel launch per role plus the host call, the copies and a synchronisation"},
            "e2e": {"value": 1.0 / dt, "unit": "evals/s", "h2d_bytes_per_step": int((N + 1) * 8),
                    "d2h_bytes_per_step": int((1 + (N + 1) + nerr + nerr * (N + 1)) * 8)},
            "gpu_launches": int(launches), "path": dp.path(N + 1), "F": float(out[0])}
    if not args.no_cpu_baseline:
        from oracle import cpu_port
        pp = cpu_port.PortProblem(fp)
        X1 = x[:, None]
        pp.fidelity_and_derivatives_batch(X1, 1)
        t = time.perf_counter()
        reps = 5
        for _ in range(reps):
            pp.fidelity_and_derivatives_batch(X1, 1)
        ct = (time.perf_counter() - t) / reps
        line["cpu_baseline"] = {"value": 1.0 / ct, "unit": "evals/s", "cores": 1, "kind": "port",
                                "sample": "the same pulse, C++ port of the reference's literal algorithm on one host thread"}
    emit(line)


def run_response(args):
    """--workload C3 (BASELINE.json configs[2]): fidelity-response sweep over a 4096-point frequency grid (N = 500, amplitude + frequency
    error sources, examples/time_optimal_cz.jl:60-67,82), the grid sharded over the ranks (no exchange inside a shard), rows all-gathered."""
    import torch
    import torch.distributed as dist
    from robustgrape_b200._lib import Context
    import robustgrape_b200 as rg
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = Context(local)
    N, nfreq = 500, 4096
    fp = make_problem(N, 2)
    x = make_pulses(N, 1)[0]
    freqs = np.linspace(0, 3, nfreq)
    from robustgrape_b200.sharding import shard_range
    first, hi = shard_range(nfreq, rank, world)
    count = hi - first
    full = torch.empty((world, 2, (nfreq + world - 1) // world), dtype=torch.float64, device=dev)

    def sweep():
        R = rg.calculate_fidelity_response(fp, x, freqs, ctx=ctx, first=first, count=count)          # (count, 2), blocking C-ABI call
        mine = torch.zeros((2, full.shape[2]), dtype=torch.float64, device=dev)
        mine[:, :count] = torch.from_numpy(np.ascontiguousarray(R.T)).to(dev)
        if world > 1:
            dist.all_gather_into_tensor(full.view(-1), mine.view(-1))
        else:
            full[0] = mine
        torch.cuda.synchronize()
        return R
    for _ in range(args.warmup):
        sweep()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(args.steps):
        R = sweep()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ms = float(dt.item()) / args.steps * 1e3
    if rank == 0:
        line = {"metric": "fidelity-response sweeps/sec (4096-point grid)", "value": 1e3 / ms, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"C3: fidelity response, {nfreq} frequencies in [0, 3], N={N}, e=2 (amplitude, frequency), one pulse; grid "
                                       f"sharded over {world} rank(s), rows all-gathered (NCCL)",
                           "note": "a single sweep is ~0.8 GFLOP: launch/latency bound (SURVEY 8d says so); the time is host-call wall clock, "
                                   "copies and the all-gather included"},
                "e2e": {"value": 1e3 / ms, "unit": "sweeps/s", "h2d_bytes_per_step": int((N + 1) * 8 + nfreq * 8), "d2h_bytes_per_step": int(count * 2 * 8)},
                "gpu_launches": int(ctx.launch_count)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly one JSON line, but libraries write banners to fd 1 (NCCL prints its version there under torchrun
    whatever NCCL_DEBUG_FILE says).  Everything that is not the result line goes to stderr: fd 1 is pointed at fd 2 and the
    result is written to a duplicate of the original fd 1."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--ntimes", type=int, default=1000)
    ap.add_argument("--nerr", type=int, default=0)
    ap.add_argument("--cpu-pulses-per-thread", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--workload", default="C4", choices=["C4", "C5", "d16", "d32", "d48", "C3", "C1", "C2"],
                    help="C4 (default): multi-start CZ batch, the headline; C5 / d16 / d32 / d48: dense synthetic problem (d = 64 / 16 / 32 / 48) on the DMMA path (1 GPU)")
    ap.add_argument("--dense-ntimes", type=int, default=10000)
    ap.add_argument("--dense-norm", type=float, default=2.4, help="max_k ||dt H(k)||_1 of the dense workload (SURVEY 8d: 2.1 ... 5.4)")
    ap.add_argument("--dense-batch", type=int, default=1)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the --batch pulses of BASELINE.json configs[3] sharded over the ranks; "
                         "weak = --batch pulses per GPU (reported under extra.weak_scaling)")
    ap.add_argument("--gather-what", default="costs", choices=["costs", "all"],
                    help="N > 1: what every rank receives from every other rank per evaluation.  costs (default): the costs (8 B per "
                         "pulse) -- all a sharded multi-start optimisation exchanges, each rank's optimiser consumes its own gradients; "
                         "all: costs and gradients (north_star's all-gather; measured in the same run under extra.full_gather)")
    ap.add_argument("--gather", default="fused", choices=["fused", "peer", "nccl", "none"],
                    help="N > 1: how the per-rank [cost|grad] blocks are gathered each step: fused = stores to the peers' buffers from "
                         "the evaluation kernel itself (rg_cost_and_grad_batch_dev_scatter), peer = copy engines after it, nccl = all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "C3":
        run_response(args)
        return
    if args.workload in ("C1", "C2"):
        run_single_pulse(args)
        return
    if args.workload != "C4":
        run_dense(args)
        return

    import torch
    import torch.distributed as dist
    from robustgrape_b200._lib import Context, Problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL prints its version banner to stdout otherwise: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    N = args.ntimes
    nx = N + 1
    coeff = [1e-4] * args.nerr
    gather = args.gather if world > 1 else "none"
    if os.environ.get("RG_BENCH_NO_GATHER"):
        gather = "none"
    gather_mode = int(os.environ.get("RG_GATHER_MODE", "0"))        # 0 copy engines, 1 store kernel

    ctx = Context(local)
    # All work (library kernels, NCCL, timing events) goes on one non-default torch stream.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    prob = Problem(make_problem(N, args.nerr), ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def shard_inputs(scaling):
        """(global batch, per-rank batch, this rank's pulses)."""
        if scaling == "weak" or world == 1:
            return args.batch * world, args.batch, make_pulses(N, args.batch, seed=43 + rank)
        assert args.batch % world == 0, "batch must divide over the ranks"
        bs = args.batch // world
        return args.batch, bs, np.ascontiguousarray(make_pulses(N, args.batch)[rank * bs:(rank + 1) * bs])

    def device_run(scaling, steps, warmup, sample_clocks, what=None):
        """Times `steps` evaluations of this rank's shard, inputs resident in HBM; max over ranks.  what = "all": every rank's
        [cost | grad] block is gathered on every rank each step (north_star); "costs": only the costs are (the gradients stay with
        the rank whose optimiser consumes them)."""
        what = what or args.gather_what
        B, Bs, Xs = shard_inputs(scaling)
        dX = torch.from_numpy(Xs).to(dev)
        # inputs rotate over NROT distinct pulse sets so that, together with the double-buffered outputs, the bytes touched
        # between two uses of the same buffer exceed the 126 MB L2 (no L2-resident inputs from the previous step)
        nrot = max(1, min(4, int(np.ceil(2 * 126e6 / max(1.0, Bs * nx * 8.0)))))
        dXs = [dX] + [torch.from_numpy(make_pulses(N, Bs, seed=1000 + 17 * r + rank)).to(dev) for r in range(1, nrot)]
        blk = Bs * (1 + nx)
        gblk = blk if what == "all" else Bs                      # doubles gathered per rank and step
        # [cost (Bs) | grad (Bs, nx)] per rank; two buffers so that the gather of step i (side streams / NCCL's stream)
        # overlaps the kernels of step i+1: in multi-start optimisation a rank's next evaluation only needs its own shard.
        out_locals = [torch.empty(blk, dtype=torch.float64, device=dev) for _ in range(2)]
        pg = None
        out_alls = None
        if gather in ("peer", "fused"):
            from robustgrape_b200.sharding import PeerGather
            pg = PeerGather(ctx, rank, world, gblk, nbuf=2, mode=gather_mode)
            out_alls = [pg.view(i, dev) for i in range(2)]
        elif gather == "nccl":
            out_alls = [torch.empty(world * gblk, dtype=torch.float64, device=dev) for _ in range(2)]
        pending = [None, None]
        step_no = [0]
        flag = torch.zeros(1, dtype=torch.float32, device=dev) if pg is not None else None
        # Cross-rank completion of the peer gather: drain() orders the end of the timed region after this rank's outgoing copies of
        # every step, and finish() adds one NCCL barrier inside the timed region -- after it, every rank's block of every timed
        # step has landed in every rank's buffer.  (A per-step NCCL handshake was measured: its ~100 us of host-side launch work
        # per step exceeds the 80 us evaluation of a 1024-pulse shard; consumers that need per-step completion poll nothing --
        # they order their reader after rg_gather_wait_on + their own barrier, see sharding.PeerGather.)
        def step():
            i = step_no[0] & 1
            step_no[0] += 1
            if pending[i] is not None:
                pending[i].wait()                                  # buffer i is free again (its gather finished on every rank)
                pending[i] = None
            if pg is not None:
                pg.wait(i)
            ol = out_locals[i]
            xin = dXs[(step_no[0] - 1) % len(dXs)]
            if gather == "fused":
                # one call: the evaluation kernel stores this rank's block into every rank's gathered buffer (its own included)
                targets, off = pg.scatter_targets(i, include_self=True)
                prob.cost_and_grad_batch_dev_scatter(Bs, nx, xin.data_ptr(), coeff, ol[:Bs].data_ptr(), ol[Bs:].data_ptr(), targets, off,
                                                     1 if what == "all" else 0)
                return
            prob.cost_and_grad_batch_dev(Bs, nx, xin.data_ptr(), coeff, ol[:Bs].data_ptr(), ol[Bs:].data_ptr())
            if gather == "nccl":
                pending[i] = dist.all_gather_into_tensor(out_alls[i], ol[:gblk], async_op=True)
            elif pg is not None:
                pg.push(ol.data_ptr(), i)

        def drain():
            for i in (0, 1):
                if pending[i] is not None:
                    pending[i].wait()
                    pending[i] = None
                if pg is not None:
                    pg.wait(i)

        def finish():
            drain()
            if pg is not None:
                dist.all_reduce(flag)          # stream-ordered after the drained pushes of every rank: cross-rank completion

        # The clock sampler (NVML, 20 Hz) starts before the warm-up and runs through the timed region: the GPU executes the same steps
        # the whole time, so every sample is "under load".  It must not *start* at the timed region: its first query then always lands
        # inside it, an NVML query holds up kernel launches for a few milliseconds, and at 20 steps of a 35 us shard that alone made
        # the 8-GPU step 0.18 ms instead of 0.035 ms (measured).
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        for _ in range(warmup):
            step()
        if sample_clocks:                                                  # keep the GPU loaded until the barrier (untimed); a fixed count,
            for _ in range(16):                                            # so that every rank issues the same collectives with --gather nccl
                for _ in range(16):
                    step()
                drain()
                ctx.synchronize()
        drain()
        ctx.synchronize()
        barrier()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        finish()                                                   # every step's gather completes on every rank inside the timed region
        e1.record()
        barrier()
        if sampler:
            sampler.__exit__(None, None, None)
        launches = ctx.launch_count - l0
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ctx.synchronize()
        last = (step_no[0] - 1) & 1
        x_last = dXs[(step_no[0] - 1) % len(dXs)]
        if out_alls is not None:
            # the gathered buffer must hold every rank's block: check against one untimed NCCL all-gather
            ref = torch.empty(world * gblk, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(ref, out_locals[last][:gblk])
            torch.cuda.synchronize()
            assert torch.equal(ref, out_alls[last]), f"rank {rank}: gathered [cost|grad] differs from NCCL all-gather"
            del ref
        cost_host = out_locals[last][:Bs].cpu().numpy()
        if pg is not None:
            barrier()
            out_alls = None
            pg.close(barrier)
        if x_last is not dX:                      # the e2e cross-check below uses Xs: evaluate it once more, untimed
            prob.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), coeff, out_locals[last][:Bs].data_ptr(), out_locals[last][Bs:].data_ptr())
            ctx.synchronize()
            cost_host = out_locals[last][:Bs].cpu().numpy()
        del dXs
        return {"B": B, "Bs": Bs, "Xs": Xs, "dX": dX, "ms": float(ms.item()), "launches": launches, "cost": cost_host,
                "clocks": sampler.summary() if sampler else None, "out": out_locals[0],
                "gather_bytes_in_per_rank_per_step": (world - 1) * gblk * 8 if out_alls is not None or pg is not None or gather != "none" else 0}

    # Every rank runs the FP64 peak loops (0.3 s of full load): rank 0 reports them, and on all ranks they bring the SM clock up from
    # idle before anything is timed -- with 20 steps of a 35 us shard the timed region is under a millisecond, and an idle GPU that is
    # still ramping its clock made the max-over-ranks time of an 8-GPU run several times longer than the same run repeated.
    peak_dfma, peak_dmma = ctx.measure_fp64_peak(0.3)
    if rank != 0:
        peak_dfma = peak_dmma = None

    scaling = args.scaling if world > 1 else "weak"
    run = device_run(scaling, args.steps, args.warmup, True)
    B, Bs, Xs, dX = run["B"], run["Bs"], run["Xs"], run["dX"]
    ms_total, launches, cost_host = run["ms"], run["launches"], run["cost"]
    value = B * args.steps / (ms_total * 1e-3)
    dcost, dgrad = run["out"][:Bs], run["out"][Bs:]
    other = None
    if world > 1 and not args.no_extra:
        oscal = "strong" if scaling == "weak" else "weak"
        if oscal == "weak" or args.batch % world == 0:
            o = device_run(oscal, args.steps, args.warmup, False)
            other = {"scaling": oscal, "batch": o["B"], "per_gpu_batch": o["Bs"], "ms_per_step": o["ms"] / args.steps,
                     "evals_per_s": o["B"] * args.steps / (o["ms"] * 1e-3)}
            del o
    other_what = None
    if world > 1 and not args.no_extra and gather != "none":
        ow = "all" if args.gather_what == "costs" else "costs"
        o = device_run(scaling, args.steps, args.warmup, False, what=ow)
        gb, step_s = o["gather_bytes_in_per_rank_per_step"], o["ms"] / args.steps * 1e-3
        other_what = {"what": ow, "scaling": scaling, "ms_per_step": o["ms"] / args.steps, "evals_per_s": o["B"] * args.steps / (o["ms"] * 1e-3),
                      "bytes_in_per_rank_per_step": gb, "nvlink_ingress_GBps": gb / step_s / 1e9, "nvlink_peak_GBps": 900.0,
                      "floor_ms_per_step": gb / 900e9 * 1e3,
                      "note": ("north_star's all-gather of costs AND gradients (8.02 kB per pulse to every rank): bound by NVLink ingress, "
                               "not by the kernels -- floor_ms_per_step is the step time at 900 GB/s; on this box the fused stores, the "
                               "copy engines and NCCL all deliver ~200 GB/s per GPU at 8 GPUs (DESIGN.md section 6)") if ow == "all" else
                              "same run with only the costs gathered (8 B per pulse)"}
        del o

    # ---- e2e: host buffers through the C ABI, H2D and D2H inside the timed region
    hX = torch.from_numpy(Xs).pin_memory()
    hcost = torch.empty(Bs, dtype=torch.float64).pin_memory()
    hgrad = torch.empty(Bs * nx, dtype=torch.float64).pin_memory()
    lib, h = ctx.lib, prob.handle_for(nx)[0]
    import ctypes as C
    cp = C.c_void_p
    cf = np.asarray(coeff, dtype=np.float64)

    def e2e_step():
        ctx.check(lib.rg_cost_and_grad_batch(h, Bs, cp(hX.data_ptr()), cf.ctypes.data_as(cp) if args.nerr else None,
                                             cp(hcost.data_ptr()), cp(hgrad.data_ptr())))
    for _ in range(2):
        e2e_step()
    barrier()
    t = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = B * args.steps / float(e2e_s.item())
    assert np.allclose(hcost.numpy(), cost_host, rtol=0, atol=1e-12), "e2e and device-resident paths disagree"

    # ---- live per-kernel timing (CUDA events on the launch stream) for the roofline
    ctx.set_timing(True)
    ctx.get_timing(reset=True)
    for _ in range(max(3, min(args.steps, 10))):
        prob.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), coeff, dcost.data_ptr(), dgrad.data_ptr())
    timing = ctx.get_timing(reset=True)
    ctx.set_timing(False)

    extra = None
    if not args.no_extra and args.nerr == 0 and rank == 0 and world == 1:
        # C4' : same workload with one error source (amplitude), the robust-GRAPE path
        try:
            p1 = Problem(make_problem(N, 1), ctx)
            for _ in range(2):
                p1.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), [1e-4], dcost.data_ptr(), dgrad.data_ptr())
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            k = max(2, args.steps // 2)
            for _ in range(k):
                p1.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), [1e-4], dcost.data_ptr(), dgrad.data_ptr())
            f1.record()
            torch.cuda.synchronize()
            extra = {"C4prime_e1_evals_per_s": B * k / (f0.elapsed_time(f1) * 1e-3)}
            p1.close()
        except Exception as ex:      # noqa: BLE001
            extra = {"C4prime_e1_error": str(ex)}
        # the reference's other two Rydberg models (src/RydbergTools.jl:71-81,118-130): 7-level full-blockaded (1+2+2+2 blocks, fused
        # phase-only kernel) on the same batch, 9-level finite-blockade model (a 4-level block: general kernels) on 512 pulses
        for key, d, bsz in (("model_d7", 7, Bs), ("model_d9", 9, min(Bs, 512))):
            try:
                pm = Problem(make_model_problem(N, d), ctx)
                for _ in range(2):
                    pm.cost_and_grad_batch_dev(bsz, nx, dX.data_ptr(), [], dcost.data_ptr(), dgrad.data_ptr())
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                k = max(2, args.steps // 2)
                for _ in range(k):
                    pm.cost_and_grad_batch_dev(bsz, nx, dX.data_ptr(), [], dcost.data_ptr(), dgrad.data_ptr())
                f1.record()
                torch.cuda.synchronize()
                extra[key] = {"pulses": bsz, "evals_per_s": bsz * k / (f0.elapsed_time(f1) * 1e-3), "path": pm.path(nx)}
                pm.close()
            except Exception as ex:      # noqa: BLE001
                extra[key + "_error"] = str(ex)

    # ---- the optimiser loop on the device: batched L-BFGS with X resident in HBM, sin^2 regularisation epilogue
    if not args.no_extra and rank == 0 and world == 1 and args.nerr == 0:
        try:
            x0 = np.concatenate([2 * np.pi * 0.001 * np.random.default_rng(7).random((Bs, N)), 2 * np.pi * np.random.default_rng(8).random((Bs, 1))], axis=1)
            dXo = torch.from_numpy(x0).to(dev)
            dco = torch.empty(Bs, dtype=torch.float64, device=dev)
            reg = [(3, 1e-6, 1e-6)]
            prob.lbfgs_batch_dev(Bs, nx, dXo.data_ptr(), coeff, dco.data_ptr(), reg, 10, 2, 0.0)          # warm-up / allocations
            dXo.copy_(torch.from_numpy(x0))
            torch.cuda.synchronize()
            t = time.perf_counter()
            its, info = prob.lbfgs_batch_dev(Bs, nx, dXo.data_ptr(), coeff, dco.data_ptr(), reg, 10, 30, 0.0)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            extra = dict(extra or {})
            extra["device_lbfgs"] = {"pulses": Bs, "iterations": info["iterations"], "evaluations": info["evaluations"], "seconds": dt,
                                     "evals_per_s_inside_optimiser": info["evaluations"] * Bs / dt,
                                     "pulse_iterations_per_s": float(np.sum(its)) / dt,
                                     "median_infidelity_after": float(torch.median(dco).item()),
                                     "note": "rg_lbfgs_batch_dev: iterates, gradients and curvature history stay in HBM; only a 16-byte "
                                             "progress record per round crosses PCIe (vs 131 MB per evaluation through the host API); every "
                                             "pulse runs its own line-search state machine, one batched evaluation per round"}
            del dXo, dco
        except Exception as ex:      # noqa: BLE001
            extra = dict(extra or {})
            extra["device_lbfgs_error"] = str(ex)

    # ---- dense-Hamiltonian instantiation of the same kernels (RG_DENSE=1): FP64-roofline evidence
    dense = None
    if rank == 0 and not args.no_extra and args.nerr == 0:
        os.environ["RG_DENSE"] = "1"
        try:
            pd = Problem(make_problem(N, 0), ctx)
            for _ in range(2):
                pd.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), coeff, dcost.data_ptr(), dgrad.data_ptr())
            ctx.set_timing(True); ctx.get_timing(reset=True)
            for _ in range(3):
                pd.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), coeff, dcost.data_ptr(), dgrad.data_ptr())
            dense = ctx.get_timing(reset=True)
            ctx.set_timing(False)
            pd.close()
        finally:
            os.environ.pop("RG_DENSE", None)

    if rank == 0:
        m = taylor_degree_for(T0 / N * 0.7071067811865476 * 1.0826 * 1.001 + 2e-4 * T0 / N)   # what the dense kernels pick
        canonical = canonical_flops(5, N, 1, 1, args.nerr)
        kernel_ms = {k: (v[0] / max(1, v[1])) for k, v in timing.items() if v[1]}
        dom = max(kernel_ms, key=kernel_ms.get)
        dom_ms = kernel_ms[dom]
        hbm_peak = None
        mp = ROOT / "MEASURED_PEAKS.json"
        if mp.exists():
            try:
                hbm_peak = json.loads(mp.read_text()).get("hbm_gbs")
            except Exception:
                hbm_peak = None
        peak_src = "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
        if hbm_peak is None:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
        # Executed FP64 flops per (pulse, time step) of the dominant kernel, counted from the ncu source page of that kernel
        # (thread-level DFMA x 2 + DMUL + DADD; profiles/r02_kernel_flops.json holds the counts and how they were taken).
        fl = {}
        fp_ = ROOT / "profiles" / "r02_kernel_flops.json"
        if fp_.exists():
            try:
                fl = json.loads(fp_.read_text())
            except Exception:
                fl = {}
        path = prob.path(nx)                       # which kernel family ran (rg_problem_path)
        pcs = "_pc" if path == "fused_q_pc" else ""
        names = {"k_grad": "k_fused_q<5, CZ drive mask, fidelity role" + (", phase-only class" if pcs else "") + "> (forward sweep + scan + "
                           "fidelity algebra + backward sweep in one launch; " +
                           ("step constants of the closed-form block propagators evaluated once per problem, one sincos per step and sweep)"
                            if pcs else "closed-form block propagators recomputed in both sweeps)"),
                 "k_grad_err": "k_fused_q<5, CZ drive mask, error role" + (", phase-only class" if pcs else "") + ">",
                 "k_steps": "k_steps_t", "k_chunk_agg": "k_agg_b2", "k_scan": "k_scan"}
        key = {"k_grad": "k_fused_q" + pcs + "_e0", "k_grad_err": "k_fused_q" + pcs + "_err"}.get(dom)
        per_step = (fl.get(key) or {}).get("fp64_flops_per_pulse_step")
        traffic = (fl.get(key) or {}).get("dram_bytes_per_launch")
        exec_eval = per_step * N if per_step else None
        alg_bytes = Bs * (nx * 8.0 * (2 if dom == "k_grad" else 1) + (nx + 1) * 8.0)      # x read by both sweeps, [cost | grad] written
        ach_tf = (min(exec_eval, canonical) * Bs / (dom_ms * 1e-3) / 1e12) if exec_eval else None
        roof = {
            "bound": "fp64", "kernel": names.get(dom, dom),
            "achieved": ach_tf, "peak": peak_dfma, "unit": "TFLOP/s", "frac": (ach_tf / peak_dfma) if (ach_tf and peak_dfma) else None,
            "traffic": traffic,
            "peak_source": "DFMA-loop microbenchmark run by this process (MEASURED_PEAKS.json has no FP64 figure); DMMA m8n8k4: "
                           + (f"{peak_dmma:.1f} TFLOP/s" if peak_dmma else "n/a"),
            "flops_per_eval": {"canonical_survey_8d": canonical, "executed_dominant_kernel": exec_eval,
                               "rule": "achieved uses min(executed, canonical) (SURVEY 8d)"},
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_GBps": alg_bytes / (dom_ms * 1e-3) / 1e9, "peak_GBps": hbm_peak,
                    "frac": alg_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                    "note": "the fused path touches only x and [cost | grad] in HBM (SURVEY 8d: 16,024 B per pulse); it is FP64/issue bound"},
            "path": path,
            "note": "workspace-free fused path: one launch per role; the block structure of the Rydberg model is exploited (closed-form "
                    "2x2 propagators, quaternion state, and for phase-only drives step-independent cos/sinc), so executed flops are far "
                    "below the canonical dense count; dense_fp64 shows the dense-H instantiation of the general kernels on the same workload",
            "fp64": {"achieved_tflops": ach_tf, "peak_dfma_tflops": peak_dfma, "peak_dmma_tflops": peak_dmma},
            "kernel_ms": kernel_ms,
        }
        if roof["fp64"]["peak_dfma_tflops"] and ach_tf:
            roof["fp64"]["frac"] = ach_tf / peak_dfma
        if dense:
            dk_ms = dense["k_steps"][0] / max(1, dense["k_steps"][1])
            ex_d = k_steps_flops(N, m, False)
            ach = min(ex_d, canonical) * Bs / (dk_ms * 1e-3) / 1e12
            roof["dense_fp64"] = {"kernel": "k_steps_t<5, full mask> (RG_DENSE=1, same workload treated as a dense Hamiltonian)",
                                  "bound": "fp64", "achieved": ach, "peak": peak_dfma, "unit": "TFLOP/s",
                                  "frac": ach / peak_dfma if peak_dfma else None,
                                  "flops_per_eval": {"canonical_survey_8d": canonical, "executed_k_steps": ex_d},
                                  "kernel_ms": {k: (v[0] / max(1, v[1])) for k, v in dense.items() if v[1]},
                                  "evals_per_s": Bs / (sum(v[0] / max(1, v[1]) for v in dense.values() if v[1]) * 1e-3)}
        line = {
            "metric": "GRAPE cost+grad evals/sec (CZ, batched pulses)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, scaling, gather),
            "clocks": run["clocks"],
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": B * nx * 8, "d2h_bytes_per_step": B * (nx + 1) * 8},
            "gpu_launches": launches * world,
            "roofline": roof,
        }
        if other:
            extra = dict(extra or {})
            extra[other["scaling"] + "_scaling"] = other
        if world > 1 and gather != "none":
            # the collective's own roofline: bytes every rank must receive per step over NVLink (ingress is the bound of an all-gather,
            # with or without switch multicast) against the 900 GB/s per direction of NVLink 5
            gb = run["gather_bytes_in_per_rank_per_step"]
            step_s = ms_total / args.steps * 1e-3
            line["gather"] = {"what": "costs of every rank on every rank; gradients stay with the rank whose optimiser consumes them"
                                      if args.gather_what == "costs" else "[cost | grad] of every rank on every rank (north_star)",
                              "how": gather, "bytes_in_per_rank_per_step": gb, "nvlink_ingress_GBps": gb / step_s / 1e9,
                              "nvlink_peak_GBps": 900.0, "frac": gb / step_s / 1e9 / 900.0, "floor_ms_per_step": gb / 900e9 * 1e3,
                              "note": "extra.full_gather / extra.costs_only_gather holds the same run with the other payload"}
        if other_what:
            extra = dict(extra or {})
            extra["full_gather" if other_what["what"] == "all" else "costs_only_gather"] = other_what
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_port
            threads = os.cpu_count() or 1
            sample = min(B, max(8, args.cpu_pulses_per_thread * threads))
            pp = cpu_port.PortProblem(make_problem(N, args.nerr))
            t = time.perf_counter()
            c_cpu, g_cpu = pp.cost_and_grad_batch(Xs[:sample].T, coeff, threads)
            dt = time.perf_counter() - t
            line["cpu_baseline"] = {"value": sample / dt, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} pulses of the workload, C++ port of the reference's literal "
                                              "algorithm (oracle/cpu_port.cpp); Julia is not installed in this image",
                                    "max_abs_cost_diff_vs_gpu": float(np.abs(c_cpu - cost_host[:sample]).max())}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
