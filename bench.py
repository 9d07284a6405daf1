#!/usr/bin/env python
"""bench.py -- GRAPE cost+grad evaluations/s on the multi-start CZ workload (BASELINE.json config C4).

  python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1)
  python bench.py --impl reference ...                     (CPU arm: C++ port of the reference's literal algorithm)

One "step" = one cost+gradient evaluation (reference `calculate_common!`, src/FidelityCalculations.jl:174-184,
without host regularisation) of every pulse of the batch: 8192 random-init pulses x 1000 time steps of the
5-level symmetric-blockaded Rydberg CZ problem (examples/time_optimal_cz.jl), sharded over the ranks
(strong scaling: the batch is fixed), followed for N > 1 by one NCCL all-gather of [cost | grad].
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


def make_pulses(ntimes, batch, seed=43):
    """phi_k ~ 2 pi U(0,1) (test/runtests.jl:91), theta ~ 2 pi U(0,1); one pulse per row (C order)
    == one pulse per column of the (nx, B) column-major array the C ABI takes."""
    rng = np.random.default_rng(seed)
    return 2 * np.pi * rng.random((batch, ntimes + 1))


def flops_per_eval(d, N, p, a, e, taylor_m, nvar):
    """(canonical, executed) real flops per cost+grad evaluation.
    canonical: SURVEY.md section 8(d): 8 d^3 N (n_exp_needed c_exp + 4 + 8 e), Pade-3 regime (c_exp = 2 + 4/3).
    executed: what the kernels actually issue (one d x d complex product = 8 d^3)."""
    prod = 8.0 * d ** 3
    n_exp_needed = 1 + p + a + e * (2 + p + a) + ((p + a) if e > 0 else 0)
    canonical = prod * N * (n_exp_needed * (2 + 4.0 / 3.0) + 4 + 8 * e)
    m1 = taylor_m - 1
    k_steps = max(nvar + e, 1) * 3 * m1        # Horner passes: A*Y, A*Dl, dA*(Y+Dl) per iteration, per first-order object
    k_steps += 1 + 2 * e                       # chunk aggregates: q <- U q ; wl_e <- U wl_e + D_e q
    k_so = e * nvar * 9 * m1                   # mixed second differences
    k_grad = (2 + nvar) + e * (6 + 3 * nvar)   # backward sweeps (rewind, co-state advance, contractions)
    executed = prod * N * (k_steps + k_so + k_grad)
    return canonical, executed, prod * N * k_steps


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
            time.sleep(0.02)

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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} pulses of the workload per step (C++ port of the reference's literal "
                                   "algorithm; Julia unavailable in this image)"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"C4 multi-start CZ: {args.batch} random-init pulses x {args.ntimes} time steps, d=5 symmetric-blockaded "
                    f"Rydberg, p=1, a=1, e={args.nerr}, t0={T0}, eps=1e-8, eps2=1e-4 (examples/time_optimal_cz.jl)",
        "batch": args.batch, "ntimes": args.ntimes, "nerr": args.nerr,
        "sharding": "pulses over ranks, NCCL all-gather of [cost|grad]",
        "l2": "per-step workspace traffic (step propagators, ~0.8 MB/pulse) is far larger than the 126 MB L2; no explicit flush",
    }


def main():
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    B, N = args.batch, args.ntimes
    nx = N + 1
    assert B % world == 0, "batch must divide over the ranks"
    Bs = B // world
    Xall = make_pulses(N, B)
    Xs = np.ascontiguousarray(Xall[rank * Bs:(rank + 1) * Bs])
    coeff = [1e-4] * args.nerr

    ctx = Context(local)
    # All work (library kernels, NCCL, timing events) goes on one non-default torch stream.
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    prob = Problem(make_problem(N, args.nerr), ctx)

    dX = torch.from_numpy(Xs).to(dev)                              # inputs resident in HBM
    out_local = torch.empty(Bs * (1 + nx), dtype=torch.float64, device=dev)   # [cost (Bs) | grad (Bs, nx)]
    out_all = torch.empty(world * Bs * (1 + nx), dtype=torch.float64, device=dev) if world > 1 else out_local
    dcost, dgrad = out_local[:Bs], out_local[Bs:]

    def step():
        prob.cost_and_grad_batch_dev(Bs, nx, dX.data_ptr(), coeff, dcost.data_ptr(), dgrad.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(out_all, out_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_dfma = peak_dmma = None
    if rank == 0:
        peak_dfma, peak_dmma = ctx.measure_fp64_peak(0.3)

    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    barrier()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    launches = ctx.launch_count - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = B * args.steps / (ms_total * 1e-3)
    ctx.synchronize()
    cost_host = dcost.cpu().numpy()

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

    if rank == 0:
        m = 6 if args.ntimes >= 400 else (8 if args.ntimes >= 90 else 10)     # Taylor degree the kernel picks for dt*||H||_1
        canonical, executed, exec_k1 = flops_per_eval(5, N, 1, 1, args.nerr, m, 1)
        k1_ms, k1_n = timing["k_steps"]
        k1_avg = k1_ms / max(1, k1_n)
        achieved = min(exec_k1, canonical) * Bs / (k1_avg * 1e-3) / 1e12
        traffic = None
        tp = ROOT / "profiles" / "k_steps_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "GRAPE cost+grad evals/sec (CZ, batched pulses)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": B * nx * 8, "d2h_bytes_per_step": B * (nx + 1) * 8},
            "gpu_launches": launches * world,
            "roofline": {
                "bound": "fp64", "kernel": "k_steps<5>", "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s",
                "frac": achieved / peak_dfma if peak_dfma else None, "traffic": traffic,
                "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure); "
                               f"DMMA m8n8k4 microbenchmark: {peak_dmma:.2f} TFLOP/s",
                "flops_per_eval": {"canonical": canonical, "executed_total": executed, "executed_k_steps": exec_k1},
                "kernel_ms": {k: (v[0] / max(1, v[1])) for k, v in timing.items() if v[1]},
                "whole_step_frac_of_peak": min(executed, canonical) * B / (ms_total / args.steps * 1e-3) / 1e12 / (peak_dfma * world) if peak_dfma else None,
            },
        }
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_port
            threads = os.cpu_count() or 1
            sample = min(B, max(8, args.cpu_pulses_per_thread * threads))
            pp = cpu_port.PortProblem(make_problem(N, args.nerr))
            t = time.perf_counter()
            c_cpu, g_cpu = pp.cost_and_grad_batch(Xall[:sample].T, coeff, threads)
            dt = time.perf_counter() - t
            line["cpu_baseline"] = {"value": sample / dt, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} pulses of the workload, C++ port of the reference's literal "
                                              "algorithm (oracle/cpu_port.cpp); Julia is not installed in this image",
                                    "max_abs_cost_diff_vs_gpu": float(np.abs(c_cpu - cost_host[:sample]).max())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
