"""FidelityCalculations mirror (reference src/FidelityCalculations.jl) on the CUDA library.

Single-pulse functions keep the reference's names, argument order and return shapes; the
`*_batch` functions are the batched entry points the reference has no counterpart for
(one pulse per column of X)."""
from __future__ import annotations

import math
import time

import numpy as np

from ._lib import HStackProblem, _ptr
from .unitary_calculations import device_problem


def calculate_fidelity_and_derivatives(fidelity_problem, x, ctx=None):
    """reference src/FidelityCalculations.jl:19-119 -> (F, F_dx_tot, F_d2err, F_d2err_dx_tot)."""
    x = np.asarray(x, dtype=np.float64)
    dp = device_problem(fidelity_problem, ctx)
    if isinstance(dp, HStackProblem):                       # closure problem: host-evaluated Hamiltonian stack
        return dp.fidelity_and_derivatives(x)
    F, Fdx, F2, F2dx = dp.fidelity_and_derivatives_batch(x[:, None])
    return float(F[0]), Fdx[:, 0].copy(), F2[:, 0].copy(), F2dx[:, :, 0].copy()


def calculate_fidelity_and_derivatives_batch(fidelity_problem, X, ctx=None, want_grad=True):
    """X (nx, B) -> F (B), F_dx (nx,B), F_d2err (nerr,B), F_d2err_dx (nx,nerr,B)."""
    return device_problem(fidelity_problem, ctx).fidelity_and_derivatives_batch(X, want_grad)


def cost_and_gradient_batch(fidelity_problem, X, error_source_coeff=(), ctx=None):
    """`calculate_common!` of reference src/FidelityCalculations.jl:174-184 for every column of X,
    without the host-side regularisation terms: cost (B), grad (nx, B)."""
    return device_problem(fidelity_problem, ctx).cost_and_grad_batch(X, error_source_coeff)


def optimize_batch_device(fidelity_problem, X0, error_source_coeff=(), regularization_functions=None, regularization_coeff1=None,
                          regularization_coeff2=None, iterations=1000, g_tol=1e-8, history=10, ctx=None):
    """Multi-start optimisation with the whole loop on the device: one L-BFGS per column of X0 (nx, B), cost =
    calculate_common! (src/FidelityCalculations.jl:174-197) with enumerated regularisation (regularization.device_kinds).
    Returns (X, cost, iterations per pulse, info)."""
    from . import regularization as R
    dp = device_problem(fidelity_problem, ctx)
    reg = None
    if regularization_functions:
        kinds = R.device_kinds(regularization_functions)
        if kinds is None:
            raise TypeError("device optimisation needs enumerated regularisation functions (robustgrape_b200.regularization); "
                            "arbitrary callables run through optimize_fidelity_and_error_sources")
        reg = list(zip(kinds, regularization_coeff1, regularization_coeff2))
    return dp.lbfgs_batch(X0, error_source_coeff, reg, history, iterations, g_tol)


def optimize_fidelity_and_error_sources(fidelity_problem, fidelity_parameters, ctx=None):
    """reference src/FidelityCalculations.jl:161-218.  The cost/gradient evaluation runs on the GPU.  With
    `solver_algorithm = "device-lbfgs"` and enumerated regularisation functions the optimiser loop itself stays on the device
    (batched L-BFGS, rg_lbfgs_batch); otherwise the optimiser (Optim.jl L-BFGS in the reference) is scipy's L-BFGS-B and arbitrary
    regularisation callables are evaluated on the host.  Returns an OptimizeResult (`.x` is `Optim.minimizer`)."""
    from scipy.optimize import minimize
    if fidelity_parameters.solver_algorithm == "device-lbfgs":
        from scipy.optimize import OptimizeResult
        prm = fidelity_parameters
        ap = dict(prm.additional_parameters)
        X, cost, iters, info = optimize_batch_device(fidelity_problem, np.asarray(prm.x_initial, dtype=np.float64)[:, None],
                                                     prm.error_source_coeff, prm.regularization_functions, prm.regularization_coeff1,
                                                     prm.regularization_coeff2, int(prm.iterations), float(ap.get("g_tol", 1e-8)),
                                                     int(ap.get("history", 10)), ctx)
        return OptimizeResult(x=X[:, 0].copy(), fun=float(cost[0]), nit=int(iters[0]), nfev=info["evaluations"], success=True,
                              message="device L-BFGS", info=info)
    up = fidelity_problem.unitary_problem
    prm = fidelity_parameters
    assert len(prm.error_source_coeff) == len(up.error_sources)
    ntimes, na = up.ntimes, up.nb_additional_param
    x0 = np.asarray(prm.x_initial, dtype=np.float64)
    nparam = (len(x0) - na) // ntimes
    assert len(prm.regularization_coeff1) == nparam
    assert len(prm.regularization_coeff2) == nparam
    assert len(prm.regularization_functions) == nparam
    dp = device_problem(fidelity_problem, ctx)
    t_start = time.time()

    def fg(x):
        cost, grad = dp.cost_and_grad_batch(x[:, None], prm.error_source_coeff)
        c, g = float(cost[0]), grad[:, 0].copy()
        xm = x[: len(x) - na].reshape((nparam, ntimes), order="F")
        rg = np.zeros((nparam, ntimes))
        for i in range(nparam):                                           # :189-195
            r1, j1, r2, j2 = prm.regularization_functions[i](xm[i, :].copy())
            c += prm.regularization_coeff1[i] * r1 + prm.regularization_coeff2[i] * r2
            rg[i, :] = prm.regularization_coeff1[i] * np.asarray(j1) + prm.regularization_coeff2[i] * np.asarray(j2)
        g[: len(x) - na] += rg.reshape(nparam * ntimes, order="F")
        return c, g

    class _Timeout(Exception):
        pass

    best = {"x": x0}

    def cb(xk):
        best["x"] = xk.copy()
        if not math.isnan(prm.time_limit) and time.time() - t_start > prm.time_limit:
            raise _Timeout()

    opts = {"maxiter": int(prm.iterations)}
    ap = dict(prm.additional_parameters)
    # Optim.Options names of the reference (src/FidelityCalculations.jl:219-231 passes additional_parameters straight through).
    # g_tol -> gtol (both bound the infinity norm of the gradient).  f_abstol -> ftol: scipy's L-BFGS-B test is
    # (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= ftol, and this cost lies in [0, 1 + O(coeff)], so the denominator is 1 and the test
    # is the absolute one Optim applies.  Anything else has no scipy counterpart: rejected loudly rather than silently dropped.
    if "g_tol" in ap:
        opts["gtol"] = ap.pop("g_tol")
    if "f_abstol" in ap:
        opts["ftol"] = ap.pop("f_abstol")
    if ap:
        raise ValueError(f"unsupported Optim option(s) {sorted(ap)}: only g_tol and f_abstol are mapped")
    try:
        return minimize(fg, x0, jac=True, method=prm.solver_algorithm, callback=cb, options=opts)
    except _Timeout:
        from scipy.optimize import OptimizeResult
        c, g = fg(best["x"])
        return OptimizeResult(x=best["x"], fun=c, jac=g, success=False, message="time_limit reached")


def calculate_fidelity_response(fidelity_problem, x, normalized_frequencies, ctx=None, first=0, count=None):
    """reference src/FidelityCalculations.jl:246-280 -> (nfreq, nerr).  `first`/`count` select a
    contiguous shard of the frequency grid (multi-GPU sharding); default is the whole grid."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    freqs = np.ascontiguousarray(normalized_frequencies, dtype=np.float64)
    dp = device_problem(fidelity_problem, ctx)
    h, p = dp.handle_for(len(x))
    count = len(freqs) - first if count is None else count
    R = np.zeros((count, dp.nerr), order="F")
    dp.ctx.check(dp.ctx.lib.rg_fidelity_response(h, _ptr(x), _ptr(freqs), len(freqs), first, count, _ptr(R)))
    return R


def calculate_fidelity_response_fft(fidelity_problem, x, oversampling=1, ctx=None):
    """reference src/FidelityCalculations.jl:306-343 -> (response (N*os, nerr), norm_frequencies)."""
    assert oversampling >= 1
    x = np.ascontiguousarray(x, dtype=np.float64)
    dp = device_problem(fidelity_problem, ctx)
    h, p = dp.handle_for(len(x))
    n = dp.ntimes * oversampling
    R = np.zeros((n, dp.nerr), order="F")
    fr = np.zeros(n)
    dp.ctx.check(dp.ctx.lib.rg_fidelity_response_fft(h, _ptr(x), int(oversampling), _ptr(R), _ptr(fr)))
    return R, fr


def calculate_expectation_values(fidelity_problem, x, ctx=None):
    """reference src/FidelityCalculations.jl:368-390 -> (ntimes, nerr)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    dp = device_problem(fidelity_problem, ctx)
    h, p = dp.handle_for(len(x))
    out = np.zeros((dp.ntimes, dp.nerr), order="F")
    dp.ctx.check(dp.ctx.lib.rg_expectation_values(h, _ptr(x), _ptr(out)))
    return out
