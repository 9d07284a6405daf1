"""UnitaryCalculations mirror (reference src/UnitaryCalculations.jl) on the CUDA library."""
from __future__ import annotations

import numpy as np

from ._lib import HStackProblem, Problem, _ptr, is_descriptor_problem


def _fingerprint(problem):
    """Fields the device twin depends on: a changed struct gets a new twin (the reference re-reads its fields on every call)."""
    up = getattr(problem, "unitary_problem", problem)
    proj = getattr(problem, "projector", None)
    return (id(up.H0), tuple(id(s.Herror) for s in up.error_sources), id(getattr(problem, "target_unitary", None)),
            float(up.t0), int(up.ntimes), int(up.ndim), int(up.nb_additional_param), float(up.eps), float(up.eps2),
            None if proj is None else np.asarray(proj, dtype=np.float64).tobytes())


def device_problem(problem, ctx=None):
    """Device-resident twin of a problem struct, cached on the struct itself and rebuilt when a field changed.  Descriptor
    problems evaluate their Hamiltonians on the device (Problem); closure problems go through host-evaluated stacks (HStackProblem)."""
    dp = getattr(problem, "_rg_device_problem", None)
    fpn = _fingerprint(problem)
    if dp is None or (ctx is not None and dp.ctx is not ctx) or getattr(problem, "_rg_device_fingerprint", None) != fpn:
        dp = Problem(problem, ctx) if is_descriptor_problem(problem) else HStackProblem(problem, ctx)
        try:
            object.__setattr__(problem, "_rg_device_fingerprint", fpn)
        except Exception:
            pass
        try:
            object.__setattr__(problem, "_rg_device_problem", dp)
        except Exception:
            pass
    return dp


def _nparam(problem, x):
    n = len(x) - problem.nb_additional_param
    assert n >= 0 and n % problem.ntimes == 0, "Control parameter size must be a multiple of time steps"
    return n // problem.ntimes


def calculate_unitary_and_derivatives(problem, x, ctx=None):
    """reference src/UnitaryCalculations.jl:20-155.
    Returns (U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add) with the reference's shapes
    (d,d), (d,d,p,N), (d,d,a), (d,d,e), (d,d,p,N,e), (d,d,a,e), as complex128 arrays
    (the reference's containers have abstract eltype `Complex`; values are the same)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    dp = device_problem(problem, ctx)
    if isinstance(dp, HStackProblem):
        return dp.unitary_and_derivatives(x)
    h, p = dp.handle_for(len(x))
    d, N, a, e = problem.ndim, problem.ntimes, problem.nb_additional_param, len(problem.error_sources)
    z = lambda *s: np.zeros(s, dtype=np.complex128, order="F")
    U, U_dx, U_dx_add, U_derr = z(d, d), z(d, d, p, N), z(d, d, a), z(d, d, e)
    U_derr_dx, U_derr_dx_add = z(d, d, p, N, e), z(d, d, a, e)
    dp.ctx.check(dp.ctx.lib.rg_unitary_and_derivatives(h, _ptr(x), _ptr(U), _ptr(U_dx), _ptr(U_dx_add), _ptr(U_derr),
                                                       _ptr(U_derr_dx), _ptr(U_derr_dx_add)))
    return U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add


def calculate_interaction_error_operators(problem, x, ctx=None):
    """reference src/UnitaryCalculations.jl:180-204.  Returns (ndim, ndim, ntimes, nerr) complex128."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    dp = device_problem(problem, ctx)
    h, p = dp.handle_for(len(x))
    O = np.zeros((problem.ndim, problem.ndim, problem.ntimes, len(problem.error_sources)), dtype=np.complex128, order="F")
    dp.ctx.check(dp.ctx.lib.rg_interaction_error_operators(h, _ptr(x), _ptr(O)))
    return O
