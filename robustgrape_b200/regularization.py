"""Regularization (reference src/Regularization.jl:26-47,78-83,111-115).

The three forms the reference and its tests use are *enumerated* (`_rg_kind`): when every entry of
`FidelityRobustGRAPEParameters.regularization_functions` is one of them, the terms are evaluated by the device epilogue
k_regularize (csrc/rg_optim.cuh) and the optimiser loop never leaves the GPU; arbitrary callables (src/Types.jl:76) still work
through the host path below."""
from __future__ import annotations

import numpy as np


def regularization_cost(x, f=None, df=None):
    x = np.asarray(x, dtype=np.float64)
    if f is not None:
        r1, j1, r2, j2 = regularization_cost(f(x))
        d = df(x)
        return r1, d * j1, r2, d * j2
    n = len(x)
    dx = np.diff(x)
    ddx = np.diff(dx)
    reg1, reg2 = float(np.sum(dx ** 2)), float(np.sum(ddx ** 2))
    jac1, jac2 = np.zeros(n), np.zeros(n)
    jac1[1:n - 1] = -2 * ddx
    jac1[0] += -2 * dx[0]
    jac1[n - 1] += 2 * dx[n - 2]
    jac2[0] = 2 * (x[2] - 2 * x[1] + x[0])
    jac2[1] = 2 * (x[3] - 4 * x[2] + 5 * x[1] - 2 * x[0])
    i = np.arange(2, n - 2)
    jac2[i] = 2 * (x[i + 2] - 4 * x[i + 1] + 6 * x[i] - 4 * x[i - 1] + x[i - 2])
    jac2[n - 2] = 2 * (x[n - 4] - 4 * x[n - 3] + 5 * x[n - 2] - 2 * x[n - 1])
    jac2[n - 1] = 2 * (x[n - 3] - 2 * x[n - 2] + x[n - 1])
    return reg1, jac1, reg2, jac2


def regularization_cost_phase(phis):
    c = regularization_cost(phis, np.cos, lambda v: -np.sin(v))
    s = regularization_cost(phis, np.sin, np.cos)
    return c[0] + s[0], c[1] + s[1], c[2] + s[2], c[3] + s[3]


def regularization_cost_phase_sin2(x):
    """The sin^2-of-differences form of the reference's own tests (test/runtests.jl:9-45), quirks included: the gradient loops
    as written there leave jac1[n] at zero."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    dx = np.diff(x)
    ddx = np.diff(dx)
    reg1, reg2 = float(np.sum(np.sin(dx / 2) ** 2)), float(np.sum(np.sin(ddx / 2) ** 2))
    jac1, jac2 = np.zeros(n), np.zeros(n)
    for i in range(1, n):                     # 1-based i = 1 .. n-1
        if i < n - 1:
            jac1[i - 1] -= 0.5 * np.sin(dx[i - 1])
        if i > 1:
            jac1[i - 1] += 0.5 * np.sin(dx[i - 2])
    for i in range(1, n + 1):
        if i < n - 2:
            jac2[i - 1] -= 0.5 * np.sin(ddx[i - 1])
        if 1 < i < n - 1:
            jac2[i - 1] += np.sin(ddx[i - 2])
        if i > 2:
            jac2[i - 1] -= 0.5 * np.sin(ddx[i - 3])
    return reg1, jac1, reg2, jac2


RG_REG_NONE, RG_REG_PLAIN, RG_REG_PHASE, RG_REG_SIN2 = 0, 1, 2, 3
regularization_cost._rg_kind = RG_REG_PLAIN
regularization_cost_phase._rg_kind = RG_REG_PHASE
regularization_cost_phase_sin2._rg_kind = RG_REG_SIN2


def device_kinds(functions):
    """Enumerated kinds of a list of regularisation callables, or None when one of them is an arbitrary closure."""
    kinds = [getattr(f, "_rg_kind", None) for f in functions]
    return None if any(k is None for k in kinds) else kinds
