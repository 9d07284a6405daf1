"""Regularization mirror (reference src/Regularization.jl:26-47,78-83,111-115).
Host-side O(ntimes) vector work on values already returned to the host (out of kernel scope)."""
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
