"""ORACLE (test infrastructure, NOT product code) -- exact-semantics evaluation of the
reference's formulas in mpmath (default 50 digits).  PARITY UNPINNED (no Julia here).

What "exact semantics" means: the inputs are the FP64 numbers the reference would see --
x, the perturbed inputs fl(x+eps) / fl(x+eps2) (reference src/UnitaryCalculations.jl:50,53,
58,61,76,88), dt = fl(t0/ntimes) (:30), 1/eps and 1/eps2^2 as FP64 (:52,80), and
fl(1+err) inside the Rydberg builders (reference src/RydbergTools.jl:34-37) -- and every
operation after that (cos/sin, matrix exponential, inverse, products, traces) is carried
out without rounding.  The FP64 restatement (oracle/reference_oracle.py) and any FP64
implementation of the reference differ from these values only by rounding noise, which the
finite-difference quotients amplify by 1/eps (1e8) and 1/eps2^2 (1e8); this oracle is what
lets tests measure that noise floor and check the CUDA path below it.

Works only with descriptor callables (robustgrape_b200.descriptors), which accept `lib=mp`.
Sized for small problems (ntimes <= ~60).
"""
from __future__ import annotations

import numpy as np
import mpmath as mp


def _mpx(v):
    return [mp.mpf(float(a)) for a in v]


def _H(M):
    return M.transpose_conj()


def _tr(M):
    return sum(M[i, i] for i in range(M.rows))


def _tomp(A):
    A = np.asarray(A)
    M = mp.zeros(A.shape[0], A.shape[1])
    for i in range(A.shape[0]):
        for j in range(A.shape[1]):
            M[i, j] = mp.mpc(float(np.real(A[i, j])), float(np.imag(A[i, j])))
    return M


def calculate_unitary_and_derivatives(problem, x, dps=50):
    """Same outputs as reference src/UnitaryCalculations.jl:20-155, as nested lists of mp matrices:
    (U, U_dx[np][nt], U_dx_add[npa], U_derr[ne], U_derr_dx[np][nt][ne], U_derr_dx_add[npa][ne])."""
    mp.mp.dps = dps
    x = np.asarray(x, dtype=np.float64)
    na, N, d = problem.nb_additional_param, problem.ntimes, problem.ndim
    xm = x[: len(x) - na]
    p = len(xm) // N
    xm = xm.reshape((p, N), order="F")
    xa = x[len(x) - na:].copy()
    eps, eps2 = float(problem.eps), float(problem.eps2)
    dt = mp.mpf(float(problem.t0 / problem.ntimes))
    inv_eps = mp.mpf(float(1 / eps))
    inv_eps2sq = mp.mpf(float(1 / eps2 ** 2))
    nerr = len(problem.error_sources)
    H0 = problem.H0
    mI = mp.mpc(0, -1)

    def prop(k, xk, xadd, herr=None, err=0.0):
        Hm = H0(k, _mpx(xk), _mpx(xadd), lib=mp)
        if herr is not None:
            Hm = herr(k, _mpx(xk), _mpx(xadd), mp.mpf(float(err)), lib=mp) + Hm
        return mp.expm(mI * dt * Hm)

    C = mp.eye(d)
    A_dx = [[None] * N for _ in range(p)]
    A_dxa = [[None] * N for _ in range(na)]
    A_derr = [[None] * N for _ in range(nerr)]
    A_derr_dx = [[[None] * nerr for _ in range(N)] for _ in range(p)]
    A_derr_dxa = [[[None] * nerr for _ in range(N)] for _ in range(na)]
    for nt in range(1, N + 1):
        xk = xm[:, nt - 1]
        U = prop(nt, xk, xa)
        Cold = C
        C = U * C
        Cinv = mp.inverse(C)
        U2_dx, U2_dxa = [], []
        for i in range(p):
            xc = xk.copy(); xc[i] += eps
            A_dx[i][nt - 1] = Cinv * (inv_eps * (prop(nt, xc, xa) - U)) * Cold
            xc[i] = xk[i] + eps2
            U2_dx.append(prop(nt, xc, xa))
        for j in range(na):
            xc = xa.copy(); xc[j] += eps
            A_dxa[j][nt - 1] = Cinv * (inv_eps * (prop(nt, xk, xc) - U)) * Cold
            xc[j] = xa[j] + eps2
            U2_dxa.append(prop(nt, xk, xc))
        for e in range(nerr):
            herr = problem.error_sources[e].Herror
            A_derr[e][nt - 1] = Cinv * (inv_eps * (prop(nt, xk, xa, herr, eps) - U)) * Cold
            U2e = prop(nt, xk, xa, herr, eps2)
            for i in range(p):
                xc = xk.copy(); xc[i] += eps2
                A_derr_dx[i][nt - 1][e] = Cinv * (inv_eps2sq * (prop(nt, xc, xa, herr, eps2) + U - U2e - U2_dx[i])) * Cold
            for j in range(na):
                xc = xa.copy(); xc[j] += eps2
                A_derr_dxa[j][nt - 1][e] = Cinv * (inv_eps2sq * (prop(nt, xk, xc, herr, eps2) + U - U2e - U2_dxa[j])) * Cold

    Z = mp.zeros(d, d)
    U_dx = [[C * A_dx[i][k] for k in range(N)] for i in range(p)]
    U_dx_add = [C * sum(A_dxa[j], Z) for j in range(na)]
    U_derr, U_derr_dx, U_derr_dx_add = [], [[[None] * nerr for _ in range(N)] for _ in range(p)], [[None] * nerr for _ in range(na)]
    for e in range(nerr):
        U_derr.append(C * sum(A_derr[e], Z))
        cs, acc = [], Z
        for k in range(N):
            acc = acc + A_derr[e][k]; cs.append(acc)
        rcs, acc = [None] * N, Z
        for k in range(N - 1, -1, -1):
            acc = acc + A_derr[e][k]; rcs[k] = acc
        for i in range(p):
            for k in range(N):
                T = A_derr_dx[i][k][e]
                if k >= 1:
                    T = T + A_dx[i][k] * cs[k - 1]
                if k < N - 1:
                    T = T + rcs[k + 1] * A_dx[i][k]
                U_derr_dx[i][k][e] = C * T
        for j in range(na):
            T = Z
            for k in range(N):
                T = T + A_derr_dxa[j][k][e]
                if k >= 1:
                    T = T + A_dxa[j][k] * cs[k - 1]
                if k < N - 1:
                    T = T + rcs[k + 1] * A_dxa[j][k]
            U_derr_dx_add[j][e] = C * T
    return C, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add


def calculate_fidelity_and_derivatives(fp, x, dps=50):
    """Same outputs as reference src/FidelityCalculations.jl:19-119, as float64 arrays rounded
    from the high-precision values."""
    mp.mp.dps = dps
    up = fp.unitary_problem
    U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add = calculate_unitary_and_derivatives(up, x, dps)
    x = np.asarray(x, dtype=np.float64)
    na, N, d = up.nb_additional_param, up.ntimes, up.ndim
    p = (len(x) - na) // N
    nerr = len(up.error_sources)
    xa = x[len(x) - na:].copy()
    eps = float(up.eps)
    inv_eps = mp.mpf(float(1 / eps))
    U0 = fp.target_unitary(_mpx(xa), lib=mp)
    V = []
    for j in range(na):
        xc = xa.copy(); xc[j] += eps
        V.append(inv_eps * (fp.target_unitary(_mpx(xc), lib=mp) - U0))
    P0 = _tomp(np.asarray(fp.projector, dtype=np.float64))
    Pn = np.asarray(fp.projector, dtype=np.float64).copy(); Pn[Pn != 0] = 1
    P = _tomp(Pn)
    D = mp.re(_tr(P0))
    DD = D * (D + 1)
    trm = lambda A: _tr(P0 * A)
    U0h, Uh = _H(U0), _H(U)
    tau = trm(P * U0h * U)
    F = (mp.re(trm(P * U0h * U * P * Uh * U0)) + abs(tau) ** 2) / DD

    def dF(X, Vj=None):
        t = trm(P * U0h * X * P * Uh * U0 + P * U0h * U * P * _H(X) * U0)
        s = trm(P * U0h * X)
        if Vj is not None:
            t = t + trm(P * _H(Vj) * U * P * Uh * U0 + P * U0h * U * P * Uh * Vj)
            s = s + trm(P * _H(Vj) * U)
        return (mp.re(t) + 2 * mp.re(mp.conj(tau) * s)) / DD

    F_dx = np.zeros(p * N + na)
    for k in range(N):
        for i in range(p):
            F_dx[i + p * k] = float(dF(U_dx[i][k]))
    for j in range(na):
        F_dx[p * N + j] = float(dF(U_dx_add[j], V[j]))
    F_d2err = np.zeros(nerr)
    F_d2err_dx = np.zeros((p * N + na, nerr))
    for e in range(nerr):
        E = U_derr[e]; Eh = _H(E)
        te = trm(P * U0h * E)
        F_d2err[e] = float(2 * (mp.re(trm(P * U0h * E * P * Eh * U0 - P * Eh * E)) + abs(te) ** 2
                                - D * mp.re(trm(P * Eh * E))) / DD)

        def d2(Zm, Vj=None):
            Zh = _H(Zm)
            t = trm(P * U0h * Zm * P * Eh * U0 + P * U0h * E * P * Zh * U0 - P * Zh * E - P * Eh * Zm)
            s = trm(P * U0h * Zm)
            if Vj is not None:
                t = t + trm(P * _H(Vj) * E * P * Eh * U0 + P * U0h * E * P * Eh * Vj)
                s = s + trm(P * _H(Vj) * E)
            return 2 * (mp.re(t) + 2 * mp.re(mp.conj(te) * s) - D * mp.re(trm(P * Zh * E + P * Eh * Zm))) / DD

        for k in range(N):
            for i in range(p):
                F_d2err_dx[i + p * k, e] = float(d2(U_derr_dx[i][k][e]))
        for j in range(na):
            F_d2err_dx[p * N + j, e] = float(d2(U_derr_dx_add[j][e], V[j]))
    return float(F), F_dx, F_d2err, F_d2err_dx
