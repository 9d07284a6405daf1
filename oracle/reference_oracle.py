"""ORACLE (test infrastructure, NOT product code) -- literal CPU restatement of the
reference's GRAPE propagator / fidelity path in numpy complex128.

PARITY UNPINNED: the reference is pure Julia, Julia is not installed in this image,
and the reference ships no golden vectors (its tests are rtol=1e-3 self-consistency
checks, reference test/runtests.jl:106-111,164,289,347-352,415,523-524,613-614).
This file restates the reference's algorithm statement by statement; the arithmetic
the reference delegates to Julia's stdlib `LinearAlgebra` (`exp`, `inv`, `*`, `tr`;
version unpinned, no Manifest) is delegated here to `scipy.linalg.expm` (Pade
scaling-and-squaring, same family) and numpy.  It is pinned only by (i) reproducing
every test structure of reference test/runtests.jl (tests/test_oracle_*.py), and
(ii) the exact-semantics mpmath evaluation in oracle/exact_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Problem objects are duck-typed: anything with the reference's field names
(t0, ntimes, ndim, H0, nb_additional_param, error_sources[.Herror], eps, eps2 /
unitary_problem, projector, target_unitary) works, with plain Python closures or
the callable descriptors of robustgrape_b200.descriptors.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm as _expm


def _split(problem, x):
    """reference src/UnitaryCalculations.jl:21-26"""
    x = np.asarray(x, dtype=np.float64)
    na = problem.nb_additional_param
    x_main = x[: len(x) - na]
    assert len(x_main) % problem.ntimes == 0, "Control parameter size must be a multiple of time steps"
    nparam = len(x_main) // problem.ntimes
    x_main = x_main.reshape((nparam, problem.ntimes), order="F")
    x_add = x[len(x) - na:].copy()
    return x_main, x_add, nparam


def calculate_unitary_and_derivatives(problem, x, expm=_expm):
    """reference src/UnitaryCalculations.jl:20-155, same loop order and FD formulas."""
    x_main, x_add, nparam = _split(problem, x)
    ntimes, ndim, na = problem.ntimes, problem.ndim, problem.nb_additional_param
    x_add_copy = x_add.copy()
    cum_evo = np.eye(ndim, dtype=np.complex128)
    old_cum_evo = cum_evo.copy()
    dt = problem.t0 / problem.ntimes
    nerr = len(problem.error_sources)
    eps, eps2 = problem.eps, problem.eps2
    c = np.complex128
    infimU_dx = np.zeros((ndim, ndim, nparam, ntimes), c)
    infimU_dx_add = np.zeros((ndim, ndim, na, ntimes), c)
    infimU_derr = np.zeros((ndim, ndim, nerr, ntimes), c)
    infimU_derr_dx = np.zeros((ndim, ndim, nparam, nerr, ntimes), c)
    infimU_derr_dx_add = np.zeros((ndim, ndim, na, nerr, ntimes), c)
    infim_evo_derr_array = np.zeros((ndim, ndim, nerr), c)
    infim_evo_dx_array = np.zeros((ndim, ndim, nparam), c)
    infim_evo_dx_add_array = np.zeros((ndim, ndim, na), c)
    H0 = problem.H0

    for nt in range(1, ntimes + 1):                                              # :44
        xk = x_main[:, nt - 1]
        infim_evo = expm(-1j * dt * H0(nt, xk, x_add))                           # :45
        cum_evo = infim_evo @ cum_evo                                            # :46
        cum_evo_inv = np.linalg.inv(cum_evo)                                     # :47
        x_main_copy = xk.copy()
        for np_ in range(nparam):                                                # :49-56
            x_main_copy[np_] += eps
            infim_evo_dx = expm(-1j * dt * H0(nt, x_main_copy, x_add))
            infimU_dx[:, :, np_, nt - 1] = cum_evo_inv @ ((1 / eps) * (infim_evo_dx - infim_evo)) @ old_cum_evo
            x_main_copy[np_] = xk[np_] + eps2
            infim_evo_dx_array[:, :, np_] = expm(-1j * dt * H0(nt, x_main_copy, x_add))
            x_main_copy[np_] = xk[np_]
        for npa in range(na):                                                    # :57-64
            x_add_copy[npa] += eps
            infim_evo_dx_add = expm(-1j * dt * H0(nt, xk, x_add_copy))
            infimU_dx_add[:, :, npa, nt - 1] = cum_evo_inv @ ((1 / eps) * (infim_evo_dx_add - infim_evo)) @ old_cum_evo
            x_add_copy[npa] = x_add[npa] + eps2
            infim_evo_dx_add_array[:, :, npa] = expm(-1j * dt * H0(nt, xk, x_add_copy))
            x_add_copy[npa] = x_add[npa]
        for ne in range(nerr):                                                   # :66-98
            Herr = problem.error_sources[ne].Herror
            infim_evo_derr = expm(-1j * dt * (Herr(nt, xk, x_add, eps) + H0(nt, xk, x_add)))
            infimU_derr[:, :, ne, nt - 1] = cum_evo_inv @ ((1 / eps) * (infim_evo_derr - infim_evo)) @ old_cum_evo
            infim_evo_derr_array[:, :, ne] = expm(-1j * dt * (Herr(nt, xk, x_add, eps2) + H0(nt, xk, x_add)))
            for np_ in range(nparam):                                            # :75-85
                x_main_copy[np_] += eps2
                infim_evo_derr_dx = expm(-1j * dt * (Herr(nt, x_main_copy, x_add, eps2) + H0(nt, x_main_copy, x_add)))
                infimU_derr_dx[:, :, np_, ne, nt - 1] = cum_evo_inv @ ((1 / eps2 ** 2) * (
                    infim_evo_derr_dx + infim_evo
                    - infim_evo_derr_array[:, :, ne] - infim_evo_dx_array[:, :, np_])) @ old_cum_evo
                x_main_copy[np_] = xk[np_]
            for npa in range(na):                                                # :87-97
                x_add_copy[npa] += eps2
                infim_evo_derr_dx_add = expm(-1j * dt * (Herr(nt, xk, x_add_copy, eps2) + H0(nt, xk, x_add_copy)))
                infimU_derr_dx_add[:, :, npa, ne, nt - 1] = cum_evo_inv @ ((1 / eps2 ** 2) * (
                    infim_evo_derr_dx_add + infim_evo
                    - infim_evo_derr_array[:, :, ne] - infim_evo_dx_add_array[:, :, npa])) @ old_cum_evo
                x_add_copy[npa] = x_add[npa]
        old_cum_evo = cum_evo.copy()                                             # :99

    infimU_derr = np.transpose(infimU_derr, (0, 1, 3, 2))                        # :102
    infimU_derr_dx = np.transpose(infimU_derr_dx, (0, 1, 2, 4, 3))               # :103
    infimU_derr_dx_add = np.transpose(infimU_derr_dx_add, (0, 1, 2, 4, 3))       # :104

    U_dx = np.zeros((ndim, ndim, nparam, ntimes), c)
    U_dx_add = np.zeros((ndim, ndim, na), c)
    U_derr = np.zeros((ndim, ndim, nerr), c)
    U_derr_dx = np.zeros((ndim, ndim, nparam, ntimes, nerr), c)
    U_derr_dx_add = np.zeros((ndim, ndim, na, nerr), c)

    cumsum = np.cumsum(infimU_derr, axis=2)                                      # :112
    revcumsum = np.flip(np.cumsum(np.flip(infimU_derr, axis=2), axis=2), axis=2)  # :113
    for nt in range(ntimes):                                                     # :114-118
        for np_ in range(nparam):
            U_dx[:, :, np_, nt] = cum_evo @ infimU_dx[:, :, np_, nt]
    for npa in range(na):                                                        # :119-121
        U_dx_add[:, :, npa] = cum_evo @ infimU_dx_add[:, :, npa, :].sum(axis=2)
    for ne in range(nerr):                                                       # :122-152
        U_derr[:, :, ne] = cum_evo @ infimU_derr[:, :, :, ne].sum(axis=2)
        for nt in range(1, ntimes):
            for np_ in range(nparam):
                U_derr_dx[:, :, np_, nt, ne] += infimU_dx[:, :, np_, nt] @ cumsum[:, :, nt - 1, ne]
        for nt in range(ntimes - 1):
            for np_ in range(nparam):
                U_derr_dx[:, :, np_, nt, ne] += revcumsum[:, :, nt + 1, ne] @ infimU_dx[:, :, np_, nt]
        for nt in range(ntimes):
            for np_ in range(nparam):
                U_derr_dx[:, :, np_, nt, ne] += infimU_derr_dx[:, :, np_, nt, ne]
                U_derr_dx[:, :, np_, nt, ne] = cum_evo @ U_derr_dx[:, :, np_, nt, ne]
        for npa in range(na):
            for nt in range(1, ntimes):
                U_derr_dx_add[:, :, npa, ne] += infimU_dx_add[:, :, npa, nt] @ cumsum[:, :, nt - 1, ne]
            for nt in range(ntimes - 1):
                U_derr_dx_add[:, :, npa, ne] += revcumsum[:, :, nt + 1, ne] @ infimU_dx_add[:, :, npa, nt]
            for nt in range(ntimes):
                U_derr_dx_add[:, :, npa, ne] += infimU_derr_dx_add[:, :, npa, nt, ne]
            U_derr_dx_add[:, :, npa, ne] = cum_evo @ U_derr_dx_add[:, :, npa, ne]
    return cum_evo, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add             # :154


def calculate_interaction_error_operators(problem, x, expm=_expm):
    """reference src/UnitaryCalculations.jl:180-204.  Returns (ndim, ndim, ntimes, nerr)."""
    x_main, x_add, nparam = _split(problem, x)
    ntimes, ndim = problem.ntimes, problem.ndim
    cum_evo = np.eye(ndim, dtype=np.complex128)
    dt = problem.t0 / problem.ntimes
    nerr = len(problem.error_sources)
    out = np.zeros((ndim, ndim, nerr, ntimes), np.complex128)
    for nt in range(1, ntimes + 1):
        xk = x_main[:, nt - 1]
        cum_evo_inv = np.linalg.inv(cum_evo)
        for ne in range(nerr):
            Oerr = (1 / problem.eps) * problem.error_sources[ne].Herror(nt, xk, x_add, problem.eps)
            out[:, :, ne, nt - 1] = cum_evo_inv @ Oerr @ cum_evo
        cum_evo = expm(-1j * dt * problem.H0(nt, xk, x_add)) @ cum_evo
    return np.transpose(out, (0, 1, 3, 2))


def _projectors(fp):
    """reference src/FidelityCalculations.jl:47-51"""
    P0 = np.asarray(fp.projector).astype(np.complex128)
    P = P0.copy()
    P[P != 0] = 1
    D = np.real(np.trace(P0))
    return P0, P, D


def calculate_fidelity_and_derivatives(fp, x, expm=_expm):
    """reference src/FidelityCalculations.jl:19-119.  Returns (F, F_dx_tot, F_d2err, F_d2err_dx_tot)."""
    up = fp.unitary_problem
    ndim = up.ndim
    U, U_dx, U_dx_add, U_derr, U_derr_dx, U_derr_dx_add = calculate_unitary_and_derivatives(up, x, expm)
    ntimes, na, nerr = up.ntimes, up.nb_additional_param, len(up.error_sources)
    x_main, x_add, nparam = _split(up, x)

    U0 = np.asarray(fp.target_unitary(x_add), dtype=np.complex128)               # :32
    U0_dx_add = np.zeros((ndim, ndim, na), np.complex128)
    x_add_copy = x_add.copy()
    for npa in range(na):                                                        # :35-40
        x_add_copy[npa] += up.eps
        U0_temp = np.asarray(fp.target_unitary(x_add_copy), dtype=np.complex128)
        U0_dx_add[:, :, npa] = (1 / up.eps) * (U0_temp - U0)
        x_add_copy[npa] = x_add[npa]
    F_dx = np.zeros((nparam, ntimes))
    F_dx_add = np.zeros(na)
    F_d2err = np.zeros(nerr)
    F_d2err_dx = np.zeros((nparam, ntimes, nerr))
    F_d2err_dx_add = np.zeros((na, nerr))

    P0, P, D = _projectors(fp)
    tr_mod = lambda A: np.trace(P0 @ A)
    H = lambda A: A.conj().T
    U0h, Uh = H(U0), H(U)

    F = (np.real(tr_mod(P @ U0h @ U @ P @ Uh @ U0)) + abs(tr_mod(P @ U0h @ U)) ** 2) / (D * (D + 1))   # :54

    for nt in range(ntimes):                                                     # :56-65
        for np_ in range(nparam):
            X = U_dx[:, :, np_, nt]
            F_dx[np_, nt] = (
                np.real(tr_mod(P @ U0h @ X @ P @ Uh @ U0 + P @ U0h @ U @ P @ H(X) @ U0))
                + 2 * np.real(np.conj(tr_mod(P @ U0h @ U)) * tr_mod(P @ U0h @ X))
            ) / (D * (D + 1))
    for npa in range(na):                                                        # :67-76
        X = U_dx_add[:, :, npa]
        V = U0_dx_add[:, :, npa]
        F_dx_add[npa] = (
            np.real(tr_mod(P @ U0h @ X @ P @ Uh @ U0 + P @ U0h @ U @ P @ H(X) @ U0
                           + P @ H(V) @ U @ P @ Uh @ U0 + P @ U0h @ U @ P @ Uh @ V))
            + 2 * np.real(np.conj(tr_mod(P @ U0h @ U)) * tr_mod(P @ U0h @ X + P @ H(V) @ U))
        ) / (D * (D + 1))
    for ne in range(nerr):                                                       # :78-114
        E = U_derr[:, :, ne]
        Eh = H(E)
        F_d2err[ne] = 2 * (
            np.real(tr_mod(P @ U0h @ E @ P @ Eh @ U0 - P @ Eh @ E))
            + abs(tr_mod(P @ U0h @ E)) ** 2
            - D * np.real(tr_mod(P @ Eh @ E))
        ) / (D * (D + 1))
        for nt in range(ntimes):
            for np_ in range(nparam):
                Z = U_derr_dx[:, :, np_, nt, ne]
                Zh = H(Z)
                F_d2err_dx[np_, nt, ne] = 2 * (
                    np.real(tr_mod(P @ U0h @ Z @ P @ Eh @ U0 + P @ U0h @ E @ P @ Zh @ U0 - P @ Zh @ E - P @ Eh @ Z))
                    + 2 * np.real(np.conj(tr_mod(P @ U0h @ E)) * tr_mod(P @ U0h @ Z))
                    - D * np.real(tr_mod(P @ Zh @ E + P @ Eh @ Z))
                ) / (D * (D + 1))
        for npa in range(na):
            Z = U_derr_dx_add[:, :, npa, ne]
            Zh = H(Z)
            V = U0_dx_add[:, :, npa]
            F_d2err_dx_add[npa, ne] = 2 * (
                np.real(tr_mod(P @ H(V) @ E @ P @ Eh @ U0 + P @ U0h @ Z @ P @ Eh @ U0 + P @ U0h @ E @ P @ Zh @ U0
                               + P @ U0h @ E @ P @ Eh @ V - P @ Zh @ E - P @ Eh @ Z))
                + 2 * np.real(np.conj(tr_mod(P @ U0h @ E)) * tr_mod(P @ H(V) @ E + P @ U0h @ Z))
                - D * np.real(tr_mod(P @ Zh @ E + P @ Eh @ Z))
            ) / (D * (D + 1))

    F_dx_tot = np.concatenate([F_dx.reshape(nparam * ntimes, order="F"), F_dx_add])            # :116
    F_d2err_dx_tot = np.concatenate([F_d2err_dx.reshape((nparam * ntimes, nerr), order="F"), F_d2err_dx_add], axis=0)
    return float(F), F_dx_tot, F_d2err, F_d2err_dx_tot


def cost_and_gradient(fp, x, error_source_coeff=(), regularization_functions=None,
                      regularization_coeff1=None, regularization_coeff2=None, expm=_expm):
    """`calculate_common!` of reference src/FidelityCalculations.jl:174-197:
    buffer[0] = cost, buffer[1:] = gradient."""
    up = fp.unitary_problem
    na, ntimes = up.nb_additional_param, up.ntimes
    x = np.asarray(x, dtype=np.float64)
    nerr = len(up.error_sources)
    F, F_dx, F_d2err, F_d2err_dx = calculate_fidelity_and_derivatives(fp, x, expm)
    buffer = np.zeros(len(x) + 1)
    buffer[0] = 1 - F
    buffer[1:] = -F_dx
    if nerr > 0:
        coeff = np.asarray(error_source_coeff, dtype=np.float64)
        buffer[0] += np.sum(coeff * F_d2err ** 2)
        buffer[1:] += 2 * np.sum((coeff * F_d2err).reshape(1, nerr) * F_d2err_dx, axis=1)
    if regularization_functions:
        nparam = (len(x) - na) // ntimes
        x_main = x[: len(x) - na].reshape((nparam, ntimes), order="F")
        reg_grad = np.zeros((nparam, ntimes))
        for np_ in range(nparam):
            r1, j1, r2, j2 = regularization_functions[np_](x_main[np_, :].copy())
            buffer[0] += regularization_coeff1[np_] * r1 + regularization_coeff2[np_] * r2
            reg_grad[np_, :] = regularization_coeff1[np_] * np.asarray(j1) + regularization_coeff2[np_] * np.asarray(j2)
        buffer[1: len(x) + 1 - na] += reg_grad.reshape(nparam * ntimes, order="F")
    return buffer


def optimize_fidelity_and_error_sources(fp, params):
    """reference src/FidelityCalculations.jl:161-218 with scipy's L-BFGS-B standing in for
    Optim.LBFGS (the optimiser itself is out of the hot path; only used to reproduce the
    structure of the reference's optimisation tests)."""
    from scipy.optimize import minimize
    up = fp.unitary_problem
    assert len(params.error_source_coeff) == len(up.error_sources)
    nparam = (len(params.x_initial) - up.nb_additional_param) // up.ntimes
    assert len(params.regularization_coeff1) == nparam
    assert len(params.regularization_coeff2) == nparam
    assert len(params.regularization_functions) == nparam

    def fg(x):
        b = cost_and_gradient(fp, x, params.error_source_coeff, params.regularization_functions,
                              params.regularization_coeff1, params.regularization_coeff2)
        return b[0], b[1:]

    opts = {"maxiter": params.iterations}
    ap = dict(params.additional_parameters)
    if "g_tol" in ap:
        opts["gtol"] = ap["g_tol"]
    if "f_abstol" in ap:
        opts["ftol"] = ap["f_abstol"]
    return minimize(fg, np.asarray(params.x_initial, dtype=np.float64), jac=True, method="L-BFGS-B", options=opts)


def calculate_fidelity_response(fp, x, normalized_frequencies, expm=_expm):
    """reference src/FidelityCalculations.jl:246-280 (note the 0-based sum / 1-based weight quirk)."""
    up = fp.unitary_problem
    ntimes, nerr = up.ntimes, len(up.error_sources)
    freqs = np.asarray(normalized_frequencies, dtype=np.float64)
    dt = up.t0 / ntimes
    O = calculate_interaction_error_operators(up, x, expm)
    P0, P, D = _projectors(fp)
    tr_mod = lambda A: np.trace(P0 @ A)
    tidx = np.arange(ntimes)
    out = np.zeros((len(freqs), nerr))
    for ne in range(nerr):
        for nf, w in enumerate(freqs):
            S = np.tensordot(O[:, :, :, ne], np.exp(-1j * w * dt * tidx), axes=([2], [0]))   # :267
            r = 0.0
            for k in range(1, ntimes + 1):                                                   # :269-275
                Ok = O[:, :, k - 1, ne]
                ph = np.exp(1j * w * dt * k)
                r += (1.0 / D * np.real(ph * tr_mod(Ok @ S @ P))
                      - 1.0 / (D * (D + 1)) * np.real(ph * tr_mod(Ok @ P @ S @ P))
                      - 1.0 / (D * (D + 1)) * np.real(ph * tr_mod(Ok @ P) * tr_mod(S @ P)))
            out[nf, ne] = dt ** 2 * r
    return out


def calculate_fidelity_response_fft(fp, x, oversampling=1, expm=_expm):
    """reference src/FidelityCalculations.jl:306-343."""
    assert oversampling >= 1
    up = fp.unitary_problem
    ndim, ntimes, nerr = up.ndim, up.ntimes, len(up.error_sources)
    dt = up.t0 / ntimes
    O = calculate_interaction_error_operators(up, x, expm)
    n = ntimes * oversampling
    Opad = np.zeros((ndim, ndim, n, nerr), np.complex128)
    Opad[:, :, :ntimes, :] = O
    P0, P, D = _projectors(fp)
    tr_mod = lambda A: np.trace(P0 @ A)
    out = np.zeros((n, nerr))
    for ne in range(nerr):
        Of = np.fft.fft(Opad[:, :, :, ne], axis=2)
        Oi = n * np.fft.ifft(Opad[:, :, :, ne], axis=2)
        for nt in range(n):
            A, B = Oi[:, :, nt], Of[:, :, nt]
            out[nt, ne] = dt ** 2 * (1 / D * np.real(tr_mod(A @ B @ P))
                                     - 1 / (D * (D + 1)) * np.real(tr_mod(A @ P @ B @ P))
                                     - 1 / (D * (D + 1)) * np.real(tr_mod(A @ P) * tr_mod(B @ P)))
    freqs = (2 * np.pi / (n * dt)) * np.arange(n)
    return out, freqs


def calculate_expectation_values(fp, x, expm=_expm):
    """reference src/FidelityCalculations.jl:368-390."""
    up = fp.unitary_problem
    ntimes = up.ntimes
    O = calculate_interaction_error_operators(up, x, expm)
    nerr = O.shape[3]
    Oc = np.cumsum(O, axis=2)
    dt = up.t0 / ntimes
    P0, P, D = _projectors(fp)
    out = np.zeros((ntimes, nerr))
    for ne in range(nerr):
        for nt in range(ntimes):
            out[nt, ne] = np.real(dt * np.trace(P0 @ Oc[:, :, nt, ne]) / D)
    return out


# --- regularisation (reference src/Regularization.jl:26-47,78-83,111-115) ----------------
def regularization_cost(x, f=None, df=None):
    x = np.asarray(x, dtype=np.float64)
    if f is not None:
        r1, j1, r2, j2 = regularization_cost(f(x))
        d = df(x)
        return r1, d * j1, r2, d * j2
    n = len(x)
    dx = np.diff(x)
    ddx = np.diff(dx)
    reg1 = np.sum(dx ** 2)
    reg2 = np.sum(ddx ** 2)
    jac1 = np.zeros(n)
    jac2 = np.zeros(n)
    jac1[1:n - 1] = -2 * ddx
    jac1[0] += -2 * dx[0]
    jac1[n - 1] += 2 * dx[n - 2]
    jac2[0] = 2 * (x[2] - 2 * x[1] + x[0])
    jac2[1] = 2 * (x[3] - 4 * x[2] + 5 * x[1] - 2 * x[0])
    for i in range(2, n - 2):
        jac2[i] = 2 * (x[i + 2] - 4 * x[i + 1] + 6 * x[i] - 4 * x[i - 1] + x[i - 2])
    jac2[n - 2] = 2 * (x[n - 4] - 4 * x[n - 3] + 5 * x[n - 2] - 2 * x[n - 1])
    jac2[n - 1] = 2 * (x[n - 3] - 2 * x[n - 2] + x[n - 1])
    return reg1, jac1, reg2, jac2


def regularization_cost_phase(phis):
    c = regularization_cost(phis, np.cos, lambda v: -np.sin(v))
    s = regularization_cost(phis, np.sin, np.cos)
    return c[0] + s[0], c[1] + s[1], c[2] + s[2], c[3] + s[3]


def runtests_regularization_cost_phase(x):
    """The sin^2-of-differences helper the reference's *tests* define (test/runtests.jl:9-45)."""
    x = np.asarray(x, dtype=np.float64)
    dx = np.diff(x)
    ddx = np.diff(dx)
    reg1 = np.sum(np.sin(dx / 2) ** 2)
    reg2 = np.sum(np.sin(ddx / 2) ** 2)
    n = len(x)
    jac1 = np.zeros(n)
    jac2 = np.zeros(n)
    for i in range(1, n):            # Julia 1:n-1
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
