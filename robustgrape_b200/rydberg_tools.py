"""RydbergTools mirror (reference src/RydbergTools.jl).

Two layers:
  * the literal dense builders, same names/arguments/values as the reference
    (`rydberg_hamiltonian_symmetric_blockaded` :31-39, `..._full_blockaded` :71-81,
    `rydberg_hamiltonian_full` :118-130, `cz_with_1q_phase_symmetric` :160-162,
    `cz_with_1q_phase_full` :197-203, `unwrap_phase` :221-232);
  * descriptor factories returning callable term lists (robustgrape_b200.descriptors)
    that evaluate to those same matrices and can cross the C-ABI.
"""
from __future__ import annotations

import math

import numpy as np

from .descriptors import (Factor, Term, TermHamiltonian, TermErrorHamiltonian, TermTarget,
                          S_MAIN, S_ADD, OWNER_H0, OWNER_TARGET)

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------
# literal builders
# ----------------------------------------------------------------------------
def rydberg_hamiltonian_symmetric_blockaded(phi, eps, delta):
    """Basis |00>,|01>,|11>,|0r>,|W>  (reference src/RydbergTools.jl:31-39)."""
    em, ep = np.exp(-1j * phi), np.exp(1j * phi)
    H = np.zeros((5, 5), dtype=np.complex128)
    H[1, 3] = em * (1 + eps) / 2
    H[2, 4] = em * (1 + eps) / SQRT2
    H[3, 1] = ep * (1 + eps) / 2
    H[4, 2] = ep * (1 + eps) / SQRT2
    H[3, 3] = delta
    H[4, 4] = delta
    return H


def rydberg_hamiltonian_full_blockaded(phi, eps, delta):
    """Basis |00>,|01>,|10>,|11>,|0r>,|r0>,|W'>  (reference src/RydbergTools.jl:71-81)."""
    em, ep = np.exp(-1j * phi), np.exp(1j * phi)
    H = np.zeros((7, 7), dtype=np.complex128)
    H[1, 4] = em * (1 + eps) / 2
    H[2, 5] = em * (1 + eps) / 2
    H[3, 6] = em * (1 + eps) / SQRT2
    H[4, 1] = ep * (1 + eps) / 2
    H[5, 2] = ep * (1 + eps) / 2
    H[6, 3] = ep * (1 + eps) / SQRT2
    H[4, 4] = delta
    H[5, 5] = delta
    H[6, 6] = delta
    return H


def rydberg_hamiltonian_full(phi, Omega1, Omega2, delta1, delta2, B):
    """Basis |00>,|01>,|10>,|11>,|0r>,|r0>,|1r>,|r1>,|rr>  (reference src/RydbergTools.jl:118-130)."""
    em, ep = np.exp(-1j * phi), np.exp(1j * phi)
    H = np.zeros((9, 9), dtype=np.complex128)
    H[1, 4] = em * Omega1 / 2
    H[2, 5] = em * Omega2 / 2
    H[3, 6] = em * Omega1 / 2
    H[3, 7] = em * Omega2 / 2
    H[4, 1] = ep * Omega1 / 2
    H[4, 4] = delta1
    H[5, 2] = ep * Omega2 / 2
    H[5, 5] = delta2
    H[6, 3] = ep * Omega1 / 2
    H[6, 6] = delta1
    H[6, 8] = em * Omega2 / 2
    H[7, 3] = ep * Omega2 / 2
    H[7, 7] = delta2
    H[7, 8] = em * Omega1 / 2
    H[8, 6] = ep * Omega2 / 2
    H[8, 7] = ep * Omega1 / 2
    H[8, 8] = delta1 + delta2 + B
    return H


def cz_with_1q_phase_symmetric(theta):
    """diag(1, e^{i theta}, e^{i(2 theta + pi)}, 0, 0)  (reference src/RydbergTools.jl:160-162)."""
    return np.diag(np.array([1, np.exp(1j * theta), np.exp(1j * (2 * theta + np.pi)), 0, 0], dtype=np.complex128))


def cz_with_1q_phase_full(theta, rydberg_dimension=5):
    """reference src/RydbergTools.jl:197-203."""
    d = np.zeros(4 + rydberg_dimension, dtype=np.complex128)
    d[0] = 1
    d[1:3] = np.exp(1j * theta)
    d[3] = np.exp(1j * (2 * theta + np.pi))
    return np.diag(d)


def unwrap_phase(phi):
    """reference src/RydbergTools.jl:221-232 (plotting helper; host-only)."""
    p = np.mod(np.array(phi, dtype=float), 2 * np.pi)
    for i in range(len(p) - 1):
        if p[i + 1] - p[i] > np.pi:
            p[i + 1:] -= 2 * np.pi
        elif p[i + 1] - p[i] < -np.pi:
            p[i + 1:] += 2 * np.pi
    return p


# ----------------------------------------------------------------------------
# descriptor factories
# ----------------------------------------------------------------------------
_SYM_UP = ((1, 3, 0.5), (2, 4, 1 / SQRT2))            # entries multiplied by e^{-i phi}
_FB_UP = ((1, 4, 0.5), (2, 5, 0.5), (3, 6, 1 / SQRT2))
_SYM_RYD = (3, 4)
_FB_RYD = (4, 5, 6)


def _drive_terms(up, phase_index, extra=(), owner=OWNER_H0):
    dn = tuple((c, r, v) for r, c, v in up)
    return [
        Term(1.0, (Factor.expi(S_MAIN, phase_index, -1.0),) + tuple(extra), up, owner),
        Term(1.0, (Factor.expi(S_MAIN, phase_index, +1.0),) + tuple(extra), dn, owner),
    ]


def _model(model):
    if model == "symmetric_blockaded":
        return 5, _SYM_UP, _SYM_RYD
    if model == "full_blockaded":
        return 7, _FB_UP, _FB_RYD
    raise ValueError(model)


def rydberg_h0(model="symmetric_blockaded", phase_index=0, eps=0.0, delta=0.0):
    """H0(time_step, phi, x_add) = rydberg_hamiltonian_<model>(phi[phase_index], eps, delta)
    with constant eps/delta (the closures of examples/time_optimal_cz.jl:15 and
    test/runtests.jl:57,234-235)."""
    ndim, up, ryd = _model(model)
    amp = 1.0 + eps
    terms = _drive_terms(tuple((r, c, v * amp) for r, c, v in up), phase_index)
    if delta != 0.0:
        terms.append(Term(delta, (), tuple((r, r, 1.0) for r in ryd), OWNER_H0))
    return TermHamiltonian(ndim, terms)


def rydberg_amplitude_error(model="symmetric_blockaded", phase_index=0, source=0):
    """Herror(time_step, phi, x_add, e) = H_model(phi, e, 0) - H_model(phi, 0, 0)
    (examples/time_optimal_cz.jl:60): ((1+e) - 1) * (e^{-i phi} L + h.c.)."""
    ndim, up, _ = _model(model)
    return TermErrorHamiltonian(ndim, _drive_terms(up, phase_index, extra=(Factor.err1p_m1(),), owner=source))


def rydberg_frequency_error(model="symmetric_blockaded", source=0):
    """Herror(time_step, phi, x_add, d) = H_model(phi, 0, d) - H_model(phi, 0, 0)
    (examples/time_optimal_cz.jl:61): d * (projector on Rydberg levels)."""
    ndim, _, ryd = _model(model)
    return TermErrorHamiltonian(ndim, [Term(1.0, (Factor.err(),), tuple((r, r, 1.0) for r in ryd), source)])


def rydberg_decay_operator(model="symmetric_blockaded", source=0):
    """decay_operator(time_step, x, x_add, e) = e * diag(0,0,0,1,1) (examples/time_optimal_cz.jl:70)."""
    return rydberg_frequency_error(model, source)


def cz_target(model="symmetric_blockaded", theta_index=0):
    """cz_with_1q_phase_symmetric(x_add[theta_index]) / cz_with_1q_phase_full(...; rydberg_dimension=3)."""
    e1 = Factor.expi(S_ADD, theta_index, 1.0, 0.0)
    e2 = Factor.expi(S_ADD, theta_index, 2.0, math.pi)
    if model == "symmetric_blockaded":
        terms = [Term(1.0, (), ((0, 0, 1.0),), OWNER_TARGET),
                 Term(1.0, (e1,), ((1, 1, 1.0),), OWNER_TARGET),
                 Term(1.0, (e2,), ((2, 2, 1.0),), OWNER_TARGET)]
        return TermTarget(5, terms)
    if model == "full_blockaded":
        terms = [Term(1.0, (), ((0, 0, 1.0),), OWNER_TARGET),
                 Term(1.0, (e1,), ((1, 1, 1.0), (2, 2, 1.0)), OWNER_TARGET),
                 Term(1.0, (e2,), ((3, 3, 1.0),), OWNER_TARGET)]
        return TermTarget(7, terms)
    raise ValueError(model)


def rydberg_full_h0(Omega1=1.0, Omega2=1.0, delta1=0.0, delta2=0.0, B=0.0, phase_index=0):
    """H0 = rydberg_hamiltonian_full(phi[phase_index], Omega1, Omega2, delta1, delta2, B), constants fixed."""
    up = ((1, 4, Omega1 / 2), (2, 5, Omega2 / 2), (3, 6, Omega1 / 2), (3, 7, Omega2 / 2),
          (6, 8, Omega2 / 2), (7, 8, Omega1 / 2))
    terms = _drive_terms(up, phase_index)
    diag = ((4, 4, delta1), (5, 5, delta2), (6, 6, delta1), (7, 7, delta2), (8, 8, delta1 + delta2 + B))
    diag = tuple(e for e in diag if e[2] != 0.0)
    if diag:
        terms.append(Term(1.0, (), diag, OWNER_H0))
    return TermHamiltonian(9, terms)


def rydberg_h0_ensemble(model="symmetric_blockaded", phase_index=0, amp_index=1, delta_index=2):
    """H0(time_step, phi, x_add) = rydberg_hamiltonian_<model>(phi[phase_index], x_add[amp_index], x_add[delta_index]): the static
    amplitude deviation and detuning are *per-pulse* additional parameters, so a batch of pulses is an error ensemble -- every
    sample of (eps, delta) is one column of X, evaluated (and sharded over GPUs) like any other batch.  This is the batched form
    of the H0(+-eps2) variants of reference test/runtests.jl:228-289 (`north_star`: "error-ensemble samples")."""
    ndim, up, ryd = _model(model)
    amp = Factor.var(S_ADD, amp_index, 1.0, 1.0)          # (1 + eps)
    terms = _drive_terms(up, phase_index, extra=(amp,))
    terms.append(Term(1.0, (Factor.var(S_ADD, delta_index),), tuple((r, r, 1.0) for r in ryd), OWNER_H0))
    return TermHamiltonian(ndim, terms)
