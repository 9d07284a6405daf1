"""Problem definitions shared by the tests and the golden-fixture generator."""
from __future__ import annotations

import math

import numpy as np

import robustgrape_b200 as rg
from robustgrape_b200 import rydberg_tools as rt
from robustgrape_b200.descriptors import (Factor, Term, TermHamiltonian, TermErrorHamiltonian, TermTarget,
                                          S_MAIN, S_ADD, OWNER_H0, OWNER_TARGET)

PROJ5 = np.diag([1.0, 2.0, 1.0, 0.0, 0.0])
PROJ7 = np.diag([1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0])


def cz_problem(ntimes, t0, errors=(), model="symmetric_blockaded", eps=0.0, delta=0.0):
    """The CZ problems of reference test/runtests.jl and examples/*.jl, with descriptor closures."""
    srcs = []
    for i, e in enumerate(errors):
        if e == "amp":
            srcs.append(rg.ErrorSource(rt.rydberg_amplitude_error(model, source=i)))
        elif e in ("freq", "decay"):
            srcs.append(rg.ErrorSource(rt.rydberg_frequency_error(model, source=i)))
        else:
            raise ValueError(e)
    d = 5 if model == "symmetric_blockaded" else 7
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=d, H0=rt.rydberg_h0(model, eps=eps, delta=delta),
                                      nb_additional_param=1, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ5 if d == 5 else PROJ7, rt.cz_target(model))


def cz_problem_closures(ntimes, t0, errors=()):
    """Same problem with plain Python closures calling the literal builders (the reference's style)."""
    H0 = lambda k, phi, xa: rt.rydberg_hamiltonian_symmetric_blockaded(phi[0], 0, 0)
    srcs = []
    for e in errors:
        if e == "amp":
            srcs.append(rg.ErrorSource(lambda k, phi, xa, err: rt.rydberg_hamiltonian_symmetric_blockaded(phi[0], err, 0) - H0(k, phi, xa)))
        else:
            srcs.append(rg.ErrorSource(lambda k, phi, xa, err: rt.rydberg_hamiltonian_symmetric_blockaded(phi[0], 0, err) - H0(k, phi, xa)))
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=5, H0=H0, nb_additional_param=1, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ5, lambda xa: rt.cz_with_1q_phase_symmetric(xa[0]))


def detuned_problem(ntimes, t0, errors=("amp",)):
    """A problem whose Hamiltonian depends on an additional parameter: two controls per step
    (phase phi = x[0], amplitude scale via cos(x[1])), detuning = x_add[1], target phase = x_add[0],
    plus a per-step envelope table.  Exercises VAR/COS/TABLE factors, p=2, a=2 and dH/dx_add != 0."""
    up_ent = ((1, 3, 0.5), (2, 4, 1 / math.sqrt(2.0)))
    dn_ent = tuple((c, r, v) for r, c, v in up_ent)
    env = 0.8 + 0.2 * np.cos(np.linspace(0, 2.0, ntimes))[:, None]
    terms = [
        Term(1.0, (Factor.expi(S_MAIN, 0, -1.0), Factor.cos(S_MAIN, 1, 0.5, 0.1), Factor.table(0)), up_ent, OWNER_H0),
        Term(1.0, (Factor.expi(S_MAIN, 0, +1.0), Factor.cos(S_MAIN, 1, 0.5, 0.1), Factor.table(0)), dn_ent, OWNER_H0),
        Term(1.0, (Factor.var(S_ADD, 1, 0.3, 0.05),), ((3, 3, 1.0), (4, 4, 1.0)), OWNER_H0),
    ]
    H0 = TermHamiltonian(5, terms, table=env)
    srcs = []
    for i, e in enumerate(errors):
        if e == "amp":
            eterms = [
                Term(1.0, (Factor.expi(S_MAIN, 0, -1.0), Factor.cos(S_MAIN, 1, 0.5, 0.1), Factor.table(0), Factor.err1p_m1()), up_ent, i),
                Term(1.0, (Factor.expi(S_MAIN, 0, +1.0), Factor.cos(S_MAIN, 1, 0.5, 0.1), Factor.table(0), Factor.err1p_m1()), dn_ent, i),
            ]
            srcs.append(rg.ErrorSource(TermErrorHamiltonian(5, eterms, table=env)))
        else:
            srcs.append(rg.ErrorSource(TermErrorHamiltonian(5, [Term(1.0, (Factor.err(), Factor.sin(S_ADD, 1, 1.0, 0.7)),
                                                                      ((3, 3, 1.0), (4, 4, 1.0)), i)], table=env)))
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=5, H0=H0, nb_additional_param=2, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ5, rt.cz_target("symmetric_blockaded"))


def phase_only_problem(ntimes, t0, model="symmetric_blockaded", nerr=2, nparam=1):
    """A member of the phase-only drive class (DevProblem::pc) that is *not* the plain CZ problem: the laser phase is
    0.7 * x[nparam-1] + 0.3 plus a per-pulse offset 1.3 * x_add[1] (two EXPI factors, one of them in an additional parameter),
    the coupling constants are complex, and there are a multiplicative (1 + e) and a linear (e) amplitude-type error source with
    different complex strengths per block.  nparam = 2 adds a control the Hamiltonian does not depend on (zero gradient rows)."""
    d = 5 if model == "symmetric_blockaded" else 7
    up_ent = rt._SYM_UP if d == 5 else rt._FB_UP
    ph = (Factor.expi(S_MAIN, nparam - 1, -0.7, 0.3), Factor.expi(S_ADD, 1, 1.3, 0.0))
    phc = (Factor.expi(S_MAIN, nparam - 1, 0.7, -0.3), Factor.expi(S_ADD, 1, -1.3, 0.0))        # conjugate phase of the lower triangle
    c0 = [0.5 + 0.1j, 0.6 - 0.2j, 0.4 + 0.3j]
    c1 = [0.2 - 0.1j, 0.1 + 0.25j, -0.3 + 0.1j]

    def pair(coefs, extra, owner):
        up_t = tuple((r, c, v * coefs[i]) for i, (r, c, v) in enumerate(up_ent))
        dn_t = tuple((c, r, np.conj(v * coefs[i])) for i, (r, c, v) in enumerate(up_ent))
        return [Term(1.0, ph + extra, up_t, owner), Term(1.0, phc + extra, dn_t, owner)]

    terms = pair(c0, (), OWNER_H0)
    srcs = []
    if nerr >= 1:
        srcs.append(rg.ErrorSource(TermErrorHamiltonian(d, pair(c0, (Factor.err1p_m1(),), 0))))
    if nerr >= 2:
        srcs.append(rg.ErrorSource(TermErrorHamiltonian(d, pair(c1, (Factor.err(),), 1))))
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=d, H0=TermHamiltonian(d, terms), nb_additional_param=2, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ5 if d == 5 else PROJ7, rt.cz_target(model))


def random_pulse(fp, nparam=1, seed=0):
    up = fp.unitary_problem
    rng = np.random.default_rng(seed)
    return 2 * np.pi * rng.random(nparam * up.ntimes + up.nb_additional_param)


def golden_cases():
    cases = {}
    fp = cz_problem(24, 7.613 * 24 / 200)
    cases["cz5_e0_N24"] = (fp, random_pulse(fp, 1, 11))
    fp = cz_problem(17, 14.32 * 17 / 200, ("amp",))
    cases["cz5_e1_N17"] = (fp, random_pulse(fp, 1, 12))
    fp = cz_problem(12, 7.613 * 12 / 100, ("amp", "freq"))
    cases["cz5_e2_N12"] = (fp, random_pulse(fp, 1, 13))
    fp = cz_problem(10, 7.613 * 10 / 100, ("amp", "freq"), model="full_blockaded")
    cases["cz7_e2_N10"] = (fp, random_pulse(fp, 1, 14))
    fp = detuned_problem(9, 1.1, ("amp", "freq"))
    cases["detuned_p2_a2_e2_N9"] = (fp, random_pulse(fp, 2, 15))
    # BASELINE.json configs[1] at full size: examples/ar_cz.jl (N = 200, t0 = 14.32, amplitude error)
    fp = cz_problem(200, 14.32, ("amp",))
    cases["cz5_C2_e1_N200"] = (fp, random_pulse(fp, 1, 21))
    # dense random-Hermitian control problems on the DMMA path (BASELINE config 5 in miniature); the last needs squarings
    fp = dense_random_problem(16, 4, nparam=2, nerr=2, seed=16)
    cases["dense16_p2_e2_N4"] = (fp, np.random.default_rng(116).uniform(-1, 1, 8))
    fp = dense_random_problem(32, 3, nparam=2, nerr=1, seed=32, t0=1.2)
    cases["dense32_p2_e1_N3_squared"] = (fp, np.random.default_rng(132).uniform(-1, 1, 6))
    # detuned 5- and 7-level models (diagonal terms inside the 2 x 2 blocks), two error sources
    fp = cz_problem(31, 7.613 * 31 / 100, ("freq", "amp"), delta=0.37, eps=0.02)
    cases["cz5_detuned_e2_N31"] = (fp, random_pulse(fp, 1, 22))
    fp = cz_problem(19, 7.613 * 19 / 60, ("amp", "freq"), model="full_blockaded", delta=-0.21)
    cases["cz7_detuned_e2_N19"] = (fp, random_pulse(fp, 1, 23))
    return cases


def dense_random_problem(d, ntimes, nparam=2, nerr=1, seed=0, t0=None):
    """BASELINE config 5 in miniature: H(k) = H_0 + sum_j x_j(k) H_j with GUE draws scaled to unit spectral norm,
    Herr_e(err) = err * E_e, no additional parameters, projector = identity on the first min(d,4) levels,
    constant Haar-random target on that block."""
    from robustgrape_b200.descriptors import ConstantTarget
    rng = np.random.default_rng(seed)

    def gue():
        g = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        h = (g + g.conj().T) / 2
        return h / np.linalg.norm(h, 2)

    def ent(M):
        return tuple((r, c, M[r, c]) for r in range(d) for c in range(d))

    terms = [Term(1.0, (), ent(gue()), OWNER_H0)]
    for j in range(nparam):
        terms.append(Term(1.0, (Factor.var(S_MAIN, j),), ent(gue()), OWNER_H0))
    srcs = [rg.ErrorSource(TermErrorHamiltonian(d, [Term(1.0, (Factor.err(),), ent(gue()), e)])) for e in range(nerr)]
    nb = min(d, 4)
    proj = np.zeros((d, d)); proj[:nb, :nb] = np.eye(nb)
    q, _ = np.linalg.qr(rng.normal(size=(nb, nb)) + 1j * rng.normal(size=(nb, nb)))
    U0 = np.zeros((d, d), dtype=complex); U0[:nb, :nb] = q
    up = rg.UnitaryRobustGRAPEProblem(t0=t0 if t0 is not None else 0.05 * ntimes, ntimes=ntimes, ndim=d,
                                      H0=TermHamiltonian(d, terms), nb_additional_param=0, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, proj, ConstantTarget(U0))


def rydberg9_problem(ntimes, t0, errors=("amp",)):
    """9-level two-atom model (reference src/RydbergTools.jl:118-130) with finite blockade and detunings."""
    H0 = rt.rydberg_full_h0(1.0, 0.9, 0.05, -0.03, 8.0)
    up_ent = ((1, 4, 0.5), (2, 5, 0.45), (3, 6, 0.5), (3, 7, 0.45), (6, 8, 0.45), (7, 8, 0.5))
    dn_ent = tuple((c, r, v) for r, c, v in up_ent)
    srcs = []
    for i, e in enumerate(errors):
        if e == "amp":
            srcs.append(rg.ErrorSource(TermErrorHamiltonian(9, [
                Term(1.0, (Factor.expi(S_MAIN, 0, -1.0), Factor.err1p_m1()), up_ent, i),
                Term(1.0, (Factor.expi(S_MAIN, 0, +1.0), Factor.err1p_m1()), dn_ent, i)])))
        else:
            srcs.append(rg.ErrorSource(TermErrorHamiltonian(9, [Term(1.0, (Factor.err(),), tuple((r, r, 1.0) for r in range(4, 9)), i)])))
    proj = np.diag([1.0, 1, 1, 1, 0, 0, 0, 0, 0])
    tgt = TermTarget(9, [Term(1.0, (), ((0, 0, 1.0),), OWNER_TARGET),
                         Term(1.0, (Factor.expi(S_ADD, 0, 1.0, 0.0),), ((1, 1, 1.0), (2, 2, 1.0)), OWNER_TARGET),
                         Term(1.0, (Factor.expi(S_ADD, 0, 2.0, math.pi),), ((3, 3, 1.0),), OWNER_TARGET)])
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=9, H0=H0, nb_additional_param=1, error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, proj, tgt)


def decay_problem(ntimes, t0, gamma=0.05, errors=("amp",)):
    """Non-Hermitian effective Hamiltonian: CZ drive plus -i gamma/2 on the Rydberg levels.  Legal in the reference
    because it uses inv(), not the adjoint (src/UnitaryCalculations.jl:47,194)."""
    base = rt.rydberg_h0()
    terms = list(base.terms) + [Term(-0.5j * gamma, (), ((3, 3, 1.0), (4, 4, 1.0)), OWNER_H0)]
    srcs = []
    for i, e in enumerate(errors):
        srcs.append(rg.ErrorSource(rt.rydberg_amplitude_error(source=i) if e == "amp" else rt.rydberg_frequency_error(source=i)))
    up = rg.UnitaryRobustGRAPEProblem(t0=t0, ntimes=ntimes, ndim=5, H0=TermHamiltonian(5, terms), nb_additional_param=1,
                                      error_sources=srcs)
    return rg.FidelityRobustGRAPEProblem(up, PROJ5, rt.cz_target())
