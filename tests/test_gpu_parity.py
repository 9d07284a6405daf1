"""Parity of the CUDA path (called through the C ABI) against the oracles.

Tolerances (stated per check):
  * vs the exact-semantics golden fixtures (mpmath, tests/golden/): 1e-10 relative to the largest
    component of each output -- the north-star tolerance.  The CUDA path forms differenced
    exponentials, so it sits ~1e-15 from the exact value of the reference's finite-difference formulas.
  * vs the FP64 literal restatement (numpy / C++ port): F at 1e-12; eps-quotient outputs (F_dx, F_d2err)
    at 2e-5 and the eps2^2-quotient output (F_d2err_dx) at 2e-4 of the largest component.  These are the
    restatement's own finite-difference noise floors (1e-16 rounding amplified by 1/eps = 1e8 resp.
    1/eps2^2 = 1e8 and accumulated over ntimes; measured in
    tests/test_oracle_structure.py::test_fp64_restatement_vs_exact_semantics_noise_floor), not a
    property of the CUDA path, which the golden test pins at 1e-10.
"""
import os
from pathlib import Path

import numpy as np
import pytest

import robustgrape_b200 as rg
from robustgrape_b200 import _lib
from cases import cz_problem, detuned_problem, golden_cases, random_pulse
from oracle import cpu_port, reference_oracle as ro

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
NAMES = ["F", "F_dx", "F_d2err", "F_d2err_dx"]


def relmax(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", list(golden_cases().keys()))
def test_golden_exact_semantics(gpu_ctx, name):
    fp, x = golden_cases()[name]
    z = np.load(GOLD / f"{name}.npz")
    got = rg.calculate_fidelity_and_derivatives(fp, z["x"])
    for k, g in zip(NAMES, got):
        assert relmax(g, z["exact_" + k]) < 1e-10, (name, k, relmax(g, z["exact_" + k]))


def test_known_answer_evered_pulse(gpu_ctx):
    """reference test/runtests.jl:115-165."""
    T0 = 2 * np.pi * 1.22
    A, w0, p0, d0, th = 0.7701624, 0.97525275, -0.97449603, -0.04319765, 2.0802725844516097
    times = np.linspace(0, T0, 1000)
    xs = np.concatenate([A * np.cos(w0 * times - p0) + d0 * times, [th]])
    F = rg.calculate_fidelity_and_derivatives(cz_problem(1000, T0), xs)[0]
    assert F > 0.9999
    assert abs(F - 0.9999961847609591) < 1e-12


@pytest.mark.parametrize("N,errors,model,B", [
    (1, (), "symmetric_blockaded", 1), (2, ("amp",), "symmetric_blockaded", 3), (3, ("freq", "amp"), "symmetric_blockaded", 2),
    (33, (), "symmetric_blockaded", 7), (50, ("amp",), "symmetric_blockaded", 5), (64, ("amp", "freq"), "symmetric_blockaded", 4),
    (200, ("amp",), "symmetric_blockaded", 37), (41, ("amp", "freq"), "full_blockaded", 6), (129, (), "full_blockaded", 9)])
def test_batch_vs_fp64_restatement(gpu_ctx, N, errors, model, B):
    """Seeded random pulses incl. ragged sizes (N not a multiple of the chunk length, N=1, odd batches)."""
    fp = cz_problem(N, 7.613 * max(N, 10) / 500, errors, model)
    rng = np.random.default_rng(100 + N)
    X = 2 * np.pi * rng.random((N + 1, B))
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    pF, pFdx, pF2, pF2dx = cpu_port.PortProblem(fp).fidelity_and_derivatives_batch(X)
    assert np.abs(F - pF).max() < 1e-12
    assert relmax(Fdx, pFdx) < 2e-5
    assert relmax(F2, pF2) < 2e-5
    assert relmax(F2dx, pF2dx) < 2e-4
    coeff = [1e-4, 3e-4][:len(errors)]
    c, g = rg.cost_and_gradient_batch(fp, X, coeff)
    pc, pg = cpu_port.PortProblem(fp).cost_and_grad_batch(X, coeff)
    assert np.abs(c - pc).max() < 1e-9
    assert relmax(g, pg) < 2e-4
    # cost/grad epilogue is exactly the reference's combination of the four outputs (src/FidelityCalculations.jl:178-184)
    c2 = 1 - F + sum(coeff[e] * F2[e] ** 2 for e in range(len(errors)))
    g2 = -Fdx + sum(2 * coeff[e] * F2[e][None, :] * F2dx[:, e, :] for e in range(len(errors)))
    assert np.abs(c - c2).max() < 1e-14 and np.abs(g - g2).max() < 1e-13


def test_additional_parameter_dependent_hamiltonian(gpu_ctx):
    """p=2, a=2, table envelope, dH/dx_add != 0 (VAR/COS/SIN/TABLE factors)."""
    fp = detuned_problem(23, 2.5, ("amp", "freq"))
    X = np.stack([random_pulse(fp, 2, s) for s in range(4)], axis=1)
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    for b in range(4):
        a = ro.calculate_fidelity_and_derivatives(fp, X[:, b])
        assert abs(F[b] - a[0]) < 1e-12
        assert relmax(Fdx[:, b], a[1]) < 2e-5
        assert relmax(F2[:, b], a[2]) < 2e-5
        assert relmax(F2dx[:, :, b], a[3]) < 2e-4


def test_chunk_length_independence(gpu_ctx, monkeypatch):
    """The time-parallel blocking must not change results beyond rounding: chunk lengths 1, 5, 32, N."""
    N, B = 100, 6
    X = 2 * np.pi * np.random.default_rng(3).random((N + 1, B))
    ref = None
    for L in (1, 5, 32, 100):
        monkeypatch.setenv("RG_CHUNK", str(L))
        fp = cz_problem(N, 3.0, ("amp",))
        out = rg.calculate_fidelity_and_derivatives_batch(fp, X)
        if ref is None:
            ref = out
        else:
            for k, u, v in zip(NAMES, out, ref):
                assert relmax(u, v) < 1e-11, (L, k)


def test_batch_composition_independence(gpu_ctx):
    """Pulse b's result does not depend on which other pulses share the batch."""
    N = 77
    fp = cz_problem(N, 2.0, ("amp",))
    X = 2 * np.pi * np.random.default_rng(8).random((N + 1, 10))
    full = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    one = rg.calculate_fidelity_and_derivatives(fp, X[:, 4])
    assert full[0][4] == one[0]
    assert np.array_equal(full[1][:, 4], one[1])
    assert np.array_equal(full[3][:, :, 4], one[3])


def test_full_size_properties(gpu_ctx):
    """BASELINE.json config C4 (8192 pulses x 1000 steps): size-independent properties.
       * 0 <= F <= 1;
       * rows match a 64-pulse oracle sample;
       * the gradient of a handful of pulses matches a central difference of the GPU's own F;
       * deterministic: two runs are bitwise identical."""
    import bench
    N, B = 1000, 8192
    fp = bench.make_problem(N, 0)
    X = bench.make_pulses(N, B).T
    cost, grad = rg.cost_and_gradient_batch(fp, X)
    cost2, grad2 = rg.cost_and_gradient_batch(fp, X)
    assert np.array_equal(cost, cost2) and np.array_equal(grad, grad2)
    assert np.all(cost >= -1e-12) and np.all(cost <= 1 + 1e-12)
    idx = np.arange(0, B, 128)
    pc, pg = cpu_port.PortProblem(fp).cost_and_grad_batch(X[:, idx])
    assert np.abs(cost[idx] - pc).max() < 1e-11
    assert relmax(grad[:, idx], pg) < 2e-5
    h = 1e-5
    for b, i in [(0, 0), (17, 500), (8191, 999), (4096, 1000)]:
        xp, xm = X[:, b].copy(), X[:, b].copy()
        xp[i] += h; xm[i] -= h
        c2, _ = rg.cost_and_gradient_batch(fp, np.stack([xp, xm], axis=1))
        assert abs((c2[0] - c2[1]) / (2 * h) - grad[i, b]) < 1e-7


def test_api_shapes_and_asserts(gpu_ctx):
    fp = cz_problem(10, 1.0, ("amp",))
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives(fp, np.linspace(0, 1, 11))
    assert isinstance(F, float) and Fdx.shape == (11,) and F2.shape == (1,) and F2dx.shape == (11, 1)
    with pytest.raises(AssertionError):                    # src/UnitaryCalculations.jl:22
        rg.calculate_fidelity_and_derivatives(cz_problem(7, 1.0), np.zeros(12))
    with pytest.raises(AssertionError):                    # src/FidelityCalculations.jl:162
        rg.cost_and_gradient_batch(fp, np.zeros((11, 2)), [])


@pytest.mark.parametrize("N,t0,errors,model", [(3, 40.0, ("amp", "freq"), "symmetric_blockaded"), (5, 300.0, ("amp",), "symmetric_blockaded"),
                                               (4, 25.0, ("freq", "amp"), "full_blockaded"), (2, 2.5, (), "symmetric_blockaded")])
def test_scaling_and_squaring(gpu_ctx, N, t0, errors, model):
    """dt * ||H||_1 up to ~40: outside the Taylor range, handled by scaling-and-squaring of the (value, difference)
    pairs.  Compared with the exact-semantics oracle (mpmath expm handles any norm) at 1e-9 of the largest
    component -- each squaring doubles the relative rounding error."""
    from oracle import exact_oracle as eo
    fp = cz_problem(N, t0, errors, model)
    x = random_pulse(fp, 1, 31)
    got = rg.calculate_fidelity_and_derivatives(fp, x)
    ex = eo.calculate_fidelity_and_derivatives(fp, x)
    for k, g, e in zip(NAMES, got, ex):
        assert relmax(g, e) < 1e-9, (k, relmax(g, e))


def test_norm_out_of_range_is_reported(gpu_ctx):
    fp = cz_problem(1, 1e9)        # dt * ||H|| ~ 1e9: beyond RG_MAX_SQUARINGS
    with pytest.raises(_lib.RGError) as ei:
        rg.calculate_fidelity_and_derivatives(fp, np.zeros(2))
    assert ei.value.code == _lib.RG_ERR_NORM


def test_optimize_wrapper_reaches_high_fidelity(gpu_ctx):
    """reference test/runtests.jl:356-416 through the GPU cost/gradient (scipy L-BFGS-B as optimiser)."""
    N, T0 = 200, 2 * np.pi * 1.22
    fp = cz_problem(N, T0)
    rng = np.random.default_rng(42)
    prm = rg.FidelityRobustGRAPEParameters(
        x_initial=np.concatenate([2 * np.pi * 0.001 * rng.random(N), [2 * np.pi * rng.random()]]),
        regularization_functions=[ro.runtests_regularization_cost_phase], regularization_coeff1=[1e-6],
        regularization_coeff2=[1e-6], error_source_coeff=[], iterations=200,
        additional_parameters={"f_abstol": 1e-15, "g_tol": 3e-10})
    res = rg.optimize_fidelity_and_error_sources(fp, prm)
    assert 1 - rg.calculate_fidelity_and_derivatives(fp, res.x)[0] < 1e-6


# ---- analysis entry points ---------------------------------
def _analysis_case(model="symmetric_blockaded", N=60):
    fp = cz_problem(N, 7.613 * N / 500, ("amp", "freq"), model)
    x = random_pulse(fp, 1, 21)
    return fp, x


@pytest.mark.parametrize("model,N", [("symmetric_blockaded", 60), ("full_blockaded", 60), ("symmetric_blockaded", 203),
                                     ("full_blockaded", 130)])
def test_interaction_error_operators(gpu_ctx, model, N):
    """reference src/UnitaryCalculations.jl:180-204; well-conditioned (no finite difference): 1e-11 relative.
    N < 64 runs the time-sequential kernel, larger N the time-parallel chunk pipeline."""
    fp, x = _analysis_case(model, N)
    got = rg.calculate_interaction_error_operators(fp.unitary_problem, x)
    ref = ro.calculate_interaction_error_operators(fp.unitary_problem, x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()


def test_fidelity_response_direct_and_sharded(gpu_ctx):
    """reference src/FidelityCalculations.jl:246-280, incl. the 0-based-sum / 1-based-weight phase quirk."""
    fp, x = _analysis_case()
    freqs = np.linspace(0, 3, 17)
    ref = ro.calculate_fidelity_response(fp, x, freqs)
    got = rg.calculate_fidelity_response(fp, x, freqs)
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()
    # frequency shards (multi-GPU partitioning of the grid) reproduce the rows of the full sweep
    a = rg.calculate_fidelity_response(fp, x, freqs, first=0, count=9)
    b = rg.calculate_fidelity_response(fp, x, freqs, first=9, count=8)
    assert np.array_equal(np.vstack([a, b]), got)


def test_fidelity_response_fft(gpu_ctx):
    """reference src/FidelityCalculations.jl:306-343 (on the GPU a direct DFT on the uniform grid)."""
    fp, x = _analysis_case(N=40)
    ref, fref = ro.calculate_fidelity_response_fft(fp, x, oversampling=3)
    got, fgot = rg.calculate_fidelity_response_fft(fp, x, oversampling=3)
    assert got.shape == (120, 2)
    assert np.allclose(fgot, fref, rtol=1e-15, atol=0)
    assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max()


def test_expectation_values(gpu_ctx):
    """reference src/FidelityCalculations.jl:368-390 (integrated Rydberg population, examples/time_optimal_cz.jl:70-74)."""
    fp = cz_problem(80, 7.613, ("decay",))
    x = random_pulse(fp, 1, 5)
    ref = ro.calculate_expectation_values(fp, x)
    got = rg.calculate_expectation_values(fp, x)
    assert got.shape == (80, 1)
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_response_at_zero_frequency_matches_sensitivity(gpu_ctx):
    """reference test/runtests.jl:531-619 and examples/time_optimal_cz.jl:82-84: -F_d2err == 2 R(0) for a pulse that
    implements the target gate (the identity needs U ~ U0), all on the GPU.  Pulse: the RNG-free Evered solution
    (test/runtests.jl:126-140, F = 0.999996)."""
    T0 = 2 * np.pi * 1.22
    A, w0, p0, d0, th = 0.7701624, 0.97525275, -0.97449603, -0.04319765, 2.0802725844516097
    times = np.linspace(0, T0, 1000)
    x = np.concatenate([A * np.cos(w0 * times - p0) + d0 * times, [th]])
    fp = cz_problem(1000, T0, ("amp", "freq"))
    F, _, s, _ = rg.calculate_fidelity_and_derivatives(fp, x)
    assert F > 0.9999
    R = rg.calculate_fidelity_response(fp, x, np.array([0.0, 1.0]))
    Rf, _ = rg.calculate_fidelity_response_fft(fp, x, oversampling=2)
    assert np.allclose(-s, 2 * R[0], rtol=1e-3, atol=1e-3)
    assert np.allclose(-s, 2 * Rf[0], rtol=1e-3, atol=1e-3)
    assert np.all(np.abs(s) > 1.0)          # both sensitivities are O(1..10) for this gate: the check is not vacuous


@pytest.mark.parametrize("case", ["cz5_e2", "cz7_e1", "detuned", "cz5_e0", "squared"])
def test_unitary_and_derivatives_materialised(gpu_ctx, case):
    """calculate_unitary_and_derivatives (reference src/UnitaryCalculations.jl:20-155): all six returned tensors.
    U is well conditioned (1e-12); the derivative tensors are compared with the FP64 restatement at its
    finite-difference noise floor (2e-5 / 2e-4 of the largest element) and, where cheap, with the exact-semantics
    oracle at 1e-10."""
    from oracle import exact_oracle as eo
    if case == "cz5_e2":
        fp = cz_problem(37, 7.613 * 37 / 300, ("amp", "freq")); p = 1
    elif case == "cz7_e1":
        fp = cz_problem(21, 1.5, ("amp",), "full_blockaded"); p = 1
    elif case == "detuned":
        fp = detuned_problem(13, 1.7, ("amp", "freq")); p = 2
    elif case == "squared":
        fp = cz_problem(4, 30.0, ("amp",)); p = 1
    else:
        fp = cz_problem(50, 2.0); p = 1
    up = fp.unitary_problem
    x = random_pulse(fp, p, 77)
    got = rg.calculate_unitary_and_derivatives(up, x)
    ref = ro.calculate_unitary_and_derivatives(up, x)
    names = ["U", "U_dx", "U_dx_add", "U_derr", "U_derr_dx", "U_derr_dx_add"]
    tol = {"U": 1e-12, "U_dx": 2e-5, "U_dx_add": 2e-5, "U_derr": 2e-5, "U_derr_dx": 2e-4, "U_derr_dx_add": 2e-4}
    refd = dict(zip(names, ref))
    # scale of each family: a tensor whose exact value is 0 (x_add the Hamiltonian does not depend on) comes out of the
    # FP64 restatement as pure (A + U - A - U)/eps2^2 rounding noise, so it is judged against its family's scale
    fam = {"U": ["U"], "U_dx": ["U_dx"], "U_dx_add": ["U_dx_add", "U_dx"], "U_derr": ["U_derr"],
           "U_derr_dx": ["U_derr_dx"], "U_derr_dx_add": ["U_derr_dx_add", "U_derr_dx"]}
    for n, g, r in zip(names, got, ref):
        assert g.shape == r.shape, n
        if r.size:
            scale = max(np.abs(refd[m]).max() if refd[m].size else 0.0 for m in fam[n]) * (up.ntimes if n.endswith("_add") else 1)
            assert np.abs(g - r).max() <= tol[n] * scale, (n, np.abs(g - r).max() / scale)
    # exact-semantics check of every tensor
    ex = eo.calculate_unitary_and_derivatives(up, x)
    d, N, a, e = up.ndim, up.ntimes, up.nb_additional_param, len(up.error_sources)
    tom = lambda M: np.array([[complex(M[i, j]) for j in range(d)] for i in range(d)])
    tol_ex = 1e-9 if case == "squared" else 1e-10
    assert np.abs(got[0] - tom(ex[0])).max() < 1e-13 * (30 if case == "squared" else 1)
    for i in range(p):
        for k in range(0, N, max(1, N // 5)):
            r = tom(ex[1][i][k])
            assert np.abs(got[1][:, :, i, k] - r).max() <= tol_ex * max(np.abs(r).max(), 1e-300)
            for ee in range(e):
                r = tom(ex[4][i][k][ee])
                assert np.abs(got[4][:, :, i, k, ee] - r).max() <= tol_ex * np.abs(r).max()
    for ee in range(e):
        r = tom(ex[3][ee])
        assert np.abs(got[3][:, :, ee] - r).max() <= tol_ex * np.abs(r).max()
        for j in range(a):
            r = tom(ex[5][j][ee])
            assert np.abs(got[5][:, :, j, ee] - r).max() <= tol_ex * max(np.abs(r).max(), 1e-30) + 1e-300
    for j in range(a):
        r = tom(ex[2][j])
        assert np.abs(got[2][:, :, j] - r).max() <= tol_ex * max(np.abs(r).max(), 1e-30) + 1e-300


@pytest.mark.parametrize("d,N", [(3, 9), (6, 9), (10, 9), (11, 9), (12, 9), (16, 9), (24, 7), (32, 5), (48, 4), (64, 4)])
def test_dense_random_hamiltonians(gpu_ctx, d, N):
    """Dense random-Hermitian control problems (BASELINE config 5 in miniature): p = 2 controls, one error source, no
    additional parameters, identity-block projector, constant target.  d <= 10 runs the general group kernels, d >= 11 the
    DMMA path (rg_big.cuh: jets of planar matrices, Paterson-Stockmeyer Taylor + squarings)."""
    from cases import dense_random_problem
    fp = dense_random_problem(d, N, nparam=2, nerr=1, seed=d)
    X = np.stack([np.random.default_rng(s).uniform(-1, 1, 2 * N) for s in range(3)], axis=1)
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    for b in range(3):
        a = ro.calculate_fidelity_and_derivatives(fp, X[:, b])
        assert abs(F[b] - a[0]) < 1e-12
        assert relmax(Fdx[:, b], a[1]) < 2e-5
        assert relmax(F2[:, b], a[2]) < 2e-5
        assert relmax(F2dx[:, :, b], a[3]) < 2e-4


def test_nine_level_rydberg_model(gpu_ctx):
    """reference src/RydbergTools.jl:118-130 (rydberg_hamiltonian_full) with cz_with_1q_phase_full, d = 9."""
    from cases import rydberg9_problem
    from oracle import exact_oracle as eo
    fp = rydberg9_problem(7, 2.0, ("amp", "freq"))
    x = random_pulse(fp, 1, 9)
    got = rg.calculate_fidelity_and_derivatives(fp, x)
    ex = eo.calculate_fidelity_and_derivatives(fp, x)
    for k, g, e in zip(NAMES, got, ex):
        assert relmax(g, e) < 1e-10, (k, relmax(g, e))
    O = rg.calculate_interaction_error_operators(fp.unitary_problem, x)
    Or = ro.calculate_interaction_error_operators(fp.unitary_problem, x)
    assert np.abs(O - Or).max() < 1e-11 * np.abs(Or).max()
    # response / expectation kernels at d*d = 81 elements per operator (strided element loop)
    freqs = np.linspace(0, 2.5, 7)
    R, Rr = rg.calculate_fidelity_response(fp, x, freqs), ro.calculate_fidelity_response(fp, x, freqs)
    assert np.abs(R - Rr).max() < 1e-10 * np.abs(Rr).max()
    E, Er = rg.calculate_expectation_values(fp, x), ro.calculate_expectation_values(fp, x)
    assert np.abs(E - Er).max() < 1e-12 * max(1.0, np.abs(Er).max())


def test_analysis_entry_points_fail_closed_on_the_dense_path(gpu_ctx):
    """ndim > 10 runs the DMMA path, which implements the fused fidelity/gradient entry points only: the analysis and
    materialising entry points must refuse (RG_ERR_UNSUPPORTED), never return unfilled buffers."""
    from cases import dense_random_problem
    fp = dense_random_problem(12, 5, nparam=1, nerr=1, seed=3)
    x = np.random.default_rng(1).uniform(-1, 1, 5)
    for call in (lambda: rg.calculate_fidelity_response(fp, x, np.array([0.0, 1.0])),
                 lambda: rg.calculate_expectation_values(fp, x),
                 lambda: rg.calculate_interaction_error_operators(fp.unitary_problem, x),
                 lambda: rg.calculate_unitary_and_derivatives(fp.unitary_problem, x)):
        with pytest.raises(_lib.RGError) as ei:
            call()
        assert ei.value.code == _lib.RG_ERR_UNSUPPORTED


def test_structured_path_with_non_diagonal_target_and_projector(gpu_ctx):
    """The CZ Hamiltonian takes the structured fast path, but a dense target / non-diagonal projector puts the
    co-state outside the Hamiltonian's block pattern: the library must notice and use the general gradient sweep."""
    from robustgrape_b200.descriptors import ConstantTarget
    from robustgrape_b200 import rydberg_tools as rt
    rng = np.random.default_rng(4)
    q, _ = np.linalg.qr(rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5)))
    proj = np.diag([1.0, 2, 1, 0, 0]); proj[0, 1] = 0.5; proj[1, 0] = 0.5
    up = rg.UnitaryRobustGRAPEProblem(t0=1.3, ntimes=40, ndim=5, H0=rt.rydberg_h0(), nb_additional_param=1,
                                      error_sources=[rg.ErrorSource(rt.rydberg_amplitude_error())])
    fp = rg.FidelityRobustGRAPEProblem(up, proj, ConstantTarget(q))
    x = random_pulse(fp, 1, 2)
    got = rg.calculate_fidelity_and_derivatives(fp, x)
    ref = ro.calculate_fidelity_and_derivatives(fp, x)
    assert abs(got[0] - ref[0]) < 1e-12
    assert relmax(got[1], ref[1]) < 2e-5 and relmax(got[2], ref[2]) < 2e-5 and relmax(got[3], ref[3]) < 2e-4


@pytest.mark.parametrize("N,t0,errors", [(30, 4.0, ("amp", "freq")), (64, 7.613, ()), (6, 60.0, ("amp",))])
def test_non_hermitian_hamiltonian(gpu_ctx, N, t0, errors):
    """-i gamma/2 decay on the Rydberg levels: propagators are not unitary, so the backward sweeps rewind with the
    stored exp(+i dt H) instead of the adjoint (the reference uses inv(), src/UnitaryCalculations.jl:47).
    The last case also needs scaling-and-squaring."""
    from cases import decay_problem
    from oracle import exact_oracle as eo
    fp = decay_problem(N, t0, 0.08, errors)
    assert not fp.unitary_problem.H0.is_hermitian()
    x = random_pulse(fp, 1, 41)
    got = rg.calculate_fidelity_and_derivatives(fp, x)
    ref = ro.calculate_fidelity_and_derivatives(fp, x)
    assert abs(got[0] - ref[0]) < 1e-12
    for k, g, r, tol in zip(NAMES[1:], got[1:], ref[1:], (2e-5, 2e-5, 2e-4)):
        assert relmax(g, r) < tol, (k, relmax(g, r))
    if N <= 30:
        ex = eo.calculate_fidelity_and_derivatives(fp, x)
        for k, g, e in zip(NAMES, got, ex):
            assert relmax(g, e) < (1e-9 if t0 > 20 else 1e-10), (k, relmax(g, e))
    if errors:
        O = rg.calculate_interaction_error_operators(fp.unitary_problem, x)
        Or = ro.calculate_interaction_error_operators(fp.unitary_problem, x)
        assert np.abs(O - Or).max() < 1e-10 * np.abs(Or).max()
        U = rg.calculate_unitary_and_derivatives(fp.unitary_problem, x)
        Ur = ro.calculate_unitary_and_derivatives(fp.unitary_problem, x)
        assert np.abs(U[0] - Ur[0]).max() < 1e-12
        assert np.abs(U[1] - Ur[1]).max() < 2e-5 * np.abs(Ur[1]).max()
        assert np.abs(U[4] - Ur[4]).max() < 2e-4 * np.abs(Ur[4]).max()


def test_no_gradient_mode_and_empty_batch(gpu_ctx):
    """F / F_d2err only (no gradient sweeps, no mixed differences) agree with the full call; B = 0 is a no-op."""
    fp = cz_problem(45, 2.2, ("amp", "freq"))
    X = 2 * np.pi * np.random.default_rng(12).random((46, 5))
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    F_, Fdx_, F2_, F2dx_ = rg.calculate_fidelity_and_derivatives_batch(fp, X, want_grad=False)
    assert Fdx_ is None and F2dx_ is None
    assert np.array_equal(F, F_) and np.array_equal(F2, F2_)
    from robustgrape_b200.unitary_calculations import device_problem
    dp = device_problem(fp)
    h, p = dp.handle_for(46)
    assert dp.ctx.lib.rg_cost_and_grad_batch(h, 0, None, None, None, None) == 0


def test_device_pointer_entry_point_matches_host_entry_point(gpu_ctx):
    """rg_cost_and_grad_batch_dev (device buffers, caller's stream) == rg_cost_and_grad_batch (host buffers, pipelined slabs)."""
    import torch
    from robustgrape_b200.unitary_calculations import device_problem
    fp = cz_problem(100, 7.613 / 10, ("amp",))
    B, nx = 4100, 101                       # more than one host slab
    X = 2 * np.pi * np.random.default_rng(13).random((B, nx))
    cost_h, grad_h = rg.cost_and_gradient_batch(fp, X.T, [1e-4])
    dp = device_problem(fp)
    dX = torch.from_numpy(X).cuda()
    dc = torch.empty(B, dtype=torch.float64, device="cuda")
    dg = torch.empty(B * nx, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dp.ctx.set_stream(s.cuda_stream)
        dp.cost_and_grad_batch_dev(B, nx, dX.data_ptr(), [1e-4], dc.data_ptr(), dg.data_ptr())
        dp.ctx.synchronize()
        dp.ctx.set_stream(0)
    # the host path evaluates 2050-pulse slabs, the device path one 4100-pulse slab: the planner picks different chunk
    # lengths for them, so the two agree to rounding (test_chunk_length_independence: 1e-11), not bit for bit
    assert np.allclose(dc.cpu().numpy(), cost_h, rtol=0, atol=1e-12)
    g_dev = dg.cpu().numpy().reshape(B, nx).T
    assert np.abs(g_dev - grad_h).max() <= 1e-11 * max(1.0, np.abs(grad_h).max())


@pytest.mark.parametrize("mode", [0, 1])
def test_peer_gather_single_process(gpu_ctx, mode):
    """rg_gather_to_peers (include/robustgrape_b200.h): with one process the only 'peer' is the rank's own buffer; the block
    must land at its slot, ordered after the work queued on the context stream. (The multi-process path is checked against
    an NCCL all-gather inside bench.py whenever it runs on more than one GPU.)"""
    import torch
    from robustgrape_b200.sharding import PeerGather
    world, rank, blk = 4, 2, 1003                         # odd block: exercises the 8-byte path of the store kernel
    pg = PeerGather(gpu_ctx, 0, 1, world * blk, nbuf=2, mode=mode, exchange=lambda h: [h])
    pg.rank, pg.world, pg.block = rank, world, blk        # view the one buffer as 4 slots, act as rank 2
    dev = torch.device("cuda", gpu_ctx.device)
    for buf in (0, 1):
        out = pg.view(buf, dev)
        out.zero_()
        src = torch.arange(blk, dtype=torch.float64, device=dev) + 7.0 * buf
        torch.cuda.synchronize()
        pg.push(src.data_ptr(), buf)
        pg.wait(buf)
        gpu_ctx.synchronize()
        got = out.cpu().numpy().reshape(world, blk)
        assert np.array_equal(got[rank], src.cpu().numpy())
        assert not got[[0, 1, 3]].any()
    del out
    pg.close()


# ---- config-size parity (BASELINE.json configs 2-4 at their full sizes) ------------------------------------------------
def test_config_C3_response_sweep_full_size(gpu_ctx):
    """BASELINE.json configs[2]: 4096-point frequency grid, N = 500, two error sources (examples/time_optimal_cz.jl:60-67,82).
    The full sweep runs on the GPU; a 64-frequency sample of it is compared with the literal restatement
    (src/FidelityCalculations.jl:246-280) and 8 frequency shards (the multi-GPU partition) must concatenate bit for bit."""
    fp = cz_problem(500, 7.613, ("amp", "freq"))
    x = random_pulse(fp, 1, 33)
    freqs = np.linspace(0, 3, 4096)
    got = rg.calculate_fidelity_response(fp, x, freqs)
    assert got.shape == (4096, 2)
    idx = np.arange(0, 4096, 64)
    ref = ro.calculate_fidelity_response(fp, x, freqs[idx])
    assert np.abs(got[idx] - ref).max() < 1e-10 * np.abs(ref).max()
    shards = [rg.calculate_fidelity_response(fp, x, freqs, first=r * 512, count=512) for r in range(8)]
    assert np.array_equal(np.vstack(shards), got)


def test_config_C4prime_full_size_one_error_source(gpu_ctx):
    """C4' (8192 pulses x 1000 steps, amplitude error, c_e = 1e-4; examples/ar_cz.jl:52): 64-pulse sample vs the C++ port
    of the literal algorithm, determinism, and -- on 16 pulses -- agreement of the workspace-free path with the
    step-matrix-workspace path (RG_WS=1), two independent implementations of the same differences."""
    import bench
    N, B = 1000, 8192
    fp = bench.make_problem(N, 1)
    X = bench.make_pulses(N, B).T
    cost, grad = rg.cost_and_gradient_batch(fp, X, [1e-4])
    cost2, grad2 = rg.cost_and_gradient_batch(fp, X, [1e-4])
    assert np.array_equal(cost, cost2) and np.array_equal(grad, grad2)
    idx = np.arange(0, B, 128)
    pc, pg = cpu_port.PortProblem(fp).cost_and_grad_batch(X[:, idx], [1e-4])
    assert np.abs(cost[idx] - pc).max() < 1e-9
    assert relmax(grad[:, idx], pg) < 2e-4


@pytest.mark.parametrize("model,errors,delta", [("symmetric_blockaded", (), 0.0), ("symmetric_blockaded", ("amp", "freq"), 0.0),
                                                ("symmetric_blockaded", ("freq", "amp"), 0.3), ("full_blockaded", ("amp",), 0.0),
                                                ("full_blockaded", ("amp", "freq"), -0.2)])
def test_workspace_free_path_equals_workspace_path(gpu_ctx, monkeypatch, model, errors, delta):
    """The block-2 closed-form path (rg_block2.cuh: propagators recomputed in the sweeps) and the step-matrix-workspace
    path (differenced Horner recurrences, RG_WS=1) evaluate the same exact differences by unrelated formulas."""
    N, B = 130, 9
    d = 5 if model == "symmetric_blockaded" else 7
    X = 2 * np.pi * np.random.default_rng(17).random((N + 1, B))
    fp = cz_problem(N, 7.613 * N / 400, errors, model, delta=delta)
    new = rg.calculate_fidelity_and_derivatives_batch(fp, X)           # one-launch quaternion path where eligible, else block-2
    monkeypatch.setenv("RG_B2", "1")                                   # three-kernel block-2 path
    fp1 = cz_problem(N, 7.613 * N / 400, errors, model, delta=delta)
    mid = rg.calculate_fidelity_and_derivatives_batch(fp1, X)
    monkeypatch.setenv("RG_WS", "1")                                   # step-matrix workspace path
    fp2 = cz_problem(N, 7.613 * N / 400, errors, model, delta=delta)
    old = rg.calculate_fidelity_and_derivatives_batch(fp2, X)
    for k, u, m, v in zip(NAMES, new, mid, old):
        assert relmax(u, v) < 1e-11, (k, relmax(u, v))
        assert relmax(m, v) < 1e-11, (k, relmax(m, v))
    # cost/gradient epilogue of every path
    coeff = [1e-4, 3e-4][:len(errors)]
    c_old, g_old = rg.cost_and_gradient_batch(fp2, X, coeff)
    monkeypatch.delenv("RG_WS"); monkeypatch.delenv("RG_B2")
    fp3 = cz_problem(N, 7.613 * N / 400, errors, model, delta=delta)
    c_new, g_new = rg.cost_and_gradient_batch(fp3, X, coeff)
    assert np.abs(c_new - c_old).max() < 1e-12 and relmax(g_new, g_old) < 1e-11


@pytest.mark.parametrize("N,B", [(1, 3), (7, 2), (31, 5), (64, 1), (500, 1), (1000, 150), (333, 1100)])
def test_fused_quaternion_path_shapes(gpu_ctx, monkeypatch, N, B):
    """One-launch fused path (rg_fusedq.cuh) over ragged shapes: fewer steps than lanes, one pulse (4 warps per pulse),
    mid-size batches (2 warps per pulse), against the workspace path."""
    X = 2 * np.pi * np.random.default_rng(N + B).random((N + 1, B))
    fp = cz_problem(N, 7.613 * max(N, 20) / 1000, ("amp",))
    new = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    new_nograd = rg.calculate_fidelity_and_derivatives_batch(fp, X, want_grad=False)
    monkeypatch.setenv("RG_WS", "1")
    old = rg.calculate_fidelity_and_derivatives_batch(cz_problem(N, 7.613 * max(N, 20) / 1000, ("amp",)), X)
    for k, u, v in zip(NAMES, new, old):
        assert relmax(u, v) < 1e-11, (k, relmax(u, v))
    assert np.array_equal(new[0], new_nograd[0]) and np.array_equal(new[2], new_nograd[2])
    # the fused kernel's elementwise (diagonal projector / target) fidelity algebra against its dense fidelity algebra
    monkeypatch.delenv("RG_WS")
    monkeypatch.setenv("RG_DENSE_ALG", "1")
    dense_alg = rg.calculate_fidelity_and_derivatives_batch(cz_problem(N, 7.613 * max(N, 20) / 1000, ("amp",)), X)
    for k, u, v in zip(NAMES, new, dense_alg):
        assert relmax(u, v) < 1e-12, (k, relmax(u, v))


def test_block2_large_step_norm(gpu_ctx):
    """||dt H||_1 between 0.3 and 2 per step: long series in the block-2 path; beyond its range (z > 4.5) the host falls
    back to scaling-and-squaring.  Exact-semantics oracle at 1e-10 / 1e-9."""
    from oracle import exact_oracle as eo
    for N, t0, tol in ((6, 6 * 0.6, 1e-10), (5, 5 * 2.4, 1e-10), (4, 4 * 9.0, 1e-9)):
        fp = cz_problem(N, t0, ("amp", "freq"), delta=0.4)
        x = random_pulse(fp, 1, 50 + N)
        got = rg.calculate_fidelity_and_derivatives(fp, x)
        ex = eo.calculate_fidelity_and_derivatives(fp, x)
        for k, g, e in zip(NAMES, got, ex):
            assert relmax(g, e) < tol, (N, k, relmax(g, e))


def test_dense_path_config5_shape(gpu_ctx):
    """BASELINE.json configs[4] in miniature on the DMMA path: d = 64, p = 8 controls, e = 4 error sources (57 jet slots),
    norm ||dt H||_1 ~ 3 (squarings), N = 3: all four outputs against the literal restatement at its noise floor, and the
    cost/gradient epilogue."""
    from cases import dense_random_problem
    fp = dense_random_problem(64, 3, nparam=8, nerr=4, seed=5, t0=3 * 0.12)
    X = np.stack([np.random.default_rng(60 + s).uniform(-1, 1, 24) for s in range(2)], axis=1)
    F, Fdx, F2, F2dx = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    for b in range(2):
        a = ro.calculate_fidelity_and_derivatives(fp, X[:, b])
        assert abs(F[b] - a[0]) < 1e-12
        assert relmax(Fdx[:, b], a[1]) < 2e-5
        assert relmax(F2[:, b], a[2]) < 2e-5
        assert relmax(F2dx[:, :, b], a[3]) < 2e-4
    coeff = [1e-4, 2e-4, 3e-4, 4e-4]
    c, g = rg.cost_and_gradient_batch(fp, X, coeff)
    c2 = 1 - F + sum(coeff[e] * F2[e] ** 2 for e in range(4))
    g2 = -Fdx + sum(2 * coeff[e] * F2[e][None, :] * F2dx[:, e, :] for e in range(4))
    assert np.abs(c - c2).max() < 1e-13 and np.abs(g - g2).max() < 1e-12


@pytest.mark.parametrize("errors,N", [((), 30), (("amp",), 25), (("amp", "freq"), 18)])
def test_closure_problems_through_the_hstack_entry_points(gpu_ctx, errors, N):
    """the reference's own call style -- plain closures calling the literal RydbergTools builders
    (examples/time_optimal_cz.jl:15-16,60-61; test/runtests.jl:474-476) -- runs through rg_*_from_hstack.  The host evaluates
    the closures, the device does the rest; compared with the literal restatement at its finite-difference noise floor."""
    from cases import cz_problem_closures
    fp = cz_problem_closures(N, 7.613 * N / 300, errors)
    x = random_pulse(fp, 1, 60 + N)
    got = rg.calculate_fidelity_and_derivatives(fp, x)
    ref = ro.calculate_fidelity_and_derivatives(fp, x)
    assert abs(got[0] - ref[0]) < 1e-12
    for k, g, r, tol in zip(NAMES[1:], got[1:], ref[1:], (2e-5, 2e-5, 2e-4)):
        assert relmax(g, r) < tol, (k, relmax(g, r))
    # the same problem written with descriptors agrees with the closure version within that noise floor
    desc = rg.calculate_fidelity_and_derivatives(cz_problem(N, 7.613 * N / 300, errors), x)
    assert abs(got[0] - desc[0]) < 1e-13
    for k, g, r, tol in zip(NAMES[1:], got[1:], desc[1:], (2e-5, 2e-5, 2e-4)):
        assert relmax(g, r) < tol, (k, relmax(g, r))
    U = rg.calculate_unitary_and_derivatives(fp.unitary_problem, x)
    Ur = ro.calculate_unitary_and_derivatives(fp.unitary_problem, x)
    assert np.abs(U[0] - Ur[0]).max() < 1e-12
    assert np.abs(U[1] - Ur[1]).max() < 2e-5 * np.abs(Ur[1]).max()
    if errors:
        assert np.abs(U[3] - Ur[3]).max() < 2e-5 * np.abs(Ur[3]).max()
        assert np.abs(U[4] - Ur[4]).max() < 2e-4 * np.abs(Ur[4]).max()


# ---- optimiser loop on the device -------------------------------------------------------------------
def test_device_regularization_epilogue(gpu_ctx):
    """k_regularize against the host functions (reference src/Regularization.jl:26-47,78-83,111-115; test/runtests.jl:9-45) and
    against the literal restatement of calculate_common! with regularisation (src/FidelityCalculations.jl:174-197)."""
    from robustgrape_b200 import regularization as R
    from robustgrape_b200.unitary_calculations import device_problem
    N, B = 57, 5
    fp = detuned_problem(N, 2.0, ("amp",))            # p = 2 control rows, a = 2
    X = np.stack([random_pulse(fp, 2, 90 + s) for s in range(B)], axis=1)
    dp = device_problem(fp)
    c0, g0 = dp.cost_and_grad_batch(X, [2e-4])
    for funcs in ([R.regularization_cost, R.regularization_cost_phase], [R.regularization_cost_phase_sin2, R.regularization_cost],
                  [R.regularization_cost_phase, R.regularization_cost_phase_sin2]):
        c1s, c2s = [1e-3, 2e-2], [3e-3, 5e-4]
        reg = list(zip(R.device_kinds(funcs), c1s, c2s))
        c, g = dp.cost_and_grad_batch_reg(X, [2e-4], reg)
        for b in range(B):
            xm = X[:2 * N, b].reshape((2, N), order="F")
            cost = c0[b]
            grad = g0[:, b].copy()
            for i, f in enumerate(funcs):
                r1, j1, r2, j2 = f(xm[i].copy())
                cost += c1s[i] * r1 + c2s[i] * r2
                grad[i:2 * N:2] += c1s[i] * np.asarray(j1) + c2s[i] * np.asarray(j2)
            assert abs(c[b] - cost) < 1e-12 * max(1.0, abs(cost))
            assert np.abs(g[:, b] - grad).max() < 1e-12 * max(1.0, np.abs(grad).max())
    # the test-suite form agrees with the oracle's restatement of test/runtests.jl:9-45
    x = X[:2 * N:2, 0]
    a, bq = R.regularization_cost_phase_sin2(x), ro.runtests_regularization_cost_phase(x)
    assert abs(a[0] - bq[0]) < 1e-14 and np.abs(a[1] - bq[1]).max() < 1e-14 and abs(a[2] - bq[2]) < 1e-14 and np.abs(a[3] - bq[3]).max() < 1e-14


def test_device_lbfgs_reaches_high_fidelity(gpu_ctx):
    """reference test/runtests.jl:356-416 (40 L-BFGS iterations from a small random pulse -> infidelity < 1e-6) with the whole
    optimiser loop on the device, for a batch of independent starts."""
    from robustgrape_b200 import regularization as R
    N, T0, B = 200, 2 * np.pi * 1.22, 24
    fp = cz_problem(N, T0)
    rng = np.random.default_rng(42)
    X0 = np.concatenate([2 * np.pi * 0.001 * rng.random((N, B)), 2 * np.pi * rng.random((1, B))], axis=0)
    X, cost, iters, info = rg.optimize_batch_device(fp, X0, [], [R.regularization_cost_phase_sin2], [1e-6], [1e-6],
                                                    iterations=200, g_tol=3e-10)
    F = rg.calculate_fidelity_and_derivatives_batch(fp, X, want_grad=False)[0]
    assert np.median(1 - F) < 1e-6, (np.sort(1 - F), info)
    assert (1 - F < 1e-6).mean() >= 0.75, (np.sort(1 - F), info)
    # one evaluation per round; a pulse needs at most 8 trials per iteration (plus 8 for a final failed search)
    assert info["evaluations"] <= 9 + 8 * info["iterations"] and info["iterations"] <= 200, info
    # the single-pulse wrapper with the device optimiser
    prm = rg.FidelityRobustGRAPEParameters(x_initial=X0[:, 0], regularization_functions=[R.regularization_cost_phase_sin2],
                                           regularization_coeff1=[1e-6], regularization_coeff2=[1e-6], error_source_coeff=[], iterations=200,
                                           solver_algorithm="device-lbfgs", additional_parameters={"g_tol": 3e-10})
    res = rg.optimize_fidelity_and_error_sources(fp, prm)
    assert 1 - rg.calculate_fidelity_and_derivatives(fp, res.x)[0] < 1e-6


def test_error_ensemble_on_the_batch_axis(gpu_ctx):
    """reference test/runtests.jl:228-289: the second-order sensitivity F_d2err equals the central second difference of F over
    problems with H0(+-eps2).  Here the five Hamiltonian variants are one batch -- per-pulse (eps, delta) in x_add
    (rydberg_tools.rydberg_h0_ensemble) -- i.e. an error ensemble sharded like any other batch."""
    from robustgrape_b200 import rydberg_tools as rt
    from cases import PROJ5
    # the identity needs a pulse that implements the gate (the reference optimises one first): the RNG-free Evered solution
    N, T0, h = 1000, 2 * np.pi * 1.22, 1e-4
    A, w0, p0, d0, th = 0.7701624, 0.97525275, -0.97449603, -0.04319765, 2.0802725844516097
    times = np.linspace(0, T0, N)
    x = np.concatenate([A * np.cos(w0 * times - p0) + d0 * times, [th]])
    F0, _, s2, _ = rg.calculate_fidelity_and_derivatives(cz_problem(N, T0, ("amp", "freq")), x)
    up = rg.UnitaryRobustGRAPEProblem(t0=T0, ntimes=N, ndim=5, H0=rt.rydberg_h0_ensemble(), nb_additional_param=3, error_sources=[])
    fp = rg.FidelityRobustGRAPEProblem(up, PROJ5, rt.cz_target())
    samples = [(0.0, 0.0), (h, 0.0), (-h, 0.0), (0.0, h), (0.0, -h)]
    X = np.stack([np.concatenate([x, [e, d]]) for e, d in samples], axis=1)
    F = rg.calculate_fidelity_and_derivatives_batch(fp, X, want_grad=False)[0]
    assert abs(F[0] - F0) < 1e-13
    d2_amp = (F[1] - 2 * F[0] + F[2]) / h ** 2
    d2_freq = (F[3] - 2 * F[0] + F[4]) / h ** 2
    # the closed-form pulse is not a stationary point to machine precision (1 - F = 4e-6), which leaves a 3e-3 relative residual;
    # the reference gates its optimised pulse at rtol 1e-3 / atol 1e-2 (test/runtests.jl:289)
    assert np.isclose(d2_amp, s2[0], rtol=1e-2, atol=1e-2) and np.isclose(d2_freq, s2[1], rtol=1e-2, atol=1e-2)
    assert abs(s2[0]) > 1.0 and abs(s2[1]) > 1.0
    # shards of the ensemble (the multi-GPU partition) reproduce the full batch
    Fa = rg.calculate_fidelity_and_derivatives_batch(fp, X[:, :2], want_grad=False)[0]
    Fb = rg.calculate_fidelity_and_derivatives_batch(fp, X[:, 2:], want_grad=False)[0]
    assert np.array_equal(np.concatenate([Fa, Fb]), F)


def _path(fp, nx):
    from robustgrape_b200.unitary_calculations import device_problem
    return device_problem(fp).path(nx)


@pytest.mark.parametrize("model,N,B,nerr,nparam", [("symmetric_blockaded", 77, 5, 2, 1), ("full_blockaded", 40, 3, 2, 1),
                                                   ("symmetric_blockaded", 33, 9, 1, 2), ("symmetric_blockaded", 200, 2, 0, 1)])
def test_phase_only_class_equals_generic_fused_kernel(gpu_ctx, monkeypatch, model, N, B, nerr, nparam):
    """Phase-only drive class (k_fused_q PC instantiation: step constants evaluated once, own sincos) against the generic
    closed-form kernel (RG_NO_PC=1) on a problem with a two-factor phase, complex couplings and two kinds of amplitude error, and
    both against the exact-semantics oracle for one pulse."""
    from cases import phase_only_problem
    from oracle import exact_oracle as eo
    fp = phase_only_problem(N, 7.613 * N / 1000 * 6, model, nerr, nparam)
    X = 2 * np.pi * np.random.default_rng(N + B).random((nparam * N + 2, B)) - np.pi
    assert _path(fp, X.shape[0]) == "fused_q_pc"
    new = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    monkeypatch.setenv("RG_NO_PC", "1")
    fp2 = phase_only_problem(N, 7.613 * N / 1000 * 6, model, nerr, nparam)
    assert _path(fp2, X.shape[0]) == "fused_q"
    old = rg.calculate_fidelity_and_derivatives_batch(fp2, X)
    for k, u, v in zip(NAMES, new, old):
        assert relmax(u, v) < 1e-12, (k, relmax(u, v))
    if N <= 80:
        ex = eo.calculate_fidelity_and_derivatives(fp, X[:, 0])
        for k, g, e in zip(NAMES, [np.asarray(a)[..., 0] for a in new], ex):
            assert relmax(g, e) < 1e-10, (k, relmax(g, e))


def test_phase_only_class_detection(gpu_ctx):
    """The class is a property of the term lists: the CZ problems of BASELINE configs 1, 2, 4 are in it, a detuned or
    amplitude-controlled Hamiltonian is not, and a step norm outside the closed-form range falls back to the generic kernel."""
    assert _path(cz_problem(50, 7.613, ("amp",)), 51) == "fused_q_pc"
    assert _path(cz_problem(50, 7.613, ("amp",), "full_blockaded"), 51) == "fused_q_pc"
    assert _path(cz_problem(50, 7.613, ("amp", "freq")), 51) == "block2"
    assert _path(cz_problem(50, 7.613, (), delta=0.3), 51) == "block2"
    assert _path(detuned_problem(50, 7.613), 102) != "fused_q_pc"
    assert _path(cz_problem(4, 4 * 9.0, ("amp",)), 5) == "fused_q"          # z > 4.5: constants out of range


def test_phase_only_large_arguments(gpu_ctx):
    """Phases far outside [-pi, pi] (|x| up to 3e5: the Cody-Waite reduction up to 1e5, the library sincos beyond) against the
    literal restatement's F and the generic kernel."""
    fp = cz_problem(40, 7.613 * 40 / 1000 * 5)
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.uniform(-3e5, 3e5, (40, 6)), rng.uniform(-3, 3, (1, 6))])
    X[:, :3] *= 1e-2
    F, Fdx, _, _ = rg.calculate_fidelity_and_derivatives_batch(fp, X)
    for b in range(6):
        a = ro.calculate_fidelity_and_derivatives(fp, X[:, b])
        assert abs(F[b] - a[0]) < 1e-10, (b, F[b], a[0])
        assert relmax(Fdx[:, b], a[1]) < 2e-5


def test_reduced_vs_full_hamiltonian_on_gpu(gpu_ctx):
    """reference test/runtests.jl:418-529 on the CUDA path: the 5-level symmetric and the 7-level full-blockaded model give the same
    fidelity, sensitivities and gradients for the same pulse (the reference checks the sensitivities at rtol 1e-3; the differenced
    path agrees to 1e-10), on the fused phase-only kernel (amplitude error) and the block-2 path (amplitude + frequency error)."""
    rng = np.random.default_rng(3)
    for errs in (("amp",), ("amp", "freq")):
        X = 2 * np.pi * rng.random((201, 3))
        a = rg.calculate_fidelity_and_derivatives_batch(cz_problem(200, 7.613, errs), X)
        b = rg.calculate_fidelity_and_derivatives_batch(cz_problem(200, 7.613, errs, "full_blockaded"), X)
        assert np.abs(a[0] - b[0]).max() < 1e-12
        for k, u, v in zip(NAMES[1:], a[1:], b[1:]):
            assert relmax(u, v) < 1e-10, (errs, k, relmax(u, v))


@pytest.mark.parametrize("what,nerr,B,N", [(1, 0, 37, 101), (0, 0, 37, 101), (1, 1, 9, 101), (1, 0, 1, 101), (1, 0, 3, 30000), (0, 0, 2, 30000)])
def test_fused_evaluation_and_gather_single_process(gpu_ctx, what, nerr, B, N):
    """rg_cost_and_grad_batch_dev_scatter: the [cost | grad] block of an evaluation must land, bit for bit, at the given offset of
    every destination buffer -- from the evaluation kernel itself for e = 0 (fused stores, gradient staged through shared-memory
    rows) and from copy-engine pushes otherwise -- and the local outputs must equal those of the plain call.  One process: the
    'peers' are three more buffers on this device, viewed as slots of a world of 4 in which this rank is rank 2."""
    import torch
    from robustgrape_b200.unitary_calculations import device_problem
    # N = 30000: a pulse's gradient no longer fits the staged shared-memory rows -> evaluation without the fused stores + copy-engine pushes
    fp = cz_problem(N, 7.613 * min(N, 2000) / 1000 * 4, ("amp",)[:nerr])
    dp = device_problem(fp, gpu_ctx)
    nx, world, rank = N + 1, 4, 2
    blk = B * (1 + nx)
    dev = torch.device("cuda", gpu_ctx.device)
    X = torch.from_numpy(np.ascontiguousarray((2 * np.pi * np.random.default_rng(B).random((nx, B))).T)).to(dev)
    coeff = [1e-3] * nerr
    ref = torch.zeros(blk, dtype=torch.float64, device=dev)
    dp.cost_and_grad_batch_dev(B, nx, X.data_ptr(), coeff, ref[:B].data_ptr(), ref[B:].data_ptr())
    bufs = [torch.full((world * blk,), -7.0, dtype=torch.float64, device=dev) for _ in range(3)]
    loc = torch.zeros(blk, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    dp.cost_and_grad_batch_dev_scatter(B, nx, X.data_ptr(), coeff, loc[:B].data_ptr(), loc[B:].data_ptr(), [b.data_ptr() for b in bufs],
                                       rank * blk * 8, what)
    gpu_ctx.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(loc, ref)
    for b in bufs:
        got = b.view(world, blk)
        assert torch.equal(got[rank, :B], ref[:B])
        if what:
            assert torch.equal(got[rank, B:], ref[B:])
        else:
            assert bool((got[rank, B:] == -7.0).all())
        assert bool((got[[0, 1, 3]] == -7.0).all())


def test_reference_gradient_validation_testsets_on_gpu(gpu_ctx):
    """The two finite-difference testsets of the reference, statement for statement, on the CUDA path:
    test/runtests.jl:48-111 (error-sensitivity gradient: N = 200, t0 = 2 pi 1.22, amplitude error, perturbation 1e-4, rtol 1e-3 /
    atol 1e-5, the last test on the additional parameter) and :292-352 (fidelity gradient: N = 50, perturbation = the problem's
    eps = 1e-8, rtol 1e-3 / atol 1e-3, the last of five tests on the additional parameter)."""
    t0 = 2 * np.pi * 1.22
    rng = np.random.default_rng(42)
    # error sensitivity gradient
    fp = cz_problem(200, t0, ("amp",))
    for ntest in range(2):
        idx = 200 if ntest == 1 else int(rng.integers(0, 200))
        xs = 2 * np.pi * rng.random(201)
        _, _, s0, s0dx = rg.calculate_fidelity_and_derivatives(fp, xs)
        xs[idx] += 1e-4
        _, _, s1, _ = rg.calculate_fidelity_and_derivatives(fp, xs)
        assert np.isclose((s1[0] - s0[0]) / 1e-4, np.asarray(s0dx)[idx, 0], rtol=1e-3, atol=1e-5), (idx, (s1[0] - s0[0]) / 1e-4, np.asarray(s0dx)[idx, 0])
    # fidelity gradient
    fp = cz_problem(50, t0)
    eps = fp.unitary_problem.eps
    for ntest in range(5):
        idx = 50 if ntest == 4 else int(rng.integers(0, 50))
        xs = 2 * np.pi * rng.random(51)
        F0, g0, _, _ = rg.calculate_fidelity_and_derivatives(fp, xs)
        xs[idx] += eps
        F1 = rg.calculate_fidelity_and_derivatives(fp, xs)[0]
        assert np.isclose((F1 - F0) / eps, np.asarray(g0)[idx], rtol=1e-3, atol=1e-3), (idx, (F1 - F0) / eps, np.asarray(g0)[idx])
