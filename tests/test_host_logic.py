"""CPU-only checks of the host side: descriptors, ABI structs, exported symbols, loud failure without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import robustgrape_b200 as rg
from robustgrape_b200 import _lib, descriptors as D, rydberg_tools as rt
from cases import cz_problem, cz_problem_closures, detuned_problem
from oracle import reference_oracle as ro

ROOT = Path(__file__).resolve().parents[1]


def test_descriptors_equal_literal_builders():
    rng = np.random.default_rng(0)
    for _ in range(5):
        phi, e, dl = rng.normal(size=3)
        assert np.allclose(rt.rydberg_h0(eps=e, delta=dl)(1, [phi], [0.0]), rt.rydberg_hamiltonian_symmetric_blockaded(phi, e, dl), rtol=0, atol=1e-15)
        assert np.allclose(rt.rydberg_h0("full_blockaded", eps=e, delta=dl)(1, [phi], [0.0]), rt.rydberg_hamiltonian_full_blockaded(phi, e, dl), rtol=0, atol=1e-15)
        assert np.allclose(rt.rydberg_amplitude_error()(1, [phi], [0.0], 1e-4),
                           rt.rydberg_hamiltonian_symmetric_blockaded(phi, 1e-4, 0) - rt.rydberg_hamiltonian_symmetric_blockaded(phi, 0, 0), rtol=0, atol=3e-16)
        assert np.allclose(rt.rydberg_frequency_error()(1, [phi], [0.0], 1e-4),
                           rt.rydberg_hamiltonian_symmetric_blockaded(phi, 0, 1e-4) - rt.rydberg_hamiltonian_symmetric_blockaded(phi, 0, 0), rtol=0, atol=0)
        th = rng.normal()
        assert np.allclose(rt.cz_target()([th]), rt.cz_with_1q_phase_symmetric(th), rtol=0, atol=1e-15)
        assert np.allclose(rt.cz_target("full_blockaded")([th]), rt.cz_with_1q_phase_full(th, rydberg_dimension=3), rtol=0, atol=1e-15)
        O1, O2, d1, d2, B = rng.normal(size=5)
        assert np.allclose(rt.rydberg_full_h0(O1, O2, d1, d2, B)(1, [phi], []), rt.rydberg_hamiltonian_full(phi, O1, O2, d1, d2, B), rtol=0, atol=1e-15)


def test_literal_builders_are_hermitian_and_match_doc_shapes():
    for H in (rt.rydberg_hamiltonian_symmetric_blockaded(0.3, 0.1, 0.2), rt.rydberg_hamiltonian_full_blockaded(0.3, 0.1, 0.2),
              rt.rydberg_hamiltonian_full(0.3, 1.0, 0.9, 0.1, 0.2, 5.0)):
        assert np.allclose(H, H.conj().T)
    assert rt.rydberg_hamiltonian_full(0.0, 1, 1, 0, 0, 0).shape == (9, 9)
    assert rt.cz_with_1q_phase_full(0.2).shape == (9, 9)


def test_unwrap_phase():
    p = rt.unwrap_phase(np.array([0.1, 6.2, 0.2, 0.4]))       # reference src/RydbergTools.jl:221-232
    assert np.all(np.abs(np.diff(p)) < np.pi)


def test_hermitian_detection():
    assert rt.rydberg_h0().is_hermitian()
    assert rt.rydberg_amplitude_error().is_hermitian()
    nh = D.TermHamiltonian(2, [D.Term(-0.5j, (), ((1, 1, 1.0),))])     # -i gamma/2 decay: legal in the reference (it uses inv)
    assert not nh.is_hermitian()


def test_regularization_mirror_matches_oracle():
    x = np.random.default_rng(1).normal(size=17)
    for a, b in zip(rg.regularization_cost(x), ro.regularization_cost(x)):
        assert np.allclose(a, b, rtol=0, atol=1e-14)
    for a, b in zip(rg.regularization_cost_phase(x), ro.regularization_cost_phase(x)):
        assert np.allclose(a, b, rtol=0, atol=1e-14)
    # gradient of reg1/reg2 by central differences
    r1, j1, r2, j2 = rg.regularization_cost_phase(x)
    for i in (0, 1, 8, 15, 16):
        xp, xm = x.copy(), x.copy()
        xp[i] += 1e-6; xm[i] -= 1e-6
        a, b = rg.regularization_cost_phase(xp), rg.regularization_cost_phase(xm)
        assert abs((a[0] - b[0]) / 2e-6 - j1[i]) < 1e-6
        assert abs((a[2] - b[2]) / 2e-6 - j2[i]) < 1e-6


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.rg_factor) == 32
    assert C.sizeof(_lib.rg_term) == 8 + 16 + 4 * 32 + 8 + 3 * 8
    assert _lib.rg_problem_desc.t0.offset == 24


def test_library_exports_every_declared_symbol(built_library):
    header = (ROOT / "include" / "robustgrape_b200.h").read_text()
    declared = set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = C.CDLL(str(built_library))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/robustgrape_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_closures_take_the_hstack_path_and_never_the_descriptor_path():
    """Closures cannot be evaluated on the device: the descriptor Problem rejects them loudly, and device_problem routes them
    to the host-evaluated H-stack entry points (rg_*_from_hstack) instead."""
    with pytest.raises(_lib.DescriptorError):
        _lib.Problem(cz_problem_closures(10, 1.0), ctx=object())
    assert not _lib.is_descriptor_problem(cz_problem_closures(10, 1.0))
    assert _lib.is_descriptor_problem(cz_problem(10, 1.0, ("amp",)))


def test_hstack_layout_matches_reference_perturbation_pattern():
    """Order and arithmetic of the stacked Hamiltonians (src/UnitaryCalculations.jl:45-97)."""
    from robustgrape_b200 import rydberg_tools as rt
    fp = cz_problem_closures(4, 1.0, ("amp", "freq"))
    hp = _lib.HStackProblem.__new__(_lib.HStackProblem)
    hp.up, hp.fp = fp.unitary_problem, fp
    hp.ndim, hp.ntimes, hp.na, hp.nerr = 5, 4, 1, 2
    x = np.array([0.3, 0.7, 1.1, 1.9, 0.4])
    Hs, Ts, p = hp.stacks(x)
    assert p == 1 and Hs.shape == (5, 5, 1 + 4 + 2 * 4, 4) and Ts.shape == (5, 5, 2)
    k = 2
    assert np.array_equal(Hs[:, :, 0, k], rt.rydberg_hamiltonian_symmetric_blockaded(x[k], 0, 0))
    assert np.array_equal(Hs[:, :, 1, k], rt.rydberg_hamiltonian_symmetric_blockaded(x[k] + 1e-8, 0, 0))
    assert np.array_equal(Hs[:, :, 3, k], rt.rydberg_hamiltonian_symmetric_blockaded(x[k] + 1e-4, 0, 0))
    assert np.array_equal(Hs[:, :, 2, k], Hs[:, :, 0, k])          # the additional parameter does not enter this Hamiltonian
    H0 = rt.rydberg_hamiltonian_symmetric_blockaded(x[k], 0, 0)
    assert np.array_equal(Hs[:, :, 5, k], H0 + (rt.rydberg_hamiltonian_symmetric_blockaded(x[k], 1e-8, 0) - H0))
    assert np.array_equal(Hs[:, :, 8, k], H0 + (rt.rydberg_hamiltonian_symmetric_blockaded(x[k], 0, 1e-4) - H0))
    assert np.array_equal(Ts[:, :, 1], rt.cz_with_1q_phase_symmetric(x[4] + 1e-8))


def test_no_cpu_fallback(built_library):
    """Without a CUDA device the product path raises; it never routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RGError):
        rg.calculate_fidelity_and_derivatives(cz_problem(10, 1.0), np.zeros(11))
    src = "".join(p.read_text() for p in (ROOT / "robustgrape_b200").glob("*.py"))
    assert "oracle" not in src.replace("the oracle", "").replace("oracle/", ""), "product code must not import the oracle"


def test_term_flattening_roundtrip():
    fp = detuned_problem(7, 1.0, ("amp", "freq"))
    keep = []
    arr = _lib._terms_to_c(fp.unitary_problem.H0.terms, keep)
    assert arr[0].nfactors == 3 and arr[0].factors[0].kind == D.F_EXPI and arr[0].factors[0].scale == -1.0
    assert arr[2].factors[0].kind == D.F_VAR and arr[2].factors[0].space == D.S_ADD
    assert arr[0].nnz == 2 and arr[0].vals[0] == 0.5
