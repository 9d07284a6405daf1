"""ORACLE (test infrastructure, NOT product code) -- ctypes wrapper of oracle/cpu_port.cpp, the C++
port of the reference's literal CPU algorithm (PARITY UNPINNED, see reference_oracle.py).
Used by tests as a second checker and by bench.py as the timed CPU baseline."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from robustgrape_b200 import _lib as L          # struct layouts of include/robustgrape_b200.h only
from robustgrape_b200 import descriptors as D

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "liboracle_port.so"
_so = None


def build():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


def lib():
    global _so
    if _so is None:
        if not SO.exists():
            build()
        _so = C.CDLL(str(SO))
        vp = C.c_void_p
        _so.oracle_fidelity_and_derivatives_batch.restype = C.c_int
        _so.oracle_fidelity_and_derivatives_batch.argtypes = [C.POINTER(L.rg_problem_desc), C.c_int, vp, vp, vp, vp, vp, C.c_int]
        _so.oracle_cost_and_grad_batch.restype = C.c_int
        _so.oracle_cost_and_grad_batch.argtypes = [C.POINTER(L.rg_problem_desc), C.c_int, vp, vp, vp, vp, C.c_int]
        _so.oracle_max_threads.restype = C.c_int
    return _so


class PortProblem:
    """Host-side descriptor (same rg_problem_desc the CUDA library takes) for the C++ port."""

    def __init__(self, fp, nparam=1):
        up = fp.unitary_problem
        self.keep = []
        terms = list(up.H0.terms)
        for e, src in enumerate(up.error_sources):
            terms += [D.Term(t.coef, t.factors, t.entries, e) for t in src.Herror.terms]
        tterms = [D.Term(t.coef, t.factors, t.entries, D.OWNER_TARGET) for t in fp.target_unitary.terms]
        d = L.rg_problem_desc()
        d.ndim, d.ntimes, d.nparam = up.ndim, up.ntimes, nparam
        d.nb_additional_param, d.nerr = up.nb_additional_param, len(up.error_sources)
        d.t0, d.eps, d.eps2 = float(up.t0), float(up.eps), float(up.eps2)
        self.terms = L._terms_to_c(terms, self.keep)
        self.tterms = L._terms_to_c(tterms, self.keep)
        d.nterms, d.terms = len(terms), self.terms
        d.ntarget_terms, d.target_terms = len(tterms), self.tterms
        self.proj = np.asfortranarray(np.asarray(fp.projector, dtype=np.float64))
        d.projector = self.proj.ctypes.data_as(C.POINTER(C.c_double))
        if up.H0.table is not None:
            self.tab = np.asfortranarray(np.asarray(up.H0.table, dtype=np.float64))
            d.ntable_cols, d.table = self.tab.shape[1], self.tab.ctypes.data_as(C.POINTER(C.c_double))
        self.desc = d
        self.nx = nparam * up.ntimes + up.nb_additional_param
        self.nerr = len(up.error_sources)

    def fidelity_and_derivatives_batch(self, X, nthreads=0):
        X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(self.nx, -1))
        B = X.shape[1]
        F = np.zeros(B); Fdx = np.zeros((self.nx, B), order="F")
        F2 = np.zeros((self.nerr, B), order="F"); F2dx = np.zeros((self.nx, self.nerr, B), order="F")
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib().oracle_fidelity_and_derivatives_batch(C.byref(self.desc), B, p(X), p(F), p(Fdx), p(F2), p(F2dx), nthreads)
        return F, Fdx, F2, F2dx

    def cost_and_grad_batch(self, X, coeff=(), nthreads=0):
        X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(self.nx, -1))
        B = X.shape[1]
        cost = np.zeros(B); grad = np.zeros((self.nx, B), order="F")
        coeff = np.asarray(coeff, dtype=np.float64) if self.nerr else np.zeros(1)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib().oracle_cost_and_grad_batch(C.byref(self.desc), B, p(X), p(coeff), p(cost), p(grad), nthreads)
        return cost, grad


def max_threads():
    return int(lib().oracle_max_threads())
