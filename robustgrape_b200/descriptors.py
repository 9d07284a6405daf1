"""Declarative Hamiltonian / target descriptors.

The reference's problem structs hold opaque Julia closures (`H0::Function`,
`Herror::Function`, `target_unitary::Function`; reference src/Types.jl:13,35,55).
A GPU cannot call a closure, so the drop-in boundary carries a *term list*
instead: every operator is

    sum_t  coef_t * prod_f factor_f(variables)  *  M_t        (M_t constant, sparse)

The classes here are callable with the reference's closure signatures
(`H0(time_step, x, x_add)`, `Herror(time_step, x, x_add, err)`,
`target_unitary(x_add)`), so they are valid field values for the problem
structs *and* evaluate on the CPU to the same matrices as the reference's
`RydbergTools` builders (reference src/RydbergTools.jl:31-39,71-81,118-130,
160-162,197-203).  `to_c()` flattens them into the `rg_term` records declared in
include/robustgrape_b200.h.

The evaluation is written against a tiny numeric backend (`lib`) so the same
term list can be evaluated in numpy complex128 (product/host path) or in mpmath
at high precision (the exact-semantics checker under oracle/ passes `lib=mp`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

# factor kinds -- keep in sync with include/robustgrape_b200.h
F_VAR = 0        # value = scale * v + offset
F_COS = 1        # cos(scale * v + offset)
F_SIN = 2        # sin(scale * v + offset)
F_EXPI = 3       # exp(+i (scale * v + offset))
F_ERR = 4        # err                      (amplitude of the owning error source)
F_ERR1P_M1 = 5   # fl(1 + err) - 1          ("(1+eps)" as written in RydbergTools, minus H0's 1)
F_TABLE = 6      # table[time_step, index]  (per-step real envelope)

# variable spaces
S_MAIN = 0       # x[index] at the current time step
S_ADD = 1        # x_add[index]
S_NONE = 2

OWNER_H0 = -1
OWNER_TARGET = -2

MAX_FACTORS = 4


@dataclass(frozen=True)
class Factor:
    kind: int
    space: int = S_NONE
    index: int = 0
    scale: float = 1.0
    offset: float = 0.0

    # -- constructors -----------------------------------------------------
    @staticmethod
    def var(space, index, scale=1.0, offset=0.0):
        return Factor(F_VAR, space, index, scale, offset)

    @staticmethod
    def cos(space, index, scale=1.0, offset=0.0):
        return Factor(F_COS, space, index, scale, offset)

    @staticmethod
    def sin(space, index, scale=1.0, offset=0.0):
        return Factor(F_SIN, space, index, scale, offset)

    @staticmethod
    def expi(space, index, scale=1.0, offset=0.0):
        return Factor(F_EXPI, space, index, scale, offset)

    @staticmethod
    def err():
        return Factor(F_ERR)

    @staticmethod
    def err1p_m1():
        return Factor(F_ERR1P_M1)

    @staticmethod
    def table(col):
        return Factor(F_TABLE, S_NONE, col)

    def depends_on(self, space, index):
        return self.kind in (F_VAR, F_COS, F_SIN, F_EXPI) and self.space == space and self.index == index

    def value(self, time_step, x, x_add, err, table, lib):
        """Evaluate the factor. `time_step` is 1-based like the reference."""
        if self.kind == F_ERR:
            return err
        if self.kind == F_ERR1P_M1:
            return (1.0 + err) - 1.0 if lib is np else (lib.mpf(float(1.0 + float(err))) - 1)
        if self.kind == F_TABLE:
            return table[time_step - 1][self.index]
        v = x[self.index] if self.space == S_MAIN else x_add[self.index]
        u = self.scale * v + self.offset if (self.scale != 1.0 or self.offset != 0.0) else v
        if self.kind == F_VAR:
            return u
        if self.kind == F_COS:
            return lib.cos(u)
        if self.kind == F_SIN:
            return lib.sin(u)
        if self.kind == F_EXPI:
            if lib is np:
                return complex(np.cos(u), np.sin(u))
            return lib.mpc(lib.cos(u), lib.sin(u))
        raise ValueError(f"bad factor kind {self.kind}")


@dataclass
class Term:
    """coef * prod(factors) * M, with M given as a sparse (row, col, value) list (0-based)."""
    coef: complex
    factors: Tuple[Factor, ...]
    entries: Tuple[Tuple[int, int, complex], ...]
    owner: int = OWNER_H0

    def __post_init__(self):
        self.factors = tuple(self.factors)
        self.entries = tuple((int(r), int(c), complex(v)) for r, c, v in self.entries)
        if len(self.factors) > MAX_FACTORS:
            raise ValueError(f"at most {MAX_FACTORS} factors per term")

    def coefficient(self, time_step, x, x_add, err, table, lib):
        c = self.coef if lib is np else lib.mpc(self.coef.real, self.coef.imag)
        for f in self.factors:
            c = c * f.value(time_step, x, x_add, err, table, lib)
        return c


def _dense(ndim, terms: Sequence[Term], time_step, x, x_add, err, table, lib):
    if lib is np:
        out = np.zeros((ndim, ndim), dtype=np.complex128)
        for t in terms:
            c = t.coefficient(time_step, x, x_add, err, table, lib)
            for r, cc, v in t.entries:
                out[r, cc] += c * v
        return out
    out = lib.zeros(ndim, ndim)
    for t in terms:
        c = t.coefficient(time_step, x, x_add, err, table, lib)
        for r, cc, v in t.entries:
            out[r, cc] += c * lib.mpc(v.real, v.imag)
    return out


@dataclass
class TermOperator:
    """Base: an operator given as a list of terms."""
    ndim: int
    terms: List[Term]
    table: np.ndarray | None = None   # (ntimes, ncols) real, optional

    def matrix(self, time_step, x, x_add, err=0.0, lib=np):
        return _dense(self.ndim, self.terms, time_step, x, x_add, err, self.table, lib)

    def is_hermitian(self) -> bool:
        """True when the term list is Hermitian for all real variable values:
        checked structurally by evaluating at a few random points."""
        rng = np.random.default_rng(7)
        nmain = 1 + max([f.index for t in self.terms for f in t.factors if f.space == S_MAIN] + [0])
        nadd = 1 + max([f.index for t in self.terms for f in t.factors if f.space == S_ADD] + [0])
        steps = 1 if self.table is None else min(3, self.table.shape[0])
        for k in range(1, steps + 1):
            for _ in range(3):
                m = self.matrix(k, rng.normal(size=nmain), rng.normal(size=nadd), err=rng.normal())
                if not np.allclose(m, m.conj().T, rtol=0, atol=1e-13 * max(1.0, np.abs(m).max())):
                    return False
        return True


class TermHamiltonian(TermOperator):
    """Valid value for `UnitaryRobustGRAPEProblem.H0`: callable as H0(time_step, x, x_add)."""

    def __call__(self, time_step, x, x_add, lib=np):
        return self.matrix(time_step, x, x_add, 0.0, lib)


class TermErrorHamiltonian(TermOperator):
    """Valid value for `ErrorSource.Herror`: callable as Herror(time_step, x, x_add, err).
    Every term must contain an err-type factor so that Herror(..., 0) == 0."""

    def __post_init__(self):
        for t in self.terms:
            if not any(f.kind in (F_ERR, F_ERR1P_M1) for f in t.factors):
                raise ValueError("every error-Hamiltonian term needs an err factor")

    def __call__(self, time_step, x, x_add, err, lib=np):
        return self.matrix(time_step, x, x_add, err, lib)


class TermTarget(TermOperator):
    """Valid value for `FidelityRobustGRAPEProblem.target_unitary`: callable as U0(x_add)."""

    def __call__(self, x_add, lib=np):
        return self.matrix(1, (), x_add, 0.0, lib)


class ConstantTarget(TermTarget):
    """Target unitary that does not depend on x_add."""

    def __init__(self, U0):
        U0 = np.asarray(U0, dtype=np.complex128)
        ent = [(r, c, U0[r, c]) for r in range(U0.shape[0]) for c in range(U0.shape[1]) if U0[r, c] != 0]
        super().__init__(U0.shape[0], [Term(1.0, (), tuple(ent), OWNER_TARGET)])
