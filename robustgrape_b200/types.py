"""Problem structs -- field-for-field mirror of the reference's `Parameters.@with_kw`
structs (reference src/Types.jl:12-14, 31-40, 52-56, 74-84).  Names, defaults
(eps = 1e-8, eps2 = 1e-4, iterations = 1000, time_limit = NaN) and meaning are the
reference's; the Greek field names are spelled `eps` / `eps2`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Sequence

import numpy as np


@dataclass
class ErrorSource:
    """reference src/Types.jl:12-14.  Herror(time_step, x, x_add, err) -> matrix."""
    Herror: Callable


@dataclass
class UnitaryRobustGRAPEProblem:
    """reference src/Types.jl:31-40.  H0(time_step, x, x_add) -> matrix; time_step is 1-based."""
    t0: float
    ntimes: int
    ndim: int
    H0: Callable
    nb_additional_param: int
    error_sources: List[ErrorSource]
    eps: float = 1e-8
    eps2: float = 1e-4


@dataclass
class FidelityRobustGRAPEProblem:
    """reference src/Types.jl:52-56.  projector may be a "pseudo"-projector (real weights)."""
    unitary_problem: UnitaryRobustGRAPEProblem
    projector: np.ndarray
    target_unitary: Callable


@dataclass
class FidelityRobustGRAPEParameters:
    """reference src/Types.jl:74-84.  `solver_algorithm` names a scipy.optimize method here."""
    x_initial: np.ndarray
    regularization_functions: Sequence[Callable]
    regularization_coeff1: Sequence[float]
    regularization_coeff2: Sequence[float]
    error_source_coeff: Sequence[float]
    time_limit: float = math.nan
    iterations: int = 1000
    solver_algorithm: str = "L-BFGS-B"
    additional_parameters: Dict[str, Any] = field(default_factory=dict)
