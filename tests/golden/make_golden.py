"""Generates the golden fixtures under tests/golden/ (run in the CPU container; ~2 minutes):

    python tests/golden/make_golden.py

Each fixture holds a seeded input pulse and the outputs of `calculate_fidelity_and_derivatives`
(reference src/FidelityCalculations.jl:19-119) evaluated by BOTH checkers:
  exact_*  -- oracle/exact_oracle.py (mpmath, 50 digits): the exact value of the reference's formulas;
  fp64_*   -- oracle/reference_oracle.py (numpy/scipy complex128): the literal FP64 restatement.
The reference itself (Julia) cannot be run in this image and ships no golden vectors, so these are
restatement values: PARITY UNPINNED beyond the structure tests in tests/test_oracle_structure.py.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from cases import golden_cases  # noqa: E402
from oracle import exact_oracle as eo, reference_oracle as ro  # noqa: E402

if __name__ == "__main__":
    if "--export-inputs" in sys.argv:          # inputs of every case as CSV for tests/golden/make_golden.jl (the Julia reference run)
        for name in golden_cases():
            z = np.load(Path(__file__).parent / f"{name}.npz")
            np.savetxt(Path(__file__).parent / f"{name}_x.csv", z["x"][None, :], delimiter=",", fmt="%.17g")
        sys.exit(0)
    only = set(sys.argv[1:])
    for name, (fp, x) in golden_cases().items():
        if only and name not in only:
            continue
        e = eo.calculate_fidelity_and_derivatives(fp, x)
        a = ro.calculate_fidelity_and_derivatives(fp, x)
        out = {"x": x}
        for k, ev, av in zip(["F", "F_dx", "F_d2err", "F_d2err_dx"], e, a):
            out["exact_" + k] = np.asarray(ev, dtype=np.float64)
            out["fp64_" + k] = np.asarray(av, dtype=np.float64)
        np.savez(Path(__file__).parent / f"{name}.npz", **out)
        print(name, "F =", repr(e[0]))
