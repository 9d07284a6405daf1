// rg_analysis.cuh -- kernels for the analysis entry points (materialised unitary derivatives,
// interaction-picture error operators, fidelity response, expectation values).
#pragma once
#include "rg_common.cuh"
