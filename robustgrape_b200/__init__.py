"""robustgrape_b200 -- B200-native GRAPE propagator hot path behind the RobustGRAPE.jl API.

Exports mirror reference src/RobustGRAPE.jl:6-13.  Importing the package does not need a GPU;
calling any calculate_* function does (there is no CPU fallback)."""
from .types import (ErrorSource, UnitaryRobustGRAPEProblem, FidelityRobustGRAPEProblem,
                    FidelityRobustGRAPEParameters)
from . import rydberg_tools, descriptors
from .unitary_calculations import calculate_unitary_and_derivatives, calculate_interaction_error_operators
from .fidelity_calculations import (calculate_fidelity_and_derivatives, optimize_fidelity_and_error_sources,
                                    calculate_fidelity_response, calculate_fidelity_response_fft,
                                    calculate_expectation_values, calculate_fidelity_and_derivatives_batch,
                                    cost_and_gradient_batch, optimize_batch_device)
from .regularization import regularization_cost, regularization_cost_phase

RydbergTools = rydberg_tools
