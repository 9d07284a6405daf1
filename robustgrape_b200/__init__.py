"""robustgrape_b200 -- B200-native GRAPE propagator hot path behind the RobustGRAPE.jl API."""
from .types import (ErrorSource, UnitaryRobustGRAPEProblem, FidelityRobustGRAPEProblem,
                    FidelityRobustGRAPEParameters)
from . import rydberg_tools, descriptors
