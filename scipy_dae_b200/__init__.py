"""scipy_dae_b200 — B200-native ensemble DAE/IDE integrator (drop-in for scipy_dae's hot path).

Importing the package does not need a GPU; calling a solver does, and fails loudly when
libb200dae.so (the sm_100a CUDA library) is missing — there is no CPU fallback.
"""
from .batched import BatchedDaeResult, FunctorEvent, ThresholdEvent, solve_dae_batched, device_residual
from .host_pipeline import solve_dae_batched_host, alloc_host_result, HostResultBuffers
from .consistent_ic import consistent_initial_conditions_batched, ConsistentICResult
from .runtime import register_problem, precompile
from .problems import PROBLEMS, Problem, get_problem, perturbed_params, initial_conditions

__all__ = ["solve_dae_batched", "ThresholdEvent", "FunctorEvent", "solve_dae_batched_host", "alloc_host_result", "HostResultBuffers",
           "consistent_initial_conditions_batched", "ConsistentICResult", "BatchedDaeResult", "device_residual", "PROBLEMS", "Problem",
           "get_problem", "perturbed_params", "initial_conditions", "register_problem", "precompile"]
__version__ = "0.1.0"
