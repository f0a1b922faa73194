"""Functional ops on top of the communicator: fused-epilogue collectives and
the Allreduce->GEMM fused linear layer."""
from .functional import allreduce, allreduce_mean, allreduce_sgd_step_, average_parameters_flat
from .fused_linear import allreduce_linear, has_fused_kernel

__all__ = ["allreduce", "allreduce_mean", "allreduce_sgd_step_", "average_parameters_flat", "allreduce_linear",
           "has_fused_kernel"]
