"""Functional ops on top of the communicator: fused-epilogue collectives and
the Allreduce->GEMM fused linear layer."""
from .functional import allreduce, allreduce_mean, allreduce_sgd_step_, average_parameters_flat
from .fused_linear import (InBackwardSGD, allreduce_linear, dp_linear_mse, dp_linear_mse_supported, has_fused_kernel,
                           in_backward_sgd_supported)

__all__ = ["allreduce", "allreduce_mean", "allreduce_sgd_step_", "average_parameters_flat", "allreduce_linear",
           "has_fused_kernel", "dp_linear_mse", "dp_linear_mse_supported", "InBackwardSGD", "in_backward_sgd_supported"]
