"""Utilities: device timing, clock sampling, NVTX ranges, env-gated logging."""
from .timing import ClockSampler, L2Flusher, cuda_time_ms, max_over_ranks
from .nvtx import nvtx_range
from .log import debug, is_debug
from .affinity import bind_to_gpu_numa, gpu_local_cpus

__all__ = ["ClockSampler", "L2Flusher", "cuda_time_ms", "max_over_ranks", "nvtx_range", "debug", "is_debug", "bind_to_gpu_numa",
           "gpu_local_cpus"]
