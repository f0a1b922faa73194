"""Utilities: device timing, clock sampling, NVTX ranges, env-gated logging."""
from .timing import ClockSampler, L2Flusher, cuda_time_ms, max_over_ranks
from .nvtx import nvtx_range
from .log import debug, is_debug

__all__ = ["ClockSampler", "L2Flusher", "cuda_time_ms", "max_over_ranks", "nvtx_range", "debug", "is_debug"]
