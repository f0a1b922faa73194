"""NVTX ranges around collectives so ``ncu``/profilers show op boundaries."""
import contextlib

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    pushed = False
    if torch.cuda.is_available():
        try:
            torch.cuda.nvtx.range_push(name)
            pushed = True
        except Exception:  # pragma: no cover - nvtx missing
            pushed = False
    try:
        yield
    finally:
        if pushed:
            torch.cuda.nvtx.range_pop()
