"""Env-gated debug logging (``M4T_DEBUG=1``); the reference has only ``#if 0``
diagnostics (reference csrc/extension.cpp:1377-1392)."""
import os
import sys


def is_debug() -> bool:
    return os.environ.get("M4T_DEBUG", "0") not in ("", "0")


def debug(msg: str) -> None:
    if is_debug():
        rank = os.environ.get("RANK", "0")
        sys.stderr.write(f"[m4t:{rank}] {msg}\n")
        sys.stderr.flush()
