"""Tracing hooks (the reference has none). NVTX ranges around every phase of a round and of a client's
local step - visible in Nsight Systems / ``torch.profiler`` traces (``scripts/trace_round.py``) - plus a tiny
wall-clock section timer for host-side accounting. All no-ops on CPU."""
from __future__ import annotations

import os
from contextlib import contextmanager

import torch

_ENABLED = os.environ.get("FLPR_NVTX", "1") != "0"


@contextmanager
def nvtx_range(name: str):
    on = _ENABLED and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
