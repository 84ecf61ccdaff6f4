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


# ---- device -> host read-back accounting -------------------------------------------------------------------------------
# ``bench.py`` reports the bytes a step actually reads back from the device (losses, hit counts, herding group sizes, the
# mixing matrix when it is logged): every such site goes through :func:`host_list`, which counts what it copies.
import threading as _threading

_d2h = [0]
_d2h_lock = _threading.Lock()


def count_d2h(t: torch.Tensor) -> None:
    if t.is_cuda:
        with _d2h_lock:
            _d2h[0] += t.numel() * t.element_size()


def d2h_bytes() -> int:
    """Bytes read back from CUDA tensors through :func:`host_list` / :func:`count_d2h` so far (all threads)."""
    return _d2h[0]


def host_list(t: torch.Tensor) -> list:
    """``t.tolist()`` (a host sync when ``t`` lives on the device) with the copied bytes counted."""
    count_d2h(t)
    return t.tolist()
