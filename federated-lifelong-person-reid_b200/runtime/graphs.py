"""CUDA-graph capture of launch-bound inner loops (one training step = ~170 kernel launches whose Python dispatch
costs more than their device time). A step function is run eagerly a few times (lazy initialisation, allocator
warm-up), then captured once and replayed with its inputs copied into static buffers."""
from __future__ import annotations

import threading
from typing import Callable, Dict, Sequence, Tuple

import torch

from ..ops import native


# One capture at a time per process. Captures use the thread-local error mode so that client threads that are NOT
# capturing (ExperimentStage runs `parallel` clients per device concurrently, each on its own stream) may keep
# launching / allocating while another thread captures.
CAPTURE_LOCK = threading.RLock()


_capture_streams: dict = {}


def capture(graph: "torch.cuda.CUDAGraph", **kw):
    """``torch.cuda.graph`` in thread-local error mode on a capture stream of its own (``torch.cuda.graph`` would take
    one from the 32-stream pool, which a client thread's stream may alias - see ``native.dedicated_stream``).
    Captures are serialised by ``CAPTURE_LOCK``, so one capture stream per device is enough."""
    if "stream" not in kw:
        dev = torch.cuda.current_device()
        st = _capture_streams.get(dev)
        if st is None:
            st = _capture_streams[dev] = native.dedicated_stream(torch.device("cuda", dev))
        kw["stream"] = st
    return torch.cuda.graph(graph, capture_error_mode="thread_local", **kw)


class GraphedStep:
    """``step(*tensors)`` -> eager for the first ``warmup`` calls of a given input signature, captured on the next
    call, replayed afterwards. The callable must be free of host synchronisation and keep its outputs in
    pre-allocated device tensors (accumulators)."""

    def __init__(self, fn: Callable, warmup: int = 2, enabled: bool = True):
        self.fn = fn
        self.warmup = warmup
        self.enabled = enabled
        self._state: Dict[Tuple, dict] = {}

    @staticmethod
    def _sig(tensors: Sequence[torch.Tensor]) -> Tuple:
        return tuple((tuple(t.shape), t.dtype, t.stride()) for t in tensors)

    def __call__(self, *tensors: torch.Tensor) -> None:
        if not self.enabled or not tensors[0].is_cuda:
            self.fn(*tensors)
            return
        sig = self._sig(tensors)
        st = self._state.setdefault(sig, {"eager": 0, "graph": None})
        if st["graph"] is None:
            if st["eager"] < self.warmup:
                st["eager"] += 1
                self.fn(*tensors)
                return
            with CAPTURE_LOCK:
                st["static"] = [t.clone() for t in tensors]
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                native.capture_count_begin()
                with capture(g):
                    self.fn(*st["static"])
                st["launches"] = native.capture_count_end()
                native.count_launch(-st["launches"])          # capture recorded, did not execute
                st["graph"] = g
        for s, t in zip(st["static"], tensors):
            s.copy_(t, non_blocking=True)
        st["graph"].replay()
        native.count_launch(st["launches"])

    def reset(self) -> None:
        self._state.clear()


class GraphedForward:
    """``y = fn(x)`` for a no-grad forward (e.g. the eval-mode head that produces herding features): eager for the first
    ``warmup`` calls of an input signature, then captured and replayed; returns a tensor the caller owns."""

    def __init__(self, fn: Callable, warmup: int = 1, enabled: bool = True):
        self.fn, self.warmup, self.enabled = fn, warmup, enabled
        self._state: Dict[Tuple, dict] = {}

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not self.enabled or not x.is_cuda:
            return self.fn(x)
        sig = (tuple(x.shape), x.dtype, x.stride())
        st = self._state.setdefault(sig, {"eager": 0, "graph": None})
        if st["graph"] is None:
            if st["eager"] < self.warmup:
                st["eager"] += 1
                return self.fn(x)
            with CAPTURE_LOCK:
                st["in"] = x.clone()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                native.capture_count_begin()
                with capture(g):
                    st["out"] = self.fn(st["in"])
                st["launches"] = native.capture_count_end()
                native.count_launch(-st["launches"])
                st["graph"] = g
        st["in"].copy_(x, non_blocking=True)
        st["graph"].replay()
        native.count_launch(st["launches"])
        return st["out"].clone()

    def reset(self) -> None:
        self._state.clear()
