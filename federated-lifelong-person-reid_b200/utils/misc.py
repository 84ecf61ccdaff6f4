"""Small utilities mirroring ``tools/utils.py`` (seeding, state sizes, tensor helpers) plus device timing."""
from __future__ import annotations

import gc
import random
from contextlib import contextmanager
from typing import Any, Dict, Iterable, List, Set, Tuple

import numpy as np
import torch


def same_seeds(seed: int = 42069) -> None:
    """``tools/utils.py:92-100``."""
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


def params_state_size(state: Any) -> int:
    """Number of scalars in a nested payload (``tools/utils.py:39-48``); used for communication accounting."""
    if state is None:
        return 0
    if isinstance(state, (int, float, bool, complex, str)):
        return 1
    if isinstance(state, torch.Tensor):
        return state.numel()
    if isinstance(state, np.ndarray):
        return int(state.size)
    if isinstance(state, Dict):
        return sum(params_state_size(v) for v in state.values())
    if isinstance(state, (List, Tuple, Set)):
        return sum(params_state_size(v) for v in state)
    raise TypeError(f"unrecognized state type {type(state)} to calculate parameters size")


def tensor_reverse_permute(t: torch.Tensor | None) -> torch.Tensor | None:
    """``tools/utils.py:27-32`` (FedWeIT stores weights reverse-permuted)."""
    if t is None:
        return None
    return t.permute(*reversed(range(t.dim())))


def get_one_hot(target: torch.Tensor, num_class: int) -> torch.Tensor:
    return torch.zeros(target.shape[0], num_class, device=target.device).scatter_(1, target.long().view(-1, 1), 1.0)


def clear_cache() -> None:
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


class DeviceTimer:
    """CUDA-event timer on the current stream (host wall-clock on CPU). ``with timer('phase'): ...``"""

    def __init__(self, device: torch.device | str):
        self.device = torch.device(device)
        self.records: Dict[str, List] = {}
        self._pending: List[Tuple[str, Any, Any]] = []

    @contextmanager
    def __call__(self, name: str):
        if self.device.type == "cuda":
            from .trace import nvtx_range
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with nvtx_range(f"flpr/{name}"):
                yield
            e1.record()
            self._pending.append((name, e0, e1))
        else:
            import time
            t0 = time.perf_counter()
            yield
            self.records.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    def flush(self) -> Dict[str, List[float]]:
        if self._pending:
            torch.cuda.synchronize(self.device)
            for name, e0, e1 in self._pending:
                self.records.setdefault(name, []).append(e0.elapsed_time(e1))
            self._pending.clear()
        return self.records

    def total_ms(self, name: str) -> float:
        self.flush()
        return float(sum(self.records.get(name, [])))
