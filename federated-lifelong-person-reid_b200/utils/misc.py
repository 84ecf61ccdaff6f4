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


# ------------------------------------------------------------------------------------------------- small helpers
# The remaining helpers of the reference's ``tools/utils.py`` (API parity for user scripts; the engine itself does not
# need most of them: models are device-resident and the trunk / head split is static instead of an fx trace).
def torch_device(default_device: str | None = None, **kwargs) -> str:
    """``tools/utils.py:12-18``: 'cuda' when available, overridable by ``default_device`` or ``device=...``."""
    auto = "cuda" if torch.cuda.is_available() else "cpu"
    if default_device is not None:
        return default_device if default_device in ("cuda", "cpu") else auto
    return kwargs.get("device", auto)


def extract_kwargs(kwargs: Dict, key: str, default_value: Any = None) -> Any:
    return kwargs.get(key, default_value)


def extract_losses(losses: Any):
    """Sum of the scalar, differentiable tensors inside a nested container (``tools/utils.py:51-58``)."""
    if isinstance(losses, torch.Tensor):
        return losses if (losses.requires_grad and losses.dim() == 0) else 0.0
    if isinstance(losses, dict):
        return sum(extract_losses(v) for v in losses.values())
    if isinstance(losses, (list, tuple, set)):
        return sum(extract_losses(v) for v in losses)
    return 0.0


def random_shuffle(seed: int, items: list) -> None:
    random.Random(seed).shuffle(items)


def random_sample(seed: int, items: Iterable, num_pick: int) -> list:
    return random.Random(seed).sample(list(items), num_pick)


def random_int(seed: int, start: int, end: int) -> int:
    return random.Random(seed).randint(start, end)


def normalize(x, ord: int | None = None, axis: int = 0, keepdims: bool = True):
    return x / np.linalg.norm(x, ord=ord, axis=axis, keepdims=keepdims)


def np_save(base_dir: str, filename: str, data) -> None:
    import os
    os.makedirs(base_dir, exist_ok=True)
    np.save(os.path.join(base_dir, filename), data)


def load_task(base_dir: str, task: str):
    import os
    return np.load(os.path.join(base_dir, task), allow_pickle=True)


def tensor_value(*tensors: torch.Tensor):
    """Host values of scalar tensors with ONE synchronisation (the reference calls ``.cpu().item()`` per tensor)."""
    vals = torch.stack([t.detach().float().reshape(()) for t in tensors]).tolist()
    return vals[0] if len(vals) == 1 else tuple(vals)


class model_on_device:
    """``with model_on_device(model, device):`` (``tools/utils.py:110-121``). The reference shuttles the whole model
    CPU <-> GPU around every train / validate call; engine models are device-resident, so this only moves a model
    that is not already on ``device`` and leaves it there unless ``restore=True``."""

    def __init__(self, model: torch.nn.Module, device: str = "cpu", restore: bool = False) -> None:
        self.model, self.device, self.restore = model, torch.device(device), restore
        self._origin = None

    def __enter__(self):
        p = next(self.model.parameters(), None)
        self._origin = p.device if p is not None else None
        if self._origin is not None and self._origin != self.device:
            self.model.to(self.device)
        return self.model

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.restore and self._origin is not None and self._origin != self.device:
            self.model.to(self._origin)
        return False


def module_paths(net: torch.nn.Module, example: torch.Tensor | None = None) -> List[str]:
    """Qualified names of the leaf modules in *execution order* (what the reference's ``ModulePathTracer`` fx trace,
    ``tools/utils.py:139-182``, is used for: finding the first module inside the fine-tuned part). Uses forward hooks
    when an example input is given, definition order otherwise."""
    leaves = [(n, m) for n, m in net.named_modules() if n and not list(m.children())]
    if example is None:
        return [n for n, _ in leaves]
    order: List[str] = []
    hooks = [m.register_forward_pre_hook(lambda mod, inp, n=n: order.append(n)) for n, m in leaves]
    try:
        with torch.no_grad():
            net(example)
    finally:
        for h in hooks:
            h.remove()
    seen, out = set(), []
    for n in order:
        if n not in seen:
            seen.add(n)
            out.append(n)
    return out
