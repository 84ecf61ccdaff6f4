"""Checkpoint store with the reference's directory / key layout:

    {checkpoints_dir}/{exp_name}/{actor}/{state_name}.ckpt        (torch.save of nested dicts of CPU tensors)

The reference uses checkpoints as its *working store* (every train/validate starts with ``load_model`` from disk).
Here the working state stays resident on the device; checkpoints are snapshots written by a background thread
(device->pinned-host copy on a side stream, ``torch.save`` off the critical path), plus a real resume manifest
(round, RNG state, task-pipeline positions, server registry) that the reference lacks.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Any, Dict, Optional

import torch

os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")  # exemplar ckpts hold numpy objects


def _to_cpu(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.detach().to("cpu", copy=True)
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


class CheckpointStore:
    """One instance per process; actors address it with ``(actor_name, state_name)``."""

    def __init__(self, root: str, asynchronous: bool = True, enabled: bool = True):
        self.root = root
        self.enabled = enabled
        self.asynchronous = asynchronous and enabled
        self._q: "queue.Queue" = queue.Queue()
        self._thread: Optional[threading.Thread] = None
        self._err: Optional[BaseException] = None
        self.bytes_written = 0
        if self.asynchronous:
            self._thread = threading.Thread(target=self._worker, name="flpr-ckpt-writer", daemon=True)
            self._thread.start()

    def path(self, actor: str, state_name: str) -> str:
        return os.path.join(self.root, actor, f"{state_name}.ckpt")

    def exists(self, actor: str, state_name: str) -> bool:
        return os.path.exists(self.path(actor, state_name))

    def _worker(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            path, state = item
            try:
                self._write(path, state)
            except BaseException as ex:  # surfaced on the next save / flush
                self._err = ex
            finally:
                self._q.task_done()

    def _write(self, path: str, state: Any) -> None:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp"
        torch.save(state, tmp)
        os.replace(tmp, path)
        try:
            self.bytes_written += os.path.getsize(path)
        except OSError:
            pass

    def save(self, actor: str, state_name: Optional[str], state: Any, cover: bool = False) -> None:
        """``save_state`` of ``modules/client.py:52-63`` / ``modules/server.py:46-57``."""
        if state_name is None or not self.enabled:
            return
        if self._err is not None:
            err, self._err = self._err, None
            raise err
        path = self.path(actor, state_name)
        if cover is False and os.path.exists(path):
            raise ValueError(f"State checkpoint has already exist in '{path}'.")
        snap = _to_cpu(state)  # snapshot now; serialisation happens on the writer thread
        if self.asynchronous:
            self._q.put((path, snap))
        else:
            self._write(path, snap)

    def load(self, actor: str, state_name: str, default_value: Any = None, map_location: str = "cpu") -> Any:
        """``load_state`` (``modules/client.py:34-50``): returns ``default_value`` when the file does not exist."""
        self.flush()
        path = self.path(actor, state_name)
        if os.path.exists(path):
            return torch.load(path, map_location=map_location, weights_only=False)
        if default_value is not None:
            return default_value
        raise ValueError(f"State checkpoint does not exist in '{path}'.")

    def flush(self) -> None:
        if self.asynchronous:
            self._q.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def close(self) -> None:
        if self.asynchronous and self._thread is not None:
            self._q.put(None)
            self._thread.join()
            self._thread = None
            self.asynchronous = False
