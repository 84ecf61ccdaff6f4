"""Structured experiment log: ``{logs_dir}/{exp_name}-{time}.json`` with the reference schema
``{"config": ..., "data": {client: {round: {task: {tr_acc, tr_loss, val_rank_1/3/5/10, val_map}}}}}``
(``experiment.py:16-55,148-153,260-263,282-288``) — the input contract of :mod:`flpr_b200.analyse`.

Unlike the reference (which rewrites the whole file under a lock on every ``record``) writes are coalesced: the
file is flushed at most every ``flush_interval`` seconds and at ``close()`` / end of round.
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Any


class ExperimentLog:
    def __init__(self, save_path: str, flush_interval: float = 2.0, enabled: bool = True):
        self.records: dict = {}
        self.save_path = save_path
        self.lock = threading.Lock()
        self.flush_interval = flush_interval
        self.enabled = enabled
        self._dirty = False
        self._last = 0.0

    def _update_iter(self, key: str, value: Any) -> None:
        keys = key.split(".")
        cur = self.records
        for k in keys[:-1]:
            cur = cur.setdefault(k, {})
        last = keys[-1]
        if last not in cur:
            cur[last] = value
        elif isinstance(cur[last], list):
            cur[last].append(value)
        elif isinstance(cur[last], set):
            cur[last].add(value)
        elif isinstance(cur[last], dict) and isinstance(value, dict):
            cur[last].update(value)
        else:
            cur[last] = value

    def record(self, key: str, value: Any, flush: bool = False) -> None:
        with self.lock:
            self._update_iter(key, value)
            self._dirty = True
            if flush or time.time() - self._last >= self.flush_interval:
                self._save()

    def merge(self, other: dict) -> None:
        """Merge the ``data`` sub-tree produced by another rank (C7: metrics gathered to rank 0)."""
        def rec(dst, src):
            for k, v in src.items():
                if isinstance(v, dict) and isinstance(dst.get(k), dict):
                    rec(dst[k], v)
                else:
                    dst[k] = v
        with self.lock:
            rec(self.records, other)
            self._dirty = True

    def _save(self) -> None:
        if not self.enabled:
            self._dirty = False
            return
        os.makedirs(os.path.dirname(self.save_path) or ".", exist_ok=True)
        tmp = self.save_path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(self.records, f, indent=2, default=_jsonable)
        os.replace(tmp, self.save_path)
        self._dirty = False
        self._last = time.time()

    def flush(self) -> None:
        with self.lock:
            if self._dirty:
                self._save()

    close = flush


def _jsonable(o):
    try:
        import numpy as np
        if isinstance(o, np.generic):
            return o.item()
        if isinstance(o, np.ndarray):
            return o.tolist()
    except Exception:
        pass
    try:
        import torch
        if isinstance(o, torch.Tensor):
            return o.tolist()
    except Exception:
        pass
    return str(o)
