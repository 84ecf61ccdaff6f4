"""Checkpoint store with the reference's directory / key layout:

    {checkpoints_dir}/{exp_name}/{actor}/{state_name}.ckpt        (torch.save of nested dicts of CPU tensors)

The reference uses checkpoints as its *working store* (every train/validate starts with ``load_model`` from disk and
every payload is ``torch.save``d on the critical path). Here the working state stays resident on the device and a
checkpoint is a snapshot that leaves the critical path in three hops:

  1. device -> pinned, process-shared staging slab: ``copy_(non_blocking=True)`` on a dedicated copy stream
     (the compute stream only records an event; PCIe DMA overlaps the next client's training);
  2. a feeder thread waits for the copy event (GIL released) and hands the slab to a pool of writer *processes*
     (threads would fight the training loop for the GIL during pickling);
  3. a writer process runs ``torch.save`` into ``path.tmp`` and renames it; the slab returns to the pool.

``flush()`` drains the pipeline; ``load()`` flushes first so a reader always sees the latest snapshot.
Set ``asynchronous=False`` for the simple in-line behaviour (tests, tiny CPU runs).
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Any, Callable, Dict, List, Optional, Tuple

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


def _own(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.clone()
    if isinstance(obj, dict):
        return {k: _own(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_own(v) for v in obj)
    return obj


# ----------------------------------------------------------------------------------------------- post transforms
def expand_examplars(state: Any) -> Any:
    """Writer-side conversion of the compact exemplar memory into the reference's on-disk schema
    ``{np.int64 pid: [(ndarray prototype, class_id), ...]}`` (``methods/fedstil.py:841,846``)."""
    import numpy as np
    if isinstance(state, dict) and "_compact_gens" in state:
        out = {}
        for gen in state["_compact_gens"]:
            k = int(gen["k"])
            bank = gen["bank"][:, :k].float().numpy()
            cls = gen["cls"][:, :k].tolist()
            for gi, pid in enumerate(gen["pids"].tolist()):
                out[np.int64(pid)] = [(bank[gi, j], int(cls[gi][j])) for j in range(k)]
        return out
    if not isinstance(state, dict) or "_compact_examplars" not in state:
        return state
    out = {}
    for pid, ex in state["_compact_examplars"].items():
        bank = ex["bank"].float().numpy()
        cls = ex["cls"].tolist()
        out[np.int64(pid)] = [(bank[i], int(cls[i])) for i in ex["order"]]
    return out


POST: Dict[str, Callable[[Any], Any]] = {"expand_examplars": expand_examplars}


def _writer_main(jobs, done) -> None:  # pragma: no cover  (runs in a child process)
    torch.set_num_threads(1)
    while True:
        job = jobs.get()
        if job is None:
            return
        job_id, path, state, post = job
        err = None
        try:
            state = _own(state)          # detach from the staging slab (torch.save would serialise the whole slab)
            if post:
                state = POST[post](state)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = f"{path}.tmp{os.getpid()}"
            # legacy (non-zip) container: no per-record CRC32 pass over hundreds of MB; torch.load reads both formats
            torch.save(state, tmp, _use_new_zipfile_serialization=False)
            os.replace(tmp, path)
        except BaseException as ex:
            err = f"{type(ex).__name__}: {ex}"
        del state, job
        done.put((job_id, err))


class _Slab:
    """A process-shared, page-locked byte buffer that tensors are staged into."""

    def __init__(self, nbytes: int, cuda: bool):
        self.buf = torch.empty(nbytes, dtype=torch.uint8).share_memory_()
        self.registered = False
        if cuda:
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), nbytes, 0)
            self.registered = int(rc) == 0
        self.nbytes = nbytes

    def release(self) -> None:
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self.registered = False


class CheckpointStore:
    """One instance per process; actors address it with ``(actor_name, state_name)``."""

    def __init__(self, root: str, asynchronous: bool = True, enabled: bool = True, workers: int = 6,
                 max_inflight_bytes: int = 24 << 30):
        self.root = root
        self.enabled = enabled
        self.asynchronous = bool(asynchronous and enabled)
        self.workers = max(1, int(workers))
        self.max_inflight = int(max_inflight_bytes)
        self.bytes_written = 0
        self._started = False
        self._err: Optional[str] = None
        self._next_id = 0
        self._inflight: Dict[int, Tuple[_Slab, int]] = {}
        self._inflight_bytes = 0
        self._pool: Dict[int, List[_Slab]] = {}
        self._lock = threading.RLock()
        self._copy_stream = None
        self._last_copy_event = None

    # ------------------------------------------------------------------ paths
    def path(self, actor: str, state_name: str) -> str:
        return os.path.join(self.root, actor, f"{state_name}.ckpt")

    def exists(self, actor: str, state_name: str) -> bool:
        self.flush()
        return os.path.exists(self.path(actor, state_name))

    # ------------------------------------------------------------------ pipeline
    def _start(self) -> None:
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        self._jobs = ctx.Queue()
        self._done = ctx.Queue()
        self._procs = [ctx.Process(target=_writer_main, args=(self._jobs, self._done), daemon=True)
                       for _ in range(self.workers)]
        for p in self._procs:
            p.start()
        self._feed_q: "queue.Queue" = queue.Queue()
        self._feeder = threading.Thread(target=self._feed, name="flpr-ckpt-feeder", daemon=True)
        self._feeder.start()
        self._started = True

    def _feed(self) -> None:
        while True:
            item = self._feed_q.get()
            if item is None:
                return
            event, job = item
            if event is not None:
                event.synchronize()                      # releases the GIL while the DMA finishes
            self._jobs.put(job)

    def _get_slab(self, nbytes: int, cuda: bool) -> _Slab:
        size = 1 << max(20, (nbytes - 1).bit_length())
        lst = self._pool.get(size)
        if lst:
            return lst.pop()
        return _Slab(size, cuda)

    def _reap(self, block: bool) -> None:
        while self._inflight:
            try:
                job_id, err = self._done.get(block, timeout=600 if block else None)
            except queue.Empty:
                return
            slab, nbytes = self._inflight.pop(job_id)
            self._inflight_bytes -= nbytes
            self.bytes_written += nbytes
            self._pool.setdefault(slab.nbytes, []).append(slab)
            if err and self._err is None:
                self._err = err
            if block and self._inflight_bytes <= self.max_inflight // 2:
                block = False

    def _stage(self, state: Any, cuda_dev: Optional[torch.device]) -> Tuple[Any, Optional[_Slab], int, Any]:
        """Copy every tensor of ``state`` into one shared slab; returns the mirrored structure of CPU views."""
        tensors: List[torch.Tensor] = []

        def collect(o):
            if isinstance(o, torch.Tensor):
                tensors.append(o)
            elif isinstance(o, dict):
                for v in o.values():
                    collect(v)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    collect(v)
        collect(state)
        total = sum((t.numel() * t.element_size() + 63) // 64 * 64 for t in tensors)
        if total == 0:
            return state, None, 0, None
        slab = self._get_slab(total, cuda_dev is not None)
        event = None
        offset = 0
        views = {}
        ctx = torch.cuda.stream(self._copy_stream) if cuda_dev is not None else _null()
        if cuda_dev is not None:
            self._copy_stream.wait_stream(torch.cuda.current_stream(cuda_dev))
        with ctx:
            for t in tensors:
                nb = t.numel() * t.element_size()
                v = slab.buf[offset:offset + nb].view(t.dtype).view(t.shape)
                src = t.detach()
                if not src.is_contiguous():
                    src = src.contiguous()
                v.copy_(src, non_blocking=True)
                if src.is_cuda:
                    src.record_stream(self._copy_stream)
                views[id(t)] = v
                offset += (nb + 63) // 64 * 64
            if cuda_dev is not None:
                event = torch.cuda.Event()
                event.record(self._copy_stream)
                self._last_copy_event = event

        def mirror(o):
            if isinstance(o, torch.Tensor):
                return views[id(o)]
            if isinstance(o, dict):
                return {k: mirror(v) for k, v in o.items()}
            if isinstance(o, (list, tuple)):
                return type(o)(mirror(v) for v in o)
            return o
        return mirror(state), slab, total, event

    def fence(self) -> None:
        """Make the compute stream wait for outstanding snapshot copies (call before sources are overwritten)."""
        with self._lock:
            if self._last_copy_event is not None:
                torch.cuda.current_stream().wait_event(self._last_copy_event)
                self._last_copy_event = None

    # ------------------------------------------------------------------ public API
    def save(self, actor: str, state_name: Optional[str], state: Any, cover: bool = False,
             post: Optional[str] = None) -> None:
        """``save_state`` of ``modules/client.py:52-63`` / ``modules/server.py:46-57``."""
        if state_name is None or not self.enabled:
            return
        with self._lock:                                  # client threads checkpoint concurrently
            self._save_locked(actor, state_name, state, cover, post)

    def _save_locked(self, actor: str, state_name: str, state: Any, cover: bool, post: Optional[str]) -> None:
        self._raise_pending()
        path = self.path(actor, state_name)
        if cover is False and os.path.exists(path):
            raise ValueError(f"State checkpoint has already exist in '{path}'.")
        if not self.asynchronous:
            snap = _to_cpu(state)
            if post:
                snap = POST[post](snap)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(snap, path + ".tmp")
            os.replace(path + ".tmp", path)
            return
        if not self._started:
            self._start()
        cuda_dev = None

        def find_dev(o):
            nonlocal cuda_dev
            if isinstance(o, torch.Tensor):
                if o.is_cuda and cuda_dev is None:
                    cuda_dev = o.device
            elif isinstance(o, dict):
                for v in o.values():
                    find_dev(v)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    find_dev(v)
        find_dev(state)
        if cuda_dev is not None and self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(cuda_dev)
        self._reap(block=False)
        if self._inflight_bytes > self.max_inflight:
            self._reap(block=True)                        # back-pressure: the writers cannot keep up
        snap, slab, nbytes, event = self._stage(state, cuda_dev)
        job_id = self._next_id
        self._next_id += 1
        if slab is None:
            slab = self._get_slab(1, False)               # tensor-free payload: placeholder for the bookkeeping
        self._inflight[job_id] = (slab, nbytes)
        self._inflight_bytes += nbytes
        self._feed_q.put((event, (job_id, path, snap, post)))

    def load(self, actor: str, state_name: str, default_value: Any = None, map_location: str = "cpu") -> Any:
        """``load_state`` (``modules/client.py:34-50``): returns ``default_value`` when the file does not exist."""
        self.flush()
        path = self.path(actor, state_name)
        if os.path.exists(path):
            return torch.load(path, map_location=map_location, weights_only=False)
        if default_value is not None:
            return default_value
        raise ValueError(f"State checkpoint does not exist in '{path}'.")

    def _raise_pending(self) -> None:
        if self._err is not None:
            err, self._err = self._err, None
            raise RuntimeError(f"checkpoint writer failed: {err}")

    def flush(self) -> None:
        with self._lock:
            if self._started:
                while self._inflight:
                    self._reap(block=True)
            self._raise_pending()

    def close(self) -> None:
        if self._started:
            self.flush()
            self._feed_q.put(None)
            self._feeder.join()
            for _ in self._procs:
                self._jobs.put(None)
            for p in self._procs:
                p.join(timeout=30)
            for lst in self._pool.values():
                for slab in lst:
                    slab.release()
            self._pool.clear()
            self._started = False
        self.asynchronous = False


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
