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


class PersistentState(dict):
    """A state dict whose structure AND tensor objects stay the same from one ``save`` to the next (views of
    long-lived device buffers; only their contents change). ``plan_token`` identifies that binding: a store may cache
    whatever it derived from the structure (byte layout, copy plan) for as long as the token is unchanged."""
    plan_token: Any = None


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


_TREF = "__flpr_tensor_ref__"


def _resolve(obj: Any, arena_np) -> Any:
    """Rebuild the tensors of a staged state from ``(offset, shape, dtype)`` references into the shared arena.
    ``torch.frombuffer`` gives every tensor a storage of exactly its own size, so ``torch.save`` writes just the
    payload bytes (a view of the arena tensor would serialise the whole arena) without an intermediate copy."""
    if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == _TREF:
        _, off, shape, dtype = obj
        dt = getattr(torch, dtype)
        n = 1
        for d in shape:
            n *= d
        if n == 0:
            return torch.empty(shape, dtype=dt)
        nb = n * torch.empty(0, dtype=dt).element_size()
        return torch.frombuffer(arena_np[off:off + nb], dtype=dt).view(shape)
    if isinstance(obj, dict):
        return {k: _resolve(v, arena_np) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_resolve(v, arena_np) for v in obj)
    return obj


def _writer_main(jobs, done, arena) -> None:  # pragma: no cover  (runs in a child process)
    torch.set_num_threads(4)                 # bf16 -> fp32 expansion of the exemplar file
    arena_np = arena.numpy() if arena is not None else None
    while True:
        job = jobs.get()
        if job is None:
            return
        job_id, path, state, post = job
        err = None
        try:
            state = _resolve(state, arena_np)
            if post:
                state = POST[post](state)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = f"{path}.tmp{os.getpid()}"
            # legacy (non-zip) container: no per-record CRC32 pass over hundreds of MB; torch.load reads both formats.
            # pickle protocol 5: numpy exemplar arrays go out as zero-copy in-band buffers (protocol 2 escapes them
            # byte by byte: 1.6 s instead of 0.11 s for one client's 268 MB exemplar file); Python >= 3.8 loads it.
            torch.save(state, tmp, _use_new_zipfile_serialization=False, pickle_protocol=5)
            os.replace(tmp, path)
        except BaseException as ex:
            err = f"{type(ex).__name__}: {ex}"
        del state, job
        done.put((job_id, err))


class _SharedArena:
    """ONE process-shared, page-locked staging arena, mapped by every writer process at start-up. Regions are handed
    out first-fit and returned when the writer is done, so a checkpoint costs no shared-memory creation, no
    ``cudaHostRegister`` and no file-descriptor passing on the training thread (jobs carry offsets only)."""

    ALIGN = 4096

    def __init__(self, nbytes: int, cuda: bool):
        self.nbytes = int(nbytes)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8).share_memory_()
        self.registered = False
        if cuda:
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), self.nbytes, 0)
            self.registered = int(rc) == 0
        self.free: List[Tuple[int, int]] = [(0, self.nbytes)]          # sorted (offset, size)

    def alloc(self, n: int) -> int:
        n = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        for i, (off, size) in enumerate(self.free):
            if size >= n:
                if size == n:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + n, size - n)
                return off
        return -1

    def release_region(self, off: int, n: int) -> None:
        n = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        import bisect
        i = bisect.bisect_left(self.free, (off, 0))
        self.free.insert(i, (off, n))
        if i + 1 < len(self.free) and self.free[i][0] + self.free[i][1] == self.free[i + 1][0]:
            self.free[i] = (self.free[i][0], self.free[i][1] + self.free[i + 1][1])
            self.free.pop(i + 1)
        if i > 0 and self.free[i - 1][0] + self.free[i - 1][1] == self.free[i][0]:
            self.free[i - 1] = (self.free[i - 1][0], self.free[i - 1][1] + self.free[i][1])
            self.free.pop(i)

    def release(self) -> None:
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self.registered = False


class CheckpointStore:
    """One instance per process; actors address it with ``(actor_name, state_name)``."""

    def __init__(self, root: str, asynchronous: bool = True, enabled: bool = True, workers: int = 12,
                 arena_bytes: int = 8 << 30):
        self.root = root
        self.enabled = enabled
        self.asynchronous = bool(asynchronous and enabled)
        self.workers = max(1, int(workers))
        self.arena_bytes = int(arena_bytes)
        self.bytes_written = 0
        self.muted = False                                # engine_opts.checkpoint_interval: skip this round's files
        self._started = False
        self._err: Optional[str] = None
        self._next_id = 0
        self._inflight: Dict[int, Tuple[int, int]] = {}      # job -> (arena offset, bytes)
        self._arena: Optional[_SharedArena] = None
        self._lock = threading.RLock()
        self._copy_stream = None
        self._last_copy_event = None
        self._actor_events: Dict[str, Any] = {}              # actor -> event of its latest staged snapshot

    # ------------------------------------------------------------------ paths
    def path(self, actor: str, state_name: str) -> str:
        return os.path.join(self.root, actor, f"{state_name}.ckpt")

    def exists(self, actor: str, state_name: str) -> bool:
        self.flush()
        return os.path.exists(self.path(actor, state_name))

    # ------------------------------------------------------------------ pipeline
    def _start(self, cuda: bool) -> None:
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        self._arena = _SharedArena(self.arena_bytes, cuda)
        self._jobs = ctx.Queue()
        self._done = ctx.Queue()
        self._procs = [ctx.Process(target=_writer_main, args=(self._jobs, self._done, self._arena.buf), daemon=True)
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

    def _reap(self, block: bool) -> None:
        """Collect finished jobs (at least one when ``block``) and return their arena regions."""
        waited = 0.0
        while self._inflight:
            try:
                job_id, err = self._done.get(block, timeout=2.0 if block else None)
            except queue.Empty:
                if not block:
                    return
                waited += 2.0
                if not all(p.is_alive() for p in self._procs):
                    raise RuntimeError("a checkpoint writer process died")
                if waited > 600:
                    raise RuntimeError("checkpoint writers made no progress for 600 s")
                continue
            off, nbytes = self._inflight.pop(job_id)
            if nbytes:
                self._arena.release_region(off, nbytes)
            self.bytes_written += nbytes
            if err and self._err is None:
                self._err = err
            block = False

    def _stage(self, state: Any, cuda_dev: Optional[torch.device]) -> Tuple[Any, int, int, Any]:
        """Copy every tensor of ``state`` into one arena region; returns the structure with tensor references."""
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
        if total > self._arena.nbytes:
            return None, -1, total, None                    # larger than the whole arena: caller saves in-line
        base = 0
        event = None
        if total:
            base = self._arena.alloc(total)
            while base < 0:                                   # back-pressure: the writers cannot keep up
                if not self._inflight:
                    return None, -1, total, None
                self._reap(block=True)
                base = self._arena.alloc(total)
        offset = base
        refs = {}
        buf = self._arena.buf
        ctx = torch.cuda.stream(self._copy_stream) if cuda_dev is not None else _null()
        if cuda_dev is not None:
            self._copy_stream.wait_stream(torch.cuda.current_stream(cuda_dev))
        with ctx:
            for t in tensors:
                nb = t.numel() * t.element_size()
                if nb:
                    v = buf[offset:offset + nb].view(t.dtype).view(t.shape)
                    src = t.detach()
                    if not src.is_contiguous():
                        src = src.contiguous()
                    v.copy_(src, non_blocking=True)
                    if src.is_cuda:
                        src.record_stream(self._copy_stream)
                refs[id(t)] = (_TREF, offset, tuple(t.shape), str(t.dtype).replace("torch.", ""))
                offset += (nb + 63) // 64 * 64
            if cuda_dev is not None and total:
                event = torch.cuda.Event()
                event.record(self._copy_stream)
                self._last_copy_event = event

        def mirror(o):
            if isinstance(o, torch.Tensor):
                return refs[id(o)]
            if isinstance(o, dict):
                return {k: mirror(v) for k, v in o.items()}
            if isinstance(o, (list, tuple)):
                return type(o)(mirror(v) for v in o)
            return o
        return mirror(state), base, total, event

    def fence(self, actor: Optional[str] = None) -> None:
        """Make the compute stream wait for outstanding snapshot copies (call before sources are overwritten).
        With ``actor`` only the snapshots staged on behalf of that actor are waited for (copies are issued in order on
        one stream, so this is "everything up to that actor's latest snapshot")."""
        with self._lock:
            if actor is not None:
                ev = self._actor_events.pop(actor, None)
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                return
            if self._last_copy_event is not None:
                torch.cuda.current_stream().wait_event(self._last_copy_event)
                self._last_copy_event = None
                self._actor_events.clear()

    # ------------------------------------------------------------------ public API
    def save(self, actor: str, state_name: Optional[str], state: Any, cover: bool = False,
             post: Optional[str] = None) -> None:
        """``save_state`` of ``modules/client.py:52-63`` / ``modules/server.py:46-57``."""
        if state_name is None or not self.enabled or (self.muted and actor != "_resume"):
            return
        with self._lock:                                  # client threads checkpoint concurrently
            self._save_locked(actor, state_name, state, cover, post)

    def _save_locked(self, actor: str, state_name: str, state: Any, cover: bool, post: Optional[str]) -> None:
        self._raise_pending()
        path = self.path(actor, state_name)
        if cover is False and os.path.exists(path):
            raise ValueError(f"State checkpoint has already exist in '{path}'.")
        if not self.asynchronous:
            self._save_inline(path, state, post)
            return
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
        if not self._started:
            self._start(cuda_dev is not None or torch.cuda.is_available())
        if cuda_dev is not None and self._copy_stream is None:
            self._copy_stream = _dedicated_stream(cuda_dev)
        self._reap(block=False)
        snap, base, nbytes, event = self._stage(state, cuda_dev)
        if event is not None:
            self._actor_events[actor] = event
        if base < 0:                                           # does not fit in the staging arena
            self._save_inline(path, state, post)
            return
        job_id = self._next_id
        self._next_id += 1
        self._inflight[job_id] = (base, nbytes)
        self._feed_q.put((event, (job_id, path, snap, post)))

    @staticmethod
    def _save_inline(path: str, state: Any, post: Optional[str]) -> None:
        snap = _to_cpu(state)
        if post:
            snap = POST[post](snap)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(snap, path + ".tmp")
        os.replace(path + ".tmp", path)

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

    def track(self, future) -> None:
        """A write some actor performs on its own thread (FedSTIL's token history): ``flush`` waits for it too."""
        with self._lock:
            self._external = [f for f in getattr(self, "_external", []) if not f.done()] + [future]

    def flush(self) -> None:
        with self._lock:
            pending, self._external = list(getattr(self, "_external", [])), []
        for f in pending:
            f.result()
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
            if self._arena is not None:
                self._arena.release()
                self._arena = None
            self._started = False
        self.asynchronous = False


def _dedicated_stream(dev):
    from ..ops import native
    return native.dedicated_stream(dev)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
