"""Per-client lifelong task pipeline (``datasets/datasets_pipeline.py:10-93``).

Semantics kept: ordered task list, ``sustain_rounds`` rounds per task, ``next_task()`` advances and then stays on the
last task forever, ``get_task(idx)`` returns ``{task_name, tr_epochs, tr_loader, query_loader, gallery_loaders}``.
Difference: splits are loaded **once** and cached (the reference rebuilds 3 datasets + 3 DataLoaders on every call),
and by default the loaders are :class:`DeviceBatchLoader` (pinned uint8 -> async H2D -> GPU augmentation).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

from ..ops import native as _native
from torch.utils.data import DataLoader

from .augmentation import DeviceAugment, augmentations
from .datasets import ArrayReIDDataset, ReIDImageDataset


class DeviceBatchLoader:
    """Iterates ``(data, person_id, class_index)`` batches; data lands on ``device`` already augmented.

    Host side: index gather into a pinned staging buffer. Device side: async copy on a dedicated stream,
    double-buffered so the copy of batch i+1 overlaps the compute of batch i.
    """

    def __init__(self, dataset: ArrayReIDDataset, batch_size: int, shuffle: bool, level: str, mean, std,
                 device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32, drop_last: bool = False,
                 seed: Optional[int] = None):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.device = torch.device(device)
        self.augment = DeviceAugment(level, mean, std, dtype)
        self.num_workers = 0
        self.pin_memory = True
        self.persistent_workers = False
        self.multiprocessing_context = None
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(seed)
        self._stage: List[torch.Tensor] = []
        self._copy_stream = None
        self.h2d_bytes = 0
        self._labels_dev = None
        self.bulk_limit_bytes = 2 << 30
        self.device_generator: Optional[torch.Generator] = None     # CUDA generator of the owning model (if any)

    def to(self, device, dtype: Optional[torch.dtype] = None) -> "DeviceBatchLoader":
        self.device = torch.device(device)
        if dtype is not None:
            self.augment.dtype = dtype
        return self

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def fetch(self, indices: torch.Tensor):
        """Augmented device batch for arbitrary dataset indices (rehearsal methods mix exemplars and task data)."""
        ds = self.dataset
        idx = indices.cpu()
        u8 = ds.images[idx]
        if self.device.type == "cuda":
            u8 = u8.pin_memory().to(self.device, non_blocking=True)
            self.h2d_bytes += u8.numel()
        return self.augment(u8, self._gen_for(u8.device)), ds.pids[idx].to(self.device), ds.cidx[idx].to(self.device)

    def _order(self) -> torch.Tensor:
        n = len(self.dataset)
        return torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)

    def _bulk_ok(self) -> bool:
        ds = self.dataset
        return (self.device.type == "cuda" and ds.images.is_pinned()
                and ds.images.numel() * ds.images.element_size() <= self.bulk_limit_bytes)

    def iterate(self, batch_size: Optional[int] = None, ordered: bool = False):
        """Like ``iter(self)`` with a different batch size (the inference-only trunk wants wider batches).
        ``ordered`` skips the shuffle (the caller does not care about the sample order)."""
        return self._iter_bulk(batch_size, ordered) if self._bulk_ok() else self._iter_staged(batch_size, ordered)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        return self.iterate()

    def _iter_bulk(self, batch_size: Optional[int], ordered: bool):
        """Pinned split -> ONE async H2D copy per epoch on the copy stream (no host-side gather, no per-batch
        staging); shuffling and batching are index ops on the device; augmentation is one fused kernel per batch."""
        ds, dev = self.dataset, self.device
        bs = int(batch_size or self.batch_size)
        n = len(ds)
        if self._copy_stream is None:
            self._copy_stream = _native.dedicated_stream(dev)
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self._copy_stream):
            dev_all = ds.images.to(dev, non_blocking=True)
            if self._labels_dev is None:
                self._labels_dev = (ds.pids.to(dev, non_blocking=True), ds.cidx.to(dev, non_blocking=True))
                self.h2d_bytes += 16 * n
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self.h2d_bytes += dev_all.numel()
        pids, cidx = self._labels_dev
        shuffle = self.shuffle and not ordered
        order = self._order().to(dev, non_blocking=True) if shuffle else None
        cur.wait_event(ev)
        dev_all.record_stream(cur)
        last = n // bs * bs if (self.drop_last and batch_size is None) else n
        for s0 in range(0, last, bs):
            if shuffle:
                idx = order[s0:s0 + bs]
                yield self.augment(dev_all.index_select(0, idx), self._gen_for(dev)), pids[idx], cidx[idx]
            else:
                yield self.augment(dev_all[s0:s0 + bs], self._gen_for(dev)), pids[s0:s0 + bs], cidx[s0:s0 + bs]

    def _iter_staged(self, batch_size: Optional[int] = None, ordered: bool = False):
        ds = self.dataset
        order = self._order() if not ordered else torch.arange(len(ds))
        if batch_size is not None and int(batch_size) != self.batch_size:
            saved = self.batch_size
            self.batch_size, self._stage = int(batch_size), []
            try:
                yield from self._iter_staged(None, ordered)
            finally:
                self.batch_size, self._stage = saved, []
            return
        nb = len(self)
        cuda = self.device.type == "cuda"
        if cuda:
            if self._copy_stream is None:
                self._copy_stream = _native.dedicated_stream(self.device)
            if not self._stage:
                shape = (self.batch_size,) + tuple(ds.images.shape[1:])
                self._stage = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
                self._stage_ev = [None, None]

        def stage(bi: int):
            idx = order[bi * self.batch_size:(bi + 1) * self.batch_size]
            if not cuda:
                return ds.images[idx], ds.pids[idx], ds.cidx[idx], None
            buf = self._stage[bi % 2]
            if self._stage_ev[bi % 2] is not None:
                self._stage_ev[bi % 2].synchronize()          # previous copy out of this buffer finished
            torch.index_select(ds.images, 0, idx, out=buf[:len(idx)])
            with torch.cuda.stream(self._copy_stream):
                dev_u8 = buf[:len(idx)].to(self.device, non_blocking=True)
                pid = ds.pids[idx].to(self.device, non_blocking=True)
                cid = ds.cidx[idx].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._stage_ev[bi % 2] = ev
            self.h2d_bytes += dev_u8.numel() + 16 * len(idx)
            return dev_u8, pid, cid, ev

        nxt = stage(0) if nb else None
        for bi in range(nb):
            cur = nxt
            nxt = stage(bi + 1) if bi + 1 < nb else None
            u8, pid, cid, ev = cur
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                u8.record_stream(torch.cuda.current_stream(self.device))
            yield self.augment(u8, self._gen_for(u8.device)), pid, cid

    def _gen_for(self, device) -> Optional[torch.Generator]:
        g = self.device_generator
        return g if g is not None and g.device == torch.device(device) else None


class ReIDTaskPipeline:
    def __init__(self, task_list: List[str], task_opts: Dict, datasets_dir: str, device_loader: bool = True,
                 source_factory: Optional[Callable[[str, str], ArrayReIDDataset]] = None):
        self.task_list = task_list
        self.task_opts = task_opts
        self.datasets_dir = datasets_dir
        self.current_task_idx = -1
        self.task_round_rest = [task_opts["sustain_rounds"] for _ in task_list]
        self.device_loader = device_loader
        self.source_factory = source_factory      # (task_name, split) -> ArrayReIDDataset (synthetic / in-memory)
        self._cache: Dict[int, Dict] = {}

    # ---- reference schedule (datasets_pipeline.py:19-20,81-93) -------------------------------------------------
    def reach_final_task(self) -> bool:
        return self.current_task_idx + 1 == len(self.task_list)

    def current_task(self) -> Dict:
        if self.current_task_idx == -1:
            self.current_task_idx = 0
        return self.get_task(self.current_task_idx)

    def next_task(self) -> Dict:
        if not self.reach_final_task():
            if self.current_task_idx != -1 and self.task_round_rest[self.current_task_idx]:
                self.task_round_rest[self.current_task_idx] -= 1
            else:
                self.current_task_idx += 1
                self.task_round_rest[self.current_task_idx] -= 1
        return self.current_task()

    def state_dict(self) -> Dict:
        """Position in the task schedule + the shuffle generators of the cached train loaders (a task that sustains over
        several rounds continues ITS permutation sequence: a run resumed in the middle of a task must too)."""
        rng = {int(k): v for k, v in getattr(self, "_pending_rng", {}).items()}
        for idx, task in self._cache.items():
            gen = getattr(task["tr_loader"], "_gen", None)
            if isinstance(gen, torch.Generator):
                rng[int(idx)] = gen.get_state()
        return {"current_task_idx": self.current_task_idx, "task_round_rest": list(self.task_round_rest),
                "loader_rng": rng}

    def load_state_dict(self, sd: Dict) -> None:
        self.current_task_idx = int(sd["current_task_idx"])
        self.task_round_rest = list(sd["task_round_rest"])
        self._pending_rng = {int(k): v for k, v in (sd.get("loader_rng") or {}).items()}
        for idx in list(self._cache):
            self._restore_loader_rng(idx)

    def _restore_loader_rng(self, idx: int) -> None:
        st = getattr(self, "_pending_rng", {}).pop(idx, None)
        gen = getattr(self._cache[idx]["tr_loader"], "_gen", None)
        if st is not None and isinstance(gen, torch.Generator):
            gen.set_state(st.cpu() if isinstance(st, torch.Tensor) else st)

    # ---- loaders ----------------------------------------------------------------------------------------------
    def _split(self, task: str, split: str):
        aug = self.task_opts["augment_opts"]
        if self.source_factory is not None:
            return self.source_factory(task, split)
        path = os.path.join(self.datasets_dir, task, split)
        if self.device_loader:
            return ArrayReIDDataset.from_folder(path, aug["img_size"])
        level = aug["level"] if split == "train" else "none"
        return ReIDImageDataset(path, augmentations[level](size=aug["img_size"], mean=aug["norm_mean"],
                                                           std=aug["norm_std"]))

    def _loader(self, ds, train: bool):
        aug, lo = self.task_opts["augment_opts"], self.task_opts["loader_opts"]
        bs = lo["batch_size"]
        drop_last = len(ds) % bs == 1                                    # datasets_pipeline.py:39
        if isinstance(ds, ArrayReIDDataset):
            return DeviceBatchLoader(ds, bs, shuffle=train, level=aug["level"] if train else "none",
                                     mean=aug["norm_mean"], std=aug["norm_std"], drop_last=drop_last)
        workers = lo.get("num_workers", 0)
        return DataLoader(ds, shuffle=train, drop_last=drop_last, batch_size=bs, num_workers=workers,
                          pin_memory=lo.get("pin_memory", False),
                          persistent_workers=lo.get("persistent_workers", False) and workers > 0,
                          multiprocessing_context=lo.get("multiprocessing_context") if workers > 0 else None)

    def get_task(self, idx: int = -1) -> Dict:
        idx = idx % len(self.task_list)
        if idx not in self._cache:
            task = self.task_list[idx]
            self._cache[idx] = {
                "task_name": task,
                "tr_epochs": self.task_opts["train_epochs"],
                "tr_loader": self._loader(self._split(task, "train"), True),
                "query_loader": self._loader(self._split(task, "query"), False),
                "gallery_loaders": self._loader(self._split(task, "gallery"), False),
            }
            self._restore_loader_rng(idx)              # (a resumed run: continue the saved shuffle sequence)
        return self._cache[idx]

    def evict(self, keep: Optional[int] = None) -> None:
        """Drop cached splits (except ``keep``) – host-RAM control for long task lists."""
        for k in list(self._cache):
            if k != keep:
                del self._cache[k]
