"""ReID datasets (``datasets/datasets_loader.py:10-43``).

``ReIDImageDataset`` reads either an ``ImageFolder`` tree whose directory names are integer person ids, or an
in-memory ``{person_id: [(array, class_id), ...]}`` dict (exemplars / prototypes). Items are
``(data, person_id, class_index)``; person ids are used directly as class indices of the ``num_classes`` head.

``ArrayReIDDataset`` is the device-pipeline form: the whole split decoded once into a pinned uint8 NHWC tensor.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from .augmentation import augmentations

IMG_EXT = (".jpg", ".jpeg", ".png", ".bmp", ".ppm", ".webp")


def _scan_folder(root: str) -> Tuple[List[str], List[Tuple[str, int]]]:
    classes = sorted(d.name for d in os.scandir(root) if d.is_dir())
    samples: List[Tuple[str, int]] = []
    for ci, cname in enumerate(classes):
        for dirpath, _, files in sorted(os.walk(os.path.join(root, cname))):
            for fn in sorted(files):
                if fn.lower().endswith(IMG_EXT):
                    samples.append((os.path.join(dirpath, fn), ci))
    return classes, samples


class ReIDImageDataset(Dataset):
    def __init__(self, source: Union[str, Dict], transform: Optional[Callable] = None):
        super().__init__()
        self.reload_source(source, transform if transform is not None else augmentations["none"]())

    def reload_source(self, source, transform: Optional[Callable] = None) -> None:
        self.transform = transform
        if isinstance(source, str):
            classes, samples = _scan_folder(source)
            self.samples = samples
            self.dataset = None
            self.classes = [int(c) for c in classes]
        elif isinstance(source, dict):
            self.dataset = []
            self.classes = {}
            for person_id, protos in source.items():
                for img, class_id in protos:
                    self.dataset.append((img, class_id))
                    self.classes[class_id] = person_id
            self.samples = None
        else:
            raise ValueError("Input source should be path in disk or dictionary in memory.")

    @property
    def person_ids(self):
        return self.classes

    def __getitem__(self, index) -> Any:
        if self.samples is not None:
            from PIL import Image
            path, class_index = self.samples[index]
            with Image.open(path) as im:
                data = im.convert("RGB")
            if self.transform is not None:
                data = self.transform(data)
        else:
            data, class_index = self.dataset[index]
            if not isinstance(data, torch.Tensor):
                data = torch.as_tensor(np.asarray(data), dtype=torch.float32)
        class_index = int(class_index)
        return data, int(self.classes[class_index]), class_index

    def __len__(self) -> int:
        return len(self.samples) if self.samples is not None else len(self.dataset)


class ArrayReIDDataset(Dataset):
    """A split held as one uint8 ``[N,H,W,3]`` tensor (pinned when CUDA is present) + int64 id vectors."""

    def __init__(self, images_u8: torch.Tensor, person_ids: torch.Tensor, class_index: Optional[torch.Tensor] = None,
                 classes: Optional[Sequence[int]] = None, pin: bool = True):
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[-1] == 3
        if pin and torch.cuda.is_available() and not images_u8.is_pinned():
            images_u8 = images_u8.pin_memory()
        self.images = images_u8
        self.pids = person_ids.long()
        self.classes = list(classes) if classes is not None else sorted(set(self.pids.tolist()))
        if class_index is None:
            lut = {p: i for i, p in enumerate(self.classes)}
            class_index = torch.tensor([lut[int(p)] for p in self.pids.tolist()], dtype=torch.long)
        self.cidx = class_index.long()

    @property
    def person_ids(self):
        return self.classes

    def __len__(self) -> int:
        return self.images.shape[0]

    def __getitem__(self, i):
        return self.images[i], int(self.pids[i]), int(self.cidx[i])

    @staticmethod
    def from_folder(root: str, size: Sequence[int], workers: int = 8) -> "ArrayReIDDataset":
        """Decode + resize an ImageFolder split once (the reference re-decodes every image every epoch)."""
        from PIL import Image
        classes, samples = _scan_folder(root)
        h, w = int(size[0]), int(size[1])

        def load(item):
            path, _ = item
            with Image.open(path) as im:
                im = im.convert("RGB")
                if im.size != (w, h):
                    im = im.resize((w, h), Image.BILINEAR)
                return np.asarray(im, dtype=np.uint8)
        if samples:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                arrs = list(pool.map(load, samples))
            images = torch.from_numpy(np.stack(arrs))
        else:
            images = torch.zeros((0, h, w, 3), dtype=torch.uint8)
        cls_int = [int(c) for c in classes]
        cidx = torch.tensor([ci for _, ci in samples], dtype=torch.long)
        pids = torch.tensor([cls_int[ci] for _, ci in samples], dtype=torch.long)
        return ArrayReIDDataset(images, pids, cidx, cls_int)
