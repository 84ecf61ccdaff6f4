"""``baseline`` – plain fine-tuning (configs ``sm`` = one shared model, ``mm`` = one model per task).

Reference: ``methods/baseline.py``. No upload, no aggregation; first contact dispatches the server's full
``state_dict``. With ``model_ckpt_name`` unset the reference keeps one checkpoint *per task* and reloads it around
every train / validate call; here the per-task weights are device-resident snapshots swapped in on demand.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from ..runtime.modules import ClientModule, OperatorModule, ServerModule


class Operator(OperatorModule):
    pass


class Client(ClientModule):
    """File semantics of the reference, kept on the device (``methods/baseline.py:214-336``, ``modules/client.py:63-70``):
    ``load_model(name)`` restores checkpoint ``name`` *if it exists* and is a no-op otherwise - so a task that was never
    trained is evaluated with, and trained from, whatever weights are resident (sequential fine-tuning) - and a
    checkpoint is created or refreshed only by ``save_model`` (end of ``train`` and, with ``model_ckpt_name`` set,
    after a dispatch). A snapshot here is the device-resident twin of such a file: it exists iff the file would."""

    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self._snapshots: Dict[str, Dict[str, torch.Tensor]] = {}
        self._on_disk: Dict[str, bool] = {}          # files of an earlier run (the reference would pick them up too)
        self._active: str | None = None              # name of the snapshot the resident weights are known to equal

    def _activate(self, name: str) -> None:
        """``load_model(name)``: make checkpoint ``name`` resident if it exists, else leave the weights untouched."""
        if name is None or self._active == name:
            return
        if name in self._snapshots:
            self.model.load_full_state(self._snapshots[name])
            self._active = name
            return
        if name not in self._on_disk:
            self._on_disk[name] = bool(self.store.enabled and self.store.exists(self.name, name))
        if self._on_disk[name]:
            self.model.update_model(self.load_state(name))
            self._snapshots[name] = self.model.full_state()
            self._active = name

    def save_model(self, model_name: str) -> None:
        if model_name is None:
            return
        self._snapshots[model_name] = self.model.full_state()
        self._active = model_name
        super().save_model(model_name)

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        self._activate(self.model_ckpt_name)          # load_model(model_ckpt_name); `None` (mm) never exists
        self.model.load_full_state(state["model_params"])
        self._active = None
        if self.model_ckpt_name:                      # save_model(model_ckpt_name): the dispatch becomes the checkpoint
            self._snapshots[self.model_ckpt_name] = self.model.full_state()
            self._active = self.model_ckpt_name
        self.logger.info("Update model succeed by integrated state from server.")

    update_by_incremental_state = update_by_integrated_state

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        self._activate(self.ckpt_name(task_name))

    def validate(self, task_name, query_loader, gallery_loader, device="cpu", **kwargs):
        self._activate(self.ckpt_name(task_name))
        return super().validate(task_name, query_loader, gallery_loader, device, **kwargs)

    def inference(self, task_name, query_loader, gallery_loader, device="cpu", **kwargs):
        self._activate(self.ckpt_name(task_name))
        return super().inference(task_name, query_loader, gallery_loader, device, **kwargs)


class Server(ServerModule):
    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        return {"model_params": self.model.full_state()}
