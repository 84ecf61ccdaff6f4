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
    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self._snapshots: Dict[str, Dict[str, torch.Tensor]] = {}
        self._active: str | None = None

    def _activate(self, name: str) -> None:
        """Make the weights of checkpoint ``name`` resident (``load_model(name)`` with default = current weights)."""
        if self._active == name:
            return
        if self._active is not None:
            self._snapshots[self._active] = self.model.full_state()
        if name in self._snapshots:
            self.model.load_full_state(self._snapshots[name])
        elif self.store.exists(self.name, name):
            self.model.update_model(self.load_state(name))
        self._active = name

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        self.model.load_full_state(state["model_params"])
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
