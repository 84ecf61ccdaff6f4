"""``ewc`` – Elastic Weight Consolidation, local-only continual learning (reference ``methods/ewc.py``).

Diagonal Fisher over the *previous* tasks' train loaders (the current one is skipped), ``penalty = lam * sum F
(p - p_old)^2``, ``remember_task(task, tr_loader)`` after every ``train()``. Reference quirk kept under
``reference_compat``: the server's first-contact dispatch is a silent no-op (the client expects ``net_params``)."""
from __future__ import annotations

from typing import Any, Dict

from ..runtime.modules import ClientModule, OperatorModule, ServerModule
from .penalty import PenaltyModel


class Model(PenaltyModel):
    importance_mode = "fisher"
    skip_current_task = True


class Operator(OperatorModule):
    pass


class Client(ClientModule):
    default_ckpt_name = "ewc_model"
    remember_split = "train"

    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self.model.operator = operator
        if not self.model_ckpt_name:
            self.model_ckpt_name = self.default_ckpt_name

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        self.model.install(self.operator.optimizer)

    def after_train(self, task_name, tr_loader, val_loader, output) -> None:
        self.model.remember_task(task_name, tr_loader if self.remember_split == "train" else val_loader)

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        if getattr(self, "reference_compat", True) and "net_params" not in state.get("model_params", {}):
            self.logger.info("Update model succeed by integrated state from server.")   # silent no-op (ewc.py:373)
            return
        self.model.update_model({"net_params": state["model_params"]})

    update_by_incremental_state = update_by_integrated_state


class Server(ServerModule):
    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        return {"model_params": self.model.full_state()}
