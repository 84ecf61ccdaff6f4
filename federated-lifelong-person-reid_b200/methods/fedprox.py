"""``fedprox`` – FedAvg + proximal term ``lambda_l2 * sum (p - p_old)^2`` (reference ``methods/fedprox.py``).

Reference quirk (kept under ``reference_compat``): ``remember_params()`` runs *before* the incoming global model is
applied, so ``p_old`` is the previous **local** weights, not the global ones (``fedprox.py:351-352,363-364``).
The proximal gradient is fused into the optimizer kernel (Q == 1 is not materialised)."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..runtime.modules import ModelModule, OperatorModule
from .fedbase import FedClient, FedServer


class Model(ModelModule):
    resume_attrs = ("p_old", "has_old")
    def __init__(self, net, lambda_l2: float = 1e-2, **kwargs):
        super().__init__(net, **kwargs)
        self.lambda_l2 = float(lambda_l2)
        self.p_old: Optional[torch.Tensor] = None
        self.has_old = False

    def materialize(self, device, compute_dtype="bf16", fine_tuning=None):
        super().materialize(device, compute_dtype, fine_tuning)
        self.p_old = self.arena.master.clone()
        return self

    def remember_params(self) -> None:
        self.p_old.copy_(self.arena.master)
        self.has_old = True

    def install(self, optimizer) -> None:
        optimizer.Q, optimizer.R = None, self.p_old
        optimizer.lam2, optimizer.penalty_ones = self.lambda_l2, True

    def penalty(self) -> torch.Tensor:
        return self.lambda_l2 * ((self.arena.master - self.p_old) ** 2).sum()

    def model_state(self) -> Dict:
        return {"net_params": {k: v.detach().clone(memory_format=torch.contiguous_format)
                               for k, v in self.net.state_dict().items()},
                "params_old": self.arena.to_dict(self.p_old) if self.has_old else {}}

    def update_model(self, params_state: Dict) -> None:
        if "net_params" in params_state:
            own = self.net.state_dict()
            with torch.no_grad():
                for k, v in params_state["net_params"].items():
                    if k in own:
                        own[k].copy_(v.to(own[k].device))
            self.arena.refresh_shadow()


class Operator(OperatorModule):
    pass


class Client(FedClient):
    default_ckpt_name = "fedprox_model"
    payload_prefix = ""

    def before_global_update(self) -> None:
        if getattr(self, "reference_compat", True):
            self.model.remember_params()                  # previous *local* weights (reference behaviour)

    def update_by_incremental_state(self, state, **kwargs):
        if "_flat" in state:
            # device-resident dispatch: global -> master, bf16 copy and the proximal anchor in ONE pass (C2 fusion);
            # anchor = the weights being replaced (reference order of operations) or the incoming global model
            self.train_cnt = self.test_cnt = 0
            self.apply_global(state["_flat"], self.model.p_old, 1 if getattr(self, "reference_compat", True) else 2)
            self.model.has_old = True
            self.logger.info("Update model succeed by incremental state from server.")
            return
        super().update_by_incremental_state(state, **kwargs)
        if not getattr(self, "reference_compat", True):
            self.model.remember_params()                  # textbook FedProx: anchor = incoming global model

    def update_by_integrated_state(self, state, **kwargs):
        super().update_by_integrated_state(state, **kwargs)
        if not getattr(self, "reference_compat", True):
            self.model.remember_params()

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        self.model.install(self.operator.optimizer)


class Server(FedServer):
    payload_prefix = ""
    pass
