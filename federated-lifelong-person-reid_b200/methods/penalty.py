"""Quadratic-penalty machinery shared by EWC / MAS / FedProx / FedCurv.

All four regularisers are ``lam * sum_n Q_n p_n^2 - 2 R_n p_n + C`` for per-parameter buffers ``Q`` (curvature) and
``R`` (curvature-weighted anchor):

    EWC / MAS   Q = F                      R = F * p_old                       (ewc.py:80-85, mas.py:78-83)
    FedProx     Q = 1                      R = p_old                           (fedprox.py:52-57)
    FedCurv     Q = F + sum_j F_j          R = F * p_old + sum_j F_j p_j       (fedcurv.py:79-86)

so the gradient ``2*lam*(Q p - R)`` is folded into the fused optimizer kernel and no per-tensor Python loop (nor,
for FedCurv, the 2K parameter-sized copies per client) is ever materialised.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ..ops import fused as fops
from ..runtime.modules import ModelModule


class PenaltyModel(ModelModule):
    resume_attrs = ("F", "p_old", "Q", "R", "const")
    importance_mode = "fisher"          # "fisher": g^2, "mas": |g|
    skip_current_task = False           # EWC skips the most recent remembered loader (ewc.py:62-65)
    lambda_key = "lambda_penalty"
    lambda_default = 100.0

    def __init__(self, net, operator=None, **kwargs):
        super().__init__(net, **kwargs)
        self.operator = operator
        self.lam = float(kwargs.get(self.lambda_key, self.lambda_default))
        setattr(self, self.lambda_key, self.lam)
        self.recall_dataloaders: Dict[str, object] = {}
        self.F: Optional[torch.Tensor] = None           # own importance (flat, arena layout)
        self.p_old: Optional[torch.Tensor] = None
        self.Q: Optional[torch.Tensor] = None
        self.R: Optional[torch.Tensor] = None
        self.const = 0.0

    # ---- arena-dependent state -----------------------------------------------------------------------------------
    def materialize(self, device, compute_dtype="bf16", fine_tuning=None):
        super().materialize(device, compute_dtype, fine_tuning)
        a = self.arena
        self.F = a.new_buffer()
        self.p_old = a.master.clone()
        self.Q = a.new_buffer()
        self.R = a.new_buffer()
        self.rebuild_penalty()
        return self

    def other_terms(self):
        """(sum_j F_j, sum_j F_j p_j) of other clients, or None (FedCurv overrides)."""
        return None

    def rebuild_penalty(self) -> None:
        with torch.no_grad():
            self.Q.copy_(self.F)
            torch.mul(self.F, self.p_old, out=self.R)
            other = self.other_terms()
            if other is not None:
                self.Q.add_(other[0])
                self.R.add_(other[1])

    def install(self, optimizer) -> None:
        optimizer.Q, optimizer.R, optimizer.lam2, optimizer.penalty_ones = self.Q, self.R, self.lam, False

    def penalty(self) -> torch.Tensor:
        """Value of the regulariser (diagnostics / tests; the training step uses the fused gradient)."""
        p = self.arena.master
        val = (self.F * (p - self.p_old) ** 2).sum()
        return self.lam * val

    # ---- importance ----------------------------------------------------------------------------------------------
    def calculate(self) -> torch.Tensor:
        self._calculate_importance()
        self.p_old.copy_(self.arena.master)
        self.rebuild_penalty()
        return self.F

    def _calculate_importance(self) -> None:
        self.F.zero_()
        loaders = list(self.recall_dataloaders.values())
        if self.skip_current_task:
            if len(loaders) <= 1:
                return
            loaders = loaders[:-1]
        if not loaders:
            return
        n_batches = sum(len(ld) for ld in loaders)
        a = self.arena
        for ld in loaders:
            if hasattr(ld, "to") and hasattr(ld, "augment"):
                ld.to(self.device, torch.float32)
                if self.rng is not None:
                    # never the default CUDA generator: another client's thread may have it registered with a CUDA-graph
                    # capture in flight ("Offset increment outside graph capture")
                    ld.device_generator = self.rng
            for data, person_id, _ in ld:
                a.zero_grad()
                data, target = self.prepare_input(data), person_id.to(self.device)
                loss = self.operator._invoke_train(self, data, target)["loss"]
                loss.backward()
                fops.importance_accumulate(self.F, a.grad, float(len(data)) / n_batches, self.importance_mode)
        a.zero_grad()

    def remember_task(self, task_name: str, dataloader) -> None:
        self.recall_dataloaders[task_name] = dataloader
        self.calculate()

    # ---- checkpoint schema {'net_params', 'params_old', 'precision_matrices'} (ewc.py:117-132) ------------------------
    def model_state(self) -> Dict:
        a = self.arena
        return {"net_params": self.full_state_net(), "params_old": a.to_dict(self.p_old),
                "precision_matrices": a.to_dict(self.F)}

    def full_state_net(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in self.net.state_dict().items()}

    def update_model(self, params_state: Dict) -> None:
        if "net_params" in params_state:
            own = self.net.state_dict()
            with torch.no_grad():
                for k, v in params_state["net_params"].items():
                    if k in own:
                        own[k].copy_(v.to(own[k].device))
            if self.arena is not None:
                self.arena.refresh_shadow()
        if self.arena is not None:
            if "params_old" in params_state:
                self.arena.from_dict(params_state["params_old"], self.p_old)
            if "precision_matrices" in params_state:
                self.arena.from_dict(params_state["precision_matrices"], self.F)
            self.rebuild_penalty()
