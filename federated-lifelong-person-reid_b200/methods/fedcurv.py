"""``fedcurv`` – FedCurv (reference ``methods/fedcurv.py``).

Client penalty ``lam * sum_n [F (p-p_old)^2 + sum_j F_j (p-p_j)^2]`` over every registered client's last upload
``(F_j, p_j)``; the server aggregates parameters with plain FedAvg (the Fisher matrices are *not* used server-side,
``fedcurv.py:592-605``). The reference ships every client ``2K`` parameter-sized copies per round
(``fedcurv.py:621-646``); here the exchange is pre-reduced on the fabric to three moment buffers
``sum F_j``, ``sum F_j p_j``, ``sum F_j p_j^2`` (``FedComm.curv_moments``) which is mathematically identical because
the penalty is quadratic in ``p``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from ..runtime.modules import OperatorModule
from .fedbase import FedClient, FedServer
from .penalty import PenaltyModel


class Model(PenaltyModel):
    resume_attrs = ("F", "p_old", "Q", "R", "const", "other_f", "other_fp", "other_const", "have_moments")
    importance_mode = "fisher"
    skip_current_task = False

    def __init__(self, net, operator=None, **kwargs):
        super().__init__(net, operator, **kwargs)
        self.other_f: Optional[torch.Tensor] = None
        self.other_fp: Optional[torch.Tensor] = None
        self.other_const = 0.0

    def other_terms(self):
        if self.other_f is None:
            return None
        return self.other_f, self.other_fp

    def set_others(self, f: torch.Tensor, fp: torch.Tensor, fpp: Optional[torch.Tensor] = None) -> None:
        n = f.numel()
        if self.other_f is None:
            self.other_f, self.other_fp = self.arena.new_buffer(), self.arena.new_buffer()
        self.other_f[:n].copy_(f)
        self.other_fp[:n].copy_(fp)
        self.rebuild_penalty()

    def penalty(self) -> torch.Tensor:
        p = self.arena.master
        val = (self.F * (p - self.p_old) ** 2).sum()
        if self.other_f is not None:
            # sum_j F_j (p - p_j)^2 = p^2 sum F_j - 2 p sum F_j p_j + sum F_j p_j^2 (constant term omitted)
            val = val + (self.other_f * p * p - 2 * self.other_fp * p).sum()
        return self.lam * val


class Operator(OperatorModule):
    pass


class Client(FedClient):
    default_ckpt_name = "fedcurv_model"
    payload_prefix = ""

    @classmethod
    def declare_buffers(cls, comm, model, token_numel: int = 0) -> None:
        super().declare_buffers(comm, model, token_numel)
        n = cls._upload_numel(model)
        comm.alloc_client_buffer("fisher", n)
        for nm in ("curv_f", "curv_fp", "curv_fpp"):
            comm.alloc_rank_buffer(nm, n)

    def get_incremental_state(self, **kwargs) -> Dict:
        st = super().get_incremental_state(**kwargs)
        n = self.upload_numel()
        slot = self.comm.client_view("fisher", self.client_id)
        slot.copy_(self.model.F[:n])
        st["incremental_precision_matrices"] = self._named_prefix(slot)
        return st

    def _apply_others(self, state: Dict) -> None:
        if state.get("_curv") is not None:
            f, fp, fpp = state["_curv"]
            self.model.set_others(f, fp, fpp)

    def update_by_incremental_state(self, state: Dict, **kwargs) -> Any:
        super().update_by_incremental_state(state, **kwargs)
        self._apply_others(state)

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        super().update_by_integrated_state(state, **kwargs)
        self._apply_others(state)

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        self.model.install(self.operator.optimizer)

    def after_train(self, task_name, tr_loader, val_loader, output) -> None:
        self.model.remember_task(task_name, val_loader)          # the query loader (fedcurv.py:507)


class Server(FedServer):
    payload_prefix = ""
    def __init__(self, server_name, model, operator, ckpt_root, **kwargs):
        super().__init__(server_name, model, operator, ckpt_root, **kwargs)
        self.have_moments = False

    def calculate(self) -> Any:
        super().calculate()
        if self.uploaded:
            self.comm.curv_moments("fisher", "up", self.uploaded, "curv_f", "curv_fp", "curv_fpp")
            self.have_moments = True

    def _curv(self):
        if not self.have_moments:
            return None
        return tuple(self.comm.rank_view(nm) for nm in ("curv_f", "curv_fp", "curv_fpp"))

    def get_dispatch_incremental_state(self, client_name: str) -> Dict:
        st = super().get_dispatch_incremental_state(client_name)
        st["_curv"] = self._curv()
        return st

    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        st = super().get_dispatch_integrated_state(client_name)
        st["_curv"] = self._curv()
        return st
