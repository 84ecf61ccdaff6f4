"""``fedweit`` – Federated Weighted Inter-client Transfer (reference ``methods/fedweit.py``).

Every fully-trainable ``nn.Linear`` / ``nn.Conv2d`` leaf becomes a *decomposed* layer
``theta = mask (.) sw + aw + sum_k atten_k * aw_kb[..., k]`` (``fedweit.py:127-136``) with ``sw`` (shared weight) and
``aw_kb`` (knowledge base: ``kb_cnt`` other clients' adaptive weights stacked on a trailing dim) frozen and
``mask`` (one value per output unit, init ``sigmoid(0)``), ``aw`` (init ``(1-mask) sw``), ``atten`` (zeros) trained;
in train mode ``aw`` / ``mask`` are hard-thresholded at ``lambda_l1`` / ``lambda_mask``. Loss = criterion +
``lambda_l1 (||aw||_1 + ||mask||_1)``; the reference's ``lambda_l2`` "approx" term compares a module with itself and
is identically zero (``fedweit.py:599-619``) so it is not computed.

Exchange: upload ``theta`` ("gw") and ``aw``; server = FedAvg mean of ``gw`` -> ``sw`` plus, when at least ``kb_cnt``
clients have uploaded, ``random.sample`` of ``kb_cnt`` clients' ``aw`` stacked into ``aw_kb``
(``fedweit.py:983-1015``) – here a fused reduce (``FedComm.reduce_bcast``) and a strided all-gather that writes the
``[..., kb]`` layout directly (``FedComm.gather_strided``). After every dispatch the client re-initialises
``aw = (1-mask) sw`` and ``atten = 0`` (``fedweit.py:833-835``). Checkpoints are per task (``fedweit.py:898,918,945``).

Storage note: the reference keeps these tensors reverse-permuted; here they live in the arena's physical order and are
reverse-permuted only when a reference-schema checkpoint is written.
"""
from __future__ import annotations

import random
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops import layer as lops
from ..runtime.modules import ClientModule, ModelModule, OperatorModule, ServerModule
from ..utils.misc import tensor_reverse_permute


class Decomposed(nn.Module):
    """Decomposed replacement of one Linear / Conv2d leaf."""

    def __init__(self, src: nn.Module, kb_cnt: int, lambda_l1: float, lambda_mask: float):
        super().__init__()
        self.is_conv = isinstance(src, nn.Conv2d)
        if self.is_conv:
            self.stride, self.padding = src.stride, src.padding
        w = src.weight.detach()
        out = w.shape[0]
        self.kb_cnt, self.lambda_l1, self.lambda_mask = kb_cnt, lambda_l1, lambda_mask
        self.register_buffer("sw", w.clone())
        self.register_buffer("aw_kb", torch.zeros(*w.shape, kb_cnt))
        self.mask = nn.Parameter(torch.sigmoid(torch.zeros(out)))
        self.aw = nn.Parameter((1 - self._bmask(self.mask.detach(), w)) * w)
        self.atten = nn.Parameter(torch.zeros(kb_cnt))
        self.bias = nn.Parameter(src.bias.detach().clone()) if src.bias is not None else None

    @staticmethod
    def _bmask(mask: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
        return mask.view(-1, *([1] * (like.dim() - 1)))

    @staticmethod
    def _prune(w: torch.Tensor, thr: float) -> torch.Tensor:
        return w * (w.abs() > thr).to(w.dtype)

    def theta(self, prune: bool) -> torch.Tensor:
        if self.aw.is_cuda and self.kb_cnt <= lops.WC_MAX_K and lops.enabled("wcompose", self.aw.device):
            th = self._theta_fused(prune)
            if th is not None:
                return th
        aw = self._prune(self.aw, self.lambda_l1) if prune else self.aw
        mask = self._prune(self.mask, self.lambda_mask) if prune else self.mask
        return self._bmask(mask, self.sw) * self.sw + aw + (self.aw_kb * self.atten).sum(-1)

    def _theta_fused(self, prune: bool, use_ref: bool = False) -> Optional[torch.Tensor]:
        """One kernel instead of the ~10 element-wise passes of the expression above (``csrc/layer_ops.cu``): reads
        ``aw``, ``sw``, the ``kb`` stacked weights, writes ``theta`` in fp32 (carries the gradient) and bf16 (the
        tensor-core operand, attached as ``theta._flpr_bf16`` for the fast head); the backward kernel returns
        ``d aw`` / ``d mask`` / ``d atten`` from ``d theta`` in one sweep. ``sw`` and ``aw_kb`` are kept in ``aw``'s
        storage order (OHWI on CUDA) - they are only ever written through logical-shape ``copy_``s."""
        aw = self.aw
        if not (aw.is_contiguous() or lops.is_channels_last_4d(aw)) or lops.stack_phys(self.aw_kb, aw) is None \
                or lops.is_channels_last_4d(aw) != lops.is_channels_last_4d(self.sw) or not (
                    self.sw.is_contiguous() or lops.is_channels_last_4d(self.sw)):
            return None                                   # storage not aligned (see align_storage): original expression
        theta, t16 = lops.compose_weight(aw, self.aw_kb, self.atten, self.kb_cnt, self.sw, self.mask,
                                         self.lambda_l1, self.lambda_mask, prune, use_ref)
        theta._flpr_bf16 = t16
        return theta

    def align_storage(self) -> None:
        """Keep ``sw`` and ``aw_kb`` in ``aw``'s storage order (called once the arena has moved ``aw`` to OHWI storage;
        never from inside a step: replacing a buffer during a CUDA-graph capture would strand the old one)."""
        aw = self.aw
        with torch.no_grad():
            self.aw_kb = lops.stack_aligned(self.aw_kb, aw)
            if lops.is_channels_last_4d(aw):
                if not lops.is_channels_last_4d(self.sw):
                    self.sw = self.sw.contiguous(memory_format=torch.channels_last)
            elif not self.sw.is_contiguous():
                self.sw = self.sw.contiguous()

    def reinit(self) -> None:
        with torch.no_grad():
            self.aw.copy_((1 - self._bmask(self.mask, self.sw)) * self.sw)
            self.atten.zero_()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        th = self.theta(self.training)
        if self.is_conv:
            return F.conv2d(x, th.to(x.dtype) if x.dtype != th.dtype else th, self.bias, self.stride, self.padding)
        return F.linear(x, th.to(x.dtype) if x.dtype != th.dtype else th, self.bias)


class Model(ModelModule):
    def __init__(self, net, lambda_l1: float = 1e-3, lambda_l2: float = 1e2, lambda_mask: float = 0.0,
                 kb_cnt: int = 5, **kwargs):
        super().__init__(net, **kwargs)
        self.lambda_l1, self.lambda_l2, self.lambda_mask = float(lambda_l1), float(lambda_l2), float(lambda_mask)
        self.kb_cnt = int(kb_cnt)
        self.net_list: Dict[str, bool] = {}
        self.decomposed_names: List[str] = []
        self.layer_convert(self.net)

    def layer_convert(self, net: nn.Module) -> None:
        targets = []
        for name, mod in net.named_modules():
            if type(mod) in (nn.Linear, nn.Conv2d) and not list(mod.children()):
                ps = list(mod.parameters())
                if ps and all(p.requires_grad for p in ps):
                    targets.append((name, mod))
        for name, mod in targets:
            parent = net.get_submodule(name.rsplit(".", 1)[0]) if "." in name else net
            setattr(parent, name.rsplit(".", 1)[-1], Decomposed(mod, self.kb_cnt, self.lambda_l1, self.lambda_mask))
            self.decomposed_names.append(name)

    def materialize(self, device, compute_dtype: str = "bf16", fine_tuning=None):
        super().materialize(device, compute_dtype, fine_tuning)
        for _, layer in self.decomposed_module_leaves():
            layer.align_storage()
        if self.device.type == "cuda":
            lops.enabled("wcompose", self.device)      # one-time on-device self-check of the compose kernels, up front
        return self

    def decomposed_module_leaves(self):
        return [(n, self.net.get_submodule(n)) for n in self.decomposed_names]

    def remember_params(self, model_name: str) -> None:
        self.net_list[model_name] = True          # the reference deep-copies the net for a term that is always 0

    def upload_filter(self, name: str) -> bool:
        return name.endswith(".aw") and name[:-3] in self.decomposed_names

    # ---- flat views in arena (physical) order ------------------------------------------------------------------------
    @property
    def aw_numel(self) -> int:
        return self.arena.prefix_numel

    def flat_of(self, getter) -> torch.Tensor:
        """Concatenate ``getter(layer)`` (logical weight-shaped tensors) in the arena's ``aw`` segment order."""
        a = self.arena
        out = torch.zeros(self.aw_numel, device=self.device)
        for name, layer in self.decomposed_module_leaves():
            seg = a.segments[f"{name}.aw"]
            t = getter(layer).detach()
            out[seg.offset:seg.offset + seg.numel] = (t.permute(0, 2, 3, 1) if seg.channels_last else t).reshape(-1)
        return out

    def scatter_flat(self, flat: torch.Tensor, setter) -> None:
        a = self.arena
        for name, layer in self.decomposed_module_leaves():
            seg = a.segments[f"{name}.aw"]
            chunk = flat[seg.offset:seg.offset + seg.numel]
            if seg.channels_last:
                o, i, h, w = seg.shape
                t = chunk.view(o, h, w, i, *flat.shape[1:]).permute(0, 3, 1, 2, *range(4, 4 + flat.dim() - 1))
            else:
                t = chunk.view(*seg.shape, *flat.shape[1:])
            setter(layer, t)

    def theta_flat(self) -> torch.Tensor:
        return self.flat_of(lambda l: l.theta(False))

    # ---- reference checkpoint schema (fedweit.py:412-470), reverse-permuted like the reference ---------------------------
    def model_state(self) -> Dict:
        rp = lambda t: tensor_reverse_permute(t.detach()).clone(memory_format=torch.contiguous_format)  # noqa: E731
        layers = self.decomposed_module_leaves()
        dec_keys = set()
        for n, l in layers:
            dec_keys |= {f"{n}.{k}" for k in ("sw", "aw_kb", "mask", "aw", "atten", "bias")}
        pre = {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in self.net.state_dict().items()
               if k not in dec_keys}
        return {"sw": {f"{n}.sw": rp(l.sw) for n, l in layers},
                "aw": {f"{n}.aw": rp(l.aw) for n, l in layers},
                "mask": {f"{n}.mask": l.mask.detach().clone() for n, l in layers},
                "bias": {f"{n}.bias": l.bias.detach().clone() for n, l in layers if l.bias is not None},
                "atten": {f"{n}.atten": l.atten.detach().clone() for n, l in layers},
                "aw_kb": {f"{n}.aw_kb": torch.stack([tensor_reverse_permute(l.aw_kb[..., k].detach())
                                                     for k in range(l.kb_cnt)], dim=-1) for n, l in layers},
                "bn_params": {}, "pre_trained_params": pre}

    def update_model(self, params_state: Dict) -> None:
        with torch.no_grad():
            for group, attr in (("sw", "sw"), ("aw", "aw")):
                for key, t in (params_state.get(group) or {}).items():
                    layer = self.net.get_submodule(key[: -len(attr) - 1])
                    getattr(layer, attr).copy_(tensor_reverse_permute(t).to(self.device))
            for key, t in (params_state.get("aw_kb") or {}).items():
                layer = self.net.get_submodule(key[:-6])
                for k in range(layer.kb_cnt):
                    layer.aw_kb[..., k].copy_(tensor_reverse_permute(t[..., k]).to(self.device))
            for group in ("mask", "atten", "bias"):
                for key, t in (params_state.get(group) or {}).items():
                    layer = self.net.get_submodule(key[: -len(group) - 1])
                    getattr(layer, group).copy_(t.to(self.device))
            pre = params_state.get("pre_trained_params") or {}
            own = self.net.state_dict()
            for k, v in pre.items():
                if k in own:
                    own[k].copy_(v.to(own[k].device))
        if self.arena is not None:
            self.arena.refresh_shadow()


class Operator(OperatorModule):
    def compute_loss(self, model: Model, score, feature, target) -> torch.Tensor:
        loss = super().compute_loss(model, score, feature, target)
        sparse = 0.0
        for _, layer in model.decomposed_module_leaves():
            sparse = sparse + layer.aw.abs().sum() + layer.mask.abs().sum()
        return loss + model.lambda_l1 * sparse


class Client(ClientModule):
    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self.current_task: Optional[str] = None
        self._snapshots: Dict[str, Dict] = {}              # task -> model_state() kept when no checkpoint file is current
        self.operator.reset_lr_each_epoch = bool(getattr(self, "reference_compat", True))

    @classmethod
    def declare_buffers(cls, comm, model, token_numel: int = 0) -> None:
        n = model.aw_numel
        comm.alloc_client_buffer("gw", n)
        comm.alloc_client_buffer("aw", n)
        comm.alloc_client_buffer("cnt", 4)
        comm.alloc_rank_buffer("glob", n)

    def ckpt_name(self, task_name: str) -> str:
        return self.current_task or task_name              # one checkpoint per task

    def get_incremental_state(self, **kwargs) -> Dict:
        gw = self.comm.client_view("gw", self.client_id)
        aw = self.comm.client_view("aw", self.client_id)
        gw.copy_(self.model.theta_flat())
        aw.copy_(self.model.arena.master[:self.model.aw_numel])
        self.comm.client_view("cnt", self.client_id).fill_(float(self.train_cnt))
        a = self.model.arena
        # decomposed weights travel reverse-permuted, like everything the reference's layers hold (fedweit.py:785-802)
        named = lambda flat, suffix: {f"{n}.{suffix}": tensor_reverse_permute(a.view(flat, f"{n}.aw"))  # noqa: E731
                                      for n in self.model.decomposed_names}
        return {"train_cnt": self.train_cnt, "incremental_aw": named(aw, "aw"), "incremental_gw": named(gw, "sw"),
                "incremental_bn": {}}

    def _apply(self, state: Dict) -> None:
        sw_flat, kb_flat = state["_sw_flat"], state["_kb_flat"]
        self.model.scatter_flat(sw_flat, lambda l, t: l.sw.copy_(t))
        if kb_flat is not None:
            self.model.scatter_flat(kb_flat, lambda l, t: l.aw_kb.copy_(t))
        for _, layer in self.model.decomposed_module_leaves():
            layer.reinit()
        self.model.arena.refresh_shadow()

    def update_by_incremental_state(self, state: Dict, **kwargs) -> Any:
        with torch.no_grad():
            self._apply(state)
        self.logger.info("Update model succeed by incremental state from server.")

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        with torch.no_grad():
            self.model.update_model({"pre_trained_params": state.get("pre_trained_params") or {}})
            self._apply(state)
        self.logger.info("Update model succeed by integrated state from server.")

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        if self.current_task is not None and self.current_task != task_name:
            self.model.remember_params(task_name)
        self.current_task = task_name

    def after_epoch(self, output: Dict) -> None:
        self.train_cnt += output["data_count"]

    # ---- one model per task: other tasks are evaluated with the weights saved at the end of *their* last training
    #      round (``validate`` / ``inference`` start with ``load_model(task_name)``, fedweit.py:918,945) ----------------
    def after_train(self, task_name, tr_loader, val_loader, output) -> None:
        if self.store.enabled and not self.store.muted:
            self._snapshots.pop(task_name, None)           # the checkpoint written right after this call is current
        else:
            self._snapshots[task_name] = self.model.model_state()

    def _swap_in(self, task_name: str) -> Optional[Dict]:
        """Install the weights of another task's checkpoint; returns the state to restore afterwards."""
        if self.current_task is None or task_name == self.current_task:
            return None
        state = self._snapshots.get(task_name)
        if state is None and self.store.enabled and self.store.exists(self.name, task_name):
            state = self.load_state(task_name)
        if state is None:
            return None                                    # never trained: the resident weights (load_state default)
        resident = self.model.model_state()
        self.model.update_model(state)
        return resident

    def validate(self, task_name, query_loader, gallery_loader, device="cpu", **kwargs):
        resident = self._swap_in(task_name)
        try:
            return super().validate(task_name, query_loader, gallery_loader, device, **kwargs)
        finally:
            if resident is not None:
                self.model.update_model(resident)

    def inference(self, task_name, query_loader, gallery_loader, device="cpu", **kwargs):
        resident = self._swap_in(task_name)
        try:
            return super().inference(task_name, query_loader, gallery_loader, device, **kwargs)
        finally:
            if resident is not None:
                self.model.update_model(resident)


class Server(ServerModule):
    def __init__(self, server_name, model, operator, ckpt_root, **kwargs):
        super().__init__(server_name, model, operator, ckpt_root, **kwargs)
        self.client_ids: Dict[str, int] = {}
        self.uploaded: List[int] = []
        self.kb_flat: Optional[torch.Tensor] = None
        self.sw_flat: Optional[torch.Tensor] = None

    def bind_client(self, client_name: str, client_id: int, client=None) -> None:
        self.client_ids[client_name] = client_id

    def set_client_incremental_state(self, client_name: str, client_state: Optional[Dict]) -> None:
        if client_name not in self.clients:
            self.logger.warn(f"Collect incremental state failed from unregistered client {client_name}.")
            return
        self.clients[client_name] = client_state if client_state is not None else {"remote": True}
        cid = self.client_ids[client_name]
        if cid not in self.uploaded:
            self.uploaded.append(cid)
        self.logger.info(f"Collect incremental state successfully from client {client_name}.")

    set_client_integrated_state = set_client_incremental_state

    def calculate(self) -> Any:
        if not self.uploaded:
            return
        self.comm.reduce_bcast("gw", "glob", self.uploaded, cnt="cnt")
        self.sw_flat = self.comm.rank_view("glob")
        self.model.scatter_flat(self.sw_flat, lambda l, t: l.sw.copy_(t))
        if len(self.uploaded) >= self.model.kb_cnt:
            chosen = random.sample(list(self.uploaded), self.model.kb_cnt)     # same RNG stream on every rank
            n = self.model.aw_numel
            if self.kb_flat is None:
                self.kb_flat = torch.zeros(n, self.model.kb_cnt, device=self.model.device)
            self.comm.gather_strided("aw", chosen, self.kb_flat)
            self.model.scatter_flat(self.kb_flat, lambda l, t: l.aw_kb.copy_(t))

    def _state(self, prefix: str) -> Dict:
        ms = self.model.model_state()
        sw = self.sw_flat if self.sw_flat is not None else self.model.flat_of(lambda l: l.sw)
        return {f"{prefix}_sw": ms["sw"], f"{prefix}_aw_kb": ms["aw_kb"], "_sw_flat": sw, "_kb_flat": self.kb_flat,
                "_ms": ms}

    def get_dispatch_incremental_state(self, client_name: str) -> Dict:
        st = self._state("incremental")
        st.pop("_ms")
        return st

    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        st = self._state("integrated")
        ms = st.pop("_ms")
        st["integrated_bn"] = ms["bn_params"]
        st["pre_trained_params"] = ms["pre_trained_params"]
        return st
