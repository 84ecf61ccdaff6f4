"""``fedstil-atten`` – FedSTIL with a learned per-client attention over stacked global weights
(reference ``methods/fedstil_atten.py``).

Differences to ``fedstil``: ``global_weight`` carries a trailing **client** dimension ``K`` and
``global_weight_atten in R^K`` is *trained*: ``theta = sum_k atten_k * gw[..., k] + aw`` (``fedstil_atten.py:88-96``);
``aw`` is created once as ``(1 - atten_default) * w`` and persists across dispatches; the L1 term pulls ``atten`` and
``aw`` towards their values at the last dispatch (``:650-660``). The server does not average: it concatenates every
registered client's uploaded ``theta`` along the client dim (``:1099-1121``) and sends the stack to everyone
(``:1145-1149``) – here one strided all-gather (``FedComm.gather_strided``) that writes the ``[n, K]`` layout
directly on every rank, so dispatch to a local client is a device copy.

The composition is expressed as a ``torch.nn.utils.parametrize`` parametrization of each adaptive layer's ``weight``,
so the backbone code (including the tensor-core head) is unchanged and autograd routes ``d theta`` to ``aw`` / ``atten``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.utils.parametrize as parametrize

from . import fedstil as base
from ..ops import layer as lops


class _Compose(nn.Module):
    """weight = sum_k atten_k * gw[..., k] + aw ; ``gw`` has the weight's logical shape plus a trailing client dim
    ``Kmax`` and is *stored* in the weight's own element order (OHWI + K on CUDA, see ``Model.materialize``), so the
    fused compose kernel (``ops.layer.compose_weight``: one pass, fp32 + bf16 results, ``d atten`` by a deterministic
    two-stage reduction) walks ``aw`` and ``gw`` with the same flat index. Every other access goes through the logical
    shape, so the storage order is invisible outside this class."""

    def __init__(self, weight: torch.Tensor, k_max: int, atten_default: float):
        super().__init__()
        self.register_buffer("gw", torch.zeros(*weight.shape, k_max))
        self.gw[..., 0] = weight.detach()
        self.atten = nn.Parameter(torch.zeros(k_max))
        self.k_cur = 1
        self.k_max = int(k_max)
        self.atten_default = atten_default
        with torch.no_grad():
            self.atten[0] = atten_default
        self.register_buffer("atten0", self.atten.detach().clone())
        self.register_buffer("aw0", torch.zeros_like(weight))

    def _mix(self, atten: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
        return (self.gw.reshape(-1, self.k_max) @ atten).view(like.shape)

    def forward(self, aw: torch.Tensor) -> torch.Tensor:
        if aw.is_cuda and self.k_max <= lops.WC_MAX_K and lops.enabled("wcompose", aw.device) \
                and lops.stack_phys(self.gw, aw) is not None:
            theta, t16 = lops.compose_weight(aw, self.gw, self.atten, self.k_max)
            theta._flpr_bf16 = t16                       # the fast head's tensor-core operand: no separate cast pass
            return theta
        return self._mix(self.atten, aw) + aw

    def right_inverse(self, theta: torch.Tensor) -> torch.Tensor:
        return theta - self._mix(self.atten.detach(), theta)

    def align_storage(self, aw: torch.Tensor) -> None:
        """Re-lay ``gw`` out in ``aw``'s physical element order (after the arena moved ``aw`` to OHWI storage)."""
        with torch.no_grad():
            self.gw = lops.stack_aligned(self.gw, aw)


class Model(base.Model):
    def __init__(self, net, lambda_l1: float = 1e-4, lambda_k: int = 8000, atten_default: float = 0.80,
                 num_clients: int = 8, **kwargs):
        super().__init__(net, lambda_l1, lambda_k, atten_default, **kwargs)
        self.k_max = int(num_clients)
        self.composers: Dict[str, _Compose] = {}
        for lname in self.adaptive_names:
            mod = self.net.get_submodule(lname)
            comp = _Compose(mod.weight, self.k_max, self.atten_default)
            # registration stores ``right_inverse(w) = w - a * w`` as the trainable tensor: aw = (1 - a) * w, theta == w
            parametrize.register_parametrization(mod, "weight", comp, unsafe=True)
            comp.aw0.copy_(mod.parametrizations.weight.original.detach())
            self.composers[lname] = comp
        self._theta_params = {f"{n}.parametrizations.weight.original" for n in self.adaptive_names}

    def install(self, optimizer) -> None:              # plain optimizer: the L1 term is part of the autograd loss
        optimizer.G, optimizer.lam1, optimizer.atten = None, 0.0, 0.0
        if optimizer.stats is None:
            optimizer.stats = torch.zeros(2, dtype=torch.float32, device=self.device)

    def materialize(self, device, compute_dtype="bf16", fine_tuning=None):
        base.ModelModule.materialize(self, device, compute_dtype, fine_tuning)
        for lname, comp in self.composers.items():
            comp.align_storage(self.net.get_submodule(lname).parametrizations.weight.original)
        if self.device.type == "cuda":
            lops.enabled("wcompose", self.device)      # one-time on-device self-check of the compose kernels, up front
        self.G = None
        self.use_cuda_graphs = False                   # atten / K change between rounds: keep the step eager
        return self

    def sparseness(self) -> torch.Tensor:
        s = 0.0
        for lname, comp in self.composers.items():
            aw = self.net.get_submodule(lname).parametrizations.weight.original
            s = s + (comp.atten0 - comp.atten).abs().sum() + (comp.aw0 - aw).abs().sum()
        return s

    def theta_flat(self) -> torch.Tensor:
        """Composed weights in the arena's prefix order/layout (what a client uploads)."""
        a = self.arena
        out = torch.zeros(self.theta_numel, device=self.device)
        with torch.no_grad():
            for lname in self.adaptive_names:
                seg = a.segments[f"{lname}.parametrizations.weight.original"]
                th = self.net.get_submodule(lname).weight
                out[seg.offset:seg.offset + seg.numel] = (th.permute(0, 2, 3, 1) if seg.channels_last else th).reshape(-1)
        return out

    def set_global_stack(self, stack: torch.Tensor, k_cur: int) -> None:
        """``stack``: ``[theta_numel, Kmax]`` in arena layout; re-initialises atten and the L1 anchors."""
        a = self.arena
        with torch.no_grad():
            for lname, comp in self.composers.items():
                seg = a.segments[f"{lname}.parametrizations.weight.original"]
                chunk = stack[seg.offset:seg.offset + seg.numel]                       # [numel, Kmax], arena order
                if seg.channels_last:
                    o, i, h, w = seg.shape
                    chunk = chunk.view(o, h, w, i, -1).permute(0, 3, 1, 2, 4)          # logical view, no copy
                else:
                    chunk = chunk.view(*seg.shape, -1)
                comp.gw.copy_(chunk)                     # gw is stored in the same element order: a straight copy
                comp.k_cur = k_cur
                comp.atten.zero_()
                comp.atten[:k_cur] = self.atten_default
                comp.atten0.copy_(comp.atten)
                comp.aw0.copy_(self.net.get_submodule(lname).parametrizations.weight.original)

    def init_training_weights(self) -> None:
        """``AdaptiveLayer.init_training_weights()`` after a dispatch (``fedstil_atten.py:53-84``): attention back to
        ``atten_default`` over the ``K`` stacked weights, ``aw`` kept, L1 anchors re-taken."""
        with torch.no_grad():
            for lname, comp in self.composers.items():
                comp.atten.zero_()
                comp.atten[:comp.k_cur] = self.atten_default
                comp.atten0.copy_(comp.atten)
                comp.aw0.copy_(self.net.get_submodule(lname).parametrizations.weight.original)
        if self.arena is not None:
            self.arena.refresh_shadow()

    def model_state(self, copy: bool = True) -> Dict:
        gw, gwa, aw, ab = {}, {}, {}, {}
        for lname, comp in self.composers.items():
            mod = self.net.get_submodule(lname)
            shape = mod.parametrizations.weight.original.shape
            gw[f"{lname}.global_weight"] = comp.gw[..., :comp.k_cur].detach().clone(
                memory_format=torch.contiguous_format)
            gwa[f"{lname}.global_weight_atten"] = comp.atten[:comp.k_cur].detach().clone()
            aw[f"{lname}.adaptive_weight"] = mod.parametrizations.weight.original.detach().clone(
                memory_format=torch.contiguous_format).unsqueeze(-1)
            if getattr(mod, "bias", None) is not None:
                ab[f"{lname}.adaptive_bias"] = mod.bias.detach().clone()
        skip = set()
        for n in self.adaptive_names:
            skip |= {f"{n}.parametrizations.weight.original", f"{n}.bias", f"{n}.parametrizations.weight.0.gw",
                     f"{n}.parametrizations.weight.0.atten", f"{n}.parametrizations.weight.0.atten0",
                     f"{n}.parametrizations.weight.0.aw0"}
        pre = {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in self.net.state_dict().items()
               if k not in skip}
        return {"global_weight": gw, "global_weight_atten": gwa, "adaptive_weights": aw, "adaptive_bias": ab,
                "bn_params": {}, "pre_trained_params": pre}

    def update_model(self, params_state: Dict) -> None:
        with torch.no_grad():
            for key, g in (params_state.get("global_weight") or {}).items():
                comp = self.composers.get(key[: -len(".global_weight")])
                if comp is not None:
                    k = g.shape[-1]
                    comp.gw.zero_()
                    comp.gw[..., :k] = g.reshape(*comp.gw.shape[:-1], k).to(comp.gw.device)
                    comp.k_cur = k
            for key, w in (params_state.get("adaptive_weights") or {}).items():
                lname = key[: -len(".adaptive_weight")]
                if lname in self.composers:
                    orig = self.net.get_submodule(lname).parametrizations.weight.original
                    orig.copy_(w.squeeze(-1).to(orig.device))
            pre = params_state.get("pre_trained_params") or {}
            own = self.net.state_dict()
            for k2, v in pre.items():
                if k2 in own and own[k2].shape == v.shape:
                    own[k2].copy_(v.to(own[k2].device))
        if self.arena is not None:
            self.arena.refresh_shadow()

    def folded_trunk(self):
        return base.Model.folded_trunk(self)


class Operator(base.Operator):
    def compute_loss(self, model: Model, score, feature, target) -> torch.Tensor:
        return super().compute_loss(model, score, feature, target) + model.lambda_l1 * model.sparseness()

    def _ce_stats_ok(self, model: Model) -> bool:
        return False          # the reported loss includes the L1 term (fedstil_atten.py:650-666): accumulate the sum itself


class Client(base.Client):
    def get_incremental_state(self, **kwargs) -> Dict:
        slot = self.comm.client_view("theta_up", self.client_id)
        slot.copy_(self.model.theta_flat())
        self.comm.client_view("cnt", self.client_id).fill_(float(self.train_cnt))
        if self.task_token is not None and "token" in self.comm.bufs:
            tok = self.comm.client_view("token", self.client_id)
            tok.zero_()
            tok[:self.task_token.numel()].copy_(self.task_token)
        a = self.model.arena
        named = {f"{l}.global_weight": a.view(slot, f"{l}.parametrizations.weight.original").unsqueeze(-1)
                 for l in self.model.adaptive_names}
        return {"train_cnt": self.train_cnt, "task_token": self.task_token, "incremental_sw": named,
                "incremental_bn": {}}

    def update_by_incremental_state(self, state: Dict, **kwargs) -> Any:
        self.model.set_global_stack(state["_stack"], state["_k"])
        self.logger.info("Update model succeed by incremental state from server.")

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        self.model.update_model({"pre_trained_params": state.get("integrated_pre_trained_params") or {}})
        if state.get("_stack") is not None:
            self.model.set_global_stack(state["_stack"], state["_k"])
        else:
            # nothing uploaded yet: the server's own [..., 1] global weight becomes this client's (fedstil_atten.py:
            # 1145-1157,919-934) - theta = atten * gw_server + aw_client, NOT the client's own initial weights
            self.model.update_model({"global_weight": state["integrated_global_weight"]})
            self.model.init_training_weights()
        self.logger.info("Update model succeed by integrated state from server.")


class Server(base.Server):
    split_calculate = False         # the next dispatch needs the whole gathered stack: nothing can be deferred

    def __init__(self, server_name, model, operator, ckpt_root, **kwargs):
        super().__init__(server_name, model, operator, ckpt_root, **kwargs)
        self.stack: Optional[torch.Tensor] = None

    def calculate(self) -> Any:
        """All-gather of every registered client's theta along a trailing client dim (no averaging)."""
        if self.uploaded:
            n = self.model.theta_numel
            if self.stack is None:
                self.stack = torch.zeros(n, self.model.k_max, device=self.model.device)
            k = len(self.uploaded)
            out = torch.empty(n, k, device=self.model.device)
            self.comm.gather_strided("theta_up", self.uploaded, out)
            self.stack.zero_()
            self.stack[:, :k] = out
            self.model.set_global_stack(self.stack, k)
        if self._round_uploads and "token" in self.comm.bufs:
            ids = [self.client_ids[nm] for nm in self._round_uploads]
            d = self.comm.bufs["token"].n
            tk = torch.empty(d, len(ids), device=self.model.device)
            self.comm.gather_strided("token", ids, tk)
            toks = tk.t().contiguous()
            for i, name in enumerate(self._round_uploads):
                self.token_memory.setdefault(name, []).append(toks[i])
        self._round_uploads = []
        self.save_state(f"{self.server_name}_tokens", self.token_memory, True)

    def prepare_dispatch(self, online_names: Sequence[str], first_contact: Sequence[str]) -> None:
        return None

    def _payload(self) -> Dict:
        return {"_stack": self.stack, "_k": len(self.uploaded)}

    def get_dispatch_incremental_state(self, client_name: str) -> Optional[Dict]:
        if self.stack is None:
            return None
        ms = self.model.model_state()
        return {"incremental_shared_params": ms["global_weight"], **self._payload()}

    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        ms = self.model.model_state()
        return {"integrated_global_weight": ms["global_weight"], "integrated_bn_params": ms["bn_params"],
                "integrated_pre_trained_params": ms["pre_trained_params"], **self._payload()}
