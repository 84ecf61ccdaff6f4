"""``fedstil`` – Federated Spatial-Temporal Incremental Learning, the headline method
(reference ``methods/fedstil.py``).

Reference algorithm -> this implementation
------------------------------------------
* *Adaptive layers* ``theta = atten (.) G + A`` with ``G`` and ``atten`` frozen, ``A`` trained
  (``fedstil.py:24-129``). ``atten`` is never trained and is always re-initialised to ``atten_default``, so it is a
  scalar ``a``; and ``init_training_weights`` sets ``A0 = (1-a) G`` so the initial ``theta`` equals ``G``. Training
  ``A`` with gradient ``g_theta`` is therefore *identical* to training ``theta`` directly, with
  ``A = theta - a G`` (weight decay) and ``A - A0 = theta - G`` (L1 sparseness). We keep ONE fp32 master ``theta``
  per adaptive layer and fold weight decay on ``A``, ``lambda_l1 * sign(theta - G)`` and the bf16 refresh into the
  fused optimizer kernel: the per-step compose pass and the 6-11-layer Python L1 loop (``fedstil.py:639-644``)
  disappear. The reference's accidental training of its ``initial_*`` L1 anchors is reproduced under
  ``reference_compat`` (``engine_opts.train_l1_anchor``): the anchor and its moments are stepped inside the same fused
  optimizer kernel (``csrc/fused_ops.cu``; the CPU twin is ``ArenaOptimizer._anchor_step``).
* *Head discovery by torch.fx* (``fedstil.py:258-288``) -> static trunk / head split of the backbone.
* *Prototype pass* (``fedstil.py:558-617``): eval-mode frozen trunk over the task's train loader; the feature maps
  at the cut stay on the device (NHWC bf16) instead of bouncing through numpy; ``task_token`` = mean prototype.
* *Prototype rehearsal* (``fedstil.py:349-399``): herding runs on the device; an exemplar set is an ordered index
  list into a bank of unique prototypes (the reference stores ``m`` physical copies).
* *Server* (``fedstil.py:1047-1172``): FedAvg mean into the server model (``calculate``) and the
  similarity-weighted per-client mix ``out_i = sum_j W_ij theta_j`` (``get_dispatch_incremental_state``) run as
  peer-memory kernels (``FedComm.reduce_bcast`` / ``FedComm.mix``); the mix writes ``global_weight``, the ``theta``
  master and its bf16 copy of the receiving client in the same pass (= ``init_training_weights``).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..criterions import kl_distance
from ..utils.trace import host_list
from ..runtime.modules import ClientModule, ModelModule, OperatorModule, ServerModule


class Model(ModelModule):
    adaptive_types = (nn.Linear, nn.Conv2d)

    def __init__(self, net, lambda_l1: float = 1e-4, lambda_k: int = 8000, atten_default: float = 0.80, **kwargs):
        super().__init__(net, **kwargs)
        self.atten_default = float(atten_default)
        self.lambda_l1 = float(lambda_l1)
        self.lambda_k = int(lambda_k)
        self.adaptive_names: List[str] = self._find_adaptive_layers()
        self._theta_params = {f"{n}.weight" for n in self.adaptive_names}
        self.ids: set = set()
        self.ex_gens: List[Dict[str, Any]] = []          # exemplar memory (see "exemplar memory" below)
        self.G: Optional[torch.Tensor] = None
        self.train_l1_anchor = False                     # engine_opts.train_l1_anchor (set by the Client)
        self.anchor: Optional[torch.Tensor] = None

    # ---- adaptive layers: leaves of type Linear/Conv2d whose parameters are all trainable (fedstil.py:290-347) ------
    def _find_adaptive_layers(self) -> List[str]:
        names = []
        for name, mod in self.net.named_modules():
            if type(mod) in self.adaptive_types and len(list(mod.children())) == 0:
                ps = list(mod.parameters())
                if ps and all(p.requires_grad for p in ps):
                    names.append(name)
        return names

    def upload_filter(self, name: str) -> bool:
        return name in self._theta_params

    # ---- resume manifest ---------------------------------------------------------------------------------------------
    def resume_extra(self) -> Dict[str, Any]:
        return {"G": self.G.detach().clone() if self.G is not None else None, "ids": sorted(self.ids),
                "ex_gens": [{"pids": g["pids"], "bank": g["bank"], "cls": g["cls"], "k": int(g["k"])}
                            for g in self.ex_gens]}

    def load_resume_extra(self, extra: Dict[str, Any]) -> None:
        if extra.get("G") is not None and self.G is not None:
            self.G.copy_(extra["G"].to(self.G.device))
        self.ids = set(int(i) for i in extra.get("ids", []))
        self.ex_gens = []
        for g in extra.get("ex_gens", []):
            pids = g["pids"].to(self.device)
            self.ex_gens.append({"pids": pids, "pid_list": [int(p) for p in pids.tolist()], "k": int(g["k"]),
                                 "bank": g["bank"].to(self.device), "cls": g["cls"].to(self.device)})

    def materialize(self, device, compute_dtype="bf16", fine_tuning=None):
        super().materialize(device, compute_dtype, fine_tuning)
        self.G = self.arena.master[:self.theta_numel].clone()       # global_weight (frozen between dispatches)
        return self

    @property
    def theta_numel(self) -> int:
        return self.arena.prefix_numel

    def install(self, optimizer) -> None:
        optimizer.G, optimizer.lam1, optimizer.atten = self.G, self.lambda_l1, self.atten_default
        if self.train_l1_anchor and self.anchor is None:
            self.anchor = self.G.clone()
        optimizer.anchor = self.anchor
        if optimizer.stats is None:
            optimizer.stats = torch.zeros(2, dtype=torch.float32, device=self.device)

    def reset_anchor(self) -> None:
        """``init_training_weights()`` re-takes the L1 anchors after every dispatch (fedstil.py:71-76,879-892)."""
        if self.anchor is not None:
            self.anchor.copy_(self.G)

    def set_global_weight(self, flat: torch.Tensor) -> None:
        """New ``global_weight`` + ``init_training_weights()``: theta <- G (A = (1-a) G), bf16 copy refreshed."""
        n = self.theta_numel
        self.G.copy_(flat[:n])
        self.arena.master[:n].copy_(self.G)
        self.arena.refresh_shadow()
        self.reset_anchor()

    @property
    def m(self) -> int:
        return math.ceil(self.lambda_k / max(len(self.ids), 1))

    # ---- reference checkpoint schema (fedstil.py:444-491) ------------------------------------------------------------
    def model_state(self, copy: bool = True) -> Dict:
        """``copy=False`` returns views of the live tensors for callers that serialise immediately (the checkpoint
        store snapshots them when ``save`` is called). That dict is *persistent*: the same object, over the same
        device buffers, is handed out on every call - only the flat ``adaptive_weight = theta - a G`` buffer behind
        its ``adaptive_weights`` views is recomputed (one fused op) - so the store can keep its copy plan
        (``PersistentState.plan_token``) instead of re-analysing ~330 tensors on every snapshot."""
        if not copy:
            return self._persistent_state()
        a = self.arena
        own = lambda t: t.clone(memory_format=torch.contiguous_format)          # noqa: E731
        gw, gwa, aw, ab = {}, {}, {}, {}
        for lname in self.adaptive_names:
            theta = a.view(a.master, f"{lname}.weight").detach()
            g = a.view(self.G, f"{lname}.weight").detach()
            gw[f"{lname}.global_weight"] = own(g)
            gwa[f"{lname}.global_weight_atten"] = torch.full((theta.shape[-1],), self.atten_default,
                                                             device=theta.device)
            aw[f"{lname}.adaptive_weight"] = (theta - self.atten_default * g).contiguous()
            bias = getattr(self.net.get_submodule(lname), "bias", None)
            if bias is not None:
                ab[f"{lname}.adaptive_bias"] = own(bias.detach())
        skip = {f"{n}.weight" for n in self.adaptive_names} | {f"{n}.bias" for n in self.adaptive_names}
        pre = {k: own(v.detach()) for k, v in self.net.state_dict().items() if k not in skip}
        return {"global_weight": gw, "global_weight_atten": gwa, "adaptive_weights": aw, "adaptive_bias": ab,
                "bn_params": {}, "pre_trained_params": pre}

    def _persistent_state(self) -> Dict:
        from ..runtime.checkpoint import PersistentState
        a = self.arena
        n = self.theta_numel
        ver = getattr(self, "_static_version", 0)
        st = getattr(self, "_pstate", None)
        if st is None or st.plan_token[1] != ver or self._pstate_ptr != (a.master.data_ptr(), self.G.data_ptr()):
            self._aw_flat = torch.empty(n, dtype=torch.float32, device=a.master.device)
            gw, gwa, aw, ab = {}, {}, {}, {}
            for lname in self.adaptive_names:
                theta = a.view(a.master, f"{lname}.weight").detach()
                gw[f"{lname}.global_weight"] = a.view(self.G, f"{lname}.weight").detach()
                gwa[f"{lname}.global_weight_atten"] = torch.full((theta.shape[-1],), self.atten_default,
                                                                 device=theta.device)
                aw[f"{lname}.adaptive_weight"] = a.view(self._aw_flat, f"{lname}.weight")
                bias = getattr(self.net.get_submodule(lname), "bias", None)
                if bias is not None:
                    ab[f"{lname}.adaptive_bias"] = bias.detach()
            skip = {f"{nm}.weight" for nm in self.adaptive_names} | {f"{nm}.bias" for nm in self.adaptive_names}
            pre = {k: v.detach() for k, v in self.net.state_dict().items() if k not in skip}
            # The frozen stages never change between dispatches of pre-trained parameters (the trunk runs in eval
            # mode, fedstil.py:569): the checkpoint store keeps them in its device image instead of re-copying ~250
            # tensors on every snapshot. ``_static_version`` is bumped whenever something writes them (update_model).
            prefixes = self._frozen_prefixes()
            for k, v in pre.items():
                if k.startswith(prefixes):
                    v._flpr_static = ver
            st = self._pstate = PersistentState(
                {"global_weight": gw, "global_weight_atten": gwa, "adaptive_weights": aw, "adaptive_bias": ab,
                 "bn_params": {}, "pre_trained_params": pre})
            st.plan_token = (id(self), ver)
            self._pstate_ptr = (a.master.data_ptr(), self.G.data_ptr())
        torch.sub(a.master[:n].detach(), self.G, alpha=self.atten_default, out=self._aw_flat)   # A = theta - a G
        return st

    def _frozen_prefixes(self) -> Tuple[str, ...]:
        start = getattr(self.net, "head_start", 0)
        if not hasattr(self.net, "base") or not isinstance(start, int) or start < 1 or not hasattr(self.net.base, "conv1"):
            return ("\0",)                               # unknown backbone: nothing is declared static
        return ("base.conv1.", "base.bn1.") + tuple(f"base.layer{i}." for i in range(1, start))

    def update_model(self, params_state: Dict) -> None:
        a = self.arena
        if params_state.get("pre_trained_params"):
            self._static_version = getattr(self, "_static_version", 0) + 1
        with torch.no_grad():
            gw = params_state.get("global_weight") or {}
            aw = params_state.get("adaptive_weights") or {}
            for key, g in gw.items():
                lname = key[: -len(".global_weight")]
                pname = f"{lname}.weight"
                if pname in a.segments:
                    a.from_dict({pname: g}, self.G)
                    if f"{lname}.adaptive_weight" not in aw:
                        a.from_dict({pname: g}, a.master)              # theta = G until told otherwise
            for key, w in aw.items():
                lname = key[: -len(".adaptive_weight")]
                pname = f"{lname}.weight"
                if pname in a.segments:
                    g = a.view(self.G, pname)
                    a.view(a.master, pname).copy_(self.atten_default * g + w.to(g.device))
            for key, b in (params_state.get("adaptive_bias") or {}).items():
                lname = key[: -len(".adaptive_bias")]
                mod = self.net.get_submodule(lname)
                if mod.bias is not None:
                    mod.bias.copy_(b.to(mod.bias.device))
            pre = params_state.get("pre_trained_params") or {}
            if pre:
                own = self.net.state_dict()
                for k, v in pre.items():
                    if k in own:
                        own[k].copy_(v.to(own[k].device))
        a.refresh_shadow()

    # ---- prototypes ------------------------------------------------------------------------------------------------
    def folded_trunk(self):
        """BN-folded, CUDA-graphed frozen trunk (CUDA + bf16 + ResNet only), shared by the clients of a rank."""
        if self.compute_dtype != torch.bfloat16 or not hasattr(self.net, "base") or \
                not getattr(self, "use_folded_trunk", True):
            return None
        ft = getattr(self, "_folded", None)
        if ft is None:
            from ..models.frozen import shared_folded_trunk
            from ..models.resnet import ResNetReID
            if not isinstance(self.net, ResNetReID) or self.net.head_start < 1:
                self.use_folded_trunk = False
                return None
            ft = self._folded = shared_folded_trunk(self.net, torch.bfloat16)
        return ft

    def forward_trunk(self, data: torch.Tensor) -> torch.Tensor:
        return self.net.forward_trunk(data)

    def forward_head(self, protos: torch.Tensor):
        return self.net.forward_head(protos)

    # ---- exemplar memory (fedstil.py:349-399) -----------------------------------------------------------------------
    # Stored as *generations* (one per build_examplars call): ``{"pids": LongTensor[P], "bank": [P, m, ...],
    # "cls": [P, m], "k": kept exemplars per identity}``. The reference's per-identity python lists
    # ``{pid: [(prototype, class_id)] * m}`` are a view of this (``examplars`` property, checkpoint writer).
    def reduce_examplars(self) -> None:
        m = self.m
        for gen in self.ex_gens:
            gen["k"] = min(gen["k"], m)

    @property
    def examplars(self) -> Dict[int, Dict[str, Any]]:
        out: Dict[int, Dict[str, Any]] = {}
        for gen in self.ex_gens:
            k = gen["k"]
            for gi, pid in enumerate(gen["pid_list"]):
                out[pid] = {"bank": gen["bank"][gi, :k], "cls": gen["cls"][gi, :k], "order": list(range(k))}
        return out

    @torch.no_grad()
    def build_examplars(self, protos: torch.Tensor, pids: torch.Tensor, classes: torch.Tensor,
                        person_ids: Sequence[int], batch_size: int = 256) -> None:
        """Herding on ``forward_head`` features of the epoch's prototype set (exemplars + current task)."""
        self.eval()
        keep = sorted(set(int(p) for p in person_ids))
        m = self.m
        dev = protos.device
        if keep:
            keep_t = torch.tensor(keep, device=dev)
            rows = torch.nonzero(torch.isin(pids, keep_t)).squeeze(1)       # one host sync (sizes)
        else:
            rows = torch.arange(pids.numel(), device=dev)
        if rows.numel() == 0 or m <= 0:
            return
        feats = []
        fwd = self._herding_forward()
        for s in range(0, rows.numel(), batch_size):
            feats.append(fwd(protos[rows[s:s + batch_size]]))
        feats = torch.cat(feats)                                             # [n_sel, D], aligned with `rows`
        sel_pids = pids[rows]
        order = torch.argsort(sel_pids, stable=True)
        upid, counts = torch.unique_consecutive(sel_pids[order], return_counts=True)
        upid_l, counts_l = host_list(upid), host_list(counts)                # the second (and last) host sync
        P, nmax = len(upid_l), max(counts_l)
        starts = torch.cumsum(counts, 0) - counts
        ar = torch.arange(nmax, device=dev)
        pos = (starts[:, None] + ar[None, :]).clamp_(max=order.numel() - 1)
        idx = order[pos]                                                     # [P, nmax] rows of `feats` per identity
        picks = herding_select_batched(feats, idx, counts, m)               # [P, m] positions within each identity
        sel = rows[idx.gather(1, picks)]                                     # [P, m] rows of `protos`
        gen = {"pids": upid, "pid_list": [int(p) for p in upid_l], "k": m,
               "bank": protos[sel.reshape(-1)].reshape(P, m, *protos.shape[1:]),
               "cls": classes[sel.reshape(-1)].reshape(P, m)}
        # identities that are herded again replace their older exemplars
        fresh = set(gen["pid_list"])
        kept = []
        for old in self.ex_gens:
            if fresh.isdisjoint(old["pid_list"]):
                kept.append(old)
                continue
            mask = [p not in fresh for p in old["pid_list"]]
            if any(mask):
                mk = torch.tensor(mask, device=dev)
                kept.append({"pids": old["pids"][mk], "pid_list": [p for p, k_ in zip(old["pid_list"], mask) if k_],
                             "k": old["k"], "bank": old["bank"][mk], "cls": old["cls"][mk]})
        self.ex_gens = kept + [gen]

    def _herding_forward(self):
        """Eval-mode head forward that produces the herding features, CUDA-graph captured per batch shape: ~100 eager
        launches per 256-prototype batch cost three times their device time in Python dispatch."""
        fwd = getattr(self, "_herd_fwd", None)
        if fwd is None:
            from ..runtime.graphs import GraphedForward

            def fn(x):
                with self.autocast():
                    return self.forward_head(x).float()
            ok = (self.device.type == "cuda" and getattr(self, "use_cuda_graphs", True)
                  and getattr(self.net, "thread_safe_rng", False) and getattr(self.net, "_fast_head", None) is not None)
            fwd = self._herd_fwd = GraphedForward(fn, warmup=1, enabled=ok)
        return fwd

    def examplar_tensors(self) -> Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Expanded rehearsal set ``(protos, person_ids, class_ids)`` (duplicates included, like the reference)."""
        ps, ids, cs = [], [], []
        for gen in self.ex_gens:
            k = gen["k"]
            if k <= 0 or not gen["pid_list"]:
                continue
            P = len(gen["pid_list"])
            ps.append(gen["bank"][:, :k].reshape(P * k, *gen["bank"].shape[2:]))
            cs.append(gen["cls"][:, :k].reshape(-1))
            ids.append(gen["pids"].repeat_interleave(k))
        if not ps:
            return None
        ids_cat, cs_cat = torch.cat(ids), torch.cat(cs)
        if getattr(self, "relabel_by_class_index", False) and len(ids) > 1:
            # Reference quirk (reference_compat): the rehearsal set is a ``ReIDImageDataset(source=examplars)`` whose
            # dict branch keeps ONE person id per *class index* - ``classes[class_id] = person_id``, last insertion
            # wins (datasets_loader.py:21-27,36-38). Class indices restart at 0 in every task folder, so as soon as
            # exemplars of two tasks coexist, the older task's exemplars are served (trained, and re-herded) under the
            # person ids of the newest task that has the same class index.
            lut = getattr(self, "_cls_lut", None)
            size = max(int(getattr(self.net, "num_classes", 0) or 0), 1 << 16)
            if lut is None or lut.numel() != size or lut.device != ids_cat.device:
                lut = self._cls_lut = torch.zeros(size, dtype=torch.long, device=ids_cat.device)
            for pid_v, cls_v in zip(ids, cs):                   # generation order == insertion order of the dict
                lut[cls_v] = pid_v
            ids_cat = lut[cs_cat]
        return torch.cat(ps), ids_cat, cs_cat

    def examplars_compact(self) -> Dict:
        """Compact exemplar memory; the checkpoint writer process expands it to the reference schema."""
        return {"_compact_gens": [{"pids": g["pids"], "bank": g["bank"], "cls": g["cls"], "k": int(g["k"])}
                                  for g in self.ex_gens]}

    def examplars_state(self, max_bytes: int = 256 << 20) -> Dict:
        """``{np.int64 pid: [(ndarray proto, class_id), ...]}`` (``fedstil.py:841,846``)."""
        out = {}
        for pid, ex in self.examplars.items():
            bank = ex["bank"].float().cpu().numpy()
            cls = ex["cls"].cpu().tolist()
            out[np.int64(pid)] = [(bank[i], int(cls[i])) for i in ex["order"]]
        return out

    def load_examplars_state(self, state: Dict) -> None:
        """Inverse of :meth:`examplars_state` (resume from a reference-schema exemplar checkpoint)."""
        self.ex_gens = []
        by_len: Dict[int, List] = {}
        for pid, items in state.items():
            by_len.setdefault(len(items), []).append((int(pid), items))
        dt = torch.bfloat16 if self.compute_dtype == torch.bfloat16 else torch.float32
        for k, group in by_len.items():
            if k == 0:
                continue
            bank = torch.stack([torch.stack([torch.as_tensor(pr) for pr, _ in items]) for _, items in group])
            cls = torch.tensor([[int(c) for _, c in items] for _, items in group])
            pl = [pid for pid, _ in group]
            self.ex_gens.append({"pids": torch.tensor(pl, device=self.device), "pid_list": pl, "k": k,
                                 "bank": bank.to(self.device, dt), "cls": cls.to(self.device)})


def herding_select(feats: torch.Tensor, m: int) -> List[int]:
    """iCaRL herding: at step t pick ``argmin_i || mean - (f_i + sum selected) / (t+1) ||`` (duplicates allowed).

    ``argmin_i ||c - f_i||`` with ``c = (t+1) mean - S`` only needs ``||f_i||^2 - 2 f_i.c`` -> one mat-vec per step,
    no Python loop over samples and no host round-trip per step (indices are collected on the device).
    """
    n = feats.shape[0]
    if n == 0:
        return []
    f = feats.float()
    mean = f.mean(0)
    sq = (f * f).sum(1)
    S = torch.zeros_like(mean)
    picks = torch.empty(m, dtype=torch.long, device=f.device)
    for t in range(m):
        c = (t + 1) * mean - S
        i = torch.argmin(sq - 2 * (f @ c))
        picks[t] = i
        S = S + f[i]
    return host_list(picks)


def group_matrix(groups: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """List of per-identity row-index vectors -> padded ``[P, nmax]`` index matrix + ``[P]`` counts."""
    dev = groups[0].device
    counts = torch.tensor([int(g.numel()) for g in groups], device=dev)
    idx = torch.zeros(len(groups), int(counts.max()), dtype=torch.long, device=dev)
    for gi, g in enumerate(groups):
        idx[gi, :g.numel()] = g
    return idx, counts


def herding_select_batched(feats: torch.Tensor, idx: torch.Tensor, counts: torch.Tensor, m: int) -> torch.Tensor:
    """Herding for all identities at once. ``idx[g, :counts[g]]`` are the rows of ``feats`` that belong to identity
    g. Returns ``[P, m]`` positions *within each identity's row list* (same rule as :func:`herding_select`).
    CUDA: one kernel, one block per identity, the whole m-step loop on the device (``csrc/fused_ops.cu``)."""
    P, nmax = idx.shape
    dev = feats.device
    picks = torch.empty(P, m, dtype=torch.long, device=dev)
    from ..ops import native
    if native.on_device(feats, "flpr_herding"):
        lib = native.kernels()
        f, idx_c, cnt32 = feats.float().contiguous(), idx.contiguous(), counts.to(torch.int32)   # (kept alive over the call)
        rc = lib.flpr_herding(native.ptr(f), native.ptr(idx_c), native.ptr(cnt32),
                              native.ptr(picks), P, nmax, f.shape[1], m, native.stream_of(dev))
        if rc == 0:
            native.count_launch()
            return picks
        if rc != -3:                                   # -3: does not fit in shared memory -> tensor-op path below
            native.check(rc, "flpr_herding")
    valid = torch.arange(nmax, device=dev)[None, :] < counts[:, None]
    f = feats.float()[idx] * valid.unsqueeze(-1)                             # [P, nmax, D]
    cnt = counts.view(-1, 1).clamp(min=1)
    mean = f.sum(1) / cnt                                                    # [P, D]
    sq = (f * f).sum(2).masked_fill(~valid, float("inf"))                    # padding can never win the argmin
    S = torch.zeros_like(mean)
    ar = torch.arange(P, device=dev)
    for t in range(m):
        c = (t + 1) * mean - S
        score = sq - 2 * torch.bmm(f, c.unsqueeze(2)).squeeze(2)
        i = torch.argmin(score, dim=1)
        picks[:, t] = i
        S = S + f[ar, i]
    return picks


class Operator(OperatorModule):

    @torch.no_grad()
    def generate_prototypes(self, model: Model, source_loader) -> Dict[str, torch.Tensor]:
        """Eval-mode trunk pass; returns this epoch's rehearsal+task prototype set and the task token."""
        model.eval()
        protos, pids, cids = [], [], []
        folded = model.folded_trunk()
        chunk = int(getattr(model, "trunk_batch", 256))         # the frozen trunk is inference-only: batch it wider
        if hasattr(source_loader, "iterate") and (folded is not None or model.device.type == "cuda"):
            batches = source_loader.iterate(chunk, ordered=True)   # sample order is irrelevant here (permuted later)
        else:
            batches = source_loader
        pend: List[torch.Tensor] = []

        def flush():
            if pend:
                big = torch.cat(pend) if len(pend) > 1 else pend[0]
                protos.append(folded(big))                     # caller-owned copy of the graph's output buffer
                pend.clear()

        for data, person_id, classes_id in batches:
            data = model.prepare_input(data)
            if folded is not None:
                pend.append(data)
                if sum(d.shape[0] for d in pend) >= chunk:
                    flush()
            else:
                with model.autocast():
                    fmap = model.forward_trunk(data)
                if model.compute_dtype == torch.bfloat16:
                    fmap = fmap.to(torch.bfloat16)
                protos.append(fmap)
            pids.append(person_id.to(model.device))
            cids.append(classes_id.to(model.device))
        flush()
        protos, pids, cids = torch.cat(protos), torch.cat(pids), torch.cat(cids)
        task_token = protos.float().flatten(1).mean(0)
        ex = model.examplar_tensors()
        if ex is not None:                                   # ConcatDataset([exemplars, current]) (fedstil.py:590-593)
            protos = torch.cat([ex[0].to(protos.dtype), protos])
            pids = torch.cat([ex[1], pids])
            cids = torch.cat([ex[2], cids])
        return {"protos": protos, "pids": pids, "cids": cids, "task_token": task_token}

    def forward_train(self, model: Model, data: torch.Tensor):
        with model.autocast():
            return model.forward_head(data)

    def _graph_capable(self, model: ModelModule) -> bool:
        # the head step (prototype batch in, device-side accumulators out) is capturable for every backbone; Swin's
        # stochastic depth draws from the default generator, which torch registers with the capture - and a backbone
        # that is not ``thread_safe_rng`` is never trained on concurrent client threads
        return model.device.type == "cuda" and getattr(model, "use_cuda_graphs", True)

    def invoke_train(self, model: Model, dataloader, **kwargs) -> Dict:
        from ..utils.trace import nvtx_range
        device = model.device
        with nvtx_range("fedstil/prototype_pass"):
            pset = self.generate_prototypes(model, dataloader)
        protos, pids = pset["protos"], pset["pids"]
        n = protos.shape[0]
        bs = dataloader.batch_size
        model.train()
        self.begin_epoch()
        model.install(self.optimizer)
        self.optimizer.stats.zero_()
        perm = torch.randperm(n, device=device, generator=model.rng if device.type == "cuda" else None)
        n_batches = n // bs if (n % bs == 1) else (n + bs - 1) // bs     # drop_last iff remainder == 1
        data_cnt = 0
        step = self._graphed_step(model)
        acc = self._acc
        acc.zero_()
        if self._ce_acc is not None:
            self._ce_acc.zero_()
        self._bn_counters(model, on=False)            # 13 one-element kernels per step -> 13 per epoch (below)
        for b in range(n_batches):
            idx = perm[b * bs:(b + 1) * bs]
            step(protos[idx], pids[idx])
            data_cnt += len(idx)
        self._bn_counters(model, on=True, add=n_batches)
        src = acc if self._ce_acc is None else self._ce_acc.double()
        vals = host_list(torch.cat([src, self.optimizer.stats.double()]))    # single host sync per epoch
        loss_sum, hits, _, l1_sum = vals
        train_loss = (loss_sum + model.lambda_l1 * l1_sum) / max(n_batches, 1)
        if self.scheduler:
            self.scheduler.step()
        return {"task_token": pset["task_token"], "proto_set": pset, "accuracy": hits / max(data_cnt, 1),
                "loss": train_loss, "batch_count": n_batches, "data_count": data_cnt}


class Client(ClientModule):
    default_ckpt_name = None

    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self.current_task = None
        self.task_token: Optional[torch.Tensor] = None
        self._task_tokens: List[torch.Tensor] = []
        self.model.relabel_by_class_index = bool(getattr(self, "reference_compat", True))
        self.operator.reset_lr_each_epoch = bool(getattr(self, "reference_compat", True))
        self.model.train_l1_anchor = bool(getattr(self, "train_l1_anchor", False)) and hasattr(self.model, "G")

    # ---- symmetric buffers ---------------------------------------------------------------------------------------------
    @classmethod
    def declare_buffers(cls, comm, model, token_numel: int = 0) -> None:
        n = model.theta_numel
        comm.alloc_client_buffer("theta_up", n)
        comm.alloc_client_buffer("cnt", 4)
        comm.alloc_rank_buffer("glob", n)
        if token_numel:
            comm.alloc_client_buffer("token", (token_numel + 3) // 4 * 4)

    def resume_extra(self) -> Dict[str, Any]:
        return {"current_task": self.current_task, "task_token": self.task_token}

    def load_resume_extra(self, extra: Dict[str, Any]) -> None:
        self.current_task = extra.get("current_task")
        tok = extra.get("task_token")
        self.task_token = tok.to(self.model.device) if tok is not None else None

    def _named_theta(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        a = self.model.arena
        return {f"{l}.global_weight": a.view(flat, f"{l}.weight") for l in self.model.adaptive_names}

    # ---- checkpoints -----------------------------------------------------------------------------------------------------
    def save_model(self, model_name: str) -> None:
        self.save_state(model_name, self.model.model_state(copy=False), True)
        self.store.save(self.name, f"{model_name}_examplars", self.model.examplars_compact(), True,
                        post="expand_examplars")

    # ---- protocol ----------------------------------------------------------------------------------------------------------
    def get_incremental_state(self, **kwargs) -> Dict:
        n = self.model.theta_numel
        slot = self.comm.client_view("theta_up", self.client_id)
        slot.copy_(self.model.arena.master[:n])               # theta = a*G + A  (fedstil.py:851-854)
        self.comm.client_view("cnt", self.client_id).fill_(float(self.train_cnt))
        if self.task_token is not None and "token" in self.comm.bufs:
            tok = self.comm.client_view("token", self.client_id)
            tok.zero_()
            tok[:self.task_token.numel()].copy_(self.task_token)
        return {"train_cnt": self.train_cnt, "task_token": self.task_token,
                "incremental_sw": self._named_theta(slot), "incremental_bn": {}}

    def get_integrated_state(self, **kwargs) -> Dict:
        st = self.get_incremental_state()
        ms = self.model.model_state()
        return {"train_cnt": st["train_cnt"], "task_token": st["task_token"], "integrated_sw": st["incremental_sw"],
                "integrated_bn": {}, "pre_trained_params": ms["pre_trained_params"]}

    def update_by_incremental_state(self, state: Dict, **kwargs) -> Any:
        if state.get("_delivered"):
            self.model.reset_anchor()                           # the mix kernel already wrote G / theta / bf16 copy
        else:
            self.model.update_model({"global_weight": state["incremental_shared_params"]})
            self.model.set_global_weight(self.model.G)
        self.logger.info("Update model succeed by incremental state from server.")

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        self.model.update_model({"global_weight": state["integrated_global_weight"],
                                 "pre_trained_params": state["integrated_pre_trained_params"]})
        self.model.set_global_weight(self.model.G)
        self.logger.info("Update model succeed by integrated state from server.")

    # ---- training ----------------------------------------------------------------------------------------------------------
    def ckpt_name(self, task_name: str) -> str:
        return self.model_ckpt_name if self.model_ckpt_name else (self.current_task or task_name)

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        if self.current_task is None or self.current_task != task_name:
            self.model.ids.update(int(p) for p in tr_loader.dataset.person_ids)
        self.current_task = task_name
        self._task_tokens = []

    def after_epoch(self, output: Dict) -> None:
        self._task_tokens.append(output["task_token"])
        self.train_cnt += output["data_count"]                 # never reset

    def after_train(self, task_name, tr_loader, val_loader, output) -> None:
        from ..utils.trace import nvtx_range
        self.model.reduce_examplars()
        ps = output.get("proto_set")
        if ps is not None:
            with nvtx_range("fedstil/herding"):
                self.model.build_examplars(ps["protos"], ps["pids"], ps["cids"], tr_loader.dataset.person_ids)
            output.pop("proto_set", None)
        if self._task_tokens:
            self.task_token = torch.stack(self._task_tokens).mean(0)


class TokenFileWriter:
    """Background writer of the growing ``{server}_tokens.ckpt`` (``fedstil.py:1096`` re-writes the whole history every
    round: 8 x 512 KB more per round for ResNet-50, 200 MB at round 50).

    * **latest wins**: every request covers the whole history, so a request that arrives while a write is in flight
      replaces the one still waiting - at most one write runs and one waits, the backlog cannot grow with the round
      count; the future handed out by :meth:`submit` completes when nothing is pending any more;
    * **append-only host cache**: token tensors are never modified once appended, so each write copies only the tokens
      it has not seen yet to the host (identity-checked against the previous snapshot, reset when the history was
      replaced, e.g. by a resume)."""

    def __init__(self, device: Optional[torch.device] = None):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.device = device
        self._lock = threading.Lock()
        self._latest = None
        self._running = False
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="flpr-tokens")
        self._cache: Dict[str, List[Tuple[torch.Tensor, torch.Tensor]]] = {}       # name -> [(source, host copy)]
        self.writes = self.replaced = self.copied = 0

    def submit(self, snap: Dict[str, List[torch.Tensor]], ready, path: str):
        """Returns a future when a drain task had to be started, ``None`` when a running one will pick the request up."""
        with self._lock:
            if self._latest is not None:
                self.replaced += 1
            self._latest = (snap, ready, path)
            if self._running:
                return None
            self._running = True
        return self._pool.submit(self._drain)

    def _drain(self) -> None:
        try:
            while True:
                with self._lock:
                    req, self._latest = self._latest, None
                    if req is None:
                        self._running = False
                        return
                self._write(*req)
        except BaseException:
            with self._lock:
                self._running = False
            raise

    def _write(self, snap, ready, path: str) -> None:
        import os
        if ready is not None:
            ready.synchronize()
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        out = {}
        for name, toks in snap.items():
            have = self._cache.get(name, [])
            if len(have) > len(toks) or any(h[0] is not t for h, t in zip(have, toks)):
                have = []                                   # the history was replaced, not appended to
            for t in toks[len(have):]:
                have.append((t, t.detach().to("cpu", copy=True)))
                self.copied += 1
            self._cache[name] = have
            out[name] = [h[1] for h in have]
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(out, path + ".tmp")
        os.replace(path + ".tmp", path)
        self.writes += 1


class Server(ServerModule):
    def __init__(self, server_name, model, operator, ckpt_root, distance_calculate_step: int = 10,
                 distance_calculate_decay: float = 0.8, **kwargs):
        super().__init__(server_name, model, operator, ckpt_root, **kwargs)
        self.token_memory: Dict[str, List[torch.Tensor]] = {}
        self.distance_calculate_step = int(distance_calculate_step)
        self.distance_calculate_decay = float(distance_calculate_decay)
        self.client_ids: Dict[str, int] = {}
        self.local_clients: Dict[str, Client] = {}
        self.uploaded: List[int] = []
        self._round_uploads: List[str] = []
        self._delivered: Dict[str, bool] = {}

    def bind_client(self, client_name: str, client_id: int, client: Optional[Client] = None) -> None:
        self.client_ids[client_name] = client_id
        if client is not None:
            self.local_clients[client_name] = client

    def save_model(self, model_name: str) -> None:
        self.save_state(model_name, self.model.model_state(copy=False), True)

    def resume_extra(self) -> Dict[str, Any]:
        return {"token_memory": {k: list(v) for k, v in self.token_memory.items()}}

    def load_resume_extra(self, extra: Dict[str, Any]) -> None:
        dev = self.model.device
        self.token_memory = {k: [t.to(dev) for t in v] for k, v in (extra.get("token_memory") or {}).items()}

    # ---- uploads -----------------------------------------------------------------------------------------------------------
    def set_client_incremental_state(self, client_name: str, client_state: Optional[Dict]) -> None:
        if client_name not in self.clients:
            self.logger.warn(f"Collect incremental state failed from unregistered client {client_name}.")
            return
        self.clients[client_name] = client_state if client_state is not None else {"remote": True}
        cid = self.client_ids[client_name]
        if cid not in self.uploaded:
            self.uploaded.append(cid)
        self._round_uploads.append(client_name)
        self.logger.info(f"Collect incremental state successfully from client {client_name}.")

    set_client_integrated_state = set_client_incremental_state

    split_calculate = True      # ``calculate() == calculate_urgent() + calculate_deferred()`` (sub-classes that
                                # override ``calculate`` must clear it)

    def calculate(self) -> Any:
        """FedAvg mean of theta into the server model + token exchange (``fedstil.py:1075-1096``)."""
        self.calculate_urgent()
        self.calculate_deferred()

    # The round loop may split the aggregation (``engine_opts.overlap_aggregate``): what the NEXT dispatch depends on -
    # the task tokens, a few hundred KB - stays on the compute stream; the 125 MB FedAvg mean into the server replica
    # (needed only by first-contact dispatches and the server checkpoint) runs on the communication stream, on its own
    # flag channel, concurrently with the next round's mix and local training.
    def calculate_urgent(self) -> None:
        if self._round_uploads and "token" in self.comm.bufs:
            ids = [self.client_ids[n] for n in self._round_uploads]
            d = self.comm.bufs["token"].n
            out = torch.empty(d, len(ids), device=self.model.device)
            self.comm.gather_strided("token", ids, out)         # tokens of this round's uploads, on every rank
            toks = out.t().contiguous()
            for i, name in enumerate(self._round_uploads):
                self.token_memory.setdefault(name, []).append(toks[i])
        self._round_uploads = []
        self._save_tokens()

    def _save_tokens(self) -> None:
        """``{server}_tokens.ckpt`` (fedstil.py:1096). The token history grows by one tensor per client and round, so
        its file is re-laid out every time; token tensors are never modified once appended, which makes the snapshot a
        list of references - on CUDA it is written by a background thread of rank 0 (the server role is replicated:
        every rank would write the same bytes), off the round's critical path."""
        comm = self.comm
        if comm is not None and getattr(comm, "rank", 0) != 0:
            return
        dev = self.model.device
        if dev.type != "cuda" or not self.store.enabled or getattr(self.store, "muted", False):
            self.save_state(f"{self.server_name}_tokens", self.token_memory, True)
            return
        snap = {k: list(v) for k, v in self.token_memory.items()}
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        writer = getattr(self, "_token_writer", None)
        if writer is None:
            writer = self._token_writer = TokenFileWriter(dev)
        fut = writer.submit(snap, ready, self.store.path(self.name, f"{self.server_name}_tokens"))
        if fut is not None:
            self.store.track(fut)                          # ``store.flush()`` waits for it like for any snapshot

    def calculate_deferred(self) -> None:
        if self.uploaded:
            self.comm.reduce_bcast("theta_up", "glob", self.uploaded, cnt="cnt")
            self.model.set_global_weight(self.comm.rank_view("glob"))

    # ---- spatial-temporal integration (fedstil.py:1118-1164) ----------------------------------------------------------------
    def relevance_row(self, client_name: str) -> Tuple[List[str], torch.Tensor]:
        """Mixing weights of ``client_name`` over ``select_client`` (others in memory order, then itself)."""
        own = self.token_memory[client_name][-1].unsqueeze(0)
        select, rel = [], []
        for c_name, c_tokens in self.token_memory.items():
            if c_name == client_name:
                continue
            hist = c_tokens[::-1 * self.distance_calculate_step]
            dis = own.new_full((), 1e-8)
            for decay_cnt, other in enumerate(hist):
                dis = dis + kl_distance(own, other.unsqueeze(0)) / math.pow(self.distance_calculate_decay, decay_cnt)
            select.append(c_name)
            rel.append(1.0 / dis)
        if not rel:
            return [client_name], own.new_ones(1)
        rel = torch.stack(rel)
        rel = torch.cat([rel, rel.mean().view(1)])
        select.append(client_name)
        rel = rel / rel.sum()
        return select, torch.softmax(rel, dim=0)

    def relevance_rows(self, client_names: Sequence[str]) -> Tuple[List[str], torch.Tensor]:
        """Mixing weights of several receiving clients at once: ``(memory order, W [len(client_names), N])``.
        Same arithmetic as :meth:`relevance_row` (decayed KL of the task tokens, inverse, own = mean of the others,
        normalise, softmax) evaluated as a handful of batched tensor ops instead of ``R x N x T`` small ones."""
        order = list(self.token_memory.keys())
        N, R = len(order), len(client_names)
        dev = self.token_memory[order[0]][-1].device
        if N == 1:
            return order, torch.ones(R, 1, device=dev)
        P = torch.stack([self.token_memory[n][-1] for n in client_names]).float()          # [R, D]
        ent, owner, wgt = [], [], []
        for j, name in enumerate(order):
            for t, tok in enumerate(self.token_memory[name][::-1 * self.distance_calculate_step]):
                ent.append(tok)
                owner.append(j)
                wgt.append(1.0 / math.pow(self.distance_calculate_decay, t))
        Q = torch.stack(ent).float()                                                        # [E, D]
        logP = F.log_softmax(P, dim=-1)
        logQ = F.log_softmax(Q, dim=-1)
        Qs = logQ.exp()
        A = (Qs * logQ).sum(-1)                                                             # [E]
        kl = torch.stack([A - (Qs * logP[i]).sum(-1) for i in range(R)])                    # [R, E]
        owner_t = torch.tensor(owner, device=dev)
        w_t = torch.tensor(wgt, device=dev, dtype=kl.dtype)
        dis = torch.full((R, N), 1e-8, device=dev, dtype=kl.dtype).index_add_(1, owner_t, kl * w_t[None, :])
        rel = 1.0 / dis
        me = torch.tensor([order.index(n) for n in client_names], device=dev)
        mask = F.one_hot(me, N).bool()
        others_mean = rel.masked_fill(mask, 0.0).sum(1, keepdim=True) / (N - 1)
        rel = torch.where(mask, others_mean, rel)
        rel = rel / rel.sum(1, keepdim=True)
        return order, torch.softmax(rel, dim=1)

    def prepare_dispatch(self, online_names: Sequence[str], first_contact: Sequence[str]) -> None:
        """Collective: every rank mixes for its local, already-registered online clients in ONE kernel."""
        self._delivered = {}
        if not self.uploaded:
            return
        recv = [n for n in online_names if n not in first_contact and n in self.local_clients
                and n in self.token_memory]
        K = len(self.uploaded)
        col = {cid: j for j, cid in enumerate(self.uploaded)}
        dev = self.model.device
        rows = torch.zeros(len(recv), K, device=dev)
        if recv:
            order, W = self.relevance_rows(recv)
            cols = torch.tensor([col[self.client_ids[c]] for c in order], device=dev)
            rows[:, cols] = W.to(dev)
            if self.logger.enabled_for_info():
                flat = host_list(W)                               # ONE host sync for the whole mixing matrix
                for name, wr in zip(recv, flat):
                    for c_name, wv in zip(order, wr):
                        self.logger.info(f"Relevant ratio between {name} and {c_name}: {wv:.4f}")
        models = [self.local_clients[n].model for n in recv]
        n_theta = self.model.theta_numel
        self.comm.mix("theta_up", self.uploaded, rows, list(range(len(recv))),
                      dst_g=[m.G for m in models],
                      dst_theta=[m.arena.master[:n_theta] for m in models],
                      dst_bf16=[m.arena.shadow[:n_theta] for m in models] if models and models[0].arena.shadow
                      is not None else None)
        for n in recv:
            self._delivered[n] = True

    def get_dispatch_incremental_state(self, client_name: str) -> Optional[Dict]:
        if client_name not in self.local_clients or not self._delivered.get(client_name):
            return None
        cl = self.local_clients[client_name]
        return {"incremental_shared_params": cl._named_theta(cl.model.G), "_delivered": True}

    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        ms = self.model.model_state()
        return {"integrated_global_weight": ms["global_weight"], "integrated_bn_params": ms["bn_params"],
                "integrated_pre_trained_params": ms["pre_trained_params"]}
