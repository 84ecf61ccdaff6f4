"""``icarl`` – iCaRL-style class-incremental rehearsal, local-only (reference ``methods/icarl.py``).

* the classifier is replaced by an ``n_classes``-wide Linear that *grows* to ``max(person_id) + 1`` at the start of
  every ``train()`` (``icarl.py:52-57,68-84,459-461``);
* exemplar memory of raw (augmented) image tensors chosen by herding on the backbone's pooled features,
  ``m = ceil(k / n_classes)`` per identity (``icarl.py:64-66,97-151``);
* each epoch: (1) distillation pass over the exemplars – ``BCE(score, onehot) + BCE(score[:, :prev], sigmoid(prev
  logits))`` – then (2) the ordinary criterion over ``exemplars U new task`` (``icarl.py:216-248``).

Device-resident differences: exemplars are a bank of unique images + an ordered index list; herding is batched on the
GPU; growing the classifier re-materialises the parameter arena (the optimizer state is wiped after every ``train()``
anyway). Reference quirk: its distillation pass pairs a reshuffled exemplar loader with logits recorded in a
different shuffle order (``icarl.py:86-95,219-223``), i.e. every exemplar is distilled towards the recorded logits of a
*random* exemplar. ``reference_compat`` (default) reproduces that pairing with two independent permutations; with
``reference_compat: false`` logits and exemplars stay aligned (textbook iCaRL).
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..runtime.arena import ArenaOptimizer
from ..runtime.modules import ClientModule, ModelModule, OperatorModule, ServerModule
from ..utils.misc import get_one_hot
from .fedstil import group_matrix, herding_select_batched


class Model(ModelModule):
    def __init__(self, net, operator=None, k: float = 8000, n_classes: int = 10, **kwargs):
        super().__init__(net, **kwargs)
        self.operator = operator
        self.k = k
        self.n_classes = int(n_classes)
        bias = net.classifier.bias is not None
        net.classifier = nn.Linear(net.classifier.in_features, self.n_classes, bias)
        self.examplars: Dict[int, Dict[str, Any]] = {}
        self.previous_logits: Optional[torch.Tensor] = None
        self._mat_args: Tuple = ("cpu", "bf16", None)

    @property
    def m(self) -> int:
        return math.ceil(self.k / self.n_classes)

    def materialize(self, device, compute_dtype="bf16", fine_tuning=None):
        self._mat_args = (device, compute_dtype, fine_tuning)
        return super().materialize(device, compute_dtype, fine_tuning)

    def add_n_classes(self, n: int) -> bool:
        """Grow the classifier by ``n`` rows keeping the learned ones (``icarl.py:68-84``)."""
        if n <= 0:
            return False
        old = self.net.classifier
        self.n_classes += n
        new = nn.Linear(old.in_features, self.n_classes, old.bias is not None).to(self.device)
        with torch.no_grad():
            new.weight[: self.n_classes - n] = old.weight.detach()
            if old.bias is not None:
                new.bias[: self.n_classes - n] = old.bias.detach()
        self.net.classifier = new
        # every parameter currently views the old arena: detach them to own storage, then rebuild the arena
        for _, p in self.net.named_parameters():
            if p.requires_grad:
                p.data = p.data.clone(memory_format=torch.contiguous_format)
                p.grad = None
        self.materialize(*self._mat_args)
        return True

    # ---- exemplar memory ----------------------------------------------------------------------------------------
    def examplar_tensors(self) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
        imgs, ids = [], []
        for pid, ex in self.examplars.items():
            if ex["order"]:
                o = torch.tensor(ex["order"], device=ex["bank"].device)
                imgs.append(ex["bank"][o])
                ids.append(torch.full((len(o),), pid, dtype=torch.long, device=ex["bank"].device))
        if not imgs:
            return None
        return torch.cat(imgs), torch.cat(ids)

    def reduce_examplars(self) -> None:
        m = self.m
        for pid in self.examplars:
            self.examplars[pid]["order"] = self.examplars[pid]["order"][:m]

    @torch.no_grad()
    def build_previous_logits(self, batch_size: int = 64) -> None:
        ex = self.examplar_tensors()
        if ex is None:
            return
        self.net.train()                                          # as in the reference (icarl.py:89)
        self._mode_uniform = None                                 # (sub-modules toggled behind ModelModule.train)
        outs = []
        for s in range(0, ex[0].shape[0], batch_size):
            with self.autocast():
                score, _ = self.net(self.prepare_input(ex[0][s:s + batch_size].float()))
            outs.append(score.float())
        self.previous_logits = torch.cat(outs)

    @torch.no_grad()
    def build_examplars(self, dataloader, batch_size: int = 64) -> None:
        """Herding over ``exemplars U task`` restricted to the task's identities (``icarl.py:97-139``)."""
        self.eval()
        imgs, ids = [], []
        ex = self.examplar_tensors()
        if ex is not None:
            imgs.append(ex[0]); ids.append(ex[1])
        for data, person_id, _ in dataloader:
            imgs.append(self.prepare_input(data).to(torch.bfloat16 if self.compute_dtype == torch.bfloat16
                                                    else torch.float32))
            ids.append(person_id.to(self.device))
        imgs, ids = torch.cat([i.to(imgs[-1].dtype) for i in imgs]), torch.cat(ids)
        feats = []
        for s in range(0, imgs.shape[0], batch_size):
            with self.autocast():
                feats.append(self.net(imgs[s:s + batch_size].float()).float())
        feats = torch.cat(feats)
        keep = set(int(p) for p in dataloader.dataset.person_ids)
        upids = [p for p in torch.unique(ids).tolist() if p in keep]
        if not upids:
            return
        groups = [torch.nonzero(ids == pid).squeeze(1) for pid in upids]
        picks = herding_select_batched(feats, *group_matrix(groups), self.m).cpu()
        for gi, pid in enumerate(upids):
            uniq, inverse = torch.unique(picks[gi], return_inverse=True)
            sel = groups[gi][uniq.to(groups[gi].device)]
            self.examplars[int(pid)] = {"bank": imgs[sel].clone(), "order": inverse.tolist()}

    # ---- checkpoint schema {'net_params', 'examplars'} (icarl.py:178-193) ---------------------------------------------
    def model_state(self, *args, **kwargs) -> Dict:
        net = {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in self.net.state_dict().items()}
        return {"net_params": net, "examplars": self.examplars_state()}

    def examplars_state(self, max_bytes: int = 1 << 30) -> Dict:
        """``{np.int64 pid: [(ndarray image, pid), ...]}`` – the reference's exemplar memory (``icarl.py:135-139``).
        Beyond ``max_bytes`` of fp32 images (the reference writes ~3 GB per client and round at ``k = 8000``,
        256x128) the compact ``{pid: {"bank", "order"}}`` form is written instead; both forms load."""
        n = sum(len(e["order"]) for e in self.examplars.values())
        per = next((e["bank"][0].numel() * 4 for e in self.examplars.values() if len(e["bank"])), 0)
        if n * per > max_bytes:
            return {int(p): {"bank": e["bank"], "order": list(e["order"])} for p, e in self.examplars.items()}
        out = {}
        for pid, e in self.examplars.items():
            bank = e["bank"].float().cpu().numpy()
            out[np.int64(pid)] = [(bank[i], np.int64(pid)) for i in e["order"]]
        return out

    def update_model(self, params_state: Dict) -> None:
        if "net_params" in params_state:
            sd = params_state["net_params"]
            w = sd.get("classifier.weight")
            if w is not None and w.shape[0] != self.n_classes:
                self.add_n_classes(w.shape[0] - self.n_classes)
            own = self.net.state_dict()
            with torch.no_grad():
                for k, v in sd.items():
                    if k in own and own[k].shape == v.shape:
                        own[k].copy_(v.to(own[k].device))
            if self.arena is not None:
                self.arena.refresh_shadow()
        if "examplars" in params_state:
            dt = torch.bfloat16 if self.compute_dtype == torch.bfloat16 else torch.float32
            self.examplars = {}
            for p, e in params_state["examplars"].items():
                if isinstance(e, dict):
                    self.examplars[int(p)] = {"bank": e["bank"].to(self.device), "order": list(e["order"])}
                elif len(e):                                       # reference form: [(image ndarray, pid), ...]
                    bank = torch.stack([torch.as_tensor(img) for img, _ in e]).to(self.device, dt)
                    self.examplars[int(p)] = {"bank": bank, "order": list(range(len(e)))}


class Operator(OperatorModule):
    def rebind(self, model: Model) -> None:
        """Point the optimizer at the (possibly re-materialised) arena; keeps the hyper-parameters."""
        opt = self.optimizer
        if opt.arena is not model.arena:
            d = opt.defaults
            new = ArenaOptimizer(opt.kind, model.arena, lr=d["lr"], weight_decay=d["weight_decay"], betas=d["betas"],
                                 eps=d["eps"], momentum=d["momentum"])
            new.lr = opt.lr
            new.sync_hyper()
            self.optimizer = new
            if self.scheduler is not None:
                self.scheduler.optimizer = new

    def invoke_train(self, model: Model, dataloader, **kwargs) -> Dict:
        device = model.device
        model.train()
        self.rebind(model)
        self.begin_epoch()
        bs = dataloader.batch_size
        ex = model.examplar_tensors()
        # (1) distillation pass on the exemplars
        if model.previous_logits is not None and len(model.previous_logits) and ex is not None:
            n_ex = ex[0].shape[0]
            prev = model.previous_logits
            perm = torch.randperm(min(n_ex, prev.shape[0]), device=device)
            perm_logits = torch.randperm(len(perm), device=device) if getattr(model, "misaligned_distill", True) \
                else perm
            for s in range(0, len(perm), bs):
                idx = perm[s:s + bs]
                data, target = model.prepare_input(ex[0][idx].float()), ex[1][idx]
                pl = prev[perm_logits[s:s + bs]]
                pc = pl.shape[1]
                self.optimizer.zero_grad()
                with model.autocast():
                    score, _ = model.forward(data)
                if score.is_cuda:
                    from ..ops.fused import bce_distill          # both BCE terms + gradient in one kernel
                    loss = bce_distill(score, target, pl[:, :pc])
                else:
                    score = score.float()
                    loss = F.binary_cross_entropy_with_logits(score, get_one_hot(target, model.n_classes)) + \
                        F.binary_cross_entropy_with_logits(score[:, :pc], torch.sigmoid(pl[:, :pc]))
                loss.backward()
                self.optimizer.step()
        # (2) criterion pass over exemplars U task (ConcatDataset + shuffle)
        acc = torch.zeros(2, dtype=torch.float64, device=device)
        batch_cnt = data_cnt = 0
        n_ex = ex[0].shape[0] if ex is not None else 0
        if hasattr(dataloader, "fetch"):
            n_task = len(dataloader.dataset)
            total = n_ex + n_task
            perm = torch.randperm(total)
            nb = total // bs if total % bs == 1 else (total + bs - 1) // bs
            batches = (perm[b * bs:(b + 1) * bs] for b in range(nb))

            def materialise(idx):
                e_idx, t_idx = idx[idx < n_ex], idx[idx >= n_ex] - n_ex
                parts_d, parts_t = [], []
                if len(e_idx):
                    parts_d.append(ex[0][e_idx.to(device)].float()); parts_t.append(ex[1][e_idx.to(device)])
                if len(t_idx):
                    d, pid, _ = dataloader.fetch(t_idx)
                    parts_d.append(d.float()); parts_t.append(pid)
                return torch.cat(parts_d), torch.cat(parts_t)
            stream = (materialise(i) for i in batches)
        else:
            def gen():
                for s in range(0, n_ex, bs):
                    yield ex[0][s:s + bs].float(), ex[1][s:s + bs]
                for data, pid, _ in dataloader:
                    yield data, pid
            stream = gen()
        for data, target in stream:
            data, target = model.prepare_input(data), target.to(device)
            self.optimizer.zero_grad()
            out = self._invoke_train(model, data, target, **kwargs)
            out["loss"].backward()
            self.optimizer.step()
            with torch.no_grad():
                acc[0] += out["loss"].detach().double()
                acc[1] += (out["score"].argmax(dim=1) == target).sum()
            data_cnt += len(data)
            batch_cnt += 1
        loss_sum, hits = acc.tolist()
        if self.scheduler:
            self.scheduler.step()
        return {"accuracy": hits / max(data_cnt, 1), "loss": loss_sum / max(batch_cnt, 1), "batch_count": batch_cnt,
                "data_count": data_cnt}


class Client(ClientModule):
    def __init__(self, client_name, model, operator, ckpt_root, model_ckpt_name=None, **kwargs):
        super().__init__(client_name, model, operator, ckpt_root, model_ckpt_name, **kwargs)
        self.model.operator = operator
        self.model.misaligned_distill = bool(getattr(self, "reference_compat", True))
        self.operator.reset_lr_each_epoch = bool(getattr(self, "reference_compat", True))
        if not self.model_ckpt_name:
            self.model_ckpt_name = "icarl_model"

    def update_by_integrated_state(self, state: Dict, **kwargs) -> Any:
        # base ``update_model`` merge (modules/client.py:72-76): every ``net.*`` entry, the n_classes-wide head included
        # (server and client heads have the configured width at first contact; other widths are skipped by shape)
        sd = {k[4:] if k.startswith("net.") else k: v for k, v in state["model_params"].items()}
        w = sd.get("classifier.weight")
        if w is not None and w.shape[0] != self.model.n_classes:
            sd = {k: v for k, v in sd.items() if not k.startswith("classifier.")}
        self.model.update_model({"net_params": sd})
        self.logger.info("Update model succeed by integrated state from server.")

    update_by_incremental_state = update_by_integrated_state

    def before_train(self, task_name, tr_loader, val_loader) -> None:
        incremental = int(max(tr_loader.dataset.person_ids)) - self.model.n_classes + 1
        self.model.build_previous_logits(tr_loader.batch_size)
        self.model.add_n_classes(incremental)
        self.operator.rebind(self.model)

    def after_train(self, task_name, tr_loader, val_loader, output) -> None:
        self.model.reduce_examplars()
        self.model.build_examplars(tr_loader)


class Server(ServerModule):
    def get_dispatch_integrated_state(self, client_name: str) -> Dict:
        sd = self.model.full_state()
        # the reference's Model registers ``features_extractor = net.base`` as a second sub-module, so its state_dict
        # lists the trunk twice (``icarl.py:59,569-573``); same tensors, alias keys
        sd.update({"features_extractor." + k[len("net.base."):]: v for k, v in list(sd.items())
                   if k.startswith("net.base.")})
        return {"model_params": sd}
