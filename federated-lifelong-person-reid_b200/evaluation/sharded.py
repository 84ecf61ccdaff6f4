"""Collective validation: the ranking of ONE client's query set against its gallery, sharded over all ranks.

In the SPMD round loop every rank normally validates the clients it hosts on its own (``ClientModule.validate``).
``engine_opts.sharded_validation`` switches to the gallery-parallel form (the analogue of sequence
parallelism for this workload): the ranks walk the *global* (client, task) list in the same order; the rank that hosts
the client extracts the features (it owns the model) and calls the ranker, every other rank calls
:meth:`ShardedRanker.participate` for the same step. Features travel once (broadcast from the owner), every rank
scores and ranks a contiguous ``1 / world`` slice of the gallery, and :func:`~..ops.rank.evaluate_sharded` combines
the slices exactly - the ``Q x G`` similarity matrix, its sort and the hit bookkeeping are never materialised on one
device. Replaces ``tools.evaluate.evaluate`` inside ``methods/baseline.py:225-253`` of the reference (one process, one
``Q x G`` product on one device).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..ops.rank import evaluate, evaluate_sharded


class ShardedRanker:
    """``ranker(qf, ql, gf, gl)`` on the owning rank + ``ranker.participate(owner)`` on all others = one collective."""

    def __init__(self, device: torch.device, group=None):
        self.device = torch.device(device)
        self.group = group
        ready = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if ready else 0
        self.world = dist.get_world_size(group) if ready else 1
        self.calls = 0                                   # collectives executed (tests / logging)
        self.bytes_broadcast = 0

    # -- owner side: same signature as ``evaluate`` -----------------------------------------------------------------
    def __call__(self, query_features: torch.Tensor, query_labels: torch.Tensor, gallery_features: torch.Tensor,
                 gallery_labels: torch.Tensor) -> Tuple[np.ndarray, float]:
        if self.world == 1:
            return evaluate(query_features, query_labels, gallery_features, gallery_labels)
        return self._run(self.rank, query_features, query_labels, gallery_features, gallery_labels)

    # -- every other rank ---------------------------------------------------------------------------------------------
    def participate(self, owner: int) -> Optional[Tuple[np.ndarray, float]]:
        if self.world == 1:
            return None
        return self._run(int(owner), None, None, None, None)

    def _src(self, owner: int) -> int:
        return dist.get_global_rank(self.group, owner) if self.group is not None else owner

    def _run(self, owner, qf, ql, gf, gl):
        dev, src = self.device, self._src(owner)
        mine = self.rank == owner
        meta = torch.zeros(3, dtype=torch.long, device=dev)
        if mine:
            dim = qf.shape[1] if qf.numel() else (gf.shape[1] if gf.numel() else 1)
            meta = torch.tensor([len(qf), len(gf), dim], dtype=torch.long, device=dev)
        dist.broadcast(meta, src=src, group=self.group)
        nq, ng, dim = (int(v) for v in meta.tolist())
        if mine:
            qf = qf.reshape(nq, dim).to(dev, torch.float32).contiguous()
            gf = gf.reshape(ng, dim).to(dev, torch.float32).contiguous()
            ql = ql.to(dev, torch.long).contiguous()
            gl = gl.to(dev, torch.long).contiguous()
        else:
            qf = torch.empty(nq, dim, dtype=torch.float32, device=dev)
            gf = torch.empty(ng, dim, dtype=torch.float32, device=dev)
            ql = torch.empty(nq, dtype=torch.long, device=dev)
            gl = torch.empty(ng, dtype=torch.long, device=dev)
        for t in (qf, ql, gf, gl):
            if t.numel():
                dist.broadcast(t, src=src, group=self.group)
                self.bytes_broadcast += t.numel() * t.element_size()
        per = (ng + self.world - 1) // self.world
        lo, hi = min(ng, per * self.rank), min(ng, per * (self.rank + 1))
        self.calls += 1
        return evaluate_sharded(qf, ql, gf[lo:hi], gl[lo:hi], group=self.group)
