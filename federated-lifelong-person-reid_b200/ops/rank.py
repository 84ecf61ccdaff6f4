"""Gallery-vs-query ranking on device: one similarity GEMM + a sort-free CMC / AP kernel.

Replaces the per-query Python loop of ``tools/evaluate.py:88-142`` (one ``torch.mm`` + ``.cpu()`` + ``np.argsort`` +
numpy set ops per query) with two kernel launches and a single device->host copy of the final numbers.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import native
from .gemm import gemm


def similarity(query: torch.Tensor, gallery: torch.Tensor, precise: bool = True) -> torch.Tensor:
    """``S[q, g] = <query_q, gallery_g>`` in fp32. With ``precise`` the bf16 tensor-core GEMM is run on a
    hi/lo split of both operands (3 passes) which recovers ~fp32 accuracy for L2-normalised features."""
    if not query.is_cuda:
        return query.float() @ gallery.float().t()
    q, g = query.float().contiguous(), gallery.float().contiguous()
    k = q.shape[1]
    if k % 8:  # TMA needs 16-byte row strides
        pad = 8 - k % 8
        q = torch.nn.functional.pad(q, (0, pad))
        g = torch.nn.functional.pad(g, (0, pad))
    q_hi, g_hi = q.to(torch.bfloat16), g.to(torch.bfloat16)
    s = gemm(q_hi, g_hi, out_dtype=torch.float32)
    if precise:
        q_lo = (q - q_hi.float()).to(torch.bfloat16)
        g_lo = (g - g_hi.float()).to(torch.bfloat16)
        s += gemm(q_lo, g_hi, out_dtype=torch.float32)
        s += gemm(q_hi, g_lo, out_dtype=torch.float32)
    return s


def rank_metrics_reference(sim: torch.Tensor, q_labels: torch.Tensor, g_labels: torch.Tensor
                           ) -> Tuple[np.ndarray, float]:
    """Straightforward (sorting) CPU implementation with the semantics of ``tools/evaluate.py``."""
    sim = sim.detach().cpu().float().numpy()
    ql = q_labels.cpu().numpy()
    gl = g_labels.cpu().numpy()
    nq, ng = sim.shape
    total_cmc = np.zeros(ng, dtype=np.float64)
    total_ap = 0.0
    for i in range(nq):
        order = np.argsort(sim[i], kind="stable")[::-1]
        hits = np.flatnonzero(gl[order] == ql[i])
        if hits.size == 0:
            continue
        total_cmc[hits[0]:] += 1
        ap = 0.0
        for j, loc in enumerate(hits):
            precision = (j + 1) / (loc + 1)
            old = j / loc if loc != 0 else 1.0
            ap += (old + precision) / 2 / hits.size
        total_ap += ap
    return total_cmc / nq, total_ap / nq


def rank_metrics(sim: torch.Tensor, q_labels: torch.Tensor, g_labels: torch.Tensor) -> Tuple[np.ndarray, float]:
    """CMC curve (length G, float64) and mAP from a similarity matrix. Queries with no match are skipped but still
    count in the denominator (``tools/evaluate.py:137-142``)."""
    if not sim.is_cuda:
        return rank_metrics_reference(sim, q_labels, g_labels)
    nq, ng = sim.shape
    lib = native.load()
    sim = sim.float().contiguous()
    ql = q_labels.to(sim.device).long().contiguous()
    gl = g_labels.to(sim.device).long().contiguous()
    ap = torch.empty(nq, dtype=torch.float32, device=sim.device)
    first = torch.empty(nq, dtype=torch.int32, device=sim.device)
    rc = lib.flpr_rank_eval(native.ptr(sim), native.ptr(ql), native.ptr(gl), native.ptr(ap), native.ptr(first), nq, ng,
                            sim.stride(0), native.stream(sim.device))
    native.check(rc, "flpr_rank_eval")
    native.count_launch()
    valid = first >= 0
    hist = torch.bincount(first[valid].long(), minlength=ng).double()
    cmc = torch.cumsum(hist, 0) / nq
    m_ap = ap.double().sum() / nq
    out = torch.cat([cmc, m_ap.view(1)]).cpu().numpy()  # single D2H
    return out[:-1], float(out[-1])


def evaluate(query_features: torch.Tensor, query_labels: torch.Tensor, gallery_features: torch.Tensor,
             gallery_labels: torch.Tensor, device: str | torch.device | None = None, precise: bool = True
             ) -> Tuple[np.ndarray, float]:
    """Drop-in for ``tools.evaluate.evaluate`` (no camera-junk filtering: the reference never passes cameras)."""
    if device is not None:
        query_features = query_features.to(device)
        gallery_features = gallery_features.to(device)
    sim = similarity(query_features, gallery_features, precise=precise)
    return rank_metrics(sim, query_labels, gallery_labels)
