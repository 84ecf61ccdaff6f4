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
    if not native.on_device(sim, "flpr_rank_eval"):
        return rank_metrics_reference(sim, q_labels, g_labels)
    nq, ng = sim.shape
    lib = native.kernels()
    sim = sim.float().contiguous()
    ql = q_labels.to(sim.device).long().contiguous()
    gl = g_labels.to(sim.device).long().contiguous()
    ap = torch.empty(nq, dtype=torch.float32, device=sim.device)
    first = torch.empty(nq, dtype=torch.int32, device=sim.device)
    rc = lib.flpr_rank_eval(native.ptr(sim), native.ptr(ql), native.ptr(gl), native.ptr(ap), native.ptr(first), nq, ng,
                            sim.stride(0), native.stream_of(sim.device))
    native.check(rc, "flpr_rank_eval")
    native.count_launch()
    valid = first >= 0
    hist = torch.bincount(first[valid].long(), minlength=ng).double()
    cmc = torch.cumsum(hist, 0) / nq
    m_ap = ap.double().sum() / nq
    out = torch.cat([cmc, m_ap.view(1)]).cpu().numpy()  # single D2H
    return out[:-1], float(out[-1])


def rank_metrics_with_cameras(sim: torch.Tensor, q_labels: torch.Tensor, g_labels: torch.Tensor,
                              q_cameras: torch.Tensor, g_cameras: torch.Tensor) -> Tuple[np.ndarray, float]:
    """CMC / mAP with the reference's junk rule (``tools/evaluate.py:12-33,60-67``): gallery items of the query's
    identity *seen by the query's camera*, and mis-detections (label ``-1``), are removed from the ranking; the hits
    are the same identity under another camera. Tensor ops on ``sim``'s device (counts instead of a sort per query):

        position(hit) = #{non-junk g : S[q, g] > s_hit},   index among hits = #{hits h : s_h > s_hit}
    """
    dev = sim.device
    ql, gl = q_labels.to(dev).long().view(-1, 1), g_labels.to(dev).long().view(1, -1)
    qc, gc = q_cameras.to(dev).long().view(-1, 1), g_cameras.to(dev).long().view(1, -1)
    nq, ng = sim.shape
    same_id, same_cam = gl == ql, gc == qc
    junk = (same_id & same_cam) | (gl == -1)
    hit = same_id & ~same_cam
    neg = torch.finfo(torch.float32).min
    s = sim.float()
    keep = torch.where(junk, torch.full_like(s, neg), s)                 # junk can never outrank anything
    s_sorted, _ = torch.sort(keep, dim=1)
    h_sorted, _ = torch.sort(torch.where(hit, s, torch.full_like(s, neg)), dim=1)
    loc = (ng - torch.searchsorted(s_sorted, s.contiguous(), right=True)).double()      # [Q, G], valid where hit
    j = (ng - torch.searchsorted(h_sorted, s.contiguous(), right=True)).double()
    n_hits = hit.sum(1).double()
    precision = (j + 1) / (loc + 1)
    old = torch.where(loc > 0, j / loc.clamp(min=1), torch.ones_like(loc))
    ap = torch.where(hit, (old + precision) / 2, torch.zeros_like(loc)).sum(1) / n_hits.clamp(min=1)
    has = n_hits > 0
    first = torch.where(hit, loc, torch.full_like(loc, float(ng))).min(1).values.long()
    hist = torch.bincount(first[has], minlength=ng)[:ng].double()
    # the reference's per-query curve keeps the full gallery length even after junk removal (evaluate.py:53,72)
    cmc = torch.cumsum(hist, 0) / nq
    out = torch.cat([cmc, ((ap * has).sum() / nq).view(1)]).cpu().numpy()
    return out[:-1], float(out[-1])


def evaluate(query_features: torch.Tensor, query_labels: torch.Tensor, gallery_features: torch.Tensor,
             gallery_labels: torch.Tensor, query_camera_labels: torch.Tensor | None = None,
             gallery_camera_labels: torch.Tensor | None = None, device: str | torch.device | None = None,
             precise: bool = True) -> Tuple[np.ndarray, float]:
    """Drop-in for ``tools.evaluate.evaluate`` (same positional order). Camera labels are optional - the reference's
    runtime never passes them - and switch on its same-camera / mis-detection junk rule."""
    if device is not None:
        query_features = query_features.to(device)
        gallery_features = gallery_features.to(device)
    sim = similarity(query_features, gallery_features, precise=precise)
    if query_camera_labels is not None and gallery_camera_labels is not None:
        return rank_metrics_with_cameras(sim, query_labels, gallery_labels, query_camera_labels,
                                         gallery_camera_labels)
    return rank_metrics(sim, query_labels, gallery_labels)


# ------------------------------------------------------------------------------------------------- gallery-sharded
def evaluate_sharded(query_features: torch.Tensor, query_labels: torch.Tensor, gallery_shard: torch.Tensor,
                     gallery_shard_labels: torch.Tensor, group=None, precise: bool = True
                     ) -> Tuple[np.ndarray, float]:
    """Exact CMC / mAP when the **gallery is sharded over the ranks** of ``group`` (the dimension that
    scales in this workload is the gallery, the analogue of sequence parallelism is to give every rank a slice of it).

    Every rank scores its shard against all queries. A hit's position in the global ranking is the number of gallery
    items that score higher, and that count is a *sum over shards* - so instead of gathering the ``Q x G`` similarity
    matrix (or merging sorted lists) the ranks exchange only the similarities of the hits (``Q x P``, ``P`` = matches per
    query and shard) and all-reduce two small count tensors:

        above[r, q, p]     = #{g in shard_r : S[q, g] > s_p}            -> position of hit p in the full ranking
        hits_above[r, q, p] = #{g in shard_r : S[q, g] > s_p, g a hit}  -> index of hit p among the hits

    from which AP (the reference's trapezoid form, ``tools/evaluate.py:75-82``) and the first-hit position (CMC) follow
    locally and identically on every rank. Returns what :func:`evaluate` returns on the concatenated gallery.
    Without an initialised process group (or ``world == 1``) this *is* :func:`evaluate`.
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return evaluate(query_features, query_labels, gallery_shard, gallery_shard_labels, precise=precise)
    world = dist.get_world_size(group)
    dev = query_features.device
    sim = similarity(query_features, gallery_shard, precise=precise)                   # [Q, Gr]
    ql = query_labels.to(dev).long()
    gl = gallery_shard_labels.to(dev).long()
    nq, ngr = sim.shape
    match = gl.view(1, -1) == ql.view(-1, 1)                                           # [Q, Gr]
    # ---- exchange the hit similarities (padded to the largest per-shard hit count) -----------------------------------
    sizes = torch.tensor([int(match.sum(1).max()) if ngr else 0, ngr], device=dev, dtype=torch.long)
    size_list = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(size_list, sizes, group=group)
    pmax = max(int(s[0]) for s in size_list)
    ng_total = sum(int(s[1]) for s in size_list)
    if pmax == 0:
        return np.zeros(ng_total, dtype=np.float64), 0.0
    neg = torch.finfo(torch.float32).min
    # hits first (stable), then padding
    order = torch.argsort((~match).to(torch.int8), dim=1, stable=True)[:, :pmax] if ngr else \
        torch.zeros(nq, 0, dtype=torch.long, device=dev)
    mine = torch.full((nq, pmax), neg, dtype=torch.float32, device=dev)
    valid = torch.zeros(nq, pmax, dtype=torch.bool, device=dev)
    if ngr:
        k = order.shape[1]
        valid[:, :k] = match.gather(1, order)
        mine[:, :k] = torch.where(valid[:, :k], sim.gather(1, order), torch.full_like(mine[:, :k], neg))
    hit_list = [torch.empty_like(mine) for _ in range(world)]
    val_list = [torch.empty(nq, pmax, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(hit_list, mine.contiguous(), group=group)
    dist.all_gather(val_list, valid.to(torch.uint8).contiguous(), group=group)
    hits = torch.stack(hit_list, 1).reshape(nq, world * pmax)                           # [Q, W*P]
    hvalid = torch.stack(val_list, 1).reshape(nq, world * pmax).bool()
    # ---- local counts: sorted shard rows + searchsorted ---------------------------------------------------------------
    if ngr:
        s_sorted, _ = torch.sort(sim, dim=1)                                            # ascending
        above = ngr - torch.searchsorted(s_sorted, hits.contiguous(), right=True)       # strictly greater
        pos_only = torch.where(match, sim, torch.full_like(sim, neg))
        p_sorted, _ = torch.sort(pos_only, dim=1)
        hits_above = ngr - torch.searchsorted(p_sorted, hits.contiguous(), right=True)
    else:
        above = torch.zeros_like(hits, dtype=torch.long)
        hits_above = torch.zeros_like(hits, dtype=torch.long)
    counts = torch.stack([above, hits_above]).to(torch.int64)
    dist.all_reduce(counts, group=group)
    loc, j = counts[0].double(), counts[1].double()                                     # 0-based rank / hit index
    # ---- AP and first hit ----------------------------------------------------------------------------------------------
    n_hits = hvalid.sum(1).double()
    precision = (j + 1) / (loc + 1)
    old = torch.where(loc > 0, j / loc.clamp(min=1), torch.ones_like(loc))
    ap = (torch.where(hvalid, (old + precision) / 2, torch.zeros_like(loc)).sum(1) / n_hits.clamp(min=1))
    has = n_hits > 0
    first = torch.where(hvalid, loc, torch.full_like(loc, float(ng_total))).min(1).values.long()
    hist = torch.bincount(first[has], minlength=ng_total)[:ng_total].double()
    cmc = torch.cumsum(hist, 0) / nq
    m_ap = (ap * has).sum() / nq
    out = torch.cat([cmc, m_ap.view(1)]).cpu().numpy()
    return out[:-1], float(out[-1])
