"""Layer-level fused element-wise ops (``csrc/layer_ops.cu``): decomposed-weight composition for FedWeIT /
fedstil-atten and the Swin token kernels (LayerNorm written straight into the shifted-window layout, window merge +
residual, GELU).

Every op has a plain PyTorch formulation (``*_ref``) with the same index mathematics: it is the CPU path, the yardstick
of the tests, and the yardstick of :func:`enabled` - a one-time numerics self-check that each kernel family runs on
the device before its first use (these kernels were written after the round's last GPU session; a family whose check
fails is switched off with a loud warning and its callers keep their PyTorch formulation, so a kernel bug can cost
speed but not correctness).

Reference sites: ``methods/fedweit.py:122-136`` (decomposed layer), ``methods/fedstil_atten.py:88-96`` (stacked global
weights), ``models/swin_transformer.py:118-140`` (MLP), ``:358-395`` (block: norm -> roll -> window partition ->
attention -> window reverse -> roll -> residual).
"""
from __future__ import annotations

import logging
import os
import threading
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import native

_log = logging.getLogger("flpr.ops.layer")
c_ll, c_int, c_float = native.c_ll, native.c_int, native.c_float


def _declare(lib) -> None:
    if getattr(lib, "_flpr_layer_declared", False):
        return
    P, I, Fl, L = native.c_void_p, native.c_int, native.c_float, native.c_ll
    sig = {
        "flpr_wcompose_fwd": [P, P, P, I, I, P, P, L, Fl, Fl, I, P, P, L, P],
        "flpr_wcompose_bwd": [P, P, P, I, I, P, P, L, Fl, Fl, I, P, P, P, P, L, P],
        "flpr_ln_rows": [P, P, P, P, L, I, Fl, I, I, I, I, I, P],
        "flpr_ln_rows_train": [P, P, P, P, P, L, I, Fl, I, I, I, I, I, P],
        "flpr_ln_rows_bwd": [P, P, P, P, P, P, P, L, I, I, I, I, I, I, P],
        "flpr_window_merge_add": [P, P, P, L, I, I, I, I, I, P],
        "flpr_window_merge_add_scaled": [P, P, P, P, L, I, I, I, I, I, P],
        "flpr_window_gather_scale": [P, P, P, L, I, I, I, I, I, P],
        "flpr_gelu_rows": [P, P, L, P],
        "flpr_gelu_bwd_rows": [P, P, P, L, P],
        "flpr_apply_global": [P, P, P, P, I, L, P],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = I
    lib.flpr_wcompose_max_k.argtypes = []
    lib.flpr_wcompose_max_k.restype = I
    lib.flpr_ln_rows_bwd_blocks.argtypes = [L]
    lib.flpr_ln_rows_bwd_blocks.restype = I
    lib._flpr_layer_declared = True


_emu_lib = None    # tests only: a host build of csrc/layer_ops.cu under the SIMT emulator (tests/emu), driven with CPU tensors


def use_emulated_library(path: Optional[str]):
    """Route the wrappers of this module to the emulated library at ``path`` (``None``: back to normal). With it set,
    CPU tensors take the KERNEL path (same argument marshalling, same entry points) instead of the PyTorch reference."""
    global _emu_lib
    if path is None:
        _emu_lib = None
        return None
    import ctypes
    lib = ctypes.CDLL(path)
    _declare(lib)
    _emu_lib = lib
    return lib


def _lib():
    if _emu_lib is not None:
        return _emu_lib
    lib = native.load()
    _declare(lib)
    return lib


def _native(t: torch.Tensor) -> bool:
    """The kernel path applies to ``t`` (a CUDA tensor - or any tensor while the emulated library is installed)."""
    return t.is_cuda or _emu_lib is not None


def _stream(device):
    return native.c_void_p(0) if _emu_lib is not None else native.stream(device)


WC_MAX_K = 16      # flpr_wcompose_max_k(): stacked weights per element the compose kernels keep in registers


# ===================================================================================================== physical layouts
def is_channels_last_4d(t: torch.Tensor) -> bool:
    """4-D tensor stored OHWI (``channels_last``) and not simultaneously plain-contiguous."""
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def phys_flat(t: torch.Tensor) -> torch.Tensor:
    """1-D view of ``t`` in its storage (physical) order. ``t`` must be dense: contiguous, or 4-D channels_last."""
    if t.is_contiguous():
        return t.reshape(-1)
    if is_channels_last_4d(t):
        return t.permute(0, 2, 3, 1).reshape(-1)
    raise ValueError("tensor is neither contiguous nor channels_last-dense")


def like_phys(flat: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """View the flat physical-order buffer ``flat`` with ``ref``'s logical shape and strides."""
    if is_channels_last_4d(ref):
        o, i, h, w = ref.shape
        return flat.view(o, h, w, i).permute(0, 3, 1, 2)
    return flat.view(ref.shape)


def flat_like(g: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """``g`` (logical shape of ``ref``, any strides) flattened in ``ref``'s physical order (copies only if needed)."""
    if is_channels_last_4d(ref):
        return g.permute(0, 2, 3, 1).contiguous().reshape(-1)
    return g.contiguous().reshape(-1)


def stack_phys(stack: torch.Tensor, ref: torch.Tensor) -> Optional[torch.Tensor]:
    """``stack``: ``[*ref.shape, K]``. Returns the ``[numel, K]`` view whose row order is ``ref``'s physical order, or
    ``None`` when the stack is not stored that way (see :func:`stack_aligned`)."""
    if is_channels_last_4d(ref):
        v = stack.permute(0, 2, 3, 1, 4)
        return v.reshape(-1, stack.shape[-1]) if v.is_contiguous() else None
    return stack.reshape(-1, stack.shape[-1]) if stack.is_contiguous() else None


def stack_aligned(stack: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """A tensor equal to ``stack`` (same logical shape) stored in ``ref``'s physical element order + trailing K."""
    if stack_phys(stack, ref) is not None:
        return stack
    if is_channels_last_4d(ref):
        return stack.permute(0, 2, 3, 1, 4).contiguous().permute(0, 3, 1, 2, 4)
    return stack.contiguous()


# ===================================================================================================== compose
def wcompose_fwd_ref(aw, stack, atten, kb, sw, mask, row_len, thr_aw, thr_mask, prune):
    a = aw * (aw.abs() > thr_aw).to(aw.dtype) if prune else aw
    th = a
    if kb > 0:
        th = th + (stack[:, :kb] * atten[:kb]).sum(-1)
    if sw is not None:
        m = mask * (mask.abs() > thr_mask).to(mask.dtype) if prune else mask
        th = th + m.repeat_interleave(row_len)[:aw.numel()] * sw
    return th


def wcompose_fwd(aw: torch.Tensor, stack: Optional[torch.Tensor], atten: Optional[torch.Tensor], kb: int,
                 sw: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, row_len: int = 1,
                 thr_aw: float = 0.0, thr_mask: float = 0.0, prune: bool = False, want_f32: bool = True,
                 want_bf16: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """``theta[e] = prune(aw[e]) + sum_{k<kb} atten[k] stack[e, k] + prune(mask[e // row_len]) sw[e]`` over flat fp32
    buffers; returns ``(theta_fp32, theta_bf16)`` (either may be skipped)."""
    n = aw.numel()
    if not _native(aw):
        th = wcompose_fwd_ref(aw, stack, atten, kb, sw, mask, row_len, thr_aw, thr_mask, prune)
        return (th if want_f32 else None), (th.to(torch.bfloat16) if want_bf16 else None)
    lib = _lib()
    assert aw.dtype == torch.float32 and aw.is_contiguous() and 0 <= kb <= WC_MAX_K
    ks = 0
    if kb > 0:
        assert stack.dtype == torch.float32 and stack.is_contiguous() and stack.shape[0] == n and stack.shape[1] >= kb
        assert atten.dtype == torch.float32 and atten.is_contiguous() and atten.numel() >= kb
        ks = stack.shape[1]
    if sw is not None:
        assert sw.dtype == torch.float32 and sw.is_contiguous() and sw.numel() == n
        assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.numel() * row_len >= n
    o32 = torch.empty(n, dtype=torch.float32, device=aw.device) if want_f32 else None
    o16 = torch.empty(n, dtype=torch.bfloat16, device=aw.device) if want_bf16 else None
    rc = lib.flpr_wcompose_fwd(native.ptr(aw), native.ptr(stack if kb > 0 else None),
                               native.ptr(atten if kb > 0 else None), int(kb), int(ks), native.ptr(sw),
                               native.ptr(mask if sw is not None else None), int(row_len), float(thr_aw),
                               float(thr_mask), int(bool(prune)), native.ptr(o32), native.ptr(o16), n,
                               _stream(aw.device))
    native.check(rc, "flpr_wcompose_fwd")
    native.count_launch()
    return o32, o16


def wcompose_bwd_ref(dth, aw, stack, kb, sw, mask, row_len, thr_aw, thr_mask, prune):
    d_aw = dth * (aw.abs() > thr_aw).to(dth.dtype) if prune else None
    d_att = (dth[:, None] * stack[:, :kb]).sum(0) if kb > 0 else None
    d_mask = None
    if sw is not None:
        rows = mask.numel()
        prod = dth * sw
        pad = rows * row_len - prod.numel()
        if pad:
            prod = torch.cat([prod, prod.new_zeros(pad)])
        d_mask = prod.view(rows, row_len).sum(1)
        if prune:
            d_mask = d_mask * (mask.abs() > thr_mask).to(d_mask.dtype)
    return d_aw, d_att, d_mask


def wcompose_bwd(dth: torch.Tensor, aw: torch.Tensor, stack: Optional[torch.Tensor], kb: int,
                 sw: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, row_len: int = 1,
                 thr_aw: float = 0.0, thr_mask: float = 0.0, prune: bool = False, chunk: int = 8192):
    """Gradients of :func:`wcompose_fwd` w.r.t. ``aw`` (``None`` = identical to ``dth``: nothing was pruned),
    ``atten[:kb]`` and ``mask``. Without ``sw`` the rows of the reduction are arbitrary ``chunk``-element pieces."""
    if not _native(dth):
        return wcompose_bwd_ref(dth, aw, stack, kb, sw, mask, row_len, thr_aw, thr_mask, prune)
    lib = _lib()
    n = dth.numel()
    assert dth.dtype == torch.float32 and dth.is_contiguous() and aw.is_contiguous() and 0 <= kb <= WC_MAX_K
    if sw is None:
        row_len = int(chunk)
    rows = (n + row_len - 1) // row_len
    ks = stack.shape[1] if kb > 0 else 0
    if kb > 0:
        assert stack.dtype == torch.float32 and stack.is_contiguous() and stack.shape[0] == n
    d_aw = torch.empty_like(dth) if prune else None
    part = torch.empty(rows, max(kb, 1), dtype=torch.float32, device=dth.device) if kb > 0 else None
    d_att = torch.empty(kb, dtype=torch.float32, device=dth.device) if kb > 0 else None
    d_mask = torch.empty(rows, dtype=torch.float32, device=dth.device) if sw is not None else None
    if sw is not None:
        assert mask.numel() == rows and sw.is_contiguous() and mask.is_contiguous()
    rc = lib.flpr_wcompose_bwd(native.ptr(dth), native.ptr(aw), native.ptr(stack if kb > 0 else None), int(kb),
                               int(ks), native.ptr(sw), native.ptr(mask if sw is not None else None), int(row_len),
                               float(thr_aw), float(thr_mask), int(bool(prune)), native.ptr(d_aw), native.ptr(part),
                               native.ptr(d_mask), native.ptr(d_att), n, _stream(dth.device))
    native.check(rc, "flpr_wcompose_bwd")
    native.count_launch(2 if kb > 0 else 1)
    return d_aw, d_att, d_mask


class _WComposeFn(torch.autograd.Function):
    """``theta = prune(aw) + sum_k atten_k stack[..., k] (+ prune(mask)[o] * sw)`` on the weight's own storage order.

    ``aw`` (trainable, weight-shaped) fixes the physical layout; ``sw`` and ``stack[..., k]`` must be stored in the same
    element order (:func:`stack_aligned`). Returns ``(theta_fp32, theta_bf16)`` shaped / strided like ``aw``; the bf16
    tensor is the tensor-core operand (non-differentiable), the fp32 one carries the gradient."""

    @staticmethod
    def forward(ctx, aw, mask, atten, sw, stack, kb, thr_aw, thr_mask, prune, use_ref):
        aw_f = phys_flat(aw.detach())
        st = stack_phys(stack, aw) if kb > 0 else None
        assert kb == 0 or st is not None, "stack is not stored in the weight's physical order"
        sw_f = phys_flat(sw) if sw is not None else None
        if sw_f is not None:
            assert is_channels_last_4d(sw) == is_channels_last_4d(aw) or sw.numel() == sw.shape[0], "sw / aw layouts differ"
        row_len = aw.numel() // aw.shape[0]
        at = atten.detach() if atten is not None else None
        mk = mask.detach().contiguous() if mask is not None else None
        if use_ref:
            th = wcompose_fwd_ref(aw_f, st, at, kb, sw_f, mk, row_len, thr_aw, thr_mask, prune)
            th32, th16 = th, th.to(torch.bfloat16)
        else:
            th32, th16 = wcompose_fwd(aw_f, st, at, kb, sw_f, mk, row_len, thr_aw, thr_mask, prune)
        ctx.save_for_backward(aw, mask, atten, sw, stack)
        ctx.cfg = (kb, thr_aw, thr_mask, prune, use_ref, row_len)
        theta, theta16 = like_phys(th32, aw), like_phys(th16, aw)
        ctx.mark_non_differentiable(theta16)
        return theta, theta16

    @staticmethod
    def backward(ctx, g, _g16):
        aw, mask, atten, sw, stack = ctx.saved_tensors
        kb, thr_aw, thr_mask, prune, use_ref, row_len = ctx.cfg
        if g is None:
            return (None,) * 10
        gf = flat_like(g.float(), aw)
        aw_f = phys_flat(aw.detach())
        st = stack_phys(stack, aw) if kb > 0 else None
        sw_f = phys_flat(sw) if sw is not None else None
        mk = mask.detach().contiguous() if mask is not None else None
        fn = wcompose_bwd_ref if use_ref else wcompose_bwd
        d_aw, d_att, d_mask = fn(gf, aw_f, st, kb, sw_f, mk, row_len, thr_aw, thr_mask, prune)
        g_aw = like_phys(gf if d_aw is None else d_aw, aw) if ctx.needs_input_grad[0] else None
        g_mask = d_mask.view(mask.shape) if (mask is not None and ctx.needs_input_grad[1] and d_mask is not None) else None
        g_att = None
        if atten is not None and ctx.needs_input_grad[2]:
            g_att = torch.zeros_like(atten)
            if kb > 0:
                g_att[:kb] = d_att
        return g_aw, g_mask, g_att, None, None, None, None, None, None, None


def compose_weight(aw: torch.Tensor, stack: Optional[torch.Tensor], atten: Optional[torch.Tensor], kb: int,
                   sw: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, thr_aw: float = 0.0,
                   thr_mask: float = 0.0, prune: bool = False, use_ref: bool = False):
    """Differentiable fused composition; returns ``(theta_fp32, theta_bf16)`` (see :class:`_WComposeFn`)."""
    return _WComposeFn.apply(aw, mask, atten, sw, stack, int(kb), float(thr_aw), float(thr_mask), bool(prune),
                             bool(use_ref or not _native(aw)))


# ===================================================================================================== Swin token ops
def window_src_rows(rows: int, H: int, W: int, ws: int, shift: int, device) -> torch.Tensor:
    """Image-layout row read by each window-layout row: window token ``(b, wh, ww, ph, pw)`` of the cyclically shifted
    map is image token ``(b, (wh*ws+ph+shift) % H, (ww*ws+pw+shift) % W)`` (``roll(-shift)`` then window partition)."""
    r = torch.arange(rows, device=device)
    ws2, nww, nwh = ws * ws, W // ws, H // ws
    win, pos = r // ws2, r % ws2
    ph, pw = pos // ws, pos % ws
    wwi, t = win % nww, win // nww
    whi, b = t % nwh, t // nwh
    hh, ww = (whi * ws + ph + shift) % H, (wwi * ws + pw + shift) % W
    return (b * H + hh) * W + ww


def image_src_rows(rows: int, H: int, W: int, ws: int, shift: int, device) -> torch.Tensor:
    """Window-layout row that lands on each image-layout row (window reverse then ``roll(+shift)``)."""
    r = torch.arange(rows, device=device)
    ws2, nww, nwh = ws * ws, W // ws, H // ws
    b, rem = r // (H * W), r % (H * W)
    hh, ww = rem // W, rem % W
    h2, w2 = (hh - shift + H) % H, (ww - shift + W) % W
    win = (b * nwh + h2 // ws) * nww + w2 // ws
    return win * ws2 + (h2 % ws) * ws + (w2 % ws)


def ln_rows_ref(x, gamma, beta, eps, window=None):
    rows, c = x.shape
    src = x if window is None else x[window_src_rows(rows, *window, device=x.device)]
    return F.layer_norm(src.float(), (c,), gamma.float(), beta.float(), eps).to(x.dtype)


def ln_rows(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
            window: Optional[Tuple[int, int, int, int]] = None) -> torch.Tensor:
    """LayerNorm over the last dim of ``x`` ``[rows, C]`` (bf16, fp32 statistics, fp32 ``gamma`` / ``beta``).
    ``window = (H, W, ws, shift)``: ``x`` is in image layout ``[B*H*W, C]`` and the result is written in the layout of
    the shifted windows ``[B*nW*ws*ws, C]`` (norm -> roll -> window partition in one pass)."""
    if not _native(x):
        return ln_rows_ref(x, gamma, beta, eps, window)
    lib = _lib()
    rows, c = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and c % 8 == 0 and c <= 2048
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.is_contiguous() and beta.is_contiguous()
    out = torch.empty_like(x)
    h, w, ws, sh = window if window is not None else (0, 0, 0, 0)
    rc = lib.flpr_ln_rows(native.ptr(x), native.ptr(gamma), native.ptr(beta), native.ptr(out), rows, c, float(eps),
                          int(window is not None), int(h), int(w), int(ws), int(sh), _stream(x.device))
    native.check(rc, "flpr_ln_rows")
    native.count_launch()
    return out


# ---- trainable LayerNorm (Swin blocks of the trainable stage): forward keeps (mean, rstd), backward is one sweep ----------
def ln_rows_train(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                  window: Optional[Tuple[int, int, int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """:func:`ln_rows` that also returns the per-destination-row ``[rows, 2]`` fp32 ``(mean, rstd)`` pairs."""
    rows, c = x.shape
    if not _native(x):
        src = x if window is None else x[window_src_rows(rows, *window, device=x.device)]
        xf = src.float()
        mean = xf.mean(1)
        rstd = torch.rsqrt(xf.var(1, unbiased=False) + eps)
        y = ((xf - mean[:, None]) * rstd[:, None]) * gamma.float() + beta.float()
        return y.to(x.dtype), torch.stack([mean, rstd], 1)
    lib = _lib()
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and c % 8 == 0 and c <= 2048
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.is_contiguous() and beta.is_contiguous()
    assert gamma.data_ptr() % 16 == 0 and beta.data_ptr() % 16 == 0
    out = torch.empty_like(x)
    stats = torch.empty(rows, 2, dtype=torch.float32, device=x.device)
    h, w, ws, sh = window if window is not None else (0, 0, 0, 0)
    rc = lib.flpr_ln_rows_train(native.ptr(x), native.ptr(gamma), native.ptr(beta), native.ptr(out), native.ptr(stats),
                                rows, c, float(eps), int(window is not None), int(h), int(w), int(ws), int(sh),
                                _stream(x.device))
    native.check(rc, "flpr_ln_rows_train")
    native.count_launch()
    return out, stats


def ln_rows_bwd_ref(dy, x, gamma, stats, window=None):
    rows, c = x.shape
    src_rows = None if window is None else window_src_rows(rows, *window, device=x.device)
    xf = (x if src_rows is None else x[src_rows]).float()
    xh = (xf - stats[:, :1]) * stats[:, 1:2]
    dyf = dy.float()
    g = dyf * gamma.float()
    dxw = stats[:, 1:2] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    if src_rows is None:
        dx = dxw
    else:
        dx = torch.empty_like(dxw)
        dx[src_rows] = dxw
    return dx.to(x.dtype), (dyf * xh).sum(0), dyf.sum(0)


def ln_rows_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, stats: torch.Tensor,
                window: Optional[Tuple[int, int, int, int]] = None):
    """Backward of :func:`ln_rows_train`: ``(dx [layout of x], dgamma, dbeta)``; ``dy`` is in the forward's destination
    layout. Per-block partials of dgamma / dbeta are folded in a fixed order (deterministic)."""
    if not _native(dy):
        return ln_rows_bwd_ref(dy, x, gamma, stats, window)
    lib = _lib()
    rows, c = x.shape
    assert dy.shape == x.shape and dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    assert dy.is_contiguous() and x.is_contiguous() and stats.is_contiguous() and stats.dtype == torch.float32
    assert gamma.dtype == torch.float32 and gamma.is_contiguous() and gamma.data_ptr() % 16 == 0
    dx = torch.empty_like(x)
    blocks = int(lib.flpr_ln_rows_bwd_blocks(rows))
    part = torch.empty(blocks, 2, c, dtype=torch.float32, device=x.device)
    dgb = torch.empty(2, c, dtype=torch.float32, device=x.device)
    h, w, ws, sh = window if window is not None else (0, 0, 0, 0)
    rc = lib.flpr_ln_rows_bwd(native.ptr(dy), native.ptr(x), native.ptr(gamma), native.ptr(stats), native.ptr(dx),
                              native.ptr(part), native.ptr(dgb), rows, c, int(window is not None), int(h), int(w),
                              int(ws), int(sh), _stream(x.device))
    native.check(rc, "flpr_ln_rows_bwd")
    native.count_launch(2)
    return dx, dgb[0], dgb[1]


class _LnRowsFn(torch.autograd.Function):
    """LayerNorm over bf16 token rows with fp32 statistics and fp32 affine parameters (``models/swin_transformer.py:
    358-395``, ``norm1`` / ``norm2`` of a trainable block): bf16 in, bf16 out - the consumer is a tensor-core Linear that
    would cast the fp32 result of the autocast ``layer_norm`` to bf16 anyway, so the values reaching the GEMM are the same
    and the fp32 round trip of the token stream (forward and backward) disappears."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, window):
        y, stats = ln_rows_train(x, gamma.detach(), beta.detach(), eps, window)
        ctx.save_for_backward(x, gamma, stats)
        ctx.window = window
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, stats = ctx.saved_tensors
        dyc = dy if (dy.dtype == x.dtype and dy.is_contiguous()) else dy.to(x.dtype).contiguous()
        dx, dg, db = ln_rows_bwd(dyc, x, gamma.detach(), stats, ctx.window)
        return (dx if ctx.needs_input_grad[0] else None, dg.to(gamma.dtype) if ctx.needs_input_grad[1] else None,
                db.to(gamma.dtype) if ctx.needs_input_grad[2] else None, None, None)


def layer_norm_rows(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                    window: Optional[Tuple[int, int, int, int]] = None) -> torch.Tensor:
    """Differentiable :func:`ln_rows` (``x``: ``[rows, C]`` bf16 contiguous; fp32 ``gamma`` / ``beta``)."""
    return _LnRowsFn.apply(x, gamma, beta, float(eps), window)


def window_merge_add_ref(win, shortcut, H, W, ws, shift):
    rows = shortcut.shape[0]
    return (shortcut.float() + win[image_src_rows(rows, H, W, ws, shift, win.device)].float()).to(shortcut.dtype)


def window_merge_add(win: torch.Tensor, shortcut: torch.Tensor, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """``shortcut + roll(window_reverse(win), +shift)`` over ``[B*H*W, C]`` bf16 rows in one pass."""
    if not _native(win):
        return window_merge_add_ref(win, shortcut, H, W, ws, shift)
    lib = _lib()
    rows, c = shortcut.shape
    assert win.shape == shortcut.shape and win.dtype == torch.bfloat16 and shortcut.dtype == torch.bfloat16
    assert win.is_contiguous() and shortcut.is_contiguous() and c % 8 == 0
    out = torch.empty_like(shortcut)
    rc = lib.flpr_window_merge_add(native.ptr(win), native.ptr(shortcut), native.ptr(out), rows, c, int(H), int(W),
                                   int(ws), int(shift), _stream(win.device))
    native.check(rc, "flpr_window_merge_add")
    native.count_launch()
    return out


# ---- trainable block: merge + residual with the per-sample drop-path factor, and its backward ---------------------------------
def window_merge_add_scaled_ref(win, shortcut, scale, H, W, ws, shift):
    rows = shortcut.shape[0]
    g = win[image_src_rows(rows, H, W, ws, shift, win.device)].float()
    if scale is not None:
        g = g * scale.float().repeat_interleave(H * W)[:, None]
    return (shortcut.float() + g).to(shortcut.dtype)


def window_merge_add_scaled(win: torch.Tensor, shortcut: torch.Tensor, scale: Optional[torch.Tensor], H: int, W: int,
                            ws: int, shift: int) -> torch.Tensor:
    """``shortcut + scale[sample] * roll(window_reverse(win), +shift)`` over ``[B*H*W, C]`` bf16 rows (``scale``: ``[B]``
    fp32 drop-path factors 0 or ``1 / keep``; ``None`` = 1)."""
    if not _native(win):
        return window_merge_add_scaled_ref(win, shortcut, scale, H, W, ws, shift)
    lib = _lib()
    rows, c = shortcut.shape
    assert win.shape == shortcut.shape and win.dtype == torch.bfloat16 and shortcut.dtype == torch.bfloat16
    assert win.is_contiguous() and shortcut.is_contiguous() and c % 8 == 0
    assert scale is None or (scale.dtype == torch.float32 and scale.is_contiguous() and scale.numel() == rows // (H * W))
    out = torch.empty_like(shortcut)
    rc = lib.flpr_window_merge_add_scaled(native.ptr(win), native.ptr(shortcut), native.ptr(scale), native.ptr(out), rows,
                                          c, int(H), int(W), int(ws), int(shift), _stream(win.device))
    native.check(rc, "flpr_window_merge_add_scaled")
    native.count_launch()
    return out


def window_gather_scale_ref(dy, scale, H, W, ws, shift):
    rows = dy.shape[0]
    src = window_src_rows(rows, H, W, ws, shift, dy.device)
    g = dy[src].float()
    if scale is not None:
        g = g * scale.float().repeat_interleave(H * W)[src][:, None]
    return g.to(dy.dtype)


def window_gather_scale(dy: torch.Tensor, scale: Optional[torch.Tensor], H: int, W: int, ws: int, shift: int
                        ) -> torch.Tensor:
    """Backward of :func:`window_merge_add_scaled` w.r.t. ``win``: ``scale[sample] * dy`` gathered into window layout."""
    if not _native(dy):
        return window_gather_scale_ref(dy, scale, H, W, ws, shift)
    lib = _lib()
    rows, c = dy.shape
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and c % 8 == 0
    out = torch.empty_like(dy)
    rc = lib.flpr_window_gather_scale(native.ptr(dy), native.ptr(scale), native.ptr(out), rows, c, int(H), int(W),
                                      int(ws), int(shift), _stream(dy.device))
    native.check(rc, "flpr_window_gather_scale")
    native.count_launch()
    return out


class _WindowMergeAddFn(torch.autograd.Function):
    """``out = shortcut + scale[sample] * merge(win)`` (``models/swin_transformer.py:383-391`` incl. the block's
    ``drop_path``): window reverse, roll back, stochastic-depth scaling and the residual add in one pass; backward: the
    shortcut gradient is ``dy`` itself, the window gradient one gather pass."""

    @staticmethod
    def forward(ctx, win, shortcut, scale, H, W, ws, shift):
        ctx.cfg = (H, W, ws, shift)
        ctx.save_for_backward(scale)
        return window_merge_add_scaled(win, shortcut, scale, H, W, ws, shift)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        dyc = dy if dy.is_contiguous() else dy.contiguous()
        dwin = window_gather_scale(dyc, scale, *ctx.cfg) if ctx.needs_input_grad[0] else None
        return dwin, (dyc if ctx.needs_input_grad[1] else None), None, None, None, None, None


def window_merge_residual(win: torch.Tensor, shortcut: torch.Tensor, scale: Optional[torch.Tensor], H: int, W: int,
                          ws: int, shift: int) -> torch.Tensor:
    """Differentiable :func:`window_merge_add_scaled`."""
    return _WindowMergeAddFn.apply(win, shortcut, scale, int(H), int(W), int(ws), int(shift))


def gelu_rows(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU over a bf16 tensor."""
    if not _native(x):
        return F.gelu(x.float()).to(x.dtype)
    lib = _lib()
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty_like(x)
    native.check(lib.flpr_gelu_rows(native.ptr(x), native.ptr(out), x.numel(), _stream(x.device)),
                 "flpr_gelu_rows")
    native.count_launch()
    return out


def gelu_bwd_rows_ref(x, dy):
    xf = x.float()
    cdf = 0.5 * (1 + torch.erf(xf * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * xf * xf)
    return (dy.float() * (cdf + xf * pdf)).to(x.dtype)


def gelu_bwd_rows(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """``dy * GELU'(x)`` (exact form) over bf16 tensors; ``x`` is the saved pre-activation."""
    if not _native(x):
        return gelu_bwd_rows_ref(x, dy)
    lib = _lib()
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    assert x.shape == dy.shape and x.numel() % 8 == 0
    dx = torch.empty_like(x)
    native.check(lib.flpr_gelu_bwd_rows(native.ptr(x), native.ptr(dy), native.ptr(dx), x.numel(), _stream(x.device)),
                 "flpr_gelu_bwd_rows")
    native.count_launch()
    return dx


class _GeluFn(torch.autograd.Function):
    """Exact GELU of the Swin MLP (``models/swin_transformer.py:118-140``) in the trainable stage: one pass forward, one
    pass backward from the saved pre-activation."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return gelu_rows(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return gelu_bwd_rows(x, dy if (dy.dtype == x.dtype and dy.is_contiguous()) else dy.to(x.dtype).contiguous())


def gelu_act(x: torch.Tensor) -> torch.Tensor:
    """Differentiable :func:`gelu_rows` (``x``: bf16, contiguous, numel % 8 == 0)."""
    return _GeluFn.apply(x)


# ===================================================================================================== dispatch apply
def apply_global_ref(flat, master, shadow, p_old, snap_mode):
    n = flat.numel()
    if p_old is not None and snap_mode == 1:
        p_old[:n].copy_(master[:n])
    master[:n].copy_(flat)
    if p_old is not None and snap_mode == 2:
        p_old[:n].copy_(flat)
    if shadow is not None:
        shadow[:n].copy_(flat)


def apply_global(flat: torch.Tensor, master: torch.Tensor, shadow: Optional[torch.Tensor] = None,
                 p_old: Optional[torch.Tensor] = None, snap_mode: int = 0) -> None:
    """Receiving end of a FedAvg-family dispatch: ``master[:n] <- flat`` with the bf16 compute copy refreshed and the
    FedProx anchor snapshotted in the same pass (``snap_mode`` 1: the weights being replaced, 2: the incoming ones)."""
    n = flat.numel()
    ok = _native(flat) and n % 4 == 0 and flat.dtype == torch.float32 and flat.is_contiguous() and \
        all(t is None or (t.is_contiguous() and t.data_ptr() % 16 == 0) for t in (flat, master, shadow, p_old))
    if not ok:
        return apply_global_ref(flat, master, shadow, p_old, snap_mode)
    lib = _lib()
    assert master.dtype == torch.float32 and master.numel() >= n
    assert shadow is None or (shadow.dtype == torch.bfloat16 and shadow.numel() >= n)
    assert p_old is None or (p_old.dtype == torch.float32 and p_old.numel() >= n)
    rc = lib.flpr_apply_global(native.ptr(flat), native.ptr(master), native.ptr(shadow),
                               native.ptr(p_old if snap_mode else None), int(snap_mode), n, _stream(flat.device))
    native.check(rc, "flpr_apply_global")
    native.count_launch()


# ===================================================================================================== self-checks
_state: Dict[str, bool] = {}
_lock = threading.Lock()


def _close(a: torch.Tensor, b: torch.Tensor, rtol: float, atol_frac: float) -> bool:
    a, b = a.float(), b.float()
    if a.shape != b.shape or not bool(torch.isfinite(a).all()):
        return False
    atol = atol_frac * float(b.abs().max()) + 1e-6
    return bool(torch.allclose(a, b, rtol=rtol, atol=atol))


def _check_wcompose(dev) -> bool:
    g = torch.Generator(device="cpu").manual_seed(11)
    ok = True
    for (rows, row_len, kb, ks, with_sw, prune) in ((24, 200, 5, 5, True, True), (7, 8192, 8, 8, False, False),
                                                    (5, 333, 0, 0, True, True), (3, 1000, 3, 6, True, False)):
        n = rows * row_len - (3 if not with_sw else 0)
        aw = (torch.randn(n, generator=g) * 0.01).to(dev)
        stack = torch.randn(n, max(ks, 1), generator=g).to(dev) if kb else None
        atten = torch.randn(max(ks, 1), generator=g).to(dev) if kb else None
        sw = torch.randn(n, generator=g).to(dev) if with_sw else None
        mask = torch.rand(rows, generator=g).to(dev) if with_sw else None
        thr = 0.008
        r32 = wcompose_fwd_ref(aw, stack, atten, kb, sw, mask, row_len, thr, 0.5, prune)
        o32, o16 = wcompose_fwd(aw, stack, atten, kb, sw, mask, row_len, thr, 0.5, prune)
        ok = ok and _close(o32, r32, 1e-5, 1e-6) and _close(o16, r32, 1e-2, 1e-2)
        dth = torch.randn(n, generator=g).to(dev)
        ra, rt, rm = wcompose_bwd_ref(dth, aw, stack, kb, sw, mask, row_len if with_sw else 8192, thr, 0.5, prune)
        da, dt, dm = wcompose_bwd(dth, aw, stack, kb, sw, mask, row_len, thr, 0.5, prune)
        ok = ok and ((da is None) == (ra is None)) and (da is None or bool(torch.equal(da, ra)))
        ok = ok and (kb == 0 or _close(dt, rt, 1e-4, 1e-5)) and (not with_sw or _close(dm, rm, 1e-4, 1e-5))
    return ok


def _check_swin_tokens(dev) -> bool:
    from . import gemm as gops
    g = torch.Generator(device="cpu").manual_seed(12)
    ok = True
    for (b, h, w, ws, shift, c) in ((2, 14, 14, 7, 3, 96), (1, 8, 4, 4, 0, 192), (3, 7, 7, 7, 0, 768),
                                    (1, 14, 7, 7, 2, 1536)):
        rows = b * h * w
        x = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        gamma, beta = (1 + 0.1 * torch.randn(c, generator=g)).to(dev), (0.1 * torch.randn(c, generator=g)).to(dev)
        for window in (None, (h, w, ws, shift)):
            ok = ok and _close(ln_rows(x, gamma, beta, 1e-5, window), ln_rows_ref(x, gamma, beta, 1e-5, window),
                               2e-2, 1e-2)
        win = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        ok = ok and bool(torch.equal(window_merge_add(win, x, h, w, ws, shift),
                                     window_merge_add_ref(win, x, h, w, ws, shift)))
        ok = ok and _close(gelu_rows(x), F.gelu(x.float()), 1e-2, 1e-2)
    # fc2 of the Swin MLP with bias and the residual stream folded into the GEMM epilogue (N = 96: three 32-column chunks)
    for (m, n, k) in ((392, 96, 384), (98, 768, 3072)):
        a = torch.randn(m, k, generator=g).to(dev).to(torch.bfloat16)
        wt = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev).to(torch.bfloat16)
        bias = torch.randn(n, generator=g).to(dev)
        res = torch.randn(m, n, generator=g).to(dev).to(torch.bfloat16)
        y = gops.gemm(a, wt, bias_n=bias, residual=res)
        ref = a.float() @ wt.float().t() + bias + res.float()
        ok = ok and _close(y, ref, 2e-2, 1e-2)
    return ok


def _check_ln_train(dev) -> bool:
    g = torch.Generator(device="cpu").manual_seed(14)
    ok = True
    for (b, h, w, ws, shift, c) in ((2, 8, 4, 4, 0, 768), (3, 14, 14, 7, 3, 96), (40, 7, 7, 7, 0, 1024),
                                    (5, 8, 4, 4, 2, 1536), (1, 4, 4, 4, 0, 384)):
        rows = b * h * w
        x = (torch.randn(rows, c, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16)
        gamma, beta = (1 + 0.1 * torch.randn(c, generator=g)).to(dev), (0.1 * torch.randn(c, generator=g)).to(dev)
        dy = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        for window in (None, (h, w, ws, shift)):
            y, stats = ln_rows_train(x, gamma, beta, 1e-5, window)
            ok = ok and _close(y, ln_rows_ref(x, gamma, beta, 1e-5, window), 2e-2, 1e-2)
            # yardstick: fp32 autograd through F.layer_norm on the (gathered) rows
            xr = x.float().requires_grad_(True)
            gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
            with torch.enable_grad():
                src = xr if window is None else xr[window_src_rows(rows, *window, device=dev)]
                F.layer_norm(src, (c,), gr, br, 1e-5).backward(dy.float())
            dx, dg, db = ln_rows_bwd(dy, x, gamma, stats, window)
            ok = ok and _close(dx, xr.grad, 3e-2, 1e-2) and _close(dg, gr.grad, 1e-2, 5e-3) and \
                _close(db, br.grad, 1e-3, 1e-4)
        ok = ok and _close(gelu_bwd_rows(x, dy), gelu_bwd_rows_ref(x, dy), 2e-2, 1e-2)
        # merge + residual with the per-sample drop-path factor, and the gather that is its backward
        win = torch.randn(rows, c, generator=g).to(dev).to(torch.bfloat16)
        for scale in (None, (torch.rand(b, generator=g) < 0.6).float().div(0.6).to(dev)):
            ok = ok and _close(window_merge_add_scaled(win, x, scale, h, w, ws, shift),
                               window_merge_add_scaled_ref(win, x, scale, h, w, ws, shift), 1e-2, 1e-2)
            ok = ok and _close(window_gather_scale(dy, scale, h, w, ws, shift),
                               window_gather_scale_ref(dy, scale, h, w, ws, shift), 1e-2, 1e-2)
    return ok


def _check_apply(dev) -> bool:
    g = torch.Generator(device="cpu").manual_seed(13)
    ok = True
    for n, total, mode, with_shadow in ((4096, 5000, 1, True), (1 << 20, 1 << 20, 2, True), (64, 64, 0, False),
                                        (12, 16, 1, False)):
        flat = torch.randn(n, generator=g).to(dev)
        m0 = torch.randn(total, generator=g).to(dev)
        res = []
        for fn in (apply_global_ref, apply_global):
            master, p_old = m0.clone(), torch.zeros(total, device=dev)
            shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev) if with_shadow else None
            fn(flat, master, shadow, p_old if mode else None, mode)
            res.append((master, p_old, shadow))
        for a, b in zip(*res):
            ok = ok and (a is None) == (b is None) and (a is None or bool(torch.equal(a, b)))
    return ok


_CHECKS = {"wcompose": _check_wcompose, "swin_tokens": _check_swin_tokens, "apply": _check_apply,
           "ln_train": _check_ln_train}


def run_checks_inprocess(device, families=None) -> Dict[str, bool]:
    """Run the numerics checks of ``families`` (default: all) on ``device`` in THIS process (what the isolated children
    of :func:`enabled` execute; also ``FLPR_LAYER_SELFCHECK=inprocess``)."""
    dev = torch.device(device)
    out: Dict[str, bool] = {}
    for family in (families or list(_CHECKS)):
        fn = _CHECKS[family]
        try:
            with torch.no_grad(), torch.autocast(device_type=dev.type, enabled=False):
                ok = bool(fn(dev))
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
        except Exception as ex:  # noqa: BLE001  (a failed launch must not take the experiment down)
            _log.error("flpr layer kernels '%s': self-check raised %s: %s", family, type(ex).__name__, ex)
            ok = False
        out[family] = ok
    return out


def _cache_path(dev: torch.device) -> str:
    import hashlib
    import tempfile
    from .. import _build
    try:
        with open(_build.STAMP) as f:
            stamp = f.read().strip()[:16]
    except OSError:
        stamp = "nostamp"
    name = torch.cuda.get_device_name(dev) if dev.type == "cuda" else "cpu"      # (cpu: plumbing tests only)
    gpu = hashlib.sha1(name.encode()).hexdigest()[:8]
    return os.path.join(tempfile.gettempdir(), f"flpr_layer_selfcheck_{stamp}_{gpu}.json")


def _checks_isolated(dev: torch.device) -> Dict[str, bool]:
    """The checks in child processes with their own CUDA contexts: a kernel fault there (illegal address = sticky
    context error) costs the child, not the experiment. The verdict is cached per (library build, GPU model) in the
    temp dir."""
    import json
    import subprocess
    import sys
    path = _cache_path(dev)
    try:
        with open(path) as f:
            return {k: bool(v) for k, v in json.load(f).items()}
    except (OSError, ValueError):
        pass
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, FLPR_LAYER_SELFCHECK="inprocess",
               PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    if dev.type == "cuda":
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        dev_s, pre = f"cuda:{idx}", f"torch.cuda.set_device({idx});"
    else:
        dev_s, pre = "cpu", ""
    verdict = {k: False for k in _CHECKS}
    procs = {}
    try:
        for fam in _CHECKS:                      # one child per family, side by side: a fault in one family's kernels
            code = ("import json,sys,torch;from flpr_b200.ops import layer as L;"      # (dead context) cannot fail another
                    f"{pre}r=L.run_checks_inprocess('{dev_s}',['{fam}']);"
                    "print('FLPR_SELFCHECK '+json.dumps(r))")
            procs[fam] = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                          text=True, env=env, cwd=root)
        for fam, pr in procs.items():
            try:
                so, se = pr.communicate(timeout=300)
            except subprocess.TimeoutExpired:
                pr.kill()
                so, se = pr.communicate()
            for line in so.splitlines():
                if line.startswith("FLPR_SELFCHECK "):
                    verdict[fam] = bool(json.loads(line[len("FLPR_SELFCHECK "):]).get(fam, False))
                    break
            else:
                _log.error("flpr layer kernels '%s': self-check child exited with code %s: %s", fam, pr.returncode,
                           (se or "")[-2000:])
    except Exception as ex:  # noqa: BLE001
        _log.error("flpr layer kernels: self-check children failed: %s: %s", type(ex).__name__, ex)
        for pr in procs.values():
            if pr.poll() is None:
                pr.kill()
    try:
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "w") as f:
            json.dump(verdict, f)
        os.replace(tmp, path)
    except OSError:
        pass
    return verdict


def enabled(family: str, device=None) -> bool:
    """True when the kernel ``family`` (``wcompose`` / ``swin_tokens``) passed its one-time on-device numerics check
    (run in an isolated child process, cached per library build + GPU model). ``FLPR_LAYER_OPS=0`` switches every
    family off, ``FLPR_LAYER_OPS=force`` on without a check; ``FLPR_LAYER_SELFCHECK=inprocess`` runs the checks in the
    calling process (not during a CUDA-graph capture: the check synchronises)."""
    hit = _state.get(family)
    if hit is not None:
        return hit
    mode = os.environ.get("FLPR_LAYER_OPS", "1")
    if mode == "0" or not torch.cuda.is_available():
        _state[family] = False
        return False
    if mode == "force":
        _state[family] = True
        return True
    inproc = os.environ.get("FLPR_LAYER_SELFCHECK", "isolated") == "inprocess"
    if inproc and torch.cuda.is_current_stream_capturing():
        return False
    with _lock:
        hit = _state.get(family)
        if hit is not None:
            return hit
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        verdict = run_checks_inprocess(dev) if inproc else _checks_isolated(dev)
        for fam in _CHECKS:
            ok = bool(verdict.get(fam, False))
            if not ok:
                _log.error("flpr layer kernels '%s' FAILED their on-device numerics self-check: the PyTorch "
                           "formulation is used instead (slower, same results)", fam)
            _state[fam] = ok
        return _state[family]


def status() -> Dict[str, bool]:
    return dict(_state)
