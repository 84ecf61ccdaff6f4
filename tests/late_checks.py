"""Device-agnostic bodies of the checks for the components added after the round's last GPU session (stride-2
convolution gradients on the stride-1 kernels, fused compose kernels of FedWeIT / fedstil-atten, Swin token kernels).
``test_cpu_late.py`` runs them on the CPU reference paths of the ops, ``test_zz_gpu_late.py`` on the native kernels
(that file sorts last on purpose: these kernels have not been on a GPU yet, a failure there must not hide the rest of
the suite behind ``-x``). Every check compares against a plain fp32 PyTorch formulation of the same op."""
import torch
import torch.nn.functional as F


def close(a, b, rtol=3e-2, atol=None):
    a, b = a.float(), b.float()
    if atol is None:
        atol = 3e-2 * b.abs().max().item() + 1e-6
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max().item()


# ------------------------------------------------------------------------------------------------ stride-2 convolutions
def check_conv_stride2(device, k, n, h, w, cin, cout, want_stats=False, slot=False):
    from flpr_b200.ops.gemm import conv_s2_supported, conv_stride2
    assert conv_s2_supported(h, w, cin, cout, k)
    g = torch.Generator().manual_seed(k * 1000 + h)
    x = torch.randn(n, h, w, cin, generator=g).to(device).bfloat16().requires_grad_(True)
    wt = (torch.randn(cout, k, k, cin, generator=g) / (k * cin ** 0.5)).to(device).requires_grad_(True)
    grad_slot = torch.zeros(cout, k, k, cin, device=device) if slot else None
    out = conv_stride2(x, wt, None, grad_slot, want_stats)
    y = out[0] if want_stats else out
    gy = torch.randn(y.shape, generator=g).to(device).to(y.dtype)
    y.backward(gy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().bfloat16().float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=k // 2)
    yr.backward(gy.float().permute(0, 3, 1, 2))
    close(y, yr.permute(0, 2, 3, 1))
    close(x.grad, xr.grad.permute(0, 2, 3, 1))
    gw = grad_slot if (slot and device != "cpu") else wt.grad
    close(gw, wr.grad.permute(0, 2, 3, 1), rtol=2e-2, atol=0.05 * wr.grad.abs().max().item())
    if want_stats:
        part = out[1]
        yf = y.detach().float().reshape(-1, cout)
        close(part[:, 0].sum(0), yf.sum(0), atol=0.02 * yf.abs().sum(0).max().item() + 1e-3)
        close(part[:, 1].sum(0), (yf * yf).sum(0), atol=0.02 * (yf * yf).sum(0).max().item() + 1e-3)


def check_fast_head_last_stride2(device):
    """ResNet head with ``last_stride: 2`` (the strided 3x3 and the strided 1x1 downsample of layer4.0 need
    gradients): fast head vs the fp32 module."""
    import copy
    from flpr_b200.models import nets
    from flpr_b200.models.resnet import FastResNetHead
    torch.manual_seed(0)
    net = nets["resnet50"](num_classes=16, last_stride=2, neck="bnneck")
    net.configure_split(["base.layer4", "classifier"])
    net = net.to(device).to(memory_format=torch.channels_last)
    net.train()
    ref = copy.deepcopy(net).float()
    fast = FastResNetHead(net)
    proto = (torch.randn(8, 1024, 16, 8) * 0.5).to(device)
    tgt = torch.randint(0, 16, (8,)).to(device)
    s, f = fast(proto.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    loss = F.cross_entropy(s.float(), tgt)
    loss.backward()
    x = ref.base.run_stages(proto, ref.head_start, 5)
    rf = x.mean(dim=(2, 3))
    rs = ref.classifier(ref.bottleneck(rf))
    rl = F.cross_entropy(rs, tgt)
    rl.backward()
    close(f, rf, rtol=5e-2, atol=5e-2 * rf.abs().max().item())
    assert abs(loss.item() - rl.item()) < 5e-2 * abs(rl.item())
    for name in ("base.layer4.0.conv2.weight", "base.layer4.0.downsample.0.weight", "base.layer4.0.conv1.weight"):
        a = net.get_parameter(name).grad
        b = ref.get_parameter(name).grad
        assert a is not None, name
        cos = F.cosine_similarity(a.float().flatten(), b.flatten(), dim=0).item()
        assert cos > 0.97, (name, cos)


# ------------------------------------------------------------------------------------------------ fused weight composition
def check_compose_function(device, channels_last, use_ref):
    """``ops.layer.compose_weight`` (kernel or its flat reference) vs autograd through the original expression, on
    plain and on channels_last (OHWI, the CUDA arena layout) storage."""
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(3)
    for shape, kb in (((16, 8, 3, 3), 5), ((24, 40), 3), ((8, 16, 1, 1), 2)):
        def mk(*s, scale=1.0):
            return (torch.randn(*s, generator=g) * scale).to(device)
        aw, sw, stack = mk(*shape, scale=0.01), mk(*shape), mk(*shape, kb)
        mask, atten = torch.rand(shape[0], generator=g).to(device), mk(kb)
        if channels_last and len(shape) == 4:
            aw = aw.contiguous(memory_format=torch.channels_last)
            sw = sw.contiguous(memory_format=torch.channels_last)
        stack = lops.stack_aligned(stack, aw)
        assert lops.stack_phys(stack, aw) is not None
        thr_aw, thr_mask = 0.008, 0.4
        outs = []
        for fused in (False, True):
            a, m, t = (x.detach().clone().requires_grad_(True) for x in (aw, mask, atten))
            if channels_last and len(shape) == 4:
                a = aw.detach().clone(memory_format=torch.channels_last).requires_grad_(True)
            if fused:
                th, th16 = lops.compose_weight(a, stack, t, kb, sw, m, thr_aw, thr_mask, True, use_ref)
                assert th16.dtype == torch.bfloat16 and th16.shape == th.shape and th16.stride() == th.stride()
                close(th16, th, rtol=1e-2, atol=1e-2 * th.abs().max().item())
            else:
                pa = a * (a.abs() > thr_aw).to(a.dtype)
                pm = m * (m.abs() > thr_mask).to(m.dtype)
                th = pm.view(-1, *([1] * (len(shape) - 1))) * sw + pa + (stack * t).sum(-1)
            gy = torch.randn(shape, generator=torch.Generator().manual_seed(9)).to(device)
            (th * gy).sum().backward()
            outs.append((th.detach(), a.grad, m.grad, t.grad))
        for x, y in zip(*outs):
            close(x, y, rtol=1e-4, atol=1e-5 * y.abs().max().item() + 1e-6)
    # stacked form without a shared weight (fedstil-atten): theta = aw + stack @ atten, d aw is the incoming gradient
    aw, stack, atten = mk(10, 6, 3, 3), mk(10, 6, 3, 3, 4), mk(4)
    if channels_last:
        aw = aw.contiguous(memory_format=torch.channels_last)
    stack = lops.stack_aligned(stack, aw)
    a, t = aw.clone().requires_grad_(True), atten.clone().requires_grad_(True)
    th, _ = lops.compose_weight(a, stack, t, 4, use_ref=use_ref)
    th.sum().backward()
    close(th, aw + (stack * atten).sum(-1), rtol=1e-5, atol=1e-5)
    close(a.grad, torch.ones_like(aw), rtol=0, atol=0)
    close(t.grad, stack.reshape(-1, 4).sum(0), rtol=1e-4, atol=1e-4)


def check_fedweit_layer_fused(device, channels_last, use_ref):
    """``Decomposed._theta_fused`` vs the original ``theta`` expression of the same layer (values + gradients), incl.
    the storage re-alignment of ``sw`` / ``aw_kb`` and logical-shape writes into them afterwards."""
    import torch.nn as nn
    from flpr_b200.methods.fedweit import Decomposed
    torch.manual_seed(5)
    for src in (nn.Conv2d(16, 32, 3, padding=1, bias=False), nn.Linear(48, 24, bias=False)):
        layer = Decomposed(src, kb_cnt=3, lambda_l1=0.02, lambda_mask=0.1).to(device)
        with torch.no_grad():
            layer.aw_kb.copy_(torch.randn_like(layer.aw_kb) * 0.05)
            layer.atten.copy_(torch.randn(3) * 0.3)
            layer.mask.copy_(torch.rand_like(layer.mask))
        if channels_last and layer.is_conv:
            layer.aw.data = layer.aw.data.contiguous(memory_format=torch.channels_last)
        layer.align_storage()
        layer.train()
        ref = layer.theta(True) if device == "cpu" else None
        if ref is None:                                   # on CUDA theta() is the fused path: rebuild the expression
            pa = layer.aw * (layer.aw.abs() > layer.lambda_l1)
            pm = layer.mask * (layer.mask.abs() > layer.lambda_mask)
            ref = layer._bmask(pm, layer.sw) * layer.sw + pa + (layer.aw_kb * layer.atten).sum(-1)
        gy = torch.randn(ref.shape).to(device)
        (ref * gy).sum().backward()
        want = [p.grad.clone() for p in (layer.aw, layer.mask, layer.atten)]
        for p in (layer.aw, layer.mask, layer.atten):
            p.grad = None
        th = layer._theta_fused(True, use_ref)
        assert th is not None and th._flpr_bf16.dtype == torch.bfloat16
        (th * gy).sum().backward()
        close(th, ref, rtol=1e-5, atol=1e-6)
        for p, w in zip((layer.aw, layer.mask, layer.atten), want):
            close(p.grad, w, rtol=1e-4, atol=1e-5 * w.abs().max().item() + 1e-7)
        # logical-shape writes (dispatch) keep working on the re-aligned buffers
        new_kb = torch.randn(*layer.aw.shape, 3).to(device)
        with torch.no_grad():
            layer.aw_kb.copy_(new_kb)
            layer.reinit()
        th2 = layer._theta_fused(False, use_ref)
        close(th2, layer.mask.view(-1, *([1] * (layer.sw.dim() - 1))) * layer.sw + layer.aw + (new_kb * layer.atten).sum(-1),
              rtol=1e-5, atol=1e-6)


def check_atten_composer_storage(device, use_ref):
    """fedstil-atten ``_Compose`` with the weight in OHWI storage (what the CUDA arena does): ``gw`` re-aligned, logical
    accessors unchanged, arena-order stack chunks land correctly, fused composition == the matmul formulation."""
    from flpr_b200.methods.fedstil_atten import _Compose
    from flpr_b200.ops import layer as lops
    torch.manual_seed(7)
    w = torch.randn(12, 8, 3, 3).to(device)
    comp = _Compose(w, 4, 0.8).to(device)
    aw = comp.right_inverse(w).contiguous(memory_format=torch.channels_last)          # (1 - a) * w, OHWI storage
    before = comp._mix(comp.atten.detach(), aw).clone()
    comp.align_storage(aw)
    assert lops.stack_phys(comp.gw, aw) is not None and comp.gw.shape == (12, 8, 3, 3, 4)
    close(comp._mix(comp.atten.detach(), aw), before, rtol=0, atol=0)
    close(comp.gw[..., 0], w, rtol=0, atol=0)
    # a dispatched stack arrives in arena (physical) order: [numel, Kmax] rows ordered (o, h, w, i)
    logical = torch.randn(12, 8, 3, 3, 4).to(device)
    chunk = logical.permute(0, 2, 3, 1, 4).reshape(-1, 4).contiguous()
    comp.gw.copy_(chunk.view(12, 3, 3, 8, 4).permute(0, 3, 1, 2, 4))
    close(comp.gw, logical, rtol=0, atol=0)
    assert lops.stack_phys(comp.gw, aw) is not None                                   # the copy kept the storage order
    close(lops.stack_phys(comp.gw, aw), chunk, rtol=0, atol=0)
    with torch.no_grad():
        comp.atten.copy_(torch.tensor([0.8, 0.8, 0.1, 0.0]))
    a = aw.clone(memory_format=torch.channels_last).requires_grad_(True)
    th, t16 = lops.compose_weight(a, comp.gw, comp.atten, comp.k_max, use_ref=use_ref)
    ref = (logical * comp.atten.detach()).sum(-1) + aw
    close(th, ref, rtol=1e-5, atol=1e-5)
    assert th.stride() == aw.stride() and t16.stride() == aw.stride()
    gy = torch.randn(12, 8, 3, 3).to(device)
    (th * gy).sum().backward()
    close(a.grad, gy, rtol=0, atol=0)
    close(comp.atten.grad, (logical * gy[..., None]).reshape(-1, 4).sum(0), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ Swin token kernels
def check_swin_token_ops(device):
    """``ln_rows`` (plain / shifted-window destination), ``window_merge_add`` and ``gelu_rows`` against the block's
    original formulation: ``LayerNorm -> roll -> window_partition`` and ``window_reverse -> roll -> + shortcut``."""
    from flpr_b200.models.swin import window_partition, window_reverse
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(21)
    for (b, h, w, ws, shift, c) in ((2, 14, 14, 7, 3, 96), (1, 8, 4, 4, 0, 192), (3, 7, 7, 7, 0, 768),
                                    (1, 14, 7, 7, 2, 1536), (2, 8, 8, 4, 1, 384)):
        x = torch.randn(b * h * w, c, generator=g).to(device).to(torch.bfloat16)
        gamma = (1 + 0.1 * torch.randn(c, generator=g)).to(device)
        beta = (0.1 * torch.randn(c, generator=g)).to(device)
        ln = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
        close(lops.ln_rows(x, gamma, beta, 1e-5), ln, rtol=2e-2, atol=2e-2)
        img = ln.view(b, h, w, c)
        if shift:
            img = torch.roll(img, shifts=(-shift, -shift), dims=(1, 2))
        want = window_partition(img, ws).reshape(-1, c)
        close(lops.ln_rows(x, gamma, beta, 1e-5, (h, w, ws, shift)), want, rtol=2e-2, atol=2e-2)
        win = torch.randn(b * h * w, c, generator=g).to(device).to(torch.bfloat16)
        back = window_reverse(win.view(-1, ws * ws, c), ws, h, w)
        if shift:
            back = torch.roll(back, shifts=(shift, shift), dims=(1, 2))
        want = (x.float() + back.reshape(-1, c).float()).to(torch.bfloat16)
        got = lops.window_merge_add(win, x, h, w, ws, shift)
        assert torch.equal(got, want)
        close(lops.gelu_rows(x), F.gelu(x.float()), rtol=1e-2, atol=1e-2)


def check_swin_block_fused(device):
    """``SwinTransformerBlock._forward_fused`` (frozen bf16 path) vs the plain fp32 block."""
    import copy
    from flpr_b200.models.swin import SwinTransformerBlock
    torch.manual_seed(3)
    for (dim, res, heads, ws, shift) in ((96, (14, 14), 3, 7, 3), (192, (14, 14), 6, 7, 0), (384, (7, 7), 12, 7, 0)):
        blk = SwinTransformerBlock(dim, res, heads, ws, shift).to(device).eval()
        for p in blk.parameters():
            p.requires_grad_(False)
        with torch.no_grad():
            for p in blk.parameters():
                if p.dim() == 1:
                    p.add_(torch.randn_like(p) * 0.05)
        x = torch.randn(2, res[0] * res[1], dim).to(device)
        with torch.no_grad():
            ref = blk(x)
            fast = copy.deepcopy(blk)
            if device == "cpu":
                fast = fast.to(torch.bfloat16)
            else:
                from flpr_b200.models.swin import use_tensor_core_linears
                use_tensor_core_linears(fast)
                assert fast._fused_ok(x.to(torch.bfloat16)), "fused path not taken"
            out = fast._forward_fused(x.to(torch.bfloat16))
        assert out.dtype == torch.bfloat16 and out.shape == ref.shape
        close(out, ref, rtol=5e-2, atol=5e-2 * ref.abs().max().item())


# ------------------------------------------------------------------------------------------------ trainable LayerNorm
def check_layer_norm_rows(device):
    """``ops.layer.layer_norm_rows`` (bf16 rows, fp32 statistics, optional shifted-window destination) forward and
    backward vs fp32 autograd through ``F.layer_norm`` + the window gather."""
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(41)
    for (b, h, w, ws, shift, c) in ((2, 8, 4, 4, 0, 768), (2, 14, 14, 7, 3, 96), (9, 8, 4, 4, 1, 1536), (1, 4, 4, 4, 0, 8)):
        rows = b * h * w
        x0 = (torch.randn(rows, c, generator=g) * 2 + 0.5).to(device).to(torch.bfloat16)
        gamma0 = (1 + 0.2 * torch.randn(c, generator=g)).to(device)
        beta0 = (0.2 * torch.randn(c, generator=g)).to(device)
        dy = torch.randn(rows, c, generator=g).to(device).to(torch.bfloat16)
        for window in (None, (h, w, ws, shift)):
            x = x0.clone().requires_grad_(True)
            gm, bt = gamma0.clone().requires_grad_(True), beta0.clone().requires_grad_(True)
            y = lops.layer_norm_rows(x, gm, bt, 1e-5, window)
            assert y.dtype == torch.bfloat16 and y.shape == x.shape
            y.backward(dy)
            xr = x0.float().requires_grad_(True)
            gr, br = gamma0.clone().requires_grad_(True), beta0.clone().requires_grad_(True)
            src = xr if window is None else xr[lops.window_src_rows(rows, *window, device=xr.device)]
            yr = F.layer_norm(src, (c,), gr, br, 1e-5)
            yr.backward(dy.float())
            close(y, yr)
            close(x.grad, xr.grad)
            close(gm.grad, gr.grad, rtol=2e-2, atol=1e-2 * gr.grad.abs().max().item() + 1e-4)
            close(bt.grad, br.grad, rtol=1e-3, atol=1e-3 * br.grad.abs().max().item() + 1e-4)


def check_swin_train_block_norms(device):
    """A trainable Swin block whose ``norm1`` / ``norm2`` were switched to ``TrainLayerNorm``: loss and parameter
    gradients of one step vs the fp32 module (on the CPU the class falls through to ``nn.LayerNorm``: plumbing only)."""
    import copy
    from flpr_b200.models.swin import SwinTransformerBlock, TrainLayerNorm, use_tensor_core_linears
    torch.manual_seed(5)
    blk = SwinTransformerBlock(192, (8, 4), 6, window_size=4, shift_size=2, drop_path=0.0).to(device)
    ref = copy.deepcopy(blk).float()
    n = use_tensor_core_linears(blk)
    assert n == 4 and type(blk.norm1) is TrainLayerNorm and type(blk.norm2) is TrainLayerNorm
    x = torch.randn(6, 32, 192).to(device)
    tgt = torch.randn(6, 32, 192).to(device)
    blk.train()
    ref.train()
    if device == "cpu":
        out = blk(x)
    else:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = blk(x.to(torch.bfloat16))
    loss = F.mse_loss(out.float(), tgt)
    loss.backward()
    rl = F.mse_loss(ref(x), tgt)
    rl.backward()
    assert abs(loss.item() - rl.item()) <= 3e-2 * abs(rl.item()) + 1e-4
    for name in ("norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "attn.qkv.weight"):
        a, b = blk.get_parameter(name).grad, ref.get_parameter(name).grad
        assert a is not None, name
        cos = F.cosine_similarity(a.float().flatten(), b.flatten(), dim=0).item()
        assert cos > 0.98, (name, cos)


def check_gelu_act(device):
    """``ops.layer.gelu_act`` (exact GELU, native forward / backward from the saved pre-activation) vs ``F.gelu`` autograd."""
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(61)
    for shape in ((37, 384), (5, 32, 3072), (8,)):
        x0 = (torch.randn(*shape, generator=g) * 2).to(device).to(torch.bfloat16)
        dy = torch.randn(*shape, generator=g).to(device).to(torch.bfloat16)
        x = x0.clone().requires_grad_(True)
        y = lops.gelu_act(x)
        y.backward(dy)
        xr = x0.float().requires_grad_(True)
        yr = F.gelu(xr)
        yr.backward(dy.float())
        close(y, yr, rtol=2e-2, atol=2e-2)
        close(x.grad, xr.grad, rtol=2e-2, atol=2e-2)


def check_window_merge_residual(device):
    """``ops.layer.window_merge_residual`` (window reverse + roll back + per-sample drop-path factor + residual, one op)
    forward and both gradients vs the block's original formulation."""
    from flpr_b200.models.swin import window_reverse
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(51)
    for (b, h, w, ws, shift, c) in ((3, 8, 8, 4, 2, 96), (2, 14, 14, 7, 0, 192), (4, 8, 4, 4, 1, 768)):
        rows = b * h * w
        win0 = torch.randn(rows, c, generator=g).to(device).to(torch.bfloat16)
        sc0 = torch.randn(rows, c, generator=g).to(device).to(torch.bfloat16)
        dy = torch.randn(rows, c, generator=g).to(device).to(torch.bfloat16)
        for scale in (None, (torch.rand(b, generator=g) < 0.7).float().div(0.7).to(device)):
            win, sc = win0.clone().requires_grad_(True), sc0.clone().requires_grad_(True)
            out = lops.window_merge_residual(win, sc, scale, h, w, ws, shift)
            out.backward(dy)
            wr, sr = win0.float().requires_grad_(True), sc0.float().requires_grad_(True)
            back = window_reverse(wr.view(-1, ws * ws, c), ws, h, w)
            if shift:
                back = torch.roll(back, shifts=(shift, shift), dims=(1, 2))
            back = back.reshape(b, h * w, c)
            if scale is not None:
                back = back * scale.view(b, 1, 1)
            ref = sr + back.reshape(rows, c)
            ref.backward(dy.float())
            close(out, ref, rtol=2e-2, atol=2e-2)
            close(win.grad, wr.grad, rtol=2e-2, atol=2e-2)
            close(sc.grad, sr.grad, rtol=0, atol=0)


def check_swin_train_block_fused(device):
    """``SwinTransformerBlock._forward_fused_train`` (norm1 into the window layout, merge + drop-path + residual in one op)
    vs the plain block: outputs and every parameter / input gradient, without and with stochastic depth (same seed ->
    same per-sample masks). On the CPU both run in fp32 through the reference forms of the ops: a tight comparison of
    the autograd glue and the index mathematics; on CUDA the fused side runs the kernels in bf16."""
    import copy
    from flpr_b200.models.swin import SwinTransformerBlock
    for (dim, res, heads, ws, shift, dp) in ((96, (8, 8), 3, 4, 2, 0.0), (192, (8, 4), 6, 4, 0, 0.3), (96, (14, 14), 3, 7, 3, 0.2)):
        torch.manual_seed(7)
        blk = SwinTransformerBlock(dim, res, heads, ws, shift, drop_path=dp).to(device).train()
        ref = copy.deepcopy(blk)
        bsz = 5
        x0 = torch.randn(bsz, res[0] * res[1], dim).to(device)
        tgt = torch.randn(bsz, res[0] * res[1], dim).to(device)
        xr = x0.clone().requires_grad_(True)
        torch.manual_seed(99)
        out_r = ref(xr)
        F.mse_loss(out_r, tgt).backward()
        if device == "cpu":
            xf = x0.clone().requires_grad_(True)
            torch.manual_seed(99)
            out_f = blk._forward_fused_train(xf)
            tol, gtol = 1e-5, 1e-4
        else:
            from flpr_b200.models.swin import use_tensor_core_linears
            use_tensor_core_linears(blk)
            xf = x0.to(torch.bfloat16).requires_grad_(True)
            assert blk._fused_train_ok(xf) or not __import__("flpr_b200.ops.layer", fromlist=["x"]).enabled("ln_train", device)
            torch.manual_seed(99)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out_f = blk._forward_fused_train(xf)
            tol, gtol = 5e-2, 5e-2
        F.mse_loss(out_f.float(), tgt).backward()
        if device != "cpu" and dp > 0.0:
            # the two sides draw their stochastic-depth masks from tensors of different dtype (fp32 module vs bf16 path):
            # whether the device generator then yields the same mask is an implementation detail of ``bernoulli_`` - the
            # mask logic itself is compared exactly on the CPU; here the kernels only have to run and stay finite
            assert torch.isfinite(out_f.float()).all() and torch.isfinite(xf.grad.float()).all()
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in blk.parameters())
            continue
        close(out_f, out_r, rtol=tol, atol=tol * out_r.abs().max().item())
        close(xf.grad, xr.grad, rtol=gtol * 10, atol=gtol * xr.grad.abs().max().item())
        for (name, pf), (_, pr) in zip(blk.named_parameters(), ref.named_parameters()):
            assert pf.grad is not None, name
            if device == "cpu":
                close(pf.grad, pr.grad, rtol=1e-3, atol=gtol * pr.grad.abs().max().item() + 1e-7)
            else:
                cos = F.cosine_similarity(pf.grad.float().flatten(), pr.grad.flatten(), dim=0).item()
                assert cos > 0.97, (name, cos)


# ------------------------------------------------------------------------------------------------ dispatch apply
def check_apply_global(device):
    """``ops.layer.apply_global`` (master <- global + bf16 copy + FedProx anchor in one pass) vs the three copies."""
    from flpr_b200.ops import layer as lops
    g = torch.Generator().manual_seed(31)
    for n, total, mode in ((4096, 5000, 1), (1 << 18, 1 << 18, 2), (64, 64, 0)):
        flat = torch.randn(n, generator=g).to(device)
        m0 = torch.randn(total, generator=g).to(device)
        res = []
        for fn in (lops.apply_global_ref, lops.apply_global):
            master, p_old = m0.clone(), torch.zeros(total).to(device)
            shadow = torch.zeros(total, dtype=torch.bfloat16).to(device)
            fn(flat, master, shadow, p_old if mode else None, mode)
            res.append((master, p_old, shadow))
        for a, b in zip(*res):
            assert torch.equal(a, b)
        master, p_old, shadow = res[1]
        assert torch.equal(master[:n], flat) and torch.equal(master[n:], m0[n:])
        assert torch.equal(shadow[:n], flat.to(torch.bfloat16))
        if mode == 1:
            assert torch.equal(p_old[:n], m0[:n])
        elif mode == 2:
            assert torch.equal(p_old[:n], flat)


# ------------------------------------------------------------------------------------------------ convergence
def run_tiny_rounds(tmp: str, method: str, device: str, dtype: str, rounds: int = 10):
    """One tiny deterministic experiment (every split fits one batch, no augmentation, SGD, shared initial weights):
    returns ``{client: [tr_loss of round 1..rounds]}``."""
    import os
    from helpers import tiny_common, tiny_experiment
    from flpr_b200.data.synthetic import synthetic_source_factory
    from flpr_b200.models import nets
    from flpr_b200.runtime.experiment import ExperimentStage
    os.makedirs(tmp, exist_ok=True)
    common = tiny_common(tmp, device=device)
    d = common["defaults"]
    d["exp_opts"].update(comm_rounds=rounds, val_interval=rounds)
    d["optimizer_opts"] = {"name": "sgd", "lr": 0.01, "momentum": 0.9, "weight_decay": 1e-4}
    d["task_opts"]["train_epochs"] = 2
    d["task_opts"]["augment_opts"].update(level="none", img_size=[64, 32])
    d["task_opts"]["loader_opts"]["batch_size"] = 64
    torch.manual_seed(1234)
    init = os.path.join(tmp, "init.pt")
    torch.save({"*": nets["resnet18"](num_classes=8000, last_stride=1, neck="bnneck").state_dict()}, init)
    cfg = tiny_experiment(common, method, engine_opts={"compute_dtype": dtype, "init_state": init,
                                                       "val_at_round0": False, "client_threads": False})
    with ExperimentStage(common, [cfg], source_factory=synthetic_source_factory(
            num_ids=4, train_per_id=4, size=(64, 32))) as stage:
        log = stage.run_experiment(cfg)
    out = {}
    for client, per_round in log.records["data"].items():
        rows = sorted(((int(r), tasks) for r, tasks in per_round.items()), key=lambda x: x[0])
        out[client] = [float(next(v["tr_loss"] for v in tasks.values() if "tr_loss" in v)) for _, tasks in rows
                       if any("tr_loss" in v for v in tasks.values())]
    return out


_REF_CURVES: dict = {}


def check_bf16_engine_tracks_fp32(tmp: str, method: str, device: str, per_round: float = 0.10, mean_tol: float = 0.05,
                                  rounds: int = 10, dtype: str = "bf16"):
    """The bf16 engine on ``device`` (native kernels) follows the fp32 CPU engine round by round on the same
    experiment: a wrong gradient / optimizer / aggregation kernel shows up as a diverging loss curve within a few
    rounds, which the finite-metric assertions of the e2e tests cannot see."""
    import os
    key = (method, rounds)
    if key not in _REF_CURVES:                       # the fp32 CPU curve is shared by the bf16 and the fp32 device arms
        _REF_CURVES[key] = run_tiny_rounds(os.path.join(tmp, "fp32"), method, "cpu", "fp32", rounds)
    ref = _REF_CURVES[key]
    got = run_tiny_rounds(os.path.join(tmp, "dev"), method, device, dtype if device != "cpu" else "fp32", rounds)
    assert ref.keys() == got.keys() and all(len(v) == rounds for v in ref.values()), (ref, got)
    for client in ref:
        a, b = torch.tensor(ref[client]), torch.tensor(got[client])
        assert a.shape == b.shape, (client, ref[client], got[client])
        rel = (a - b).abs() / a.abs().clamp_min(1e-6)
        assert float(rel.max()) <= per_round and float(rel.mean()) <= mean_tol, (method, client, ref[client], got[client])
    return ref, got
