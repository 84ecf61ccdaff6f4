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
