"""Frozen-trunk forward (256 images, 256x128, ResNet-50 stem + layer1-3): library vs native residual stages vs fully
native, CUDA-graphed like in the engine; plus the per-kernel breakdown of the fully native path."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flpr_b200.models import resnet as R
from flpr_b200.models.frozen import FoldedTrunk


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    net = R.resnet50(num_classes=8000, last_stride=1, neck="bnneck").cuda().eval()
    net.configure_split(["base.layer4", "classifier"])
    x = torch.randn(256, 3, 256, 128, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    for label, native, stem in (("library (cuDNN)", False, False), ("native stages, library stem", True, False),
                                ("fully native", True, True)):
        ft = FoldedTrunk(net, torch.bfloat16, use_graphs=True)
        ft.native, ft.native_stem = native, stem
        ms = timeit(lambda: ft(x))
        print(f"{label:32s} {ms:8.3f} ms / 256 images")
    ft = FoldedTrunk(net, torch.bfloat16, use_graphs=False)
    ft(x)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ft(x)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            agg[e.name[:90]][0] += 1
            agg[e.name[:90]][1] += e.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"fully native, eager: {tot / 1e3:.3f} ms of kernels")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {us / 1e3:8.3f} ms x{c:3d}  {n}")


if __name__ == "__main__":
    main()
