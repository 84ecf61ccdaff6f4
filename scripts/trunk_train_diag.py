"""Per-stage agreement of the native train-mode trunk with the nn.Module (fp32 and bf16-autocast references)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flpr_b200.models import resnet as R

torch.manual_seed(4)
net = R.resnet50(num_classes=10, last_stride=1, neck="bnneck").cuda()
net.configure_split(["base.layer4", "classifier"])
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
ref = copy.deepcopy(net).float()
x = torch.randn(16, 3, 256, 128, device="cuda").contiguous(memory_format=torch.channels_last)
net.train(); ref.train()
trunk = R.NativeTrunk(net)


def stats(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (f"cos {torch.nn.functional.cosine_similarity(a, b, dim=0).item():.5f} rel-fro "
            f"{((a - b).norm() / b.norm()).item():.4f} max-err {(a - b).abs().max().item():.3f} ref-max {b.abs().max().item():.2f}")


with torch.no_grad():
    for stop in (1, 2, 3, 4):
        net2 = copy.deepcopy(net); net2.head_start = stop
        t2 = R.NativeTrunk(net2)
        y = t2(x)
        r32 = copy.deepcopy(ref).base.run_stages(x, 0, stop)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            r16 = copy.deepcopy(ref).base.run_stages(x, 0, stop)
        print(f"stages 0..{stop - 1}: native vs fp32: {stats(y, r32)}")
        print(f"             bf16-autocast vs fp32: {stats(r16, r32)}")
