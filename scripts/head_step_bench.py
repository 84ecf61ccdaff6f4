"""Head training step (layer4 + BNNeck + classifier, batch 64 prototypes 1024x16x8) timing: eager vs CUDA graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flpr_b200.runtime.builder import parser_model, parser_criterion
from flpr_b200.runtime.arena import ArenaOptimizer
from flpr_b200.runtime.graphs import GraphedStep
from flpr_b200.ops import native

def main():
    steps = int(os.environ.get("STEPS", "10"))
    use_graph = os.environ.get("GRAPH", "1") == "1"
    dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
    cfg = {"name": "resnet50", "num_classes": 8000, "last_stride": 1, "neck": "bnneck", "atten_default": 0.9,
           "lambda_l1": 1e-3, "lambda_k": 64, "fine_tuning": ["base.layer4", "classifier"]}
    model = parser_model("fedstil", cfg, dev, {"compute_dtype": "bf16"})
    crit = parser_criterion({"name": "cross_entropy", "num_classes": 8000, "epsilon": 0.1})[0]
    opt = ArenaOptimizer("adam", model.arena, lr=1e-3, weight_decay=1e-5)
    model.install(opt)
    model.train()
    B = int(os.environ.get("BATCH", "64"))
    proto = torch.randn(B, 1024, 16, 8, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 8000, (B,), device=dev)
    def fn(data, target):
        opt.zero_grad()
        with model.autocast():
            score, feat = model.forward_head(data)
        loss = crit(score=score, feature=feat, target=target)
        loss.backward()
        opt.step()
    step = GraphedStep(fn, warmup=2, enabled=use_graph)
    for _ in range(4):
        step(proto, y)
    torch.cuda.synchronize()
    l0 = native.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps):
        step(proto, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = 3 * 2 * B * 128 * (1024*512 + 9*512*512 + 512*2048 + 1024*2048 + 2*(2048*512 + 9*512*512 + 512*2048)) + 3*2*B*2048*8000
    print(f"head step: {ms:.3f} ms/step device, {(time.perf_counter()-t0)*1e3/steps:.3f} ms/step wall, graph={use_graph}, "
          f"{flops/ms/1e9:.1f} TFLOP/s, native launches/step={(native.launches()-l0)/steps:.0f}")

if __name__ == "__main__":
    main()
