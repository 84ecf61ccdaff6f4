import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flpr_b200.ops import native
from flpr_b200.ops.gemm import gemm
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from gemm_bench import timeit
lib = native.load()
for (m, n, k) in [(8192, 2048, 512), (8192, 2048, 2048)]:
    a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    for bn in (256, 128):
        for dbg in (0, 1, 2, 4, 5):
            os.environ["FLPR_GEMM_DEBUG"] = str(dbg)
            t = timeit(lambda: gemm(a, b, bn=bn))
            t32 = timeit(lambda: gemm(a, b, bn=bn, out_dtype=torch.float32))
            print(f"{(m,n,k)} bn={bn} debug={dbg}: bf16-out {t:7.1f} us ({2.0*m*n*k/t/1e6:7.1f} TF)   fp32-out {t32:7.1f} us")
os.environ["FLPR_GEMM_DEBUG"] = "0"
for (m, n, k) in [(8192, 2048, 512), (512, 2048, 8192)]:
    a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    for gen in ("", "1"):
        if gen:
            os.environ["FLPR_GEMM_GENERIC_EPI"] = "1"
        else:
            os.environ.pop("FLPR_GEMM_GENERIC_EPI", None)
        for bn in (256, 128):
            t = timeit(lambda: gemm(a, b, bn=bn))
            t32 = timeit(lambda: gemm(a, b, bn=bn, out_dtype=torch.float32))
            o = torch.zeros(m, n, device="cuda")
            ta = timeit(lambda: gemm(a, b, bn=bn, out_dtype=torch.float32, split_k=4, out=o))
            print(f"{(m,n,k)} bn={bn} generic_epi={gen or 0}: bf16 {t:7.1f} us  fp32 {t32:7.1f} us  fp32 split-K4 atomic {ta:7.1f} us")
os.environ.pop("FLPR_GEMM_GENERIC_EPI", None)
