"""2-rank CPU (gloo) check of the gallery-sharded evaluation: every rank holds a slice of the gallery (uneven slices,
queries without any match, matches on one shard only) and must obtain the CMC / mAP of the full gallery."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flpr_b200.ops.rank import evaluate_sharded, rank_metrics_reference, similarity  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for seed, (nq, ng, ids) in enumerate([(40, 157, 12), (7, 30, 40), (16, 64, 3)]):
        g = torch.Generator().manual_seed(seed)
        qf = torch.nn.functional.normalize(torch.randn(nq, 32, generator=g), dim=1)
        gf = torch.nn.functional.normalize(torch.randn(ng, 32, generator=g), dim=1)
        ql = torch.randint(0, ids, (nq,), generator=g)
        gl = torch.randint(0, ids, (ng,), generator=g)
        ql[0] = ids + 5                                        # a query nobody matches
        cut = [0, ng // 3, ng] if world == 2 else [ng * r // world for r in range(world + 1)]
        lo, hi = cut[rank], cut[rank + 1]
        cmc, m_ap = evaluate_sharded(qf, ql, gf[lo:hi], gl[lo:hi])
        ref_cmc, ref_map = rank_metrics_reference(similarity(qf, gf), ql, gl)
        ok &= bool(np.allclose(cmc, ref_cmc, atol=1e-12)) and abs(m_ap - ref_map) < 1e-9
    flag = torch.tensor([int(ok)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_EVAL", "OK" if int(flag) else "FAILED", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
