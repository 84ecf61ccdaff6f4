"""Multi-rank check of FedComm (run under torchrun). CPU: gloo plumbing mode. CUDA: NVLink P2P kernels.

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 2 tests/dist_comm_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flpr_b200.parallel.comm import FedComm  # noqa: E402


def main():
    use_cuda = torch.cuda.is_available() and os.environ.get("FLPR_FORCE_CPU", "0") != "1"
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dist.init_process_group("gloo")
        dev = torch.device("cpu")
    K = int(os.environ.get("FLPR_K", "8"))
    n = int(os.environ.get("FLPR_N", str(4096 * 33)))
    comm = FedComm(dev, K, arena_bytes=max(64 << 20, n * 4 * (4 * ((K + world - 1) // world) + 8)))
    comm.alloc_client_buffer("up", n)
    comm.alloc_client_buffer("cnt", 4)
    comm.alloc_client_buffer("fisher", n)
    for nm in ("glob", "f", "fp", "fpp"):
        comm.alloc_rank_buffer(nm, n)
    gen = torch.Generator().manual_seed(1234)
    U = torch.randn(K, n, generator=gen)
    Fs = torch.rand(K, n, generator=gen)
    cnts = torch.arange(K, dtype=torch.float32) + 10
    for c in comm.local_clients():
        comm.client_view("up", c).copy_(U[c])
        comm.client_view("fisher", c).copy_(Fs[c])
        comm.client_view("cnt", c).fill_(float(cnts[c]))
    clients = list(range(K))
    ok = True

    def close(x, ref, tol=1e-4, what=""):
        nonlocal ok
        err = (x.float().cpu() - ref).abs().max().item()
        if not err <= tol:
            ok = False
            print(f"[rank {rank}] MISMATCH {what}: max err {err}", flush=True)

    if rank == 0:
        print(f"COMM backend={comm.backend} mc={'yes' if getattr(comm, '_mc_base', 0) else 'no'} "
              f"note={getattr(comm, 'backend_note', '')!r}", flush=True)
    for it in range(3):  # repeat: exercises the epoch counters
        w = cnts / cnts.sum()
        for nvls in ((True, False) if getattr(comm, "_mc_base", 0) else (False,)):
            comm.nvls, comm.nvls_min_bytes = nvls, 0
            comm.rank_view("glob").zero_()
            comm.reduce_bcast("up", "glob", clients, cnt="cnt")
            close(comm.rank_view("glob"), (w[:, None] * U).sum(0), 1e-4, f"reduce_bcast nvls={nvls}")
            sub = clients[1::2]                                        # stale / partial participation, explicit weights
            ws = [0.5 / len(sub)] * len(sub)
            comm.reduce_bcast("up", "glob", sub, weights=ws)
            close(comm.rank_view("glob"), 0.5 * U[sub].mean(0), 1e-4, f"reduce_bcast weights nvls={nvls}")
        # a second channel (a collective on a communication stream next to one on the compute stream)
        if use_cuda:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                comm.set_channel(1)
                comm.reduce_bcast("up", "glob", clients, cnt="cnt")
                comm.set_channel(0)
            torch.cuda.current_stream().wait_stream(side)
            close(comm.rank_view("glob"), (w[:, None] * U).sum(0), 1e-4, "reduce_bcast channel 1")
        rows = torch.softmax(torch.randn(K, K, generator=torch.Generator().manual_seed(it)), dim=1)
        loc = comm.local_clients()
        og = [torch.empty(n, device=dev) for _ in loc]
        ob = [torch.empty(n, device=dev, dtype=torch.bfloat16) for _ in loc]
        comm.mix("up", clients, rows[loc], loc, og, None, ob)
        ref = rows @ U
        for i, c in enumerate(loc):
            close(og[i], ref[c], 1e-4, "mix")
            close(ob[i], ref[c], 5e-2, "mix bf16")
        comm.curv_moments("fisher", "up", clients, "f", "fp", "fpp")
        close(comm.rank_view("f"), Fs.sum(0), 1e-4, "curv f")
        close(comm.rank_view("fp"), (Fs * U).sum(0), 1e-3, "curv fp")
        close(comm.rank_view("fpp"), (Fs * U * U).sum(0), 1e-3, "curv fpp")
        out = torch.empty(n, K, device=dev)
        comm.gather_strided("up", clients, out)
        close(out, U.t(), 0.0, "gather")
        d = torch.empty(n, device=dev)
        comm.pull("up", (it * 3 + 1) % K, d)
        close(d, U[(it * 3 + 1) % K], 0.0, "pull")
        # the next iteration overwrites nothing, but uploads change to make stale reads visible
        comm.barrier()
        U = U + 1.0
        for c in comm.local_clients():
            comm.client_view("up", c).copy_(U[c])

    if use_cuda:
        torch.cuda.synchronize()
        comm.check_errors()
        # bandwidth probe: FedAvg-sized reduce (ResNet-50 head = 31.3 M floats)
        nb = int(os.environ.get("FLPR_BW_N", str(31_326_208)))
        nb -= nb % 4
        comm2 = FedComm(dev, K, arena_bytes=nb * 4 * (comm.slots + 2) + (16 << 20))
        comm2.alloc_client_buffer("theta", nb)
        comm2.alloc_rank_buffer("g", nb)
        for c in comm2.local_clients():
            comm2.client_view("theta", c).normal_()
        loc = comm2.local_clients()
        dst = [torch.empty(nb, device=dev) for _ in loc]
        rows = torch.softmax(torch.randn(K, K), dim=1)
        res = {}

        def rb(nvls):
            comm2.nvls = nvls
            comm2.reduce_bcast("theta", "g", clients, weights=[1.0 / K] * K)
        cases = [("reduce_bcast", lambda: rb(False)), ("mix", lambda: comm2.mix("theta", clients, rows[loc], loc, dst,
                                                                               None, None))]
        if getattr(comm2, "_mc_base", 0):
            cases.insert(1, ("reduce_bcast_nvls", lambda: rb(True)))
        for name, fn in cases:
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[name] = ms.item()
        comm2.check_errors()
        if rank == 0:
            S = nb * 4
            rb_bytes = 2 * S * (world - 1) / world            # ingress: pulled slice + pushed slices
            mix_bytes = S * (K - len(loc))
            print(f"BW world={world} K={K} S={S/1e6:.1f}MB reduce_bcast {res['reduce_bcast']:.3f} ms "
                  f"({rb_bytes/res['reduce_bcast']/1e6:.1f} GB/s ingress/rank) mix {res['mix']:.3f} ms "
                  f"({mix_bytes/res['mix']/1e6:.1f} GB/s ingress/rank)", flush=True)
            if "reduce_bcast_nvls" in res:
                print(f"BW world={world} K={K} S={S/1e6:.1f}MB reduce_bcast_nvls {res['reduce_bcast_nvls']:.3f} ms "
                      f"(switch does the {world}-way add; each GPU sends and receives S once: "
                      f"{S/res['reduce_bcast_nvls']/1e6:.1f} GB/s per direction)", flush=True)
        comm2.close()
    comm.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("DIST_COMM_CHECK", "OK" if flag.item() == 0 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
