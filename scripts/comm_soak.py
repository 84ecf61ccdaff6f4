"""Soak of the cross-rank flag protocol (run under torchrun on >= 2 GPUs; gloo plumbing mode on CPU):

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 8 scripts/comm_soak.py [iters]

Randomised rank arrival (every rank sleeps a random time on the host AND delays its stream with a spin kernel of random
length before every collective), randomised participation (``online < K`` with explicit weights, stale slots of the
others), randomly interleaved collectives of different sizes (so consecutive launches use different grids and both the
two-shot, one-shot and NVLS paths), a second flag channel on a side stream in a random subset of the iterations. Every
result is checked against a host-side fp64 reference built from the same seeds. Reports iterations, mismatches and the
watchdog state; exit code 1 on any mismatch.
"""
import os
import random
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flpr_b200.parallel.comm import FedComm  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    use_cuda = torch.cuda.is_available() and os.environ.get("FLPR_FORCE_CPU", "0") != "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dev = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    K = 8
    sizes = {"tok": 4096 * 2, "mid": 4096 * 40, "big": 4096 * 900}          # 32 KB (one-shot), 640 KB, 14.7 MB
    comm = FedComm(dev, K, arena_bytes=sum(sizes.values()) * 4 * (((K + world - 1) // world) + 2) + (32 << 20))
    for nm, n in sizes.items():
        comm.alloc_client_buffer(nm, n)
        comm.alloc_rank_buffer("g_" + nm, n)
    shared = random.Random(1234)                      # identical decisions on every rank
    private = random.Random(99 + rank)                # arrival jitter differs per rank
    vals = {nm: torch.zeros(K) for nm in sizes}       # slot c of buffer nm currently holds the constant vals[nm][c]
    bad = 0
    side = torch.cuda.Stream() if use_cuda else None

    def jitter():
        time.sleep(private.random() * 0.004)
        if use_cuda and private.random() < 0.5:
            torch.cuda._sleep(int(private.random() * 2e6))      # up to ~1 ms of device-side delay

    for it in range(iters):
        nm = shared.choice(list(sizes))
        # refresh a random subset of the slots (the others stay stale, like offline clients)
        online = sorted(shared.sample(range(K), shared.randint(1, K)))
        for c in online:
            v = float(shared.randint(-50, 50))
            vals[nm][c] = v
            if comm.owner(c) == rank:
                comm.client_view(nm, c).fill_(v)
        part = sorted(shared.sample(range(K), shared.randint(1, K)))
        w = [shared.random() + 0.1 for _ in part]
        tot = sum(w)
        w = [x / tot for x in w]
        use_side = use_cuda and shared.random() < 0.3
        comm.nvls = shared.random() < 0.5
        comm.nvls_min_bytes = 0
        jitter()
        if use_side:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                comm.set_channel(1)
                comm.reduce_bcast(nm, "g_" + nm, part, weights=w)
                comm.set_channel(0)
            torch.cuda.current_stream().wait_stream(side)
        else:
            comm.reduce_bcast(nm, "g_" + nm, part, weights=w)
        want = sum(wi * float(vals[nm][c]) for wi, c in zip(w, part))
        got = comm.rank_view("g_" + nm)
        lo, hi = float(got.min()), float(got.max())
        if abs(lo - want) > 1e-3 or abs(hi - want) > 1e-3:
            bad += 1
            if bad <= 5:
                print(f"[rank {rank}] it {it} {nm} part={part} want {want:.5f} got [{lo:.5f}, {hi:.5f}] nvls={comm.nvls}",
                      flush=True)
        if it % 7 == 0:                                # a mix in between: different kernel, different grid
            loc = comm.local_clients()
            rows = torch.tensor([[shared.random() for _ in range(K)] for _ in range(K)])
            outs = [torch.empty(sizes[nm], device=dev) for _ in loc]
            jitter()
            comm.mix(nm, list(range(K)), rows[loc].to(dev) if use_cuda else rows[loc], loc, outs, None, None)
            for i, c in enumerate(loc):
                want = float((rows[c].double() * vals[nm].double()).sum())
                lo, hi = float(outs[i].min()), float(outs[i].max())
                if abs(lo - want) > 2e-3 * max(1.0, abs(want)) or abs(hi - want) > 2e-3 * max(1.0, abs(want)):
                    bad += 1
        comm.poll_errors()
    if use_cuda:
        torch.cuda.synchronize()
    comm.check_errors()
    flag = torch.tensor([bad], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print(f"COMM_SOAK {'OK' if flag.item() == 0 else 'FAILED'} iters={iters} world={world} "
              f"backend={getattr(comm, 'backend', comm.mode)} mismatches={int(flag.item())}", flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
