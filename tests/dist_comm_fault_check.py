"""Fault injection for the peer-memory collectives (run under torchrun on >= 2 GPUs): one rank "dies" - it skips a
collective the others enter. The survivors' kernels must not hang and must not publish a partial aggregate: the flag
watchdog fires after ``timeout_s``, the store phase is skipped (destination buffers untouched), the sticky error word /
host mailbox turns the next ``poll_errors`` / ``check_errors`` into a ``NativeError``.

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 2 tests/dist_comm_fault_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from flpr_b200.ops.native import NativeError  # noqa: E402
from flpr_b200.parallel.comm import FedComm  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for n, label in ((1 << 16, "small (one-shot / two-shot path)"), (1 << 21, "8 MB (NVLS path when the switch offers it)")):
        comm = FedComm(dev, world, arena_bytes=64 << 20, timeout_s=1.0)
        comm.alloc_client_buffer("up", n)
        comm.alloc_rank_buffer("glob", n)
        clients = list(range(world))
        for c in comm.local_clients():
            comm.client_view("up", c).fill_(float(c + 1))
        w = [1.0 / world] * world
        # a healthy collective first: every rank takes part
        comm.reduce_bcast("up", "glob", clients, weights=w)
        torch.cuda.synchronize()
        comm.check_errors()
        want = sum(c + 1 for c in clients) / world
        if not torch.allclose(comm.rank_view("glob"), torch.full((n,), want, device=dev)):
            ok = False
            print(f"[rank {rank}] healthy reduce wrong ({label})", flush=True)
        comm.rank_view("glob").fill_(-7.0)
        torch.cuda.synchronize()
        dist.barrier()
        # the last rank dies: it never enters the next collective
        if rank != world - 1:
            t0 = time.perf_counter()
            raised = False
            try:
                comm.reduce_bcast("up", "glob", clients, weights=w)
                torch.cuda.synchronize()
                comm.check_errors()
            except NativeError as ex:
                raised = "timed out" in str(ex)
            dt = time.perf_counter() - t0
            untouched = bool((comm.rank_view("glob") == -7.0).all().item())
            if not (raised and dt < 30.0 and untouched):
                ok = False
                print(f"[rank {rank}] fault not handled ({label}): raised={raised} after {dt:.1f}s, "
                      f"destination untouched={untouched}", flush=True)
            elif rank == 0:
                print(f"FAULT {label}: watchdog fired after {dt:.2f}s, no partial aggregate published", flush=True)
        dist.barrier()                       # the "dead" rank waits here while the survivors time out
        comm.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_COMM_FAULT_CHECK OK" if flag.item() else "DIST_COMM_FAULT_CHECK FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
