"""Randomised soak of the peer-memory collectives under the SIMT emulator (no GPU): the CPU twin of
``scripts/comm_soak.py``.

    python scripts/emu_comm_soak.py --seeds 200 [--world 8] [--seed0 1]

Every seed draws a world size, a client placement, participation, sizes, grid size, one-shot threshold, launch skew
and a slow rank, queues a few federated rounds (device-ordered producers -> collective -> consumers, mixing the
two-shot / one-shot / NVLS mean with the FedSTIL mix on a second channel) and checks every archived result against
the PyTorch arithmetic. Prints one line per failing seed and a summary; exit code 1 on any mismatch or deadlock."""
import argparse
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from emu import comm_harness as H  # noqa: E402
from emu.build_emu import build  # noqa: E402


def one_seed(lib, seed: int, world_arg: int) -> str:
    rng = random.Random(seed)
    world = world_arg or rng.choice([2, 3, 4, 8])
    K = rng.randint(world, min(2 * world, 12))
    owner = [c % world for c in range(K)]
    rng.shuffle(owner)
    n = 4 * rng.randint(64, 700)
    blocks = rng.choice([1, 2, 3])
    reps = rng.randint(2, 3)
    mode = rng.choice(["two_shot", "one_shot", "nvls"])
    lib.flpr_comm_set_one_shot_bytes((1 << 30) if mode == "one_shot" else 0)
    w = H.EmuWorld(lib, world, blocks=blocks)
    g = torch.Generator().manual_seed(seed)
    up = [torch.zeros(n) for _ in range(K)]
    cnt = [torch.tensor([float(rng.randint(1, 9))]) for _ in range(K)]
    dst = [torch.zeros(n) for _ in range(world)]
    partial = [torch.zeros(n) for _ in range(world)]
    mc_partial, mc_dst = w.multicast(partial), w.multicast(dst)
    mixed = [torch.zeros(n) for _ in range(world)]
    staging = [[torch.randn(n, generator=g) for _ in range(K)] for _ in range(reps)]
    part = [sorted(rng.sample(range(K), rng.randint(1, K))) for _ in range(reps)]
    rows = [[torch.softmax(torch.randn(K, generator=g), 0) for _ in range(world)] for _ in range(reps)]
    arch_mean = [[torch.zeros(n) for _ in range(world)] for _ in range(reps)]
    arch_mix = [[torch.zeros(n) for _ in range(world)] for _ in range(reps)]
    for r in range(world):
        for rep in range(reps):
            for c in range(K):
                if owner[c] == r:
                    w.local_copy(r, staging[rep][c], up[c])
            p = part[rep]
            tot = sum(float(cnt[c]) for c in p)
            if mode == "nvls":
                mine = [c for c in p if owner[c] == r]
                w.reduce_bcast_nvls(r, [up[c] for c in mine], [cnt[c] for c in mine], None, [cnt[c] for c in p], tot,
                                    partial[r], mc_partial, mc_dst, len(p))
            else:
                w.reduce_bcast(r, [up[c] for c in p], dst, cnt=[cnt[c] for c in p])
            w.local_copy(r, dst[r], arch_mean[rep][r])
            # the mix of the same uploads (every rank one receiver); in the product it runs after the uploads of the round
            # as well - same stream here, so it is ordered after the mean and before the next round's producers
            w.mix(r, up, [rows[rep][r].tolist()], [mixed[r]], [None], [None])
            w.local_copy(r, mixed[r], arch_mix[rep][r])
    for r in range(world):
        w.set_start_delay(r, rng.randint(0, 40))
    if rng.random() < 0.6:
        w.set_slowdown(rng.randrange(world), rng.randint(2, 6))
    rc = w.run(seed, max_passes=400000, stall_one_in=rng.choice([2, 3, 5]))
    msg = ""
    if rc != 0:
        msg = "deadlock"
    else:
        for rep in range(reps):
            p = part[rep]
            tot = sum(float(cnt[c]) for c in p)
            ref = sum(staging[rep][c] * (float(cnt[c]) / tot) for c in p)
            for r in range(world):
                if float((arch_mean[rep][r] - ref).abs().max()) > 1e-4:
                    msg = f"mean mismatch rep {rep} rank {r}"
                refm = sum(rows[rep][r][j] * staging[rep][j] for j in range(K))
                if float((arch_mix[rep][r] - refm).abs().max()) > 1e-4:
                    msg = msg or f"mix mismatch rep {rep} rank {r}"
        if any(w.error_word(r) for r in range(world)):
            msg = msg or "watchdog fired"
    w.close()
    cfg = f"world={world} K={K} n={n} blocks={blocks} reps={reps} mode={mode}"
    return f"seed {seed}: {msg} ({cfg})" if msg else ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=50)
    ap.add_argument("--seed0", type=int, default=1)
    ap.add_argument("--world", type=int, default=0)
    a = ap.parse_args()
    lib = H.load(build("fedcomm.cu"))
    t0, bad = time.time(), []
    for seed in range(a.seed0, a.seed0 + a.seeds):
        msg = one_seed(lib, seed, a.world)
        if msg:
            bad.append(msg)
            print(msg, flush=True)
    print(f"EMU_COMM_SOAK {'OK' if not bad else 'FAILED'}: {a.seeds} seeds, {len(bad)} failures, "
          f"{time.time() - t0:.0f} s, deadlocks reported by the emulator: {lib.flpr_emu_deadlocks()}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
