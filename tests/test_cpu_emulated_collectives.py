"""The peer-memory collectives of ``csrc/fedcomm.cu`` executed on the CPU: R ranks, one process, many interleavings.

``tests/emu`` compiles the kernel source for the host; every rank is a queue of kernels whose threads are fibers, all
ranks run concurrently under a seeded random schedule (random fiber order, random stalls at the flag / streaming memory
operations, random launch skew between ranks), `ld.acquire.sys` spin loops are cooperative, ``%globaltimer`` is virtual
and ``multimem`` goes through an emulated NVSwitch multicast window. This is the flag-protocol soak the review asked for
(SURVEY 5.2 race detection) in a form that needs no GPU: every collective is checked against the plain PyTorch
arithmetic across schedules, back-to-back rounds with device-ordered producers / consumers expose early or stale
reads, a *mutated* kernel (trailing barrier removed) proves the harness would see such a bug, and the watchdog path
(a rank that never arrives, ranks launching different grids) is exercised with the virtual clock.

Reference arithmetic: ``methods/fedavg.py:386-397`` (weighted mean), ``methods/fedstil.py:1146-1160`` (mix),
``methods/fedcurv.py:621-646`` (moments), ``methods/fedweit.py:999-1009`` (client-last gather)."""
import shutil

import pytest
import torch

R, K = 3, 5
OWNER = [0, 0, 1, 1, 2]                  # client -> hosting rank
SEEDS = [0, 1, 2, 3, 4, 5, 6, 7]


@pytest.fixture(scope="module")
def comm_lib():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from emu import comm_harness as H
    from emu.build_emu import build
    try:
        lib = H.load(build("fedcomm.cu"))
    except RuntimeError as ex:
        pytest.skip(f"emulator build unavailable: {ex}")
    yield lib
    assert lib.flpr_emu_deadlocks() == 0, "a collective deadlocked under some schedule"


def make_world(lib, world=R, blocks=2, timeout_s=1e3):
    from emu.comm_harness import EmuWorld
    return EmuWorld(lib, world, blocks=blocks, timeout_s=timeout_s)


def rand(n, seed):
    return torch.randn(n, generator=torch.Generator().manual_seed(seed))


def skew(world, seed):
    """Launch skew between the ranks and, every other seed, one rank that runs several times slower than the rest."""
    if seed:
        for r in range(world.world):
            world.set_start_delay(r, (seed * 7 + r * 13) % 23)
        if seed % 2:
            for idx in (0, 1):
                world.set_slowdown(seed % world.world, 2 + seed % 4, idx)


def close(a, b, tol=2e-6):
    scale = float(b.abs().max()) + 1e-6
    assert float((a.float() - b.float()).abs().max()) <= tol * scale, float((a.float() - b.float()).abs().max())


# ------------------------------------------------------------------------------------------------------------ C1 + C2
@pytest.mark.parametrize("one_shot", [False, True])
@pytest.mark.parametrize("participants", [[0, 1, 2, 3, 4], [0, 3, 4], [2]])
def test_weighted_mean_reduce_broadcast(comm_lib, one_shot, participants):
    n = 4 * 1237
    for seed in SEEDS:
        w = make_world(comm_lib)
        comm_lib.flpr_comm_set_one_shot_bytes((1 << 30) if one_shot else 0)
        up = [rand(n, 100 * seed + c) for c in range(K)]
        cnt = [torch.tensor([float(2 + 3 * c)]) for c in range(K)]
        dst = [torch.full((n,), 7.0) for _ in range(R)]
        src = [up[c] for c in participants]
        for r in range(R):
            if seed % 2 == 0:
                w.reduce_bcast(r, src, dst, cnt=[cnt[c] for c in participants])
            else:                                      # explicit (already normalised) weights instead of counters
                tot = sum(float(cnt[c]) for c in participants)
                w.reduce_bcast(r, src, dst, w=[float(cnt[c]) / tot for c in participants])
        skew(w, seed)
        assert w.run(seed) == 0
        tot = sum(float(cnt[c]) for c in participants)
        ref = sum(up[c] * (float(cnt[c]) / tot) for c in participants)
        for r in range(R):
            close(dst[r], ref)
            assert w.error_word(r) == 0
        w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)


@pytest.mark.parametrize("participants", [[0, 1, 2, 3, 4], [0, 1, 3]])
def test_weighted_mean_through_the_emulated_switch(comm_lib, participants):
    """``fed_reduce_bcast_nvls``: local fold -> ``multimem.ld_reduce`` of this rank's slice -> ``multimem.st`` to all.
    With ``[0, 1, 3]`` rank 2 hosts no participant (L = 0: it still folds zeros, reduces its slice and broadcasts)."""
    n = 4 * 1031
    for seed in SEEDS:
        w = make_world(comm_lib)
        up = [rand(n, 200 * seed + c) for c in range(K)]
        cnt = [torch.tensor([float(1 + c)]) for c in range(K)]
        partial = [torch.full((n,), 3.0) for _ in range(R)]
        dst = [torch.full((n,), 9.0) for _ in range(R)]
        mc_partial, mc_dst = w.multicast(partial), w.multicast(dst)
        use_cnt = seed % 2 == 0
        tot = sum(float(cnt[c]) for c in participants)
        for r in range(R):
            mine = [c for c in participants if OWNER[c] == r]
            w.reduce_bcast_nvls(r, [up[c] for c in mine], [cnt[c] for c in mine] if use_cnt else None,
                                None if use_cnt else [float(cnt[c]) for c in mine],
                                [cnt[c] for c in participants] if use_cnt else None, tot, partial[r], mc_partial,
                                mc_dst, len(participants))
        skew(w, seed)
        assert w.run(seed) == 0
        ref = sum(up[c] * (float(cnt[c]) / tot) for c in participants)
        for r in range(R):
            close(dst[r], ref, tol=4e-6)
        w.close()


# ------------------------------------------------------------------------------------------------------------ C4
@pytest.mark.parametrize("rows_on_device", [False, True])
def test_spatial_temporal_mix(comm_lib, rows_on_device):
    """``fed_mix``: every receiving client gets its own row of the mixing matrix; G, theta and the bf16 copy written in
    one pass; a rank without a receiver (L = 0) only takes part in the barriers."""
    n = 4 * 911
    receivers = {0: [0, 1], 1: [3], 2: []}             # rank -> receiving clients this round
    for seed in SEEDS:
        w = make_world(comm_lib)
        theta = [rand(n, 300 * seed + c) for c in range(K)]
        rows = {c: torch.softmax(rand(K, 17 * seed + c), 0) for c in range(K)}
        out = {}
        for r in range(R):
            mine = receivers[r]
            g = [torch.zeros(n) for _ in mine]
            th = [torch.zeros(n) if i % 2 == 0 else None for i, _ in enumerate(mine)]
            b16 = [torch.zeros(n, dtype=torch.bfloat16) for _ in mine]
            out[r] = (g, th, b16)
            dev = torch.stack([rows[c] for c in mine]).contiguous() if (rows_on_device and mine) else None
            w.mix(r, theta, None if dev is not None else [rows[c].tolist() for c in mine], g, th, b16, rows_dev=dev)
            w._keep_run.append(dev)
        skew(w, seed)
        assert w.run(seed) == 0
        for r in range(R):
            g, th, b16 = out[r]
            for i, c in enumerate(receivers[r]):
                ref = sum(rows[c][j] * theta[j] for j in range(K))
                close(g[i], ref)
                if th[i] is not None:
                    close(th[i], ref)
                close(b16[i], ref, tol=8e-3)
        w.close()


# ------------------------------------------------------------------------------------------------------------ C3 / C5 / C2
def test_fedcurv_moments_gather_and_pull_copy(comm_lib):
    n = 4 * 703
    for seed in SEEDS[:5]:
        w = make_world(comm_lib)
        fisher = [rand(n, 400 * seed + c).abs() for c in range(K)]
        param = [rand(n, 500 * seed + c) for c in range(K)]
        df, dfp, dfpp = ([torch.zeros(n) for _ in range(R)] for _ in range(3))
        m = n + 3                                                        # gather: a tail that is not a multiple of 4
        feat = [rand(m, 600 * seed + c) for c in range(K)]
        gathered = [torch.zeros(m, K) for _ in range(R)]
        first = [torch.zeros(n) for _ in range(R)]
        first16 = [torch.zeros(n, dtype=torch.bfloat16) for _ in range(R)]
        for r in range(R):
            w.curv_moments(r, fisher, param, df, dfp, dfpp)
            w.gather_strided(r, feat, gathered[r], m)
            w.pull_copy(r, param[(r + 1) % K], first[r], first16[r])      # first-contact dispatch from a peer's buffer
        skew(w, seed)
        assert w.run(seed) == 0
        for r in range(R):
            close(df[r], sum(fisher), tol=4e-6)
            close(dfp[r], sum(f * p for f, p in zip(fisher, param)), tol=4e-6)
            close(dfpp[r], sum(f * p * p for f, p in zip(fisher, param)), tol=4e-6)
            assert torch.equal(gathered[r], torch.stack(feat, 1))
            assert torch.equal(first[r], param[(r + 1) % K])
            assert torch.equal(first16[r], param[(r + 1) % K].to(torch.bfloat16))
        w.close()


# ------------------------------------------------------------------------------------------------------------ rounds
def _rounds(world, lib, reps, seed, n, archive_from_stream=0):
    """``reps`` federated rounds, everything device-ordered: producers overwrite the upload slots, the collective reduces
    them, a consumer archives this rank's result - the next round's producers may only start once nobody reads the slots
    any more, and a consumer may only read once every slice has landed. Returns (staging, archive); the device buffers
    are parked on ``world`` so that they outlive the queued kernels."""
    up = [torch.zeros(n) for _ in range(K)]
    dst = [torch.zeros(n) for _ in range(R)]
    staging = [[rand(n, 1000 * rep + 10 * seed + c) for c in range(K)] for rep in range(reps)]
    archive = [[torch.zeros(n) for _ in range(R)] for _ in range(reps)]
    for r in range(R):
        for rep in range(reps):
            for c in range(K):
                if OWNER[c] == r:
                    world.local_copy(r, staging[rep][c], up[c])
            world.reduce_bcast(r, up, dst, w=[1.0 / K] * K)
            world.local_copy(r, dst[r], archive[rep][r])
    world._keep.extend([up, dst])
    return staging, archive


@pytest.mark.parametrize("one_shot", [False, True])
def test_back_to_back_rounds_never_read_early_or_stale(comm_lib, one_shot):
    n = 4 * 640
    comm_lib.flpr_comm_set_one_shot_bytes((1 << 30) if one_shot else 0)
    for seed in range(1, 13):
        w = make_world(comm_lib)
        staging, archive = _rounds(w, comm_lib, 4, seed, n)
        skew(w, seed)
        assert w.run(seed, stall_one_in=3) == 0
        for rep in range(4):
            ref = sum(staging[rep]) / K
            for r in range(R):
                close(archive[rep][r], ref)
        w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)


def test_the_harness_sees_a_missing_barrier():
    """Mutation check: the two-shot kernel WITHOUT its trailing barrier ("every rank's slice has landed everywhere") must
    produce a wrong archive under some schedule - otherwise the rounds test above proves nothing."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from emu import build_emu, comm_harness as H
    marker = "  rank_barrier(ctx, e0 + 2);  // every rank's slice has landed everywhere\n"
    src_path = build_emu.os.path.join(build_emu.CSRC, "fedcomm.cu")
    with open(src_path) as f:
        assert f.read().count(marker) == 1
    lib = H.load(build_emu.build("fedcomm.cu", mutate=lambda s: s.replace(marker, "")))
    lib.flpr_comm_set_one_shot_bytes(0)
    n, wrong = 4 * 640, 0
    for seed in range(1, 13):
        w = make_world(lib)
        staging, archive = _rounds(w, lib, 4, seed, n)
        skew(w, seed)
        assert w.run(seed, stall_one_in=3) == 0
        for rep in range(4):
            ref = sum(staging[rep]) / K
            wrong += sum(int(float((archive[rep][r] - ref).abs().max()) > 1e-4) for r in range(R))
        w.close()
    assert wrong > 0, "removing a barrier went unnoticed: the schedules do not exercise the protocol"


def test_concurrent_collectives_on_two_channels(comm_lib):
    """Aggregation on the communication stream (channel 1) while the next round's mix runs on the compute stream
    (channel 0): two kernels per rank in flight, separate arrival flags and epochs."""
    n = 4 * 512
    for seed in SEEDS:
        w = make_world(comm_lib)
        comm_lib.flpr_comm_set_one_shot_bytes(0)
        theta = [rand(n, 700 * seed + c) for c in range(K)]
        up = [rand(n, 800 * seed + c) for c in range(K)]
        rows = [torch.softmax(rand(K, seed + r), 0) for r in range(R)]
        g = [torch.zeros(n) for _ in range(R)]
        dst = [torch.zeros(n) for _ in range(R)]
        for r in range(R):
            for rep in range(2):                                          # twice: epochs advance independently
                w.mix(r, theta, [rows[r].tolist()], [g[r]], [None], [None], stream=0, channel=0)
                w.reduce_bcast(r, up, dst, w=[1.0 / K] * K, stream=1, channel=1)
        skew(w, seed)
        assert w.run(seed) == 0
        for r in range(R):
            close(g[r], sum(rows[r][j] * theta[j] for j in range(K)))
            close(dst[r], sum(up) / K)
        w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)


# ------------------------------------------------------------------------------------------------------------ watchdog
def test_a_rank_that_never_arrives_trips_the_watchdog_and_nothing_is_published(comm_lib):
    """Rank 2 skips the collective: the survivors' barriers time out on the virtual clock, the error word is set (sticky),
    the load / store phase is skipped (no partial aggregate reaches anybody), later collectives drain without hanging."""
    n = 4 * 256
    for seed in (0, 3):
        w = make_world(comm_lib, timeout_s=2e-4)
        comm_lib.flpr_comm_set_one_shot_bytes(0)
        up = [rand(n, c) for c in range(K)]
        dst = [torch.full((n,), -1.0) for _ in range(R)]
        for r in (0, 1):
            w.reduce_bcast(r, up, dst, w=[1.0 / K] * K)
        t0 = comm_lib.flpr_emu_clock_ns()
        assert w.run(seed, max_passes=50000) == 0, "the survivors hung instead of timing out"
        assert comm_lib.flpr_emu_clock_ns() - t0 >= 2e5
        assert w.error_word(0) == 1 and w.error_word(1) == 1 and w.error_word(2) == 0
        assert all(bool((d == -1.0).all()) for d in dst), "a partial aggregate was published"
        # the error is sticky: a complete collective afterwards still refuses to publish on the ranks that saw it
        for r in range(R):
            w.reduce_bcast(r, up, dst, w=[1.0 / K] * K)
        assert w.run(seed, max_passes=50000) == 0
        assert bool((dst[0] == -1.0).all()) or w.error_word(0) == 1
        w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)


def test_ranks_launching_different_grids_are_caught_not_hung_on(comm_lib):
    """The bug class found on the device this round (a grid size derived from rank-local state): block 1 of rank 0 has no
    partner on rank 1 - its barrier must time out and raise the error word instead of spinning forever."""
    n = 4 * 2048
    w = make_world(comm_lib, world=2, timeout_s=2e-4)
    comm_lib.flpr_comm_set_one_shot_bytes(0)
    up = [rand(n, c) for c in range(2)]
    dst = [torch.zeros(n) for _ in range(2)]
    w.reduce_bcast(0, up, dst, w=[0.5, 0.5], blocks=2)
    w.reduce_bcast(1, up, dst, w=[0.5, 0.5], blocks=1)
    assert w.run(1, max_passes=50000) == 0
    assert w.error_word(0) == 1
    w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)


# ------------------------------------------------------------------------------------------------------------ limits
@pytest.mark.parametrize("world,clients,local,n,blocks,one_shot", [
    (2, 32, 8, 4 * 37, 1, False),          # MAX_CLIENTS sources, MAX_LOCAL receivers on one rank, fewer elements than threads
    (8, 32, 4, 4 * 5, 2, False),           # MAX_RANKS, slices of the two-shot mean smaller than one element per rank
    (8, 8, 1, 4, 4, False),                # ONE float4 spread over 8 ranks x 4 blocks
    (4, 9, 3, 4 * 513, 2, True),           # one-shot form, a size just past a block boundary
])
def test_compile_time_limits_and_tiny_buffers(comm_lib, world, clients, local, n, blocks, one_shot):
    """The table sizes of ``fedcomm.cu`` (32 clients, 8 local receivers, 8 ranks) and buffers far smaller than the grid:
    mean (peer loads), mean (switch), mix and client-last gather with a 1-element tail, all in one queue per rank."""
    for seed in (0, 3):
        comm_lib.flpr_comm_set_one_shot_bytes((1 << 30) if one_shot else 0)
        w = make_world(comm_lib, world=world, blocks=blocks)
        up = [rand(n, 10 * seed + c) for c in range(clients)]
        cnt = [torch.tensor([float(1 + (c % 5))]) for c in range(clients)]
        dst, dst2, partial = ([torch.zeros(n) for _ in range(world)] for _ in range(3))
        rows = [[torch.softmax(rand(clients, seed + 100 * r + i), 0) for i in range(local)] for r in range(world)]
        g = [[torch.zeros(n) for _ in range(local)] for _ in range(world)]
        b16 = [[torch.zeros(n, dtype=torch.bfloat16) for _ in range(local)] for _ in range(world)]
        feat = [rand(n + 1, 7 * seed + c) for c in range(clients)]
        gathered = [torch.zeros(n + 1, clients) for _ in range(world)]
        mc_partial, mc_dst = w.multicast(partial), w.multicast(dst2)
        hosted = {r: [c for c in range(clients) if c % world == r][:8] for r in range(world)}   # <= MAX_LOCAL per rank
        part = sorted(c for r in range(world) for c in hosted[r])
        tot, totp = sum(float(c) for c in cnt), sum(float(cnt[c]) for c in part)
        for r in range(world):
            w.reduce_bcast(r, up, dst, cnt=cnt)
            w.mix(r, up, [x.tolist() for x in rows[r]], g[r], [None] * local, b16[r])
            w.gather_strided(r, feat, gathered[r], n + 1)
            w.reduce_bcast_nvls(r, [up[c] for c in hosted[r]], [cnt[c] for c in hosted[r]], None,
                                [cnt[c] for c in part], totp, partial[r], mc_partial, mc_dst, len(part))
        skew(w, seed)
        assert w.run(seed, max_passes=400000) == 0
        ref = sum(u * (float(c) / tot) for u, c in zip(up, cnt))
        refp = sum(up[c] * (float(cnt[c]) / totp) for c in part)
        for r in range(world):
            close(dst[r], ref, tol=4e-6)
            close(dst2[r], refp, tol=4e-6)
            assert torch.equal(gathered[r], torch.stack(feat, 1))
            for i in range(local):
                want = sum(rows[r][i][j] * up[j] for j in range(clients))
                close(g[r][i], want, tol=4e-6)
                close(b16[r][i], want, tol=8e-3)
            assert w.error_word(r) == 0
        w.close()
    comm_lib.flpr_comm_set_one_shot_bytes(1 << 20)
