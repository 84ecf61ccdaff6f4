"""The product's ``FedComm`` (``parallel/comm.py``) in ``p2p`` mode on the CPU: R instances in one process on host arenas,
launching the emulated ``fedcomm.cu`` kernels (``tests/emu/comm_harness.make_fedcomm_world``).

What runs unchanged is the Python half of the communication layer - symmetric buffer placement, owner / slot maps,
peer pointer tables, the kernel choice (two-shot / one-shot / NVLS, decided from rank-invariant quantities), grid
sizes, the launch loop of the mix when a rank hosts more receivers than one kernel takes - i.e. the code between the
method plug-ins and the kernels that only ever executed on multi-GPU boxes. Results are compared with the plain
arithmetic on every rank under seeded random interleavings."""
import shutil

import pytest
import torch


@pytest.fixture(scope="module")
def lib():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from emu import comm_harness as H
    from emu.build_emu import build
    try:
        handle = H.load(build("fedcomm.cu"))
    except RuntimeError as ex:
        pytest.skip(f"emulator build unavailable: {ex}")
    yield handle
    assert handle.flpr_emu_deadlocks() == 0


def rand(n, seed):
    return torch.randn(n, generator=torch.Generator().manual_seed(seed))


def world_of(lib, world, clients, **kw):
    from emu import comm_harness as H
    return H.make_fedcomm_world(lib, world, clients, **kw)


def run(lib, seed):
    from emu import comm_harness as H
    assert H.run(lib, seed) == 0


def done(lib):
    from emu import comm_harness as H
    H.end(lib)
    lib.flpr_comm_set_one_shot_bytes(1 << 20)


def close(a, b, tol=4e-6):
    assert float((a.float() - b.float()).abs().max()) <= tol * (float(b.abs().max()) + 1e-6)


@pytest.mark.parametrize("form", ["two_shot", "one_shot", "nvls"])
@pytest.mark.parametrize("world,clients", [(3, 7), (2, 8)])
def test_weighted_mean_through_fedcomm(lib, form, world, clients):
    n = 4 * 300
    comms = world_of(lib, world, clients)
    lib.flpr_comm_set_one_shot_bytes(0 if form == "two_shot" else (1 << 30))
    for c in comms:
        c.alloc_client_buffer("up", n)
        c.alloc_client_buffer("cnt", 4)
        c.alloc_rank_buffer("glob", n)
        c.nvls_min_bytes = 0 if form == "nvls" else (1 << 40)
    ups = [rand(n, c) for c in range(clients)]
    cnts = [float(2 + c) for c in range(clients)]
    for c in comms:
        assert c.local_clients() == [i for i in range(clients) if i % world == c.rank]
        for cid in c.local_clients():
            c.client_view("up", cid).copy_(ups[cid])
            c.client_view("cnt", cid).fill_(cnts[cid])
    for seed, part, use_cnt in ((0, list(range(clients)), True), (3, [0, 2, clients - 1], True), (5, [1, 2, 3], False)):
        tot = sum(cnts[i] for i in part)
        for c in comms:
            if use_cnt:
                c.reduce_bcast("up", "glob", part, cnt="cnt")
            else:
                c.reduce_bcast("up", "glob", part, weights=[cnts[i] / tot for i in part])
        run(lib, seed)
        ref = sum(ups[i] * (cnts[i] / tot) for i in part)
        for c in comms:
            close(c.rank_view("glob"), ref)
            assert c.error_word() == 0
            assert (c.nvls_launches > 0) == (form == "nvls")
    done(lib)


def test_mix_launch_loop_when_a_rank_hosts_more_receivers_than_one_kernel_takes(lib):
    """20 clients on 2 ranks: 10 receivers per rank > MAX_LOCAL = 8 - ``FedComm.mix`` issues two launches per rank (the
    same number on every rank: the kernels barrier across ranks), the second with the remaining receivers."""
    world, clients, n = 2, 20, 4 * 120
    comms = world_of(lib, world, clients)
    for c in comms:
        c.alloc_client_buffer("theta_up", n)
    theta = [rand(n, 10 + c) for c in range(clients)]
    for c in comms:
        for cid in c.local_clients():
            c.client_view("theta_up", cid).copy_(theta[cid])
    order = list(range(clients))
    outs = {}
    for c in comms:
        mine = c.local_clients()
        rows = torch.stack([torch.softmax(rand(clients, 100 + cid), 0) for cid in mine])
        g = [torch.zeros(n) for _ in mine]
        th = [torch.zeros(n) for _ in mine]
        b16 = [torch.zeros(n, dtype=torch.bfloat16) for _ in mine]
        outs[c.rank] = (mine, rows, g, th, b16)
        c.mix("theta_up", order, rows, mine, g, th, b16)
    run(lib, 4)
    for c in comms:
        mine, rows, g, th, b16 = outs[c.rank]
        for i, cid in enumerate(mine):
            ref = sum(rows[i][j] * theta[j] for j in range(clients))
            close(g[i], ref)
            close(th[i], ref)
            close(b16[i], ref, tol=8e-3)
    done(lib)


def test_moments_gather_and_first_contact_pull_through_fedcomm(lib):
    world, clients, n = 3, 5, 4 * 200
    comms = world_of(lib, world, clients)
    for c in comms:
        c.alloc_client_buffer("fisher", n)
        c.alloc_client_buffer("param", n)
        for name in ("f", "fp", "fpp"):
            c.alloc_rank_buffer(name, n)
    fisher = [rand(n, c).abs() for c in range(clients)]
    param = [rand(n, 50 + c) for c in range(clients)]
    for c in comms:
        for cid in c.local_clients():
            c.client_view("fisher", cid).copy_(fisher[cid])
            c.client_view("param", cid).copy_(param[cid])
    part = [0, 1, 3, 4]
    gathered = [torch.zeros(n, len(part)) for _ in comms]
    pulled = [torch.zeros(n) for _ in comms]
    pulled16 = [torch.zeros(n, dtype=torch.bfloat16) for _ in comms]
    for c in comms:
        c.curv_moments("fisher", "param", part, "f", "fp", "fpp")
        c.gather_strided("param", part, gathered[c.rank])
        c.pull("param", 4, pulled[c.rank], pulled16[c.rank])
    run(lib, 6)
    for c in comms:
        close(c.rank_view("f"), sum(fisher[i] for i in part))
        close(c.rank_view("fp"), sum(fisher[i] * param[i] for i in part))
        close(c.rank_view("fpp"), sum(fisher[i] * param[i] ** 2 for i in part))
        assert torch.equal(gathered[c.rank], torch.stack([param[i] for i in part], 1))
        assert torch.equal(pulled[c.rank], param[4]) and torch.equal(pulled16[c.rank], param[4].to(torch.bfloat16))
        assert c.bytes_moved > 0
    done(lib)


def test_a_missing_rank_surfaces_as_native_error_on_the_host(lib):
    """Rank 2 never enters the collective: the watchdog of the survivors' kernels fires on the virtual clock, mirrors the
    error word into the host mailbox, and ``FedComm.poll_errors`` - what the engine calls after every collective and at
    every round boundary - raises on exactly those ranks; nothing was published."""
    from flpr_b200.ops.native import NativeError
    world, clients, n = 3, 6, 4 * 100
    comms = world_of(lib, world, clients, timeout_s=2e-4)
    lib.flpr_comm_set_one_shot_bytes(0)
    for c in comms:
        c.alloc_client_buffer("up", n)
        c.alloc_rank_buffer("glob", n)
        for cid in c.local_clients():
            c.client_view("up", cid).copy_(rand(n, cid))
        c.rank_view("glob").fill_(-1.0)
    for c in comms[:2]:
        c.reduce_bcast("up", "glob", list(range(clients)), weights=[1.0 / clients] * clients)
        c.poll_errors()                                   # nothing has run yet: clean
    run(lib, 2)
    for c in comms[:2]:
        with pytest.raises(NativeError, match="timed out waiting for a peer rank"):
            c.poll_errors()
        assert bool((c.rank_view("glob") == -1.0).all())
    comms[2].poll_errors()                                # the absent rank saw nothing
    done(lib)
