"""Asynchronous checkpoint pipeline (shared pinned arena + writer processes) on CPU tensors."""
import os

import pytest
import numpy as np
import torch

from flpr_b200.runtime.checkpoint import CheckpointStore, _SharedArena


def test_arena_allocator_coalesces():
    a = _SharedArena(1 << 20, cuda=False)
    offs = [a.alloc(100_000) for _ in range(10)]
    assert all(o >= 0 for o in offs) and a.alloc(200_000) < 0
    for o in offs[::2]:
        a.release_region(o, 100_000)
    assert a.alloc(150_000) < 0                      # fragmented
    for o in offs[1::2]:
        a.release_region(o, 100_000)
    assert a.free == [(0, 1 << 20)]                  # fully coalesced
    assert a.alloc(1 << 20) == 0


def test_async_store_roundtrip(tmp_path):
    st = CheckpointStore(str(tmp_path), asynchronous=True, workers=3, arena_bytes=16 << 20)
    state = {"a": torch.randn(1000, 33), "k": 3,
             "n": {"b": torch.arange(10), "c": [torch.ones(3, dtype=torch.bfloat16), 5, "x"], "e": torch.zeros(0)}}
    for i in range(40):                               # 40 x 132 KB through a 16 MB arena: regions are recycled
        st.save("client-0", f"s{i}", state, True)
    gens = {"_compact_gens": [{"pids": torch.tensor([7, 9]), "bank": torch.randn(2, 3, 4, 2, 2).bfloat16(),
                               "cls": torch.tensor([[1, 2, 3], [4, 5, 6]]), "k": 2}]}
    st.save("client-0", "ex", gens, True, post="expand_examplars")
    st.save("client-0", "big", {"x": torch.zeros(20 << 20, dtype=torch.uint8)}, True)   # > arena: written in-line
    st.flush()
    assert st._arena.free == [(0, 16 << 20)]
    out = st.load("client-0", "s7")
    assert torch.equal(out["a"], state["a"]) and torch.equal(out["n"]["b"], state["n"]["b"])
    assert out["n"]["c"][1:] == [5, "x"] and out["k"] == 3
    assert out["n"]["c"][0].dtype == torch.bfloat16 and out["n"]["e"].numel() == 0
    # the file holds the payload only, not the whole arena
    assert os.path.getsize(os.path.join(str(tmp_path), "client-0", "s7.ckpt")) < 200_000
    ex = st.load("client-0", "ex")
    assert sorted(int(k) for k in ex) == [7, 9] and len(ex[np.int64(7)]) == 2 and ex[np.int64(9)][1][1] == 5
    assert isinstance(ex[np.int64(7)][0][0], np.ndarray) and ex[np.int64(7)][0][0].dtype == np.float32
    assert st.load("client-0", "big")["x"].numel() == 20 << 20
    st.close()


def test_async_store_concurrent_writers_with_back_pressure(tmp_path):
    """Client threads snapshot concurrently through an arena that is much smaller than the total volume: saves block on
    the writers (back-pressure), nothing is lost or torn, the arena is fully returned."""
    import threading
    st = CheckpointStore(str(tmp_path), asynchronous=True, workers=4, arena_bytes=8 << 20)
    errors = []

    def client(cid):
        try:
            g = torch.Generator().manual_seed(cid)
            for i in range(25):
                state = {"w": torch.randn(50_000 + 1000 * cid, generator=g), "i": i, "cid": cid,
                         "nested": {"v": torch.full((1000,), float(i * 100 + cid))}}
                st.save(f"client-{cid}", f"s{i}", state, True)          # ~200 KB each, 20 MB in total per client
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=client, args=(c,)) for c in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    st.flush()
    assert st._arena.free == [(0, 8 << 20)]
    for cid in range(4):
        g = torch.Generator().manual_seed(cid)
        for i in range(25):
            want = torch.randn(50_000 + 1000 * cid, generator=g)
            got = st.load(f"client-{cid}", f"s{i}")
            assert got["i"] == i and got["cid"] == cid and torch.equal(got["w"], want)
            assert float(got["nested"]["v"][0]) == float(i * 100 + cid)
    st.close()


def test_dead_writer_is_detected_not_waited_for(tmp_path):
    """Fault injection: every writer process is killed while jobs are in flight; ``flush`` must report it
    within seconds instead of waiting for completions that will never come."""
    import time
    st = CheckpointStore(str(tmp_path), asynchronous=True, workers=2, arena_bytes=8 << 20)
    st.save("client-0", "warm", {"x": torch.zeros(10)}, True)
    st.flush()                                                     # writers are up
    for p in st._procs:
        p.kill()                                                   # exact processes this test started
    for p in st._procs:
        p.join(10)
    for i in range(4):
        st.save("client-0", f"lost{i}", {"x": torch.randn(1000)}, True)
    t0 = time.time()
    with pytest.raises(RuntimeError, match="writer process died"):
        st.flush()
    assert time.time() - t0 < 30
    st._inflight.clear()                                           # nothing left to wait for; release the rest
    try:
        st.close()
    except Exception:
        pass
