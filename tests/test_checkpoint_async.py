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


# ------------------------------------------------------------------------------------------------- mapped store
def _mapped(tmp_path, **kw):
    from flpr_b200.runtime.mapped_store import MappedCheckpointStore
    return MappedCheckpointStore(str(tmp_path), asynchronous=False, force_mapped=True, **kw)


def test_mapped_store_files_are_plain_torch_checkpoints(tmp_path):
    """Layout emitted by ``legacy_layout`` = a legacy ``torch.save`` container: loadable by ``torch.load`` (also with
    ``weights_only=True``), every dtype / empty tensor / nested container / scalar intact, channels-last views saved in
    logical order."""
    st = _mapped(tmp_path)
    cl = torch.randn(4, 6, 3, 3).contiguous(memory_format=torch.channels_last)
    state = {"train_cnt": 5, "a": torch.randn(1000, 33), "cl": cl, "tok": None, "f": 0.25, "flag": True,
             "n": {"b": torch.arange(10), "c": [torch.ones(3, dtype=torch.bfloat16), 5, "x"], "e": torch.zeros(0),
                   "u8": torch.arange(7, dtype=torch.uint8), "t": (torch.tensor(3.5), torch.tensor([True, False]))}}
    st.save("c0", "m", state, True)
    for wo in (False, True):
        out = torch.load(st.path("c0", "m"), weights_only=wo)
        assert torch.equal(out["a"], state["a"]) and torch.equal(out["cl"], cl) and out["train_cnt"] == 5
        assert out["tok"] is None and out["f"] == 0.25 and out["flag"] is True
        assert out["n"]["c"][0].dtype == torch.bfloat16 and out["n"]["c"][1:] == [5, "x"] and out["n"]["e"].numel() == 0
        assert torch.equal(out["n"]["u8"], state["n"]["u8"]) and out["n"]["t"][0].item() == 3.5
        assert out["n"]["t"][1].tolist() == [True, False]
    st.close()


def test_mapped_store_overwrites_in_place_and_survives_growing_scalars(tmp_path):
    st = _mapped(tmp_path)
    path = st.path("c0", "m")
    state = {"train_cnt": 5, "w": torch.randn(257, 3)}
    st.save("c0", "m", state, True)
    ino, start = os.stat(path).st_ino, st._files[path].layout.data_start
    for cnt in (200, 70_000, 10 ** 12):                          # BININT1 -> BININT2 -> BININT -> LONG1
        state = {"train_cnt": cnt, "w": torch.randn(257, 3)}
        st.save("c0", "m", state, True)
        assert os.stat(path).st_ino == ino and st._files[path].layout.data_start == start
        out = st.load("c0", "m")
        assert out["train_cnt"] == cnt and torch.equal(out["w"], state["w"])
    st.save("c0", "m", {"train_cnt": 1, "w": torch.randn(300, 3)}, True)       # structure change: new mapping
    assert st.load("c0", "m")["w"].shape == (300, 3)
    with pytest.raises(ValueError):
        st.save("c0", "m", state, False)                         # cover=False refuses to overwrite
    st.close()


def test_mapped_store_persistent_state_plan(tmp_path):
    """``PersistentState``: the same dict object over the same buffers - the second save reuses the cached layout (no
    re-analysis), static tensors are not copied again, the file holds a plain dict."""
    from flpr_b200.runtime.checkpoint import PersistentState
    st = _mapped(tmp_path)
    w, frozen = torch.randn(64, 8), torch.randn(32)
    frozen._flpr_static = 0
    state = PersistentState({"w": {"a": w}, "pre": {"f": frozen}})
    state.plan_token = ("m", 0)
    st.save("c0", "m", state, True)
    w.mul_(2.0)
    st.save("c0", "m", state, True)
    out = torch.load(st.path("c0", "m"), weights_only=True)
    assert type(out) is dict and torch.equal(out["w"]["a"], w) and torch.equal(out["pre"]["f"], frozen)
    st.close()


def test_mapped_store_payload_ring_recycles_old_rounds(tmp_path):
    st = _mapped(tmp_path, payload_ring=2)
    for r in range(1, 7):
        st.save("client-0", f"{r}-client-0-server", {"train_cnt": r, "w": torch.full((100,), float(r))}, True)
        st.save("server", f"{r}-server-client-0", {"g": torch.full((50,), float(-r))}, True)
    assert sorted(os.listdir(tmp_path / "client-0")) == ["5-client-0-server.ckpt", "6-client-0-server.ckpt"]
    assert sorted(os.listdir(tmp_path / "server")) == ["5-server-client-0.ckpt", "6-server-client-0.ckpt"]
    assert st.load("client-0", "5-client-0-server")["w"][0] == 5 and st.load("server", "6-server-client-0")["g"][0] == -6
    os.remove(tmp_path / "client-0" / "5-client-0-server.ckpt")            # a janitor got there first
    st.save("client-0", "7-client-0-server", {"train_cnt": 7, "w": torch.full((100,), 7.0)}, True)
    assert st.load("client-0", "7-client-0-server")["train_cnt"] == 7
    st.close()


def test_mapped_store_exemplar_file_keeps_the_reference_schema(tmp_path):
    st = _mapped(tmp_path)
    gens = {"_compact_gens": [
        {"pids": torch.tensor([7, 9]), "bank": torch.randn(2, 3, 4, 2, 2).bfloat16(),
         "cls": torch.tensor([[1, 2, 3], [4, 5, 6]]), "k": 2},
        {"pids": torch.tensor([300]), "bank": torch.randn(1, 2, 4, 2, 2).bfloat16(), "cls": torch.tensor([[1, 700]]),
         "k": 2}]}
    for _ in range(2):                                           # second save: same structure, in place
        gens["_compact_gens"][0]["bank"] = torch.randn(2, 3, 4, 2, 2).bfloat16()
        st.save("c0", "ex", gens, True, post="expand_examplars")
        ex = st.load("c0", "ex")
        assert sorted(int(k) for k in ex) == [7, 9, 300] and [len(ex[np.int64(p)]) for p in (7, 9, 300)] == [2, 2, 2]
        assert ex[np.int64(9)][1][1] == 5 and ex[np.int64(300)][1][1] == 700
        for gi, g in enumerate(gens["_compact_gens"]):
            for pi, pid in enumerate(g["pids"].tolist()):
                for j in range(2):
                    arr = ex[np.int64(pid)][j][0]
                    assert arr.dtype == np.float32 and np.array_equal(arr, g["bank"][pi, j].float().numpy())
    gens["_compact_gens"][0]["k"] = 1                            # reduce_examplars: structure change -> re-laid out
    st.save("c0", "ex", gens, True, post="expand_examplars")
    assert len(st.load("c0", "ex")[np.int64(7)]) == 1
    st.close()


def test_mapped_store_runs_a_whole_experiment(tmp_path):
    """The CPU fedstil experiment through the mapped store: same files, loadable, same values as the in-line store."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import tiny_common, tiny_experiment, tiny_factory
    from flpr_b200.runtime import experiment as E
    from flpr_b200.runtime.mapped_store import MappedCheckpointStore
    outs = {}
    for kind in ("inline", "mapped"):
        common = tiny_common(str(tmp_path / kind))
        cfg = tiny_experiment(common, "fedstil")
        cfg["engine_opts"]["val_at_round0"] = False
        orig = E.CheckpointStore
        if kind == "mapped":
            E.CheckpointStore = lambda root, **kw: MappedCheckpointStore(root, force_mapped=True, **kw)
        try:
            with E.ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
                stage.run_experiment(cfg)
        finally:
            E.CheckpointStore = orig
        root = os.path.join(common["checkpoints_dir"], cfg["exp_name"])
        files = {}
        for d, _, names in os.walk(root):
            for n in names:
                files[os.path.relpath(os.path.join(d, n), root)] = torch.load(os.path.join(d, n), weights_only=False)
        outs[kind] = files
    assert set(outs["inline"]) == set(outs["mapped"])

    def same(a, b):
        if isinstance(a, dict):
            assert set(map(str, a)) == set(map(str, b))
            bk = {str(k): v for k, v in b.items()}
            for k, v in a.items():
                same(v, bk[str(k)])
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                same(x, y)
        elif isinstance(a, (torch.Tensor, np.ndarray)):
            assert torch.allclose(torch.as_tensor(a).float(), torch.as_tensor(b).float(), atol=1e-6)
        else:
            assert a == b or (a != a and b != b)
    for name in outs["inline"]:
        same(outs["inline"][name], outs["mapped"][name])
