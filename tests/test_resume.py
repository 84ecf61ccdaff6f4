"""Resume manifest: N rounds in one go == K rounds, restart from the manifest, N-K more rounds."""
import copy
import os

import pytest
import torch

from flpr_b200.runtime.experiment import ExperimentStage
from helpers import tiny_common, tiny_experiment, tiny_factory


def _run(tmp, method, rounds, resume=False, interval=0):
    common = tiny_common(tmp)
    common["defaults"]["exp_opts"].update(comm_rounds=rounds, val_interval=100)
    common["defaults"]["task_opts"]["sustain_rounds"] = 2
    cfg = tiny_experiment(common, method)
    cfg["engine_opts"].update(resume=resume, resume_interval=interval, val_at_round0=False)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        from flpr_b200.utils.misc import same_seeds
        same_seeds(cfg["random_seed"])
        store, comm, server, clients, names = stage.build(cfg)
        from flpr_b200.runtime import resume as R
        from flpr_b200.runtime.explog import ExperimentLog
        from flpr_b200.utils.misc import DeviceTimer
        log = ExperimentLog(os.path.join(tmp, "log.json"), enabled=False)
        timer = DeviceTimer(stage.device)
        first = 1
        if resume and R.available(store, 0):
            first = R.load(stage, store, server, clients, comm) + 1
        for r in range(first, rounds + 1):
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
            if interval and r % interval == 0:
                R.save(stage, store, r, server, clients, comm)
        store.flush()
        out = {"server": server.model.arena.master.clone(),
               "clients": [c.model.arena.master.clone() for c in clients],
               "pipe": [{k: v for k, v in c.task_pipeline.state_dict().items() if k != "loader_rng"} for c in clients],
               "cnt": [c.train_cnt for c in clients]}
        store.close()
        if comm is not None:
            comm.close()
    return out, first


@pytest.mark.parametrize("method", ["fedavg", "fedstil", "fedprox"])
def test_resume_continues_where_it_stopped(tmp_path, method):
    full, _ = _run(str(tmp_path / "a"), method, 4)
    part, first = _run(str(tmp_path / "b"), method, 2, interval=2)
    assert first == 1
    cont, first = _run(str(tmp_path / "b"), method, 4, resume=True, interval=2)
    assert first == 3                                            # continued after the manifest of round 2
    assert cont["pipe"] == full["pipe"] and cont["cnt"] == full["cnt"]
    assert torch.allclose(cont["server"], full["server"], atol=1e-5)
    for a, b in zip(cont["clients"], full["clients"]):
        assert torch.allclose(a, b, atol=1e-5)


# ---------------------------------------------------------------------------------------------------- rank loss
def _torchrun(tmp, rounds, fault=None, timeout=600):
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(__file__), "dist_resume_check.py")
    env = dict(os.environ, FLPR_TMP=str(tmp), OMP_NUM_THREADS="2")
    env.pop("FLPR_FAULT_EXIT", None)
    if fault:
        env["FLPR_FAULT_EXIT"] = fault
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                           "--nproc-per-node", "2", script, "fedavg", str(rounds)], env=env, capture_output=True,
                          text=True, timeout=timeout)


def _results(tmp):
    return [torch.load(os.path.join(str(tmp), f"result_rank{r}.pt"), weights_only=False) for r in (0, 1)]


def _flat(state):
    return torch.cat([v.float().flatten() for _, v in sorted(state.items()) if torch.is_tensor(v)])


@pytest.fixture(scope="module")
def uninterrupted(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("resume_full")
    r = _torchrun(tmp, 4)
    assert r.returncode == 0 and r.stdout.count("DIST_RESUME") == 2, r.stdout[-2000:] + r.stderr[-2000:]
    return _results(tmp)


@pytest.mark.parametrize("phase", ["round", "saving", "saved"])
def test_rank_loss_then_restart_resumes_consistently(tmp_path, uninterrupted, phase):
    """World 2 (gloo): rank 1 exits hard in round 3 - after the round but before its manifest (``round``), with the
    snapshot written but not committed (``saving``), or right after its commit marker (``saved``); the job dies with
    it. A restart on the same directories must pick the newest manifest committed on BOTH ranks (never a torn or a
    one-sided one) and finish with exactly the weights of the uninterrupted run."""
    crashed = _torchrun(tmp_path, 4, fault=f"1:3:{phase}")
    assert crashed.returncode != 0, "the fault was not injected"
    assert "DIST_RESUME" not in crashed.stdout
    again = _torchrun(tmp_path, 4)
    assert again.returncode == 0 and again.stdout.count("DIST_RESUME") == 2, again.stdout[-2000:] + again.stderr[-2000:]
    assert "Resumed from the manifest of round" in again.stdout + again.stderr
    for full, got in zip(uninterrupted, _results(tmp_path)):
        assert got["cnt"] == full["cnt"]
        assert torch.allclose(_flat(got["server"]), _flat(full["server"]), atol=1e-5)
        for name in full["clients"]:
            assert torch.allclose(_flat(got["clients"][name]), _flat(full["clients"][name]), atol=1e-5)


def test_manifest_generations_and_commit_markers(tmp_path):
    """Single process: two generations alternate, the generation being overwritten loses its marker first, a file without
    a marker (a crash between snapshot and commit) is ignored by ``load``."""
    from flpr_b200.runtime import resume as R
    from flpr_b200.runtime.checkpoint import CheckpointStore
    _run(str(tmp_path), "fedavg", 3, interval=1)
    store = CheckpointStore(os.path.join(str(tmp_path), "ckpts", "t-fedavg"), asynchronous=False)
    committed = R._committed(store, 0)
    assert sorted(committed) == [2, 3] and sorted(committed.values()) == [0, 1]
    os.remove(R._marker(store, 0, committed[3]))                     # as if the process had died before committing round 3
    assert sorted(R._committed(store, 0)) == [2]

    class Stage:
        world, rank = 1, 0
    assert R.agreed_round(Stage(), store) == 2
    cont, first = _run(str(tmp_path), "fedavg", 4, resume=True, interval=1)
    assert first == 3
    full, _ = _run(str(tmp_path / "full"), "fedavg", 4)
    assert torch.allclose(cont["server"], full["server"], atol=1e-5)


def test_resume_reshards_to_a_different_world_size(tmp_path):
    """A job written by 2 ranks is continued by 1 (a GPU lost for good): the single rank reads both old ranks' committed
    manifests from the shared directory, takes over both clients (models, counters, task positions, loader shuffle
    state, upload slots) and the replicated server state, trains on, and its first commit retires the old world's files.
    (Bitwise equality with the 2-rank continuation is not defined - host RNG streams are per process - so the check is
    that what was restored IS what was saved.)"""
    import subprocess
    import sys
    from flpr_b200.runtime import resume as R
    from flpr_b200.runtime.explog import ExperimentLog
    from flpr_b200.utils.misc import DeviceTimer, same_seeds
    script = os.path.join(os.path.dirname(__file__), "dist_resume_check.py")
    env = dict(os.environ, FLPR_TMP=str(tmp_path), FLPR_SHARED_ROOT="1", OMP_NUM_THREADS="2")
    env.pop("FLPR_FAULT_EXIT", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", "2", script, "fedavg", "2"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("DIST_RESUME") == 2, r.stdout[-2000:] + r.stderr[-2000:]

    tmp = os.path.join(str(tmp_path), "shared")
    common = tiny_common(tmp)
    common["defaults"]["exp_opts"].update(comm_rounds=4, val_interval=100)
    common["defaults"]["task_opts"]["sustain_rounds"] = 2
    cfg = tiny_experiment(common, "fedavg", n_clients=2, n_tasks=2)
    cfg["engine_opts"].update(resume=True, resume_interval=1, val_at_round0=False)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        same_seeds(cfg["random_seed"])
        store, comm, server, clients, names = stage.build(cfg)
        assert R.written_world(store) == 2 and stage.world == 1
        assert R.agreed_round(stage, store) == 2
        saved = [store.load(R.ACTOR, R._gen_name(rk, R._committed(store, rk)[2])) for rk in (0, 1)]
        assert R.load(stage, store, server, clients, comm) == 2
        by_name = {c.client_name: c for c in clients}
        assert set(by_name) == {"client-0", "client-1"}
        for st in saved:
            for name, cs in st["clients"].items():                   # every client of every old rank landed here
                c = by_name[name]
                assert c.train_cnt == cs["train_cnt"]
                pipe = c.task_pipeline.state_dict()
                assert pipe["current_task_idx"] == cs["pipeline"]["current_task_idx"]
                assert pipe["task_round_rest"] == cs["pipeline"]["task_round_rest"]
                mine = c.model.full_state()
                for k, v in cs["model"].items():
                    if torch.is_tensor(v):
                        assert torch.equal(mine[k].cpu().float(), v.float()), (name, k)
            for bname, val in (st.get("comm") or {}).items():
                if isinstance(val, dict):
                    for cid, t in val.items():
                        assert torch.equal(comm.client_view(bname, int(cid)).cpu(), t), (bname, cid)
        sv = server.model.full_state()
        for k, v in saved[0]["server"]["model"].items():
            if torch.is_tensor(v):
                assert torch.equal(sv[k].cpu().float(), v.float()), k
        log = ExperimentLog(os.path.join(tmp, "log.json"), enabled=False)
        timer = DeviceTimer(stage.device)
        for rnd in (3, 4):
            stage._process_one_round(rnd, server, clients, names, cfg, log, timer, comm)
            R.save(stage, store, rnd, server, clients, comm)
        assert all(torch.isfinite(c.model.arena.master).all() for c in clients)
        assert R.written_world(store) == 1 and sorted(R._committed(store, 0)) == [3, 4]
        assert not R._committed(store, 1), "the old world's manifests were not retired"
        store.close()
        if comm is not None:
            comm.close()
