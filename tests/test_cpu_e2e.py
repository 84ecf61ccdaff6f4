"""End-to-end CPU runs of every method through ``ExperimentStage`` (tiny synthetic workload), log / checkpoint
schema checks, and the world_size=2 gloo plumbing mode (BASELINE.json config 1)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from flpr_b200.runtime.experiment import ExperimentStage, VirtualContainer
from helpers import tiny_common, tiny_experiment, tiny_factory

ALL = ["baseline", "ewc", "mas", "icarl", "fedavg", "fedprox", "fedcurv", "fedweit", "fedstil", "fedstil-atten"]


def _run(tmp_path, method, **over):
    common = tiny_common(str(tmp_path))
    cfg = tiny_experiment(common, method, **over)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        log = stage.run_experiment(cfg)
    return common, cfg, log


@pytest.mark.parametrize("method", ALL)
def test_method_runs_and_logs(tmp_path, method):
    common, cfg, log = _run(tmp_path, method)
    data = log.records["data"]
    assert set(data) == {"client-0", "client-1"}
    for client, rounds in data.items():
        assert set(rounds) == {"0", "1", "2"}
        r1 = rounds["1"]
        task = next(iter(t for t in r1 if "tr_acc" in r1[t]))
        assert {"tr_acc", "tr_loss", "val_rank_1", "val_rank_3", "val_rank_5", "val_rank_10", "val_map"} <= set(r1[task])
        assert all(0.0 <= r1[task][k] <= 1.0 for k in ("tr_acc", "val_rank_1", "val_map"))
    saved = json.load(open(log.save_path))
    assert saved["config"]["exp_method"] == method and "perf" in saved


def test_fedavg_checkpoint_and_payload_schemas(tmp_path):
    common, cfg, _ = _run(tmp_path, "fedavg")
    root = os.path.join(common["checkpoints_dir"], cfg["exp_name"])
    model = torch.load(os.path.join(root, "client-0", "fedavg_model.ckpt"), weights_only=False)
    assert "net.base.layer4.0.conv1.weight" in model and "net.classifier.weight" in model        # 'net.' prefix
    up = torch.load(os.path.join(root, "client-0", "2-client-0-server.ckpt"), weights_only=False)
    assert set(up) == {"train_cnt", "incremental_model_params"}
    assert "net.base.layer4.1.conv2.weight" in up["incremental_model_params"]          # fedavg.py:232-237
    assert up["incremental_model_params"]["net.classifier.weight"].shape == (8000, 512)
    down1 = torch.load(os.path.join(root, "server", "1-server-client-0.ckpt"), weights_only=False)
    down2 = torch.load(os.path.join(root, "server", "2-server-client-0.ckpt"), weights_only=False)
    assert set(down1) == {"integrated_model_params"} and set(down2) == {"incremental_model_params"}


def test_fedstil_checkpoint_schema(tmp_path):
    common, cfg, _ = _run(tmp_path, "fedstil")
    root = os.path.join(common["checkpoints_dir"], cfg["exp_name"])
    m = torch.load(os.path.join(root, "client-1", "fedstil_model.ckpt"), weights_only=False)
    assert set(m) == {"global_weight", "global_weight_atten", "adaptive_weights", "adaptive_bias", "bn_params",
                      "pre_trained_params"}
    assert "classifier.global_weight" in m["global_weight"] and "base.conv1.weight" in m["pre_trained_params"]
    assert m["global_weight_atten"]["base.layer4.0.conv1.global_weight_atten"].shape == (3,)      # kernel width
    g, a = m["global_weight"]["classifier.global_weight"], m["adaptive_weights"]["classifier.adaptive_weight"]
    assert g.shape == a.shape == (8000, 512)
    ex = torch.load(os.path.join(root, "client-1", "fedstil_model_examplars.ckpt"), weights_only=False)
    pid, protos = next(iter(ex.items()))
    assert protos[0][0].shape == (256, 2, 1) and isinstance(protos[0][1], int)
    up = torch.load(os.path.join(root, "client-1", "2-client-1-server.ckpt"), weights_only=False)
    assert set(up) == {"train_cnt", "task_token", "incremental_sw", "incremental_bn"}
    assert up["task_token"].numel() == 256 * 2 * 1
    toks = torch.load(os.path.join(root, "server", "server_tokens.ckpt"), weights_only=False)
    assert set(toks) == {"client-0", "client-1"} and len(toks["client-0"]) == 2


def test_fedavg_aggregation_uses_stale_uploads(tmp_path):
    """online_clients < K: the mean runs over every registered client's LAST upload (fedavg.py:386-397)."""
    from flpr_b200.methods.fedavg import Server
    common = tiny_common(str(tmp_path))
    cfg = tiny_experiment(common, "fedavg", n_clients=3)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        store, comm, server, clients, names = stage.build(cfg)
        n = clients[0].upload_numel()
        vals = {0: 1.0, 1: 2.0, 2: 4.0}
        for c in clients:
            server.register_client(c.client_name)
        for c in clients[:2]:
            c.model.arena.master.fill_(vals[c.client_id]); c.train_cnt = 10
            server.set_client_incremental_state(c.client_name, c.get_incremental_state())
        server.calculate()
        assert torch.allclose(comm.rank_view("glob"), torch.full((n,), 1.5))
        c = clients[2]; c.model.arena.master.fill_(4.0); c.train_cnt = 20
        server.set_client_incremental_state(c.client_name, c.get_incremental_state())
        server.calculate()                                  # clients 0/1 did not upload again: stale states count
        assert torch.allclose(comm.rank_view("glob"), torch.full((n,), (10 * 1 + 10 * 2 + 20 * 4) / 40))
        store.close()


def test_virtual_container():
    vc = VirtualContainer(["cuda:0", "cuda:1"], parallel=2)
    assert vc.max_worker() == 4
    with vc.possess_device() as d0:
        with vc.possess_device(vc.max_worker()) as d1:      # clamped: never negative, never None
            assert d0 is not None and d1 is not None and d0 != d1
    assert vc.devices == vc.capacity


def test_gloo_world2_fedavg_plumbing(tmp_path):
    """BASELINE.json config 1: fedavg, 2 clients, 1 task, world_size=2 on CPU."""
    script = os.path.join(os.path.dirname(__file__), "dist_e2e_check.py")
    env = dict(os.environ, FLPR_TMP=str(tmp_path), OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", "2", script, "fedavg"], env=env, capture_output=True, text=True, timeout=600)
    assert "DIST_E2E OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("world", [2, 3])
def test_gallery_sharded_evaluation_matches_full_gallery(world):
    """gallery sharded over ranks, exact CMC / mAP from all-reduced rank counts (``ops/rank.py``)."""
    script = os.path.join(os.path.dirname(__file__), "dist_eval_check.py")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", str(world), script], env=env, capture_output=True, text=True, timeout=600)
    assert "DIST_EVAL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("method,world", [("fedavg", 2), ("fedweit", 3)])
def test_sharded_validation_logs_the_same_metrics(tmp_path, method, world):
    """``engine_opts.sharded_validation``: all ranks rank every client's gallery in ``1 / world`` slices
    (``evaluation/sharded.py``); CMC / mAP of every client, round and task equal the per-rank validation's (fedweit: the
    per-task checkpoint swap around ``validate`` happens on the owner only)."""
    script = os.path.join(os.path.dirname(__file__), "dist_sharded_val_check.py")
    env = dict(os.environ, FLPR_TMP=str(tmp_path), OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", str(world), script, method], env=env, capture_output=True, text=True,
                       timeout=900)
    assert "DIST_SHARDED_VAL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_comm_soak_script_in_plumbing_mode():
    """``scripts/comm_soak.py`` (randomised arrival / participation / sizes / channels; meant for >= 2 GPUs) stays
    runnable: three gloo ranks, every result checked against the host-side reference."""
    script = os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts", "comm_soak.py")
    env = dict(os.environ, FLPR_FORCE_CPU="1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
                        "--nproc-per-node", "3", script, "40"], env=env, capture_output=True, text=True, timeout=600)
    assert "COMM_SOAK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cli_synthetic(tmp_path):
    import yaml
    common = tiny_common(str(tmp_path))
    common["defaults"]["exp_opts"]["comm_rounds"] = 1
    cp = tmp_path / "common.yaml"
    yaml.safe_dump(common, open(cp, "w"))
    exp = {"exp_name": "cli", "exp_method": "fedavg", "server": {"server_name": "server"},
           "clients": [{"client_name": "client-0", "tasks": ["task-0-0"]}, {"client_name": "client-1", "tasks": ["task-1-0"]}]}
    ep = tmp_path / "exp.yaml"
    yaml.safe_dump(exp, open(ep, "w"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "main.py"), "--common", str(cp), "--experiments", str(ep),
                        "--synthetic"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    assert any(f.startswith("cli-") for f in os.listdir(tmp_path / "logs"))


def test_checkpoint_interval_skips_rounds(tmp_path):
    """engine_opts.checkpoint_interval = 2: payload files only for even rounds; model files still exist at the end."""
    import glob
    from flpr_b200.runtime.experiment import ExperimentStage
    common = tiny_common(str(tmp_path))
    common["defaults"]["exp_opts"].update(comm_rounds=4, val_interval=100)
    cfg = tiny_experiment(common, "fedavg")
    cfg["engine_opts"].update(checkpoint_interval=2, val_at_round0=False)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        stage.run_experiment(cfg)
    root = os.path.join(common["checkpoints_dir"], cfg["exp_name"])
    rounds = sorted({int(os.path.basename(f).split("-")[0]) for f in glob.glob(os.path.join(root, "*", "[0-9]*-*.ckpt"))})
    assert rounds == [2, 4], rounds
    assert os.path.exists(os.path.join(root, "client-0", "fedavg_model.ckpt"))


@pytest.mark.parametrize("method", ["fedstil", "fedavg", "ewc"])
def test_client_threads_on_cpu(tmp_path, method):
    """`parallel: 3` client threads (forced on CPU): the client path shares no mutable state across clients."""
    from flpr_b200.runtime.experiment import ExperimentStage
    common = tiny_common(str(tmp_path))
    common["parallel"] = 3
    common["defaults"]["exp_opts"].update(comm_rounds=3, online_clients=4, val_interval=100)
    cfg = tiny_experiment(common, method, n_clients=4)
    cfg["engine_opts"].update(client_threads="force", val_at_round0=False)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        log = stage.run_experiment(cfg)
    data = log.records["data"]
    assert len(data) == 4
    for client in data.values():
        assert len(client) == 3
        for tasks in client.values():
            for vals in tasks.values():
                assert all(v == v and 0.0 <= v < 1e4 for v in vals.values()), vals


def test_fedweit_per_task_evaluation_without_checkpoint_files(tmp_path):
    """With checkpoint files disabled (or muted by ``checkpoint_interval``) FedWeIT keeps a device snapshot per finished
    task, so an older task is still evaluated with the weights it had at the end of its own last round."""
    common, cfg, log = _run(tmp_path, "fedweit", engine_opts={"checkpoints": False})
    data = log.records["data"]["client-0"]
    assert {"val_map", "val_rank_1"} <= set(data["2"]["task-0-0"])          # older task evaluated in round 2
    root = os.path.join(common["checkpoints_dir"], cfg["exp_name"])
    assert not os.path.exists(os.path.join(root, "client-0", "task-0-0.ckpt"))
    # same numbers as the run that reads the per-task checkpoint files, and with the files muted on odd rounds
    _, _, with_files = _run(tmp_path / "files", "fedweit")
    _, _, muted = _run(tmp_path / "muted", "fedweit", engine_opts={"checkpoint_interval": 2})
    for other in (with_files, muted):
        for client in ("client-0", "client-1"):
            for task, metrics in other.records["data"][client]["2"].items():
                for k, v in metrics.items():
                    assert abs(log.records["data"][client]["2"][task][k] - v) < 1e-6, (client, task, k)
