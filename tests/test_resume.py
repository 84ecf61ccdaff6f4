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
               "pipe": [c.task_pipeline.state_dict() for c in clients],
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
