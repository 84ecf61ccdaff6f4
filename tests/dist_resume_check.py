"""Rank loss under ``torchrun --max-restarts``: a rank exits hard in the middle of an experiment
(``FLPR_FAULT_EXIT=rank:round:phase``), the elastic agent restarts the group, every rank resumes from the newest
manifest that is committed on EVERY rank (``runtime/resume.py``) and the experiment finishes. Each rank then dumps the
final server / client weights found in its last manifest; the test compares them with an uninterrupted run.

    FLPR_TMP=/tmp/x [FLPR_FAULT_EXIT=1:3:round] torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 \
        [--max-restarts 3] tests/dist_resume_check.py fedavg 4

``tests/test_resume.py`` performs the restart itself (a second ``torchrun`` on the same directories): re-forming a gloo
group inside one elastic agent proved flaky on 127.0.0.1 in the build sandbox (``connectFullMesh: connection refused``
on some restarts - the agent then simply restarts again, which is why ``--max-restarts`` should be > 1 in production).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from flpr_b200.runtime import resume as R  # noqa: E402
from flpr_b200.runtime.experiment import ExperimentStage  # noqa: E402
from helpers import tiny_common, tiny_experiment, tiny_factory  # noqa: E402


def main():
    method, rounds = sys.argv[1], int(sys.argv[2])
    rank = int(os.environ.get("RANK", "0"))
    attempt = int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))
    base = os.environ["FLPR_TMP"]
    tmp = os.path.join(base, "shared" if os.environ.get("FLPR_SHARED_ROOT") else f"r{rank}")
    common = tiny_common(tmp)
    common["defaults"]["exp_opts"].update(comm_rounds=rounds, val_interval=100)
    common["defaults"]["task_opts"]["sustain_rounds"] = 2
    cfg = tiny_experiment(common, method, n_clients=2, n_tasks=2)
    cfg["engine_opts"].update(resume=True, resume_interval=1, val_at_round0=False)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        stage.run_experiment(cfg)
        from flpr_b200.runtime.checkpoint import CheckpointStore
        store = CheckpointStore(os.path.join(common["checkpoints_dir"], cfg["exp_name"]), asynchronous=False)
        committed = R._committed(store, rank)
        assert rounds in committed, (rank, committed)
        st = torch.load(store.path(R.ACTOR, R._gen_name(rank, committed[rounds])), weights_only=False)
        out = {"attempt": attempt, "server": st["server"]["model"],
               "clients": {n: c["model"] for n, c in st["clients"].items()},
               "cnt": {n: c["train_cnt"] for n, c in st["clients"].items()}}
        torch.save(out, os.path.join(base, f"result_rank{rank}.pt"))
    print(f"DIST_RESUME rank {rank} done on attempt {attempt}", flush=True)


if __name__ == "__main__":
    main()
