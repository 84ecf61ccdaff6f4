"""Multi-rank (gloo) check of ``engine_opts.sharded_validation``: the same experiment run with per-rank validation and
with the collective gallery-sharded validation must log identical CMC / mAP for every client, round and task (3 clients
on 2 ranks: uneven hosting, so one rank participates in collectives it does not own)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from flpr_b200.runtime.experiment import ExperimentStage  # noqa: E402
from helpers import tiny_common, tiny_experiment, tiny_factory  # noqa: E402


def run(tag: str, sharded: bool, method: str):
    tmp = os.path.join(os.environ.get("FLPR_TMP", "/tmp/flpr_dist"), tag, f"r{os.environ.get('RANK', '0')}")
    common = tiny_common(tmp)
    common["defaults"]["exp_opts"].update(comm_rounds=2, val_interval=1)
    cfg = tiny_experiment(common, method, n_clients=3, n_tasks=2)
    cfg["engine_opts"].update(sharded_validation=sharded, val_at_round0=True)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        log = stage.run_experiment(cfg)
        calls = getattr(getattr(stage, "_ranker_obj", None), "calls", 0)
    return log.records.get("data", {}), calls, stage.rank


def main():
    import torch.distributed as dist
    method = sys.argv[1] if len(sys.argv) > 1 else "fedavg"
    dist.init_process_group("gloo")               # ONE group for both runs (the stages then neither create nor destroy
                                                  # it: re-initialising on the same store races on the old mesh keys)
    plain, c0, rank = run("plain", False, method)
    shard, c1, _ = run("sharded", True, method)
    ok = c0 == 0 and c1 == 3 * 2 * 3            # 3 clients x 2 tasks x (round 0 + 2 rounds), on every rank
    bad = []
    if rank == 0:                                 # rank 0 owns the gathered log
        vals = 0
        for client, rounds in plain.items():
            for rnd, tasks in rounds.items():
                for task, rec in tasks.items():
                    for k, v in rec.items():
                        if k.startswith("val_"):
                            vals += 1
                            w = shard.get(client, {}).get(rnd, {}).get(task, {}).get(k)
                            if w is None or abs(w - v) > 1e-6:
                                bad.append((client, rnd, task, k, v, w))
        ok = ok and not bad and vals >= 3 * 3 * 2 * 5 and len(plain) == 3
        print("DIST_SHARDED_VAL", "OK" if ok else "FAILED", f"metrics={vals} collectives={c1}", json.dumps(bad[:5]),
              flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
