"""2-rank CPU (gloo) end-to-end run: every rank hosts one client, the server role is replicated; after the run the
aggregated global parameters must be bit-identical on both ranks and equal to the single-process result."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from flpr_b200.runtime.experiment import ExperimentStage  # noqa: E402
from helpers import tiny_common, tiny_experiment, tiny_factory  # noqa: E402


def main():
    method = sys.argv[1] if len(sys.argv) > 1 else "fedavg"
    tmp = os.path.join(os.environ.get("FLPR_TMP", "/tmp/flpr_dist"), f"r{os.environ.get('RANK', '0')}")
    common = tiny_common(tmp)
    cfg = tiny_experiment(common, method, n_clients=2, n_tasks=1)
    with ExperimentStage(common, [cfg], source_factory=tiny_factory()) as stage:
        store, comm, server, clients, names = stage.build(cfg)
        assert len(clients) == 1 and comm.world == 2 and comm.mode == "gloo"
        from flpr_b200.runtime.explog import ExperimentLog
        from flpr_b200.utils.misc import DeviceTimer, same_seeds
        same_seeds(cfg["random_seed"])
        log = ExperimentLog(os.path.join(tmp, "log.json"), enabled=False)
        timer = DeviceTimer(stage.device)
        for r in (1, 2):
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
        glob = comm.rank_view("glob").clone()
        gathered = [torch.empty_like(glob) for _ in range(2)]
        dist.all_gather(gathered, glob)
        same = torch.equal(gathered[0], gathered[1])
        finite = bool(torch.isfinite(glob).all())
        moved = (glob - clients[0].model.arena.master[:glob.numel()]).abs().max().item()
        store.close()
    if stage.rank == 0:
        print("DIST_E2E", "OK" if (same and finite) else "FAILED", f"identical={same} finite={finite} drift={moved:.3e}",
              flush=True)


if __name__ == "__main__":
    main()
