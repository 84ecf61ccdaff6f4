"""Run warm-up FedSTIL bench rounds, then ONE round between cudaProfilerStart/Stop so that
``ncu --profile-from-start off --metrics gpu__time_duration.sum`` lists every launch of a steady-state round."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    a = bench.parse_args()
    from flpr_b200.data.synthetic import random_array_split
    from flpr_b200.runtime.config import merge_experiment
    from flpr_b200.runtime.experiment import ExperimentStage
    from flpr_b200.runtime.explog import ExperimentLog
    from flpr_b200.utils.misc import DeviceTimer
    common, exp = bench.build_config(a, "flpr", 1)
    cfg = merge_experiment(common, exp)

    def factory(task, split):
        cid, tid = int(task.split("-")[1]), int(task.split("-")[2])
        n = a.images if split == "train" else 64
        return random_array_split(n, a.ids, (a.height, a.width), id_offset=(cid * 5 + tid) * a.ids % (8000 - a.ids), seed=cid)

    with ExperimentStage(common, [cfg], source_factory=factory) as stage:
        store, comm, server, clients, names = stage.build(cfg)
        log = ExperimentLog("/tmp/x.json", enabled=False)
        timer = DeviceTimer(stage.device)
        r = 0
        for _ in range(a.warmup):
            r += 1
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
            torch.cuda.synchronize(); store.flush()
            bench.cleanup_payloads(common["checkpoints_dir"])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        for _ in range(a.steps):
            r += 1
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        store.flush(); store.close(); comm.close()


if __name__ == "__main__":
    main()
