"""cProfile of the HOST side of steady-state FedSTIL bench rounds (where does the Python / launch time go)."""
import cProfile, io, os, pstats, sys, time
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
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        for _ in range(a.steps):
            r += 1
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
        torch.cuda.synchronize()
        pr.disable()
        print(f"wall per round: {(time.perf_counter() - t0) * 1e3 / a.steps:.1f} ms")
        store.flush(); store.close(); comm.close()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(70)
    txt = out.getvalue()
    print(txt[:14000])


if __name__ == "__main__":
    main()
