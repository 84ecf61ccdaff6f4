"""Section-level timing of one FedSTIL bench round on one GPU (host wall clock with device syncs around sections)."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

acc = collections.defaultdict(float); cnt = collections.defaultdict(int)

def timed(obj, name, label):
    fn = getattr(obj, name)
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[label] += (time.perf_counter() - t0) * 1e3; cnt[label] += 1
        return out
    setattr(obj, name, wrap)

def main():
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    a = bench.parse_args()
    from flpr_b200.data.synthetic import random_array_split
    from flpr_b200.runtime.config import merge_experiment
    from flpr_b200.runtime.experiment import ExperimentStage
    from flpr_b200.runtime.explog import ExperimentLog
    from flpr_b200.utils.misc import DeviceTimer
    from flpr_b200.methods import fedstil
    from flpr_b200.runtime import modules, checkpoint
    common, exp = bench.build_config(a, "flpr", 1)
    cfg = merge_experiment(common, exp)
    def factory(task, split):
        cid, tid = int(task.split("-")[1]), int(task.split("-")[2])
        n = a.images if split == "train" else 64
        return random_array_split(n, a.ids, (a.height, a.width), id_offset=(cid * 5 + tid) * a.ids % (8000 - a.ids), seed=cid)
    timed(fedstil.Operator, "generate_prototypes", "proto_pass")
    timed(fedstil.Operator, "invoke_train", "epoch_total")
    timed(fedstil.Model, "build_examplars", "herding")
    timed(fedstil.Client, "save_model", "save_model")
    timed(fedstil.Client, "get_incremental_state", "upload_copy")
    timed(fedstil.Server, "prepare_dispatch", "mix")
    timed(fedstil.Server, "calculate", "calculate")
    timed(checkpoint.CheckpointStore, "save", "ckpt_save_call")
    timed(fedstil.Model, "model_state", "model_state")
    timed(fedstil.Model, "examplars_state", "examplars_state")
    with ExperimentStage(common, [cfg], source_factory=factory) as stage:
        store, comm, server, clients, names = stage.build(cfg)
        log = ExperimentLog("/tmp/x.json", enabled=False); timer = DeviceTimer(stage.device)
        for r in range(1, 4):
            acc.clear(); cnt.clear()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
            torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
            store.flush()
            print(f"--- round {r}: {tot:.1f} ms total")
            for k in sorted(acc, key=lambda k: -acc[k]):
                print(f"   {k:18s} {acc[k]:9.1f} ms  x{cnt[k]}")
        store.close(); comm.close()

if __name__ == "__main__":
    main()
