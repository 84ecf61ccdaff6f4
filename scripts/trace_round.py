"""Timeline of steady-state FedSTIL bench rounds with torch.profiler (CUPTI activity tracing, no kernel replay):
which kernels / copies run when, how busy the GPU is, where the host sits in a wait. Writes a chrome trace and prints
an interval summary (union busy time per category, idle gaps, top kernels)."""
import gzip, json, os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def union(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def main():
    a = bench.parse_args()
    from torch.profiler import profile, ProfilerActivity
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
        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
            for _ in range(a.steps):
                r += 1
                stage._process_one_round(r, server, clients, names, cfg, log, timer, comm)
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        store.flush(); store.close(); comm.close()
    tag = f"p{a.parallel}"
    path = os.path.join(ROOT, "gpurun_out", f"trace_round_{tag}.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    with gzip.open(path + ".gz", "wt") as f:
        json.dump({"traceEvents": [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]}, f)
    os.remove(path)
    kern = [(e["ts"], e["ts"] + e["dur"], e["name"], e["args"].get("stream")) for e in ev if e.get("cat") == "kernel"]
    cpy = [(e["ts"], e["ts"] + e["dur"], e["name"], e["args"].get("bytes", 0)) for e in ev if e.get("cat") == "gpu_memcpy"]
    t_lo = min(k[0] for k in kern); t_hi = max(k[1] for k in kern)
    print(f"[{tag}] wall {wall:.1f} ms for {a.steps} round(s) (profiler on); GPU span {(t_hi - t_lo) / 1e3:.1f} ms")
    print(f"kernels: {len(kern)}  sum {sum(k[1]-k[0] for k in kern)/1e3:.1f} ms  union-busy {union([(k[0], k[1]) for k in kern])/1e3:.1f} ms")
    by_kind = collections.defaultdict(lambda: [0, 0.0, 0])
    for s, e, n, b in cpy:
        by_kind[n][0] += 1; by_kind[n][1] += (e - s) / 1e3; by_kind[n][2] += b
    for n, (c, ms, b) in by_kind.items():
        print(f"  {n:28s} x{c:5d}  {ms:8.1f} ms  {b/1e6:9.1f} MB  {b/1e6/max(ms,1e-9):7.1f} GB/s")
    print(f"copies union-busy {union([(c[0], c[1]) for c in cpy])/1e3:.1f} ms; kernels+copies union {union([(k[0], k[1]) for k in kern] + [(c[0], c[1]) for c in cpy])/1e3:.1f} ms")
    streams = collections.defaultdict(float)
    for s, e, n, st in kern:
        streams[st] += (e - s) / 1e3
    print("kernel ms per stream:", {k: round(v, 1) for k, v in sorted(streams.items(), key=lambda kv: -kv[1])[:12]})
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n, st in kern:
        agg[n[:80]][0] += 1; agg[n[:80]][1] += (e - s) / 1e3
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  {ms:8.2f} ms x{c:5d}  {n}")
    # idle gaps of the whole device (all streams): where the GPU waits for the host
    allk = sorted([(k[0], k[1], k[2]) for k in kern] + [(c[0], c[1], c[2]) for c in cpy])
    gaps, cur_end, last_name = [], allk[0][1], allk[0][2]
    for s_, e_, n_ in allk[1:]:
        if s_ > cur_end + 200:                       # > 0.2 ms idle
            gaps.append((s_ - cur_end, cur_end - t_lo, last_name[:50], n_[:50]))
        if e_ > cur_end:
            cur_end, last_name = e_, n_
    print(f"idle gaps > 0.2 ms: {len(gaps)}, total {sum(g[0] for g in gaps) / 1e3:.1f} ms")
    for g in sorted(gaps, reverse=True)[:40]:
        print(f"  gap {g[0] / 1e3:7.2f} ms at t={g[1] / 1e3:9.2f}  after [{g[2]}]  before [{g[3]}]")
    # timeline of the big copies and of the comm kernels (where does the copy engine sit relative to fed_* ?)
    big = sorted([(e["ts"], e["dur"], e["name"], e["args"].get("bytes", 0), e["args"].get("stream")) for e in ev
                  if e.get("cat") == "gpu_memcpy" and e["args"].get("bytes", 0) >= (8 << 20)])
    fed = sorted([(e["ts"], e["dur"], e["name"][:40], 0, e["args"].get("stream")) for e in ev
                  if e.get("cat") == "kernel" and ("fed_" in e["name"] or "herding" in e["name"])])
    print("timeline (ms since first kernel): big copies >= 8 MB and fed_* / herding kernels")
    for ts, dur, name, b, st in sorted(big + fed)[-120:]:
        print(f"  t={(ts - t_lo) / 1e3:9.2f}  dur={dur / 1e3:8.3f}  {b / 1e6:8.1f} MB  stream {st}  {name}")
    # host-side: time inside synchronising calls
    sync = [(e["name"], e["dur"]) for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver") and
            ("Synchronize" in e["name"] or "cudaMemcpy" == e["name"])]
    sagg = collections.defaultdict(lambda: [0, 0.0])
    for n, d in sync:
        sagg[n][0] += 1; sagg[n][1] += d / 1e3
    print("host time in sync calls:", {k: (v[0], round(v[1], 1)) for k, v in sagg.items()})


if __name__ == "__main__":
    main()
