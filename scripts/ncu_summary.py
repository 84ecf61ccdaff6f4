"""Summarise an ncu report (``--set full``) into a markdown table under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/x.md "title" [flops-per-launch-json]

Reads the raw page with ``ncu -i ... --page raw --csv`` (works without a GPU) and prints, per profiled launch: duration,
DRAM bytes, DRAM %, tensor-pipe %, achieved occupancy, registers, smem – next to the measured peaks of
MEASURED_PEAKS.json (roofline denominators).
"""
import csv, io, json, os, subprocess, sys

COLS = [("gpu__time_duration.sum", "dur_us", lambda v, u: v / 1e3 if u in ("ns", "nsecond") else v),
        ("dram__bytes_read.sum", "dram_rd_MB", None), ("dram__bytes_write.sum", "dram_wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", None),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", None),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_pipe_%", None),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", None),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", None),
        ("launch__registers_per_thread", "regs", None),
        ("launch__shared_mem_per_block_dynamic", "smem_KB", None),
        ("launch__waves_per_multiprocessor", "waves", None)]


def to_mb(v, unit):
    u = unit.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    lines = [f"# {title}", "", f"source: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on), "
             f"read with `ncu -i … --page raw --csv`.", ""]
    if peaks:
        lines += [f"Roofline denominators (MEASURED_PEAKS.json): HBM copy {peaks.get('hbm_gbs')} GB/s, cuBLAS bf16 "
                  f"{peaks.get('bf16_tflops')} TFLOP/s burst / {peaks.get('bf16_tflops_sustained')} sustained.", ""]
    names = [c[1] for c in COLS]
    lines.append("| # | kernel | grid | block | " + " | ".join(names) + " |")
    lines.append("|" + "---|" * (4 + len(names)))
    for n, r in enumerate(body):
        vals = []
        for key, label, fn in COLS:
            if key not in ix:
                vals.append("-")
                continue
            try:
                v = float(r[ix[key]].replace(",", ""))
            except ValueError:
                vals.append(r[ix[key]])
                continue
            u = units[ix[key]]
            if label.endswith("_MB"):
                v = to_mb(v, u)
            elif label == "dur_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            elif label == "smem_KB":
                v = v / 1e3 if u.lower() == "byte" else v
            vals.append(f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}")
        k = r[ix["Kernel Name"]]
        k = k.replace("flpr::", "").split("(")[0][:70]
        lines.append(f"| {n} | `{k}` | {r[ix['Grid Size']]} | {r[ix['Block Size']]} | " + " | ".join(vals) + " |")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
