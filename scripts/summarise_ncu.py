#!/usr/bin/env python
"""Turns the ncu outputs a GPU call left in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarise_ncu.py <tag>      e.g. r01b

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list) and every
gpurun_out/full_*.ncu-rep (ncu --set full), writes profiles/<tag>_launch_summary.txt,
profiles/<tag>_launches.csv and profiles/<tag>_ncu_full.txt.
"""
import collections
import csv
import glob
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
]


def launch_summary(tag):
    src = os.path.join(OUT, "launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    d = collections.OrderedDict()
    for r in rows:
        n = r["Kernel Name"]
        n = n.replace("pfx::<unnamed>::", "").replace("pfx::", "")
        n = n.split("(")[0][:90]
        v = d.setdefault(n, [0, 0.0])
        v[0] += 1
        v[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in d.values())
    with open(os.path.join(PROF, f"{tag}_launch_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {len(rows)} launches, "
                f"{tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)\n")
        f.write(f"{'us total':>12} {'launches':>8} {'share':>7}  kernel\n")
        for n, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:12.1f} {v[0]:8d} {100 * v[1] / tot:6.1f}%  {n}\n")
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches.csv"))


def full_summary(tag):
    reps = sorted(glob.glob(os.path.join(OUT, "full_*.ncu-rep")))
    if not reps:
        return
    with open(os.path.join(PROF, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; one block per captured launch\n")
        for rep in reps:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(raw)))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                name = r[hdr.index("Kernel Name")]
                f.write(f"\n== {os.path.basename(rep)} :: {name[:100]}\n")
                for m in METRICS:
                    if m in hdr:
                        i = hdr.index(m)
                        f.write(f"   {m:75s} {r[i]:>16s} {units[i]}\n")


def traffic_summary(tag, config="C2"):
    """gpurun_out/traffic.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
    over the pass / reduce / fit launches of one step) -> profiles/traffic.json, the per-launch DRAM bytes
    bench.py reports as roofline.traffic."""
    import json
    src = os.path.join(OUT, "traffic.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    acc = collections.defaultdict(lambda: collections.defaultdict(float))
    n = collections.Counter()
    for r in rows:
        name = r["Kernel Name"]
        cls = next((c for c, pat in (("pass1", "pass1"), ("pass2", "pass2"), ("reduce", "reduce_kernel"),
                                     ("fit", "fit_cell")) if pat in name), None)
        if cls is None:
            continue
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(r["Metric Unit"], 1)
        acc[cls][r["Metric Name"]] += float(r["Metric Value"]) * mult
        n[cls] += r["Metric Name"] == "gpu__time_duration.sum"
    path = os.path.join(PROF, "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[config] = {c: (acc[c]["dram__bytes_read.sum"] + acc[c]["dram__bytes_write.sum"]) / n[c] for c in acc}
    data[config + "_detail"] = {c: {"launches": n[c], "read_bytes_per_launch": acc[c]["dram__bytes_read.sum"] / n[c],
                                    "write_bytes_per_launch": acc[c]["dram__bytes_write.sum"] / n[c],
                                    "avg_launch_us_under_ncu": acc[c]["gpu__time_duration.sum"] / n[c],
                                    "capture": tag} for c in acc}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    shutil.copy(src, os.path.join(PROF, f"{tag}_traffic.csv"))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launch_summary(tag)
    full_summary(tag)
    traffic_summary(tag)
    for name in ("bench.log", "bench_ref.log", "pytest_gpu.log", "smoke.log", "bench_C1.log", "bench_C3.log",
                 "bench_C4.log", "bench_C5.log"):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            with open(p) as src, open(os.path.join(PROF, f"{tag}_{name}"), "w") as dst:
                dst.write("".join(src.readlines()[-12:]))
