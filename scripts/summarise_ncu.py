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


def full_summary_from_csv(tag):
    """Raw pages exported on the GPU box (scripts/gpu_r2g.sh: gpurun_out/<tag>_full_*_raw.csv) ->
    profiles/<tag>_ncu_full.txt; the stall-reason totals of the source pages are appended per kernel."""
    raws = sorted(glob.glob(os.path.join(OUT, f"{tag}_full_*_raw.csv")))
    if not raws:
        return
    extra = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
             "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    with open(os.path.join(PROF, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one launch each; C3 = 4096^2 unless noted);\n"
                "# raw page metrics, then the warp-stall sampling totals of the source page\n")
        for raw in raws:
            rows = list(csv.reader(open(raw)))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                name = r[hdr.index("Kernel Name")]
                f.write(f"\n== {os.path.basename(raw)} :: {name[:110]}\n")
                for m in METRICS + extra:
                    if m in hdr:
                        i = hdr.index(m)
                        f.write(f"   {m:85s} {r[i]:>18s} {units[i]}\n")
            src = raw.replace("_raw.csv", "_source.csv")
            if os.path.exists(src):
                srows = list(csv.reader(open(src)))
                h = srows[1]
                idx = {c: i for i, c in enumerate(h)}
                data = [x for x in srows[2:] if len(x) == len(h)]
                stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
                tot = {c: sum(int(x[idx[c]] or 0) for x in data) for c in stalls}
                T = sum(tot.values()) or 1
                f.write("   warp stall samples: " + "  ".join(f"{c[6:]}={100 * v / T:.1f}%" for c, v in
                                                              sorted(tot.items(), key=lambda kv: -kv[1]) if 100 * v / T >= 0.5) + "\n")
                ops = collections.Counter()
                for x in data:
                    parts = x[idx["Source"]].split()
                    op = next((q for q in parts if not q.startswith("@")), "?").split(".")[0]
                    ops[op] += int(x[idx["Instructions Executed"]] or 0)
                n = sum(ops.values()) or 1
                f.write("   instruction mix:    " + "  ".join(f"{o}={100 * c / n:.1f}%" for o, c in ops.most_common(12)) + "\n")
        sass = os.path.join(OUT, f"{tag}_sass_excerpt.txt")
        if os.path.exists(sass):
            f.write("\n" + open(sass).read())


def traffic_summary(tag, config="C2", src=None):
    """ncu --metrics gpu__time_duration.sum,dram__bytes_*,sm__pipe_fp64_cycles_active,...op_{dfma,dadd,dmul}...
    over the pass / reduce / fit launches of one step -> profiles/traffic.json (per-launch DRAM bytes and FP64
    pipe share that bench.py reports as roofline.traffic / fp64_pipe_pct) and profiles/fit_flops.json (FP64
    flops of the fit kernel: 2*dfma + dadd + dmul thread instructions)."""
    import json
    src = src or os.path.join(OUT, "traffic.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    acc = collections.defaultdict(lambda: collections.defaultdict(float))
    n = collections.Counter()
    fp = "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"
    launches = collections.OrderedDict()  # launch ID -> (class, {metric: value})
    for r in rows:
        name = r["Kernel Name"]
        cls = next((c for c, pat in (("pass1", "pass1"), ("pass2", "pass2"), ("reduce", "reduce_kernel"),
                                     ("fit", "fit_")) if pat in name), None)
        if cls is None:
            continue
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "ms": 1e3,
                "s": 1e6}.get(r["Metric Unit"], 1)
        launches.setdefault(r["ID"], (cls, {}))[1][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * mult
    for cls, m in launches.values():
        t = m.get("gpu__time_duration.sum", 0.0)
        if cls == "fit" and t < 20.0:  # a form of the kernel that returned at once (the fit launches two, a device flag picks one)
            continue
        n[cls] += 1
        for k, v in m.items():
            acc[cls][k] += v * t if k == fp else v  # the pipe share is averaged over time, everything else summed
    for cls in acc:
        if fp in acc[cls]:
            acc[cls][fp] = acc[cls][fp] / acc[cls]["gpu__time_duration.sum"] * n[cls]
    path = os.path.join(PROF, "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[config] = {c: (acc[c]["dram__bytes_read.sum"] + acc[c]["dram__bytes_write.sum"]) / n[c] for c in acc if c != "fit"}
    flops = lambda c: (2 * acc[c]["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] +  # noqa: E731
                       acc[c]["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] +
                       acc[c]["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"])
    data[config + "_detail"] = {c: {"launches": n[c], "read_bytes_per_launch": acc[c]["dram__bytes_read.sum"] / n[c],
                                    "write_bytes_per_launch": acc[c]["dram__bytes_write.sum"] / n[c],
                                    "avg_launch_us_under_ncu": acc[c]["gpu__time_duration.sum"] / n[c],
                                    "fp64_pipe_pct": acc[c][fp] / n[c] if fp in acc[c] else None,
                                    "fp64_flops_per_launch": flops(c) / n[c],
                                    "capture": tag} for c in acc}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    shutil.copy(src, os.path.join(PROF, f"{tag}_traffic_{config}.csv"))
    if "fit" in acc and flops("fit") > 0:
        import bench
        cfg = bench.CONFIGS[config]
        cells = (cfg["nx"] - 1) * (cfg["ny"] - 1)
        fpath = os.path.join(PROF, "fit_flops.json")
        fd = json.load(open(fpath)) if os.path.exists(fpath) else {}
        fd["alph" if cfg["alph"] else "noalph"] = {
            "flops_per_cell": flops("fit") / cells, "config": config, "cells": cells, "launches": n["fit"],
            "dfma": acc["fit"]["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"],
            "dadd": acc["fit"]["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"],
            "dmul": acc["fit"]["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"],
            "fp64_pipe_pct": acc["fit"][fp] / n["fit"] if fp in acc["fit"] else None,
            "us_under_ncu": acc["fit"]["gpu__time_duration.sum"], "capture": tag,
            "how": "ncu thread-instruction counts over the fit kernels of one step (lane kernel + finisher): 2*dfma + dadd + dmul"}
        json.dump(fd, open(fpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    sys.path.insert(0, ROOT)
    launch_summary(tag)
    full_summary(tag)
    full_summary_from_csv(tag)
    traffic_summary(tag)
    for cfgname in ("C1", "C2", "C3", "C4", "C5"):
        traffic_summary(tag, cfgname, os.path.join(OUT, f"traffic_{cfgname}.csv"))
    fp = os.path.join(OUT, "fp64_peak.log")
    if os.path.exists(fp):
        import json
        line = [l for l in open(fp) if l.startswith("{")]
        if line:
            json.dump(json.loads(line[-1]), open(os.path.join(PROF, "fp64_peak.json"), "w"), indent=1)
    for name in ("bench.log", "bench_ref.log", "pytest_gpu.log", "smoke.log", "bench_C1.log", "bench_C2.log", "bench_C3.log",
                 "bench_C4.log", "bench_C5.log", "bench_n2.log", "bench_n4.log", "bench_n8.log"):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            with open(p) as src, open(os.path.join(PROF, f"{tag}_{name}"), "w") as dst:
                dst.write("".join(src.readlines()[-12:]))
