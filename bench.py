#!/usr/bin/env python
"""bench.py -- headline benchmark of the PlateFlex hot path on B200.

Workload (BASELINE.json configs[1]): synthetic 1024x1024 topography + Bouguer pair, 5 km
spacing, ns = 20 scales (the first 20 of the grid's natural wavenumber ladder), 11 azimuths.
One step = one pass CWT -> admittance / coherence with jackknife errors for the pair:
nx*ny*ns*11 = 2.31e8 grid-cell.scale.azimuth units, two grids in, four (nx,ny,ns) maps out.

  value      units/s with the two input grids already resident in HBM (device-pointer C ABI)
  e2e        same metric through the host-pointer C ABI call (pinned host buffers, H2D of the
             two grids and D2H of the four maps inside the timed region)
  roofline   algorithmic bytes (547 B/unit for the two-grid path) / measured time
  fit        per-cell L2 (Te, F) fit of the resulting maps, cells/s (second metric of BASELINE.json)
  cpu_baseline  the CPU oracle port (reference precision, OpenMP over planes) on a bounded sample

`--impl reference` times only the CPU oracle port (the reference's Fortran cannot be built here).
Multi-GPU (torchrun): the scale axis is sharded over ranks, no data-path collective; the timed
region is bracketed by barriers and the maximum over ranks is reported (strong scaling).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 1024
DX = DY = 5.0
NS = 20
NA = 11
BYTES_PER_UNIT = 2 * (4 * 64 + 16) + 4 * 8 / 11.0  # SURVEY.md 8d, two-grid admittance/coherence path
METRIC = "grid-cell*scale*azimuth/s (CWT->admittance/coherence)"


def workload(nx=NX, ny=NY, ns=NS):
    import plateflex_b200 as pf
    from plateflex_b200 import classes, synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    topo, grav = synthetic.make_pair(nx, ny, DX, DY, seed=20240000 + nx, Te=40.0, F=0.5)
    k = classes._lam2k(nx, ny, DX, DY)[1][:ns].copy()
    return topo, grav, k, pf.cf_wt.k0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference(topo, grav, k, k0, sample_scales, threads):
    """The CPU oracle port in the reference's precision (float32, radix-2 fourn) on `sample_scales`."""
    from oracle import oracle as ora
    ora.lib()
    ks = k[list(sample_scales)]
    t0 = time.perf_counter()
    w1 = ora.wlet_transform(topo, DX, DY, ks, k0=k0, prec=32, nthreads=threads)
    w2 = ora.wlet_transform(grav, DX, DY, ks, k0=k0, prec=32, nthreads=threads)
    ora.wlet_admit_coh(w1, w2, prec=32)
    dt = time.perf_counter() - t0
    units = topo.shape[0] * topo.shape[1] * NA * len(ks)
    return units / dt, dt, units


def run_reference(args, rank):
    if rank != 0:
        return
    topo, grav, k, k0 = workload()
    threads = os.cpu_count() or 1
    sample = [0, NS - 1]
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, units = cpu_reference(topo, grav, k, k0, sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    what = f"scales {sample} of {NS} ({units:.3g} units per step), float32 oracle port of cpwt.f90, OpenMP over planes"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic {NX}x{NY} topo+Bouguer, {DX:g} km, ns={NS}, na={NA} (BASELINE configs[1])",
                   "sample": what},
        "cpu_baseline": {"value": value, "unit": "units/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default=os.environ.get("PLATEFLEX_B200_ENGINE", "pruned"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from plateflex_b200 import _lib

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    topo, grav, k, k0 = workload()
    engine = _lib.ENGINE_DENSE if args.engine == "dense" else _lib.ENGINE_PRUNED
    plan = _lib.Plan(NX, NY, DX, DY, k, k0, engine=engine, device=local)
    stream = torch.cuda.current_stream()
    plan.set_stream(stream.cuda_stream)

    # strong scaling: contiguous block of scales per rank
    per = [NS // world + (1 if r < NS % world else 0) for r in range(world)]
    is0 = sum(per[:rank])
    is1 = is0 + per[rank]
    nsl = is1 - is0
    nxy = NX * NY

    d_topo = torch.from_numpy(topo).to(dev)
    d_grav = torch.from_numpy(grav).to(dev)
    outs = [torch.empty(max(nsl, 1) * nxy, dtype=torch.float64, device=dev) for _ in range(4)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def step_dev():
        if nsl > 0:
            _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, d_topo.data_ptr(), d_grav.data_ptr(), is0, is1,
                                                       *[o.data_ptr() for o in outs]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = plan.launch_count()
    times = []
    with ClockSampler(local) as clocks:
        barrier()
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            e0.record(stream)
            step_dev()
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        barrier()
    launches = plan.launch_count() - launches0
    t = torch.tensor(times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, per step
    ms_per_step = float(t.mean().item())
    units = nxy * NS * NA
    value = units / (ms_per_step * 1e-3)

    # ---- e2e: host-pointer C ABI with pinned buffers (H2D + D2H inside the timed region) ----
    h_in = [torch.from_numpy(a).pin_memory() for a in (topo, grav)]
    h_out = [torch.empty(max(nsl, 1) * nxy, dtype=torch.float64).pin_memory() for _ in range(4)]

    def step_host():
        if nsl > 0:
            _lib.check(_lib.lib.pfx_wlet_admit_coh_host(plan.handle, h_in[0].data_ptr(), h_in[1].data_ptr(), is0, is1,
                                                        *[o.data_ptr() for o in h_out]))

    for _ in range(2):
        step_host()
    barrier()
    te = []
    for _ in range(max(3, args.steps // 2)):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        step_host()  # synchronous: returns after the D2H copies completed
        te.append((time.perf_counter() - t0) * 1e3)
    t2 = torch.tensor(te, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.mean().item())
    e2e = {"value": units / (e2e_ms * 1e-3), "unit": "units/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 2 * nxy * 8, "d2h_bytes_per_step": 4 * nxy * nsl * 8 * (world if world > 1 else 1)}

    # ---- second metric: per-cell L2 fit over the maps this rank holds (full-ns maps only at N=1) ----
    fit = None
    if not args.no_fit and world == 1:
        from plateflex_b200.flex import conf_flex
        conf = conf_flex.as_struct()
        d_k = torch.from_numpy(k).to(dev)
        d_wd = torch.zeros(nxy, dtype=torch.float64, device=dev)
        fo = [torch.empty(nxy, dtype=torch.float64, device=dev) for _ in range(5)]
        st = torch.empty(nxy, dtype=torch.int32, device=dev)

        def step_fit():
            _lib.check(_lib.lib.pfx_l2_fit_dev(local, stream.cuda_stream, C.byref(conf), d_k.data_ptr(), NS, NX, NY,
                                               *[o.data_ptr() for o in outs], d_wd.data_ptr(), None, None, None, 1, 0,
                                               _lib.ATYPES["joint"], fo[0].data_ptr(), fo[1].data_ptr(),
                                               fo[2].data_ptr(), fo[3].data_ptr(), None, None, fo[4].data_ptr(),
                                               st.data_ptr()))
        step_fit()
        torch.cuda.synchronize()
        ft = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fit()
            e1.record(stream)
            e1.synchronize()
            ft.append(e0.elapsed_time(e1))
        cells = (NX - 1) * (NY - 1)
        te_map = fo[0].cpu().numpy().reshape(NY, NX)[:-1, :-1]
        fit = {"cells_per_s": cells / (np.mean(ft) * 1e-3), "ms": float(np.mean(ft)), "cells": cells, "atype": "joint",
               "alph": False, "converged_frac": float((st.cpu().numpy() == 1).sum() / cells),
               "median_Te_km": float(np.median(te_map)), "true_Te_km": 40.0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = BYTES_PER_UNIT * units / (ms_per_step * 1e-3) / 1e9 / world  # per GPU
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, dt, u = cpu_reference(topo, grav, k, k0, [0, NS - 1], threads)
        cpu = {"value": v, "unit": "units/s", "cores": threads, "kind": "port",
               "sample": f"scales [0, {NS - 1}] of {NS} ({u:.3g} units, {dt:.1f} s), float32 oracle port, OpenMP over planes"}
    line = {
        "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {NX}x{NY} topo+Bouguer, {DX:g} km, ns={NS}, na={NA} (BASELINE configs[1])",
                   "engine": args.engine, "l2": "flushed between timed steps (256 MiB write)",
                   "sharding": f"scales over {world} rank(s)"},
        "clocks": clocks.summary(),
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "note": "whole step: algorithmic 547 B/unit x units / step time, per GPU"},
        "cpu_baseline": cpu, "fit": fit,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
