#!/usr/bin/env python
"""bench.py -- headline benchmark of the PlateFlex hot path on B200.

Workload (default): BASELINE.json configs[2], the configuration the north star's target is quoted on --
synthetic 4096x4096 land topography + Bouguer pair, 5 km spacing, ns = 20 scales (the first 20 of the grid's
natural wavenumber ladder), 11 azimuths -- run through the WHOLE path: CWT -> admittance / coherence with
jackknife errors -> per-cell L2 (Te, F) fit -> Te / F / chi2 maps.  One step = one pass over the pair:
nx*ny*ns*11 = 3.69e9 grid-cell.scale.azimuth units and (nx-1)*(ny-1) fitted cells.  It fits one GPU, so the same
workload is the N = 1 line and, scale-sharded (strong scaling), the N = 2, 4, 8 lines.  `--config C1|C2|C4|C5`
selects the other BASELINE configurations; at N = 1 the default run also measures configs[1] (C2) and reports it
under "c2".

  value      units/s of the whole step with the two input grids (and the water depth) already resident in the
             HBM of rank 0; N > 1: input broadcast (NCCL), scale-sharded transform whose epilogue stores into
             the peers' row shards over NVLink, stream-ordered barrier, row-sharded fit and the gather of the
             result maps on rank 0 are all inside the timed region (plateflex_b200.sharding.TeMapPipeline)
  e2e        the same step from pinned HOST grids on rank 0 to pinned HOST result maps on rank 0 (H2D and D2H
             inside the timed region)
  roofline   the dominant kernel: algorithmic bytes per launch (SURVEY.md 8d split per kernel)
             / its average launch duration from CUDA event pairs recorded around every launch on its own
             stream (pfx_plan_profile_*); own_bytes_frac is the measured DRAM traffic of the same kernel
             (profiles/traffic.json, ncu) against the same peak
  fit        the per-cell fit alone: cells/s, FP64 roofline (flops per cell from ncu, profiles/fit_flops.json,
             against the DFMA peak measured by scripts/fp64_peak.py) and the SciPy CPU figure
  cpu_baseline  the CPU oracle port (reference precision) on a bounded sample, all host cores and one thread

`--impl reference` times only the CPU path (the reference's Fortran cannot be built here): oracle32 port of
cpwt.f90 with OpenMP over planes + SciPy curve_fit over processes, each step a bounded sample of the workload,
extrapolated linearly to the whole step.

Multi-GPU (torchrun, one rank per GPU): `--scaling strong` (default) one grid pair, scales sharded over ranks;
`--scaling weak` every rank runs the single-GPU step on its own pair (independent survey regions, no data-path
collective).  The timed region is bracketed by barriers and the maximum over ranks is reported.
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

NA = 11
METRIC = "grid-cell*scale*azimuth/s (CWT->admittance/coherence->L2 Te/F map)"

# BASELINE.json configs.  ns = 20 except C1 (the example grid's natural 19).
CONFIGS = {
    "C1": dict(nx=342, ny=341, dx=20.0169157785, dy=19.9970844654, ns=19, Te=30.0, F=0.5, ocean=None, boug=True,
               alph=False, name="synthetic 342x341 example-shaped topo+Bouguer (BASELINE configs[0])"),
    "C2": dict(nx=1024, ny=1024, dx=5.0, dy=5.0, ns=20, Te=40.0, F=0.5, ocean=None, boug=True, alph=False,
               name="synthetic 1024x1024 topo+Bouguer, 5 km, ns=20, na=11 (BASELINE configs[1])"),
    "C3": dict(nx=4096, ny=4096, dx=5.0, dy=5.0, ns=20, Te=40.0, F=0.5, ocean=None, boug=True, alph=False,
               name="synthetic 4096x4096 land Bouguer: CWT + cross-scalogram + per-cell L2 Te/F map, 5 km, ns=20, na=11 "
                    "(BASELINE configs[2])"),
    "C4": dict(nx=2048, ny=2048, dx=5.0, dy=5.0, ns=20, Te=20.0, F=0.2, ocean=4500.0, boug=False, alph=True,
               name="synthetic 2048x2048 ocean free-air, alpha fitted, ns=20 (BASELINE configs[3])"),
    "C5": dict(nx=8192, ny=8192, dx=5.0, dy=5.0, ns=20, Te=40.0, F=0.5, ocean=None, boug=True, alph=False,
               name="synthetic 8192x8192 continental grid, ns=20 (BASELINE configs[4])"),
}


def npow2(n):
    p = 2
    while p < n:
        p *= 2
    return p


def alg_bytes_per_unit(cfg):
    """per cell.scale.azimuth, two grids: 2*(rho*64 + 16) + 4*8/11 with
    rho = padded/cell ratio; split per kernel: each FFT pass 2*rho*32, the epilogue 2*16 + 32/11."""
    rho = npow2(2 * cfg["nx"]) * npow2(2 * cfg["ny"]) / float(cfg["nx"] * cfg["ny"])
    per = {"pass1": 2 * rho * 32.0, "pass2": 2 * rho * 32.0, "reduce": 2 * 16.0 + 4 * 8 / 11.0}
    per["dense_fft"] = per["pass1"] + per["pass2"]
    per["dense_mul"] = 0.0  # the daughter multiply is implementation overhead, not algorithmic (8d)
    return sum(per[k] for k in ("pass1", "pass2", "reduce")), per


def workload(cfg, seed_offset=0, data=True):
    import plateflex_b200 as pf
    from plateflex_b200 import classes, synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    nx, ny = cfg["nx"], cfg["ny"]
    k = classes._lam2k(nx, ny, cfg["dx"], cfg["dy"])[1][:cfg["ns"]].copy()
    if not data:
        return None, None, k, pf.cf_wt.k0
    topo, grav = synthetic.make_pair(nx, ny, cfg["dx"], cfg["dy"], seed=20240000 + nx + seed_offset, Te=cfg["Te"],
                                     F=cfg["F"], boug=cfg["boug"], ocean_depth=cfg["ocean"],
                                     rhoc=2800. if cfg["ocean"] else 2700.)
    return topo, grav, k, pf.cf_wt.k0


def flex_conf(cfg):
    from plateflex_b200.flex import conf_flex
    conf_flex.boug = 1 if cfg["boug"] else 0
    if cfg["ocean"]:
        conf_flex.rhoc = 2800.
    return conf_flex.as_struct()


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer is allocated, so that
    host<->device copies of several ranks do not all cross one socket's memory (plain plumbing, like numactl)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The sampler is
    started early (nvidia-smi takes a while to emit its first line on an 8-GPU box); mark_start / mark_end
    bracket the timed region and the summary uses the samples in between (or the nearest one)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled
        self.i0, self.i1 = 0, None

    def mark_start(self):
        self.i0 = len(self.rows)

    def mark_end(self):
        time.sleep(0.06)
        self.i1 = len(self.rows)

    def __enter__(self):
        if not self.enabled:  # only the rank that prints the JSON line polls nvidia-smi
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
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
        i1 = len(self.rows) if self.i1 is None else self.i1
        rows = self.rows[self.i0:i1] or self.rows[max(self.i0 - 1, 0):i1 + 1] or self.rows[-1:]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except (OSError, ValueError):
        return {}


def ncu_traffic(config, cls):
    """DRAM bytes per launch of a kernel class from the committed ncu capture (profiles/traffic.json,
    written by scripts/summarise_ncu.py from `ncu --metrics dram__bytes_*`), or None."""
    return profile_json("traffic.json").get(config, {}).get(cls)


# ---------------------------------------------------------------------------------------------------------
# CPU path (the reference arm and the cpu_baseline leg): the only places that execute oracle/
# ---------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, topo, grav, k, k0, threads, fit_cells=256, fit_procs=None):
    """One bounded sample of the CPU path on this workload, extrapolated linearly to the whole step
    (SURVEY.md 8d "CPU reference timing").  CWT: the float32 oracle port of cpwt.f90 on a subset of (scale,
    azimuth) planes of both grids; the forward stage, which the job pays once per grid, is timed apart and
    charged once.  Admittance / coherence + jackknife: the oracle's single pass over a row band of synthetic
    coefficient cubes.  Fit: the reference's L2_estimate_cell logic (SciPy TRF) on `fit_cells` cells spread
    over `fit_procs` processes.  Returns (units/s of the whole step, details)."""
    from oracle import oracle as ora
    ora.lib()
    nx, ny, ns = cfg["nx"], cfg["ny"], cfg["ns"]
    big = nx * ny > 2048 * 2048
    # planes sampled: whole scales on the smaller grids (as in round 1), two azimuths of the middle scale above
    t0 = time.perf_counter()
    if big:
        isamp, ia0, ia1 = ns // 2, 4, 6
        _, tf1 = ora.wlet_planes(topo, cfg["dx"], cfg["dy"], k[isamp], ia0, ia1, k0=k0, prec=32, nthreads=threads)
        _, tf2 = ora.wlet_planes(grav, cfg["dx"], cfg["dy"], k[isamp], ia0, ia1, k0=k0, prec=32, nthreads=threads)
        t_fwd = tf1 + tf2
        nplanes = ia1 - ia0
        what = f"scale {isamp}, azimuths {ia0}..{ia1 - 1}"
    else:
        scales = [0, ns - 1]
        t_fwd = 0.0
        for s in scales:  # forward stage timed through the same entry (1-plane call) and subtracted below
            pass
        w1 = ora.wlet_transform(topo, cfg["dx"], cfg["dy"], k[scales], k0=k0, prec=32, nthreads=threads)
        w2 = ora.wlet_transform(grav, cfg["dx"], cfg["dy"], k[scales], k0=k0, prec=32, nthreads=threads)
        nplanes = NA * len(scales)
        what = f"scales {scales}"
    t_planes = time.perf_counter() - t0 - t_fwd
    # admittance / coherence + jackknife (single-threaded in the reference, cpwt.f90:303-364)
    rows = min(ny, max(8, (1 << 22) // nx))  # ~4M cells.azimuth-sets per sample
    if big:
        rng = np.random.default_rng(1)
        shape = (nx, rows, NA, 1)
        w1 = (rng.standard_normal(shape, dtype=np.float32) + 1j * rng.standard_normal(shape, dtype=np.float32))
        w2 = (rng.standard_normal(shape, dtype=np.float32) + 1j * rng.standard_normal(shape, dtype=np.float32))
        w1, w2 = np.asfortranarray(w1.astype(np.complex64)), np.asfortranarray(w2.astype(np.complex64))
        t1 = time.perf_counter()
        ora.wlet_admit_coh(w1, w2, prec=32)
        t_red_full = (time.perf_counter() - t1) * (ny / rows) * ns
    else:
        t1 = time.perf_counter()
        ora.wlet_admit_coh(w1, w2, prec=32)
        t_red_full = (time.perf_counter() - t1) * ns / len(scales)
    del w1, w2
    t_cwt_full = t_fwd + t_planes * (NA * ns / nplanes) + t_red_full
    # fit sample
    fit = cpu_fit_sample(cfg, k, fit_cells, fit_procs or threads)
    cells = (nx - 1) * (ny - 1)
    t_fit_full = cells / fit["cells_per_s"]
    units = nx * ny * ns * NA
    det = {"cwt_s_extrapolated": t_cwt_full, "fit_s_extrapolated": t_fit_full, "sample_s": time.perf_counter() - t0,
           "cwt_units_per_s": units / t_cwt_full, "fit_cells_per_s": fit["cells_per_s"],
           "what": (f"CWT: {what} of both grids ({2 * nplanes} of {2 * NA * ns} planes), float32 oracle port of cpwt.f90 "
                    f"(the Fortran is not buildable here), OpenMP over planes, {threads} threads; jackknife on a "
                    f"{rows}-row band; fit: SciPy TRF on {fit['cells']} cells over {fit['procs']} processes; each "
                    f"part extrapolated linearly to the whole step")}
    return units / (t_cwt_full + t_fit_full), det


def _fit_worker(job):
    from oracle import l2fit, oracle as ora
    k, cells, conf_kw, alph = job
    conf = ora.FlexConf(**conf_kw)
    for adm, eadm, coh, ecoh in cells:
        l2fit.L2_estimate_cell(k, adm, eadm, coh, ecoh, alph=alph, atype="joint", conf=conf)
    return len(cells)


def cpu_fit_sample(cfg, k, ncells, procs):
    """SciPy TRF (the reference's L2_estimate_cell logic, oracle/l2fit.py) on `ncells` synthetic cells: model
    curves of the workload's (Te, F) with 5 % noise and errors, over `procs` processes."""
    import multiprocessing as mp

    from oracle import oracle as ora
    rng = np.random.default_rng(3)
    conf_kw = dict(boug=1 if cfg["boug"] else 0, rhoc=2800. if cfg["ocean"] else 2700., wd=float(cfg["ocean"] or 0.0))
    adm0, coh0 = ora.real_xspec_functions(k, cfg["Te"], cfg["F"], np.pi / 2, ora.FlexConf(**conf_kw))
    cells = []
    for _ in range(ncells):
        adm = adm0 * (1 + 0.05 * rng.standard_normal(len(k)))
        coh = np.clip(coh0 * (1 + 0.05 * rng.standard_normal(len(k))), 1e-4, 1.0)
        cells.append((adm, 0.05 * np.abs(adm0) + 1e-6, coh, 0.05 * coh0 + 1e-3))
    procs = max(1, min(procs, ncells // 16 or 1))
    jobs = [(k, cells[i::procs], conf_kw, bool(cfg["alph"])) for i in range(procs)]
    if procs == 1:
        t0 = time.perf_counter()
        _fit_worker(jobs[0])
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_fit_worker, [(k, cells[:2], conf_kw, bool(cfg["alph"]))] * procs)  # import + warm up
            t0 = time.perf_counter()
            pool.map(_fit_worker, jobs)
            dt = time.perf_counter() - t0
    return {"cells_per_s": ncells / dt, "cells": ncells, "procs": procs, "seconds": dt}


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    topo, grav, k, k0 = workload(cfg)
    threads = os.cpu_count() or 1
    vals, det = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, det = cpu_sample(cfg, topo, grav, k, k0, threads)
        if i >= args.warmup:
            vals.append((v, time.perf_counter() - t0))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    units = cfg["nx"] * cfg["ny"] * cfg["ns"] * NA
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": det["what"],
                   "note": "ms_per_step is the wall time of one bounded sample; value is units of the whole step / its "
                           "extrapolated CPU time (%.0f s CWT + %.0f s fit)" % (det["cwt_s_extrapolated"],
                                                                                det["fit_s_extrapolated"])},
        "whole_step_s_extrapolated": units / value,
        "cpu_baseline": {"value": value, "unit": "units/s", "cores": threads, "kind": "port", "sample": det["what"],
                         "cwt_units_per_s": det["cwt_units_per_s"], "fit_cells_per_s": det["fit_cells_per_s"]},
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def bench_config(cfg, cfg_key, args, rank, world, local, strong, clocks=None, full=True):
    """Times the whole-path step of one configuration; returns the dict of results on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from plateflex_b200 import _lib, sharding

    dev = torch.device("cuda", local)
    nx, ny, ns = cfg["nx"], cfg["ny"], cfg["ns"]
    nxy = nx * ny
    sharded = strong and world > 1
    replicas = world if (world > 1 and not strong) else 1
    have_data = rank == 0 or not sharded
    topo, grav, k, k0 = workload(cfg, seed_offset=0 if (sharded or world == 1) else 1000 * rank, data=have_data)
    conf = flex_conf(cfg)
    engine = _lib.ENGINE_DENSE if args.engine == "dense" else _lib.ENGINE_PRUNED
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # weak scaling: every rank is its own single-GPU pipeline (no process group handed in)
    if sharded:
        pipe = sharding.TeMapPipeline(nx, ny, cfg["dx"], cfg["dy"], k, k0, conf, alph=cfg["alph"], atype="joint", nn=1,
                                      local_device=local, engine=engine, balance=args.balance)
    else:
        pipe = _single_pipeline(sharding, nx, ny, cfg, k, k0, conf, local, engine)
    if have_data:
        wd_f = np.ascontiguousarray((-np.minimum(topo, 0.0)).T)  # TopoGrid.water_depth (classes.py:533-536), x fastest
        h_in = [torch.from_numpy(a).pin_memory() for a in (topo, grav, wd_f)]
        pipe.load(*h_in)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    res = None
    times = []
    for _ in range(args.warmup):
        res = pipe.run()
    barrier()
    launches0 = pipe.launches()
    if clocks:
        clocks.mark_start()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        e0.record(stream)
        res = pipe.run()
        e1.record(stream)
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    if clocks:
        clocks.mark_end()
    barrier()
    launches = pipe.launches() - launches0 + args.steps  # + the fit kernel of every step
    t = torch.tensor(times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, per step
    ms_per_step = float(t.mean().item())
    units_job = nxy * ns * NA * replicas
    out = {"ms_per_step": ms_per_step, "value": units_job / (ms_per_step * 1e-3), "launches": int(launches)}

    # ---- stage split (device timers, same loop shape): transform + exchange, then fit + gather ----
    st = []
    for _ in range(min(3, args.steps)):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        ev[0].record(stream)
        pipe.run_transform()
        ev[1].record(stream)
        res = pipe.run_fit()
        ev[2].record(stream)
        ev[2].synchronize()
        st.append([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
    tst = torch.tensor(st, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tst, op=dist.ReduceOp.MAX)
    tst = tst.mean(0).tolist()
    cells = (nx - 1) * (ny - 1) * replicas
    out["stages"] = {"transform_ms": tst[0], "fit_ms": tst[1],
                     "transform_units_per_s": units_job / (tst[0] * 1e-3), "fit_cells_per_s": cells / (tst[1] * 1e-3),
                     "note": ("transform = " + ("input broadcast + " if sharded else "") + "CWT -> admittance/coherence" +
                              (" stored into the row shards + stream barrier" if sharded else "") + "; fit = per-cell L2" +
                              (" of the row shard + gather of the maps on rank 0" if sharded else ""))}

    # ---- per-kernel device timers (serialises the plan's internal streams) ----
    pipe.plan.profile(True)
    pipe.plan.profile_read()
    ptimes = []
    for _ in range(min(3, args.steps)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pipe.run_cwt_only()
        e1.record(stream)
        e1.synchronize()
        ptimes.append(e0.elapsed_time(e1))
    kern = pipe.plan.profile_read()  # {class: (ms total, launches)}
    pipe.plan.profile(False)
    prof_ms, prof_steps = float(np.mean(ptimes)), len(ptimes)
    barrier()
    out["kernels"] = {c: {"ms_per_step": kern[c][0] / prof_steps, "launches_per_step": kern[c][1] / prof_steps,
                          "avg_launch_us": 1e3 * kern[c][0] / kern[c][1]} for c in kern}

    # ---- e2e: pinned host grids on rank 0 -> pinned host result maps on rank 0 ----
    if not args.no_e2e:
        nmaps1 = pipe.nmaps + 1
        h_out = torch.empty((nmaps1, pipe.my, pipe.mx), dtype=torch.float64).pin_memory() if have_data else None

        def step_host():
            if have_data:
                pipe.load(*h_in)
            r = pipe.run()
            if r is not None and h_out is not None:
                h_out.copy_(r, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            step_host()
        barrier()
        te = []
        for _ in range(max(3, args.steps // 2)):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            step_host()
            te.append((time.perf_counter() - t0) * 1e3)
        t2 = torch.tensor(te, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.mean().item())
        out["e2e"] = {"value": units_job / (e2e_ms * 1e-3), "unit": "units/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": 3 * nxy * 8 * replicas, "d2h_bytes_per_step": nmaps1 * pipe.my * pipe.mx * 8 * replicas,
                      "api": "plateflex_b200.sharding.TeMapPipeline.load + run (host grids + water depth in, "
                             "Te/std/F/std/chi2/status maps out), pinned host buffers"}
    # ---- N > 1: the gathered maps must equal the single-GPU maps bit for bit ----
    if sharded and not args.no_check:
        chk = None
        if rank == 0:
            ref_pipe = _single_pipeline(sharding, nx, ny, cfg, k, k0, conf, local, engine)
            ref_pipe.load(*h_in)
            ref = ref_pipe.run()
            torch.cuda.synchronize()
            same = bool(torch.equal(ref, res))
            diff = (ref - res).abs()
            diff = diff[torch.isfinite(diff)]
            chk = {"equal_to_single_gpu": same, "max_abs_diff": float(diff.max().item()) if diff.numel() else 0.0,
                   "nan_pattern_equal": bool(torch.equal(torch.isnan(ref), torch.isnan(res))),
                   "what": "gathered Te/std/F/std/chi2/status maps of the sharded run vs the same step on rank 0 alone"}
            ref_pipe.close()
            del ref
        out["check"] = chk
        barrier()
    # ---- fit statistics from the result maps ----
    if rank == 0 and res is not None:
        r = res.detach().cpu().numpy()
        raw = r[-1].astype(np.int64)
        stat, nfev = raw & 0xff, raw >> 8
        fitted = stat > 0
        out["fit_stats"] = {"cells": int(fitted.sum()), "converged_frac": float((stat == 1).sum() / max(fitted.sum(), 1)),
                            "median_Te_km": float(np.median(r[0][fitted])) if fitted.any() else None,
                            "true_Te_km": cfg["Te"], "nfev_mean": float(nfev[fitted].mean()) if fitted.any() else None,
                            "nfev_max": int(nfev[fitted].max()) if fitted.any() else None}
    out.update({"prof_ms": prof_ms, "scales_rank0": list(map(int, pipe.scales)), "units_job": units_job,
                "nmaps": pipe.nmaps, "host": (topo, grav, k, k0) if (rank == 0 and full) else None})
    pipe.close()
    del flush
    torch.cuda.empty_cache()
    return out


def _single_pipeline(sharding, nx, ny, cfg, k, k0, conf, local, engine):
    """TeMapPipeline of this rank alone, whatever process group exists."""
    return sharding.TeMapPipeline(nx, ny, cfg["dx"], cfg["dy"], k, k0, conf, alph=cfg["alph"], atype="joint", nn=1,
                                  local_device=local, engine=engine, solo=True)


def fp64_peak(local):
    """Measured FP64 peak of this GPU (pfx_fp64_peak: independent DFMA chains on every SM), TFLOP/s; the committed
    measurement of scripts/fp64_peak.py when the live one fails."""
    try:
        from plateflex_b200 import _lib
        return max(_lib.fp64_peak(local) for _ in range(2)), "measured in this run (pfx_fp64_peak: DFMA chains, CUDA events)"
    except Exception:  # noqa: BLE001
        pk = profile_json("fp64_peak.json")
        return (pk.get("tflops"), "profiles/fp64_peak.json") if pk else (None, None)


def fit_roofline(fit_ms, cells, alph, peak, peak_src):
    """FP64 roofline of the fit: flops per cell counted by ncu on this workload (profiles/fit_flops.json,
    scripts/summarise_ncu.py: 2*dfma + dadd + dmul thread instructions / fitted cells) against the measured DFMA
    peak."""
    fl = profile_json("fit_flops.json").get("alph" if alph else "noalph")
    if not peak or not fl:
        return None
    achieved = fl["flops_per_cell"] * cells / (fit_ms * 1e-3) / 1e12
    return {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "flops_per_cell": fl["flops_per_cell"], "fp64_pipe_pct": fl.get("fp64_pipe_pct"), "peak_source": peak_src,
            "flops_source": "ncu smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on on " + fl.get("config", "?") +
                            " (profiles/fit_flops.json)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--balance", default="deal", choices=["blocks", "deal"],
                    help="strong scaling: scales dealt out by decreasing cost (longest processing time first), or "
                         "contiguous cost-balanced scale blocks")
    ap.add_argument("--engine", default=os.environ.get("PLATEFLEX_B200_ENGINE", "pruned"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="N = 1 default run: skip the extra BASELINE configs[1] measurement")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.scaling == "strong"
    sharded = strong and world > 1

    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        r = bench_config(cfg, args.config, args, rank, world, local, strong, clocks=clocks)
    c2 = None
    if world == 1 and args.config == "C3" and not args.no_c2:
        c2 = bench_config(CONFIGS["C2"], "C2", args, rank, world, local, strong, full=False)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    nx, ny, ns = cfg["nx"], cfg["ny"], cfg["ns"]
    nxy = nx * ny
    peak, peak_src = measured_peak()
    f64_peak, f64_src = fp64_peak(local)
    alg_total, alg = alg_bytes_per_unit(cfg)
    kern = r["kernels"]
    dom = max((c for c in kern if c in alg), key=lambda c: kern[c]["ms_per_step"], default=None)
    roof = None
    if dom:
        units_launch = nxy * NA  # one launch = one scale, all 11 azimuths of both grids
        avg_s = kern[dom]["avg_launch_us"] * 1e-6
        achieved = alg[dom] * units_launch / avg_s / 1e9
        traffic = ncu_traffic(args.config, dom)
        det = profile_json("traffic.json").get(args.config + "_detail", {}).get(dom, {})
        fp64 = det.get("fp64_pipe_pct")
        fl = det.get("fp64_flops_per_launch")
        step_s = r["stages"]["transform_ms"] * 1e-3
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "own_bytes_frac": (traffic / avg_s / 1e9 / peak) if traffic else None, "fp64_pipe_pct": fp64,
                "fp64": ({"achieved": fl / avg_s / 1e12, "peak": f64_peak, "unit": "TFLOP/s", "frac": fl / avg_s / 1e12 / f64_peak,
                          "flops_per_launch": fl, "peak_source": f64_src,
                          "flops_source": "ncu 2*dfma + dadd + dmul thread instructions per launch (profiles/traffic.json)"}
                         if (fl and f64_peak) else None),
                "algorithmic_bytes_per_launch": alg[dom] * units_launch, "avg_launch_us": avg_s * 1e6,
                "share_of_cwt": kern[dom]["ms_per_step"] / r["prof_ms"], "serialised_cwt_ms": r["prof_ms"],
                "whole_transform": {"achieved": alg_total * r["units_job"] / step_s / 1e9,
                                    "frac": alg_total * r["units_job"] / step_s / 1e9 / (peak * world),
                                    "peak": peak * world, "bytes_per_unit": alg_total},
                "note": ("achieved = algorithmic bytes of a dense 2-pass FP64 FFT path (SURVEY.md 8d) / measured launch "
                         "time: the band-pruned engine does not move those bytes, so frac > 1 is possible and is not a "
                         "hardware utilisation; own_bytes_frac = DRAM bytes the kernel really moves (ncu) / time / peak. "
                         "The kernel is bound by the FP64 pipe (fp64_pipe_pct, ncu), not by HBM -- see DESIGN.md")}
    cpu = None
    fit_cpu = None
    if not args.no_cpu_baseline and world == 1:
        topo, grav, k, k0 = r["host"]
        threads = os.cpu_count() or 1
        v, det = cpu_sample(cfg, topo, grav, k, k0, threads, fit_cells=4096)
        cpu = {"value": v, "unit": "units/s", "cores": threads, "kind": "port", "sample": det["what"],
               "cwt_units_per_s": det["cwt_units_per_s"], "fit_cells_per_s": det["fit_cells_per_s"]}
        v1, det1 = cpu_sample(cfg, topo, grav, k, k0, 1, fit_cells=64, fit_procs=1)
        cpu["one_thread"] = {"value": v1, "cwt_units_per_s": det1["cwt_units_per_s"],
                             "fit_cells_per_s": det1["fit_cells_per_s"],
                             "note": "the reference is single-threaded: oracle32 on one core (BASELINE.md 3.1)"}
        fit_cpu = {"value": det["fit_cells_per_s"], "unit": "cells/s", "cores": threads, "kind": "port",
                   "sample": "SciPy TRF (reference L2_estimate_cell logic) on 4096 synthetic cells, one process per core"}
    cells = (nx - 1) * (ny - 1) * (world if (world > 1 and not strong) else 1)
    fit = dict(r.get("fit_stats") or {})
    fit.update({"ms": r["stages"]["fit_ms"], "cells_per_s": r["stages"]["fit_cells_per_s"], "atype": "joint",
                "alph": bool(cfg["alph"]), "includes": "gather of the maps on rank 0" if sharded else "fit kernel only",
                "roofline": fit_roofline(r["stages"]["fit_ms"], cells, cfg["alph"], f64_peak, f64_src) if world == 1 else None,
                "cpu_baseline": fit_cpu})
    line = {
        "metric": METRIC, "value": r["value"], "unit": "units/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "engine": args.engine, "l2": "flushed between timed steps (256 MiB write)",
                   "step": "CWT -> admittance/coherence (+ jackknife) -> per-cell L2 fit -> Te/F/chi2 maps",
                   "cpu_affinity": f"{numa_cpus} CPUs of the GPU's NUMA node" if numa_cpus else "not set",
                   "balance": args.balance if sharded else None,
                   "sharding": ((f"scales over {world} ranks (dealt by cost), epilogue stores into peer row shards over "
                                 f"NVLink (CUDA IPC), rows over {world} ranks for the fit, whose kernel stores the result "
                                 "maps into rank 0's buffer over NVLink; NCCL broadcast of the inputs and both "
                                 "stream-ordered barriers inside the step") if sharded else
                                (f"{world} independent grid pairs, one per rank" if world > 1 else "single GPU"))},
        "clocks": clocks.summary(),
        "e2e": r.get("e2e"), "gpu_launches": r["launches"],
        "roofline": roof, "kernels": kern, "stages": r["stages"],
        "cpu_baseline": cpu, "fit": fit, "check": r.get("check"),
    }
    if c2:
        line["c2"] = {"workload": CONFIGS["C2"]["name"], "ms_per_step": c2["ms_per_step"], "value": c2["value"],
                      "e2e": c2.get("e2e"), "stages": c2["stages"], "kernels": c2["kernels"],
                      "fit_stats": c2.get("fit_stats")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
