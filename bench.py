#!/usr/bin/env python
"""bench.py -- headline benchmark of the PlateFlex hot path on B200.

Workload (default): BASELINE.json configs[1] -- synthetic 1024x1024 topography + Bouguer pair,
5 km spacing, ns = 20 scales (the first 20 of the grid's natural wavenumber ladder), 11 azimuths.
One step = one pass CWT -> admittance / coherence with jackknife errors for the pair:
nx*ny*ns*11 = 2.31e8 grid-cell.scale.azimuth units, two grids in, four (nx,ny,ns) maps out.
`--config C1|C3|C4|C5` selects the other BASELINE configurations (parity / sizing cases).

  value      units/s with the two input grids already resident in HBM (device-pointer C ABI)
  e2e        same metric through the host-pointer C ABI call (pinned host buffers, H2D of the
             two grids and D2H of the four maps inside the timed region)
  roofline   the dominant kernel: algorithmic bytes per launch (SURVEY.md 8d split per kernel,
             see DESIGN.md) / its average launch duration, from CUDA event pairs recorded around
             every launch on its own stream inside the timed region (pfx_plan_profile_*)
  fit        per-cell L2 (Te, F[, alpha]) fit of the resulting maps, cells/s (second metric)
  cpu_baseline  the CPU oracle port (reference precision, OpenMP over planes) on a bounded sample

`--impl reference` times only the CPU oracle port (the reference's Fortran cannot be built here).

Multi-GPU (torchrun, one rank per GPU):
  --scaling weak   (default) every rank transforms its own grid pair (independent survey regions,
                   no data-path collective); value = N * units / max-over-ranks time.
  --scaling strong one grid pair, scales sharded over ranks, the four maps re-sharded by grid rows
                   (NCCL all-to-all) for the cell-sharded fit; see plateflex_b200/sharding.py.
The timed region is bracketed by barriers and the maximum over ranks is reported.
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
METRIC = "grid-cell*scale*azimuth/s (CWT->admittance/coherence)"

# BASELINE.json configs.  ns = 20 except C1 (the example grid's natural 19).
CONFIGS = {
    "C1": dict(nx=342, ny=341, dx=20.0169157785, dy=19.9970844654, ns=19, Te=30.0, F=0.5, ocean=None, boug=True,
               alph=False, name="synthetic 342x341 example-shaped topo+Bouguer (BASELINE configs[0])"),
    "C2": dict(nx=1024, ny=1024, dx=5.0, dy=5.0, ns=20, Te=40.0, F=0.5, ocean=None, boug=True, alph=False,
               name="synthetic 1024x1024 topo+Bouguer, 5 km, ns=20, na=11 (BASELINE configs[1])"),
    "C3": dict(nx=4096, ny=4096, dx=5.0, dy=5.0, ns=20, Te=40.0, F=0.5, ocean=None, boug=True, alph=False,
               name="synthetic 4096x4096 land Bouguer, 5 km, ns=20 (BASELINE configs[2])"),
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


def workload(cfg, seed_offset=0):
    import plateflex_b200 as pf
    from plateflex_b200 import classes, synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    nx, ny = cfg["nx"], cfg["ny"]
    topo, grav = synthetic.make_pair(nx, ny, cfg["dx"], cfg["dy"], seed=20240000 + nx + seed_offset, Te=cfg["Te"],
                                     F=cfg["F"], boug=cfg["boug"], ocean_depth=cfg["ocean"],
                                     rhoc=2800. if cfg["ocean"] else 2700.)
    k = classes._lam2k(nx, ny, cfg["dx"], cfg["dy"])[1][:cfg["ns"]].copy()
    return topo, grav, k, pf.cf_wt.k0


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


def ncu_traffic(config, cls):
    """DRAM bytes per launch of a kernel class from the committed ncu capture (profiles/traffic.json,
    written by scripts/summarise_ncu.py from `ncu --metrics dram__bytes_*`), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(config, {}).get(cls)
    except (ValueError, OSError):
        return None


def cpu_reference(cfg, topo, grav, k, k0, sample_scales, threads):
    """The CPU oracle port in the reference's precision (float32, radix-2 fourn) on `sample_scales`."""
    from oracle import oracle as ora
    ora.lib()
    ks = k[list(sample_scales)]
    t0 = time.perf_counter()
    w1 = ora.wlet_transform(topo, cfg["dx"], cfg["dy"], ks, k0=k0, prec=32, nthreads=threads)
    w2 = ora.wlet_transform(grav, cfg["dx"], cfg["dy"], ks, k0=k0, prec=32, nthreads=threads)
    ora.wlet_admit_coh(w1, w2, prec=32)
    dt = time.perf_counter() - t0
    units = topo.shape[0] * topo.shape[1] * NA * len(ks)
    return units / dt, dt, units


def cpu_sample_scales(cfg):
    """Bounded CPU sample: first and last scale up to 2048^2-cell grids, one scale above."""
    return [0, cfg["ns"] - 1] if cfg["nx"] * cfg["ny"] <= 2048 * 2048 else [cfg["ns"] // 2]


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    topo, grav, k, k0 = workload(cfg)
    threads = os.cpu_count() or 1
    sample = cpu_sample_scales(cfg)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, units = cpu_reference(cfg, topo, grav, k, k0, sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    what = (f"scales {sample} of {cfg['ns']} ({units:.3g} units per step), float32 oracle port of cpwt.f90 "
            f"(the Fortran is not buildable here), OpenMP over planes")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": what},
        "cpu_baseline": {"value": value, "unit": "units/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--balance", default="blocks", choices=["blocks", "deal"],
                    help="strong scaling with --exchange p2p: contiguous cost-balanced scale blocks, or scales dealt "
                         "out by decreasing cost (longest processing time first)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="strong scaling: epilogue stores into peer row shards over NVLink (p2p) or NCCL all-to-all")
    ap.add_argument("--engine", default=os.environ.get("PLATEFLEX_B200_ENGINE", "pruned"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, cfg, rank)

    import torch
    import torch.distributed as dist
    from plateflex_b200 import _lib, sharding

    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    strong = args.scaling == "strong" and world > 1
    nx, ny, ns = cfg["nx"], cfg["ny"], cfg["ns"]
    nxy = nx * ny

    # weak: every rank owns a different pair (seed + rank); strong: the same pair everywhere
    topo, grav, k, k0 = workload(cfg, seed_offset=0 if (strong or world == 1) else 1000 * rank)
    engine = _lib.ENGINE_DENSE if args.engine == "dense" else _lib.ENGINE_PRUNED
    plan = _lib.Plan(nx, ny, cfg["dx"], cfg["dy"], k, k0, engine=engine, device=local)
    stream = torch.cuda.current_stream()
    plan.set_stream(stream.cuda_stream)

    cost = plan.scale_cost() if strong else None
    if strong:
        is0, is1 = sharding.scale_range(ns, world, rank, cost)
    else:
        is0, is1 = 0, ns
    nsl = is1 - is0

    d_topo = torch.from_numpy(topo).to(dev)
    d_grav = torch.from_numpy(grav).to(dev)
    outs = [torch.empty(max(nsl, 1) * nxy, dtype=torch.float64, device=dev) for _ in range(4)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    p2p = strong and args.exchange == "p2p"
    resharder = sharding.RowResharder(nx, ny, ns, world, rank, dev, cost=cost) if (strong and not p2p) else None
    peers = sharding.PeerShards(nx, ny, ns, world, rank, local) if p2p else None

    my_scales = sharding.deal_scales(cost, world)[rank] if (p2p and args.balance == "deal") else None

    def step_dev():
        if p2p:  # the epilogue of every scale stores its rows straight into the owners' shards (NVLink)
            if my_scales is not None:
                peers.transform_scales(plan, d_topo.data_ptr(), d_grav.data_ptr(), my_scales)
            else:
                peers.transform(plan, d_topo.data_ptr(), d_grav.data_ptr(), is0, is1)
            return
        if nsl > 0:
            _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, d_topo.data_ptr(), d_grav.data_ptr(), is0, is1,
                                                       *[o.data_ptr() for o in outs]))
        if strong:  # the exchange step of the sharded pipeline: scale-major -> row shards holding all scales
            resharder.exchange(outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    times = []
    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        for _ in range(args.warmup):
            step_dev()
        barrier()
        launches0 = plan.launch_count()
        clocks.mark_start()
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
        clocks.mark_end()
        barrier()
    launches = plan.launch_count() - launches0
    # Second timed loop, same steps, with the per-kernel device timers on: an event pair around every
    # launch on its own stream; the plan serialises its internal streams meanwhile so that each pair
    # times one kernel alone.  prof_ms is this loop's own step time (the denominator of the shares).
    plan.profile(True)
    plan.profile_read()
    ptimes = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if nsl > 0:
            _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, d_topo.data_ptr(), d_grav.data_ptr(), is0, is1,
                                                       *[o.data_ptr() for o in outs]))
        e1.record(stream)
        e1.synchronize()
        ptimes.append(e0.elapsed_time(e1))
    kern = plan.profile_read()  # {class: (ms total, launches)}
    plan.profile(False)
    prof_ms = float(np.mean(ptimes))
    barrier()
    t = torch.tensor(times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, per step
    ms_per_step = float(t.mean().item())
    units_rank = nxy * nsl * NA
    units_job = nxy * ns * NA * (1 if (strong or world == 1) else world)
    value = units_job / (ms_per_step * 1e-3)

    # ---- e2e: host-pointer C ABI with pinned buffers (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        h_in = [torch.from_numpy(a).pin_memory() for a in (topo, grav)]
        h_out = [torch.empty(max(nsl, 1) * nxy, dtype=torch.float64).pin_memory() for _ in range(4)]

        def step_host():
            if nsl > 0:
                _lib.check(_lib.lib.pfx_wlet_admit_coh_host(plan.handle, h_in[0].data_ptr(), h_in[1].data_ptr(), is0,
                                                            is1, *[o.data_ptr() for o in h_out]))

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
        per_rank_out = torch.tensor([4 * nxy * nsl * 8], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(per_rank_out)
        e2e = {"value": units_job / (e2e_ms * 1e-3), "unit": "units/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": 2 * nxy * 8 * world, "d2h_bytes_per_step": int(per_rank_out.item()),
               "api": "pfx_wlet_admit_coh_host (what Project.wlet_admit_coh calls), pinned host buffers"}
        # same call with float32 results (the reference's own output dtype; opt-in, set_output_dtype(float32))
        h_out32 = [torch.empty(max(nsl, 1) * nxy, dtype=torch.float32).pin_memory() for _ in range(4)]

        def step_host32():
            if nsl > 0:
                _lib.check(_lib.lib.pfx_wlet_admit_coh_host_f32(plan.handle, h_in[0].data_ptr(), h_in[1].data_ptr(),
                                                                is0, is1, *[o.data_ptr() for o in h_out32]))

        for _ in range(2):
            step_host32()
        barrier()
        te = []
        for _ in range(max(3, args.steps // 2)):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            step_host32()
            te.append((time.perf_counter() - t0) * 1e3)
        t3 = torch.tensor(te, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        e2e["float32_outputs"] = {"value": units_job / (float(t3.mean().item()) * 1e-3), "unit": "units/s",
                                  "ms_per_step": float(t3.mean().item()),
                                  "d2h_bytes_per_step": int(per_rank_out.item()) // 2}
        del h_out32

    # ---- second metric: per-cell L2 fit of the maps (N=1: whole grid; strong: this rank's row shard) ----
    fit = None
    if not args.no_fit and (world == 1 or strong):
        from plateflex_b200.flex import conf_flex
        conf_flex.boug = 1 if cfg["boug"] else 0
        if cfg["ocean"]:
            conf_flex.rhoc = 2800.
        conf = conf_flex.as_struct()
        alph = 1 if cfg["alph"] else 0
        d_k = torch.from_numpy(k).to(dev)
        wd_full = np.asfortranarray(-np.minimum(topo, 0.0))  # TopoGrid.water_depth, classes.py:533-536
        if p2p:
            row0, nrows = peers.row0, peers.nrows
            map_ptrs = peers.local_map_ptrs()
        elif strong:
            row0, nrows = resharder.row0, resharder.nrows
            map_ptrs = [o.data_ptr() for o in resharder.local_maps()]
        else:
            row0, nrows = 0, ny
            map_ptrs = [o.data_ptr() for o in outs]
        d_wd = torch.from_numpy(np.ascontiguousarray(wd_full.T[row0:row0 + nrows])).to(dev).reshape(-1)
        cj0, myl = C.c_int(), C.c_int()
        _lib.check(_lib.lib.pfx_l2_fit_rows_shape(ny, row0, nrows, 1, C.byref(cj0), C.byref(myl)))
        nout = max(nx * myl.value, 1)
        fo = [torch.empty(nout, dtype=torch.float64, device=dev) for _ in range(7)]
        st = torch.empty(nout, dtype=torch.int32, device=dev)

        def step_fit():
            if nrows > 0:
                _lib.check(_lib.lib.pfx_l2_fit_rows_dev(
                    local, stream.cuda_stream, C.byref(conf), d_k.data_ptr(), ns, nx, ny, row0, nrows,
                    *map_ptrs, d_wd.data_ptr(), None, None, None, 1, alph, _lib.ATYPES["joint"],
                    fo[0].data_ptr(), fo[1].data_ptr(), fo[2].data_ptr(), fo[3].data_ptr(),
                    fo[4].data_ptr() if alph else None, fo[5].data_ptr() if alph else None, fo[6].data_ptr(),
                    st.data_ptr()))
        step_fit()
        barrier()
        ft = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            e0.record(stream)
            step_fit()
            e1.record(stream)
            e1.synchronize()
            ft.append(e0.elapsed_time(e1))
        tf = torch.tensor(ft, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        cells = (nx - 1) * (ny - 1)
        fit_ms = float(tf.mean().item())
        te_map = fo[0][:nx * myl.value].cpu().numpy().reshape(myl.value, nx)
        raw = st[:nx * myl.value].cpu().numpy().reshape(myl.value, nx)
        stat, nfev = raw & 0xff, raw >> 8  # low byte: status, upper bits: model evaluations of the cell
        fitted = stat > 0
        fit = {"cells_per_s": cells / (fit_ms * 1e-3), "ms": fit_ms, "cells": cells, "atype": "joint",
               "alph": bool(alph), "converged_frac": float((stat == 1).sum() / max(fitted.sum(), 1)),
               "median_Te_km": float(np.median(te_map[fitted])) if fitted.any() else None, "true_Te_km": cfg["Te"],
               "nfev_mean": float(nfev[fitted].mean()) if fitted.any() else None,
               "nfev_max": int(nfev[fitted].max()) if fitted.any() else None,
               "sharding": f"rows over {world} rank(s)" if strong else "none"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    alg_total, alg = alg_bytes_per_unit(cfg)
    # dominant kernel class by device time inside the timed region
    dom = max((c for c in kern if c in alg), key=lambda c: kern[c][0], default=None)
    kernels = {c: {"ms_per_step": kern[c][0] / args.steps, "launches_per_step": kern[c][1] / args.steps,
                   "avg_launch_us": 1e3 * kern[c][0] / kern[c][1]} for c in kern}
    roof = None
    if dom:
        units_launch = nxy * NA  # one launch = one scale, all 11 azimuths of both grids
        avg_s = kern[dom][0] / kern[dom][1] * 1e-3
        achieved = alg[dom] * units_launch / avg_s / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(args.config, dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom] * units_launch, "avg_launch_us": avg_s * 1e6,
                "share_of_step": kern[dom][0] / args.steps / prof_ms, "serialised_step_ms": prof_ms,
                "whole_step": {"achieved": alg_total * units_rank / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_total * units_rank / (ms_per_step * 1e-3) / 1e9 / peak,
                               "bytes_per_unit": alg_total},
                "note": ("algorithmic bytes of a dense 2-pass FP64 FFT path (SURVEY.md 8d); the band-pruned engine "
                         "moves fewer bytes, so frac > 1 is possible; traffic = measured DRAM bytes per launch (ncu)")}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sample = cpu_sample_scales(cfg)
        v, dt, u = cpu_reference(cfg, topo, grav, k, k0, sample, threads)
        cpu = {"value": v, "unit": "units/s", "cores": threads, "kind": "port",
               "sample": f"scales {sample} of {ns} ({u:.3g} units, {dt:.1f} s), float32 oracle port, OpenMP over planes"}
    line = {
        "metric": METRIC, "value": value, "unit": "units/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "engine": args.engine, "l2": "flushed between timed steps (256 MiB write)",
                   "cpu_affinity": f"{numa_cpus} CPUs of the GPU's NUMA node" if numa_cpus else "not set",
                   "balance": args.balance if p2p else None,
                   "sharding": ((f"scales over {world} ranks, epilogue stores into peer row shards over NVLink (CUDA IPC)"
                                 if p2p else f"scales over {world} ranks + NCCL all-to-all to row shards") if strong else
                                (f"{world} independent grid pairs, one per rank" if world > 1 else "single GPU"))},
        "clocks": clocks.summary(),
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roof, "kernels": kernels,
        "cpu_baseline": cpu, "fit": fit,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
