#!/usr/bin/env python
"""Step and per-kernel-class times of the CWT -> admittance/coherence stage for one configuration, with whatever
PLATEFLEX_B200_* environment knobs are set: the A/B tool for kernel experiments (not a benchmark line)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plateflex_b200 import _lib  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = bench.CONFIGS[cfgname]
rng = np.random.default_rng(1)
nx, ny, ns = cfg["nx"], cfg["ny"], cfg["ns"]
_, _, k, k0 = bench.workload(cfg, data=False)
dev = torch.device("cuda", 0)
d_t = torch.from_numpy(rng.standard_normal((nx, ny))).to(dev)
d_g = torch.from_numpy(rng.standard_normal((nx, ny))).to(dev)
plan = _lib.Plan(nx, ny, cfg["dx"], cfg["dy"], k, k0, device=0)
stream = torch.cuda.current_stream()
plan.set_stream(stream.cuda_stream)
outs = [torch.empty(ns * nx * ny, dtype=torch.float64, device=dev) for _ in range(4)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def step():
    _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, d_t.data_ptr(), d_g.data_ptr(), 0, ns, *[o.data_ptr() for o in outs]))


for _ in range(2):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step()
    e1.record(stream)
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
plan.profile(True)
plan.profile_read()
for _ in range(steps):
    step()
torch.cuda.synchronize()
kern = plan.profile_read()
knobs = {k: v for k, v in os.environ.items() if k.startswith("PLATEFLEX_B200_")}
print(cfgname, knobs, "step %.3f ms (min %.3f)" % (np.mean(ts), np.min(ts)),
      {c: round(v[0] / steps, 3) for c, v in kern.items()}, "checksum %.12e" % float(outs[0][::997].nan_to_num().sum().item()))
