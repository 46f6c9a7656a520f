#!/usr/bin/env python
"""Per-call times of the per-cell fit on the maps of one BASELINE configuration (A/B tool for fit experiments):
the transform runs once, then `reps` fits are timed one by one with CUDA events and every sample is printed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plateflex_b200 import _lib, sharding  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = bench.CONFIGS[cfgname]
topo, grav, k, k0 = bench.workload(cfg)
conf = bench.flex_conf(cfg)
pipe = bench._single_pipeline(sharding, cfg["nx"], cfg["ny"], cfg, k, k0, conf, 0, _lib.ENGINE_PRUNED)
wd_f = np.ascontiguousarray((-np.minimum(topo, 0.0)).T)
pipe.load(*[torch.from_numpy(a).pin_memory() for a in (topo, grav, wd_f)])
pipe.run_transform()
stream = torch.cuda.current_stream()
res = pipe.run_fit()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    res = pipe.run_fit()
    e1.record(stream)
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
te = res[0]
print(cfgname, os.environ.get("PLATEFLEX_B200_LIB", "product").split("/")[-1], {k: v for k, v in os.environ.items() if k.startswith("PLATEFLEX_B200_F")},
      "fit ms min %.3f median %.3f" % (min(ts), float(np.median(ts))), ["%.2f" % t for t in ts],
      "median Te %.10f" % float(te[torch.isfinite(te)].median().item()))
