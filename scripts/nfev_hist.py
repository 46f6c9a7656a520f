"""Distribution of the number of trust-region trials per cell (status >> 8) for a bench configuration."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import plateflex_b200 as pf
from plateflex_b200 import estimate
from plateflex_b200.cpwt import fused
cfgname = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = bench.CONFIGS[cfgname]
topo, grav, k, k0 = bench.workload(cfg)
pf.cf_fl.boug = 1 if cfg["boug"] else 0
if cfg["ocean"]:
    pf.cf_fl.rhoc = 2800.
maps = fused.wlet_admit_coh(topo, grav, cfg["dx"], cfg["dy"], k)
wd = -np.minimum(topo, 0.0)
res = estimate.L2_estimate_grid(k, *maps, wd, nn=1, alph=cfg["alph"], atype="joint")
nf = res["nfev"][res["status"] == 1]
print(cfgname, "cells", nf.size, "mean", nf.mean(), "max", nf.max())
for t in (12, 16, 24, 32, 48, 64, 100, 200, 400):
    print("  nfev >", t, ":", int((nf > t).sum()), "cells  (%.4f %%)" % (100.0 * (nf > t).mean()), " trials in them: %.2f %%" % (100.0 * nf[nf > t].sum() / nf.sum()))
print("  percentiles 50/90/99/99.9:", np.percentile(nf, [50, 90, 99, 99.9]))
