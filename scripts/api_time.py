"""Wall time of the user-level calls (fresh pageable numpy outputs) at a bench configuration."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import bench
import plateflex_b200 as pf
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
topo, grav, k, k0 = bench.workload(cfg)
for dt in (np.float64, np.float32):
    pf.set_output_dtype(dt)
    for rep in range(4):
        proj = pf.Project(grids=[pf.TopoGrid(topo.copy(), cfg["dx"], cfg["dy"]), pf.BougGrid(grav, cfg["dx"], cfg["dy"])])
        proj.init()
        for g in proj.grids:
            g.k, g.ns = k, len(k)
        proj.k, proj.ns = k, len(k)
        t0 = time.perf_counter()
        proj.wlet_admit_coh()
        t1 = time.perf_counter()
        proj.estimate_grid(nn=1)
        t2 = time.perf_counter()
    print(np.dtype(dt).name, "Project.wlet_admit_coh %.1f ms, estimate_grid(nn=1) %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
