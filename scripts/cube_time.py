import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import plateflex_b200 as pf
from plateflex_b200 import synthetic, classes
from plateflex_b200.cpwt import cpwt
topo, grav = synthetic.make_pair(1024, 1024, 5., 5., seed=1)
k = classes._lam2k(1024, 1024, 5., 5.)[1][:20]
for dt in (np.float64, np.float32):
    pf.set_output_dtype(dt)
    for rep in range(3):
        t = time.perf_counter(); w = cpwt.wlet_transform(topo, 5., 5., k); e = time.perf_counter() - t
    print(dt.__name__, w.dtype, w.nbytes / 1e9, "GB", round(e * 1e3, 1), "ms", round(w.nbytes / e / 1e9, 1), "GB/s")
