#!/usr/bin/env python
"""The reference's notebook flow (Ex1, Ex2, Ex5/Ex6) on the B200 path.

  python examples/quickstart.py [topo.xyz grav.xyz nx ny dx_km dy_km]

Without arguments a synthetic 512 x 512 land Bouguer pair (true Te = 40 km, F = 0.5) is used -- the
reference's own example grids are not distributed with its source tree.  With arguments the two
files are read the way the notebooks do (docs/getting_started.rst:50-56, Ex1 cell 2: three columns
x, y, value; the value column reshaped to (nx, ny) and flipped along the first axis).
"""
import sys
import time

import numpy as np

import plateflex_b200 as plateflex   # the only line that differs from the reference's notebooks
from plateflex_b200 import synthetic


def read_xyz(path, nx, ny):
    import pandas as pd
    return pd.read_csv(path, sep='\t', header=None)[2].values.reshape(nx, ny)[::-1]


def main():
    if len(sys.argv) == 7:
        nx, ny, dx, dy = int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), float(sys.argv[6])
        topodata, gravdata = read_xyz(sys.argv[1], nx, ny), read_xyz(sys.argv[2], nx, ny)
    else:
        nx = ny = 512
        dx = dy = 5.0
        topodata, gravdata = synthetic.make_pair(nx, ny, dx, dy, seed=1, Te=40.0, F=0.5)

    topo = plateflex.TopoGrid(topodata, dx, dy)
    boug = plateflex.BougGrid(gravdata, dx, dy)
    project = plateflex.Project(grids=[topo, boug])
    project.init()

    t0 = time.perf_counter()
    project.wlet_admit_coh()                       # CWT of both grids -> admittance, coherence, jackknife errors
    t1 = time.perf_counter()
    project.estimate_grid(nn=1, atype='joint')     # (Te, F) for every cell
    t2 = time.perf_counter()
    print(f"{nx} x {ny} grid, {project.ns} scales: wavelet analysis {1e3 * (t1 - t0):.1f} ms, "
          f"Te / F map {1e3 * (t2 - t1):.1f} ms")
    fitted = project.status_grid == 1
    print(f"median Te = {np.median(project.mean_Te_grid[fitted]):.2f} km, "
          f"median F = {np.median(project.mean_F_grid[fitted]):.3f}")
    project.save_results(mean_Te=True, std_Te=True, mean_F=True, chi2=True, filter=True, sigma=1,
                         prefix='PlateFlex_')


if __name__ == '__main__':
    main()
