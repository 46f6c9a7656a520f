"""Generates the committed fixtures in tests/golden/ from the CPU oracle (float64).

The reference package cannot be imported or built in this image (Fortran sources, no Fortran
compiler), so these vectors pin the ORACLE and the CUDA path against regressions;
the only vectors that come from the reference itself are the _lam2k known answers in
lam2k_kat.json (classes.py:1787-1796 docstring and Ex1_making_grids.ipynb cell 12 output).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import l2fit, oracle as ora  # noqa: E402
from plateflex_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # --- CWT -> scalogram / admittance / coherence on a small non-power-of-two grid
    nx, ny, dx, dy = 44, 37, 20.0, 18.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240001, Te=25.0, F=0.4)
    k = np.array([1.2e-5, 2.5e-5, 5.0e-5, 9.0e-5])
    k0 = ora.K0_F32
    w1 = ora.wlet_transform(topo, dx, dy, k, k0=k0)
    w2 = ora.wlet_transform(grav, dx, dy, k, k0=k0)
    adm, eadm, coh, ecoh = ora.wlet_admit_coh(w1, w2)
    sg, esg = ora.wlet_scalogram(w1)
    xsg, exsg = ora.wlet_xscalogram(w1, w2)
    ia_sel = np.array([0, 5, 10])
    np.savez_compressed(os.path.join(HERE, "cwt_small.npz"), topo=topo, grav=grav, dx=dx, dy=dy, k=k, k0=k0,
                        ia_sel=ia_sel, wl_trans_topo_sel=w1[:, :, ia_sel, :], wl_admit=adm, ewl_admit=eadm,
                        wl_coh=coh, ewl_coh=ecoh, wl_sg=sg, ewl_sg=esg, wl_xsg=xsg, ewl_xsg=exsg)

    # --- flexural model curves
    kk = np.geomspace(2e-6, 2e-4, 24)
    rng = np.random.default_rng(5)
    n = 16
    te, f, al = rng.uniform(2, 250, n), rng.uniform(0.001, 0.99, n), rng.uniform(0.01, np.pi - 0.01, n)
    wd = np.where(np.arange(n) % 2 == 0, 0.0, rng.uniform(100, 5000, n))
    rows_a, rows_c = [], []
    for c in range(n):
        conf = ora.FlexConf(wd=float(wd[c]), boug=0, rhoc=2800.0, zc=30.0e3)
        a, co = ora.real_xspec_functions(kk, te[c], f[c], al[c], conf)
        rows_a.append(a)
        rows_c.append(co)
    np.savez_compressed(os.path.join(HERE, "flex_curves.npz"), k=kk, te=te, f=f, alpha=al, wd=wd, boug=0, rhoc=2800.0,
                        zc=30.0e3, admit=np.array(rows_a), coh=np.array(rows_c))

    # --- per-cell fits by SciPy TRF on the oracle model (scipy version recorded)
    kfit = np.geomspace(3e-6, 1.2e-4, 20)
    cells = []
    rng = np.random.default_rng(11)
    for c in range(24):
        te_t, f_t = rng.uniform(5, 120), rng.uniform(0.1, 0.8)
        conf = ora.FlexConf()
        a, co = ora.real_xspec_functions(kfit, te_t, f_t, np.pi / 2, conf)
        ea = 0.05 * np.abs(a) + 2e-3
        ec = 0.05 + 0.0 * co
        an = a + ea * rng.standard_normal(len(kfit)) * 0.5
        cn = np.clip(co + ec * rng.standard_normal(len(kfit)) * 0.5, 1e-6, 1.0)
        row = dict(adm=an, eadm=ea, coh=cn, ecoh=ec)
        for atype in ("joint", "admit", "coh"):
            for alph in (False, True):
                r = l2fit.L2_estimate_cell(kfit, an, ea, cn, ec, alph=alph, atype=atype, conf=ora.FlexConf())
                tag = f"{atype}_{int(alph)}"
                row[tag + "_mean"] = np.pad(r["mean"], (0, 3 - len(r["mean"])))
                row[tag + "_std"] = np.pad(r["std"], (0, 3 - len(r["std"])))
                row[tag + "_chi2"] = r["chi2"]
        cells.append(row)
    out = {key: np.array([c[key] for c in cells]) for key in cells[0]}
    np.savez_compressed(os.path.join(HERE, "l2fit_cells.npz"), k=kfit, scipy_version=l2fit.scipy_version, **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
