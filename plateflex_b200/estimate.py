"""L2 estimation of (Te, F[, alpha]) -- the GPU counterpart of the reference's
plateflex/estimate.py:227-474 (``L2_estimate_cell``, ``get_L2_estimates``,
``real_xspec_functions``).

Where the reference calls scipy.optimize.curve_fit once per cell from a Python
loop, ``L2_estimate_grid`` fits every requested cell of a grid in one kernel
launch (pfx_l2_fit_host / pfx_l2_fit_dev); ``L2_estimate_cell`` is the same
kernel on a single cell and returns the reference's summary DataFrame.  The
Bayesian path (estimate.py:65-224) is out of scope and stays the reference's.
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .flex import conf_flex, flex

__all__ = ["L2_estimate_cell", "L2_estimate_grid", "L2_estimate_grid_device", "get_L2_estimates", "real_xspec_functions"]

STATUS_TEXT = {0: "not fitted", 1: "converged", 2: "max_nfev reached", 3: "non-finite data or zero sigma"}


def _check_args(alph, atype):
    if not isinstance(alph, bool):
        raise Exception("'alph' should be a boolean: defaults to False")  # classes.py:1070-1073
    if atype not in _lib.ATYPES:
        raise Exception("'atype' should be one among: 'admit', 'coh', or 'joint'")  # classes.py:1075-1078


def L2_estimate_grid(k, wl_admit, ewl_admit, wl_coh, ewl_coh, wd, rhoc=None, zc=None, mask=None, nn=1, alph=False,
                     atype="joint", device=None, strict=False):
    """Fit all cells ``(i, j)`` with i in range(0, nx-nn, nn), j in range(0, ny-nn, nn)
    (classes.py:1218-1219).  Inputs are (nx, ny, ns) maps and (nx, ny) grids; returns a dict of
    (nx//nn, ny//nn) maps: mean_Te, std_Te, mean_F, std_F, chi2[, mean_a, std_a], status, nfev.

    ``strict=True`` raises like SciPy would for the first cell that did not converge
    (RuntimeError) or holds unusable data (ValueError); otherwise such cells are flagged
    in ``status`` (2 / 3) and, for bad data, set to NaN.
    """
    _check_args(alph, atype)
    k = np.ascontiguousarray(k, dtype=np.float64).ravel()
    maps = [np.asfortranarray(a, dtype=np.float64) for a in (wl_admit, ewl_admit, wl_coh, ewl_coh)]
    nx, ny, ns = maps[0].shape
    if any(m.shape != (nx, ny, ns) for m in maps) or ns != len(k):
        raise ValueError("admittance / coherence maps must all have shape (nx, ny, len(k))")

    def grid2(a, name, dtype=np.float64):
        if a is None:
            return None
        a = np.asfortranarray(a, dtype=dtype)
        if a.shape != (nx, ny):
            raise ValueError(f"{name} must have shape ({nx}, {ny})")
        return a

    wd_a, rhoc_a, zc_a = grid2(wd, "wd"), grid2(rhoc, "rhoc"), grid2(zc, "zc")
    mask_a = grid2(mask, "mask", np.uint8) if mask is not None else None
    mx, my = nx // nn, ny // nn
    names = ["mean_Te", "std_Te", "mean_F", "std_F", "mean_a", "std_a", "chi2"]
    out = {n: np.zeros((mx, my), order="F") for n in names if alph or not n.endswith("_a")}
    status = np.zeros((mx, my), dtype=np.int32, order="F")
    conf = conf_flex.as_struct()
    dev = _lib.default_device() if device is None else device
    _lib.check(_lib.lib.pfx_l2_fit_host(
        dev, C.byref(conf), _lib.ptr(k), ns, nx, ny, *[_lib.ptr(m) for m in maps], _lib.ptr(wd_a), _lib.ptr(rhoc_a),
        _lib.ptr(zc_a), _lib.ptr(mask_a), int(nn), int(alph), _lib.ATYPES[atype],
        *[_lib.ptr(out.get(n)) for n in names], _lib.ptr(status)))
    out["status"] = status & 0xff
    out["nfev"] = status >> 8  # model evaluations per cell (least_squares' nfev)
    _report(out["status"], strict)
    return out


def L2_estimate_grid_device(k, dmaps, wd, rhoc=None, zc=None, mask=None, nn=1, alph=False, atype="joint",
                            strict=False):
    """``L2_estimate_grid`` on maps that are already on the GPU (a cpwt.DeviceMaps, what
    Project.wlet_admit_coh leaves behind): only wd / rhoc / zc / mask go up and only the
    (nx//nn, ny//nn) result maps come back."""
    _check_args(alph, atype)
    k = np.ascontiguousarray(k, dtype=np.float64).ravel()
    nx, ny, ns, dev = dmaps.nx, dmaps.ny, dmaps.ns, dmaps.device
    if ns != len(k):
        raise ValueError("admittance / coherence maps must all have shape (nx, ny, len(k))")

    def up(a, name, dtype=np.float64):
        if a is None:
            return None
        a = np.asarray(a, dtype=dtype)
        if a.shape != (nx, ny):
            raise ValueError(f"{name} must have shape ({nx}, {ny})")
        return _lib.DeviceBuffer.from_host_as_fortran(a, dev)

    mx, my = nx // nn, ny // nn
    names = ["mean_Te", "std_Te", "mean_F", "std_F", "mean_a", "std_a", "chi2"]
    want = [n for n in names if alph or not n.endswith("_a")]
    bufs = {"k": _lib.DeviceBuffer.from_host(k, dev), "wd": up(wd, "wd"), "rhoc": up(rhoc, "rhoc"), "zc": up(zc, "zc"),
            "mask": up(mask, "mask", np.uint8) if mask is not None else None,
            "out": _lib.DeviceBuffer(max(mx * my, 1) * 8 * len(want), dev), "status": _lib.DeviceBuffer(max(mx * my, 1) * 4, dev)}
    try:
        p = lambda b: C.c_void_p(b.ptr) if b is not None else None
        optr = {n: C.c_void_p(bufs["out"].ptr + i * mx * my * 8) for i, n in enumerate(want)}
        conf = conf_flex.as_struct()
        _lib.check(_lib.lib.pfx_l2_fit_dev(
            dev, None, C.byref(conf), p(bufs["k"]), ns, nx, ny, *[C.c_void_p(q) for q in dmaps.ptrs()], p(bufs["wd"]),
            p(bufs["rhoc"]), p(bufs["zc"]), p(bufs["mask"]), int(nn), int(alph), _lib.ATYPES[atype],
            *[optr.get(n) for n in names], p(bufs["status"])))
        out = {}
        for i, n in enumerate(want):
            out[n] = bufs["out"].download(np.zeros((mx, my), order="F"), i * mx * my * 8) if mx * my else np.zeros((mx, my), order="F")
        status = np.zeros((mx, my), dtype=np.int32, order="F")
        if mx * my:
            bufs["status"].download(status)
    finally:
        for b in bufs.values():
            if b is not None:
                b.close()
    out["status"] = status & 0xff
    out["nfev"] = status >> 8
    _report(out["status"], strict)
    return out


def _report(status, strict):
    bad, slow = int((status == 3).sum()), int((status == 2).sum())
    if strict and bad:
        raise ValueError("array must not contain infs or NaNs / sigma must be non-zero "
                         f"({bad} cell(s) with unusable data)")
    if strict and slow:
        raise RuntimeError("Optimal parameters not found: The maximum number of function evaluations is exceeded.")
    if bad or slow:
        warnings.warn(f"L2 fit: {bad} cell(s) with unusable data, {slow} cell(s) stopped at max_nfev", RuntimeWarning)


def L2_estimate_cell(k, adm, eadm, coh, ecoh, alph=False, atype="joint"):
    """estimate.py:227-409 for one cell.  Model constants (rhoc, zc, wd, boug, ...) come from
    ``conf_flex`` exactly as in the reference, where Project.estimate_cell writes wd / rhoc / zc
    into the Fortran module before calling (classes.py:1087-1091).  Returns the summary
    DataFrame with rows Te, F[, alpha] and columns mean, std, chi2."""
    import pandas as pd

    _check_args(alph, atype)
    ns = len(k)
    # the grid kernel never fits the last row / column (classes.py:1218-1219), so embed the cell in 2 x 2
    def embed(v):
        a = np.ones((2, 2, ns), order="F")
        a[0, 0, :] = np.asarray(v, dtype=np.float64)
        return a

    wd = np.full((2, 2), float(conf_flex.wd), order="F")
    res = L2_estimate_grid(k, embed(adm), embed(eadm), embed(coh), embed(ecoh), wd, nn=1, alph=alph, atype=atype,
                           strict=True)
    conf_flex.rhof = conf_flex.rhow if conf_flex.wd > 0.0 else conf_flex.rhoa  # flex.f90:207-211 side effect
    rows = ["Te", "F"] + (["alpha"] if alph else [])
    mean = [res["mean_Te"][0, 0], res["mean_F"][0, 0]] + ([res["mean_a"][0, 0]] if alph else [])
    std = [res["std_Te"][0, 0], res["std_F"][0, 0]] + ([res["std_a"][0, 0]] if alph else [])
    chi2 = res["chi2"][0, 0]
    return pd.DataFrame(data={"mean": mean, "std": std, "chi2": [chi2] * len(rows)}, index=rows)


def get_L2_estimates(summary):
    """estimate.py:412-448: (mean_te, std_te, mean_F, std_F[, mean_a, std_a], rchi2)."""
    mean_te, std_te, rchi2 = summary.loc["Te", "mean"], summary.loc["Te", "std"], summary.loc["Te", "chi2"]
    mean_F, std_F = summary.loc["F", "mean"], summary.loc["F", "std"]
    if "alpha" in summary.index:
        return mean_te, std_te, mean_F, std_F, summary.loc["alpha", "mean"], summary.loc["alpha", "std"], rchi2
    return mean_te, std_te, mean_F, std_F, rchi2


def real_xspec_functions(k, Te, F, alpha=np.pi / 2.0):
    """estimate.py:451-474: predicted (admittance, coherence) on wavenumbers ``k`` (rad/m)."""
    return flex.real_xspec_functions(k, Te, F, alpha)
