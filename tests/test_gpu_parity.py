"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the float64 CPU oracle on the same seeded inputs, against the committed golden
fixtures, and through size-independent properties at BASELINE sizes."""
import os

import numpy as np
import pytest

from helpers import A_TOL, CHI2_RTOL, F_TOL, STD_RTOL, TE_TOL, assert_maps_close, red_field

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ENGINES = ["dense", "pruned"]


@pytest.fixture(params=ENGINES)
def engine(request, monkeypatch, pf):
    from plateflex_b200 import _lib
    monkeypatch.setenv("PLATEFLEX_B200_ENGINE", request.param)
    _lib.clear_plans()
    yield request.param
    _lib.clear_plans()


CASES = [  # nx, ny, dx, dy, wavenumbers (rad/m)
    (32, 32, 5.0, 5.0, [2e-5, 8e-5, 3e-4]),
    (44, 37, 20.0, 18.0, [1.2e-5, 2.5e-5, 5.0e-5, 9.0e-5]),
    (17, 50, 20.0, 7.0, [1.5e-5, 6e-5]),
    (64, 48, 10.0, 10.0, None),  # natural wavenumber ladder of the grid (classes._lam2k)
    (2, 3, 10.0, 10.0, [5e-5]),
    # padded axes of 2048 / 4096 / 8192 samples: the second-generation pruned kernels (pfx_pruned_v2.cuh),
    # one part per line, two and four parts per line; natural ladders reach every FFT length up to N
    (520, 24, 5.0, 5.0, None),
    (24, 520, 5.0, 5.0, None),
    (1030, 20, 5.0, 5.0, ("ladder", 2)),
    (18, 2050, 5.0, 5.0, ("ladder", 4)),
    # power-of-two grids whose padded y axis reaches 2048 / 4096: spectra and the admissibility plane through
    # cosine transforms of the unmirrored grid (pfx_dct.cu), one and two parts per line
    (16, 1024, 5.0, 5.0, ("ladder", 3)),
    (8, 2048, 5.0, 5.0, ("ladder", 4)),
]


def _wavenumbers(kf, nx, ny, dx, dy):
    """Explicit list, the grid's natural ladder (None), or every n-th scale of it (("ladder", n))."""
    from plateflex_b200 import classes
    if kf is None:
        return classes._lam2k(nx, ny, dx, dy)[1]
    if isinstance(kf, tuple):
        return classes._lam2k(nx, ny, dx, dy)[1][::-1][::kf[1]][::-1].copy()
    return np.array(kf)


@pytest.mark.parametrize("nx,ny,dx,dy,kf", CASES)
def test_wlet_transform_matches_oracle(pf, ora, engine, nx, ny, dx, dy, kf):
    from plateflex_b200 import classes
    from plateflex_b200.cpwt import cpwt
    g = red_field(nx, ny, seed=nx * 100 + ny) * 400.0 + 37.0
    kf = _wavenumbers(kf, nx, ny, dx, dy)
    got = cpwt.wlet_transform(g, dx, dy, kf)
    ref = ora.wlet_transform(g, dx, dy, kf, k0=pf.cf_wt.k0, prec=64)
    assert got.shape == ref.shape == (nx, ny, 11, len(kf)) and got.dtype == np.complex128
    assert got.flags.f_contiguous
    # per scale, all 11 azimuth planes together: max|d| / max|ref| <= 1e-9 and element-wise
    flat = lambda w: w.reshape(nx * ny * 11, len(kf), order="F")  # noqa: E731
    for part in (np.real, np.imag):
        assert_maps_close(part(flat(got)), part(flat(ref)), "wl_trans", rtol=1e-9, atol_rel=1e-11)


@pytest.mark.parametrize("nx,ny,dx,dy,kf", CASES[:4] + CASES[5:7])
def test_fused_reductions_match_oracle(pf, ora, engine, nx, ny, dx, dy, kf):
    from plateflex_b200 import classes
    from plateflex_b200.cpwt import fused
    t = red_field(nx, ny, seed=1) * 500.0
    g = 0.08 * t + red_field(nx, ny, seed=2) * 20.0
    kf = _wavenumbers(kf, nx, ny, dx, dy)
    w1 = ora.wlet_transform(t, dx, dy, kf, k0=pf.cf_wt.k0)
    w2 = ora.wlet_transform(g, dx, dy, kf, k0=pf.cf_wt.k0)
    for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), fused.wlet_admit_coh(t, g, dx, dy, kf),
                          ora.wlet_admit_coh(w1, w2)):
        assert_maps_close(a, b, name)
    for name, a, b in zip(("wl_sg", "ewl_sg"), fused.wlet_scalogram(t, dx, dy, kf), ora.wlet_scalogram(w1)):
        assert_maps_close(a, b, name)
    xs, exs = fused.wlet_xscalogram(t, g, dx, dy, kf)
    rxs, rexs = ora.wlet_xscalogram(w1, w2)
    assert_maps_close(xs.real, rxs.real, "wl_xsg.real", atol_rel=1e-11)
    assert_maps_close(xs.imag, rxs.imag, "wl_xsg.imag", atol_rel=1e-11)
    assert_maps_close(exs, rexs, "ewl_xsg")


@pytest.mark.parametrize("st,sg", [(1e-80, 1.0), (1.0, 1e-90), (1e-74, 1e-74), (1e-78, 1e-78)])
def test_extreme_amplitudes_keep_the_reference_guard(pf, ora, st, sg):
    """cpwt_sub.f90:329: the jackknife's guard is (p1 /= 0 .and. p2 /= 0), not "p1*p2 /= 0", and ad = xp/p1 does not
    involve p2.  With grids so small that the product of the two running sums leaves the normal float64 range the
    epilogue switches, per cell, to the reference's guard and IEEE divisions formed without that product.  The first
    two cases also pin that the two grids never share a transform (their amplitudes differ by 80-90 orders of
    magnitude; a packed complex transform would drown the small one in the rounding error of the large one).  In
    the last two cases the single-azimuth powers straddle / lie below 1e-150 and their products reach the subnormal
    range; the admittance maps and the coherence error stay well defined in the reference and are compared, the mean
    coherence (cpwt.f90:355 divides by the product of two powers near the underflow threshold) is not."""
    from plateflex_b200.cpwt import fused
    nx, ny, dx, dy = 40, 36, 10.0, 10.0
    t = red_field(nx, ny, seed=5) * 500.0 * st
    g = (0.08 * red_field(nx, ny, seed=5) * 500.0 + red_field(nx, ny, seed=6) * 20.0) * sg
    kf = np.array([1.5e-5, 4e-5, 1.2e-4])
    w1 = ora.wlet_transform(t, dx, dy, kf, k0=pf.cf_wt.k0)
    w2 = ora.wlet_transform(g, dx, dy, kf, k0=pf.cf_wt.k0)
    with np.errstate(all="ignore"):
        ref = ora.wlet_admit_coh(w1, w2)
    got = fused.wlet_admit_coh(t, g, dx, dy, kf)
    names = ("wl_admit", "ewl_admit", "wl_coh", "ewl_coh")
    for name, a, b in zip(names, got, ref):
        if st == sg and "coh" in name:
            continue
        assert np.isfinite(b).all(), name
        assert_maps_close(a, b, name)
    if st == sg:  # coherence: the jackknife error stays a ratio of like powers and must agree as well
        assert np.isfinite(ref[3]).all()
        assert_maps_close(got[3], ref[3], "ewl_coh")


def test_scale_subrange(pf, ora, engine):
    from plateflex_b200.cpwt import fused
    t, g = red_field(40, 40, 3) * 500, red_field(40, 40, 4) * 30
    kf = np.array([1e-5, 2e-5, 4e-5, 8e-5, 1.6e-4])
    full = fused.wlet_admit_coh(t, g, 10., 10., kf)
    part = fused.wlet_admit_coh(t, g, 10., 10., kf, scales=(1, 4))
    for a, b in zip(full, part):
        assert b.shape == (40, 40, 3) and np.array_equal(a[..., 1:4], b)


def test_cube_input_forms_match_oracle(pf, ora):
    from plateflex_b200.cpwt import cpwt
    rng = np.random.default_rng(9)
    shp = (13, 9, 11, 3)
    w1 = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
    w2 = 0.4 * w1 + 0.2 * (rng.standard_normal(shp) + 1j * rng.standard_normal(shp))
    for a, b in zip(cpwt.wlet_admit_coh(w1, w2), ora.wlet_admit_coh(w1, w2)):
        assert a.flags.f_contiguous and np.allclose(a, b, rtol=1e-12, atol=1e-14)
    for a, b in zip(cpwt.wlet_scalogram(w1), ora.wlet_scalogram(w1)):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)
    for a, b in zip(cpwt.wlet_xscalogram(w1, w2), ora.wlet_xscalogram(w1, w2)):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        cpwt.wlet_scalogram(w1[:, :, :5, :])


def test_golden_cwt_fixture(pf, engine):
    from plateflex_b200.cpwt import cpwt, fused
    z = np.load(os.path.join(GOLD, "cwt_small.npz"))
    pf.cf_wt.k0 = float(z["k0"])
    dx, dy = float(z["dx"]), float(z["dy"])
    w = cpwt.wlet_transform(z["topo"], dx, dy, z["k"])
    sel = w[:, :, z["ia_sel"], :]
    assert np.abs(sel - z["wl_trans_topo_sel"]).max() <= 1e-9 * np.abs(z["wl_trans_topo_sel"]).max()
    for name, got in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"),
                         fused.wlet_admit_coh(z["topo"], z["grav"], dx, dy, z["k"])):
        assert_maps_close(got, z[name], name)
    for name, got in zip(("wl_sg", "ewl_sg"), fused.wlet_scalogram(z["topo"], dx, dy, z["k"])):
        assert_maps_close(got, z[name], name)
    pf.set_conf_cpwt()


def test_zero_grid_gives_nan_admittance_and_zero_errors(pf, engine):
    from plateflex_b200.cpwt import fused
    z = np.zeros((16, 16))
    adm, eadm, coh, ecoh = fused.wlet_admit_coh(z, z, 10., 10., np.array([5e-5]))
    assert np.isnan(adm).all() and np.isnan(coh).all() and (eadm == 0).all() and (ecoh == 0).all()


def test_flex_model_matches_oracle(pf, ora):
    from plateflex_b200.flex import conf_flex, flex
    f = np.load(os.path.join(GOLD, "flex_curves.npz"))
    pf.set_conf_flex()
    conf_flex.boug, conf_flex.rhoc, conf_flex.zc = int(f["boug"]), float(f["rhoc"]), float(f["zc"])
    adm, coh, ja, jc = flex.real_xspec_batch(f["k"], f["te"], f["f"], f["alpha"], f["wd"], jac=True)
    assert np.allclose(adm, f["admit"], rtol=1e-12, atol=0) and np.allclose(coh, f["coh"], rtol=1e-12, atol=1e-300)
    # the single-curve f2py signature, with the rhof side effect of flex.f90:207-211
    conf_flex.wd = float(f["wd"][1])
    a1, c1 = flex.real_xspec_functions(f["k"], f["te"][1], f["f"][1], f["alpha"][1])
    assert np.allclose(a1, f["admit"][1], rtol=1e-12) and conf_flex.rhof == conf_flex.rhow
    # analytic Jacobian against central differences of the oracle
    for c in (0, 3, 7):
        conf = ora.FlexConf(wd=float(f["wd"][c]), boug=int(f["boug"]), rhoc=float(f["rhoc"]), zc=float(f["zc"]))
        p = np.array([f["te"][c], f["f"][c], f["alpha"][c]])
        for j in range(3):
            h = 1e-6 * max(1.0, abs(p[j]))
            pp, pm = p.copy(), p.copy()
            pp[j] += h
            pm[j] -= h
            ap, cp = ora.real_xspec_functions(f["k"], *pp, conf)
            am, cm = ora.real_xspec_functions(f["k"], *pm, conf)
            da, dc = (ap - am) / (2 * h), (cp - cm) / (2 * h)
            assert np.allclose(ja[c, j], da, rtol=2e-6, atol=1e-7 * np.abs(da).max() + 1e-300)
            assert np.allclose(jc[c, j], dc, rtol=2e-6, atol=1e-7 * np.abs(dc).max() + 1e-300)
    pf.set_conf_flex()


def _compare_fit(got_mean, got_std, got_chi2, ref_mean, ref_std, ref_chi2, npar):
    lo, hi = np.array([2., 1e-4, 1e-4])[:npar], np.array([250., 0.9999, np.pi - 1e-3])[:npar]
    near = lambda m: np.any((np.asarray(m)[:npar] - lo < 1e-3 * (hi - lo)) | (hi - np.asarray(m)[:npar] < 1e-3 * (hi - lo)))  # noqa: E731
    # a cell is a "bound cell" when either optimiser ends on a bound: in the flat valleys that lead there (std of
    # Te several times Te) the two stop, both within ftol, at points whose chi2 differ by a few 1e-4
    on_bound = near(ref_mean) or near(got_mean)
    assert got_chi2 == pytest.approx(ref_chi2, rel=CHI2_RTOL if not on_bound else 1e-3)
    if on_bound:
        return False
    tol = [TE_TOL, F_TOL, A_TOL][:npar]
    # Interior optimum.  Where a parameter is poorly determined (its standard deviation many times the tolerance:
    # a flat valley of chi2, typical of Te -- and of F along with it -- at low coherence) two optimisers that both
    # stop within ftol = 1e-8 stop at different points of the valley: there chi2 is the criterion (checked above to
    # 1e-4) and the parameters must merely lie within one standard deviation of each other (the larger of the two
    # optimisers' sigmas; they vary along the valley).  Cells whose parameters agree to the plain tolerances are
    # "sharp": their standard deviations are compared too, and the callers require at least half of a sample to be
    # sharp.
    sharp = True
    for j in range(npar):
        d = abs(got_mean[j] - ref_mean[j])
        assert d < max(tol[j], max(ref_std[j], got_std[j])), (j, got_mean, ref_mean, ref_std, got_std)
        sharp = sharp and d < tol[j]
    if sharp:
        for j in range(npar):
            assert got_std[j] == pytest.approx(ref_std[j], rel=STD_RTOL), (j, got_std, ref_std)
    return sharp


@pytest.mark.parametrize("atype", ["joint", "admit", "coh"])
@pytest.mark.parametrize("alph", [False, True])
def test_fit_matches_scipy_trf_golden(pf, atype, alph):
    """Golden vectors: SciPy curve_fit (TRF) on the oracle model, tests/golden/make_golden.py."""
    from plateflex_b200 import estimate
    z = np.load(os.path.join(GOLD, "l2fit_cells.npz"))
    pf.set_conf_flex()
    n, ns = z["adm"].shape
    # all cells at once: lay them along x of an (n+1, 2, ns) grid (last row/column is never fitted)
    def lay(a):
        m = np.ones((n + 1, 2, ns), order="F")
        m[:n, 0, :] = a
        return m
    res = estimate.L2_estimate_grid(z["k"], lay(z["adm"]), lay(z["eadm"]), lay(z["coh"]), lay(z["ecoh"]),
                                    wd=np.zeros((n + 1, 2)), nn=1, alph=alph, atype=atype)
    assert (res["status"][:n, 0] == 1).all() and (res["status"][n:, :] == 0).all()
    tag = f"{atype}_{int(alph)}"
    npar = 3 if alph else 2
    interior = 0
    for c in range(n):
        gm = [res["mean_Te"][c, 0], res["mean_F"][c, 0]] + ([res["mean_a"][c, 0]] if alph else [])
        gs = [res["std_Te"][c, 0], res["std_F"][c, 0]] + ([res["std_a"][c, 0]] if alph else [])
        interior += _compare_fit(gm, gs, res["chi2"][c, 0], z[tag + "_mean"][c], z[tag + "_std"][c],
                                 float(z[tag + "_chi2"][c]), npar)
    assert interior >= n // 2
    # the single-cell entry point returns the reference's summary frame
    s = estimate.L2_estimate_cell(z["k"], z["adm"][0], z["eadm"][0], z["coh"][0], z["ecoh"][0], alph, atype)
    assert list(s.index) == ["Te", "F"] + (["alpha"] if alph else []) and list(s.columns) == ["mean", "std", "chi2"]
    assert s.loc["Te", "mean"] == res["mean_Te"][0, 0]
    assert len(estimate.get_L2_estimates(s)) == (7 if alph else 5)


def test_fit_flags_bad_cells(pf):
    from plateflex_b200 import estimate
    z = np.load(os.path.join(GOLD, "l2fit_cells.npz"))
    ns = len(z["k"])
    def lay(v):
        m = np.ones((3, 2, ns), order="F")
        m[0, 0], m[1, 0] = v, v
        return m
    eadm = lay(z["eadm"][0])
    eadm[1, 0, 3] = 0.0  # zero sigma: SciPy would raise (estimate.py:274)
    with pytest.warns(RuntimeWarning):
        res = estimate.L2_estimate_grid(z["k"], lay(z["adm"][0]), eadm, lay(z["coh"][0]), lay(z["ecoh"][0]),
                                        wd=np.zeros((3, 2)))
    assert res["status"][0, 0] == 1 and res["status"][1, 0] == 3 and np.isnan(res["mean_Te"][1, 0])
    with pytest.raises(ValueError):
        estimate.L2_estimate_grid(z["k"], lay(z["adm"][0]), eadm, lay(z["coh"][0]), lay(z["ecoh"][0]),
                                  wd=np.zeros((3, 2)), strict=True)


def test_project_end_to_end_matches_oracle_pipeline(pf, ora, engine):
    """Grid -> Project.init -> wlet_admit_coh -> estimate_grid against oracle CWT + SciPy TRF."""
    from oracle import l2fit
    from plateflex_b200 import synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    nx, ny, dx, dy = 48, 40, 20.0, 20.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240002, Te=30.0, F=0.5)
    t, b = pf.TopoGrid(topo.copy(), dx, dy), pf.BougGrid(grav, dx, dy)
    mask = np.zeros((nx, ny), dtype=bool)
    mask[8:16, 8:16] = True
    proj = pf.Project(grids=[t, b])
    proj.mask = mask
    proj.init()
    proj.wlet_admit_coh()
    ref = ora.admit_coh_from_grids(t.data, b.data, dx, dy, proj.k, k0=pf.cf_wt.k0)
    for name, r in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), ref):
        assert_maps_close(getattr(proj, name), r, name)
    nn = 8
    proj.estimate_grid(nn=nn, atype="joint")
    assert proj.mean_Te_grid.shape == (nx // nn, ny // nn) == proj.new_mask_grid.shape
    assert proj.new_mask_grid[1, 1] and not proj.new_mask_grid[0, 0]
    checked = 0
    for ci, i in enumerate(range(0, nx - nn, nn)):
        for cj, j in enumerate(range(0, ny - nn, nn)):
            if mask[i, j]:
                assert proj.mean_Te_grid[ci, cj] == 0 and proj.status_grid[ci, cj] == 0
                continue
            r = l2fit.L2_estimate_cell(proj.k, ref[0][i, j], ref[1][i, j], ref[2][i, j], ref[3][i, j],
                                       conf=ora.FlexConf())
            gm = [proj.mean_Te_grid[ci, cj], proj.mean_F_grid[ci, cj]]
            gs = [proj.std_Te_grid[ci, cj], proj.std_F_grid[ci, cj]]
            checked += _compare_fit(gm, gs, proj.chi2_grid[ci, cj], r["mean"], r["std"], r["chi2"], 2)
    assert checked >= 5
    # decimated maps keep zeros in the never-visited last row / column when nn divides the size
    assert (proj.mean_Te_grid[-1, :] == 0).all() and (proj.mean_Te_grid[:, -1] == 0).all()
    s = proj.estimate_cell(cell=(0, 0), returned=True)
    assert s.loc["Te", "mean"] == proj.mean_Te_grid[0, 0]


def test_wlet_scalogram_paths_agree_like_the_reference_docstring(pf, engine):
    # classes.py:251-273: np.allclose(grid1.wl_sg, grid2.wl_sg) with and without a prior wlet_transform()
    rng = np.random.default_rng(0)
    data = rng.standard_normal((40, 40))
    g1, g2 = pf.Grid(data, 10., 10.), pf.Grid(data, 10., 10.)
    g1.wlet_scalogram()
    g2.wlet_transform()
    g2.wlet_scalogram()
    assert g2.wl_trans.shape == (40, 40, 11, g2.ns) and g1.wl_sg.shape == (40, 40, g1.ns)
    assert np.allclose(g1.wl_sg, g2.wl_sg, rtol=1e-12) and np.allclose(g1.ewl_sg, g2.ewl_sg, rtol=1e-10)


@pytest.mark.parametrize("n", [1024, 2048])
def test_full_size_properties(pf, engine, n):
    """BASELINE configs[1] size (1024^2, 5 km): properties that need no oracle.
    Linearity of the transform, scalogram == azimuth mean of |W|^2, admittance of a scaled copy."""
    from plateflex_b200 import classes
    from plateflex_b200.cpwt import cpwt, fused
    from plateflex_b200 import synthetic
    topo, grav = synthetic.make_pair(n, n, 5.0, 5.0, seed=20240000 + n)
    k = classes._lam2k(n, n, 5.0, 5.0)[1][[0, 9, 19]]
    wa = cpwt.wlet_transform(topo, 5.0, 5.0, k)
    wb = cpwt.wlet_transform(grav, 5.0, 5.0, k)
    wc = cpwt.wlet_transform(2.0 * topo - 3.0 * grav, 5.0, 5.0, k)
    for s in range(3):
        scale = max(np.abs(wa[..., s]).max() * 2, np.abs(wb[..., s]).max() * 3)
        assert np.abs(wc[..., s] - (2.0 * wa[..., s] - 3.0 * wb[..., s])).max() <= 1e-11 * scale
    sg, _ = fused.wlet_scalogram(topo, 5.0, 5.0, k)
    assert np.allclose(sg, (np.abs(wa) ** 2).mean(axis=2), rtol=1e-11, atol=1e-13 * sg.max())
    adm, eadm, coh, ecoh = fused.wlet_admit_coh(topo, 0.05 * topo, 5.0, 5.0, k)
    assert np.allclose(adm, 0.05, rtol=1e-10) and np.allclose(coh, 1.0, rtol=1e-10)
    assert np.abs(eadm).max() < 1e-10 and np.abs(ecoh).max() < 1e-10


def test_c3_size_properties_without_oracle(pf):
    """BASELINE configs[2] size (4096^2, padded 8192^2: four parts per line in the second-generation kernels),
    fused path only: gravity = c * topography must give admittance c, coherence 1 and vanishing jackknife errors
    at every cell; the scalogram of a sum of two fields is checked against its parts through the parallelogram
    identity |a+b|^2 + |a-b|^2 = 2|a|^2 + 2|b|^2 (azimuth means are linear in |W|^2)."""
    from plateflex_b200 import _lib, classes, synthetic
    from plateflex_b200.cpwt import fused
    n = 4096
    _lib.clear_plans()
    topo, grav = synthetic.make_pair(n, n, 5.0, 5.0, seed=20240000 + n)
    k = classes._lam2k(n, n, 5.0, 5.0)[1][:20][[0, 10, 19]]
    adm, eadm, coh, ecoh = fused.wlet_admit_coh(topo, -0.07 * topo, 5.0, 5.0, k)
    assert np.allclose(adm, -0.07, rtol=1e-10) and np.allclose(coh, 1.0, rtol=1e-10)
    assert np.abs(eadm).max() < 1e-10 and np.abs(ecoh).max() < 1e-10
    del adm, eadm, coh, ecoh
    sa, _ = fused.wlet_scalogram(topo, 5.0, 5.0, k)
    sb, _ = fused.wlet_scalogram(40.0 * grav, 5.0, 5.0, k)
    sp, _ = fused.wlet_scalogram(topo + 40.0 * grav, 5.0, 5.0, k)
    sm, _ = fused.wlet_scalogram(topo - 40.0 * grav, 5.0, 5.0, k)
    for s in range(3):
        lhs, rhs = sp[..., s] + sm[..., s], 2.0 * (sa[..., s] + sb[..., s])
        assert np.abs(lhs - rhs).max() <= 1e-11 * rhs.max()
    _lib.clear_plans()


@pytest.mark.skipif(len(ENGINES) < 2, reason="needs both engines")
def test_dense_and_pruned_engines_agree_at_full_size(pf, monkeypatch):
    from plateflex_b200 import _lib, classes, synthetic
    from plateflex_b200.cpwt import fused
    n = 1024
    topo, grav = synthetic.make_pair(n, n, 5.0, 5.0, seed=20240000 + n)
    k = classes._lam2k(n, n, 5.0, 5.0)[1][:20][[0, 5, 10, 15, 19]]
    res = {}
    for eng in ENGINES:
        monkeypatch.setenv("PLATEFLEX_B200_ENGINE", eng)
        _lib.clear_plans()
        res[eng] = fused.wlet_admit_coh(topo, grav, 5.0, 5.0, k)
    _lib.clear_plans()
    for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), res["pruned"], res["dense"]):
        # The jackknife errors are differences of nearly equal ratios: over 5e6 cells a few elements amplify
        # the 1e-16 rounding differences of two FFT algorithms beyond 1e-9 of their own value; the plane-level
        # bound stays 1e-9, the element-wise one is 1e-8 for the error maps.
        assert_maps_close(a, b, name, rtol=1e-8 if name.startswith("e") else 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("nn,world,alph", [(1, 3, False), (3, 4, True), (7, 2, False)])
def test_row_sharded_fit_equals_full_grid_fit(pf, nn, world, alph):
    """pfx_l2_fit_rows_dev on row shards (the multi-GPU cell sharding, emulated on one device) must
    reproduce the single-call grid fit bit for bit, decimation and unfitted last row included."""
    import ctypes as C
    import torch
    from plateflex_b200 import _lib, estimate, sharding, synthetic
    from plateflex_b200.cpwt import fused
    from plateflex_b200.flex import conf_flex
    pf.set_conf_flex()
    nx, ny, dx, dy = 45, 38, 20.0, 20.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240011, Te=30.0, F=0.4)
    k = np.array([1.0e-5, 1.6e-5, 2.5e-5, 4.0e-5, 6.0e-5, 9.0e-5])
    maps = fused.wlet_admit_coh(topo, grav, dx, dy, k)
    wd = np.asfortranarray(np.abs(np.random.default_rng(3).standard_normal((nx, ny))) * 50.0)
    full = estimate.L2_estimate_grid(k, *maps, wd, nn=nn, alph=alph, atype="joint")
    names = ["mean_Te", "std_Te", "mean_F", "std_F"] + (["mean_a", "std_a"] if alph else []) + ["chi2"]
    dev = torch.device("cuda", _lib.default_device())
    conf = conf_flex.as_struct()
    d_k = torch.from_numpy(k).to(dev)
    mx = nx // nn
    pieces = {n: [] for n in names + ["status"]}
    for r in range(world):
        row0, row1 = sharding.row_range(ny, world, r)
        nrows = row1 - row0
        if nrows == 0:
            continue
        cj0, myl = C.c_int(), C.c_int()
        _lib.check(_lib.lib.pfx_l2_fit_rows_shape(ny, row0, nrows, nn, C.byref(cj0), C.byref(myl)))
        # (nx, nrows, ns) Fortran-ordered shard == C-ordered [s][j][i]
        sh = [torch.from_numpy(np.ascontiguousarray(np.transpose(m, (2, 1, 0))[:, row0:row1, :])).to(dev) for m in maps]
        d_wd = torch.from_numpy(np.ascontiguousarray(wd.T[row0:row1])).to(dev)
        o = [torch.full((max(mx * myl.value, 1),), -7.0, dtype=torch.float64, device=dev) for _ in range(7)]
        st = torch.full((max(mx * myl.value, 1),), -7, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib.pfx_l2_fit_rows_dev(
            dev.index, None, C.byref(conf), d_k.data_ptr(), len(k), nx, ny, row0, nrows, *[t.data_ptr() for t in sh],
            d_wd.data_ptr(), None, None, None, nn, int(alph), _lib.ATYPES["joint"], o[0].data_ptr(), o[1].data_ptr(),
            o[2].data_ptr(), o[3].data_ptr(), o[4].data_ptr() if alph else None, o[5].data_ptr() if alph else None,
            o[6].data_ptr(), st.data_ptr()))
        torch.cuda.synchronize()
        idx = {"mean_Te": 0, "std_Te": 1, "mean_F": 2, "std_F": 3, "mean_a": 4, "std_a": 5, "chi2": 6}
        for n in names:
            pieces[n].append(o[idx[n]][:mx * myl.value].cpu().numpy().reshape(myl.value, mx))
        pieces["status"].append(st[:mx * myl.value].cpu().numpy().reshape(myl.value, mx))
    raw = np.concatenate(pieces["status"], axis=0).T
    assert np.array_equal(raw >> 8, full["nfev"]) and (full["nfev"][full["status"] == 1] >= 2).all()
    pieces["status"] = [p & 0xff for p in pieces["status"]]
    for n in names + ["status"]:
        got = np.concatenate(pieces[n], axis=0).T  # (mx, my)
        assert got.shape == full[n].shape, (n, got.shape, full[n].shape)
        assert np.array_equal(got, full[n], equal_nan=True), n
    assert (full["status"] == 1).sum() > 0


@pytest.mark.parametrize("with_grids", [False, True])
def test_ocean_free_air_project_with_alpha_matches_scipy(pf, ora, with_grids):
    """BASELINE configs[3] in small: free-air gravity over ocean (wd > 0 per cell, water-loaded model),
    3-parameter fit (alpha), optionally per-cell crustal density / thickness grids (classes.py:1087-1091),
    against SciPy TRF on the oracle model with each cell's own constants."""
    from oracle import l2fit
    from plateflex_b200 import _lib, synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    _lib.clear_plans()
    nx, ny, dx, dy = 40, 36, 20.0, 20.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240004, Te=20.0, F=0.2, boug=False, ocean_depth=4000.0,
                                     rhoc=2800.0)
    grids = [pf.TopoGrid(topo.copy(), dx, dy), pf.FairGrid(grav, dx, dy)]
    rhoc = zc = None
    if with_grids:
        rhoc = 2800.0 + 60.0 * red_field(nx, ny, seed=5)
        zc = 30.0e3 + 3.0e3 * red_field(nx, ny, seed=6)
        grids += [pf.RhocGrid(rhoc.copy(), dx, dy), pf.ZcGrid(zc.copy(), dx, dy)]
    else:
        pf.cf_fl.rhoc = 2800.0
    proj = pf.Project(grids=grids)
    proj.init()
    assert pf.cf_fl.boug == 0 and (proj.water_depth > 0).all()
    proj.wlet_admit_coh()
    nn = 7
    proj.estimate_grid(nn=nn, alph=True, atype="joint")
    maps = (proj.wl_admit, proj.ewl_admit, proj.wl_coh, proj.ewl_coh)
    checked = 0
    for ci, i in enumerate(range(0, nx - nn, nn)):
        for cj, j in enumerate(range(0, ny - nn, nn)):
            conf = ora.FlexConf(wd=float(proj.water_depth[i, j]), boug=0,
                                rhoc=float(rhoc[i, j]) if with_grids else 2800.0,
                                zc=float(zc[i, j]) if with_grids else 35.0e3)
            r = l2fit.L2_estimate_cell(proj.k, *[m[i, j] for m in maps], alph=True, atype="joint", conf=conf)
            gm = [proj.mean_Te_grid[ci, cj], proj.mean_F_grid[ci, cj], proj.mean_a_grid[ci, cj]]
            gs = [proj.std_Te_grid[ci, cj], proj.std_F_grid[ci, cj], proj.std_a_grid[ci, cj]]
            assert proj.status_grid[ci, cj] == 1
            checked += _compare_fit(gm, gs, proj.chi2_grid[ci, cj], r["mean"], r["std"], r["chi2"], 3)
    assert checked >= 5
    pf.set_conf_flex()


@pytest.mark.parametrize("nshards", [1, 3, 8])
def test_sharded_epilogue_writes_the_row_shard_layout(pf, engine, nshards):
    """pfx_wlet_admit_coh_sharded_dev (the multi-GPU epilogue that stores into the owners' row shards; here all
    shards are local buffers) must reproduce the plain outputs bit for bit in the [map][scale][row][x] layout
    that pfx_l2_fit_rows_dev reads, for scale blocks enqueued separately as different ranks would."""
    import ctypes as C
    import torch
    from plateflex_b200 import _lib, classes, sharding, synthetic
    from plateflex_b200.cpwt import fused
    nx, ny, dx, dy = 37, 29, 20.0, 20.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240021, Te=30.0, F=0.4)
    k = classes._lam2k(nx, ny, dx, dy)[1][:7]
    ns = len(k)
    want = fused.wlet_admit_coh(topo, grav, dx, dy, k)  # (nx, ny, ns) Fortran-ordered
    dev = torch.device("cuda", _lib.default_device())
    plan = _lib.get_plan(nx, ny, dx, dy, k, pf.cf_wt.k0)
    d_t, d_g = torch.from_numpy(topo).to(dev), torch.from_numpy(grav).to(dev)
    bufs = []
    for q in range(nshards):
        nbytes = _lib.lib.pfx_shard_bytes(nx, ny, ns, nshards, q)
        r0, r1 = sharding.row_range(ny, nshards, q)
        assert nbytes == 4 * ns * (r1 - r0) * nx * 8
        bufs.append(torch.full((max(nbytes // 8, 1),), np.nan, dtype=torch.float64, device=dev))
    bases = (C.c_void_p * nshards)(*[b.data_ptr() for b in bufs])
    for r in range(2):  # two "ranks", each with its block of scales
        is0, is1 = sharding.scale_range(ns, 2, r)
        _lib.check(_lib.lib.pfx_wlet_admit_coh_sharded_dev(plan.handle, d_t.data_ptr(), d_g.data_ptr(), is0, is1,
                                                           nshards, bases))
    torch.cuda.synchronize()

    def check():
        for q in range(nshards):
            r0, r1 = sharding.row_range(ny, nshards, q)
            if r1 == r0:
                continue
            got = bufs[q].cpu().numpy()[:4 * ns * (r1 - r0) * nx].reshape(4, ns, r1 - r0, nx)
            for m in range(4):
                ref = np.transpose(want[m], (2, 1, 0))[:, r0:r1, :]
                assert np.array_equal(got[m], ref, equal_nan=True), (q, m)

    check()
    # scales dealt out by cost instead of cut into blocks (non-contiguous lists, any order)
    for b in bufs:
        b.fill_(float("nan"))
    dealt = sharding.deal_scales(plan.scale_cost(), 3)
    assert sorted(s for part in dealt for s in part) == list(range(ns))
    for part in dealt:
        arr = (C.c_int * len(part))(*part[::-1])
        _lib.check(_lib.lib.pfx_wlet_admit_coh_sharded_list_dev(plan.handle, d_t.data_ptr(), d_g.data_ptr(), arr,
                                                                len(part), nshards, bases))
    torch.cuda.synchronize()
    check()


def test_float32_outputs_are_the_float64_results_rounded_once(pf, engine):
    """set_output_dtype(float32): what numpy.f2py returns for the reference's REAL arrays."""
    from plateflex_b200 import synthetic
    nx, ny, dx, dy = 33, 47, 10.0, 12.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240031)
    try:
        pf.set_output_dtype(np.float64)
        p64 = pf.Project(grids=[pf.TopoGrid(topo.copy(), dx, dy), pf.BougGrid(grav, dx, dy)])
        p64.init()
        p64.wlet_admit_coh()
        p64.grids[1].wlet_scalogram()
        pf.set_output_dtype(np.float32)
        p32 = pf.Project(grids=[pf.TopoGrid(topo.copy(), dx, dy), pf.BougGrid(grav, dx, dy)])
        p32.init()
        p32.wlet_admit_coh()
        p32.grids[1].wlet_scalogram()
    finally:
        pf.set_output_dtype(np.float64)
    for name in ("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"):
        a32, a64 = getattr(p32, name), getattr(p64, name)
        assert a32.dtype == np.float32 and a32.flags.f_contiguous and a64.dtype == np.float64
        assert np.array_equal(a32, a64.astype(np.float32), equal_nan=True), name
    p64.grids[0].wlet_transform()
    pf.set_output_dtype(np.float32)
    try:
        p32.grids[0].wlet_transform()
    finally:
        pf.set_output_dtype(np.float64)
    assert p32.grids[0].wl_trans.dtype == np.complex64 and p32.grids[0].wl_trans.flags.f_contiguous
    assert np.array_equal(p32.grids[0].wl_trans, p64.grids[0].wl_trans.astype(np.complex64))
    assert np.array_equal(p32.grids[1].wl_sg, p64.grids[1].wl_sg.astype(np.float32))
    assert np.array_equal(p32.grids[1].ewl_sg, p64.grids[1].ewl_sg.astype(np.float32))
    p32.estimate_grid(nn=6)  # the fit takes float32 maps like SciPy takes the reference's
    p64.estimate_grid(nn=6)
    assert np.allclose(p32.mean_Te_grid, p64.mean_Te_grid, atol=0.05)


# ---------------------------------------------------------------------------------------------------------
# Round 2: parity at the BASELINE sizes against the oracle itself (not against properties)
# ---------------------------------------------------------------------------------------------------------
def _bench_pair(cfgname):
    """The exact synthetic pair and wavenumbers bench.py uses for a BASELINE configuration."""
    import bench
    cfg = bench.CONFIGS[cfgname]
    topo, grav, k, k0 = bench.workload(cfg)
    return cfg, topo, grav, k, k0


@pytest.mark.parametrize("cfgname,scales", [("C2", [0, 10, 19]), ("C4", [4, 18])])
def test_pruned_engine_matches_oracle_at_baseline_sizes(pf, ora, cfgname, scales):
    """BASELINE configs[1] (1024^2 land Bouguer) on scales {0, 10, 19} and configs[3] (2048^2 ocean free-air,
    wd > 0) on two scales: the default engine through the C ABI against ora64 on the very grids bench.py times.
    Tolerance (tests/helpers.py): per scale plane max|d|/max|ref| <= 1e-9 and element-wise rtol 1e-9 for all four
    maps.  For the two jackknife ERROR maps the element-wise bound carries one more term.  The reference's jackknife
    divides by RUNNING PREFIX sums (cpwt_sub.f90:310-342), the first of which is the power of a single azimuth;
    where that one coefficient is weak (down to 1e-9 of the plane's maximum power at the smallest scale) its
    relative accuracy -- a float64 FFT is accurate to ~3e-14 of the plane's maximum AMPLITUDE, not of the cell --
    bounds the accuracy of the whole error estimate.  Term: 2e-14 * sqrt(max power of the plane / weakest single
    azimuth power of the cell) (times the natural amplitude sqrt(P2/P1) for the admittance).  Calibrated on two
    independent float64 CPU transforms of this very grid (ora64's Numerical-Recipes fourn and numpy.fft): without
    the term they disagree in 10 cells of ewl_coh at scale 19 (up to 2.1e-9 relative), with it the worst cell
    sits at 0.2 of the bound; every other map and scale agrees to rtol 1e-9 without it -- a property of the
    estimator's conditioning, not of the GPU path."""
    from plateflex_b200 import _lib
    from plateflex_b200.cpwt import fused
    cfg, topo, grav, k, k0 = _bench_pair(cfgname)
    _lib.clear_plans()
    ks = k[scales]
    got = fused.wlet_admit_coh(topo, grav, cfg["dx"], cfg["dy"], ks)
    nt = os.cpu_count() or 1
    for n, s in enumerate(scales):
        w1 = ora.wlet_transform(topo, cfg["dx"], cfg["dy"], k[[s]], k0=k0, prec=64, nthreads=nt)
        w2 = ora.wlet_transform(grav, cfg["dx"], cfg["dy"], k[[s]], k0=k0, prec=64, nthreads=nt)
        ref = ora.wlet_admit_coh(w1, w2, prec=64)
        p1, p2 = np.abs(w1) ** 2, np.abs(w2) ** 2
        del w1, w2
        weak = np.sqrt(np.maximum(p1.max(axis=(0, 1, 2)) / p1.min(axis=2), p2.max(axis=(0, 1, 2)) / p2.min(axis=2)))
        amp = np.sqrt(p2.mean(axis=2) / p1.mean(axis=2))
        del p1, p2
        floors = {"ewl_admit": 2e-14 * weak * amp, "ewl_coh": 2e-14 * weak}
        for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), got, ref):
            assert_maps_close(a[..., n:n + 1], b, f"{cfgname} {name} scale {s}", floor=floors.get(name))
    _lib.clear_plans()


@pytest.mark.parametrize("cfgname", ["C2", "C4"])
def test_fit_matches_scipy_on_sampled_cells_of_the_baseline_maps(pf, ora, cfgname):
    """>= 512 random cells of the admittance / coherence maps of BASELINE configs[1] (Te, F) and configs[3]
    (ocean, free air, alpha fitted): the device fit of the whole grid against the reference's L2_estimate_cell
    logic on SciPy TRF (oracle/l2fit.py).  Interior optima: |dTe| < 0.1 km, |dF| < 1e-3, |dalpha| < 1e-2,
    chi2 rel 1e-4, std rel 1e-2; cells where SciPy sits on a bound are compared on chi2 only and counted."""
    from oracle import l2fit
    from plateflex_b200 import _lib, estimate
    from plateflex_b200.cpwt import fused
    from plateflex_b200.flex import conf_flex
    cfg, topo, grav, k, k0 = _bench_pair(cfgname)
    pf.set_conf_flex()
    conf_flex.boug = 1 if cfg["boug"] else 0
    rhoc = 2800. if cfg["ocean"] else 2700.
    conf_flex.rhoc = rhoc
    _lib.clear_plans()
    maps = fused.wlet_admit_coh(topo, grav, cfg["dx"], cfg["dy"], k)
    wd = -np.minimum(topo, 0.0)
    alph = bool(cfg["alph"])
    res = estimate.L2_estimate_grid(k, *maps, wd, nn=1, alph=alph, atype="joint")
    rng = np.random.default_rng(2026)
    nx, ny = topo.shape
    cells = np.stack([rng.integers(0, nx - 1, 512), rng.integers(0, ny - 1, 512)], axis=1)
    npar = 3 if alph else 2
    interior = bound = 0
    for i, j in cells:
        conf = ora.FlexConf(boug=conf_flex.boug, rhoc=rhoc, wd=float(wd[i, j]))
        r = l2fit.L2_estimate_cell(k, maps[0][i, j], maps[1][i, j], maps[2][i, j], maps[3][i, j], alph=alph,
                                   atype="joint", conf=conf)
        assert res["status"][i, j] == 1
        gm = [res["mean_Te"][i, j], res["mean_F"][i, j]] + ([res["mean_a"][i, j]] if alph else [])
        gs = [res["std_Te"][i, j], res["std_F"][i, j]] + ([res["std_a"][i, j]] if alph else [])
        if _compare_fit(gm, gs, res["chi2"][i, j], r["mean"], r["std"], r["chi2"], npar):
            interior += 1
        else:
            bound += 1
    assert interior + bound == 512 and interior >= 256, (interior, bound)
    pf.set_conf_flex()
    _lib.clear_plans()


@pytest.mark.parametrize("variant", ["static", "group"])
@pytest.mark.parametrize("alph", [False, True])
def test_fit_kernel_variants_match_the_default(pf, monkeypatch, variant, alph):
    """PLATEFLEX_B200_FIT=static (one cell per thread; also the fallback when the work counter cannot be
    allocated) and =group (8 lanes per cell) run the same TRF arithmetic per cell as the lane-persistent
    default: golden cells must agree with it to rounding."""
    from plateflex_b200 import estimate
    z = np.load(os.path.join(GOLD, "l2fit_cells.npz"))
    pf.set_conf_flex()
    n, ns = z["adm"].shape

    def lay(a):
        m = np.ones((n + 1, 2, ns), order="F")
        m[:n, 0, :] = a
        return m

    args = (z["k"], lay(z["adm"]), lay(z["eadm"]), lay(z["coh"]), lay(z["ecoh"]))
    ref = estimate.L2_estimate_grid(*args, wd=np.zeros((n + 1, 2)), nn=1, alph=alph, atype="joint")
    monkeypatch.setenv("PLATEFLEX_B200_FIT", variant)
    got = estimate.L2_estimate_grid(*args, wd=np.zeros((n + 1, 2)), nn=1, alph=alph, atype="joint")
    assert np.array_equal(got["status"], ref["status"])
    for name in ["mean_Te", "mean_F", "chi2", "std_Te", "std_F"] + (["mean_a", "std_a"] if alph else []):
        tol = dict(rtol=1e-12, atol=0) if variant == "static" else dict(rtol=1e-6, atol=1e-9)
        assert np.allclose(got[name], ref[name], **tol), (variant, name)


@pytest.mark.parametrize("alph,boug", [(False, 1), (True, 0)])
def test_dry_land_fit_kernel_equals_general_form_bit_for_bit(pf, monkeypatch, alph, boug):
    """The lane kernel has a dry-land form (no fitted cell under water, one crustal thickness: no per-lane array of
    depth factors, ten warps per SM instead of eight) chosen per launch by a device flag.  It performs the general
    form's operations in the general form's order, so a launch may take either -- ranks of a row-sharded fit do --
    without changing a bit of the result."""
    from plateflex_b200 import estimate
    from plateflex_b200.flex import conf_flex
    z = np.load(os.path.join(GOLD, "l2fit_cells.npz"))
    pf.set_conf_flex()
    conf_flex.boug = boug
    n, ns = z["adm"].shape

    def lay(a):
        m = np.ones((n + 1, 2, ns), order="F")
        m[:n, 0, :] = a
        return m

    args = (z["k"], lay(z["adm"]), lay(z["eadm"]), lay(z["coh"]), lay(z["ecoh"]))
    dry = estimate.L2_estimate_grid(*args, wd=np.zeros((n + 1, 2)), nn=1, alph=alph, atype="joint")
    monkeypatch.setenv("PLATEFLEX_B200_FIT_DRY", "off")
    gen = estimate.L2_estimate_grid(*args, wd=np.zeros((n + 1, 2)), nn=1, alph=alph, atype="joint")
    wd = np.zeros((n + 1, 2))
    wd[0, 0] = 10.0  # one fitted cell under water: the device flag sends the whole launch to the general form
    monkeypatch.delenv("PLATEFLEX_B200_FIT_DRY")
    wet = estimate.L2_estimate_grid(*args, wd=wd, nn=1, alph=alph, atype="joint")
    for name in ["status", "mean_Te", "mean_F", "chi2", "std_Te", "std_F"] + (["mean_a", "std_a"] if alph else []):
        assert np.array_equal(dry[name], gen[name], equal_nan=True), name
        assert np.array_equal(dry[name][1:], wet[name][1:], equal_nan=True), name
    pf.set_conf_flex()


@pytest.mark.parametrize("nx,ny", [(8200, 12), (12, 8200)])
def test_padded_axis_16384_matches_oracle(pf, ora, nx, ny):
    """BASELINE configs[4] (8192^2) pads both axes to 16384 samples (eight parts per line).  The full grid is
    beyond what the CPU oracle finishes in a test, so the 16384-sample axis is exercised on a ragged strip, once
    per axis, against ora64; the full-size run itself is covered by bench.py --config C5 (no oracle)."""
    from plateflex_b200 import _lib, classes
    from plateflex_b200.cpwt import fused
    _lib.clear_plans()
    t = red_field(nx, ny, seed=5) * 500.0
    g = 0.06 * t + red_field(nx, ny, seed=6) * 25.0
    k = classes._lam2k(8192, 8192, 5.0, 5.0)[1][:20][[2, 11, 19]]
    got = fused.wlet_admit_coh(t, g, 5.0, 5.0, k)
    ref = ora.admit_coh_from_grids(t, g, 5.0, 5.0, k, k0=pf.cf_wt.k0, prec=64, nthreads=os.cpu_count() or 1)
    for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), got, ref):
        assert_maps_close(a, b, name)
    _lib.clear_plans()


def test_project_keeps_maps_on_the_device_until_read(pf, ora):
    """Project.wlet_admit_coh leaves the four maps on the GPU; estimate_grid fits them there; reading an
    attribute copies it once; assigning one falls back to the host path with identical results."""
    from plateflex_b200 import synthetic
    pf.set_conf_cpwt()
    pf.set_conf_flex()
    nx, ny, dx, dy = 40, 36, 20.0, 20.0
    topo, grav = synthetic.make_pair(nx, ny, dx, dy, seed=20240011, Te=25.0, F=0.4)

    def project():
        p = pf.Project(grids=[pf.TopoGrid(topo.copy(), dx, dy), pf.BougGrid(grav, dx, dy)])
        p.init()
        return p

    a = project()
    assert not hasattr(a, "wl_admit")
    a.wlet_admit_coh()
    assert a.__dict__["_dmaps"] is not None and a.__dict__["_host_wl_admit"] is None
    a.estimate_grid(nn=4)  # device path: no map has crossed to the host yet
    assert a.__dict__["_host_wl_admit"] is None
    s = a.estimate_cell(cell=(4, 8), returned=True)  # one cell fetched from the device
    assert a.__dict__["_host_wl_coh"] is None and s.loc["Te", "mean"] == a.mean_Te_grid[1, 2]
    adm = a.wl_admit  # first read copies
    assert adm.shape == (nx, ny, a.ns) and adm.flags.f_contiguous and a.wl_admit is adm
    ref = ora.admit_coh_from_grids(a.grids[0].data, grav, dx, dy, a.k, k0=pf.cf_wt.k0)
    for name, r in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), ref):
        assert_maps_close(getattr(a, name), r, name)
    b = project()
    b.wlet_admit_coh()
    b.wl_admit = b.wl_admit.copy()  # assignment drops the device copy: host path from here on
    assert b.__dict__["_dmaps"] is None
    b.estimate_grid(nn=4)
    for name in ("mean_Te_grid", "std_Te_grid", "mean_F_grid", "std_F_grid", "chi2_grid", "status_grid"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def _ipc_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bench
        import plateflex_b200 as pf
        from plateflex_b200 import sharding
        pf.set_conf_cpwt()
        pf.set_conf_flex()
        cfg = dict(bench.CONFIGS["C1"], nx=96, ny=70, ns=9)
        topo, grav, k, k0 = bench.workload(cfg)
        conf = bench.flex_conf(cfg)
        wd_f = np.ascontiguousarray((-np.minimum(topo, 0.0)).T)
        pipe = sharding.TeMapPipeline(cfg["nx"], cfg["ny"], cfg["dx"], cfg["dy"], k, k0, conf, local_device=rank)
        if rank == 0:
            pipe.load(torch.from_numpy(topo), torch.from_numpy(grav), torch.from_numpy(wd_f))
        got = pipe.run()
        got2 = pipe.run()  # a second step reuses the shards: the broadcast orders it behind the first gather
        torch.cuda.synchronize()
        ok = True
        if rank == 0:
            solo = sharding.TeMapPipeline(cfg["nx"], cfg["ny"], cfg["dx"], cfg["dy"], k, k0, conf, local_device=0, solo=True)
            solo.load(torch.from_numpy(topo), torch.from_numpy(grav), torch.from_numpy(wd_f))
            ref = solo.run()
            torch.cuda.synchronize()
            ok = bool(torch.equal(torch.nan_to_num(got), torch.nan_to_num(ref))) and bool(torch.equal(got, got2) or
                     torch.equal(torch.nan_to_num(got), torch.nan_to_num(got2)))
            ok = ok and float(ref[0].abs().max().item()) > 0
            solo.close()
        pipe.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny", [(44, 37), (16, 1024)])
def test_cutoff_change_rebuilds_the_plan(pf, nx, ny):
    """pfx_plan_set_cutoff re-plans the pruned engine (supports, FFT lengths, tables, the cosine-transform state on
    power-of-two grids): 8.5 sigma instead of 9.5 changes the maps by less than the parity budget (the neglected
    tail is e^{-36} of the peak), 5 sigma (e^{-12.5}) visibly -- so the new supports are really in use."""
    import torch
    from plateflex_b200 import _lib
    dx = dy = 5.0
    t = red_field(nx, ny, seed=31) * 500.0
    g = 0.08 * t + red_field(nx, ny, seed=32) * 20.0
    kf = _wavenumbers(("ladder", 3), nx, ny, dx, dy)
    ns = len(kf)
    dev = torch.device("cuda", 0)
    d_t, d_g = torch.from_numpy(t).to(dev), torch.from_numpy(g).to(dev)
    plan = _lib.Plan(nx, ny, dx, dy, kf, pf.cf_wt.k0, device=0)
    plan.set_stream(torch.cuda.current_stream().cuda_stream)

    def run():
        outs = [torch.empty(ns * nx * ny, dtype=torch.float64, device=dev) for _ in range(4)]
        _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, d_t.data_ptr(), d_g.data_ptr(), 0, ns, *[o.data_ptr() for o in outs]))
        torch.cuda.synchronize()
        return [o.cpu().numpy().reshape(ns, ny, nx) for o in outs]

    ref = run()
    plan.set_cutoff(8.5)
    near = run()
    plan.set_cutoff(5.0)
    far = run()
    plan.set_cutoff(9.5)
    again = run()
    plan.close()
    for a, b, c, d in zip(ref, near, far, again):
        mx = np.abs(a).max(axis=(1, 2), keepdims=True)
        assert (np.abs(a - b) / mx).max() < 1e-9
        assert np.array_equal(a, d)
    assert (np.abs(ref[0] - far[0]) / np.abs(ref[0]).max(axis=(1, 2), keepdims=True)).max() > 1e-9


def test_second_device_in_one_process_matches_the_first(pf, monkeypatch):
    """One process, two devices: the host-pointer API on device 1 after device 0 has been used -- pageable host
    copies above 4 MiB go through pinned staging chunks whose events belong to one device each (ADVICE r01:
    a process-global staging area made any second device fail with an invalid resource handle)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from plateflex_b200 import _lib, estimate
    from plateflex_b200.cpwt import fused
    nx, ny, dx, dy = 600, 520, 5.0, 5.0  # 2.5 MB per scale and map: four scales make every copy exceed 4 MiB
    t = red_field(nx, ny, seed=21) * 500.0
    g = 0.08 * t + red_field(nx, ny, seed=22) * 20.0
    kf = _wavenumbers(("ladder", 5), nx, ny, dx, dy)
    outs = {}
    for dev in ("0", "1"):
        monkeypatch.setenv("PLATEFLEX_B200_DEVICE", dev)
        _lib.clear_plans()
        maps = fused.wlet_admit_coh(t, g, dx, dy, kf)
        fit = estimate.L2_estimate_grid(kf, *maps, wd=np.zeros((nx, ny)), nn=8, alph=False, atype="joint")
        outs[dev] = (maps, fit)
    _lib.clear_plans()
    for a, b in zip(outs["0"][0], outs["1"][0]):
        assert np.array_equal(a, b, equal_nan=True)
    for name in ("mean_Te", "mean_F", "chi2", "status"):
        assert np.array_equal(outs["0"][1][name], outs["1"][1][name], equal_nan=True), name


def test_two_gpu_pipeline_equals_single_gpu_bit_for_bit(pf):
    """Two processes, two GPUs: scale-sharded transform with the epilogue storing into the peer's row shard
    through a CUDA-IPC mapping, stream-ordered barrier, row-sharded fit, gather -- the maps on rank 0 must be
    bit-identical to the single-GPU maps.  Skipped on a one-GPU box (the driver's test box has one)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket

    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ipc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


@pytest.mark.parametrize("knobs", [{"PLATEFLEX_B200_V3": "on"}, {"PLATEFLEX_B200_V3": "1", "PLATEFLEX_B200_V3_OCC": "2"},
                                   {"PLATEFLEX_B200_V3": "2", "PLATEFLEX_B200_V3_P1": "off"}, {"PLATEFLEX_B200_V2H": "on"},
                                   {"PLATEFLEX_B200_V4": "on"}, {"PLATEFLEX_B200_V4": "p2", "PLATEFLEX_B200_V4_WAVES": "2"},
                                   {"PLATEFLEX_B200_V4": "off"}, {"PLATEFLEX_B200_DCT": "off"},
                                   {"PLATEFLEX_B200_DCT": "off", "PLATEFLEX_B200_V4": "off"}])
@pytest.mark.parametrize("nx,ny,dx,dy,kf", CASES[5:])
def test_third_generation_pass_kernels_match_oracle(pf, ora, monkeypatch, knobs, nx, ny, dx, dy, kf):
    """The opt-in bulk-copy (cp.async.bulk + mbarrier) pass kernels of pfx_pruned_v3.cuh: transposed intermediate,
    one / two line buffers, 2 or 3 CTAs per SM, with and without the third-generation pass 1; and the 128-thread
    form of the second-generation pass 2 (v2h); the radix-16 fourth generation (pfx_pruned_v4.cuh) for both passes
    or pass 2 only, and switched off (second generation); the cosine-transform form of the spectra (pfx_dct.cu,
    default on power-of-two grids) switched off, i.e. the padded complex forward transform."""
    from plateflex_b200 import _lib
    from plateflex_b200.cpwt import fused
    for name, value in knobs.items():
        monkeypatch.setenv(name, value)
    monkeypatch.setenv("PLATEFLEX_B200_ENGINE", "pruned")
    _lib.clear_plans()
    t = red_field(nx, ny, seed=1) * 500.0
    g = 0.08 * t + red_field(nx, ny, seed=2) * 20.0
    kf = _wavenumbers(kf, nx, ny, dx, dy)
    ref = ora.admit_coh_from_grids(t, g, dx, dy, kf, k0=pf.cf_wt.k0, prec=64, nthreads=os.cpu_count() or 1)
    for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), fused.wlet_admit_coh(t, g, dx, dy, kf), ref):
        assert_maps_close(a, b, name)
    _lib.clear_plans()


def test_nan_fill_and_gaussian_match_scipy(pf):
    """SURVEY.md 8f row 2 on the GPU: Grid.__init__'s nearest-neighbour NaN fill (classes.py:143-153) against
    scipy.interpolate.griddata(method='nearest') -- every filled value must come from a finite cell at the minimal
    distance, and equal SciPy's wherever that cell is unique -- and TopoGrid.filter_water_depth's Gaussian
    (classes.py:538-541) against scipy.ndimage.gaussian_filter(mode='nearest', truncate=4), 1e-12 relative."""
    from scipy import ndimage
    from scipy.interpolate import griddata
    from plateflex_b200 import io
    rng = np.random.default_rng(12)
    nx, ny = 61, 47
    g = rng.standard_normal((nx, ny))
    holes = g.copy()
    holes[rng.random((nx, ny)) < 0.08] = np.nan
    holes[:9, :7] = np.nan      # a corner block, as outside a map projection
    holes[30:38, 20:31] = np.nan
    got = io.fill_nan_nearest(holes)
    good = np.where(np.isfinite(holes))
    xx, yy = np.mgrid[0:nx, 0:ny]
    ref = griddata(np.array([xx[good], yy[good]]).T, holes[good], (xx, yy), method='nearest')
    assert np.isfinite(got).all() and np.array_equal(got[good], holes[good])
    gi, gj = good
    unique = same = 0
    for i, j in zip(*np.where(~np.isfinite(holes))):
        d2 = (gi - i) ** 2 + (gj - j) ** 2
        near = np.where(d2 == d2.min())[0]
        assert got[i, j] in holes[gi[near], gj[near]]
        if len(near) == 1:
            unique += 1
            same += got[i, j] == ref[i, j]
    assert unique > 50 and same == unique
    grid = pf.Grid(holes, 10., 10.)  # the class uses it
    assert np.array_equal(grid.data, got)
    for sigma in (1, 2.5, 10):
        out = io.gaussian_filter(g * 1000. - 2000., sigma)
        want = ndimage.gaussian_filter(g * 1000. - 2000., sigma, mode='nearest', truncate=4.0)
        assert np.abs(out - want).max() <= 1e-12 * np.abs(want).max()
    topo = pf.TopoGrid(g * 1000. - 200., 10., 10.)
    wd = topo.filter_water_depth(sigma=3, returned=True)
    want = ndimage.gaussian_filter(topo.data, 3, mode='nearest', truncate=4.0)
    want[topo.data > 0.] = 0.
    assert np.allclose(wd, want, rtol=1e-12, atol=1e-9) and np.array_equal(topo.water_depth, -1. * wd)


def test_f2py_helper_exports_match_oracle(pf, ora):
    """The subroutines numpy.f2py exports besides the drivers: flex.flexfilter_top/bot, decon,
    tr_func (flex.f90:74-169) and cpwt_sub's npow2, defk, wave_function, against the oracle at 1e-12."""
    import ctypes as C
    from plateflex_b200.cpwt import cpwt
    from plateflex_b200.flex import conf_flex, flex
    pf.set_conf_flex()
    conf_flex.wd, conf_flex.rhof, conf_flex.boug = 2500.0, conf_flex.rhow, 0
    oc = ora.FlexConf(wd=2500.0, rhof=conf_flex.rhow, boug=0)
    k = np.geomspace(3e-6, 4e-4, 17)
    psi = 1e11 * (25e3) ** 3 / 12 / (1 - 0.25 ** 2) * k ** 4
    lib = ora.lib()

    def call(name, *arrays_and_scalars):
        return getattr(lib, name)(C.byref(oc), C.c_int(len(k)), *arrays_and_scalars)

    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    th_ref, ph_ref = np.empty_like(k), np.empty_like(k)
    call("ora_flex_flexfilter_top", p(psi), p(th_ref))
    call("ora_flex_flexfilter_bot", p(psi), p(ph_ref))
    th, ph = flex.flexfilter_top(psi), flex.flexfilter_bot(psi)
    assert np.allclose(th, th_ref, rtol=1e-13) and np.allclose(ph, ph_ref, rtol=1e-13)
    refs = [np.empty_like(k) for _ in range(4)]
    call("ora_flex_decon", p(th_ref), p(ph_ref), p(k), C.c_double(1.0), *[p(r) for r in refs])
    for a, b in zip(flex.decon(th, ph, k, 1.0), refs):
        assert np.allclose(a, b, rtol=1e-12)
    adm, coh = flex.tr_func(*refs, 0.3, 1.1)
    a_ref, c_ref = ora.real_xspec_functions(k, 25.0, 0.3, 1.1, oc)
    assert np.allclose(adm.real, a_ref, rtol=1e-12) and np.allclose(coh, c_ref, rtol=1e-12) and np.abs(adm.imag).max() > 0
    assert [cpwt.npow2(n) for n in (1, 3, 684)] == [2, 4, 1024]
    kx, ky = cpwt.defk(16, 10, 5.0, 7.0)
    rkx, rky = ora.defk(16, 10, 5.0, 7.0)
    assert np.allclose(kx, rkx, rtol=1e-15) and np.allclose(ky, rky, rtol=1e-15) and kx[8] > 0
    w = cpwt.wave_function(kx, ky, pf.cf_wt.k0, 30.0, 0.4)
    rw = ora.wave_function(kx, ky, pf.cf_wt.k0, 30.0, 0.4)
    # (the DC element is an exact cancellation of the two Gaussians: compare it on the scale of the array)
    assert w.shape == (16, 10) and np.allclose(w, rw, rtol=1e-12, atol=1e-15 * np.abs(rw).max())
    pf.set_conf_flex()
