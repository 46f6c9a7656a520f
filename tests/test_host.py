"""CPU tests of the host-side mirror of the reference API (no GPU work is issued)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(GOLD, "lam2k_kat.json")))


def test_lam2k_docstring_known_answer(pf, kat):
    from plateflex_b200 import classes
    d = kat["docstring"]  # classes.py:1787-1796
    ns, k = classes._lam2k(d["nx"], d["ny"], d["dx"], d["dy"])
    assert ns == d["ns"] == len(k)
    assert np.allclose(k, d["k"], rtol=5e-8, atol=0)  # printed with 9 digits; k0 is float32-valued


def test_lam2k_notebook_known_answer(pf, kat):
    from plateflex_b200 import classes
    d = kat["notebook_ex1"]  # Ex1_making_grids.ipynb cell 12
    ns, k = classes._lam2k(d["nx"], d["ny"], d["dx"], d["dy"])
    assert ns == d["ns"]
    assert np.allclose(k, d["k"], rtol=6e-9, atol=0)  # matches only with the float32 k0
    for s in kat["docstring_shapes"]:  # classes.py:209-220, 251-273
        assert classes._lam2k(s["nx"], s["ny"], s["dx"], s["dy"])[0] == s["ns"]


def test_k0_is_float32_valued(pf):
    assert pf.cf_wt.k0 == float(np.float32(5.336)) != 5.336
    assert pf.cf_wt.na == 11


def test_conf_flex_defaults(pf):
    pf.set_conf_flex()
    c = pf.cf_fl
    assert (c.zc, c.rhom, c.rhoc, c.rhow, c.rhoa, c.rhof, c.wd, c.boug) == (35e3, 3200., 2700., 1030., 0., 0., 0., 1)
    assert (c.pi, c.g, c.em, c.nu, c.gc) == (3.141592653589793, 9.81, 1e11, 0.25, 6.67e-6)


def test_grid_attributes_and_unit_scaling(pf):
    rng = np.random.default_rng(0)
    data = rng.standard_normal((30, 20))
    g = pf.Grid(data, 10.0, 12.0)
    assert (g.nx, g.ny, g.dx, g.dy) == (30, 20, 10.0, 12.0) and g.ns == len(g.k)
    assert g.data is not data and np.array_equal(g.data, data)
    assert not hasattr(g, "wl_trans")
    topo_km = rng.standard_normal((30, 20))  # std < 20 -> treated as km (classes.py:530-531)
    keep = topo_km.copy()
    t = pf.TopoGrid(topo_km, 10.0, 12.0)
    assert np.allclose(t.data, keep * 1e3)
    # water depth from the caller's array, in its original units, mutating it (classes.py:533-536)
    assert np.array_equal(t.water_depth, -np.minimum(keep, 0.0))
    assert np.array_equal(topo_km, np.minimum(keep, 0.0))
    assert pf.RhocGrid(np.full((4, 4), 2.7) + rng.standard_normal((4, 4)) * 0.01, 1, 1).data.mean() > 2000
    assert isinstance(pf.BougGrid(data, 1, 1), pf.GravGrid) and pf.FairGrid(data, 1, 1).title == "Free-air anomaly"


@pytest.mark.gpu
def test_nan_fill(pf, capsys):
    """The fill runs on the GPU (io.fill_nan_nearest); without a device the call fails loudly, no CPU fallback."""
    data = np.arange(12.0).reshape(3, 4)
    data[1, 1] = np.nan
    g = pf.Grid(data, 1.0, 1.0)
    assert np.isfinite(g.data).all() and "NaN" in capsys.readouterr().out


def test_project_validation_messages(pf):
    rng = np.random.default_rng(1)
    t = pf.TopoGrid(rng.standard_normal((16, 16)) * 100 + 500, 10., 10.)
    b = pf.BougGrid(rng.standard_normal((16, 16)) * 30, 10., 10.)
    with pytest.raises(Exception, match="There needs to be one GravGrid object in Project"):
        pf.Project(grids=[t, t]).init()
    with pytest.raises(Exception, match="There needs to be one TopoGrid object in Project"):
        pf.Project(grids=[b]).init()
    with pytest.raises(Exception, match="Grids do not have the same shape"):
        pf.Project(grids=[t, pf.BougGrid(np.zeros((8, 8)), 10., 10.)]).init()
    with pytest.raises(Exception, match="same sampling intervals"):
        pf.Project(grids=[t, pf.BougGrid(np.zeros((16, 16)), 5., 10.)]).init()
    p = pf.Project(grids=[t, b])
    with pytest.raises(Exception, match="Project not yet initialized - Abort"):
        p.wlet_admit_coh()
    with pytest.raises(Exception, match="Project not initialized. Aborting"):
        p.estimate_cell()
    p.init()
    assert p.initialized and pf.cf_fl.boug == 1 and p.rhoc is None and p.zc is None
    assert (p.nx, p.ny, p.ns) == (16, 16, t.ns) and p.water_depth is t.water_depth
    with pytest.raises(Exception, match="'alph' should be a boolean"):
        p.estimate_grid(alph=1)
    with pytest.raises(Exception, match="'atype' should be one among"):
        p.estimate_grid(atype="both")
    with pytest.raises(Exception, match="Parallel implementation does not work"):
        p.estimate_grid(parallel=True)
    p2 = pf.Project(grids=[t, pf.FairGrid(np.zeros((16, 16)), 10., 10.)])
    p2.init()
    assert pf.cf_fl.boug == 0
    pf.set_conf_flex()


def test_project_container_protocol(pf):
    g1, g2 = pf.Grid(np.zeros((4, 4)), 1, 1), pf.Grid(np.ones((4, 4)), 1, 1)
    p = pf.Project()
    p += g1
    assert (p + g2).grids == [g1, g2] and list(pf.Project([g1, g2])) == [g1, g2]
    assert p.append(g2) is p and p.extend([g1]).grids == [g1, g2, g1]
    with pytest.raises(TypeError):
        p.append(3)
    with pytest.raises(TypeError):
        p.extend([3])
    with pytest.raises(TypeError):
        p + 3
    assert p.inverse == "L2" and p.mask is None and p.initialized is False


def test_save_results_writes_the_reference_csv_layout(pf, tmp_path, capsys):
    """classes.py:1621-1754: x = dy*nn*col, y = dx*nn*row, row-major; optional Gaussian smoothing."""
    import pandas as pd
    from scipy import ndimage
    proj = pf.Project()
    proj.dx, proj.dy, proj.nn = 20.0, 10.0, 5
    rng = np.random.default_rng(4)
    proj.mean_Te_grid = rng.uniform(5, 80, (4, 3))
    proj.std_Te_grid = rng.uniform(1, 5, (4, 3))
    proj.mean_F_grid = rng.uniform(0, 1, (4, 3))
    proj.std_F_grid = rng.uniform(0, 0.1, (4, 3))
    proj.chi2_grid = rng.uniform(0.5, 5, (4, 3))
    proj.mean_a_grid = rng.uniform(0, 3, (4, 3))
    proj.new_mask_grid = rng.uniform(0, 1, (4, 3)) > 0.5
    prefix = str(tmp_path / "PF_")
    proj.save_results(mean_Te=True, chi2=True, mask=True, mean_a=True, filter=False, prefix=prefix)
    df = pd.read_csv(prefix + "mean_Te.csv")
    assert list(df.columns) == ["x", "y", "mean_Te"] and len(df) == 12
    assert np.allclose(df["mean_Te"].values.reshape(4, 3), proj.mean_Te_grid)
    assert np.allclose(df["x"].values.reshape(4, 3)[0], [0.0, 50.0, 100.0])    # dy*nn*column
    assert np.allclose(df["y"].values.reshape(4, 3)[:, 0], [0.0, 100.0, 200.0, 300.0])  # dx*nn*row
    assert np.allclose(pd.read_csv(prefix + "mean_a.csv")["mean_a"].values.reshape(4, 3), proj.mean_a_grid)
    assert np.array_equal(pd.read_csv(prefix + "mask.csv")["mask"].values.reshape(4, 3).astype(bool), proj.new_mask_grid)
    # filtered: skimage.filters.gaussian == ndimage.gaussian_filter(mode='nearest', truncate=4)
    proj.save_results(std_Te=True, mean_a=True, filter=True, sigma=1, prefix=prefix + "f_")
    got = pd.read_csv(prefix + "f_std_Te.csv")["std_Te"].values.reshape(4, 3)
    assert np.allclose(got, ndimage.gaussian_filter(proj.std_Te_grid, 1, mode="nearest", truncate=4.0))
    assert "parameter 'alpha' was not estimated" in capsys.readouterr().out  # the reference's own quirk
    assert not (tmp_path / "PF_f_mean_a.csv").exists()
    with pytest.raises(AttributeError):
        proj.save_results(MAP_Te=True, prefix=prefix)  # exists only after a Bayesian estimation


@pytest.mark.gpu
def test_filter_water_depth(pf):
    topo = np.linspace(-3000.0, 800.0, 30 * 20).reshape(30, 20)
    t = pf.TopoGrid(topo.copy(), 10.0, 10.0)
    out = t.filter_water_depth(sigma=2, returned=True)
    assert out.shape == (30, 20) and (t.water_depth >= 0).all()
    assert (t.water_depth[t.data > 0] == 0).all() and (t.water_depth[t.data < -500] > 0).all()


def test_project_maps_are_lazy_attributes(pf):
    """wl_admit & co. behave like the reference's plain attributes (classes.py:968-971, probed with
    try/except AttributeError at classes.py:1001-1008): absent before wlet_admit_coh, assignable, deletable."""
    import numpy as np
    topo = np.abs(np.random.default_rng(0).standard_normal((6, 5))) * 100. + 50.
    proj = pf.Project(grids=[pf.TopoGrid(topo.copy(), 10., 10.), pf.BougGrid(topo * 0.1, 10., 10.)])
    proj.init()
    for name in ("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"):
        assert not hasattr(proj, name)
        try:
            getattr(proj, name)
            raise AssertionError("expected AttributeError")
        except AttributeError as exc:
            assert name in str(exc)
    a = np.ones((6, 5, proj.ns))
    proj.wl_admit = a
    assert proj.wl_admit is a and hasattr(proj, "wl_admit") and not hasattr(proj, "wl_coh")
    del proj.wl_admit
    assert not hasattr(proj, "wl_admit")
    # two projects do not share state through the class-level properties
    other = pf.Project(grids=[pf.TopoGrid(topo.copy(), 10., 10.), pf.BougGrid(topo * 0.1, 10., 10.)])
    proj.ewl_coh = a
    assert not hasattr(other, "ewl_coh")


def test_read_xyz_matches_the_notebooks_pandas_recipe(pf, tmp_path):
    """Ex1_making_grids.ipynb cells 2-3: header fields from the first line, z column reshaped (ny, nx)[::-1]."""
    import numpy as np
    import pandas as pd
    from plateflex_b200 import io
    nx, ny = 37, 23
    rng = np.random.default_rng(0)
    z = rng.standard_normal(nx * ny) * 1000.
    z[[5, 99]] = np.nan
    p = str(tmp_path / "Topo_NA.xyz")
    with open(p, "w") as f:
        f.write("# Topo_NA.grd\t-3380.0\t%g\t-3400.5\t%g\t-5300.5\t4200.25\t20.0169157785\t19.9970844654\t%d\t%d\n"
                % (3440.0, 3380.0, nx, ny))
        c = 0
        for j in range(ny):
            for i in range(nx):
                f.write("%.4f\t%.4f\t%s\n" % (i * 20.0169, (ny - 1 - j) * 19.997, "NaN" if np.isnan(z[c]) else "%.3f" % z[c]))
                c += 1
    xmin, xmax, ymin, ymax, zmin, zmax, dx, dy, nnx, nny = \
        pd.read_csv(p, sep='\t', nrows=0).columns[1:].values.astype(float)
    want = pd.read_csv(p, sep='\t', skiprows=1, names=['x', 'y', 'z'])['z'].values.reshape(int(nny), int(nnx))[::-1]
    data, gdx, gdy, h = io.read_xyz(p)
    assert (gdx, gdy, h["nx"], h["ny"]) == (dx, dy, nx, ny) and h["xmin"] == xmin and h["zmax"] == zmax
    assert data.shape == (ny, nx) and data.flags.c_contiguous and np.array_equal(data, want, equal_nan=True)
    assert np.array_equal(io.read_xyz(p, nthreads=3)[0], want, equal_nan=True)
    # errors: missing header, wrong count
    bad = str(tmp_path / "bad.xyz")
    open(bad, "w").write("0 0 1\n")
    try:
        io.read_xyz(bad)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
