"""CPU tests of the oracle itself (no GPU): it must agree with an independent numpy.fft
restatement, with literal pure-Python jackknives, with the analytic limits of the flexural
model, and with the committed golden fixtures."""
import os

import numpy as np
import pytest

from helpers import numpy_jackknife, numpy_transform, red_field

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_npow2(ora):
    # cpwt_sub.f90:26-33: smallest power of two >= n, never below 2
    assert [ora.npow2(n) for n in (1, 2, 3, 4, 5, 683, 684, 1024, 1025)] == [2, 2, 4, 4, 8, 1024, 1024, 1024, 2048]


def test_defk_positive_nyquist(ora):
    kx, ky = ora.defk(8, 4, 2.0, 3.0)
    dk = 2 * np.pi / (8 * 2.0)
    assert np.allclose(kx, np.array([0, 1, 2, 3, 4, -3, -2, -1]) * dk, rtol=1e-15)
    assert kx[4] > 0 and ky[2] > 0  # unlike numpy.fft.fftfreq


def test_wave_function_unscaled_second_term(ora):
    kx, ky = ora.defk(16, 8, 5.0, 5.0)
    k0, s, a = ora.K0_F32, 37.0, 0.3
    d = ora.wave_function(kx, ky, k0, s, a)
    kx0, ky0 = k0 * np.cos(a), k0 * np.sin(a)
    ref = s / np.sqrt(np.pi) * (np.exp(-0.5 * (s * kx[:, None] - kx0) ** 2) * np.exp(-0.5 * (s * ky[None, :] - ky0) ** 2)
                                - np.exp(-0.5 * (kx[:, None] ** 2 + ky[None, :] ** 2 + kx0 ** 2 + ky0 ** 2)))
    assert np.abs(d.imag).max() == 0
    assert np.allclose(d.real, ref, rtol=1e-12, atol=1e-15 * s)  # difference of two exponentials: absolute bound
    assert d[0, 0].real == 0.0 or abs(d[0, 0].real) < 1e-20  # admissibility: psi(0) = 0


@pytest.mark.parametrize("isign", [1, -1])
def test_fourn_sign_convention(ora, isign):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((16, 8)) + 1j * rng.standard_normal((16, 8))
    y = ora.fourn2(x, isign)
    ref = np.fft.ifft2(x) * x.size if isign == 1 else np.fft.fft2(x)
    assert np.abs(y - ref).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("shape,dx,dy", [((40, 27), 10.0, 12.0), ((32, 32), 5.0, 5.0), ((17, 50), 20.0, 7.0)])
def test_transform_matches_numpy(ora, shape, dx, dy):
    g = red_field(*shape, seed=1)
    kf = np.array([2e-5, 6e-5, 1.5e-4])
    w = ora.wlet_transform(g, dx, dy, kf, prec=64)
    ref = numpy_transform(g, dx, dy, kf, ora.K0_F32)
    assert w.shape == ref.shape == shape + (11, 3)
    for s in range(3):
        assert np.abs(w[..., s] - ref[..., s]).max() <= 1e-12 * np.abs(ref[..., s]).max()


def test_one_sample_shift(ora):
    # W[m,n] = IFFT2(psi*X)[m+1, n+1]: shifting is visible against an unshifted numpy evaluation
    g = red_field(24, 24, seed=5)
    kf = np.array([8e-5])
    w = ora.wlet_transform(g, 10.0, 10.0, kf, prec=64)[..., 0]
    ref = numpy_transform(g, 10.0, 10.0, kf, ora.K0_F32)[..., 0]
    assert np.abs(w - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.abs(w[:-1, :-1] - ref[1:, 1:]).max() > 1e-3 * np.abs(ref).max()


def test_float32_restatement_is_close(ora):
    g = red_field(48, 40, seed=2) * 500.0
    kf = np.array([1.5e-5, 4e-5, 9e-5])
    w64 = ora.wlet_transform(g, 10.0, 10.0, kf, prec=64)
    w32 = ora.wlet_transform(g, 10.0, 10.0, kf, prec=32)
    assert w32.dtype == np.complex64
    for s in range(3):
        assert np.abs(w32[..., s] - w64[..., s]).max() <= 1e-4 * np.abs(w64[..., s]).max()
    a64 = ora.wlet_admit_coh(w64, w64 * (0.2 - 0.1j) + w64[::-1] * 0.05, prec=64)
    a32 = ora.wlet_admit_coh(w64, w64 * (0.2 - 0.1j) + w64[::-1] * 0.05, prec=32)
    for x64, x32 in zip(a64, a32):
        assert np.abs(x32 - x64).max() <= 2e-4 * np.abs(x64).max()


def test_reductions_match_literal_python(ora):
    rng = np.random.default_rng(7)
    nx, ny, ns = 3, 2, 2
    w1 = rng.standard_normal((nx, ny, 11, ns)) + 1j * rng.standard_normal((nx, ny, 11, ns))
    w2 = 0.5 * w1 + 0.3 * (rng.standard_normal((nx, ny, 11, ns)) + 1j * rng.standard_normal((nx, ny, 11, ns)))
    sg, esg = ora.wlet_scalogram(w1)
    xsg, exsg = ora.wlet_xscalogram(w1, w2)
    adm, eadm, coh, ecoh = ora.wlet_admit_coh(w1, w2)
    for i in range(nx):
        for j in range(ny):
            for s in range(ns):
                s1 = np.abs(w1[i, j, :, s]) ** 2
                s2 = np.abs(w2[i, j, :, s]) ** 2
                x = w1[i, j, :, s] * np.conj(w2[i, j, :, s])
                assert np.isclose(sg[i, j, s], s1.mean(), rtol=1e-14)
                assert np.isclose(esg[i, j, s], numpy_jackknife(lambda p: p[0], [s1]), rtol=1e-12)
                assert np.isclose(xsg[i, j, s], x.mean(), rtol=1e-13)
                er = numpy_jackknife(lambda p: p[0], [x.real])
                ei = numpy_jackknife(lambda p: p[0], [x.imag])
                assert np.isclose(exsg[i, j, s], np.hypot(er, ei), rtol=1e-12)
                assert np.isclose(adm[i, j, s], x.real.mean() / s1.mean(), rtol=1e-13)
                assert np.isclose(coh[i, j, s], x.real.mean() ** 2 / (s1.mean() * s2.mean()), rtol=1e-13)
                assert np.isclose(eadm[i, j, s], numpy_jackknife(lambda p: p[2] / p[0], [s1, s2, x.real]), rtol=1e-11)
                assert np.isclose(ecoh[i, j, s], numpy_jackknife(lambda p: p[2] ** 2 / (p[0] * p[1]), [s1, s2, x.real]),
                                  rtol=1e-11)


def test_jackknife_zero_power_gives_zero_not_nan(ora):
    w = np.zeros((2, 2, 11, 1), complex)
    adm, eadm, coh, ecoh = ora.wlet_admit_coh(w, w)
    assert np.isnan(adm).all() and np.isnan(coh).all()  # 0/0 flows through (cpwt.f90:354-355)
    assert (eadm == 0).all() and (ecoh == 0).all()      # jackknife guards zero sums (cpwt_sub.f90:329-335)


def test_admit_coh_from_grids_equals_cube_path(ora):
    t, g = red_field(20, 24, 11) * 300, red_field(20, 24, 12) * 30
    kf = np.array([3e-5, 9e-5])
    a = ora.admit_coh_from_grids(t, g, 10.0, 10.0, kf)
    b = ora.wlet_admit_coh(ora.wlet_transform(t, 10.0, 10.0, kf), ora.wlet_transform(g, 10.0, 10.0, kf))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_flex_limits(ora):
    # Bouguer admittance -> 2*pi*G*(rhom-rhoc)*theta -> -2 pi G rhoc at long wavelength for top loading,
    # coherence 1 -> 0 from long to short wavelengths (flex.f90:115-169)
    k = np.array([1e-7, 1e-3])
    adm, coh = ora.real_xspec_functions(k, 30.0, 1e-6, np.pi / 2)
    assert np.isclose(adm[0], -2 * np.pi * 6.67e-6 * 2700.0 * np.exp(-1e-7 * 35e3), rtol=1e-4)
    _, coh = ora.real_xspec_functions(k, 30.0, 0.5, np.pi / 2)  # both loads present: coherence rolls over
    assert coh[0] > 0.999 and coh[1] < 1e-6
    # free air over land: short-wavelength admittance -> +2 pi G rhoc exp(-k*0)
    conf = ora.FlexConf(boug=0)
    adm, _ = ora.real_xspec_functions(np.array([5e-4]), 30.0, 1e-6, np.pi / 2, conf)
    assert np.isclose(adm[0], 2 * np.pi * 6.67e-6 * 2700.0, rtol=1e-3)


def test_flex_rhof_follows_water_depth(ora):
    conf = ora.FlexConf(wd=3000.0)
    ora.real_xspec_functions(np.array([1e-5]), 20.0, 0.3, 1.0, conf)
    assert conf.rhof == conf.rhow  # flex.f90:207-211
    conf.wd = 0.0
    ora.real_xspec_functions(np.array([1e-5]), 20.0, 0.3, 1.0, conf)
    assert conf.rhof == conf.rhoa


def test_oracle_matches_golden_fixtures(ora):
    z = np.load(os.path.join(GOLD, "cwt_small.npz"))
    w1 = ora.wlet_transform(z["topo"], float(z["dx"]), float(z["dy"]), z["k"], k0=float(z["k0"]))
    w2 = ora.wlet_transform(z["grav"], float(z["dx"]), float(z["dy"]), z["k"], k0=float(z["k0"]))
    assert np.abs(w1[:, :, z["ia_sel"], :] - z["wl_trans_topo_sel"]).max() <= 1e-12 * np.abs(z["wl_trans_topo_sel"]).max()
    for name, got in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), ora.wlet_admit_coh(w1, w2)):
        assert np.allclose(got, z[name], rtol=1e-10, atol=1e-13 * np.abs(z[name]).max()), name
    for name, got in zip(("wl_sg", "ewl_sg"), ora.wlet_scalogram(w1)):
        assert np.allclose(got, z[name], rtol=1e-10, atol=1e-13 * np.abs(z[name]).max()), name
    f = np.load(os.path.join(GOLD, "flex_curves.npz"))
    for c in range(len(f["te"])):
        conf = ora.FlexConf(wd=float(f["wd"][c]), boug=int(f["boug"]), rhoc=float(f["rhoc"]), zc=float(f["zc"]))
        a, co = ora.real_xspec_functions(f["k"], f["te"][c], f["f"][c], f["alpha"][c], conf)
        assert np.allclose(a, f["admit"][c], rtol=1e-13) and np.allclose(co, f["coh"][c], rtol=1e-13)


# ---- property tests (hypothesis): the oracle against the algebra of the reference's formulas -------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=12, deadline=None)
@given(nx=st.integers(3, 24), ny=st.integers(3, 24), seed=st.integers(0, 10_000), a=st.floats(-3.0, 3.0),
       b=st.floats(-3.0, 3.0))
def test_transform_is_linear_and_blind_to_the_mean(ora, nx, ny, seed, a, b):
    # cpwt.f90:110-111 removes the mean before the FFT; everything after is linear in the grid
    rng = np.random.default_rng(seed)
    g1, g2 = rng.standard_normal((nx, ny)), rng.standard_normal((nx, ny))
    kf = np.array([3e-5, 1.1e-4])
    w1, w2 = ora.wlet_transform(g1, 9.0, 11.0, kf), ora.wlet_transform(g2, 9.0, 11.0, kf)
    wc = ora.wlet_transform(a * g1 + b * g2 + 123.0, 9.0, 11.0, kf)
    scale = max(np.abs(w1).max() * abs(a), np.abs(w2).max() * abs(b), 1e-300)
    assert np.abs(wc - (a * w1 + b * w2)).max() <= 1e-11 * scale + 1e-13


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10_000), c=st.floats(0.01, 50.0))
def test_admittance_of_a_scaled_copy_is_the_scale_factor(ora, seed, c):
    # cpwt.f90:330-355: with grid 2 = c * grid 1, xsg = c*sg1 and sg2 = c^2*sg1 at every azimuth, so admittance = c,
    # coherence = 1 and every prefix ratio of the jackknife (cpwt_sub.f90:329-335) is the same: zero spread
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((12, 10))
    kf = np.array([4e-5, 9e-5])
    w = ora.wlet_transform(g, 10.0, 10.0, kf)
    adm, eadm, coh, ecoh = ora.wlet_admit_coh(w, c * w)
    assert np.allclose(adm, c, rtol=1e-12) and np.allclose(coh, 1.0, rtol=1e-12)
    assert np.abs(eadm).max() <= 1e-12 * c and np.abs(ecoh).max() <= 1e-12


@settings(max_examples=20, deadline=None)
@given(te=st.floats(2.0, 250.0), f=st.floats(0.001, 0.99), alpha=st.floats(0.01, 3.1), wd=st.floats(0.0, 6000.0),
       boug=st.integers(0, 1))
def test_flex_model_against_the_formulas_of_flex_f90(ora, te, f, alpha, wd, boug):
    # SURVEY.md appendix A8 / flex.f90:199-231 written out with numpy; rhof follows the water depth
    k = np.geomspace(3e-6, 2e-4, 9)
    conf = ora.FlexConf(wd=wd, boug=boug, rhoc=2750.0, zc=33.0e3)
    adm, coh = ora.real_xspec_functions(k, te, f, alpha, conf)
    g, em, nu, gc = 9.81, 1.0e11, 0.25, 6.67e-6
    rhom, rhoc, rhof = 3200.0, 2750.0, (1030.0 if wd > 0 else 0.0)
    A = 0.0 if boug == 1 else 1.0
    drho, rcf = rhom - rhoc, rhoc - rhof
    psi = em * (te * 1.0e3) ** 3 / 12.0 / (1.0 - nu ** 2) * k ** 4
    r = rcf / drho
    theta = -r / (1.0 + psi / drho / g)
    phi = -r * (1.0 + psi / rcf / g)
    mu_h, mu_w = 1.0 / (1.0 - theta), 1.0 / (phi - 1.0)
    top, deep = A * rcf * np.exp(-k * wd), drho * np.exp(-k * (33.0e3 + wd))
    nu_h = 2 * np.pi * gc * (top + deep * theta) / (1.0 - theta)
    nu_w = 2 * np.pi * gc * (top + deep * phi) / (phi - 1.0)
    ff = f / (1.0 - f) * r
    ea = np.exp(1j * alpha)
    hg = nu_h * mu_h + nu_w * mu_w * ff ** 2 + (nu_h * mu_w * ea + nu_w * mu_h * np.conj(ea)) * ff
    hh = mu_h ** 2 + (mu_w * ff) ** 2 + 2 * mu_h * mu_w * ff * np.cos(alpha)
    gg = nu_h ** 2 + (nu_w * ff) ** 2 + 2 * nu_h * nu_w * ff * np.cos(alpha)
    assert np.allclose(adm, np.real(hg) / hh, rtol=1e-10, atol=1e-16)
    assert np.allclose(coh, np.real(hg / np.sqrt(hh) / np.sqrt(gg)) ** 2, rtol=1e-10, atol=1e-16)


# ---- pin: an independent float32 restatement (tests/np32_ref.py) must agree BIT FOR BIT -------------------
@pytest.mark.parametrize("shape", [(64, 64), (32, 128), (2, 8)])
@pytest.mark.parametrize("isign", [1, -1])
def test_ora32_fourn_bit_exact_against_independent_numpy_float32(ora, shape, isign):
    """cpwt_sub.f90:522-594 written twice from the Fortran (C in oracle/, numpy float32 in tests/np32_ref.py):
    same bit-reversal, same double-precision twiddle recurrence truncated with sngl() per butterfly, same
    single-precision butterflies -- so every output bit must match."""
    import np32_ref
    rng = np.random.default_rng(5)
    d = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    re, im = np32_ref.fourn2(d, isign)
    o = ora.fourn2(d, isign, prec=32)
    assert np.array_equal(o.real, re) and np.array_equal(o.imag, im)


def test_ora32_admit_coh_and_jackknife_bit_exact_against_independent_numpy_float32(ora):
    """cpwt.f90:330-361 + cpwt_sub.f90:298-366 (azimuth means, admittance, coherence, running-prefix jackknife
    with its zero guard and NaN -> 0) restated in numpy float32: all four maps bit for bit, including a cell
    whose topography coefficients are all zero (NaN admittance, zero errors)."""
    import np32_ref
    rng = np.random.default_rng(6)
    shape = (64, 64, 11, 2)
    w1 = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    w2 = (0.3 * w1 + rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    w1[3, 4, :, 0] = 0
    ref = np32_ref.admit_coh(w1, w2)
    got = ora.wlet_admit_coh(w1, w2, prec=32)
    for name, a, b in zip(("wl_admit", "ewl_admit", "wl_coh", "ewl_coh"), ref, got):
        assert np.array_equal(a, b, equal_nan=True), name
    assert np.isnan(got[0][3, 4, 0]) and got[1][3, 4, 0] == 0
