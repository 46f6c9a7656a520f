"""Shared helpers of the test-suite (inputs, tolerances, numpy restatement)."""
import numpy as np

# Parity tolerances.  CWT-derived maps: per scale plane max|d|/max|ref| <= 1e-9
# and element-wise rtol 1e-9 with atol 1e-12*max|ref|.
PLANE_TOL = 1e-9
RTOL = 1e-9
ATOL_REL = 1e-12
# fit tolerances against SciPy TRF on the oracle model (interior optimum)
TE_TOL, F_TOL, A_TOL, CHI2_RTOL, STD_RTOL = 0.1, 1e-3, 1e-2, 1e-4, 1e-2


def assert_maps_close(got, ref, name="", plane_tol=PLANE_TOL, rtol=RTOL, atol_rel=ATOL_REL, elementwise=True, floor=None):
    """floor (optional, same shape): an absolute accuracy floor per element added to the element-wise bound."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{name}: NaN pattern differs"
    for s in range(ref.shape[-1]):
        g, r = got[..., s], ref[..., s]
        ok = np.isfinite(r)
        if not ok.any():
            continue
        mx = np.abs(r[ok]).max()
        if mx == 0:
            assert np.abs(g[ok]).max() == 0, f"{name}[{s}] expected zeros"
            continue
        err = np.abs(g[ok] - r[ok])
        assert err.max() / mx <= plane_tol, f"{name}[..., {s}]: max|d|/max|ref| = {err.max() / mx:.3e}"
        extra = floor[..., s][ok] if floor is not None else 0.0
        assert not elementwise or np.all(err <= rtol * np.abs(r[ok]) + atol_rel * mx + extra), \
            f"{name}[..., {s}]: element-wise tolerance exceeded ({(err / (np.abs(r[ok]) + 1e-300)).max():.3e})"


def red_field(nx, ny, seed, slope=-1.5):
    """A smooth-ish random field with a red spectrum (non-periodic after cropping)."""
    rng = np.random.default_rng(seed)
    n = 2 * max(nx, ny)
    kx = np.fft.fftfreq(n)[:, None]
    ky = np.fft.fftfreq(n)[None, :]
    k = np.sqrt(kx ** 2 + ky ** 2)
    k[0, 0] = 1.0
    f = np.real(np.fft.ifft2(np.fft.fft2(rng.standard_normal((n, n))) * k ** slope))
    f = f[:nx, :ny]
    return np.ascontiguousarray(f / f.std())


def numpy_transform(grid, dx, dy, kf, k0):
    """Independent numpy.fft restatement of cpwt.wlet_transform."""
    nx, ny = grid.shape

    def npow2(n):
        p = 2
        while p < n:
            p *= 2
        return p

    NX, NY = npow2(2 * nx), npow2(2 * ny)
    m = np.zeros((2 * nx, 2 * ny))
    m[:nx, :ny] = grid
    m[nx:, :ny] = grid[::-1, :]
    m[:nx, ny:] = grid[:, ::-1]
    m[nx:, ny:] = grid[::-1, ::-1]
    xp = np.zeros((NX, NY))
    xp[:2 * nx, :2 * ny] = m - m.sum() / (4 * nx * ny)
    X = np.fft.fft2(xp)

    def kax(n, d):
        i = np.arange(n)
        return 2 * np.pi * np.where(i <= n // 2, i, i - n).astype(float) / (n * d)

    kx, ky = kax(NX, dx), kax(NY, dy)
    da = 2 * np.sqrt(-2 * np.log(0.75)) / k0
    out = np.zeros((nx, ny, 11, len(kf)), complex)
    for i_s, k in enumerate(kf):
        s = k0 / (k * 1e3)
        for ia in range(11):
            a = ia * da - np.pi / 2
            kx0, ky0 = k0 * np.cos(a), k0 * np.sin(a)
            psi = s / np.sqrt(np.pi) * (
                np.exp(-0.5 * (s * kx[:, None] - kx0) ** 2) * np.exp(-0.5 * (s * ky[None, :] - ky0) ** 2)
                - np.exp(-0.5 * (kx[:, None] ** 2 + ky[None, :] ** 2 + kx0 ** 2 + ky0 ** 2)))
            w = np.roll(np.fft.ifft2(psi * X), (-1, -1), (0, 1))
            out[:, :, ia, i_s] = w[:nx, :ny]
    return out


def numpy_jackknife(vals_fn, parts):
    """Literal running-prefix jackknife for one cell.
    parts: list of per-azimuth arrays; vals_fn(prefix sums) -> statistic."""
    na = len(parts[0])
    v2 = []
    for a in range(na):
        pre = [0.0] * len(parts)
        rec = []
        for a2 in range(na):
            if a2 == a:
                continue
            pre = [p + q[a2] for p, q in zip(pre, parts)]
            rec.append(vals_fn(pre))
        v2.append(sum(rec) / (na - 1))
    m = sum(v2) / na
    d = np.sqrt(sum((v - m) ** 2 for v in v2) * (na - 1) / na)
    return 0.0 if np.isnan(d) else d
