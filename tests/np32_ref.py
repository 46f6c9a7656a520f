"""A second, independent restatement of the reference's float32 arithmetic, written from the Fortran in
numpy float32 (no shared code with oracle/): Numerical-Recipes ``fourn`` (cpwt_sub.f90:522-594) and
``wlet_admit_coh`` + ``jackknife_admit_coh`` (cpwt.f90:330-361, cpwt_sub.f90:298-366).

With no Fortran compiler in the image this is the strongest pin available for the oracle: the float32 build of
the C oracle (``ora32_*``) must reproduce these results BIT FOR BIT (tests/test_oracle.py).  numpy float32
array operations round every product, sum and quotient to single precision exactly like default-kind REAL
arithmetic compiled without fused multiply-add or extended precision."""
import numpy as np

F32 = np.float32


def fourn_axis(re, im, axis, isign):
    """One dimension of fourn on float32 arrays re / im (modified in place): bit-reversal permutation along
    ``axis`` (the i2rev loop, cpwt_sub.f90:541-563), then the Danielson-Lanczos stages with the trigonometric
    recurrence carried in DOUBLE PRECISION and truncated with sngl() at every butterfly (:564-589)."""
    n = re.shape[axis]
    bits = n.bit_length() - 1
    assert 1 << bits == n
    rev = np.array([int(format(i, "0%db" % bits)[::-1], 2) if bits else 0 for i in range(n)])
    re[...] = np.take(re, rev, axis=axis)
    im[...] = np.take(im, rev, axis=axis)
    re_m, im_m = np.moveaxis(re, axis, 0), np.moveaxis(im, axis, 0)  # views
    half = 1
    while half < n:  # ifp1 = ip1*half, ifp2 = 2*ifp1
        theta = isign * 6.28318530717959 / (2 * half)
        wpr = -2.0 * np.sin(0.5 * theta) ** 2
        wpi = np.sin(theta)
        wr, wi = 1.0, 0.0
        for m in range(half):  # i3 loop
            k1 = np.arange(m, n, 2 * half)
            k2 = k1 + half
            swr, swi = F32(wr), F32(wi)  # sngl(wr), sngl(wi)
            tempr = swr * re_m[k2] - swi * im_m[k2]
            tempi = swr * im_m[k2] + swi * re_m[k2]
            re_m[k2] = re_m[k1] - tempr
            im_m[k2] = im_m[k1] - tempi
            re_m[k1] = re_m[k1] + tempr
            im_m[k1] = im_m[k1] + tempi
            wtemp = wr
            wr = wr * wpr - wi * wpi + wr
            wi = wi * wpr + wtemp * wpi + wi
        half *= 2


def fourn2(data, isign):
    """2-D fourn of a complex (n1, n2) array, first axis fastest as in the Fortran: dimension 1 first."""
    re = np.ascontiguousarray(data.real, dtype=F32).copy()
    im = np.ascontiguousarray(data.imag, dtype=F32).copy()
    fourn_axis(re, im, 0, isign)
    fourn_axis(re, im, 1, isign)
    return re, im


def admit_coh(w1, w2):
    """wlet_admit_coh on (nx, ny, na, ns) complex64 cubes -> wl_admit, ewl_admit, wl_coh, ewl_coh (float32)."""
    a1, b1 = w1.real.astype(F32), w1.imag.astype(F32)
    a2, b2 = w2.real.astype(F32), w2.imag.astype(F32)
    na = w1.shape[2]
    # w*CONJG(w) = (a*a - b*(-b)) + i(a*(-b) + b*a); CABS of a number whose imaginary part is exactly 0
    sg1 = np.abs(a1 * a1 - b1 * (-b1))
    sg2 = np.abs(a2 * a2 - b2 * (-b2))
    xr = a1 * a2 - b1 * (-b2)  # REAL(w1*CONJG(w2))
    m1 = np.zeros(sg1.shape[:2] + sg1.shape[3:], F32)
    m2, mx = m1.copy(), m1.copy()
    for ia in range(na):  # accumulation order of cpwt.f90:339-341
        m1 = m1 + sg1[:, :, ia]
        m2 = m2 + sg2[:, :, ia]
        mx = mx + xr[:, :, ia]
    m1, m2, mx = m1 / F32(na), m2 / F32(na), mx / F32(na)
    with np.errstate(all="ignore"):
        wl_admit = mx / m1
        wl_coh = (mx * mx) / (m1 * m2)
        # running-prefix jackknife, cpwt_sub.f90:310-357
        ad2, co2 = [], []
        for ia in range(na):
            p1 = np.zeros_like(m1)
            p2, xp = p1.copy(), p1.copy()
            sad, sco = p1.copy(), p1.copy()
            for ia2 in range(na):
                if ia2 == ia:
                    continue
                p1 = p1 + sg1[:, :, ia2]
                p2 = p2 + sg2[:, :, ia2]
                xp = xp + xr[:, :, ia2]
                ok = (p1 != 0) & (p2 != 0)
                co = np.where(ok, (xp * xp) / (p1 * p2), F32(0))
                ad = np.where(ok, xp / p1, F32(0))
                sco = sco + co  # SUM(co): left to right
                sad = sad + ad
            co2.append(sco / F32(na - 1))
            ad2.append(sad / F32(na - 1))

        def spread(v2):
            m = np.zeros_like(v2[0])
            for v in v2:
                m = m + v
            m = m / F32(na)
            d = np.zeros_like(m)
            for v in v2:
                d = d + (v - m) * (v - m)
            d = np.sqrt(d * F32(na - 1) / F32(na))
            return np.where(np.isnan(d), F32(0), d)

        return wl_admit, spread(ad2), wl_coh, spread(co2)
