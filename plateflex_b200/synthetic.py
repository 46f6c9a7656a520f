"""Synthetic topography / gravity pairs for tests and bench.py.

Two independent fractal (beta = 3) random fields are taken as the initial
surface and subsurface loads of a thin elastic plate of known (Te, F); the
final topography and gravity anomaly follow from the deconvolution
coefficients of the flexural model (the same algebra as flex.f90:115-132,
evaluated here in numpy at |k| of the FFT grid, uncorrelated loads).  Pure host
code: it only manufactures inputs and is not on the measured path.
"""
import numpy as np

PI = 3.141592653589793
G_ACC, EM, NU, GC = 9.81, 1.e11, 0.25, 6.67e-6  # flex.f90:39-43


def decon_coefficients(k, Te_km, rhoc=2700., rhom=3200., rhof=0., zc=35.e3, wd=0., boug=True):
    A = 0. if boug else 1.
    D = EM * (Te_km * 1.e3) ** 3 / 12. / (1. - NU ** 2)
    psi = D * k ** 4
    r = (rhoc - rhof) / (rhom - rhoc)
    theta = -r / (1. + psi / (rhom - rhoc) / G_ACC)
    phi = -r * (1. + psi / (rhoc - rhof) / G_ACC)
    mu_h, mu_w = 1. / (1. - theta), 1. / (phi - 1.)
    top, deep = A * (rhoc - rhof) * np.exp(-k * wd), (rhom - rhoc) * np.exp(-k * (zc + wd))
    nu_h = 2. * PI * GC * (top + deep * theta) / (1. - theta)
    nu_w = 2. * PI * GC * (top + deep * phi) / (phi - 1.)
    return mu_h, mu_w, nu_h, nu_w, r


def make_pair(nx, ny, dx_km, dy_km, seed, Te=40., F=0.5, boug=True, rhoc=2700., rhom=3200., zc=35.e3, ocean_depth=None,
              topo_std=500.):
    """Returns (topo [m], grav [mGal]) float64 (nx, ny) arrays.

    Land case (ocean_depth None): topography is shifted to be strictly positive so that TopoGrid
    finds no water (classes.py:533-536).  Ocean case: mean depth ``ocean_depth`` (m, positive
    number), bathymetry clipped below -100 m, loads computed with rhof = 1030 and that mean depth.
    """
    rng = np.random.default_rng(seed)
    kx = 2. * PI * np.fft.fftfreq(nx, d=dx_km * 1.e3)
    ky = 2. * PI * np.fft.fftfreq(ny, d=dy_km * 1.e3)
    k = np.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)
    k[0, 0] = 1.
    amp = k ** -1.5
    amp[0, 0] = 0.
    hi = np.fft.fft2(rng.standard_normal((nx, ny))) * amp
    wi = np.fft.fft2(rng.standard_normal((nx, ny))) * amp
    rhof, wd = (1030., float(ocean_depth)) if ocean_depth is not None else (0., 0.)
    mu_h, mu_w, nu_h, nu_w, r = decon_coefficients(k, Te, rhoc, rhom, rhof, zc, wd, boug)
    wi *= (F / (1. - F)) * r
    topo = np.real(np.fft.ifft2(mu_h * hi + mu_w * wi))
    grav = np.real(np.fft.ifft2(nu_h * hi + nu_w * wi))
    scale = topo_std / topo.std()
    topo *= scale
    grav *= scale
    if ocean_depth is None:
        topo += 10. - topo.min()
    else:
        topo = np.minimum(topo - float(ocean_depth), -100.)
    return np.ascontiguousarray(topo), np.ascontiguousarray(grav)
