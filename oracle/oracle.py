"""ctypes front-end of the CPU oracle (oracle/libpfx_oracle.so).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs.  Nothing under
plateflex_b200/ imports this module.

PARITY STATUS: "parity unpinned" for the numeric CWT / jackknife / flexure
values (no golden vectors in the reference, Fortran not buildable here); see
oracle/pfx_oracle.c.

The Python-visible signatures follow what numpy.f2py generates for the
reference: ``wlet_transform(grid, dx, dy, kf)`` returns a
Fortran-ordered (nx, ny, 11, ns) complex cube, and so on.  ``prec`` selects the
float32 restatement (the reference's own precision) or the float64 one (the
parity target of the CUDA path).
"""
import ctypes as C
import os
import subprocess

import numpy as np

NA = 11  # cpwt.f90:37
K0_F32 = float(np.float32(5.336))  # conf_cpwt.k0 is a float32 module variable (cpwt.f90:41)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile libpfx_oracle.so with the committed Makefile (gcc -O2 -fopenmp)."""
    so = os.path.join(_HERE, "libpfx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pfx_oracle.c", "pfx_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libpfx_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


class FlexConf(C.Structure):
    """conf_flex module variables (flex.f90:47-48); defaults from plateflex/__init__.py:150-158."""

    _fields_ = [(n, C.c_double) for n in ("rhoc", "rhom", "rhof", "rhow", "rhoa", "wd", "zc")] + [("boug", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        self.zc, self.rhom, self.rhoc, self.rhow, self.rhoa = 35.0e3, 3200.0, 2700.0, 1030.0, 0.0
        self.rhof, self.wd, self.boug = 0.0, 0.0, 1
        for k, v in kw.items():
            setattr(self, k, v)


def _rt(prec):
    if prec == 32:
        return np.float32, np.complex64, C.c_float, "ora32_"
    if prec == 64:
        return np.float64, np.complex128, C.c_double, "ora64_"
    raise ValueError("prec must be 32 or 64")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def npow2(n):
    return int(lib().ora64_npow2(C.c_int(n)))


def defk(nx, ny, dx, dy, prec=64):
    rt, _, cr, pre = _rt(prec)
    kx, ky = np.empty(nx, rt), np.empty(ny, rt)
    getattr(lib(), pre + "defk")(C.c_int(nx), C.c_int(ny), cr(dx), cr(dy), _p(kx), _p(ky))
    return kx, ky


def wave_function(kx, ky, k0, s, a, prec=64):
    rt, ct, cr, pre = _rt(prec)
    kx, ky = np.ascontiguousarray(kx, rt), np.ascontiguousarray(ky, rt)
    out = np.empty((len(kx), len(ky)), ct, order="F")
    getattr(lib(), pre + "wave_function")(C.c_int(len(kx)), C.c_int(len(ky)), _p(kx), _p(ky), cr(k0), cr(s), cr(a),
                                          _p(out))
    return out


def fourn2(data, isign, prec=64):
    """2-D ``fourn`` of a complex (n1, n2) array (first axis fastest, as in Fortran)."""
    _, ct, _, pre = _rt(prec)
    a = np.array(data, dtype=ct, order="F", copy=True)
    nn = (C.c_int * 2)(a.shape[0], a.shape[1])
    getattr(lib(), pre + "fourn")(_p(a), nn, C.c_int(2), C.c_int(isign))
    return a


def wlet_transform(grid, dx, dy, kf, k0=K0_F32, prec=64, nthreads=1):
    rt, ct, cr, pre = _rt(prec)
    g = np.asfortranarray(grid, dtype=rt)
    kf = np.ascontiguousarray(kf, rt)
    nx, ny = g.shape
    out = np.empty((nx, ny, NA, len(kf)), ct, order="F")
    getattr(lib(), pre + "wlet_transform")(_p(g), C.c_int(nx), C.c_int(ny), cr(dx), cr(dy), _p(kf), C.c_int(len(kf)),
                                           cr(k0), _p(out), C.c_int(nthreads))
    return out


def wlet_planes(grid, dx, dy, kscale, ia0, ia1, k0=K0_F32, prec=32, nthreads=1):
    """Timing sample: azimuth planes [ia0, ia1) of one scale -> ((nx, ny, ia1-ia0) complex, seconds spent in
    the forward stage that a full job pays once per grid)."""
    rt, ct, cr, pre = _rt(prec)
    g = np.asfortranarray(grid, dtype=rt)
    nx, ny = g.shape
    out = np.empty((nx, ny, ia1 - ia0), ct, order="F")
    tf = C.c_double(0.0)
    getattr(lib(), pre + "wlet_planes")(_p(g), C.c_int(nx), C.c_int(ny), cr(dx), cr(dy), cr(kscale), cr(k0), C.c_int(ia0),
                                        C.c_int(ia1), _p(out), C.c_int(nthreads), C.byref(tf))
    return out, tf.value


def _cube(w, ct):
    w = np.asfortranarray(w, dtype=ct)
    assert w.ndim == 4 and w.shape[2] == NA
    return w


def wlet_scalogram(wl_trans, prec=64):
    rt, ct, _, pre = _rt(prec)
    w = _cube(wl_trans, ct)
    nx, ny, _, ns = w.shape
    sg, esg = np.empty((nx, ny, ns), rt, order="F"), np.empty((nx, ny, ns), rt, order="F")
    getattr(lib(), pre + "wlet_scalogram")(_p(w), C.c_int(nx), C.c_int(ny), C.c_int(ns), _p(sg), _p(esg))
    return sg, esg


def wlet_xscalogram(wl_trans1, wl_trans2, prec=64):
    rt, ct, _, pre = _rt(prec)
    w1, w2 = _cube(wl_trans1, ct), _cube(wl_trans2, ct)
    nx, ny, _, ns = w1.shape
    xsg, exsg = np.empty((nx, ny, ns), ct, order="F"), np.empty((nx, ny, ns), rt, order="F")
    getattr(lib(), pre + "wlet_xscalogram")(_p(w1), _p(w2), C.c_int(nx), C.c_int(ny), C.c_int(ns), _p(xsg), _p(exsg))
    return xsg, exsg


def wlet_admit_coh(wl_trans1, wl_trans2, prec=64):
    rt, ct, _, pre = _rt(prec)
    w1, w2 = _cube(wl_trans1, ct), _cube(wl_trans2, ct)
    nx, ny, _, ns = w1.shape
    outs = [np.empty((nx, ny, ns), rt, order="F") for _ in range(4)]
    getattr(lib(), pre + "wlet_admit_coh")(_p(w1), _p(w2), C.c_int(nx), C.c_int(ny), C.c_int(ns), *[_p(o) for o in outs])
    return tuple(outs)  # wl_admit, ewl_admit, wl_coh, ewl_coh


def admit_coh_from_grids(topo, grav, dx, dy, kf, k0=K0_F32, prec=64, nthreads=1):
    """wlet_transform x2 + wlet_admit_coh, one scale at a time (no full cube in RAM)."""
    rt, _, cr, pre = _rt(prec)
    t, g = np.asfortranarray(topo, dtype=rt), np.asfortranarray(grav, dtype=rt)
    kf = np.ascontiguousarray(kf, rt)
    nx, ny = t.shape
    outs = [np.empty((nx, ny, len(kf)), rt, order="F") for _ in range(4)]
    getattr(lib(), pre + "admit_coh_from_grids")(_p(t), _p(g), C.c_int(nx), C.c_int(ny), cr(dx), cr(dy), _p(kf),
                                                 C.c_int(len(kf)), cr(k0), *[_p(o) for o in outs], C.c_int(nthreads))
    return tuple(outs)


def real_xspec_functions(k, Te, F, alpha, conf=None):
    """flex.real_xspec_functions (flex.f90:178-235) -> (admit, coh), float64."""
    conf = conf if conf is not None else FlexConf()
    k = np.ascontiguousarray(k, np.float64)
    adm, coh = np.empty_like(k), np.empty_like(k)
    lib().ora_flex_real_xspec_functions(C.byref(conf), C.c_int(len(k)), _p(k), C.c_double(Te), C.c_double(F),
                                        C.c_double(alpha), _p(adm), _p(coh))
    return adm, coh
