"""Stand-in for the f2py extension module ``plateflex.flex`` (built by the
reference from src/flex/flex.f90, meson.build:46-58).

``from plateflex_b200.flex import flex, conf_flex``: ``conf_flex`` carries the
module variables of flex.f90:32-50 under their f2py (lower-case) names and
``flex.real_xspec_functions(k, te, f, alpha)`` evaluates flex.f90:178-235 on
the GPU through the C ABI (pfx_flex_xspec_host).
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["flex", "conf_flex"]


class _ConfFlex:
    """conf_flex (flex.f90:32-50); defaults are installed by plateflex_b200.set_conf_flex()."""

    # hard-coded parameters, flex.f90:39-43 (f2py --lower: Em -> em, Gc -> gc)
    pi = 3.141592653589793
    g = 9.81
    em = 1.e11
    nu = 0.25
    gc = 6.67e-6

    def __init__(self):
        self.rhoc = self.rhom = self.rhof = self.rhow = self.rhoa = self.wd = self.zc = 0.0
        self.boug = 0

    def as_struct(self, rhoc=None, zc=None):
        return _lib.FlexConf(rhoc=float(self.rhoc if rhoc is None else rhoc), rhom=float(self.rhom),
                             rhow=float(self.rhow), rhoa=float(self.rhoa),
                             zc=float(self.zc if zc is None else zc), boug=int(self.boug), pad_=0)


conf_flex = _ConfFlex()


class _Flex:
    """The ``flex`` Fortran module (flex.f90:64-237) as seen through f2py."""

    def real_xspec_functions(self, k, te, f, alpha):
        """flex.f90:178-235: k (ns) [rad/m], Te [km], F, alpha [rad] -> (admit, coh), float64 (ns).
        Reads rhoc, rhom, rhow, rhoa, wd, zc, boug from ``conf_flex`` and, like the Fortran
        (flex.f90:207-211), leaves ``conf_flex.rhof`` set from the sign of ``wd``."""
        adm, coh = self.real_xspec_batch(k, [te], [f], [alpha], [conf_flex.wd])
        conf_flex.rhof = conf_flex.rhow if conf_flex.wd > 0.0 else conf_flex.rhoa
        return adm[0], coh[0]

    # -- the helper subroutines f2py also exports (flex.f90:74-169); the package itself never calls them
    def _helper(self, op, ins, p0=0.0, p1=0.0, nout=1):
        ins = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in ins]
        ns = len(ins[0])
        if any(len(a) != ns for a in ins):
            raise ValueError("input arrays must have the same length")
        outs = [np.empty(ns) for _ in range(nout)]
        conf = conf_flex.as_struct()
        pin = [_lib.ptr(a) for a in ins] + [None] * (4 - len(ins))
        pout = [_lib.ptr(a) for a in outs] + [None] * (4 - nout)
        _lib.check(_lib.lib.pfx_flex_helper_host(_lib.default_device(), C.byref(conf), float(conf_flex.rhof),
                                                 float(conf_flex.wd), op, ns, *pin, float(p0), float(p1), *pout))
        return outs

    def flexfilter_top(self, psi):
        """flex.f90:74-86 (reads rhoc, rhom, rhof from conf_flex as the Fortran does)."""
        return self._helper(0, [psi])[0]

    def flexfilter_bot(self, psi):
        """flex.f90:94-106."""
        return self._helper(1, [psi])[0]

    def decon(self, theta, phi, k, a):
        """flex.f90:115-132 -> (mu_h, mu_w, nu_h, nu_w)."""
        return tuple(self._helper(2, [theta, phi, k], p0=a, nout=4))

    def tr_func(self, mu_h, mu_w, nu_h, nu_w, f, alpha):
        """flex.f90:141-169 -> (admit complex128, coh)."""
        re, im, coh = self._helper(3, [mu_h, mu_w, nu_h, nu_w], p0=f, p1=alpha, nout=3)
        return re + 1j * im, coh

    def real_xspec_batch(self, k, te, f, alpha, wd=None, jac=False, device=None):
        """Many parameter sets at once: returns (admit, coh) of shape (ncurves, ns); with
        ``jac`` also the analytic d/d(Te, F, alpha) arrays of shape (ncurves, 3, ns)."""
        k = np.ascontiguousarray(k, dtype=np.float64).ravel()
        te = np.ascontiguousarray(te, dtype=np.float64).ravel()
        f = np.ascontiguousarray(f, dtype=np.float64).ravel()
        alpha = np.ascontiguousarray(alpha, dtype=np.float64).ravel()
        n, ns = len(te), len(k)
        if not (len(f) == n and len(alpha) == n):
            raise ValueError("te, f and alpha must have the same length")
        wd_arr = None if wd is None else np.ascontiguousarray(wd, dtype=np.float64).ravel()
        if wd_arr is not None and len(wd_arr) != n:
            raise ValueError("wd must have one value per parameter set")
        adm, coh = np.empty((n, ns)), np.empty((n, ns))
        ja = np.empty((n, 3, ns)) if jac else None
        jc = np.empty((n, 3, ns)) if jac else None
        conf = conf_flex.as_struct()
        dev = _lib.default_device() if device is None else device
        _lib.check(_lib.lib.pfx_flex_xspec_host(dev, C.byref(conf), _lib.ptr(k), ns, n, _lib.ptr(te), _lib.ptr(f),
                                                _lib.ptr(alpha), _lib.ptr(wd_arr), _lib.ptr(adm), _lib.ptr(coh),
                                                _lib.ptr(ja), _lib.ptr(jc)))
        return (adm, coh, ja, jc) if jac else (adm, coh)


flex = _Flex()
