"""Stand-in for the f2py extension module ``plateflex.cpwt`` (built by the
reference from src/cpwt/cpwt.f90 + cpwt_sub.f90, meson.build:60-75).

``from plateflex_b200.cpwt import cpwt, conf_cpwt`` gives two objects with the
attribute names f2py generates (lower-case): the four drivers on ``cpwt`` and
the module variables ``k0`` / ``na`` / ``pi`` on ``conf_cpwt``.  Every call goes
through the C ABI of libplateflex_b200.so and runs on the GPU.

Differences from the f2py module, all documented in DESIGN.md:
  * arithmetic and outputs are float64 / complex128 (the Fortran is float32);
  * inputs are not truncated to float32;
  * no progress text is written to stdout (cpwt.f90:142,152,189,360).
Output arrays are Fortran-ordered exactly like f2py's ``intent(out)`` arrays.
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["cpwt", "conf_cpwt", "fused", "set_output_dtype", "get_output_dtype"]

# dtype of the (nx, ny, ns) maps returned by the fused grid-in forms: float64 (default: what the kernels
# compute) or float32 (what numpy.f2py returns for the reference's REAL arrays; rounded once on the device,
# half the device-to-host copy)
_OUT_DTYPE = np.dtype(np.float64)


def set_output_dtype(dtype):
    global _OUT_DTYPE
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("output dtype must be float32 or float64")
    _OUT_DTYPE = dt


def get_output_dtype():
    return _OUT_DTYPE


class _ConfCpwt:
    """conf_cpwt (cpwt.f90:32-43).  ``k0`` is a float32 Fortran variable: a value written
    from Python is rounded to float32 and reads back as such (5.336 -> 5.335999965667725),
    which is what classes._lam2k sees (classes.py:1811)."""

    pi = float(np.float32(3.141592653589793))  # REAL, PARAMETER (cpwt.f90:36)
    na = _lib.NA  # cpwt.f90:37

    def __init__(self):
        self._k0 = np.float32(0.0)

    @property
    def k0(self):
        return float(self._k0)

    @k0.setter
    def k0(self, value):
        self._k0 = np.float32(value)


conf_cpwt = _ConfCpwt()


def _grid(a, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError(f"{name} must be a 2-D array, got shape {a.shape}")  # f2py: rank mismatch
    return a


def _cube(a, name):
    a = np.asarray(a)
    if a.ndim != 4 or a.shape[2] != _lib.NA:
        raise ValueError(f"{name} must have shape (nx, ny, {_lib.NA}, ns), got {a.shape}")
    return np.asfortranarray(a, dtype=np.complex128)


def _maps(nx, ny, ns, n, dtype=np.float64):
    return [np.empty((nx, ny, ns), dtype=dtype, order="F") for _ in range(n)]


class _Cpwt:
    """The ``cpwt`` Fortran module (cpwt.f90:54-366) as seen through f2py."""

    def wlet_transform(self, grid, dx, dy, kf):
        """cpwt.f90:65-191.  grid (nx,ny), dx, dy [km], kf (ns) [rad/m] -> complex (nx,ny,11,ns)."""
        g = _grid(grid, "grid")
        kf = np.ascontiguousarray(kf, dtype=np.float64).ravel()
        nx, ny = g.shape
        plan = _lib.get_plan(nx, ny, dx, dy, kf, conf_cpwt.k0)
        c64 = _OUT_DTYPE == np.float32  # complex64 cube, as numpy.f2py returns the reference's COMPLEX array
        out = np.empty((nx, ny, _lib.NA, len(kf)), dtype=np.complex64 if c64 else np.complex128, order="F")
        fn = _lib.lib.pfx_wlet_transform_host_c64 if c64 else _lib.lib.pfx_wlet_transform_host
        _lib.check(fn(plan.handle, _lib.ptr(g), 0, len(kf), _lib.ptr(out)))
        return out

    def wlet_scalogram(self, wl_trans):
        """cpwt.f90:200-241.  complex (nx,ny,11,ns) -> (wl_sg, ewl_sg), each (nx,ny,ns)."""
        w = _cube(wl_trans, "wl_trans")
        nx, ny, _, ns = w.shape
        sg, esg = _maps(nx, ny, ns, 2)
        _lib.check(_lib.lib.pfx_cube_scalogram_host(_lib.default_device(), nx, ny, ns, _lib.ptr(w), _lib.ptr(sg),
                                                    _lib.ptr(esg)))
        return sg, esg

    def wlet_xscalogram(self, wl_trans1, wl_trans2):
        """cpwt.f90:250-293.  -> (wl_xsg complex (nx,ny,ns), ewl_xsg (nx,ny,ns))."""
        w1, w2 = _cube(wl_trans1, "wl_trans1"), _cube(wl_trans2, "wl_trans2")
        if w1.shape != w2.shape:
            raise ValueError("wl_trans1 and wl_trans2 must have the same shape")
        nx, ny, _, ns = w1.shape
        xsg = np.empty((nx, ny, ns), dtype=np.complex128, order="F")
        (exsg,) = _maps(nx, ny, ns, 1)
        _lib.check(_lib.lib.pfx_cube_xscalogram_host(_lib.default_device(), nx, ny, ns, _lib.ptr(w1), _lib.ptr(w2),
                                                     _lib.ptr(xsg), _lib.ptr(exsg)))
        return xsg, exsg

    def wlet_admit_coh(self, wl_trans1, wl_trans2):
        """cpwt.f90:303-364.  -> (wl_admit, ewl_admit, wl_coh, ewl_coh), each (nx,ny,ns)."""
        w1, w2 = _cube(wl_trans1, "wl_trans1"), _cube(wl_trans2, "wl_trans2")
        if w1.shape != w2.shape:
            raise ValueError("wl_trans1 and wl_trans2 must have the same shape")
        nx, ny, _, ns = w1.shape
        outs = _maps(nx, ny, ns, 4)
        _lib.check(_lib.lib.pfx_cube_admit_coh_host(_lib.default_device(), nx, ny, ns, _lib.ptr(w1), _lib.ptr(w2),
                                                    *[_lib.ptr(o) for o in outs]))
        return tuple(outs)


    # -- helper subroutines of cpwt_sub.f90 that f2py exports at module level (unused by the package)
    def npow2(self, n):
        """cpwt_sub.f90:26-33."""
        return int(_lib.lib.pfx_npow2(int(n)))

    def defk(self, nx, ny, dx, dy):
        """cpwt_sub.f90:41-62 -> (kx, ky), float64, positive Nyquist."""
        kx, ky = np.empty(int(nx)), np.empty(int(ny))
        _lib.check(_lib.lib.pfx_defk_host(_lib.default_device(), int(nx), int(ny), float(dx), float(dy), _lib.ptr(kx),
                                          _lib.ptr(ky)))
        return kx, ky

    def wave_function(self, kx, ky, k0, scales, angle):
        """cpwt_sub.f90:73-99 -> daughter wavelet (nx, ny) complex128, Fortran order."""
        kx = np.ascontiguousarray(kx, dtype=np.float64).ravel()
        ky = np.ascontiguousarray(ky, dtype=np.float64).ravel()
        out = np.empty((len(kx), len(ky)), dtype=np.complex128, order="F")
        _lib.check(_lib.lib.pfx_wave_function_host(_lib.default_device(), len(kx), len(ky), _lib.ptr(kx), _lib.ptr(ky),
                                                   float(k0), float(scales), float(angle), _lib.ptr(out)))
        return out


cpwt = _Cpwt()


class _Fused:
    """Grid-in forms that never materialise the (nx,ny,11,ns) cube: what Grid.wlet_scalogram and
    Project.wlet_admit_coh use when no ``wl_trans`` attribute exists (classes.py:276-280, 950-967)."""

    def wlet_scalogram(self, grid, dx, dy, kf, scales=None):
        g = _grid(grid, "grid")
        kf = np.ascontiguousarray(kf, dtype=np.float64).ravel()
        nx, ny = g.shape
        is0, is1 = scales if scales is not None else (0, len(kf))
        plan = _lib.get_plan(nx, ny, dx, dy, kf, conf_cpwt.k0)
        sg, esg = _maps(nx, ny, is1 - is0, 2, _OUT_DTYPE)
        fn = _lib.lib.pfx_wlet_scalogram_host_f32 if _OUT_DTYPE == np.float32 else _lib.lib.pfx_wlet_scalogram_host
        _lib.check(fn(plan.handle, _lib.ptr(g), is0, is1, _lib.ptr(sg), _lib.ptr(esg)))
        return sg, esg

    def wlet_xscalogram(self, grid1, grid2, dx, dy, kf, scales=None):
        g1, g2 = _grid(grid1, "grid1"), _grid(grid2, "grid2")
        if g1.shape != g2.shape:
            raise ValueError("grids must have the same shape")
        kf = np.ascontiguousarray(kf, dtype=np.float64).ravel()
        nx, ny = g1.shape
        is0, is1 = scales if scales is not None else (0, len(kf))
        plan = _lib.get_plan(nx, ny, dx, dy, kf, conf_cpwt.k0)
        xsg = np.empty((nx, ny, is1 - is0), dtype=np.complex128, order="F")
        (exsg,) = _maps(nx, ny, is1 - is0, 1)
        _lib.check(_lib.lib.pfx_wlet_xscalogram_host(plan.handle, _lib.ptr(g1), _lib.ptr(g2), is0, is1, _lib.ptr(xsg),
                                                     _lib.ptr(exsg)))
        return xsg, exsg

    def wlet_admit_coh(self, topo, grav, dx, dy, kf, scales=None, out=None):
        g1, g2 = _grid(topo, "topo"), _grid(grav, "grav")
        if g1.shape != g2.shape:
            raise ValueError("grids must have the same shape")
        kf = np.ascontiguousarray(kf, dtype=np.float64).ravel()
        nx, ny = g1.shape
        is0, is1 = scales if scales is not None else (0, len(kf))
        plan = _lib.get_plan(nx, ny, dx, dy, kf, conf_cpwt.k0)
        outs = out if out is not None else _maps(nx, ny, is1 - is0, 4, _OUT_DTYPE)
        f32 = outs[0].dtype == np.float32
        if any(o.dtype != outs[0].dtype or o.shape != (nx, ny, is1 - is0) or not o.flags.f_contiguous for o in outs):
            raise ValueError("output maps must be Fortran-ordered (nx, ny, ns) arrays of one dtype")
        fn = _lib.lib.pfx_wlet_admit_coh_host_f32 if f32 else _lib.lib.pfx_wlet_admit_coh_host
        _lib.check(fn(plan.handle, _lib.ptr(g1), _lib.ptr(g2), is0, is1, *[_lib.ptr(o) for o in outs]))
        return tuple(outs)


    def wlet_admit_coh_device(self, topo, grav, dx, dy, kf):
        """Same transform, results left on the GPU: returns a DeviceMaps holding wl_admit, ewl_admit, wl_coh,
        ewl_coh as (nx, ny, ns) float64 device arrays (x fastest).  Project.wlet_admit_coh uses it so that
        Project.estimate_grid can fit the maps where they are; a map crosses to the host only when read."""
        g1, g2 = _grid(topo, "topo"), _grid(grav, "grav")
        if g1.shape != g2.shape:
            raise ValueError("grids must have the same shape")
        kf = np.ascontiguousarray(kf, dtype=np.float64).ravel()
        nx, ny = g1.shape
        plan = _lib.get_plan(nx, ny, dx, dy, kf, conf_cpwt.k0)
        maps = DeviceMaps(nx, ny, len(kf), plan.device)
        d_in = [_lib.DeviceBuffer.from_host(g, plan.device) for g in (g1, g2)]
        try:
            _lib.check(_lib.lib.pfx_wlet_admit_coh_dev(plan.handle, C.c_void_p(d_in[0].ptr), C.c_void_p(d_in[1].ptr), 0,
                                                       len(kf), *[C.c_void_p(p) for p in maps.ptrs()]))
            maps.synchronize()
        finally:
            for b in d_in:
                b.close()
        return maps


class DeviceMaps:
    """Four (nx, ny, ns) float64 maps resident on one GPU, in the order wl_admit, ewl_admit, wl_coh, ewl_coh."""

    NAMES = ("wl_admit", "ewl_admit", "wl_coh", "ewl_coh")

    def __init__(self, nx, ny, ns, device):
        self.nx, self.ny, self.ns, self.device = int(nx), int(ny), int(ns), int(device)
        self.map_bytes = self.nx * self.ny * self.ns * 8
        self.buf = _lib.DeviceBuffer(4 * self.map_bytes, device)
        self.out_dtype = _OUT_DTYPE  # the output dtype in force when the maps were computed (set_output_dtype)

    def ptrs(self):
        return [self.buf.ptr + i * self.map_bytes for i in range(4)]

    def synchronize(self):
        # a zero-byte-free way to wait for the plan's (legacy default) stream: a tiny synchronous copy
        probe = np.empty(1)
        self.buf.download(probe)

    def to_host(self, index):
        out = np.empty((self.nx, self.ny, self.ns), dtype=np.float64, order="F")
        self.buf.download(out, index * self.map_bytes)
        return out.astype(np.float32) if self.out_dtype == np.float32 else out

    def cell(self, i, j):
        """The ns values of every map at grid cell (i, j): four length-ns vectors, without moving the maps."""
        i, j = int(i) % self.nx, int(j) % self.ny
        out = np.empty((4, self.ns))
        one = np.empty(1)
        for m in range(4):
            for s in range(self.ns):
                out[m, s] = self.buf.download(one, m * self.map_bytes + ((s * self.ny + j) * self.nx + i) * 8)[0]
        return out.astype(np.float32).astype(np.float64) if self.out_dtype == np.float32 else out

    def close(self):
        self.buf.close()


fused = _Fused()
