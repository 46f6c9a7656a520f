"""ctypes binding of libplateflex_b200.so (C ABI in include/plateflex_b200.h).

The shared library is the product: there is no CPU or pure-PyTorch fallback.
If it is missing or cannot be loaded this module raises ImportError loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLATEFLEX_B200_LIB: another build of the library (kernel experiments: scripts/exp_build.sh)
LIB_PATH = os.environ.get("PLATEFLEX_B200_LIB") or os.path.join(_HERE, "libplateflex_b200.so")

NA = 11
ENGINE_PRUNED, ENGINE_DENSE = 0, 1
ATYPES = {"admit": 0, "coh": 1, "joint": 2}

_c_double_p = C.POINTER(C.c_double)


class FlexConf(C.Structure):
    """pfx_flex_conf -- conf_flex module variables (flex.f90:47-48)."""

    _fields_ = [("rhoc", C.c_double), ("rhom", C.c_double), ("rhow", C.c_double), ("rhoa", C.c_double),
                ("zc", C.c_double), ("boug", C.c_int32), ("pad_", C.c_int32)]


class PlateflexError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C plateflex_b200/csrc`. plateflex_b200 has no CPU fallback.")
    try:
        return C.CDLL(LIB_PATH)
    except OSError as exc:  # missing libcufft / libcudart etc.
        raise ImportError(f"cannot load {LIB_PATH}: {exc}") from exc


lib = _load()

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_SIGS = {
    "pfx_version": ([], _i),
    "pfx_last_error": ([], C.c_char_p),
    "pfx_device_count": ([C.POINTER(_i)], _i),
    "pfx_npow2": ([_i], _i),
    "pfx_fp64_peak": ([_i, C.POINTER(_d)], _i),
    "pfx_dev_alloc": ([_i, C.c_size_t, C.POINTER(_vp)], _i),
    "pfx_dev_free": ([_i, _vp], _i),
    "pfx_dev_h2d": ([_i, _vp, _vp, C.c_size_t], _i),
    "pfx_dev_h2d_transposed": ([_i, _vp, _vp, _i, _i, _i], _i),
    "pfx_dev_d2h": ([_i, _vp, _vp, C.c_size_t], _i),
    "pfx_xyz_header": ([C.c_char_p, _vp], _i),
    "pfx_xyz_read": ([C.c_char_p, C.c_long, _vp, _i], _i),
    "pfx_nan_fill_nearest_host": ([_i, _i, _i, _vp, C.POINTER(C.c_long)], _i),
    "pfx_gaussian_filter_host": ([_i, _i, _i, _d, _vp, _vp], _i),
    "pfx_plan_create": ([C.POINTER(_vp), _i, _i, _d, _d, _vp, _i, _d, _i, _i], _i),
    "pfx_plan_destroy": ([_vp], _i),
    "pfx_plan_set_stream": ([_vp, _vp], _i),
    "pfx_plan_info": ([_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_size_t)], _i),
    "pfx_plan_launch_count": ([_vp, C.POINTER(C.c_longlong)], _i),
    "pfx_plan_set_cutoff": ([_vp, _d], _i),
    "pfx_plan_scale_cost": ([_vp, _vp], _i),
    "pfx_plan_profile_enable": ([_vp, _i], _i),
    "pfx_plan_profile_read": ([_vp, _vp, _vp, _i], _i),
    "pfx_wlet_transform_host": ([_vp, _vp, _i, _i, _vp], _i),
    "pfx_wlet_transform_dev": ([_vp, _vp, _i, _i, _vp], _i),
    "pfx_wlet_scalogram_host": ([_vp, _vp, _i, _i, _vp, _vp], _i),
    "pfx_wlet_scalogram_dev": ([_vp, _vp, _i, _i, _vp, _vp], _i),
    "pfx_wlet_xscalogram_host": ([_vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    "pfx_wlet_xscalogram_dev": ([_vp, _vp, _vp, _i, _i, _vp, _vp], _i),
    "pfx_wlet_admit_coh_host": ([_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "pfx_wlet_admit_coh_dev": ([_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "pfx_wlet_transform_host_c64": ([_vp, _vp, _i, _i, _vp], _i),
    "pfx_wlet_scalogram_host_f32": ([_vp, _vp, _i, _i, _vp, _vp], _i),
    "pfx_wlet_admit_coh_host_f32": ([_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "pfx_wlet_admit_coh_sharded_dev": ([_vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "pfx_wlet_admit_coh_sharded_list_dev": ([_vp, _vp, _vp, _vp, _i, _i, _vp], _i),
    "pfx_shard_bytes": ([_i, _i, _i, _i, _i], C.c_size_t),
    "pfx_ipc_alloc": ([_i, C.c_size_t, C.POINTER(_vp), _vp], _i),
    "pfx_ipc_free": ([_i, _vp], _i),
    "pfx_ipc_open": ([_i, _vp, C.POINTER(_vp)], _i),
    "pfx_ipc_close": ([_i, _vp], _i),
    "pfx_cube_scalogram_host": ([_i, _i, _i, _i, _vp, _vp, _vp], _i),
    "pfx_cube_xscalogram_host": ([_i, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "pfx_cube_admit_coh_host": ([_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "pfx_flex_xspec_host": ([_i, C.POINTER(FlexConf), _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "pfx_flex_helper_host": ([_i, C.POINTER(FlexConf), _d, _d, _i, _i, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp], _i),
    "pfx_defk_host": ([_i, _i, _i, _d, _d, _vp, _vp], _i),
    "pfx_wave_function_host": ([_i, _i, _i, _vp, _vp, _d, _d, _d, _vp], _i),
    "pfx_l2_fit_host": ([_i, C.POINTER(FlexConf), _vp, _i, _i, _i] + [_vp] * 8 + [_i, _i, _i] + [_vp] * 8, _i),
    "pfx_l2_fit_rows_shape": ([_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)], _i),
    "pfx_l2_fit_rows_dev": ([_i, _vp, C.POINTER(FlexConf), _vp, _i, _i, _i, _i, _i] + [_vp] * 8 + [_i, _i, _i] +
                            [_vp] * 8, _i),
    "pfx_l2_fit_dev": ([_i, _vp, C.POINTER(FlexConf), _vp, _i, _i, _i] + [_vp] * 8 + [_i, _i, _i] + [_vp] * 8, _i),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here means header and library disagree
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTS = tuple(_SIGS)


def check(rc):
    if rc != 0:
        msg = lib.pfx_last_error().decode(errors="replace")
        kinds = {1: ValueError, 4: MemoryError}
        raise kinds.get(rc, PlateflexError)(f"libplateflex_b200 error {rc}: {msg}")


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int(0)
    check(lib.pfx_device_count(C.byref(n)))
    return n.value


def fp64_peak(device=None):
    """Measured FP64 peak (TFLOP/s) of a device: independent DFMA chains, all SMs full."""
    t = C.c_double()
    check(lib.pfx_fp64_peak(default_device() if device is None else int(device), C.byref(t)))
    return t.value


def default_device():
    """CUDA ordinal used by the host-pointer API: PLATEFLEX_B200_DEVICE, else LOCAL_RANK, else 0."""
    for var in ("PLATEFLEX_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


class DeviceBuffer:
    """A block of device memory owned by the library (pfx_dev_alloc): lets the Python layer keep
    intermediate maps on the GPU between calls without depending on a tensor package."""

    def __init__(self, nbytes, device=None):
        self.device = default_device() if device is None else int(device)
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib.pfx_dev_alloc(self.device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, arr, device=None):
        buf = cls(arr.nbytes, device)
        buf.upload(arr)
        return buf

    def upload(self, arr):
        assert arr.nbytes <= self.nbytes and (arr.flags.c_contiguous or arr.flags.f_contiguous)
        check(lib.pfx_dev_h2d(self.device, C.c_void_p(self.ptr), ptr(arr), arr.nbytes))

    @classmethod
    def from_host_as_fortran(cls, arr, device=None):
        """Device copy of a 2-D array in Fortran order (first index fastest) whatever its order on the host: a
        C-ordered array is transposed on the GPU instead of through a strided numpy copy."""
        if arr.ndim != 2 or arr.flags.f_contiguous or not arr.flags.c_contiguous or arr.itemsize not in (1, 8):
            return cls.from_host(np.asfortranarray(arr), device)
        buf = cls(arr.nbytes, device)
        check(lib.pfx_dev_h2d_transposed(buf.device, C.c_void_p(buf.ptr), ptr(arr), arr.shape[0], arr.shape[1], arr.itemsize))
        return buf

    def download(self, out, offset=0):
        assert offset + out.nbytes <= self.nbytes and (out.flags.c_contiguous or out.flags.f_contiguous)
        check(lib.pfx_dev_d2h(self.device, ptr(out), C.c_void_p(self.ptr + offset), out.nbytes))
        return out

    def close(self):
        if getattr(self, "ptr", None):
            lib.pfx_dev_free(self.device, C.c_void_p(self.ptr))
            self.ptr = None

    __del__ = close


class Plan:
    """Owns one pfx_plan (sizes, wavenumbers, k0, engine, device workspaces)."""

    def __init__(self, nx, ny, dx, dy, k, k0, engine=ENGINE_PRUNED, device=None):
        self.nx, self.ny, self.dx, self.dy = int(nx), int(ny), float(dx), float(dy)
        self.k = np.ascontiguousarray(k, dtype=np.float64)
        self.ns, self.k0, self.engine = len(self.k), float(k0), int(engine)
        self.device = default_device() if device is None else int(device)
        h = C.c_void_p()
        check(lib.pfx_plan_create(C.byref(h), self.nx, self.ny, self.dx, self.dy, ptr(self.k), self.ns, self.k0,
                                  self.engine, self.device))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            lib.pfx_plan_destroy(self.handle)
            self.handle = None

    __del__ = close

    def set_stream(self, cuda_stream):
        check(lib.pfx_plan_set_stream(self.handle, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def set_cutoff(self, nsigma):
        check(lib.pfx_plan_set_cutoff(self.handle, float(nsigma)))

    def info(self):
        a, b, s = C.c_int(), C.c_int(), C.c_size_t()
        check(lib.pfx_plan_info(self.handle, C.byref(a), C.byref(b), C.byref(s)))
        return dict(nnx=a.value, nny=b.value, device_bytes=s.value)

    def launch_count(self):
        n = C.c_longlong()
        check(lib.pfx_plan_launch_count(self.handle, C.byref(n)))
        return n.value

    KERNEL_CLASSES = ("forward", "prepare", "pass1", "pass2", "reduce", "dense_mul", "dense_fft", "crop")

    def scale_cost(self):
        cost = np.empty(self.ns)
        check(lib.pfx_plan_scale_cost(self.handle, ptr(cost)))
        return cost

    def profile(self, on=True):
        check(lib.pfx_plan_profile_enable(self.handle, int(bool(on))))

    def profile_read(self):
        """{kernel class: (milliseconds, launches)} since the last read (device timers, see the header)."""
        n = len(self.KERNEL_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_longlong * n)()
        check(lib.pfx_plan_profile_read(self.handle, ms, cnt, n))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.KERNEL_CLASSES) if cnt[i]}

    def key(self):
        return (self.nx, self.ny, self.dx, self.dy, self.k.tobytes(), self.k0, self.engine, self.device)


_PLANS = {}
_PLAN_ORDER = []
_MAX_PLANS = 2


def get_plan(nx, ny, dx, dy, k, k0, engine=None, device=None):
    """Small LRU of plans so that the topography and gravity grids of a project share one."""
    engine = default_engine() if engine is None else engine
    device = default_device() if device is None else device
    k = np.ascontiguousarray(k, dtype=np.float64)
    key = (int(nx), int(ny), float(dx), float(dy), k.tobytes(), float(k0), int(engine), int(device))
    if key in _PLANS:
        _PLAN_ORDER.remove(key)
        _PLAN_ORDER.append(key)
        return _PLANS[key]
    while len(_PLAN_ORDER) >= _MAX_PLANS:
        _PLANS.pop(_PLAN_ORDER.pop(0)).close()
    _PLANS[key] = Plan(nx, ny, dx, dy, k, k0, engine, device)
    _PLAN_ORDER.append(key)
    return _PLANS[key]


def clear_plans():
    while _PLAN_ORDER:
        _PLANS.pop(_PLAN_ORDER.pop(0)).close()


def default_engine():
    name = os.environ.get("PLATEFLEX_B200_ENGINE", "pruned").lower()
    if name not in ("pruned", "dense"):
        raise ValueError("PLATEFLEX_B200_ENGINE must be 'pruned' or 'dense'")
    return ENGINE_DENSE if name == "dense" else ENGINE_PRUNED
