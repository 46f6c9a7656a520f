"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/plateflex_b200.h declares; the ctypes struct matches the C layout."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "plateflex_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(pf):
    from plateflex_b200 import _lib
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/plateflex_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)


def test_basic_calls_without_gpu(pf):
    from plateflex_b200 import _lib
    assert _lib.lib.pfx_version() >= 100
    assert [_lib.lib.pfx_npow2(n) for n in (1, 2, 3, 684, 2048, 2049)] == [2, 2, 4, 1024, 2048, 4096]
    assert C.sizeof(_lib.FlexConf) == 48  # 5 doubles + 2 int32, as in pfx_flex_conf
    assert isinstance(_lib.lib.pfx_last_error(), bytes)


def test_bad_arguments_are_rejected_before_any_device_work(pf):
    import numpy as np
    from plateflex_b200 import _lib
    h = C.c_void_p()
    k = np.array([1e-5])
    assert _lib.lib.pfx_plan_create(C.byref(h), 1, 8, 1.0, 1.0, _lib.ptr(k), 1, 5.336, 0, 0) == 1  # PFX_ERR_ARG
    assert b"nx, ny" in _lib.lib.pfx_last_error()
    assert _lib.lib.pfx_plan_create(C.byref(h), 8, 8, 1.0, 1.0, None, 1, 5.336, 0, 0) == 1
    assert _lib.lib.pfx_plan_create(C.byref(h), 8, 8, -1.0, 1.0, _lib.ptr(k), 1, 5.336, 0, 0) == 1
    assert _lib.lib.pfx_plan_destroy(None) == 0
