"""Grid ingest: the step in front of the hot path.

``read_xyz`` loads a GMT ".xyz" dump the way the reference's notebooks do with pandas (Ex1_making_grids.ipynb
cells 2-3; the format is produced by the recipe of docs/getting_started.rst:50-56): header line
``# name<TAB>xmin xmax ymin ymax zmin zmax dx dy nx ny`` then one ``x y z`` line per node, parsed by the native
multi-threaded reader of libplateflex_b200 and returned as the (ny, nx) array flipped along its first axis --
exactly ``values.reshape(ny, nx)[::-1]``.  ``fill_nan_nearest`` and ``gaussian_filter`` are the GPU forms of
the two O(nx*ny) host steps of Grid.__init__ / TopoGrid.filter_water_depth (classes.py:143-153, 538-541).
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["read_xyz", "read_xyz_header", "fill_nan_nearest", "gaussian_filter"]

_FIELDS = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "dx", "dy", "nx", "ny")


def read_xyz_header(path):
    """The ten grdinfo numbers of the header line as a dict (nx, ny as int)."""
    h = np.empty(10)
    _lib.check(_lib.lib.pfx_xyz_header(str(path).encode(), _lib.ptr(h)))
    out = dict(zip(_FIELDS, h.tolist()))
    out["nx"], out["ny"] = int(out["nx"]), int(out["ny"])
    return out


def read_xyz(path, nthreads=0):
    """Returns (data, dx, dy, header): ``data`` is what the notebooks hand to Grid -- the z column reshaped to
    (ny, nx) and flipped along the first axis -- and dx, dy the header's sampling intervals (km)."""
    h = read_xyz_header(path)
    z = np.empty(h["nx"] * h["ny"])
    _lib.check(_lib.lib.pfx_xyz_read(str(path).encode(), z.size, _lib.ptr(z), int(nthreads)))
    return np.ascontiguousarray(z.reshape(h["ny"], h["nx"])[::-1]), h["dx"], h["dy"], h


def fill_nan_nearest(grid, device=None):
    """classes.py:143-153 on the GPU: a copy of ``grid`` (2-D) with every non-finite cell replaced by the value
    of the nearest finite cell in index space."""
    a = np.array(grid, dtype=np.float64, order="C", copy=True)
    if a.ndim != 2:
        raise ValueError("grid must be 2-D")
    n = C.c_long()
    _lib.check(_lib.lib.pfx_nan_fill_nearest_host(_lib.default_device() if device is None else int(device), a.shape[0],
                                                  a.shape[1], _lib.ptr(a), C.byref(n)))
    return a


def gaussian_filter(image, sigma, device=None):
    """skimage.filters.gaussian(image, sigma) for a float image (edge replication, truncate 4) on the GPU."""
    a = np.ascontiguousarray(image, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("image must be 2-D")
    out = np.empty_like(a)
    _lib.check(_lib.lib.pfx_gaussian_filter_host(_lib.default_device() if device is None else int(device), a.shape[0],
                                                 a.shape[1], float(sigma), _lib.ptr(a), _lib.ptr(out)))
    return out
