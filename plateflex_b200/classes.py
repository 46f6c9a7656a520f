"""Grid / Project object API -- the host-side mirror of the reference's
plateflex/classes.py for the hot path: same class names, method names,
attribute names, argument meaning and exception messages, with the numerical
work routed to the GPU library instead of the f2py modules.

Mirrored (reference file:line):
  Grid                         classes.py:71-286   (wlet_transform, wlet_scalogram)
  GravGrid/BougGrid/FairGrid   classes.py:399-477
  TopoGrid                     classes.py:480-546
  RhocGrid / ZcGrid            classes.py:571-632
  Project                      classes.py:635-1332 (init, wlet_admit_coh, estimate_cell, estimate_grid)
  _lam2k                       classes.py:1757-1828
  save_results / filter_water_depth   classes.py:538-546, 1621-1754 (host-side post-processing)
Out of scope here (presentation / Bayesian path): the plot_* methods and
``inverse='bayes'``; they raise NotImplementedError with a pointer to the
reference package.
"""
import numpy as np

from . import estimate
from .cpwt import conf_cpwt as cf_w
from .cpwt import cpwt, fused
from .flex import conf_flex as cf_f

__all__ = ["Grid", "TopoGrid", "GravGrid", "BougGrid", "FairGrid", "RhocGrid", "ZcGrid", "Project"]


def _gaussian(image, sigma=1):
    """skimage.filters.gaussian as the reference calls it (classes.py:539-541, 1464, 1663) for float
    images: scipy.ndimage.gaussian_filter with mode='nearest' and truncate=4.0 -- which is what
    scikit-image itself evaluates; used directly when scikit-image is not installed."""
    try:
        from skimage.filters import gaussian
        return gaussian(image, sigma=sigma)
    except ImportError:
        from scipy import ndimage
        return ndimage.gaussian_filter(np.asarray(image, dtype=np.float64), sigma, mode='nearest', truncate=4.0)


def _out_of_scope(name):
    raise NotImplementedError(
        f"{name} is presentation / Bayesian code outside the GPU hot path; use the reference plateflex package "
        "on the numpy attributes of this object")


class Grid(object):
    """2-D data grid with spacing (dx, dy) in km (classes.py:71-157).

    Attributes as in the reference: data, dx, dy, nx, ny, units, sg_units, logsg_units, title,
    ns, k; after the corresponding calls also wl_trans, wl_sg, ewl_sg.
    """

    def __init__(self, grid, dx, dy):
        nx, ny = grid.shape
        self.nx, self.ny = nx, ny
        self.dx, self.dy = dx, dy
        self.units = None
        self.sg_units = None
        self.logsg_units = None
        self.title = None
        self.ns, self.k = _lam2k(nx, ny, dx, dy)  # classes.py:141

        if np.any(np.isnan(np.array(grid))):
            # nearest-neighbour fill of NaNs, classes.py:143-153: scipy.interpolate.griddata(method='nearest')
            # in the reference, an exact nearest-finite-cell search on the GPU here (io.fill_nan_nearest).  Where
            # two finite cells are exactly equidistant SciPy's pick depends on its k-d tree; ours is the
            # smallest (row, column).
            print('grid contains NaN values. Performing interpolation...')
            from . import io
            self.data = io.fill_nan_nearest(np.array(grid, dtype=np.float64))
        else:
            self.data = np.array(grid)

    def wlet_transform(self):
        """classes.py:195-230: stores the (nx, ny, 11, ns) complex cube as ``wl_trans``."""
        self.wl_trans = cpwt.wlet_transform(self.data, self.dx, self.dy, self.k)
        return

    def wlet_scalogram(self):
        """classes.py:232-286: stores ``wl_sg`` and ``ewl_sg`` (nx, ny, ns).

        With a ``wl_trans`` attribute present the cube is reduced as in the reference.  Without
        one the reference first builds and keeps the cube (classes.py:276-280); here the fused
        GPU path produces the same two maps without materialising it, and ``wl_trans`` stays
        unset."""
        if hasattr(self, 'wl_trans'):
            wl_sg, ewl_sg = cpwt.wlet_scalogram(self.wl_trans)
        else:
            wl_sg, ewl_sg = fused.wlet_scalogram(self.data, self.dx, self.dy, self.k)
        self.wl_sg = wl_sg
        self.ewl_sg = ewl_sg
        return

    def make_contours(self, level=0.):
        """classes.py:159-178 (host-side helper, unchanged behaviour)."""
        try:
            from skimage import measure
        except Exception:
            raise (Exception("Package 'scikit-image' not available to make contours."))
        return measure.find_contours(self.data, level)

    def plot(self, *args, **kwargs):
        _out_of_scope('Grid.plot')

    def plot_transform(self, *args, **kwargs):
        _out_of_scope('Grid.plot_transform')

    def plot_scalogram(self, *args, **kwargs):
        _out_of_scope('Grid.plot_scalogram')


class GravGrid(Grid):
    """classes.py:399-431."""

    def __init__(self, grid, dx, dy):
        Grid.__init__(self, grid, dx, dy)
        self.units = 'mGal'
        self.sg_units = r'mGal$^2$/|k|'
        self.logsg_units = r'log(mGal$^2$/|k|)'
        self.title = 'Gravity anomaly'


class BougGrid(GravGrid):
    """classes.py:434-454."""

    def __init__(self, grid, dx, dy):
        GravGrid.__init__(self, grid, dx, dy)
        self.title = 'Bouguer anomaly'


class FairGrid(GravGrid):
    """classes.py:457-477."""

    def __init__(self, grid, dx, dy):
        GravGrid.__init__(self, grid, dx, dy)
        self.title = 'Free-air anomaly'


class TopoGrid(Grid):
    """classes.py:480-546.  Converts to metres when the standard deviation is below 20
    (classes.py:530-531) and derives ``water_depth`` from the input."""

    def __init__(self, grid, dx, dy):
        Grid.__init__(self, grid, dx, dy)
        self.units = 'm'
        self.sg_units = r'm$^2$/|k|'
        self.logsg_units = r'log(m$^2$/|k|)'
        self.title = 'Topography'
        if np.std(self.data) < 20.:
            self.data *= 1.e3
        # classes.py:533-536: the reference builds the water depth IN the caller's array (positive
        # topography zeroed in place, original units).  Kept as is: results must be identical.
        water_depth = grid
        water_depth[grid > 0.] = 0.
        self.water_depth = -1. * water_depth

    def filter_water_depth(self, sigma=10, returned=False):
        """classes.py:538-546; the Gaussian (skimage.filters.gaussian in the reference) runs on the GPU."""
        from . import io
        water_depth = io.gaussian_filter(self.data, sigma)
        water_depth[self.data > 0.] = 0.
        self.water_depth = -1. * water_depth
        if returned:
            return water_depth

    def plot_water_depth(self, *args, **kwargs):
        _out_of_scope('TopoGrid.plot_water_depth')


class RhocGrid(Grid):
    """classes.py:571-600: crustal density per cell (kg/m^3; scaled by 1e3 if std < 10)."""

    def __init__(self, grid, dx, dy):
        Grid.__init__(self, grid, dx, dy)
        self.units = r'kg/m$^3$'
        self.title = 'Crustal density'
        if np.std(self.data) < 10.:
            self.data *= 1.e3


class ZcGrid(Grid):
    """classes.py:603-632: crustal thickness per cell (m; scaled by 1e3 if std < 20)."""

    def __init__(self, grid, dx, dy):
        Grid.__init__(self, grid, dx, dy)
        self.units = 'm'
        self.title = 'Crustal thickness'
        if np.std(self.data) < 20.:
            self.data *= 1.e3


def _lazy_map(index, name):
    """One of the four admittance / coherence maps of a Project as the reference's plain attribute
    (classes.py:968-971): reading it before ``wlet_admit_coh`` raises AttributeError -- the reference's own
    probing idiom is ``try: self.wl_admit ... except AttributeError`` (classes.py:1001-1008) -- and the first
    read after it copies the map from the GPU, where ``estimate_grid`` uses it in place.  Assigning a numpy
    array replaces the map; the device copy is then dropped and the host arrays are used."""
    key = "_host_" + name

    def fget(self):
        h = self.__dict__.get(key)
        if h is None:
            d = self.__dict__.get("_dmaps")
            if d is None:
                raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")
            h = self.__dict__[key] = d.to_host(index)
        return h

    def fset(self, value):
        d = self.__dict__.get("_dmaps")
        if d is not None:  # keep what the device holds reachable, then let go of it
            for i, n in enumerate(d.NAMES):
                if n != name and self.__dict__.get("_host_" + n) is None:
                    self.__dict__["_host_" + n] = d.to_host(i)
            d.close()
            self.__dict__["_dmaps"] = None
        self.__dict__[key] = value

    def fdel(self):
        if self.__dict__.get(key) is None and self.__dict__.get("_dmaps") is None:
            raise AttributeError(name)
        self.__dict__[key] = None

    return property(fget, fset, fdel)


class Project(object):
    """Container of Grid objects (classes.py:635-1332): ``init`` validates, ``wlet_admit_coh``
    computes wavelet admittance / coherence with jackknife errors, ``estimate_cell`` /
    ``estimate_grid`` fit the flexural model.  Attributes: grids, inverse, mask, initialized and,
    after the calls, k, ns, nx, ny, dx, dy, water_depth, rhoc, zc, wl_admit, ewl_admit, wl_coh,
    ewl_coh, mean_Te_grid, std_Te_grid, mean_F_grid, std_F_grid, chi2_grid[, mean_a_grid,
    std_a_grid], new_mask_grid, plus ``status_grid`` (per-cell fit status, new)."""

    wl_admit = _lazy_map(0, "wl_admit")
    ewl_admit = _lazy_map(1, "ewl_admit")
    wl_coh = _lazy_map(2, "wl_coh")
    ewl_coh = _lazy_map(3, "ewl_coh")

    def __init__(self, grids=None):
        self.inverse = 'L2'
        self.grids = []
        self.mask = None
        self.initialized = False
        if isinstance(grids, Grid):
            grids = [grids]
        if grids:
            self.grids.extend(grids)

    def __add__(self, other):
        if isinstance(other, Grid):
            other = Project([other])
        if not isinstance(other, Project):
            raise TypeError
        return self.__class__(grids=self.grids + other.grids)

    def __iter__(self):
        return list(self.grids).__iter__()

    def append(self, grid):
        if isinstance(grid, Grid):
            self.grids.append(grid)
        else:
            raise TypeError('Append only supports a single Grid object as an argument.')
        return self

    def extend(self, grid_list):
        if isinstance(grid_list, list):
            for _i in grid_list:
                if not isinstance(_i, Grid):
                    raise TypeError('Extend only accepts a list of Grid objects.')
            self.grids.extend(grid_list)
        elif isinstance(grid_list, Project):
            self.grids.extend(grid_list.grids)
        else:
            raise TypeError('Extend only supports a list of Grid objects as argument.')
        return self

    def init(self):
        """classes.py:818-914."""
        if not any(isinstance(g, TopoGrid) for g in self.grids):
            raise (Exception('There needs to be one TopoGrid object in Project'))
        if not any(isinstance(g, GravGrid) for g in self.grids):
            raise (Exception('There needs to be one GravGrid object in Project'))
        shape = [grid.data.shape for grid in self.grids]
        if not (len(set(shape)) == 1):
            raise (Exception('Grids do not have the same shape - aborting:' + str(shape)))
        dd = [(grid.dx, grid.dy) for grid in self.grids]
        if not (len(set(dd)) == 1):
            raise (Exception('Grids do not have the same sampling intervals - ' + 'aborting:' + str(dd)))

        # Bouguer or free air selects A = 0 / 1 in the model (classes.py:888-891, flex.f90:200-204)
        if any(isinstance(g, BougGrid) for g in self.grids):
            cf_f.boug = 1
        elif any(isinstance(g, FairGrid) for g in self.grids):
            cf_f.boug = 0

        self.rhoc = None
        self.zc = None
        for grid in self.grids:
            if isinstance(grid, RhocGrid):
                self.rhoc = grid.data
            elif isinstance(grid, ZcGrid):
                self.zc = grid.data
            elif isinstance(grid, TopoGrid):
                self.water_depth = grid.water_depth

        first = self.grids[0]
        self.k, self.ns = first.k, first.ns
        self.nx, self.ny, self.dx, self.dy = first.nx, first.ny, first.dx, first.dy
        self.initialized = True

    def _topo_grav(self):
        topo = grav = None
        for grid in self.grids:  # last one of each kind wins, as in classes.py:950-962
            if isinstance(grid, TopoGrid):
                topo = grid
            elif isinstance(grid, GravGrid):
                grav = grid
        return topo, grav

    def wlet_admit_coh(self):
        """classes.py:916-975.  When neither grid carries a ``wl_trans`` cube the fused GPU path
        is used (two grids in, four (nx, ny, ns) maps out, no cube); when cubes exist they are
        reduced exactly as cpwt.wlet_admit_coh(wl_trans_topo, wl_trans_grav) would."""
        if not self.initialized:
            raise (Exception("Project not yet initialized - Abort"))
        topo, grav = self._topo_grav()
        have = [hasattr(g, 'wl_trans') for g in (topo, grav)]
        if not any(have):
            # maps stay on the GPU; the four attributes are filled from there when first read
            old = self.__dict__.get("_dmaps")
            if old is not None:
                old.close()
            self.__dict__["_dmaps"] = None
            dmaps = fused.wlet_admit_coh_device(topo.data, grav.data, self.dx, self.dy, self.k)
            for n in dmaps.NAMES:
                self.__dict__["_host_" + n] = None
            self.__dict__["_dmaps"] = dmaps
            return
        for g, h in zip((topo, grav), have):
            if not h:
                g.wlet_transform()
        res = cpwt.wlet_admit_coh(topo.wl_trans, grav.wl_trans)
        self.wl_admit, self.ewl_admit, self.wl_coh, self.ewl_coh = res
        return

    def _check_estimate_args(self, alph, atype):
        if not isinstance(alph, bool):
            raise (Exception("'alph' should be a boolean: defaults to False"))
        if atype not in ['admit', 'coh', 'joint']:
            raise (Exception("'atype' should be one among: 'admit', 'coh', or 'joint'"))

    def estimate_cell(self, cell=(0, 0), alph=False, atype='joint', returned=False):
        """classes.py:1031-1123 for ``inverse='L2'``."""
        if not self.initialized:
            raise (Exception('Project not initialized. Aborting'))
        self._check_estimate_args(alph, atype)
        dmaps = self.__dict__.get("_dmaps")
        if dmaps is not None and all(self.__dict__.get("_host_" + n) is None for n in dmaps.NAMES):
            adm, eadm, coh, ecoh = dmaps.cell(cell[0], cell[1])  # one cell from the GPU, the maps stay there
        else:
            adm = self.wl_admit[cell[0], cell[1], :]
            eadm = self.ewl_admit[cell[0], cell[1], :]
            coh = self.wl_coh[cell[0], cell[1], :]
            ecoh = self.ewl_coh[cell[0], cell[1], :]
        # per-cell model parameters go through the module globals, classes.py:1087-1091
        cf_f.wd = self.water_depth[cell[0], cell[1]]
        if self.rhoc is not None:
            cf_f.rhoc = self.rhoc[cell[0], cell[1]]
        if self.zc is not None:
            cf_f.zc = self.zc[cell[0], cell[1]]
        if self.inverse == 'L2':
            summary = estimate.L2_estimate_cell(self.k, adm, eadm, coh, ecoh, alph, atype)
            if returned:
                return summary
            self.alph, self.atype, self.cell, self.summary = alph, atype, cell, summary
        elif self.inverse == 'bayes':
            _out_of_scope("Project.estimate_cell(inverse='bayes')")

    def estimate_grid(self, nn=10, alph=False, atype='joint', parallel=False):
        """classes.py:1125-1332 for ``inverse='L2'``: one kernel launch instead of the Python
        double loop.  Same decimation (cells range(0, nx-nn, nn) x range(0, ny-nn, nn) written
        into zero-initialised (int(nx/nn), int(ny/nn)) maps), same mask handling, same per-cell
        wd / rhoc / zc.  Side effect kept from the reference: the module globals end up holding
        the values of the last fitted cell."""
        self.nn = nn
        self._check_estimate_args(alph, atype)
        self.alph = alph
        self.atype = atype
        if self.inverse == 'bayes':
            _out_of_scope("Project.estimate_grid(inverse='bayes')")
        elif self.inverse != 'L2':
            raise (Exception('Inverse method invalid: ' + str(self.inverse)))
        if parallel:
            raise (Exception('Parallel implementation does not work - ' + 'check again later'))

        dmaps = self.__dict__.get("_dmaps")
        if dmaps is not None:  # the maps of wlet_admit_coh are still on the GPU: fit them there
            res = estimate.L2_estimate_grid_device(self.k, dmaps, self.water_depth, rhoc=self.rhoc, zc=self.zc,
                                                   mask=self.mask, nn=nn, alph=alph, atype=atype)
        else:
            res = estimate.L2_estimate_grid(self.k, self.wl_admit, self.ewl_admit, self.wl_coh, self.ewl_coh,
                                            self.water_depth, rhoc=self.rhoc, zc=self.zc, mask=self.mask, nn=nn,
                                            alph=alph, atype=atype)
        if self.mask is not None:
            # classes.py:1185-1187,1229: decimated copy of the mask over the visited cells
            new_mask_grid = np.zeros((int(self.nx / nn), int(self.ny / nn)), dtype=bool)
            ii, jj = np.arange(0, self.nx - nn, nn), np.arange(0, self.ny - nn, nn)
            if len(ii) and len(jj):
                new_mask_grid[:len(ii), :len(jj)] = np.asarray(self.mask, dtype=bool)[np.ix_(ii, jj)]
            self.new_mask_grid = new_mask_grid
        self._leave_globals_like_reference(nn)
        self.chi2_grid = res['chi2']
        self.mean_Te_grid = res['mean_Te']
        self.std_Te_grid = res['std_Te']
        self.mean_F_grid = res['mean_F']
        self.std_F_grid = res['std_F']
        if self.alph:
            self.mean_a_grid = res['mean_a']
            self.std_a_grid = res['std_a']
        self.status_grid = res['status']

    def _leave_globals_like_reference(self, nn):
        ii, jj = range(0, self.nx - nn, nn), range(0, self.ny - nn, nn)
        last = None
        for i in reversed(ii):
            for j in reversed(jj):
                if self.mask is None or not self.mask[i, j]:
                    last = (i, j)
                    break
            if last:
                break
        if last is None:
            return
        cf_f.wd = self.water_depth[last]
        if self.rhoc is not None:
            cf_f.rhoc = self.rhoc[last]
        if self.zc is not None:
            cf_f.zc = self.zc[last]
        cf_f.rhof = cf_f.rhow if cf_f.wd > 0. else cf_f.rhoa

    def plot_admit_coh(self, *args, **kwargs):
        _out_of_scope('Project.plot_admit_coh')

    def plot_bayes_stats(self, *args, **kwargs):
        _out_of_scope('Project.plot_bayes_stats')

    def plot_functions(self, *args, **kwargs):
        _out_of_scope('Project.plot_functions')

    def plot_results(self, *args, **kwargs):
        _out_of_scope('Project.plot_results')

    def save_results(self, mean_Te=False, MAP_Te=False, std_Te=False, mean_F=False, MAP_F=False, std_F=False,
                     mean_a=False, MAP_a=False, std_a=False, chi2=False, mask=False, contours=None, filter=True,
                     sigma=1, prefix='PlateFlex_'):
        """classes.py:1621-1754: writes ``<prefix><label>.csv`` with columns x, y, <label> for every
        requested grid (x = dy*nn*column index, y = dx*nn*row index, row-major order, exactly the
        reference's ``save_grid``), Gaussian-smoothed first when ``filter`` is set.

        Behaviour kept from the reference, oddities included: the MAP_* grids exist only after a
        Bayesian estimation (an AttributeError propagates for MAP_Te / MAP_F as it does there); for the
        alpha grids and chi2 any failure prints the reference's message instead of raising -- and with
        ``filter=True`` the alpha grids always take that path, because the reference smooths a local
        variable it never assigned (classes.py:1713-1714, 1723-1724, 1733-1734)."""
        def save_grid(pars, grid, label, prefix):
            import pandas as pd
            xx, yy = np.mgrid[0:pars[0], 0:pars[1]]
            data = np.array([pars[3] * yy.flatten(), pars[2] * xx.flatten(), np.asarray(grid).flatten()]).T
            pd.DataFrame(data=data, columns=['x', 'y', label]).to_csv(prefix + label + '.csv', index=False)

        (nnx, nny) = self.mean_Te_grid.shape
        pars = [nnx, nny, self.dx * self.nn, self.dy * self.nn]
        if mask:
            try:
                save_grid(pars, self.new_mask_grid, 'mask', prefix)
            except Exception:
                print('No new mask found. Skipping')
        if contours is not None:
            contours = np.array(contours) / self.nn
        for want, label in ((mean_Te, 'mean_Te'), (MAP_Te, 'MAP_Te'), (std_Te, 'std_Te'), (mean_F, 'mean_F'),
                            (MAP_F, 'MAP_F'), (std_F, 'std_F')):
            if want:
                grid = getattr(self, label + '_grid')
                save_grid(pars, _gaussian(grid, sigma=sigma) if filter else grid, label, prefix)
        for want, label in ((mean_a, 'mean_a'), (MAP_a, 'MAP_a'), (std_a, 'std_a')):
            if want:
                try:
                    if filter:
                        raise UnboundLocalError(label + '_grid')  # the reference's unassigned local
                    save_grid(pars, getattr(self, label + '_grid'), label, prefix)
                except Exception:
                    print("parameter 'alpha' was not estimated")
        if chi2:
            try:
                save_grid(pars, _gaussian(self.chi2_grid, sigma=sigma) if filter else self.chi2_grid, 'chi2', prefix)
            except Exception:
                print("parameter 'chi2' was not estimated")


def _lam2k(nx, ny, dx, dy, p=0.85):
    """classes.py:1757-1828: geometric ladder of equivalent Fourier wavenumbers (rad/m) between
    a quarter of the grid diagonal and twice the cell diagonal, neighbouring wavelets
    overlapping at fractional amplitude ``p``.  Uses ``conf_cpwt.k0`` -- a float32 value.
    Known answer (classes.py:1787-1796): _lam2k(300, 300, 20., 20.) -> ns = 18."""
    k0 = cf_w.k0
    half_width = np.sqrt(-2. * np.log(p))
    maxlam = np.sqrt((nx * dx) ** 2. + (ny * dy) ** 2.) / 4. * 1.e3
    minlam = np.sqrt((2. * dx) ** 2. + (2. * dy) ** 2.) * 1.e3
    lam = [maxlam]
    k = [2. * np.pi / lam[0]]
    dk = half_width / (k0 / k[0])
    while lam[-1] > minlam:
        s = (k0 - half_width) / (k[-1] + dk)
        k.append(k0 / s)
        lam.append(2. * np.pi / k[-1])
        dk = half_width / s
    lam = np.array(lam)
    return len(lam), np.array(2. * np.pi / lam)
