"""plateflex_b200 -- B200-native drop-in for PlateFlex's data-parallel hot path.

Same public names as the reference package (plateflex/__init__.py:137-224):
Grid, TopoGrid, GravGrid, BougGrid, FairGrid, RhocGrid, ZcGrid, Project,
``estimate``, ``set_conf_cpwt``, ``set_conf_flex``, ``get_conf_cpwt``,
``get_conf_flex``; the f2py modules ``plateflex.cpwt`` / ``plateflex.flex`` are
replaced by ``plateflex_b200.cpwt`` / ``plateflex_b200.flex`` which call
hand-written sm_100a CUDA kernels through libplateflex_b200.so.  Importing this
package without the built library raises ImportError: there is no CPU path.
"""
__version__ = '0.1.0'

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import estimate
from .classes import Grid, TopoGrid, GravGrid, BougGrid, FairGrid
from .classes import RhocGrid, ZcGrid, Project
from .cpwt import conf_cpwt as cf_wt
from .cpwt import get_output_dtype, set_output_dtype
from .flex import conf_flex as cf_fl


def set_conf_cpwt():
    """__init__.py:146-147."""
    cf_wt.k0 = 5.336


def set_conf_flex():
    """__init__.py:150-158."""
    cf_fl.zc = 35. * 1.e3
    cf_fl.rhom = 3200.
    cf_fl.rhoc = 2700.
    cf_fl.rhow = 1030.
    cf_fl.rhoa = 0.
    cf_fl.rhof = cf_fl.rhoa
    cf_fl.wd = 0.
    cf_fl.boug = 1


set_conf_cpwt()
set_conf_flex()


def get_conf_cpwt():
    """__init__.py:165-186."""
    print('\n'.join((
        'Wavelet parameter used in plateflex.cpwt:',
        '-----------------------------------------',
        '[Internal wavenumber]      k0 (float):     {0:.3f}'.format(cf_wt.k0))))


def get_conf_flex():
    """__init__.py:189-224."""
    rows = (('[Crustal thickness]        zc (float):     {0:.0f} m', cf_fl.zc),
            ('[Mantle density]           rhom (float):   {0:.0f} kg/m^3', cf_fl.rhom),
            ('[Crustal density]          rhoc (float):   {0:.0f} kg/m^3', cf_fl.rhoc),
            ('[Water density]            rhow (float):   {0:.0f} kg/m^3', cf_fl.rhow),
            ('[Air density]              rhoa (float):   {0:.0f} kg/m^3', cf_fl.rhoa),
            ('[Fluid density]            rhof (float):   {0:.0f} kg/m^3', cf_fl.rhof),
            ('[Water depth]              wd (float):     {0:.0f} m', cf_fl.wd))
    print('\n'.join(['Global model parameters currently in use by plateflex:',
                     '------------------------------------------------------'] +
                    [fmt.format(val) for fmt, val in rows] +
                    ['[Bouguer analysis?]        boug (int):     {0} ; {1}'.format(cf_fl.boug, bool(cf_fl.boug))]))
