"""Oracle for the per-cell L2 fit: the reference's estimate.L2_estimate_cell
logic (plateflex/estimate.py:227-409) driven by SciPy's bounded curve_fit
(method 'trf', 2-point Jacobian), with the model callback bound to the oracle's
flex.real_xspec_functions restatement.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl
reference).  Third-party arithmetic: scipy.optimize.curve_fit -- an un-vendored
and unpinned dependency of the reference (not even listed in its
pyproject.toml:18); the version importable here is recorded by ``scipy_version``.
No reference test pins results at this boundary => "parity unpinned".
"""
import numpy as np
import scipy
from scipy.optimize import curve_fit

from . import oracle as ora

scipy_version = scipy.__version__

# estimate.py:271,280,290,299
THETA0 = (20.0, 0.5, np.pi / 2.0)
LOWER = (2.0, 0.0001, 0.0001)
UPPER = (250.0, 0.9999, np.pi - 0.001)


def L2_estimate_cell(k, adm, eadm, coh, ecoh, alph=False, atype="joint", conf=None):
    """Returns dict(mean=[Te,F(,alpha)], std=[...], chi2=float, nfev=int).

    Mirrors estimate.py:267-407: y_obs/y_err selection by ``atype`` (joint =
    concatenation [adm, coh], estimate.py:350-351), alpha pinned to pi/2 unless
    ``alph`` (estimate.py:292,333,374), absolute_sigma=True, max_nfev=1000, the
    reduced chi-square of estimate.py:286-287 and std = sqrt(diag(pcov))
    (estimate.py:390).
    """
    conf = conf if conf is not None else ora.FlexConf()
    k = np.asarray(k, float)
    nfev = [0]

    def model(kk, Te, F, alpha=np.pi / 2.0):
        nfev[0] += 1
        a, c = ora.real_xspec_functions(kk, Te, F, alpha, conf)
        if atype == "admit":
            return a
        if atype == "coh":
            return c
        return np.array([a, c]).flatten()

    if atype == "admit":
        y_obs, y_err = np.asarray(adm), np.asarray(eadm)
    elif atype == "coh":
        y_obs, y_err = np.asarray(coh), np.asarray(ecoh)
    elif atype == "joint":
        y_obs = np.array([adm, coh]).flatten()
        y_err = np.array([eadm, ecoh]).flatten()
    else:
        raise ValueError(atype)
    npar = 3 if alph else 2
    if alph:
        def function(kk, Te, F, alpha):
            return model(kk, Te, F, alpha)
    else:
        def function(kk, Te, F):
            return model(kk, Te, F)
    p1fit, p1cov = curve_fit(function, k, y_obs, p0=np.array(THETA0[:npar]), sigma=y_err, absolute_sigma=True,
                             max_nfev=1000, bounds=(list(LOWER[:npar]), list(UPPER[:npar])))
    pred = model(k, *p1fit)
    rchi2 = np.sum((pred - y_obs) ** 2 / y_err ** 2) / (len(pred) - len(p1fit))
    return dict(mean=np.array(p1fit), std=np.sqrt(np.diag(p1cov)), chi2=float(rchi2), nfev=nfev[0])
