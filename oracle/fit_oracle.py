"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

Restatement of the per-residue multi-exponential C(t) fit of fitting_Ct_functions.py.  The arithmetic
of the solver lives in third-party SciPy (`curve_fit` -> `least_squares(method='trf')`, version
unpinned upstream: requirements.txt:2 `scipy>=0.17.1`); the oracle calls the SciPy installed in this
image with the reference's exact call (fitting_Ct_functions.py:322-324).  Pinned by tests/golden.
"""
import numpy as np
from scipy.optimize import curve_fit


def model_curve(t, *p):
    """curvefit_exponential, fitting_Ct_functions.py:419-427. p = C_1..C_n, tau_1..tau_n [, S2]."""
    n = len(p)
    nc = n // 2
    C = np.array(p[:nc], dtype=float)
    tau = np.array(p[nc:2 * nc], dtype=float)
    S2 = p[-1] if n % 2 == 1 else 1.0 - np.sum(C)
    return S2 + np.sum(C[:, None] * np.exp(-t[None, :] / tau[:, None]), axis=0)


def initial_guess(t, y, n_params, n_sample=10):
    """set_nParams (:376-382) + initialise_for_fit_advanced (:359-374)."""
    nc = n_params // 2
    free_s2 = (n_params % 2 == 1)
    tau = np.logspace(np.log10(np.mean(t[1:] - t[:-1])), np.log10(t[-1] * 2.0), nc + 2)[1:-1]
    beg, end = np.mean(y[:n_sample]), np.mean(y[-n_sample:])
    C = [np.fabs(beg - end) / nc] * nc
    S2 = end if free_s2 else 1.0 - np.mean(C)
    p0 = list(C) + list(tau) + ([S2] if free_s2 else [])
    return p0, S2, np.array(C), tau


def bounds(n_params, tau_max):
    """get_bounds_as_list, :412-416."""
    nc = n_params // 2
    hi = [1.0] * nc + [tau_max] * nc + ([1.0] if n_params % 2 == 1 else [])
    return (0.0, hi)


def fit_once(t, y, dy, n_params):
    """conduct_curve_fitting(bReInitialise=True), fitting_Ct_functions.py:306-345.

    Returns dict(chi, quality[3], C, tau, S2, dC, dtau, dS2).  Quirk G6: the S2+sum(C)>1 flag (:336) and,
    for even n_params, S2 = 1 - sum(C) (:330-331) are evaluated on the INITIAL guess before the optimum
    is stored (:340)."""
    nc = n_params // 2
    free_s2 = (n_params % 2 == 1)
    p0, S2_init, C_init, _ = initial_guess(t, y, n_params)
    quality = [True, True, True]
    try:
        popt, pcov = curve_fit(model_curve, t, y, sigma=dy, p0=p0, bounds=bounds(n_params, t[-1] * 10))
    except Exception:
        quality[0] = False
        return dict(chi=np.inf, quality=quality)
    dp = np.sqrt(np.diag(pcov))
    S2_chk = S2_init if free_s2 else 1.0 - np.sum(C_init)
    if np.any(dp > popt):
        quality[1] = False
    if S2_chk + np.sum(C_init) > 1.0:
        quality[2] = False
    C, tau = popt[:nc], popt[nc:2 * nc]
    S2 = popt[-1] if free_s2 else 1.0 - np.sum(C)
    model = S2 + np.sum(C[:, None] * np.exp(-t[None, :] / tau[:, None]), axis=0)   # eval, zeta = 1 (:266-270)
    chi = np.mean(np.square(model - y) / dy) if dy is not None else np.mean(np.square(model - y))  # :272-276
    order = np.argsort(tau)
    return dict(chi=chi, quality=quality, C=C[order], tau=tau[order], S2=S2, dC=dp[:nc][order],
                dtau=dp[nc:2 * nc][order], dS2=(dp[-1] if free_s2 else 0.0), popt=popt, pcov=pcov,
                n_params=n_params)


def fit_ladder(t, y, dy, dof_list=(2, 3, 5, 7, 9), chi_threshold=0.5):
    """optimised_curve_fitting, fitting_Ct_functions.py:278-304: walk the parameter ladder, keep the last
    model that passed all quality flags and improved chi by the threshold factor."""
    prev = None
    first = True
    for n_params in dof_list:
        cur = fit_once(t, y, dy, n_params)
        if first:
            if all(cur["quality"]):
                prev = cur
                first = False
            continue
        if not all(cur["quality"]):
            break
        if cur["chi"] >= prev["chi"] * chi_threshold:
            break
        prev = cur
    return prev
