"""Host mirror of fitting_Ct_functions.py: containers for C(t) = S2 + sum_i C_i exp(-t/tau_i) models, with the
least-squares solve moved to the GPU (sr_ct_fit_trf, one CTA per residue) and batched over residues.

Kept from the reference, by name: autoCorrelations (add_model, add_target, import_target_array, export, ...),
autoCorrelationModel (conduct_curve_fitting, optimised_curve_fitting, eval, calc_chiSq, set_nParams, report,
...), curvefit_exponential, read_fittedCt_parameters.  `fit_all_residues` runs the whole 2-3-5-7-9 ladder of
calculate-fitted-Ct.py:162-178 for every residue with one kernel launch per rung; the selection rules
(fitting_Ct_functions.py:278-304) and the quality flags with their evaluation-order quirk (G6, :329-340)
are applied per residue on the host exactly as the reference does.
"""
import sys
from collections import OrderedDict

import numpy as np

from . import _lib

# scipy.optimize.curve_fit's settings for the bounded ('trf') solver: SURVEY Appendix B
FTOL = XTOL = GTOL = 1e-8
MAX_NFEV = 0              # 0 = SciPy's default, 100 * nParams


def curvefit_exponential(DeltaT, *params):
    n = len(params)
    nn = int(n / 2)
    C = np.array(params[0:nn], dtype=float)
    tau = np.array(params[nn:2 * nn], dtype=float)
    S2 = params[-1] if n % 2 == 1 else 1.0 - np.sum(C)
    return S2 + np.sum(C[:, np.newaxis] * np.exp(-1.0 * DeltaT[np.newaxis, :] / tau[:, np.newaxis]), axis=0)


# ---- GPU solve ----------------------------------------------------------------------------------------
KERNEL_EVENTS = None      # set to a list to collect (start, end) CUDA event pairs around every sr_ct_fit_trf launch


class _ResidentCurves:
    """The (nR, L) curve arrays of a ladder, uploaded once per device and sliced there for every rung (the ladder
    revisits a shrinking subset of the same residues five times)."""

    def __init__(self, T, Y, SG):
        self.T, self.Y, self.SG = T, Y, SG
        self._dev = {}

    def on(self, d):
        torch = _lib.require_cuda()
        if d not in self._dev:
            dev = torch.device("cuda", d)
            f = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)  # noqa: E731
            self._dev[d] = (f(self.T), f(self.Y), f(self.SG))
        return self._dev[d]


def _device_solve(t, y, sigma, p0, lo, hi, resident=None, rows=None):
    """sr_ct_fit_trf on (nR, L) curves: returns popt (nR,nP), the R factor of the Jacobian at the solution (nR,nP,nP),
    cost (nR,), status (nR,2) = (SciPy termination status, nfev) and the reference's chi^2 at the solution (nR,).
    Residues are independent: with several GPUs selected (multigpu.devices) every device solves a contiguous block of
    them.  With `resident` (a _ResidentCurves) and `rows` the curves are taken from the device copies instead of t/y/sigma."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    from . import multigpu
    nR, nP = p0.shape
    if resident is None:
        L = y.shape[1]
        t = np.broadcast_to(t, (nR, L))
        sigma = None if sigma is None else np.broadcast_to(sigma, (nR, L))
    else:
        L = resident.Y.shape[1]
        rows = np.asarray(rows, dtype=np.int64)
    lo, hi = np.broadcast_to(lo, (nR, nP)), np.broadcast_to(hi, (nR, nP))

    def work(d, a, b):
        dev = torch.device("cuda", d)
        f = lambda x: torch.from_numpy(np.array(x[a:b], dtype=np.float64, order='C')).to(dev)   # noqa: E731
        n = b - a
        if resident is None:
            td, yd = f(t), f(y)
            sd = None if sigma is None else f(sigma)
        else:
            Td, Yd, Sd = resident.on(d)
            idx = torch.from_numpy(rows[a:b]).to(dev)
            whole = n == Yd.shape[0] and bool(np.array_equal(rows[a:b], np.arange(n)))
            td, yd = (Td, Yd) if whole else (Td.index_select(0, idx), Yd.index_select(0, idx))
            sd = None if Sd is None else (Sd if whole else Sd.index_select(0, idx))
        p0d, lod, hid = f(p0), f(lo), f(hi)
        popt = torch.empty((n, nP), dtype=torch.float64, device=dev)
        R = torch.empty((n, nP, nP), dtype=torch.float64, device=dev)
        cost = torch.empty(n, dtype=torch.float64, device=dev)
        chi = torch.empty(n, dtype=torch.float64, device=dev)
        status = torch.empty((n, 2), dtype=torch.int32, device=dev)
        wbytes = int(lib.sr_ct_fit_workspace_bytes(n, L, nP))
        work_buf = torch.empty(max(wbytes, 8) // 8, dtype=torch.float64, device=dev) if wbytes else None
        if KERNEL_EVENTS is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.check(lib.sr_ct_fit_trf(td.data_ptr(), yd.data_ptr(), 0 if sd is None else sd.data_ptr(), n, L, nP,
                                     p0d.data_ptr(), lod.data_ptr(), hid.data_ptr(), MAX_NFEV, FTOL, XTOL, GTOL,
                                     popt.data_ptr(), R.data_ptr(), cost.data_ptr(), status.data_ptr(), chi.data_ptr(),
                                     0 if work_buf is None else work_buf.data_ptr(), wbytes, _lib.current_stream_ptr()),
                   "sr_ct_fit_trf")
        if KERNEL_EVENTS is not None:
            ev[1].record()
            KERNEL_EVENTS.append(ev)
        return popt.cpu().numpy(), R.cpu().numpy(), cost.cpu().numpy(), status.cpu().numpy(), chi.cpu().numpy()

    res = multigpu.run(multigpu.plan(nR, min_per_device=32), work)
    return tuple(np.concatenate([r[k] for r in res], axis=0) for k in range(5))


def pcov_from_R(R, cost, L):
    """Covariance as scipy.optimize.curve_fit forms it from svd(J) (singular values <= eps*max(M,n)*s0 dropped,
    pcov = V S^-2 V^T * 2 cost/(M-n)), batched over the residues.  R is the triangular factor of J = Q R: same
    singular values and right singular vectors as J, without squaring the condition number."""
    nP = R.shape[-1]
    bad = ~np.all(np.isfinite(R), axis=(1, 2))         # keep LAPACK away from NaNs; those rows are marked below
    if np.any(bad):
        R = R.copy()
        R[bad] = 0.0
    _, sv, VT = np.linalg.svd(R)
    keep = sv > (np.finfo(float).eps * max(L, nP)) * sv[:, :1]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(keep, 1.0 / (sv * sv), 0.0)
    pcov = np.einsum("rki,rk,rkj->rij", VT, inv, VT)
    if L > nP:
        pcov *= (2.0 * cost / (L - nP))[:, None, None]
    else:
        pcov[:] = np.inf
    pcov[bad] = np.inf                                  # curve_fit: indeterminate covariance is filled with inf
    return pcov


def fit_succeeded(status):
    """curve_fit raises (and the reference's bare `except` turns that into chi = inf, bQuality[0] = False,
    fitting_Ct_functions.py:325-328) when least_squares does not report success: termination status 0 = max_nfev
    reached, or the ValueErrors for an infeasible p0 / non-finite residuals (negative codes here)."""
    return np.asarray(status)[..., 0] > 0


def gpu_curve_fit(t, y, sigma, p0, lo, hi, **resident):
    """Batched bounded least squares.  t, y, sigma: (nR, L) (sigma may be None); p0, lo, hi: (nR, nP).
    Returns popt (nR,nP), pcov (nR,nP,nP) formed like scipy.optimize.curve_fit, cost (nR,), status (nR,2) and the
    reference's chi^2 of the solution for zeta = 1 (nR,); rows with fit_succeeded(status) False are the ones curve_fit
    would have raised on.  `resident=..., rows=...`: see _device_solve."""
    p0 = np.atleast_2d(p0)
    if not resident:
        t, y = np.atleast_2d(t), np.atleast_2d(y)
    out = _device_solve(t, y, sigma, p0, lo, hi, **resident)
    popt, R, cost, status = out[:4]
    L = resident["resident"].Y.shape[1] if resident else y.shape[1]
    return (popt, pcov_from_R(R, cost, L), cost, status) + tuple(out[4:5])


_REAL_SOLVERS = None      # (gpu_curve_fit, _device_solve) as defined here; tests swap either for a CPU stand-in


_REAL_SOLVERS = (gpu_curve_fit, _device_solve)


# ---- containers ------------------------------------------------------------------------------------
class autoCorrelationModel:
    """One multi-exponential C(t) model (fitting_Ct_functions.py:128-416)."""
    dictGreek = np.array(['a', 'b', 'g', 'd', 'e', 'z', 'h'])

    def __init__(self, name='Fit', listC=[], listTau=[], S2=None, bS2Fast=False, bSort=True):
        self.name = name
        self.tau = np.array(listTau, dtype=float)
        self.C = np.array(listC, dtype=float)
        self.bS2Fast = bS2Fast
        self.S2 = S2
        self.nComps = len(self.C)
        self.nParams = len(self.C) + len(self.tau) + (1 if bS2Fast else 0)
        self.bHasFit = False
        self.zeta = 1.0
        if bS2Fast and self.S2 is None:
            print("= = = ERROR: S2 must be given in fitPatam initialisation is bS2Fast is set to True!")
            sys.exit(1)
        if self.S2 is None:
            self.S2 = 1.0 - np.sum(self.C)
        self.check_consistency()
        if self.nComps > 1 and bSort:
            self.sort_components()

    def check_consistency(self):
        if self.nComps < 1:
            return
        if len(self.C) != len(self.tau):
            print("= = = ERROR: transient components in fitParam initialisation do not have matching number of parameters!")
            sys.exit(1)
        if not self.bS2Fast and not np.all(np.isclose(self.S2 + np.sum(self.C), 1.0, rtol=1e-6)):
            print("= = = ERROR: Contribution of components in fitParam initialisation do not sum sufficeintly close to 1.00!")
            sys.exit(1)

    def copy(self):
        new = autoCorrelationModel()
        new.copy_from(self)
        return new

    def copy_from(self, src):
        self.name, self.nParams, self.nComps = src.name, src.nParams, src.nComps
        self.tau, self.C = np.copy(src.tau), np.copy(src.C)
        self.bS2Fast, self.S2, self.bHasFit = src.bS2Fast, src.S2, src.bHasFit
        if src.bHasFit:
            self.set_uncertainties_from_list(src.get_uncertainties_as_list())
            self.chiSq = src.chiSq

    def add_transient_component(self, C, tau):
        self.tau, self.C = np.append(self.tau, tau), np.append(self.C, C)
        self.nComps += 1
        self.nParams += 2

    def calc_S2Fast(self):
        return 1.0 - self.S2 - np.sum(self.C) if self.bS2Fast else 0.0

    def sort_components(self):
        inds = np.argsort(self.tau)
        self.tau, self.C = self.tau[inds], self.C[inds]
        if self.bHasFit:
            self.dtau, self.dC = self.dtau[inds], self.dC[inds]

    def set_zeta(self, zeta):
        self.zeta = zeta

    def get_zeta(self):
        return self.zeta

    def report(self, style='stdout', fp=sys.stdout):
        g = autoCorrelationModel.dictGreek
        if style == 'stdout':
            print("Name: %s" % self.name, file=fp)
            if self.bHasFit:
                print('  chi-Square: %g ' % self.chiSq, file=fp)
            if self.bS2Fast:
                print("  S2_fast: %g" % self.calc_S2Fast(), file=fp)
            for i in range(self.nComps):
                if self.bHasFit:
                    print("  component %s, const.: %g +- %g" % (g[i], self.C[i], self.dC[i]), file=fp)
                    print("  component %s, tau: %g +- %g" % (g[i], self.tau[i], self.dtau[i]), file=fp)
                else:
                    print("  component %s, const.: %g " % (g[i], self.C[i]), file=fp)
                    print("  component %s, tau: %g " % (g[i], self.tau[i]), file=fp)
            print(("  S2_0: %g +- %g" % (self.S2, self.dS2)) if self.bHasFit else ("  S2_0: %g" % self.S2), file=fp)
        elif style == 'xmgrace':
            print('# Residue: %s ' % self.name, file=fp)
            if self.bHasFit:
                print('# Chi-Square: %g ' % self.chiSq, file=fp)
                if self.bS2Fast:
                    print('# Param S2_fast: %g +- 0.0' % self.calc_S2Fast(), file=fp)
                    print('# Param S2_0: %g +- %g' % (self.S2, self.dS2), file=fp)
                else:
                    print('# Param S2_0: %g +- 0.0' % self.S2, file=fp)
                for i in range(self.nComps):
                    print('# Param C_%s: %g +- %g' % (g[i], self.C[i], self.dC[i]), file=fp)
                    print('# Param tau_%s: %g +- %g' % (g[i], self.tau[i], self.dtau[i]), file=fp)
            else:
                if self.bS2Fast:
                    print('# Param S2_fast: %g' % self.calc_S2Fast(), file=fp)
                print('# Param S2_0: %g' % self.S2, file=fp)
                for i in range(self.nComps):
                    print('# Param C_%s: %g' % (g[i], self.C[i]), file=fp)
                    print('# Param tau_%s: %g' % (g[i], self.tau[i]), file=fp)
        else:
            print("= = = ERROR: fitParam.report() does not recognise the style argument! "
                  "Choices are: stdout, xmgrace", file=sys.stderr)

    def eval(self, DeltaT):
        return self.zeta * (self.S2 + np.sum(self.C[:, np.newaxis] * np.exp(-1.0 * DeltaT[np.newaxis, :]
                                                                             / self.tau[:, np.newaxis]), axis=0))

    def calc_chiSq(self, DeltaT, Decay, dDecay=None):
        if dDecay is None:
            return np.mean(np.square(self.eval(DeltaT) - Decay))
        return np.mean(np.square(self.eval(DeltaT) - Decay) / dDecay)

    # -- parameter plumbing (:376-416) --
    def set_nParams(self, n):
        self.nParams = n
        self.nComps = int(n / 2)
        self.bS2Fast = (n % 2 == 1)

    def get_params_as_list(self):
        return list(self.C) + list(self.tau) + ([self.S2] if self.bS2Fast else [])

    def set_params_from_list(self, l):
        self.C = l[0:self.nComps]
        self.tau = l[self.nComps:2 * self.nComps]
        self.S2 = l[-1] if self.bS2Fast else 1.0 - np.sum(self.C)

    def get_uncertainties_as_list(self):
        return list(self.dC) + list(self.dtau) + ([self.dS2] if self.bS2Fast else [])

    def set_uncertainties_from_list(self, l):
        self.dC = np.array(l[0:self.nComps], dtype=float)
        self.dtau = np.array(l[self.nComps:2 * self.nComps], dtype=float)
        self.dS2 = l[-1] if self.bS2Fast else 0.0

    def get_bounds_as_list(self, tauMax=np.inf):
        return (0.0, [1.0] * self.nComps + [tauMax] * self.nComps + ([1.0] if self.bS2Fast else []))

    def initialise_for_fit_basic(self, tMax, tStep, nParams=None):
        if nParams is not None:
            self.set_nParams(nParams)
        self.tau = np.logspace(np.log10(tStep), np.log10(tMax * 2.0), self.nComps + 2)[1:-1]
        self.C = [1.0 / (self.nComps + 1)] * self.nComps
        self.S2 = 1.0 / (self.nComps + 1)
        self.bHasFit = False

    def initialise_for_fit_advanced(self, DeltaT, Decay, nParams=None, nSample=10):
        if nParams is not None:
            self.set_nParams(nParams)
        self.tau = np.logspace(np.log10(np.mean(DeltaT[1:] - DeltaT[:-1])), np.log10(DeltaT[-1] * 2.0),
                               self.nComps + 2)[1:-1]
        avgBeg, avgEnd = np.mean(Decay[:nSample]), np.mean(Decay[-nSample:])
        self.C = [np.fabs(avgBeg - avgEnd) / self.nComps] * self.nComps
        self.S2 = avgEnd if self.bS2Fast else 1.0 - np.mean(self.C)
        self.bHasFit = False

    # -- fitting (:278-345) --
    def _absorb_fit(self, paramOpt, dParamMatrix, DeltaT, Decay, dDecay, fp):
        """Everything conduct_curve_fitting does after curve_fit returns (:329-345), order preserved."""
        bQuality = [True, True, True]
        dParam = np.sqrt(np.diag(dParamMatrix))
        if not self.bS2Fast:
            self.S2 = 1.0 - np.sum(self.C)
        if np.any(dParam > paramOpt):
            print("= = = WARNING, curve fitting of %s with %i params indicates overfitting." % (self.name, self.nParams), file=fp)
            bQuality[1] = False
        if self.S2 + np.sum(self.C) > 1.0:
            print("= = = WARNING, curve fitting of %s with %i params returns sum>1." % (self.name, self.nParams), file=fp)
            bQuality[2] = False
        self.set_params_from_list(paramOpt)
        self.set_uncertainties_from_list(dParam)
        self.bHasFit = True
        self.chiSq = self.calc_chiSq(DeltaT, Decay, dDecay)
        self.sort_components()
        return self.chiSq, bQuality

    def conduct_curve_fitting(self, DeltaT, Decay, dDecay=None, bReInitialise=False, fp=sys.stdout):
        if bReInitialise:
            self.initialise_for_fit_advanced(DeltaT, Decay)
        lo, hi = self.get_bounds_as_list(tauMax=DeltaT[-1] * 10)
        p0 = np.array(self.get_params_as_list(), dtype=float)
        popt, pcov, cost, status = gpu_curve_fit(DeltaT, Decay, dDecay, p0[None, :], np.full_like(p0, lo)[None, :],
                                                 np.array(hi, dtype=float)[None, :])[:4]
        if not fit_succeeded(status)[0]:
            print("= = = WARNING, curve fitting of %s with %i params failed!" % (self.name, self.nParams), file=fp)
            return np.inf, [False, True, True]
        return self._absorb_fit(popt[0], pcov[0], DeltaT, Decay, dDecay, fp)

    def optimised_curve_fitting(self, DeltaT, Decay, dDecay=None, listDoG=[2, 3, 5, 7, 9], chiSqThreshold=0.5, fp=sys.stdout):
        print("= = = Conducting optimised fit for %s with %s degrees of freedoms..." % (self.name, str(listDoG)), file=fp)
        bFirst = True
        prev = self.copy()
        for nParams in listDoG:
            self.set_nParams(nParams)
            chiSq, bQuality = self.conduct_curve_fitting(DeltaT, Decay, dDecay, bReInitialise=True, fp=fp)
            print("    ...fit with %i params yield chiSq of %g" % (nParams, chiSq), file=fp)
            if bFirst:
                if np.all(bQuality):
                    prev.copy_from(self)
                    bFirst = False
                continue
            if not np.all(bQuality):
                print("    ...fit with %i params failed >0 quality checks, will stop." % nParams, file=fp)
                break
            if chiSq >= prev.chiSq * chiSqThreshold:
                print("    ...fit with %i params did not show sufficiently improved chi values. Will stop." % nParams, file=fp)
                break
            prev.copy_from(self)
        if bFirst:
            print("    ...ERROR: fit with %i params has never generated a satisfactory outcome!" % nParams, file=fp)
        else:
            self.copy_from(prev)
        return self.chiSq


class autoCorrelations:
    """Set of models + target curves (fitting_Ct_functions.py:12-126)."""

    def __init__(self):
        self.nModels = 0
        self.model = OrderedDict()
        self.nTargets = 0
        self.DeltaT, self.Decay, self.dDecay = OrderedDict(), OrderedDict(), OrderedDict()

    def get_names(self):
        return np.array([k for k in self.model.keys()])

    def get_params_as_list(self):
        ms = list(self.model.values())
        return [m.S2 for m in ms], [m.C for m in ms], [m.tau for m in ms], [m.calc_S2Fast() for m in ms]

    def set_zeta(self, zeta):
        for m in self.model.values():
            m.set_zeta(zeta)

    def get_zeta(self):
        for m in self.model.values():
            return m.get_zeta()

    def add_model(self, key, name=None, listC=[], listTau=[], S2=None, bS2Fast=False, bSort=True):
        self.model[key] = autoCorrelationModel(key if name is None else name, listC, listTau, S2, bS2Fast, bSort)
        self.nModels = len(self.model)
        return self.model[key]

    def get_nth_model(self, n):
        return self.model[self.get_names()[n]]

    def remove_model(self, key=None, index=None):
        if key is not None:
            self.model.pop(key)
        elif index is not None:
            self.model.pop(list(self.model.keys())[index])
        else:
            print("= = = ERROR in autoCorrelations.remove_model(); it needs at least one optional argument!)", file=sys.stderr)
            return
        self.nModels = len(self.model)

    def rename_models(self, listNames):
        if len(listNames) != len(self.model):
            print("= = = ERROR in autoCorrelations.rename_model(); length of lists are not equal!", file=sys.stderr)
            return
        for k, n in zip(self.model.keys(), listNames):
            self.model[k].name = n

    def report(self):
        print("Number of C(t) models loaded:", self.nModels)
        print("Number of targets loaded:", self.nTargets)

    def report_all_models(self):
        for m in self.model.values():
            m.report()

    def add_target(self, key, DeltaT, Decay, dDecay):
        self.DeltaT[key], self.Decay[key], self.dDecay[key] = DeltaT, Decay, dDecay
        self.nTargets = len(self.DeltaT)

    def import_target_array(self, keys, DeltaT, Decay, dDecay=None):
        for i, k in enumerate(keys):
            self.add_target(k, DeltaT[i], Decay[i], None if dDecay is None else dDecay[i])

    def rescale_time(self, f):
        for k in self.model.keys():
            self.model[k].tau *= f
            if self.nTargets > 0:
                self.DeltaT[k] *= f

    def export(self, fileName, style='xmgrace'):
        with open(fileName, 'w') as fp:
            s = 0
            for k, m in self.model.items():
                m.report(style='xmgrace', fp=fp)
                dt, Ct = self.DeltaT[k], self.Decay[k]
                ymodel = m.eval(dt)
                print("@s%d legend \"Res %d\"" % (s, m.name), file=fp)
                for j in range(len(ymodel)):
                    print("%8g %8g" % (dt[j], ymodel[j]), file=fp)
                print('&', file=fp)
                for j in range(len(ymodel)):
                    print("%8g %8g" % (dt[j], Ct[j]), file=fp)
                print('&', file=fp)
                s += 2

    # -- batched ladder: one kernel launch per rung for all residues still climbing --
    def fit_all_residues(self, listDoG=(2, 3, 5, 7, 9), chiSqThreshold=0.5, fp=sys.stdout, single=False):
        """calculate-fitted-Ct.py:162-178 for every target at once: one kernel launch per rung for all residues still
        climbing the ladder, and the reference's per-residue bookkeeping (initial guesses :359-374, quality flags and
        chi^2 :329-345, selection :278-304) evaluated as array operations over those residues.  Adds/overwrites one
        model per target key.  `fit_all_residues_loop` is the same ladder residue by residue through the model
        objects (kept as the specification; the two are tested to agree exactly)."""
        keys = list(self.DeltaT.keys())
        if single or len({len(self.DeltaT[k]) for k in keys}) != 1:
            return self.fit_all_residues_loop(listDoG, chiSqThreshold, fp, single)
        n = len(keys)
        T = np.array([self.DeltaT[k] for k in keys], dtype=float)
        Y = np.array([self.Decay[k] for k in keys], dtype=float)
        has_sig = self.dDecay[keys[0]] is not None
        SG = np.array([self.dDecay[k] for k in keys], dtype=float) if has_sig else None
        work = {k: self.model[k] if k in self.model else self.add_model(k, name=_as_name(k)) for k in keys}
        zeta = np.array([work[k].zeta for k in keys], dtype=float)
        names = [work[k].name for k in keys]
        dtm = np.mean(T[:, 1:] - T[:, :-1], axis=1)                  # initialise_for_fit_advanced (:365-373)
        avgBeg, avgEnd = np.mean(Y[:, :10], axis=1), np.mean(Y[:, -10:], axis=1)
        first = np.ones(n, dtype=bool)
        maxc = max(int(p / 2) for p in listDoG)
        best = dict(nParams=np.zeros(n, dtype=int), C=np.zeros((n, maxc)), tau=np.zeros((n, maxc)), S2=np.zeros(n),
                    dC=np.zeros((n, maxc)), dtau=np.zeros((n, maxc)), dS2=np.zeros(n), chi=np.full(n, np.inf),
                    fit=np.zeros(n, dtype=bool))
        last = {key: val.copy() for key, val in best.items()}        # state of the last rung tried (for residues never accepted)
        active = np.arange(n)
        lines = []
        resident = None
        used_device_chi = False
        for nParams in listDoG:
            if active.size == 0:
                break
            a = active
            nc, fast = int(nParams / 2), (nParams % 2 == 1)
            tau0 = np.power(10.0, np.linspace(np.log10(dtm[a]), np.log10(T[a, -1] * 2.0), nc + 2, axis=-1))[:, 1:-1]
            C0 = np.repeat((np.fabs(avgBeg[a] - avgEnd[a]) / nc)[:, None], nc, axis=1)
            S20 = avgEnd[a] if fast else 1.0 - np.mean(C0, axis=1)
            p0 = np.concatenate((C0, tau0) + ((S20[:, None],) if fast else ()), axis=1)
            hi = np.concatenate((np.ones((a.size, nc)), np.repeat((T[a, -1] * 10)[:, None], nc, axis=1)) +
                                ((np.ones((a.size, 1)),) if fast else ()), axis=1)
            if (gpu_curve_fit, _device_solve) == _REAL_SOLVERS:       # curves stay on the device across the rungs
                if resident is None:
                    resident = _ResidentCurves(T, Y, SG)
                res = gpu_curve_fit(None, None, None, p0, np.zeros_like(p0), hi, resident=resident, rows=a)
            else:
                res = gpu_curve_fit(T[a], Y[a], None if SG is None else SG[a], p0, np.zeros_like(p0), hi)
            popt, pcov, cost, status = res[:4]
            chi_dev = res[4] if len(res) > 4 and np.all(zeta[a] == 1.0) else None
            used_device_chi = used_device_chi or chi_dev is not None
            finite = fit_succeeded(status)                            # False where curve_fit raises upstream (:325-328)
            with np.errstate(invalid="ignore"):
                dParam = np.sqrt(np.diagonal(pcov, axis1=1, axis2=2))
                # quality flags as conduct_curve_fitting raises them, i.e. on the initial-guess state (:329-337, quirk G6)
                sumC0 = np.sum(C0, axis=1)
                S2_state = S20 if fast else 1.0 - sumC0
                q_over = ~np.any(dParam > popt, axis=1)
                q_sum = ~((S2_state + sumC0) > 1.0)
                C, tau = popt[:, :nc], popt[:, nc:2 * nc]
                S2 = popt[:, -1] if fast else 1.0 - np.sum(C, axis=1)
                chi = chi_dev if chi_dev is not None else _chi_rows(
                    T if a.size == n else T[a], Y if a.size == n else Y[a],
                    None if SG is None else (SG if a.size == n else SG[a]), C, tau, S2, zeta[a])
            chi = np.where(finite, chi, np.inf)
            q_over, q_sum = np.where(finite, q_over, True), np.where(finite, q_sum, True)
            ok = finite & q_over & q_sum
            order = np.argsort(tau, axis=1)                           # sort_components
            take = lambda x: np.take_along_axis(x, order, axis=1)     # noqa: E731
            Cs, taus = take(C), take(tau)
            dCs, dtaus = take(dParam[:, :nc]), take(dParam[:, nc:2 * nc])
            dS2 = dParam[:, -1] if fast else np.zeros(a.size)
            for i, fin, qo, qs, c in zip(a.tolist(), finite.tolist(), q_over.tolist(), q_sum.tolist(), chi.tolist()):
                if not fin:
                    lines.append("= = = WARNING, curve fitting of %s with %i params failed!" % (names[i], nParams))
                else:
                    if not qo:
                        lines.append("= = = WARNING, curve fitting of %s with %i params indicates overfitting." % (names[i], nParams))
                    if not qs:
                        lines.append("= = = WARNING, curve fitting of %s with %i params returns sum>1." % (names[i], nParams))
                lines.append("    ...%s: fit with %i params yield chiSq of %g" % (names[i], nParams, c))
            # selection ladder (optimised_curve_fitting :288-304)
            was_first = first[a]
            accept = np.where(was_first, ok, ok & ~(chi >= best["chi"][a] * chiSqThreshold))
            still = was_first | accept
            first[a[was_first & ok]] = False

            def store(dst, rows, sel):
                dst["nParams"][rows] = nParams
                dst["C"][rows, :nc], dst["tau"][rows, :nc] = Cs[sel], taus[sel]
                dst["dC"][rows, :nc], dst["dtau"][rows, :nc] = dCs[sel], dtaus[sel]
                dst["S2"][rows], dst["dS2"][rows], dst["chi"][rows], dst["fit"][rows] = S2[sel], dS2[sel], chi[sel], True

            store(best, a[accept], accept)
            # residues that have never produced an acceptable fit keep the state of their latest attempt
            pend = first[a]
            got = pend & finite
            store(last, a[got], got)
            miss = pend & ~finite
            if np.any(miss):                                          # failed fit: the model stays at its initial guess
                rows = a[miss]
                last["nParams"][rows], last["fit"][rows] = nParams, False
                last["C"][rows, :nc], last["tau"][rows, :nc], last["S2"][rows] = C0[miss], tau0[miss], S20[miss]
            active = a[still]
        if lines:
            print("\n".join(lines), file=fp)
        if used_device_chi:
            # the rung decisions and the log used the kernel's chi^2 (same quantity, summed in another order); the value
            # kept with each model is the reference's own arithmetic, evaluated once for the rung that was kept
            for src in (best, last):
                for nParams in listDoG:
                    rows = np.nonzero(src["fit"] & (src["nParams"] == nParams))[0]
                    if rows.size:
                        nc = int(nParams / 2)
                        src["chi"][rows] = _chi_rows(T[rows], Y[rows], None if SG is None else SG[rows], src["C"][rows, :nc],
                                                     src["tau"][rows, :nc], src["S2"][rows], zeta[rows])
        out = np.full(n, np.inf)
        for i, k in enumerate(keys):
            src = last if first[i] else best
            if first[i]:
                print("    ...ERROR: fit of %s has never generated a satisfactory outcome!" % str(k), file=fp)
            m, nc = work[k], int(src["nParams"][i] / 2)
            m.set_nParams(int(src["nParams"][i]))
            m.C, m.tau, m.S2 = src["C"][i, :nc].copy(), src["tau"][i, :nc].copy(), float(src["S2"][i])
            m.bHasFit = bool(src["fit"][i])
            if m.bHasFit:
                m.dC, m.dtau = src["dC"][i, :nc].copy(), src["dtau"][i, :nc].copy()
                m.dS2 = float(src["dS2"][i]) if m.bS2Fast else 0.0
                m.chiSq = float(src["chi"][i])
                out[i] = m.chiSq
        return out

    def fit_all_residues_loop(self, listDoG=(2, 3, 5, 7, 9), chiSqThreshold=0.5, fp=sys.stdout, single=False):
        """calculate-fitted-Ct.py:162-178 for every target at once.  Adds/overwrites one model per target key.
        single=True reproduces the fixed-parameter-count branch (:174-178): one conduct_curve_fitting per residue,
        result kept whatever its quality flags."""
        keys = list(self.DeltaT.keys())
        L = {len(self.DeltaT[k]) for k in keys}
        if len(L) != 1:
            raise ValueError("fit_all_residues: all target curves must have the same number of points")
        T = np.array([self.DeltaT[k] for k in keys], dtype=float)
        Y = np.array([self.Decay[k] for k in keys], dtype=float)
        has_sig = self.dDecay[keys[0]] is not None
        SG = np.array([self.dDecay[k] for k in keys], dtype=float) if has_sig else None
        work = {k: self.model[k] if k in self.model else self.add_model(k, name=_as_name(k)) for k in keys}
        prev = {k: work[k].copy() for k in keys}
        first = {k: True for k in keys}
        active = list(range(len(keys)))
        for nParams in listDoG:
            if not active:
                break
            p0s, his = [], []
            for i in active:
                m = work[keys[i]]
                m.set_nParams(nParams)
                m.initialise_for_fit_advanced(T[i], Y[i])
                p0s.append(m.get_params_as_list())
                his.append(m.get_bounds_as_list(tauMax=T[i][-1] * 10)[1])
            p0s, his = np.array(p0s, dtype=float), np.array(his, dtype=float)
            popt, pcov, cost, status = gpu_curve_fit(T[active], Y[active], None if SG is None else SG[active], p0s,
                                                     np.zeros_like(p0s), his)[:4]
            still = []
            for j, i in enumerate(active):
                k = keys[i]
                m = work[k]
                if not fit_succeeded(status)[j]:
                    print("= = = WARNING, curve fitting of %s with %i params failed!" % (m.name, nParams), file=fp)
                    chiSq, bQ = np.inf, [False, True, True]
                else:
                    chiSq, bQ = m._absorb_fit(popt[j], pcov[j], T[i], Y[i], None if SG is None else SG[i], fp)
                print("    ...%s: fit with %i params yield chiSq of %g" % (m.name, nParams, chiSq), file=fp)
                if single:
                    prev[k].copy_from(m)
                    first[k] = False
                    continue
                if first[k]:
                    if np.all(bQ):
                        prev[k].copy_from(m)
                        first[k] = False
                    still.append(i)
                    continue
                if not np.all(bQ) or chiSq >= prev[k].chiSq * chiSqThreshold:
                    continue
                prev[k].copy_from(m)
                still.append(i)
            active = still
        for k in keys:
            if first[k]:
                print("    ...ERROR: fit of %s has never generated a satisfactory outcome!" % str(k), file=fp)
            else:
                work[k].copy_from(prev[k])
        return np.array([work[k].chiSq if work[k].bHasFit else np.inf for k in keys])


def _chi_rows(Ta, Ya, SGa, C, tau, S2, zeta):
    """calc_chiSq (fitting_Ct_functions.py:272-276) for a stack of residues: mean((zeta model - y)^2 [/ sigma]) with
    model = S2 + sum_c C_c exp(-t / tau_c), one component at a time in two (n, L) arrays -- the same operations in the
    same order as the reference's (nc, L) broadcast per residue, without its temporaries."""
    nc = C.shape[1]
    ntau = -tau                                   # t / (-tau) == (-1.0 * t) / tau exactly
    model = np.empty_like(Ta)
    buf = np.empty_like(Ta) if nc > 1 else None
    with np.errstate(all="ignore"):
        for c in range(nc):
            dst = model if c == 0 else buf
            np.divide(Ta, ntau[:, c:c + 1], out=dst)
            np.exp(dst, out=dst)
            np.multiply(C[:, c:c + 1], dst, out=dst)
            if c:
                model += buf
        model += S2[:, None]
        model *= zeta[:, None]
        model -= Ya
        np.square(model, out=model)
        if SGa is not None:
            model /= SGa
        return np.mean(model, axis=1)


def _as_name(k):
    try:
        return int(k)
    except (TypeError, ValueError):
        return k


def _get_key(index, var):
    return str(index) + "-" + var


def read_fittedCt_parameters(fileName):
    """Parser of *_fittedCt.dat (fitting_Ct_functions.py:432-481)."""
    obj = autoCorrelations()
    index = S2_slow = S2_fast = None
    tmpC, tmpTau = OrderedDict(), OrderedDict()
    bParamSection = False
    with open(fileName) as fp:
        for line in fp.readlines():
            if line.startswith("#"):
                l = line.split()
                if l[1].startswith("Residue"):
                    if bParamSection:
                        print("= = = ERROR in read_fittedCt_parameters: New parameter section detected when old "
                              "parameter section is still being read! %s " % fileName, file=sys.stderr)
                        sys.exit(1)
                    bParamSection = True
                    index = str(l[-1])
                elif l[1].startswith("Param"):
                    parName, value = l[2], float(l[-3])
                    if parName.startswith("S2_0"):
                        S2_slow = value
                    elif parName.startswith("S2_fast"):
                        S2_fast = value
                    elif parName.startswith("C_"):
                        tmpC[_get_key(index, parName[2])] = value
                    elif parName.startswith("tau_"):
                        tmpTau[_get_key(index, parName[4])] = value
            elif bParamSection:
                obj.add_model(index, S2=S2_slow, listC=[tmpC[k] for k in tmpC.keys()],
                              listTau=[tmpTau[k] for k in tmpC.keys()], bS2Fast=S2_fast is not None)
                bParamSection = False
                tmpC, tmpTau, S2_fast, S2_slow, index = OrderedDict(), OrderedDict(), None, None, None
    return obj
