"""Host mirror of the spectral-density / relaxation entry points of spectral_densities.py, GPU-backed.

Same class and method names as the reference for the part of the module that is on the hot path:
  gyromag, gyromagMultiCSA                          :23-124
  angularFrequencies                                :136-250
  globalRotationalDiffusion_Isotropic / _Axisymmetric  :392-606 (import_frame_vectors[_npz], calc_Jomega[_one])
  spinRelaxationR1 / R2 / NOE (.eval(ind=None), .values, .errors)  :611-907
  _do_Jsum, D_coefficients_symmtop, convert_LambertCylindricalHist_to_vecs
plus `relax_grid`, the batched (residue x field x CSA) evaluation the rsCSA optimiser generates.
All J(omega) / R1 / R2 / NOE arithmetic and the weighted bin averaging run in sr_relax_a_moments and
sr_relax_eval; `_do_Jsum` and `calc_Jomega` go through the GPU `npufunc.Jomega`.
"""
import ctypes
import sys
from collections import OrderedDict

import numpy as np

from . import _lib, npufunc

_TIME = {'ps': 1.0e-12, 'ns': 1.0e-9, 'us': 1.0e-6, 'ms': 1.0e-3, 's': 1.0e-0}
_DIST = {'pm': 1.0e-12, 'A': 1.0e-10, 'nm': 1.0e-9, 'um': 1.0e-6, 'mm': 1.0e-3, 'm': 1.0e-0}
_GAMMA = {'1H': 267.513e6, '13C': 67.262e6, '15N': -27.116e6, '17O': -36.264e6, '19F': 251.662e6, '31P': 108.291e6}
_CSA0 = {'15N': -170e-6, '13C': -130e-6}


def _return_time_fact(tu):
    if tu not in _TIME:
        print("= = ERROR in relaxationModel: invalid time unit definition!", file=sys.stderr)
        return None
    return _TIME[tu]


def _return_dist_fact(du):
    if du not in _DIST:
        print("= = ERROR in relaxationModel: invalid distance unit definition!", file=sys.stderr)
        return None
    return _DIST[du]


class gyromag:
    """Gyromagnetic ratio (rad s^-1 T^-1) and CSA of one nucleus type (:23-79)."""

    def __init__(self, isotope, csa=None):
        self.num = 1
        self.isotope = isotope
        self.timeUnit = 's'
        self.time_fact = _return_time_fact('s')
        self.gamma = _GAMMA[isotope] * self.time_fact
        self.csa = _CSA0.get(isotope, 0.0) if csa is None else csa

    def reset_csa(self, name):
        self.set_csa(_CSA0.get(name, 0.0))

    def set_csa(self, csa, i=None):
        self.csa = csa

    def get_csa(self, i=None):
        return self.csa

    def set_time_unit(self, tu):
        old = self.time_fact
        self.time_fact = _return_time_fact(tu)
        self.gamma *= self.time_fact / old


class gyromagMultiCSA(gyromag):
    """Same with one CSA value per residue (:81-124)."""

    def __init__(self, isotope, n, csa=None):
        gyromag.__init__(self, isotope)
        self.num = n
        self.set_csa(np.repeat(_CSA0.get(isotope, 0.0), n) if csa is None else csa)

    def reset_csa(self, name):
        self.set_csa(np.repeat(_CSA0.get(name, 0.0), self.num))

    def set_csa(self, csa, ind=None):
        if ind is None:
            if self.num != len(csa):
                print("= = ERROR: attempting to set CSA array in gyromagMultiCSA, but the lengths to not match!")
                sys.exit(1)
            self.csa = np.array(csa)
        else:
            self.csa[ind] = csa

    def get_csa(self, ind=None):
        return self.csa if ind is None else self.csa[ind]


class angularFrequencies:
    """The five frequencies J is sampled at and the DD / CSA prefactors (:136-250)."""
    iOm0, iOmA, iOmBmA, iOmB, iOmBpA = 0, 1, 2, 3, 4

    def __init__(self, nucleiA='15N', nucleiB='1H', fieldStrength=600, fieldUnit='MHz', timeUnit='ps'):
        self.timeUnit = timeUnit
        self.time_fact = _return_time_fact(timeUnit)
        self.distUnit = 'nm'
        self.dist_fact = _return_dist_fact('nm')
        self.gA, self.gB = gyromag(nucleiA), gyromag(nucleiB)
        self.rAB = 1.02e-1
        self.B0 = None
        self.set_magnetic_field(fieldStrength, fieldUnit)
        self.nOmega = 5
        self.omegaNames = OrderedDict((k, i) for i, k in enumerate(
            ('0', nucleiA, nucleiB + '-' + nucleiA, nucleiB, nucleiB + '+' + nucleiA)))
        om = np.zeros(5)
        om[1] = -1.0 * self.gA.gamma * self.B0 * self.time_fact
        om[3] = -1.0 * self.gB.gamma * self.B0 * self.time_fact
        om[2] = om[3] - om[1]
        om[4] = om[3] + om[1]
        self.omega = om

    def set_magnetic_field(self, inp, unit):
        if unit == 'Hz':
            self.B0 = 2.0 * np.pi * inp / 267.513e6
        elif unit == 'MHz':
            self.B0 = 2.0 * np.pi * inp / 267.513
        elif unit == 'T':
            self.B0 = inp
        else:
            raise ValueError("set_magnetic_field: incorrect field units given ( %s )" % unit)

    def get_magnetic_field(self, unit='T'):
        return {'T': self.B0, 'MHz': self.B0 * 267.513 / (2.0 * np.pi), 'Hz': self.B0 * 267.513e6 / (2.0 * np.pi)}[unit]

    def set_time_unit(self, tu):
        old = self.time_fact
        self.time_fact = _return_time_fact(tu)
        self.timeUnit = tu
        self.omega *= self.time_fact / old

    def get_frequencies(self):
        return self.omega

    def get_nuclei_names(self):
        return [self.gA.isotope, self.gB.isotope]

    def get_factor_DD(self):
        return 0.10 * 1.1121216813552401e-82 * self.gA.gamma ** 2.0 * self.gB.gamma ** 2.0 * (self.rAB * self.dist_fact) ** -6.0

    def get_factor_CSA(self, i=None):
        return 2.0 / 15.0 * self.gA.get_csa(i) ** 2.0 * (self.gA.gamma * self.B0) ** 2

    def initialise_CSA_array(self, numCSAs, CSAvalues=None):
        self.gA = gyromagMultiCSA(self.gA.isotope, numCSAs, CSAvalues)

    def report(self):
        print("Field: %g T" % self.get_magnetic_field())
        print("Angular frequencies (rad %s^-1 T^-1): %s" % (self.timeUnit, str(self.omega)))


# ---- helpers kept under their reference names -----------------------------------------------------
def D_coefficients_symmtop(D):
    """(Dpar, Dperp) -> (5Dperp+Dpar, 2Dperp+4Dpar, 6Dperp) (:1874-1884)."""
    Dpar, Dperp = D[0], D[1]
    return np.array([5 * Dperp + Dpar, 2 * Dperp + 4 * Dpar, 6 * Dperp])


def A_coefficients_symmtop(v, bProlate=True):
    """(:1886-1905) A0 = 3 z^2 (1-z^2), A1 = 3/4 (1-z^2)^2, A2 = 1/4 (3 z^2 - 1)^2 with z the unique axis."""
    v = np.asarray(v)
    z2 = np.square(v.take(-1 if bProlate else 0, axis=-1))
    omz = 1 - z2
    return np.stack((3.0 * (z2 * omz), 0.75 * np.square(omz), 0.25 * np.square(3.0 * z2 - 1.0)), axis=-1)   # same rounding order


def _do_Jsum(om, A_J, D_J):
    """J = A_ij D_j/(D_j^2+om_k^2) (:1961-1972) with the Lorentzian table from the GPU ufunc."""
    return np.einsum('...j,jk', A_J, npufunc.Jomega.outer(D_J, om))


def convert_LambertCylindricalHist_to_vecs(hist, edges):
    """Bin-centre unit vectors and weights of a (nR, nphi, ncos) histogram (:2334-2350)."""
    phis = 0.5 * (edges[0][:-1] + edges[0][1:])
    thetas = np.arccos(0.5 * (edges[1][:-1] + edges[1][1:]))
    P, T = np.meshgrid(phis, thetas, indexing='ij')
    binVecs = np.stack((np.cos(P) * np.sin(T), np.sin(P) * np.sin(T), np.cos(T)), axis=-1)
    nRes, nPts = hist.shape[0], hist[0].shape[0] * hist[0].shape[1]
    return np.repeat(binVecs.reshape(nPts, 3)[np.newaxis, ...], nRes, axis=0), np.reshape(hist, (nRes, nPts))


# ---- global tumbling models -------------------------------------------------------------------------
class globalRotationalDiffusion_Base:
    def __init__(self):
        self.name = 'base'
        self.D = self.D_J = self.A_J = None
        self.bVecs = False
        self.axisAvg = self.vecNames = self.vecXH = self.vecWeights = None

    def import_frame_vectors_npz(self, fileName):
        obj = np.load(fileName, allow_pickle=True)
        if not obj['bHistogram']:
            print("= = = Only histogram-type vector distributions are supported on this path.", file=sys.stderr)
            sys.exit(1)
        if obj['dataType'] != 'LambertCylindrical':
            print("= = = Histogram projection not supported! %s" % obj['dataType'], file=sys.stderr)
            sys.exit(1)
        vecs, weights = convert_LambertCylindricalHist_to_vecs(obj['data'], obj['edges'])
        self.set_frame_vectors(obj['names'], vecs, weights)

    def set_frame_vectors(self, names, vecs, weights):
        """vecs (nR, B, 3), weights (nR, B): stored swapped to (B, nR, ...) like the reference (:302-306)."""
        self.bVecs = True
        self.vecNames = names
        vecs = np.asarray(vecs)
        # histogram-derived distributions use the same bin vectors for every residue (:2349): the A_J coefficients
        # are then needed once, not per residue
        self._shared_bins = bool(vecs.shape[0] > 0 and np.array_equal(vecs, np.broadcast_to(vecs[:1], vecs.shape)))
        self.vecXH = np.swapaxes(vecs, 0, 1)
        self.vecWeights = np.swapaxes(weights, 0, 1)
        self.axisAvg = 0
        self.update_A_coefficients()

    def import_frame_vectors(self, fileName):
        if not fileName.endswith('.npz'):
            print("= = = Only .npz vector distributions are supported on this path.", file=sys.stderr)
            sys.exit(1)
        self.import_frame_vectors_npz(fileName)

    def get_names(self):
        return [str(x) for x in self.vecNames]


class globalRotationalDiffusion_Isotropic(globalRotationalDiffusion_Base):
    def __init__(self, D=None, tau=None):
        globalRotationalDiffusion_Base.__init__(self)
        self.name = 'isotropic'
        self.D = D if D is not None else 1.0 / (6.0 * tau)
        self.D_J = self.D

    def get_Diso(self):
        return self.D

    def set_Diso(self, Diso):
        self.D = self.D_J = Diso

    def get_Daniso(self):
        return 1.0

    def set_Daniso(self, Daniso):
        return

    def update_A_coefficients(self):
        return

    def get_A_coefficients(self):
        return 1.0

    def get_D_coefficients(self):
        return self.D

    def transform_D(self):
        return self.D

    def calc_Jomega_one(self, omega, CtModel, ind=None):
        out = _gpu_relax(self, [CtModel], omega[None, :], np.zeros((1, 1)), 0.0, want="J")
        return out[0, 0]

    def calc_Jomega(self, omega, Autocorrs):
        return _gpu_relax(self, list(Autocorrs.model.values()), omega[None, :], np.zeros((1, 1)), 0.0, want="J")[:, 0]

    def calc_Jomega_rigid(self, omega):
        return npufunc.Jomega(np.full(len(omega), 6.0 * self.D), np.asarray(omega, dtype=float))


class globalRotationalDiffusion_Axisymmetric(globalRotationalDiffusion_Base):
    def __init__(self, D=None, bConvert=False, tau=None, aniso=None):
        globalRotationalDiffusion_Base.__init__(self)
        self.name = 'axisymmetric'
        if D is not None:
            self.D = np.array([(2.0 * D[1] + D[0]) / 3.0, D[0] / D[1]] if bConvert else D, dtype=float)
        else:
            self.D = np.array([1.0 / (6.0 * tau), aniso], dtype=float)
        self.bProlate = bool(self.D[1] > 1)
        self.update_D_coefficients()

    def set_Diso(self, Diso):
        self.D[0] = Diso
        self.update_D_coefficients()

    def set_Daniso(self, Daniso):
        self.D[1] = Daniso
        self.update_D_coefficients()

    def get_Diso(self):
        return self.D[0]

    def get_Daniso(self):
        return self.D[1]

    def update_A_coefficients(self):
        self.A_J = A_coefficients_symmtop(self.vecXH, self.bProlate)
        self._amom = None

    def get_A_coefficients(self, ind=None):
        return self.A_J if ind is None else np.take(self.A_J, ind, axis=-2)

    def transform_D(self):
        tmp = 3.0 * self.D[0] / (2.0 + self.D[1])
        return self.D[1] * tmp, tmp

    def update_D_coefficients(self):
        self.D_J = D_coefficients_symmtop(self.transform_D())

    def get_D_coefficients(self):
        return self.D_J

    def calc_Jomega_one(self, omega, CtModel, ind):
        D_J, A_J = self.get_D_coefficients(), self.get_A_coefficients(ind)
        Jmat = _do_Jsum(omega, CtModel.zeta * CtModel.S2 * A_J, D_J)
        for j in range(CtModel.nComps):
            Jmat += _do_Jsum(omega, CtModel.zeta * CtModel.C[j] * A_J, D_J + 1. / CtModel.tau[j])
        return Jmat

    def calc_Jomega(self, omega, Autocorrs, bSearch=False):
        sh = list(self.vecXH.shape)
        sh[-1] = len(omega)
        Jmat = np.zeros(sh)
        for i, model in enumerate(Autocorrs.model.values()):
            Jmat[..., i, :] = self.calc_Jomega_one(omega, model, i)
        return Jmat

    def calc_Jomega_rigid(self, omega):
        return self.get_A_coefficients() * self.D_J / (np.power(self.D_J, 2.0) + np.power(omega, 2.0))

    def a_moments(self, rows=None):
        """Per-residue weighted mean / covariance of A_J over the bins, computed once on the GPU.  `rows` = (first, last)
        computes (uncached) the moments of that block of residues on the current device: the multi-GPU grid."""
        if rows is None and getattr(self, "_amom", None) is not None:
            return self._amom
        torch = _lib.require_cuda()
        lib = _lib.load()
        shared = getattr(self, "_shared_bins", False)
        a, b = (0, self.vecWeights.shape[1]) if rows is None else rows
        if shared:
            A = np.ascontiguousarray(self.A_J[:, 0, :], dtype=np.float64)                    # (B, 3)
        else:
            A = np.ascontiguousarray(np.swapaxes(self.A_J[:, a:b], 0, 1), dtype=np.float64)  # (nR, B, 3)
        W = np.ascontiguousarray(np.swapaxes(self.vecWeights[:, a:b], 0, 1), dtype=np.float64)   # (nR, B)
        Ad, Wd = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
        out = torch.empty((W.shape[0], 10), dtype=torch.float64, device=Ad.device)
        _lib.check(lib.sr_relax_a_moments(Ad.data_ptr(), 0 if shared else 1, Wd.data_ptr(), W.shape[0], W.shape[1],
                                          out.data_ptr(), _lib.current_stream_ptr()), "sr_relax_a_moments")
        if rows is None:
            self._amom = out
        return out


# ---- batched GPU evaluation -----------------------------------------------------------------------
def _pack_models(models):
    nR = len(models)
    mc = max(1, max(m.nComps for m in models))
    S2 = np.array([m.S2 for m in models], dtype=np.float64)
    C = np.zeros((nR, mc))
    tau = np.ones((nR, mc))
    nc = np.zeros(nR, dtype=np.int32)
    for i, m in enumerate(models):
        nc[i] = m.nComps
        C[i, :m.nComps] = m.C
        tau[i, :m.nComps] = m.tau
    return S2, C, tau, nc, mc


def _gpu_relax(rotdif, models, omega, f_csa, f_dd, gammaA=-27.116e6, gammaB=267.513e6, time_fact=1e-12,
               csa_per_residue=False, want="R", subset=None, rows=None):
    """omega (nField,5); f_csa (nField,nCSA) or (nField,nR).  Returns (nR,nField,nCSA,6), or J (nR,nField,5)
    for the isotropic `want='J'` path (evaluated through R-linearity is not possible, so J is computed with
    the GPU Lorentzian table)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    S2, C, tau, nc, mc = _pack_models(models)
    zeta = models[0].zeta if models else 1.0
    nR, nField = len(models), omega.shape[0]
    if want == "J":     # isotropic J(omega): Lorentzians from the GPU ufunc, summed on the host (tiny)
        tg = 1.0 / (6.0 * rotdif.D)
        J = np.zeros((nR, nField, 5))
        for i, m in enumerate(models):
            k = np.concatenate(([1.0 / tg], 1.0 / tg + 1.0 / np.asarray(m.tau, dtype=float)))
            amp = np.concatenate(([m.S2], np.asarray(m.C, dtype=float))) * m.zeta
            J[i] = np.einsum('c,cfk->fk', amp, npufunc.Jomega(k[:, None, None], omega[None, :, :]))
        return J
    iso = isinstance(rotdif, globalRotationalDiffusion_Isotropic)
    nCSA = 1 if csa_per_residue else f_csa.shape[1]
    dev = torch.device("cuda")
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)   # noqa: E731
    amom = None
    if not iso:
        amom = rotdif.a_moments(rows)
        if subset is not None:
            amom = amom[torch.as_tensor(subset, device=dev, dtype=torch.long)].contiguous()
    D_J = (ctypes.c_double * 3)(*([rotdif.D, 0, 0] if iso else [float(x) for x in rotdif.D_J]))
    out = torch.empty((nR, nField, nCSA, 6), dtype=torch.float64, device=dev)
    S2d, Cd, taud, ncd, omd, fcd = t(S2), t(C), t(tau), t(nc, torch.int32), t(omega), t(f_csa)
    _lib.check(lib.sr_relax_eval(1 if iso else 0, D_J, zeta, time_fact, gammaA, gammaB, f_dd, nR, nField, nCSA,
                                 1 if csa_per_residue else 0, mc, 0 if amom is None else amom.data_ptr(),
                                 S2d.data_ptr(), Cd.data_ptr(), taud.data_ptr(), ncd.data_ptr(), omd.data_ptr(),
                                 fcd.data_ptr(), out.data_ptr(), _lib.current_stream_ptr()), "sr_relax_eval")
    return out.cpu().numpy()


def relax_grid(rotdif, Autocorrs, fields_mhz, csa_grid=None, nucleiA='15N', nucleiB='1H'):
    """R1, R2, NOE (+ sigmas) for every residue x magnetic field x CSA value in one GPU launch.
    Returns dict name -> (values, errors), each (nR, nField, nCSA)."""
    models = list(Autocorrs.model.values())
    ws = [angularFrequencies(nucleiA, nucleiB, f, 'MHz', 'ps') for f in fields_mhz]
    omega = np.array([w.omega for w in ws])
    csa = np.atleast_1d(np.asarray(ws[0].gA.csa if csa_grid is None else csa_grid, dtype=float))
    f_csa = np.array([2.0 / 15.0 * csa ** 2.0 * (w.gA.gamma * w.B0) ** 2 for w in ws])
    from . import multigpu
    blocks = multigpu.plan(len(models), min_per_device=128)
    if len(blocks) == 1:
        out = _gpu_relax(rotdif, models, omega, f_csa, ws[0].get_factor_DD(), ws[0].gA.gamma, ws[0].gB.gamma,
                         ws[0].time_fact)
    else:       # residues are independent: every selected GPU evaluates a block of them (moments pass included)
        out = np.concatenate(multigpu.run(blocks, lambda d, a, b: _gpu_relax(
            rotdif, models[a:b], omega, f_csa, ws[0].get_factor_DD(), ws[0].gA.gamma, ws[0].gB.gamma, ws[0].time_fact,
            rows=(a, b))), axis=0)
    iso = isinstance(rotdif, globalRotationalDiffusion_Isotropic)
    return {n: (out[..., i], None if iso else out[..., 3 + i]) for i, n in enumerate(("R1", "R2", "NOE"))}


# ---- experiments ------------------------------------------------------------------------------------
class spinRelaxationBase:
    """One NMR observable at one field over all residues (:611-818). eval() runs on the GPU."""
    _col = None

    def __init__(self, name, timeUnit='ps', angFreq=None, globalRotDif=None, localCtModels=None):
        self.name = name
        self.values = self.errors = None
        self.timeUnit = timeUnit
        self.time_fact = _return_time_fact(timeUnit)
        self.angFreq, self.globalRotDif, self.localCtModels = angFreq, globalRotDif, localCtModels
        if globalRotDif is not None and localCtModels is not None:
            self.reset_values()

    def get_num(self):
        return self.localCtModels.nModels

    def get_name(self):
        return self.name

    def set_magnetic_field(self, fieldStrength, fieldUnit):
        self.angFreq.set_magnetic_field(fieldStrength, fieldUnit)

    def get_magnetic_field(self):
        return self.angFreq.get_magnetic_field()

    def set_zeta(self, zeta):
        self.localCtModels.set_zeta(zeta)

    def get_zeta(self):
        return self.localCtModels.get_zeta()

    def reset_values(self):
        n = self.get_num()
        self.values = np.zeros(n)
        if self.globalRotDif.axisAvg is not None:
            self.errors = np.zeros(n)

    def update_values(self, values, errors=None, ind=None):
        if ind is None:
            self.values = values
            if errors is not None:
                self.errors = errors
        else:
            self.values[ind] = values
            if errors is not None:
                self.errors[ind] = errors

    def get_values(self, ind=None):
        return self.values if ind is None else self.values[ind]

    def get_errors(self, ind=None):
        if self.errors is None:
            return None
        return self.errors if ind is None else self.errors[ind]

    def calc_Jomega(self, ind=None):
        if ind is None:
            return self.globalRotDif.calc_Jomega(self.angFreq.omega, self.localCtModels)
        return self.globalRotDif.calc_Jomega_one(self.angFreq.omega, self.localCtModels.get_nth_model(ind), ind)

    def eval(self, ind=None, bVerbose=False):
        """(:831-853, :866-875, :894-907) value (and sigma over the vector distribution) for all residues,
        or for residue `ind` only (the rsCSA inner loop)."""
        w = self.angFreq
        allm = list(self.localCtModels.model.values())
        models = allm if ind is None else [allm[ind]]
        multi = isinstance(w.gA, gyromagMultiCSA)
        csa = np.atleast_1d(np.asarray(w.gA.get_csa(ind) if multi else w.gA.get_csa(), dtype=float))
        f_csa = (2.0 / 15.0 * csa ** 2.0 * (w.gA.gamma * w.B0) ** 2)[None, :]
        per_res = multi and ind is None
        out = _gpu_relax(self.globalRotDif, models, w.omega[None, :], f_csa, w.get_factor_DD(), w.gA.gamma, w.gB.gamma,
                         self.time_fact, csa_per_residue=per_res, subset=None if ind is None else [ind])
        v = out[:, 0, 0, self._col]
        e = out[:, 0, 0, 3 + self._col] if self.globalRotDif.axisAvg is not None else None
        if ind is not None:
            v, e = v[0], (None if e is None else e[0])
        self.update_values(v, e, ind=ind)
        return v

    def calc_chisq(self, Target, dTarget=None, indices=None):
        """(:803-818) mean squared deviation weighted by experimental and predicted variances."""
        v, e = self.values, self.errors
        if indices is not None:
            v = v[indices]
            if e is not None:
                e = e[indices]
        if e is not None and dTarget is not None:
            return np.mean(np.square(v - Target) / (np.square(dTarget) + np.square(e)))
        if e is None:
            return np.mean(np.square(v - Target) / np.square(dTarget))
        return np.mean(np.square(v - Target) / np.square(e))

    def get_suffix_from_conditions(self):
        return '_%s%s_%iMHz_%s' % (self.angFreq.gA.isotope, self.angFreq.gB.isotope,
                                   round(self.angFreq.get_magnetic_field(unit='MHz')), self.name)

    def print_metadata(self, style='stdout', fp=sys.stdout):
        if style == 'xmgrace':
            print('# Type %s' % self.name, file=fp)
            print('# NucleiA %s' % (self.angFreq.gA.isotope), file=fp)
            print('# NucleiB %s' % (self.angFreq.gB.isotope), file=fp)
            print('# Frequency %g %s' % (self.angFreq.get_magnetic_field(unit='MHz'), 'MHz'), file=fp)
        else:
            print("# %s Experiment at %sT over %i vectors" % (self.name, self.get_magnetic_field(), self.get_num()), file=fp)

    def print_values(self, style='stdout', fp=sys.stdout):
        names = self.localCtModels.get_names()
        if style == 'xmgrace':
            print("@type xy" if self.errors is None else "@type xydy", file=fp)
        if self.errors is None:
            for x, y in zip(names, self.values):
                print("%s %g" % (x, y), file=fp)
        else:
            for x, y, dy in zip(names, self.values, self.errors):
                print("%s %g %g" % (x, y, dy), file=fp)
        print("&" if style == 'xmgrace' else '', file=fp)


class spinRelaxationR1(spinRelaxationBase):
    _col = 0


class spinRelaxationR2(spinRelaxationBase):
    _col = 1


class spinRelaxationNOE(spinRelaxationBase):
    _col = 2


def _BAIL(functionName, message):
    print("= = ERROR in function %s : %s" % (functionName, message), file=sys.stderr)
    sys.exit(1)


class spinRelaxationExperiments:
    """Container for several experiments sharing one tumbling model and one set of C(t) models
    (spectral_densities.py:909-1420): add_experiment (:935-1010), the peak-name maps (:1050-1097), eval_all (:1145-1150),
    initialise_CSA_array, export_xvg (:1178-1194) and the optimisation against experiment over Diso / Daniso / zeta /
    CSA / rsCSA (:1196-1447).  Global steps are SciPy's Powell on the host exactly as in the reference, every function
    evaluation being one batched GPU evaluation of all residues; the residue-specific CSA step is solved for all
    residues at once (`local_mode='batched'`: vectorised bracketing + golden section, one GPU launch per experiment and
    iteration) or, for step-by-step parity with the reference, residue by residue with Powell (`local_mode='powell'`)."""

    listAllowedOptimisationVariables = ['Diso', 'Daniso', 'CSA', 'zeta', 'rsCSA']
    dictStepSizes = {'Diso': 1e-5, 'Daniso': 0.1, 'zeta': 0.1, 'CSA': 1e-5, 'rsCSA': 1e-5}
    dictExportScaling = {'Diso': 1.0, 'Daniso': 1.0, 'zeta': 1.0, 'CSA': 1e6, 'rsCSA': 1e6}
    dictExportUnits = {'Diso': 'ps^-1', 'Daniso': 'a.u.', 'zeta': 'a.u.', 'CSA': 'ppm', 'rsCSA': 'ppm'}

    def __init__(self, globalRotDif=None, localCtModels=None, local_mode='batched'):
        self.numExpts = 0
        self.spinrelax = []
        self.data = []
        self.globalRotDif = globalRotDif
        self.localCtModels = localCtModels
        self.mapModelNames = []
        self.mapExptCoverage = []
        self.bOptInitialised = False
        self.bOptCompleted = False
        self.bDoLocalOpt = False
        self.listUpdateVariables = []
        self.chisq = None
        self.local_mode = local_mode
        self.bVerboseOpt = True
        self.bAliasQuirk = True

    def add_experiment(self, fileName, bIgnoreErrors=False):
        strType = nucleiA = nucleiB = freq = None
        freqUnit = 'MHz'
        names, values, errors = [], [], []
        for line in open(fileName, 'r'):
            l = line.split()
            if len(l) == 0:
                continue
            if line[0] in '#@':
                if len(l) > 2:
                    key = l[1]
                    if key == 'Type':
                        strType = l[2]
                    elif key == 'NucleiA':
                        nucleiA = l[2]
                    elif key == 'NucleiB':
                        nucleiB = l[2]
                    elif key == 'Frequency':
                        freq = float(l[2])
                    elif key == 'FrequencyUnit':
                        freqUnit = l[2]
                continue
            if len(l) == 1 or len(l) > 3:
                print("ERROR in spinRelaxationExperiments.add_experiment(): data line does not obey expected "
                      "conventions of 2 or 3 space-separated values!", l, file=sys.stderr)
                sys.exit()
            names.append(l[0])
            values.append(float(l[1]))
            errors.append(float(l[2]) if len(l) > 2 else None)
        if nucleiB is None and strType in ('R1', 'R2'):
            nucleiB = '1H'
        if strType is None or nucleiA is None or nucleiB is None or freq is None:
            print("ERROR in spinRelaxationExperiments.add_experiment(): not all metadata has been read! "
                  "Require: Type, NucleiA, NucleiB, Frequency", file=sys.stderr)
            sys.exit(1)
        nMissing = sum(x is None for x in errors)
        if nMissing == len(errors):
            errors = None
        elif nMissing > 0:
            print("ERROR in spinRelaxationExperiments.add_experiment(): either all entries must have uncertainties "
                  "or none!", file=sys.stderr)
            sys.exit(1)
        wObj = angularFrequencies(nucleiA=nucleiA, nucleiB=nucleiB, fieldStrength=freq, fieldUnit=freqUnit)
        cls = {'R1': spinRelaxationR1, 'R2': spinRelaxationR2, 'NOE': spinRelaxationNOE}[strType]
        self.spinrelax.append(cls(strType, angFreq=wObj, globalRotDif=self.globalRotDif, localCtModels=self.localCtModels))
        self.data.append(dict(names=np.array(names), y=np.array(values, dtype=float), dy=errors))
        self.numExpts += 1

    # ---- peak-name maps (:1050-1097) -------------------------------------------------------------------
    def map_experiment_peaknames_to_models(self):
        if self.localCtModels is None:
            print("ERROR in spinRelaxationExperiments.map_peak_names: need a local C(t) model to map experimental "
                  "peak names!", file=sys.stderr)
            sys.exit(1)
        namesCt = np.array([str(x) for x in self.localCtModels.get_names()])
        if self.globalRotDif is not None and getattr(self.globalRotDif, 'bVecs', False):
            namesRotdif = [str(x) for x in self.globalRotDif.get_names()]
            if list(namesCt) != namesRotdif:
                print("ERROR in spinRelaxationExperiments.map_peak_names: local C(t) model and global "
                      "rotational-diffusion model do not have matching peak names!", file=sys.stderr)
                sys.exit(1)
        print("    ....mapping peak sets (residue IDs) between simulated localCtModels and expeirmental datasets")
        self.mapModelNames = []
        for i in range(self.numExpts):        # gm.list_get_map(namesCt, data names): model index of every peak found
            tmp = []
            for x in self.data[i]['names']:
                t = np.where(namesCt == x)[0]
                if len(t) > 0:
                    tmp.append(t[0])
            self.mapModelNames.append(tmp)
        self.mapExptCoverage = []
        for i in range(self.localCtModels.nModels):
            l = []
            for exptID in range(self.numExpts):
                ret = np.where(self.data[exptID]['names'] == namesCt[i])[0]
                if len(ret) > 0:
                    l.append((exptID, ret[0]))
            self.mapExptCoverage.append(l)

    def report_maps(self):
        print("Number of simulation residues covered by each experiment:",
              ''.join(' %i' % len(x) for x in self.mapModelNames))
        print("Number of Experiments covering each simulation residue:",
              ''.join(' %i' % len(x) for x in self.mapExptCoverage))

    # ---- parameter access (:1196-1210, :1244-1266) ----------------------------------------------------
    def set_global_Diso(self, Diso):
        self.globalRotDif.set_Diso(Diso)

    def get_global_Diso(self):
        return self.globalRotDif.get_Diso()

    def set_global_Daniso(self, Daniso):
        self.globalRotDif.set_Daniso(Daniso)

    def get_global_Daniso(self):
        return self.globalRotDif.get_Daniso()

    def set_global_zeta(self, zeta):
        self.localCtModels.set_zeta(zeta)

    def get_global_zeta(self):
        return self.localCtModels.get_zeta()

    def get_zeta(self):
        return self.localCtModels.get_zeta()

    def set_all_csa(self, csa, ind=None):
        for sp in self.spinrelax:
            sp.angFreq.gA.set_csa(csa, ind)

    def get_first_csa(self, ind=None):
        return self.spinrelax[0].angFreq.gA.get_csa(ind)

    def return_get_function(self, param):
        return {'Diso': self.get_global_Diso, 'Daniso': self.get_global_Daniso, 'zeta': self.get_global_zeta,
                'CSA': self.get_first_csa}.get(param)

    def return_set_function(self, param):
        return {'Diso': self.set_global_Diso, 'Daniso': self.set_global_Daniso, 'zeta': self.set_global_zeta,
                'CSA': self.set_all_csa}.get(param)

    def initialise_CSA_array(self, namesCSA, CSAValues):
        namesCSA = [str(x) for x in namesCSA]
        namesCt = [str(x) for x in self.localCtModels.get_names()]
        default = self.get_first_csa()
        vals = [CSAValues[namesCSA.index(n)] if n in namesCSA else default for n in namesCt]
        for sp in self.spinrelax:
            sp.angFreq.initialise_CSA_array(len(vals), vals)

    def eval_all(self, ind=None, bVerbose=False):
        for sp in self.spinrelax:
            if bVerbose:
                print('...evaluating experiment %s at %g T.' % (sp.name, sp.angFreq.B0))
            sp.eval(ind=ind, bVerbose=bVerbose)

    def get_all_values(self, ind=None):
        out = []
        for sp in self.spinrelax:
            v, e = sp.get_values(ind=ind), sp.get_errors(ind=ind)
            out.append([v, e] if e is not None else [v])
        return out

    # ---- optimisation against experiment (:1268-1447) --------------------------------------------------
    def parse_optimisation_params(self, listOpts):
        self.bOptInitialised = False
        self.bDoLocalOpt = False
        self.listUpdateVariables, self.listStepSizes, self.listSetFunctions, self.listGetFunctions = [], [], [], []
        if 'CSA' in listOpts and 'rsCSA' in listOpts:
            _BAIL("parse_optimisation_params", "Cannot run both global CSA as well as residue-specific CSA optimisation!")
        for o in listOpts:
            if o not in spinRelaxationExperiments.listAllowedOptimisationVariables:
                _BAIL("parse_optimisation_params", "Optimisation variable %s not found in list!\nPossibilities are: %s "
                      % (o, spinRelaxationExperiments.listAllowedOptimisationVariables))
            if o == 'rsCSA':
                csa = self.get_first_csa()
                if not type(csa) is np.ndarray:
                    print("    ... NOTE: CSA values have not been preset but residue-specific CSA optimisation is being "
                          "performed. Reinitialising CSA variables as being residue-specific.")
                    self.initialise_CSA_array(self.localCtModels.get_names(), np.repeat(csa, self.localCtModels.nModels))
                self.bDoLocalOpt = True          # does not trigger global optimisation procedures
                continue
            self.listGetFunctions.append(self.return_get_function(o))
            self.listSetFunctions.append(self.return_set_function(o))
            self.listStepSizes.append(spinRelaxationExperiments.dictStepSizes[o])
            self.listUpdateVariables.append(o)
        self.bOptInitialised = True

    def perform_optimisation(self, maxCycles=10, tol=1e-6):
        """(:1302-1358) global Powell rounds, residue-specific CSA rounds, or alternating cycles of both."""
        if not self.bOptInitialised:
            _BAIL("perform_fit", "You must first run parse_optimisation_params to tell the script what to optimise.")
        if len(self.mapExptCoverage) == 0:
            self.map_experiment_peaknames_to_models()
        bDoGlobalOpt = len(self.optimisation_loop_get_globals()) > 0
        if bDoGlobalOpt and not self.bDoLocalOpt:
            self.optimisation_loop_do_global_step()
            self.bOptCompleted = True
            return self.chisq
        if self.bDoLocalOpt and not type(self.get_first_csa()) is np.ndarray:
            _BAIL("perform_optimisation", "CSA values are not an array for local optimisation!")
        if self.bDoLocalOpt and not bDoGlobalOpt:
            self.eval_all()
            self.optimisation_loop_do_local_step()
            self.bOptCompleted = True
            self.chisq = self.calc_chisq()
            return self.chisq
        if bDoGlobalOpt and self.bDoLocalOpt:
            bFirst = True
            for n in range(maxCycles):
                paramPrev = self.optimisation_loop_get_globals()
                self.optimisation_loop_do_global_step()
                paramNow = self.optimisation_loop_get_globals()
                if not bFirst and np.allclose(paramPrev, paramNow, rtol=tol):
                    self.bOptCompleted = True
                    break
                csaPrev = np.array(self.get_first_csa())
                self.optimisation_loop_do_local_step()
                csaNow = self.get_first_csa()
                # The reference's csaPrev is an alias of the array its local step updates in place (:1341-1344), so
                # its convergence test always passes from the second cycle on; bAliasQuirk reproduces that.
                if not bFirst and (self.bAliasQuirk or np.allclose(csaPrev, csaNow, rtol=tol)):
                    self.chisq = self.calc_chisq()
                    self.bOptCompleted = True
                    break
                bFirst = False
            return self.chisq
        _BAIL("perform_optimisation", "neither global or local optimisation have been successfully specified!")

    def optimisation_loop_do_global_step(self):
        from scipy.optimize import fmin_powell
        fminOut = fmin_powell(optimisation_loop_inner_function, x0=self.optimisation_loop_get_globals(),
                              direc=self.optimisation_loop_get_direc(), args=(self, '1'), full_output=True,
                              disp=self.bVerboseOpt)
        if self.bVerboseOpt:
            print("= = = Optimisation complete over variables: %s" % self.optimisation_loop_get_param_names())
            print(fminOut)
        self.chisq = fminOut[1]

    def optimisation_loop_do_local_step(self):
        """Residue-specific CSA (:1371-1384)."""
        if self.local_mode == 'powell':
            from scipy.optimize import fmin_powell
            for i in range(self.localCtModels.nModels):
                if len(self.mapExptCoverage[i]) > 0:
                    fmin_powell(optimisation_loop_rsCSA_inner_function, x0=self.get_first_csa(ind=i),
                                direc=[spinRelaxationExperiments.dictStepSizes['rsCSA']],
                                args=(self, i, self.mapExptCoverage[i]), full_output=False, disp=self.bVerboseOpt)
            return
        self._local_step_batched()

    def _rsCSA_objective(self, csa):
        """optimisation_loop_rsCSA_inner_function (:1430-1447) for ALL residues at once: chi^2_i of residue i at csa_i."""
        self.set_all_csa(np.array(csa))
        self.eval_all()
        n = self.localCtModels.nModels
        chi, cnt = np.zeros(n), np.zeros(n)
        for i, cover in enumerate(self.mapExptCoverage):
            for exptID, peakID in cover:
                sp, target = self.spinrelax[exptID], self.data[exptID]
                dv = sp.errors[i] if sp.errors is not None else 0.0
                dt = target['dy'][peakID] if target['dy'] is not None else 0.0
                w = dv ** 2 + dt ** 2
                chi[i] += (sp.values[i] - target['y'][peakID]) ** 2 / (w if w != 0 else 1.0)
                cnt[i] += 1
        return chi / np.maximum(cnt, 1)

    def _local_step_batched(self, nGolden=48):
        """All residue-specific 1-D minimisations at once: downhill bracketing from the current CSA with the
        reference's step size, then golden-section refinement (bracket shrinks by 0.618^48 ~ 1e-10)."""
        x0 = np.array(self.get_first_csa(), dtype=float)
        active = np.array([len(c) > 0 for c in self.mapExptCoverage])
        before = [(sp.values.copy(), None if sp.errors is None else sp.errors.copy()) for sp in self.spinrelax]
        gold = 1.618033988749895
        step = spinRelaxationExperiments.dictStepSizes['rsCSA']
        xa, xb = x0.copy(), x0 + step
        fa, fb = self._rsCSA_objective(xa), self._rsCSA_objective(xb)
        swap = fb > fa
        xa[swap], xb[swap] = xb[swap], xa[swap].copy()
        fa[swap], fb[swap] = fb[swap], fa[swap].copy()
        xc = xb + gold * (xb - xa)
        fc = self._rsCSA_objective(xc)
        for _ in range(40):
            grow = active & (fc < fb)
            if not grow.any():
                break
            xa[grow], fa[grow] = xb[grow], fb[grow]
            xb[grow], fb[grow] = xc[grow], fc[grow]
            xc = np.where(grow, xb + gold * (xb - xa), xc)
            fnew = self._rsCSA_objective(np.where(grow, xc, xb))
            fc = np.where(grow, fnew, fc)
        lo, hi = np.minimum(xa, xc), np.maximum(xa, xc)
        r = 1.0 / gold
        x1, x2 = hi - r * (hi - lo), lo + r * (hi - lo)
        f1, f2 = self._rsCSA_objective(x1), self._rsCSA_objective(x2)
        for _ in range(nGolden):
            left = f1 < f2
            hi = np.where(left, x2, hi)
            lo = np.where(left, lo, x1)
            x2n = np.where(left, x1, lo + r * (hi - lo))
            x1n = np.where(left, hi - r * (hi - lo), x2)
            fnew = self._rsCSA_objective(np.where(left, x1n, x2n))
            f2n = np.where(left, f1, fnew)
            f1n = np.where(left, fnew, f2)
            x1, x2, f1, f2 = x1n, x2n, f1n, f2n
        best = np.where(f1 < f2, x1, x2)
        self.set_all_csa(np.where(active, best, x0))
        self.eval_all()
        # The reference's residue loop only re-evaluates the (experiment, residue) pairs that have a measured peak
        # (:1375-1384); predictions of unresolved peaks keep the value they had before the local step.
        covered = [set() for _ in self.spinrelax]
        for i, cover in enumerate(self.mapExptCoverage):
            for exptID, _ in cover:
                covered[exptID].add(i)
        for e, sp in enumerate(self.spinrelax):
            stale = [i for i in range(self.localCtModels.nModels) if i not in covered[e]]
            if stale:
                sp.values[stale] = before[e][0][stale]
                if sp.errors is not None and before[e][1] is not None:
                    sp.errors[stale] = before[e][1][stale]

    def optimisation_loop_get_param_names(self):
        return self.listUpdateVariables

    def optimisation_loop_get_direc(self):
        n = len(self.listStepSizes)
        out = np.zeros((n, n))
        for i in range(n):
            out[i, i] = self.listStepSizes[i]
        return out

    def optimisation_loop_get_globals(self):
        return [func() for func in self.listGetFunctions]

    def optimisation_loop_set_globals(self, vNew):
        for func, v in zip(self.listSetFunctions, vNew):
            func(v)

    def calc_chisq(self):
        chisq = 0.0
        for i, sp in enumerate(self.spinrelax):
            chisq += sp.calc_chisq(self.data[i]['y'], self.data[i]['dy'], self.mapModelNames[i])
        return chisq / self.numExpts

    def print_parameters(self, style='stdout', fp=sys.stdout):
        """`# Fixed|Optimised <name>: <value> <unit>` lines of the .xvg header (:1224-1242)."""
        for x in spinRelaxationExperiments.listAllowedOptimisationVariables:
            if x == 'rsCSA':
                continue
            v = self.return_get_function(x)()
            s1 = 'Optimised' if x in self.listUpdateVariables else 'Fixed'
            if x == 'CSA' and type(v) is np.ndarray:
                v = np.mean(v)
                s1 = 'OptimisedMean' if (self.bOptCompleted and self.bDoLocalOpt) else 'FixedMean'
            print('# %s %s: %g %s' % (s1, x, v * spinRelaxationExperiments.dictExportScaling[x],
                                      spinRelaxationExperiments.dictExportUnits[x]), file=fp)
        if self.bOptCompleted:
            print('# Optimised chi: %g a.u.' % np.sqrt(self.chisq), file=fp)

    def export_xvg(self, filePrefix, bIncludeExpt=False):
        for i, sp in enumerate(self.spinrelax):
            with open('%s%s.xvg' % (filePrefix, sp.get_suffix_from_conditions()), 'w') as fp:
                sp.print_metadata('xmgrace', fp)
                self.print_parameters('xmgrace', fp)
                print('', file=fp)
                print('@target s0', file=fp)
                sp.print_values('xmgrace', fp)
                if bIncludeExpt:
                    print('@target s1', file=fp)
                    d = self.data[i]
                    print('@type xy' if d['dy'] is None else '@type xydy', file=fp)
                    for k in range(len(d['y'])):
                        print(("%s %g" % (d['names'][k], d['y'][k])) if d['dy'] is None
                              else ("%s %g %g" % (d['names'][k], d['y'][k], d['dy'][k])), file=fp)
                    print('&', file=fp)


def optimisation_loop_inner_function(params, *args):
    """(:1422-1428) objective of the global Powell step: set the globals, re-evaluate everything, chi^2."""
    objExpts = args[0]
    objExpts.optimisation_loop_set_globals(params)
    objExpts.eval_all(bVerbose=False)
    chisq = objExpts.calc_chisq()
    if objExpts.bVerboseOpt:
        print("    ....optimisation step. Params: %s chisq: %g" % (params, chisq))
    return chisq


def optimisation_loop_rsCSA_inner_function(params, *args):
    """(:1430-1447) objective of one residue's CSA step."""
    objExpts, ind, listCalcs = args[0], args[1], args[2]
    objExpts.set_all_csa(params[0], ind=ind)
    chisq = 0.0
    for exptID, peakID in listCalcs:
        sp, target = objExpts.spinrelax[exptID], objExpts.data[exptID]
        v = sp.eval(ind=ind)
        t = target['y'][peakID]
        dv = sp.errors[ind] if sp.errors is not None else 0.0
        dt = target['dy'][peakID] if target['dy'] is not None else 0.0
        w = dv ** 2 + dt ** 2
        if w == 0:
            w = 1
        chisq += (v - t) ** 2 / w
    return chisq / len(listCalcs)
