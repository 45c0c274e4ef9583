"""Host mirror of calculate-dq-distribution.py: global-tumbling statistics from an orientation-quaternion
trajectory.  The per-lag reductions (the hot loop, :554-625) run on the GPU (sr_dq_moments); what is
O(#lags) -- eigen-frames, SciPy Powell fits (:199-207), text writers -- stays on the host exactly as the
reference does it so that the fitted D tensors follow from identical curves.

Function names/signatures follow the reference: obtain_self_dq, average_LegendreP1quat,
average_anisotropic_tensor, *_chunk, conduct_exponential_fit, calculate_anisotropies, format_header,
print_model_fits_gen; `main(argv)` reproduces the CLI (flags :426-458, outputs -iso.dat, -aniso2.dat,
-aniso_q.dat, -moi.xyz, -tensor.dat).
"""
import argparse
import math
import sys
import time

import numpy as np
from scipy.optimize import fmin_powell

from . import _lib, io_formats, qs

IDENTITY = (1.0, 0.0, 0.0, 0.0)


def _sym3(m6):
    return np.array([[m6[0], m6[1], m6[2]], [m6[1], m6[3], m6[4]], [m6[2], m6[4], m6[5]]])


# ---- GPU reductions ---------------------------------------------------------------------------------
def dq_moment_sums(q, lags, nchunk=1):
    """Raw second-moment sums of the dq vector parts: returns (M (nLags, nCh, 6) float64, n (nLags,),
    counts (nLags, nCh)).  q: (N, 4) float32 (w,x,y,z) -- or (nReplicas, N, 4) / a list of equally long
    trajectories whose displacement samples are pooled the way calculate-dq-distribution-multi.py:529-540 does
    (sub-chunks are then blocks of the pooled sample list); lags: frame lags."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    q32 = np.ascontiguousarray(q, dtype=np.float32)
    if q32.ndim == 2:
        q32 = q32[None]
    if q32.ndim != 3 or q32.shape[2] != 4:
        raise ValueError("dq_moment_sums: q must be (N, 4) or (nReplicas, N, 4)")
    lags = np.ascontiguousarray(lags, dtype=np.int64)
    nRep, N = q32.shape[:2]
    if lags.size == 0 or lags.min() < 1 or lags.max() >= N:
        raise ValueError("dq_moment_sums: lags must lie in [1, N)")
    nCh = max(1, int(nchunk))
    from . import multigpu

    def work(dev, a, b):        # every device holds the (small) trajectory and reduces a block of the lag list
        sub = np.ascontiguousarray(lags[a:b])
        qd = torch.from_numpy(q32).cuda()
        ld = torch.from_numpy(sub).cuda()
        M = torch.empty((sub.size, nCh, 6), dtype=torch.float64, device=qd.device)
        for r in range(nRep):
            _lib.check(lib.sr_dq_moments_pooled(qd[r].data_ptr(), N, ld.data_ptr(), sub.size, int(sub.min()), nCh, r, nRep,
                                                1 if r > 0 else 0, M.data_ptr(), _lib.current_stream_ptr()),
                       "sr_dq_moments_pooled")
        return M.cpu().numpy()

    M = np.concatenate(multigpu.run(multigpu.plan(lags.size, min_per_device=64), work), axis=0)
    n = (N - lags) * nRep
    nb = -(-n // nCh)
    k = np.arange(nCh)[None, :]
    counts = np.clip(np.minimum(n[:, None], nb[:, None] * (k + 1)) - nb[:, None] * k, 0, None)
    return M, n, counts


def dq_histogram3d(q, delta, nbins=101):
    """np.histogramdd(v_dq, range=[(-1,1)]*3, bins=(nbins,)*3, density=True) of the --hist option
    (calculate-dq-distribution.py:527-528, 633-634; `normed=True` there is the pre-1.24 spelling of density).
    Returns (hist (nbins,)*3 float64, edges [3 x (nbins+1,)])."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    q32 = np.ascontiguousarray(q, dtype=np.float32)
    N = q32.shape[0]
    edges = np.linspace(-1.0, 1.0, nbins + 1)
    qd = torch.from_numpy(q32).cuda()
    ed = torch.from_numpy(edges).cuda()
    counts = torch.zeros((nbins, nbins, nbins), dtype=torch.int32, device=qd.device)
    cap = 1 << 16
    amb = torch.empty(cap, dtype=torch.int64, device=qd.device)
    namb = torch.zeros(1, dtype=torch.int32, device=qd.device)
    _lib.check(lib.sr_dq_hist3d(qd.data_ptr(), N, int(delta), ed.data_ptr(), nbins, counts.data_ptr(), amb.data_ptr(), cap,
                                namb.data_ptr(), _lib.current_stream_ptr()), "sr_dq_hist3d")
    c = counts.cpu().numpy().astype(np.float64)
    k = int(namb.item())
    if k > cap:
        raise _lib.SpinRelaxError("dq_histogram3d: %d edge-ambiguous samples exceed the list capacity" % k)
    if k:   # samples within 4 ulp of a bin edge (or NaN): NumPy's own arithmetic decides
        t = amb[:k].cpu().numpy()
        v = obtain_self_dq(q32, delta)[t, 1:4]
        extra, _ = np.histogramdd(v, range=[(-1, 1)] * 3, bins=(nbins,) * 3)
        c += extra
    s = c.sum()
    hist = c / (s * np.prod([np.diff(edges)[0]] * 3)) if s > 0 else c
    return hist, [edges, edges.copy(), edges.copy()]


def _vec_moment_sums(vq, nchunk=1):
    torch = _lib.require_cuda()
    lib = _lib.load()
    v = np.ascontiguousarray(vq, dtype=np.float64)
    vd = torch.from_numpy(v).cuda()
    M = torch.empty((max(1, nchunk), 6), dtype=torch.float64, device=vd.device)
    _lib.check(lib.sr_vec_second_moments(vd.data_ptr(), v.shape[0], max(1, nchunk), M.data_ptr(),
                                         _lib.current_stream_ptr()), "sr_vec_second_moments")
    return M.cpu().numpy()


def obtain_self_dq(q, delta):
    """{ q^-1(t) q(t+delta) } imaged to w >= 0 (calculate-dq-distribution.py:102-109): (N-delta, 4) float64."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    q32 = np.ascontiguousarray(q, dtype=np.float32)
    N = q32.shape[0]
    qd = torch.from_numpy(q32).cuda()
    out = torch.empty((N - delta, 4), dtype=torch.float64, device=qd.device)
    _lib.check(lib.sr_dq_self(qd.data_ptr(), N, int(delta), out.data_ptr(), _lib.current_stream_ptr()), "sr_dq_self")
    return out.cpu().numpy()


def _iso_shipped(sum6):
    """1 - (2/3) sum_t |v|^2: what average_LegendreP1quat returns as shipped (:111-112, quirk G1)."""
    return 1.0 - (2.0 / 3.0) * (sum6[..., 0] + sum6[..., 3] + sum6[..., 5])


def iso_intended(sum6, n):
    """<1 - 2|v|^2>, the quantity the reference's comments describe (:113-116)."""
    return 1.0 - 2.0 * (sum6[..., 0] + sum6[..., 3] + sum6[..., 5]) / n


def average_LegendreP1quat(ndat, vq):
    return float(_iso_shipped(_vec_moment_sums(vq, 1)[0]))


def average_LegendreP1quat_chunk(ndat, vq, nchunk):
    return _iso_shipped(_vec_moment_sums(vq, nchunk))


def _sym3_batch(m6):
    """(..., 6) packed xx xy xz yy yz zz -> (..., 3, 3)."""
    m6 = np.asarray(m6)
    idx = np.array([[0, 1, 2], [1, 3, 4], [2, 4, 5]])
    return m6[..., idx]


def _rotated_batch(M, qframe):
    """_rotated for a stack of tensors (..., 3, 3) and one frame quaternion."""
    if qs.nearly_identity(qframe):
        return M
    R = qs.rotation_matrix(qframe)
    # vec(R M R^T) = (R kron R) vec(M), row-major: one (n, 9) x (9, 9) product instead of 2n tiny 3x3 ones
    M = np.asarray(M, dtype=float)
    return (M.reshape(-1, 9) @ np.kron(R, R).T).reshape(M.shape)


def _rotated(M33, qframe):
    if qs.nearly_identity(qframe):
        return M33
    R = qs.rotation_matrix(qframe)
    return R @ M33 @ R.T


def average_anisotropic_tensor(ndat, vq, qframe=IDENTITY):
    """mean_t (R v)(R v)^T (:118-126) = R <v v^T> R^T."""
    return _rotated(_sym3(_vec_moment_sums(vq, 1)[0]) / len(vq), qframe)


def average_anisotropic_tensor_chunk(ndat, vq, nchunk, qframe=IDENTITY):
    n = len(vq)
    nb = int(math.ceil(1.0 * n / nchunk))
    M = _vec_moment_sums(vq, nchunk)
    out = np.zeros((nchunk, 3, 3))
    for i in range(nchunk):
        cnt = min(n, nb * (i + 1)) - nb * i
        out[i] = _rotated(_sym3(M[i]) / cnt, qframe)
    return out


# ---- lag bookkeeping + curves (main loop :510-650) ------------------------------------------------------
def lag_grid(times, min_dt, max_dt, skip_dt):
    ddt = times[1] - times[0]
    skip_int = max(1, int(skip_dt / ddt))
    min_int = max(skip_int, int(min_dt / ddt))
    max_int = int(max_dt / ddt)
    return min_int, max_int, skip_int, ddt


def dq_curves(q, lags, ddt, nchunk=0, do_aniso=True):
    """All per-lag outputs of the reference's main loop from one GPU pass over the lag list (q may hold several
    replica trajectories, see dq_moment_sums)."""
    lags = np.asarray(lags, dtype=np.int64)
    nl = len(lags)
    nCh = nchunk if nchunk > 1 else 1
    M, n, counts = dq_moment_sums(q, lags, nCh)
    full = M.sum(axis=1)
    out = dict(dt=lags * ddt, iso=_iso_shipped(full), iso_intended=iso_intended(full, n),
               aniso1=np.zeros((3, nl)), aniso2=np.zeros((3, nl)), qrot=np.zeros((4, nl)),
               moi_axes=np.zeros((nl, 3, 3)), moi=np.zeros((nl, 3, 3)), moiR=np.zeros((nl, 3, 3)))
    if nchunk > 1:
        out["chunk_iso"] = _iso_shipped(M).T.copy()
        out["chunk_aniso2"] = np.zeros((nchunk, 3, nl))
    # everything below is O(#lags) host arithmetic of the reference's loop body (:566-625), evaluated for the whole lag
    # list at once: stacked 3x3 eigen-decompositions, frame quaternions and rotations
    q_frame = IDENTITY
    moi_all = _sym3_batch(full) / np.asarray(n, dtype=float)[:, None, None]
    out["moi"][:] = moi_all
    if do_aniso and nl:
        eigval, eigvec = np.linalg.eigh(moi_all)
        axes = np.swapaxes(eigvec, 1, 2)
        q_rot = qs.quat_frame_transform_min_batch(axes)
        q_frame = q_rot[0]                         # the frame of the first window is kept for all windows (:569-572)
        moiR = _rotated_batch(moi_all, q_frame)
        out["aniso1"][:] = (1 - 2 * eigval).T
        out["aniso2"][:] = (1 - 2 * np.diagonal(moiR, axis1=1, axis2=2)).T
        out["qrot"][:] = q_rot.T
        out["moi_axes"][:] = axes
        out["moiR"][:] = moiR
    else:
        out["moiR"][:] = moi_all
    if nchunk > 1:
        with np.errstate(invalid="ignore", divide="ignore"):
            t2 = _rotated_batch(_sym3_batch(M) / counts[:, :, None, None], q_frame)          # (nl, nCh, 3, 3)
        out["chunk_aniso2"][:] = np.transpose(1 - 2 * np.diagonal(t2, axis1=2, axis2=3), (1, 2, 0))
    out["q_frame"] = np.array(q_frame, dtype=float)
    return out


# ---- fits and derived quantities (host, SciPy as in the reference) ------------------------------------
def powell_expdecay(pos, *args):
    """Mean squared deviation of C0 exp(-x/A) + C1 from y (calculate-dq-distribution.py:199-203 sums it point by point
    in a Python loop).  Evaluated as array operations with the terms added left to right (np.add.accumulate is the
    sequential sum, not NumPy's pairwise one), so the value is the loop's up to the last-bit difference between
    libm's and NumPy's exp -- with 1e5 lag windows and ~100 objective calls per curve the loop would dominate the
    whole stage."""
    x, y, C0, C1 = args
    A = float(np.ravel(pos)[0])
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    if x.size == 0:
        return 0.0 / len(x)
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        term = np.square(C0 * np.exp(-x / A) + C1 - y)
    return float(np.add.accumulate(term)[-1]) / len(x)


def obtain_exponential_guess(x, y, C1):
    return (x[0] - x[1]) / math.log((y[1] - C1) / (y[0] - C1))


DEVICE_OBJECTIVE_MIN_POINTS = 16384


class _DeviceObjective:
    """powell_expdecay for a long curve kept on the GPU (sr_expdecay_chi2): one small launch and an 8-byte read-back per
    trial tau instead of a pass over 1e5 points on the host.  Same signature as powell_expdecay."""

    def __init__(self, x, y):
        torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.n = len(x)
        self.x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        self.y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).cuda()
        self.work = torch.zeros(136, dtype=torch.float64, device=self.x.device)
        self.out = torch.zeros(1, dtype=torch.float64, device=self.x.device)

    def __call__(self, pos, *args):
        _, _, C0, C1 = args
        A = float(np.ravel(pos)[0])
        _lib.check(self.lib.sr_expdecay_chi2(self.x.data_ptr(), self.y.data_ptr(), self.n, float(C0), float(C1), A,
                                             self.work.data_ptr(), self.work.numel(), self.out.data_ptr(),
                                             _lib.current_stream_ptr()), "sr_expdecay_chi2")
        return float(self.out.item())


def conduct_exponential_fit(xlist, ylist, C0, C1):
    """1-parameter Powell fit of C0 exp(-x/tau) + C1 (:199-207).  SciPy's Powell runs on the host as in the reference;
    for long curves (all lag windows: 1e5 points) its objective is evaluated on the device."""
    print('= = Begin exponential fit.')
    guess = obtain_exponential_guess([xlist[0], xlist[1]], [ylist[0], ylist[1]], C1)
    print('= = = guessed initial tau: ', guess)
    objective = _DeviceObjective(xlist, ylist) if len(xlist) >= DEVICE_OBJECTIVE_MIN_POINTS else powell_expdecay
    fitOut = fmin_powell(objective, guess, args=(xlist, ylist, C0, C1), full_output=True)
    tau = np.ravel(fitOut[0])[0]
    print('= = = = Tau obtained: ', tau)
    return tau


def isotropic_decay(x, a):
    return 1.5 * np.exp(-x / a) - 0.5


def anisotropic_decay_noc(x, a):
    return 0.5 * np.exp(-x / a) + 0.5


def _aniso_tuple(D):
    ani = lambda d: 2 * d[2] / (d[1] + d[0])                      # noqa: E731
    rho = lambda d: 3 * (d[1] - d[0]) / (2 * d[2] - d[1] - d[0])  # noqa: E731
    return (np.mean(D), ani(D), rho(D), ani(D[::-1]), rho(D[::-1]))


def calculate_anisotropies(D, chunkD=[]):
    """(:70-91) anisotropy/rhombicity of the sorted D; with chunkD also the std over chunks."""
    D = np.asarray(D)
    if len(chunkD) == 0:
        return _aniso_tuple(np.sort(D))
    order = np.argsort(D)
    val = _aniso_tuple(D[order])
    errs = np.std(np.array([_aniso_tuple(np.asarray(x)[order]) for x in chunkD]), axis=0)
    return [(val[i], errs[i]) for i in range(len(val))]


def get_flex_bounds(x, samples, nsig=1):
    mean, sig = np.mean(samples), np.std(samples)
    return [x, nsig * sig + x - mean, nsig * sig + mean - x]


def format_header(style_str, tau, taus=[]):
    """Header lines of -iso.dat / -aniso2.dat (:221-272); run-all.bash:412-416 parses them by field position."""
    L = []
    if style_str == 'iso':
        L += ['# model fit, tau = %e [ps]' % tau, "# Converted D_iso = %e [s^-1]" % (0.5e12 / tau),
              "# t cos(th) P2[cos(th)] cos(th/2) th"]
    elif style_str == 'iso_err':
        b = get_flex_bounds(tau, taus)
        L.append('# model fit, tau = %e +- %e %e [ps]' % (b[0], b[1], b[2]))
        Dvals = [0.5e12 / t for t in taus]
        b = get_flex_bounds(0.5e12 / tau, Dvals)
        L.append('# Converted D_iso = %e +- %e %e [s^-1]' % (b[0], b[1], b[2]))
        L += ['# Chunk_%d D_iso = %e [s^-1]' % (i, Dvals[i]) for i in range(len(taus))]
        L.append("# t cos(th) P2[cos(th)] cos(th/2) th")
    elif style_str == 'aniso':
        Dval = 0.5e12 / tau
        for i in range(3):
            L.append("# model fit, e_%i tau = %e [ps]" % (i, tau[i]))
            L.append("# Converted D_%i = %e [s^-1]" % (i, Dval[i]))
        a = calculate_anisotropies(Dval)
        L += ["# Converted Diso = %e [s^-1]" % a[0], "# Converted Dani_L = %f" % a[1], "# Converted Drho_L = %f" % a[2],
              "# Converted Dani_S = %f" % a[3], "# Converted Drho_S = %f" % a[4], "# t <1-2x^2> <1-2y^2> <1-2z^2>"]
    elif style_str == 'aniso_err':
        Dval, Dvals = 0.5e12 / tau, 0.5e12 / taus
        for i in range(3):
            b = get_flex_bounds(tau[i], taus[:, i])
            L.append('# model fit, e_%i tau = %e +- %e %e [ps]' % (i, b[0], b[1], b[2]))
            b = get_flex_bounds(Dval[i], Dvals[:, i])
            L.append('# Converted D_%i = %e +- %e %e [s^-1]' % (i, b[0], b[1], b[2]))
        a = calculate_anisotropies(Dval, Dvals)
        L += ["# Converted Diso = %e +- %e [s^-1]" % a[0], "# Converted Dani_L = %f +- %f" % a[1],
              "# Converted Drho_L = %f +- %f" % a[2], "# Converted Dani_S = %f +- %f" % a[3],
              "# Converted Drho_S = %f +- %f" % a[4]]
        for j in range(len(taus)):
            L += ['# Chunk_%d D_%d = %e [s^-1]' % (j, i, Dvals[j, i]) for i in range(3)]
        L.append("# t <1-2x^2> <1-2y^2> <1-2z^2>")
    return L


def format_header_quat(q):
    return '# Quaternion orientation frame: %f %f %f %f' % (q[0], q[1], q[2], q[3])


def print_model_fits_gen(fname, ydims, str_header, xlist, ylist):
    """xmgrace writer of the decay curves and their model fits (:277-330)."""
    n = len(xlist)
    with open(fname, 'w') as fp:
        for line in str_header:
            print("%s" % line, file=fp)
        if ydims == 1:
            for i in range(n):
                print("%g %g" % (xlist[i], ylist[i]), file=fp)
        elif ydims == 2:
            for s, y in enumerate(ylist):
                print("@target g%d.s%d" % (0, s), file=fp)
                for i in range(n):
                    print("%g %g" % (xlist[i], y[i]), file=fp)
                print("&", file=fp)
        elif ydims == 3:
            ng = len(ylist)
            print("dim1: ", ng)
            for g in range(ng):
                print("@g%d on" % g, file=fp)
            for g in range(ng):
                print("dim2: ", len(ylist[g]))
                for s, y in enumerate(ylist[g]):
                    print("@target g%d.s%d" % (g, s), file=fp)
                    for i in range(n):
                        print("%g %g" % (xlist[i], y[i]), file=fp)
                    print("&", file=fp)
            print("@arrange(%i, %i, 0.1, 0.1, 0.1)" % (2, int(0.5 * ng + 0.5)), file=fp)
            for g in range(ng):
                print("@with g%i" % g, file=fp)
                if g == 0:
                    print("@subtitle \"Aggregate Data\"", file=fp)
                print("@autoscale", file=fp)
        else:
            print("= = = Critical ERROR: invalid dimension specifier in print_model_fits_gen!")
            sys.exit(1)


def print_axes_as_xyz(fname, moilist):
    with open(fname, 'w') as fp:
        for m in moilist:
            print("3", file=fp)
            print("AXES", file=fp)
            for name, row in zip("XYZ", m):
                print("%s %g %g %g" % (name, row[0], row[1], row[2]), file=fp)


# ---- CLI ------------------------------------------------------------------------------------------
def build_parser():
    p = argparse.ArgumentParser(description='Calculates the difference quaternions from PLUMED output: '
                                'a time-series of quaternion representation of orientations '
                                'then manipulate them in various ways',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('-f', '--infn', type=str, dest='infn', default=['colvar-q'], nargs='+',
                   help='Input file(s) in PLUMED quaternion form. Assumes that dt is identical between every frame! '
                        'Several equally long files are pooled as replicas (calculate-dq-distribution-multi.py).')
    p.add_argument('-o', '--outpref', type=str, dest='out_pref', default='out', help='Output file prefix.')
    p.add_argument('--hist', dest='bDoHist', action='store_true', default=False,
                   help='Record the 3D-histogram of rotation-quaternions dq at each delay time, dt.')
    p.add_argument('-o2', '--outtype', type=str, dest='out_suff', default='dat')
    p.add_argument('--iso', dest='bDoIso', action='store_true', default=False,
                   help='Record the isotropic decay of dq.')
    p.add_argument('--aniso', dest='bDoAniso', action='store_true', default=False,
                   help='Record an estimate of the anisotropic decay of dq.')
    p.add_argument('--fulltensor', dest='bDoFullTensor', action='store_true', default=False,
                   help='Record all nine components of the tensor <q_i q_j> in the PAF frame.')
    p.add_argument('-n', '--num_bins', type=int, dest='num_bins', default=101)
    p.add_argument('--mindt', '--min_dt', type=float, dest='min_dt', default=0.0,
                   help='Minimum interval delta_t to calculate in picoseconds [ps].')
    p.add_argument('--num_chunk', '--num_chunks', type=int, dest='num_chunk', default=0,
                   help='Uncertainty estimation from N sub-chunks of the trajectory.')
    p.add_argument('--maxdt', '--max_dt', type=float, dest='max_dt', default=1000.0,
                   help='Maximum interval delta_t to calculate in picoseconds [ps].')
    p.add_argument('--skip', '--skip_dt', type=float, dest='skip_dt', default=0.0,
                   help='Interval between successive delta_t, same units.')
    return p


def rotmatrix_to_quaternion(time_axis, matrix, bInvert=False):
    """calculate-dq-distribution.py:389-408 for `gmx rotmat` input: every row of nine matrix elements -> unit
    quaternion (transforms3d `mat2quat`: principal eigenvector of the symmetric 4x4 K matrix, w made non-negative),
    inverted when bInvert (`qinverse` = conjugate / |q|^2).  Returns the (5, n) time + w x y z table the PLUMED
    reader returns.  Input conversion on the host (one batched eigh); transforms3d is a third-party dependency absent
    from this image, restated from its published algorithm."""
    n = len(time_axis)
    if n != len(matrix):
        print("= = = ERROR in rotmatrix_to_quaternion: lengths are not the same!", file=sys.stderr)
        return None
    M = np.asarray(matrix, dtype=float).reshape(n, 9)
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = (M[:, i] for i in range(9))     # mat2quat reads M.flat in this order
    K = np.zeros((n, 4, 4))
    K[:, 0, 0] = Qxx - Qyy - Qzz
    K[:, 1, 0], K[:, 1, 1] = Qyx + Qxy, Qyy - Qxx - Qzz
    K[:, 2, 0], K[:, 2, 1], K[:, 2, 2] = Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy
    K[:, 3, 0], K[:, 3, 1], K[:, 3, 2], K[:, 3, 3] = Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz
    K /= 3.0
    vals, vecs = np.linalg.eigh(K)                     # lower triangle, as numpy's default UPLO='L'
    top = np.argmax(vals, axis=1)
    q = np.take_along_axis(vecs, top[:, None, None], axis=2)[:, [3, 0, 1, 2], 0]
    q = np.ascontiguousarray(np.where(q[:, :1] < 0, -q, q))
    if bInvert:
        # |q|^2 as np.dot(q, q) rounds it for one contiguous quaternion (the batched matmul takes the same BLAS path;
        # a plain sum over the last axis differs in the last bit for a quarter of the frames)
        n2 = np.matmul(q[:, None, :], q[:, :, None])[:, 0, :]
        q = q * np.array([1.0, -1.0, -1.0, -1.0]) / n2
    out = np.zeros((5, n))
    out[0] = time_axis
    out[1:5] = q.T
    return out


def read_quaternion_input(fn):
    """Input dispatch of calculate-dq-distribution.py:487-500: `.xvg` = GROMACS `gmx rotmat` rotation matrices
    (converted with bInvert=True), anything else = PLUMED print file.  Returns (fields, data (5+, n))."""
    if fn.endswith('.xvg'):
        x, y = io_formats.load_xys(fn)
        return ["time", "w", "x", "y", "z"], rotmatrix_to_quaternion(x, y, bInvert=True)
    return io_formats.read_from_plumedprint(fn)


def main(argv=None):
    time_start = time.time()
    args = build_parser().parse_args(argv)
    if args.out_suff not in ("dx", "dat", "none"):
        print("= = ERROR in input: histogram output type must be either dx, or dat, or none.")
        sys.exit()
    infn = [args.infn] if isinstance(args.infn, str) else list(args.infn)
    fields, data = read_quaternion_input(infn[0])
    replicas = [data]
    for fn in infn[1:]:
        _, d = read_quaternion_input(fn)
        if d.shape != data.shape:
            print("= = ERROR: replica trajectories must have identical lengths (%s vs %s)." % (d.shape, data.shape),
                  file=sys.stderr)
            sys.exit(1)
        replicas.append(d)
    nfield, ndat = data.shape
    print("= = Input data found to be %i fields and %i entries. = =" % (nfield, ndat))
    qprev = data[1:5, 0]
    print("= = Initial quaternion read: (%f %f %f %f) = =" % (qprev[0], qprev[1], qprev[2], qprev[3]))
    min_int, max_int, skip_int, ddt = lag_grid(data[0], args.min_dt, args.max_dt, args.skip_dt)
    num_int = int(np.floor((max_int - min_int) / skip_int) + 1)
    print("= = Will calculate statistics for %i intervals between %g - %g ps, every %g ps) = ="
          % (num_int, min_int * ddt, max_int * ddt, args.skip_dt))
    print("= = ...corresponding to %i - %i frames, every %i frames. = =" % (min_int, max_int, skip_int))
    if max_int * ddt > (data[0, -1] - data[0, 0]) / 2.0:
        print("= = = ERROR: max_dt requested (%g ps) is greater than half of the entire trajectory (%g ps)!"
              % (max_int * ddt, (data[0, -1] - data[0, 0]) / 2.0))
        print("             ...will refuse to calculate correlation.")
        sys.exit(1)
    lags = np.arange(min_int, max_int + 1, skip_int)
    nch = args.num_chunk
    sub = nch > 1
    qall = np.stack([np.ascontiguousarray(d[1:5].T) for d in replicas])
    res = dq_curves(qall if len(replicas) > 1 else qall[0], lags, ddt, nch, do_aniso=args.bDoAniso)
    if args.bDoHist and args.out_suff != "none":
        if len(replicas) > 1:
            print("= = ERROR: --hist works on a single trajectory.", file=sys.stderr)
            sys.exit(1)
        for delta in lags:                                                     # :632-647
            hist3, edges3 = dq_histogram3d(qall[0], int(delta), args.num_bins)
            out_file = args.out_pref + "-hist-" + str(delta * ddt) + "ps." + args.out_suff
            if out_file.endswith("dx"):
                xmin = [(e[0] + e[1]) / 2.0 for e in edges3]
                abc = np.zeros((3, 3))
                for i in range(3):
                    abc[i][i] = (edges3[i][-1] - edges3[i][0]) / args.num_bins
                io_formats.write_to_dx(out_file, hist3, (args.num_bins,) * 3, xmin, abc, 'nm')
            else:
                io_formats.print_gplot_hist(out_file, hist3, edges3)
    dt = res["dt"]
    time_chk1 = time.time()
    pref = args.out_pref
    if args.bDoIso:
        tau = conduct_exponential_fit(dt, res["iso"], 1.5, -0.5)
        model = isotropic_decay(dt, tau)
        if sub:
            chtaus = [conduct_exponential_fit(dt, res["chunk_iso"][i], 1.5, -0.5) for i in range(nch)]
            plist = [[res["iso"], model]] + [[res["chunk_iso"][i], isotropic_decay(dt, chtaus[i])] for i in range(nch)]
            print_model_fits_gen(pref + "-iso.dat", 3, format_header('iso_err', tau, chtaus), dt, plist)
        else:
            print_model_fits_gen(pref + "-iso.dat", 2, format_header('iso', tau), dt, [res["iso"], model])
    if args.bDoAniso:
        print("= = = Running exponential fitting of fully anisotropic D...")
        taus = np.array([conduct_exponential_fit(dt, res["aniso2"][i], 0.5, 0.5) for i in range(3)])
        models = anisotropic_decay_noc(dt, taus.reshape((3, 1)))
        if sub:
            print("= = = Running exponential fitting over sub-chunks as well for uncertainty analysis...")
            chtaus = np.zeros((nch, 3))
            chmodels = np.zeros((nch, 3, len(dt)))
            for i in range(nch):
                for j in range(3):
                    chtaus[i, j] = conduct_exponential_fit(dt, res["chunk_aniso2"][i][j], 0.5, 0.5)
                chmodels[i] = anisotropic_decay_noc(dt, chtaus[i].reshape((3, 1)))
            header = format_header('aniso_err', taus, chtaus) + [format_header_quat(res["q_frame"])]
            plist = [np.concatenate((res["aniso2"], models))]
            plist += [np.concatenate((res["chunk_aniso2"][i], chmodels[i])) for i in range(nch)]
            print_model_fits_gen(pref + "-aniso2.dat", 3, header, dt, plist)
        else:
            header = format_header('aniso', taus) + [format_header_quat(res["q_frame"])]
            print_model_fits_gen(pref + "-aniso2.dat", 2, header, dt, np.concatenate((res["aniso2"], models)))
        io_formats.print_xylist(pref + "-aniso_q.dat", dt, res["qrot"], bCols=True)
        print_axes_as_xyz(pref + "-moi.xyz", res["moi_axes"])
    if args.bDoFullTensor:
        io_formats.print_xylist(pref + "-tensor.dat", dt, res["moiR"].reshape(len(dt), 9).T)
    time_stop = time.time()
    print("= = Total seconds elapsed: %g" % (time_stop - time_start))
    print("= = Time of Read and fit halves: %g , %g" % (time_chk1 - time_start, time_stop - time_chk1))


if __name__ == '__main__':
    main()
