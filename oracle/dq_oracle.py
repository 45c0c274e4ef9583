"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

NumPy restatement of the quaternion-displacement statistics of calculate-dq-distribution.py and the
quaternion helpers of transforms3d_supplement.py it calls.  Pinned by tests/golden (reference run in
the build container).  The third-party `transforms3d.quaternions` calls of the reference are restated
from that package's published (w, x, y, z) Hamilton-product convention; parity at that boundary is
pinned only by our own golden vectors (the reference ships no tests).
"""
import math

import numpy as np
from scipy.optimize import fmin_powell

IDENTITY = (1.0, 0.0, 0.0, 0.0)


# ---- quaternion helpers (transforms3d_supplement.py) ---------------------------------------------
def quat_mult(q1, q2):
    """quat_mult_simd, transforms3d_supplement.py:163-183 (last axis = w,x,y,z)."""
    out = np.zeros_like(q1)
    out[..., 0] = q1[..., 0] * q2[..., 0] - np.einsum("...i,...i", q1[..., 1:4], q2[..., 1:4])
    out[..., 1:4] = (q1[..., 0, None] * q2[..., 1:4] + q2[..., 0, None] * q1[..., 1:4]
                     + np.cross(q1[..., 1:4], q2[..., 1:4]))
    return out


def quat_conj(q):
    """quat_invert, transforms3d_supplement.py:185-186 (a Python-float list promotes float32 to float64)."""
    return q * [1.0, -1.0, -1.0, -1.0]


def quat_reduce(q, qref=IDENTITY):
    """quat_reduce_simd, transforms3d_supplement.py:219-227: flip sign so that q.qref >= 0 (0 counts as +)."""
    sgn = np.sign(np.einsum("...i,i", q, qref))
    sgn[sgn == 0] = 1.0
    return q * sgn[:, None]


def self_dq(q, delta):
    """obtain_self_dq, calculate-dq-distribution.py:102-109: conj(q_t) * q_{t+delta}, imaged to w >= 0."""
    return quat_reduce(quat_mult(quat_conj(q[:-delta]), q[delta:]))


def _hamilton(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def _rot1(v, q):
    """transforms3d.quaternions.rotate_vector: vector part of q [0,v] q*."""
    qc = np.array(q) * np.array([1.0, -1.0, -1.0, -1.0])
    return _hamilton(q, _hamilton(np.concatenate(([0.0], v)), qc))[1:]


def _quat_v1v2(v1, v2):
    """quat_v1v2, transforms3d_supplement.py:71-83: minimum-angle rotation taking v1 onto v2."""
    th = math.acos(np.dot(v1, v2))
    ax = np.cross(v1, v2)
    if all(np.isnan(ax)):
        return np.array(IDENTITY)
    ax = np.array(ax, dtype=float)
    ax = ax / math.sqrt(float(np.dot(ax, ax)))      # transforms3d.quaternions.axangle2quat
    return np.concatenate(([math.cos(th / 2.0)], ax * math.sin(th / 2.0)))


def frame_transform_min(axes):
    """quat_frame_transform_min, transforms3d_supplement.py:137-149: align axes[2] with +-Z, then the
    rotated axes[0] with +-X, each time keeping the image with the larger q_w."""
    q1a, q1b = _quat_v1v2(axes[2], (0, 0, 1)), _quat_v1v2(axes[2], (0, 0, -1))
    q1 = q1a if q1a[0] > q1b[0] else q1b
    x_rot = _rot1(axes[0], q1)
    q2a, q2b = _quat_v1v2(x_rot, (1, 0, 0)), _quat_v1v2(x_rot, (-1, 0, 0))
    q2 = q2a if q2a[0] > q2b[0] else q2b
    return _hamilton(q2, q1)


def nearly_identity(q, rtol=1e-5, atol=1e-8):
    """transforms3d.quaternions.nearly_equivalent(q, (1,0,0,0))."""
    q = np.asarray(q, dtype=float)
    e = np.array(IDENTITY)
    return bool(np.allclose(q, e, rtol, atol) or np.allclose(-q, e, rtol, atol))


def rotate_rows(v, q):
    """qs.rotate_vector_simd(v, q) (transforms3d_supplement.py:270-296) for (n,3) v and one quaternion."""
    q = np.asarray(q, dtype=float)
    q = np.nan_to_num(q / np.linalg.norm(q))
    a = np.cross(q[1:4], v) + q[0] * v
    b = np.cross(q[1:4], a)
    return b + b + v


# ---- per-lag statistics (calculate-dq-distribution.py) --------------------------------------------
def iso_moment_shipped(vq):
    """average_LegendreP1quat AS SHIPPED, calculate-dq-distribution.py:111-112 (quirk G1):
    apply_along_axis(..., axis=0) sums over frames, so the value is mean_c(1 - 2 sum_t v_c(t)^2)."""
    return np.mean(1.0 - 2.0 * np.sum(np.square(vq), axis=0))


def iso_moment_intended(vq):
    """What the comment at :113-116 describes: <1 - 2 |v|^2> over frames."""
    return np.mean(1.0 - 2.0 * np.sum(np.square(vq), axis=1))


def aniso_tensor(vq, qframe=IDENTITY):
    """average_anisotropic_tensor, calculate-dq-distribution.py:118-126: mean_t (R v)(R v)^T."""
    if not nearly_identity(qframe):
        vq = rotate_rows(vq, qframe)
    return np.mean(np.einsum("ij,ik->ijk", vq, vq), axis=0)


def _blocks(ndat, nchunk):
    nblock = int(math.ceil(1.0 * ndat / nchunk))
    return [(nblock * i, min(ndat, nblock * (i + 1))) for i in range(nchunk)]


def iso_moment_chunks(vq, nchunk):
    """average_LegendreP1quat_chunk, :128-135."""
    return np.array([iso_moment_shipped(vq[a:b]) for a, b in _blocks(len(vq), nchunk)])


def aniso_tensor_chunks(vq, nchunk, qframe=IDENTITY):
    """average_anisotropic_tensor_chunk, :137-144."""
    return np.array([aniso_tensor(vq[a:b], qframe) for a, b in _blocks(len(vq), nchunk)])


def lag_grid(times, min_dt, max_dt, skip_dt):
    """Lag bookkeeping of the main script, calculate-dq-distribution.py:510-523.
    Returns (list of frame lags, frame spacing in time units)."""
    ddt = times[1] - times[0]
    skip_int = max(1, int(skip_dt / ddt))
    min_int = max(skip_int, int(min_dt / ddt))
    max_int = int(max_dt / ddt)
    return list(range(min_int, max_int + 1, skip_int)), ddt


def dq_curves(q, lags, ddt, nchunk=0, do_aniso=True):
    """Main lag loop, calculate-dq-distribution.py:554-650, for q (N,4) float32 and integer frame lags.

    Returns a dict with dt (nl,), iso (nl,), aniso1 (3,nl), aniso2 (3,nl), qrot (4,nl), moi_axes (nl,3,3),
    moi (nl,3,3) [lab-frame tensor], q_frame (4,), and if nchunk>1 chunk_iso (nchunk,nl), chunk_aniso2
    (nchunk,3,nl).  q_frame is frozen at the first lag (:586-591)."""
    nl = len(lags)
    out = dict(dt=np.zeros(nl), iso=np.zeros(nl), aniso1=np.zeros((3, nl)), aniso2=np.zeros((3, nl)),
               qrot=np.zeros((4, nl)), moi_axes=np.zeros((nl, 3, 3)), moi=np.zeros((nl, 3, 3)))
    if nchunk > 1:
        out["chunk_iso"] = np.zeros((nchunk, nl))
        out["chunk_aniso2"] = np.zeros((nchunk, 3, nl))
    q_frame = IDENTITY
    first = True
    for k, delta in enumerate(lags):
        v = self_dq(q, delta)[..., 1:4]
        out["dt"][k] = delta * ddt
        out["iso"][k] = iso_moment_shipped(v)
        moi = aniso_tensor(v)
        out["moi"][k] = moi
        moiR = aniso_tensor(v, q_frame) if not nearly_identity(q_frame) else moi
        if do_aniso:
            eigval, eigvec = np.linalg.eigh(moi)
            axes = eigvec.T
            q_rot = frame_transform_min(axes)
            if first:
                first = False
                q_frame = q_rot
                moiR = aniso_tensor(v, q_frame)
            out["aniso1"][:, k] = 1 - 2 * eigval
            out["aniso2"][:, k] = 1 - 2 * np.diag(moiR)
            out["qrot"][:, k] = q_rot
            out["moi_axes"][k] = axes
        if nchunk > 1:
            out["chunk_iso"][:, k] = iso_moment_chunks(v, nchunk)
            t2 = aniso_tensor_chunks(v, nchunk, q_frame) if not nearly_identity(q_frame) \
                else aniso_tensor_chunks(v, nchunk)
            out["chunk_aniso2"][:, :, k] = [[1 - 2 * t2[i][j, j] for j in range(3)] for i in range(nchunk)]
    out["q_frame"] = np.array(q_frame, dtype=float)
    return out


# ---- fits (host side in the product too; restated for the D-tensor parity tests) -------------------
def _expdecay_cost(pos, x, y, C0, C1):
    """powell_expdecay, :152-167 (mean squared residual of C0 exp(-x/A) + C1)."""
    A = float(np.ravel(pos)[0])
    chi2 = 0.0
    for i in range(len(x)):
        chi2 += (C0 * math.exp(-x[i] / A) + C1 - y[i]) ** 2
    return chi2 / len(x)


def exponential_fit(x, y, C0, C1):
    """conduct_exponential_fit, :199-207: guess from the first two points (:195-196), SciPy Powell defaults."""
    guess = (x[0] - x[1]) / math.log((y[1] - C1) / (y[0] - C1))
    res = fmin_powell(_expdecay_cost, guess, args=(x, y, C0, C1), full_output=True, disp=False)
    return np.ravel(res[0])[0]


def anisotropies(D):
    """calculate_anisotropies without chunks, :56-80: (iso, ani_L, rhomb_L, ani_S, rhomb_S) of sorted D."""
    D = np.sort(np.asarray(D, dtype=float))

    def ani(d):
        return 2 * d[2] / (d[1] + d[0])

    def rho(d):
        return 3 * (d[1] - d[0]) / (2 * d[2] - d[1] - d[0])

    return np.mean(D), ani(D), rho(D), ani(D[::-1]), rho(D[::-1])


def pooled_vectors(q_replicas, delta):
    """calculate-dq-distribution-multi.py:533-539: the displacement vectors of every replica, concatenated."""
    return np.concatenate([self_dq(q, delta)[..., 1:4] for q in q_replicas], axis=0)


def dq_histogram3d(q, delta, nbins=101):
    """calculate-dq-distribution.py:527-528, 633-634 (`normed=True` is NumPy < 1.24's spelling of density)."""
    v = self_dq(q, delta)[..., 1:4]
    return np.histogramdd(v, range=[(-1, 1)] * 3, bins=(nbins,) * 3, density=True)
