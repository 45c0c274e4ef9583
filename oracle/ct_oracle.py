"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU restatement, in NumPy, of the bond-vector part of SpinRelax's hot path.  Every function cites the
reference lines it follows (paths are inside the SpinRelax tree).  Pinned against outputs of the
reference itself: tests/golden/make_golden.py runs the reference functions in the build container and
tests/test_oracle_golden.py checks this file against those committed vectors.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# C(t)
# ------------------------------------------------------------------------------------------------
def ct_time_axis(dt, tau):
    """calculate_dt, calculate-Ct-from-traj.py:240-243."""
    return (np.arange(int(0.5 * tau / dt)) + 1.0) * dt


def reformat_by_tau(traj_list, dt, tau):
    """reformat_vecs_by_tau, calculate-Ct-from-traj.py:245-275: drop the remainder frames of each
    trajectory, concatenate, view as (chunks, frames_per_chunk, bonds, 3)."""
    per = int(tau / dt)
    kept = [np.asarray(t)[: (len(t) // per) * per] for t in traj_list]
    cat = np.concatenate(kept, axis=0)
    return cat.reshape(len(cat) // per, per, cat.shape[1], cat.shape[2])


def ct_palmer(vecs):
    """calculate_Ct_Palmer, calculate-Ct-from-traj.py:200-238.

    vecs (nC, nF, nR, 3).  For delta = 1..nF//2 (:222): P2 of the dot products of frames t and t+delta
    (:225), averaged over the nF-delta pairs of each chunk (:226), then mean over chunks (:227) and
    population std over chunks divided by (sqrt(nC) - 1) (:228).  Arithmetic stays in vecs.dtype (:219).
    """
    nC, nF, nR, _ = vecs.shape
    L = int(nF / 2)
    Ct = np.zeros((L, nR), dtype=vecs.dtype)
    dCt = np.zeros((L, nR), dtype=vecs.dtype)
    for lag in range(1, L + 1):
        head, tail = vecs[:, : nF - lag], vecs[:, lag:]
        cosang = np.einsum("ctrx,ctrx->ctr", head, tail)
        p2 = -0.5 + 1.5 * np.square(cosang)
        per_chunk = np.einsum("ctr->cr", p2) / (nF - lag)
        Ct[lag - 1] = per_chunk.mean(axis=0)
        dCt[lag - 1] = per_chunk.std(axis=0) / (np.sqrt(nC) - 1.0)
    return Ct, dCt


def ct_lag_body(vecs, lag):
    """One iteration of the hot loop (:223-228) for a single lag; used to time the CPU baseline on a
    lag subset of workloads whose full evaluation would take hours."""
    nC, nF, nR, _ = vecs.shape
    cosang = np.einsum("ctrx,ctrx->ctr", vecs[:, : nF - lag], vecs[:, lag:])
    per_chunk = np.einsum("ctr->cr", -0.5 + 1.5 * np.square(cosang)) / (nF - lag)
    return per_chunk.mean(axis=0), per_chunk.std(axis=0) / (np.sqrt(nC) - 1.0)


def ct_lag_sums_fft(vecs, L=None):
    """Scalable float64 oracle for big shapes: with a = the six products u_i u_j,
    sum_t (u_t.u_{t+d})^2 = sum_k w_k * autocorr(a_k)[d], w = (1,1,1,2,2,2).  Returns S (nR, nC, L)."""
    v = np.asarray(vecs, dtype=np.float64)
    nC, nF, nR, _ = v.shape
    if L is None:
        L = nF // 2
    n = 1
    while n < 2 * nF:
        n *= 2
    S = np.zeros((nR, nC, L))
    for (i, j, w) in ((0, 0, 1.0), (1, 1, 1.0), (2, 2, 1.0), (0, 1, 2.0), (0, 2, 2.0), (1, 2, 2.0)):
        a = v[..., i] * v[..., j]                       # (nC, nF, nR)
        F = np.fft.rfft(a, n=n, axis=1)
        ac = np.fft.irfft(F * np.conj(F), n=n, axis=1)[:, 1 : L + 1]   # lags 1..L
        S += w * np.transpose(ac, (2, 0, 1))
    return S


def ct_from_lag_sums(S, nF, dtype=np.float32):
    """Palmer statistics from S (nR, nC, L): (:225-228) with the affine P2 map hoisted out of the sum."""
    nR, nC, L = S.shape
    nvals = nF - np.arange(1, L + 1)
    per_chunk = -0.5 + 1.5 * (S / nvals[None, None, :])
    Ct = per_chunk.mean(axis=1).T
    dCt = (per_chunk.std(axis=1) / (np.sqrt(nC) - 1.0)).T
    return Ct.astype(dtype), dCt.astype(dtype)


# ------------------------------------------------------------------------------------------------
# PAF rotation, spherical coordinates, Lambert-cylindrical histogram
# ------------------------------------------------------------------------------------------------
def rotate_vectors(v, q):
    """qs.rotate_vector_simd(v, q) with default arguments, transforms3d_supplement.py:270-296:
    q is normalised (vecnorm_NDarray :40-52), a = q_v x v + q_w v, b = q_v x a, result b + b + v.
    A float64 quaternion promotes float32 vectors to float64."""
    q = np.asarray(q, dtype=np.float64)
    q = np.nan_to_num(q / np.linalg.norm(q))
    qw, qv = q[0], q[1:4]
    v = np.array(v)
    a = np.cross(qv, v) + qw * v
    b = np.cross(qv, a)
    return b + b + v


def xyz_to_rtp(uv, bUnit=False):
    """gm.xyz_to_rtp(uv) for the last-axis case, general_maths.py:143-158; bUnit: :131-133 (divides by phi)."""
    if bUnit:
        out = np.zeros(uv.shape[:-1] + (2,), dtype=uv.dtype)
        out[..., 0] = np.arctan2(uv[..., 1], uv[..., 0])
        out[..., 1] = np.arccos(uv[..., 2] / out[..., 0])
        return out
    out = np.zeros_like(uv)
    out[..., 0] = np.linalg.norm(uv, axis=-1)
    out[..., 1] = np.arctan2(uv[..., 1], uv[..., 0])
    out[..., 2] = np.arccos(uv[..., 2] / out[..., 0])
    return out


def spherical_by_residue(frames_vecs, q_rot=None):
    """--vecDist without --vecHist (calculate-Ct-from-traj.py:567,588,600): (nR, frames, 3) r/phi/theta."""
    w = rotate_vectors(frames_vecs, q_rot) if q_rot is not None else frames_vecs
    with np.errstate(all="ignore"):
        return np.transpose(xyz_to_rtp(w), axes=(1, 0, 2))


def sphere_histogram(frames_vecs, q_rot=None, nbins_phi=72):
    """The --vecRot / --vecHist block of calculate-Ct-from-traj.py.

    frames_vecs (frames, nR, 3) float32 (the 4-D array flattened over chunks, :535-536).  Optional
    rotation (:567), spherical coordinates (:588), transpose to vector-major and drop r (:600,611),
    cos(theta) (:613), per-vector np.histogramdd with bins (nbx, nbx//2) over ((-pi,pi),(-1,1))
    (:615-626, without the `normed` keyword that NumPy >= 1.24 rejects).  Returns (hist, edges) where
    hist is (nR, nbx, nby) in the dtype of the coordinates and edges is the two-element edge list.
    """
    v = frames_vecs
    if q_rot is not None:
        v = rotate_vectors(v, q_rot)
    rtp = xyz_to_rtp(v)
    rtp = np.transpose(rtp, axes=(1, 0, 2))
    pc = np.delete(rtp, 0, axis=2)
    pc[..., 1] = np.cos(pc[..., 1])
    nby = int(nbins_phi / 2)
    hist = np.zeros((pc.shape[0], nbins_phi, nby), dtype=pc.dtype)
    edges = None
    for i in range(pc.shape[0]):
        h, e = np.histogramdd(pc[i], bins=(nbins_phi, nby), range=((-np.pi, np.pi), (-1, 1)))
        if edges is None:
            edges = e
        hist[i] = h
    return hist, edges


def average_vector(frames_vecs):
    """--vecAvg: normalised mean over frames (:579-583, gs.normalise_vector_array)."""
    m = np.mean(frames_vecs, axis=0)
    return m / np.linalg.norm(m, axis=-1, keepdims=True)


def s2_outer_product(frames_vecs, delta_t=-1, tau_memory=-1):
    """calculate_S2_by_outerProduct for (frames, nR, 3) input, calculate-Ct-from-traj.py:121-142."""
    n = frames_vecs.shape[0]
    if delta_t < 0 or tau_memory < 0:
        m = np.einsum("trx,try->rxy", frames_vecs, frames_vecs) / n
        return 1.5 * np.einsum("rxy,rxy->r", m, m) - 0.5
    per = int(tau_memory / delta_t)
    nb = int(n / per)
    blk = frames_vecs[: nb * per].reshape(nb, per, frames_vecs.shape[1], 3)
    m = np.einsum("btrx,btry->brxy", blk, blk) / per
    s2 = 1.5 * np.einsum("brxy,brxy->br", m, m) - 0.5
    return np.stack((s2.mean(axis=0), s2.std(axis=0) / (np.sqrt(nb) - 1.0)), axis=-1)
