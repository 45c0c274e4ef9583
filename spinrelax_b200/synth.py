"""Seeded synthetic inputs for the hot path.

N-H bond vectors: random equilibrium axis per vector plus two Ornstein-Uhlenbeck tangent-plane
perturbations (tau1 = 50 ps sigma 0.25, tau2 = 1500 ps sigma 0.30) sampled every dt = 10 ps,
renormalised and cast to float32 -- C(t) decays from ~0.93 to ~0.4.  Layout is the reference's
(frames, bonds, 3) as produced by obtain_XHvecs (calculate-Ct-from-traj.py:64-86).

Quaternion trajectories: q(t+1) = q(t) * d(t), d = normalise(1, eps), eps ~ N(0, diag(sigma)^2),
float32-rounded like the PLUMED reader does (plumedcolvario.py:68).
"""
import numpy as np
from scipy.signal import lfilter

BASE_SEED = 20260101


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _ou(rng, n_frames, shape, dt, tau, sigma):
    a = np.exp(-dt / tau)
    eps = rng.standard_normal((n_frames,) + shape)
    eps[0] /= np.sqrt(1.0 - a * a)  # stationary start
    x = lfilter([np.sqrt(1.0 - a * a) * sigma], [1.0, -a], eps, axis=0)
    return x


def nh_vectors(n_frames, n_vec, seed=BASE_SEED, dt=10.0, tumbling_sigma=0.0, block=250000):
    """(n_frames, n_vec, 3) float32 unit vectors. tumbling_sigma > 0 adds a global isotropic rotational
    random walk (rotation vector per step ~ N(0, sigma^2)) to mimic an un-fitted trajectory (Ctext)."""
    rng = np.random.default_rng(seed)
    axis = _unit(rng.standard_normal((n_vec, 3)))
    helper = np.where(np.abs(axis[:, :1]) < 0.9, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
    e1 = _unit(np.cross(axis, helper))
    e2 = np.cross(axis, e1)
    p = _ou(rng, n_frames, (n_vec, 2), dt, 50.0, 0.25) + _ou(rng, n_frames, (n_vec, 2), dt, 1500.0, 0.30)
    out = np.empty((n_frames, n_vec, 3), dtype=np.float32)
    for s in range(0, n_frames, block):
        e = min(n_frames, s + block)
        v = axis[None] + p[s:e, :, 0:1] * e1[None] + p[s:e, :, 1:2] * e2[None]
        out[s:e] = _unit(v).astype(np.float32)
    if tumbling_sigma > 0.0:
        q = quaternion_walk(n_frames, seed=seed + 7, sigma=(tumbling_sigma,) * 3, dtype=np.float64)
        out = rotate_by_quats(out, q).astype(np.float32)
        out = _unit(out.astype(np.float64)).astype(np.float32)
    return out


def quat_mult(a, b):
    w1, x1, y1, z1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    w2, x2, y2, z2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2), axis=-1)


def quaternion_walk(n_frames, seed=BASE_SEED + 3, sigma=(0.01, 0.015, 0.03), dtype=np.float32):
    """(n_frames, 4) orientation quaternions (w, x, y, z) of a (possibly anisotropic) rotational random walk."""
    rng = np.random.default_rng(seed)
    d = np.empty((n_frames, 4))
    d[:, 0] = 1.0
    d[:, 1:] = rng.standard_normal((n_frames, 3)) * np.asarray(sigma)[None]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = (1.0, 0.0, 0.0, 0.0)
    # cumulative product by doubling (associative scan), O(N log N) numpy work
    q = d.copy()
    step = 1
    while step < n_frames:
        q[step:] = quat_mult(q[:-step].copy(), q[step:])
        step *= 2
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(dtype)


def rotate_by_quats(v, q):
    """Rotate v (n, m, 3) by per-frame unit quaternions q (n, 4)."""
    qw = q[:, None, 0:1]
    qv = q[:, None, 1:4]
    a = np.cross(qv, v) + qw * v
    b = np.cross(qv, a)
    return b + b + v


def write_plumed_quaternions(path, q, dt=10.0):
    """PLUMED PRINT format consumed by plumedcolvario.read_from_plumedprint (plumedcolvario.py:24-81)."""
    with open(path, "w") as fp:
        fp.write("#! FIELDS time q.w q.x q.y q.z\n")
        for i in range(len(q)):
            fp.write("%f %.9g %.9g %.9g %.9g\n" % (i * dt, q[i, 0], q[i, 1], q[i, 2], q[i, 3]))


def backbone_trajectory(n_frames, n_res, seed=BASE_SEED + 21, tumbling_sigma=0.02, noise=0.005, dt=10.0):
    """Synthetic protein-like Cartesian trajectory for the front end (obtain_XHvecs + superposition).

    Five atoms per residue in the order N, H, CA, C, O (nm).  A fixed random-coil CA trace (0.38 nm steps) carries
    rigid backbone atoms; the amide H sits 0.102 nm from N along a synthetic N-H unit vector (`nh_vectors`), every
    atom gets Gaussian positional noise, and the whole molecule tumbles (rotational random walk) and drifts.
    Returns (xyz (frames, 5 n_res, 3) float32, selections dict, ref_xyz (5 n_res, 3) float32 = noise-free,
    un-tumbled structure with the equilibrium N-H directions).
    """
    rng = np.random.default_rng(seed)
    steps = _unit(rng.standard_normal((n_res, 3))) * 0.38
    ca = np.cumsum(steps, axis=0)
    off = rng.standard_normal((n_res, 3, 3)) * 0.08                      # N, C, O offsets from CA
    nh = nh_vectors(n_frames, n_res, seed=seed + 1, dt=dt).astype(np.float64)
    base = np.empty((n_res, 5, 3))
    base[:, 0] = ca + off[:, 0]
    base[:, 2] = ca
    base[:, 3] = ca + off[:, 1]
    base[:, 4] = ca + off[:, 2]
    xyz = np.repeat(base[None], n_frames, axis=0)
    xyz[:, :, 1] = xyz[:, :, 0] + 0.102 * nh
    xyz += rng.standard_normal(xyz.shape) * noise
    xyz = xyz.reshape(n_frames, n_res * 5, 3)
    xyz -= xyz.mean(axis=1, keepdims=True)
    q = quaternion_walk(n_frames, seed=seed + 2, sigma=(tumbling_sigma,) * 3, dtype=np.float64)
    xyz = rotate_by_quats(xyz, q) + np.cumsum(rng.standard_normal((n_frames, 1, 3)) * 0.01, axis=0)
    ref = base.copy()
    ref[:, 1] = ref[:, 0] + 0.102 * _unit(nh.mean(axis=0))
    ref = ref.reshape(n_res * 5, 3)
    idx = np.arange(n_res) * 5
    sel = {"name N and not resname PRO": idx, "name H": idx + 1, "name CA": idx + 2,
           "custom occupancy": np.sort(np.concatenate((idx, idx + 2, idx + 3)))}
    return xyz.astype(np.float32), sel, ref.astype(np.float32)
