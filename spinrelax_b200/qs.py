"""Host mirror of the quaternion helpers of transforms3d_supplement.py that sit on the hot path.

Bulk operations (`rotate_vector_simd` on float32 trajectories, `obtain_self_dq`-style products) run on
the GPU through the C ABI; the 3x3 / single-quaternion helpers that the reference calls once per lag
(`quat_frame_transform_min`, transforms3d_supplement.py:137-149) are O(1) host algebra.
The third-party `transforms3d.quaternions` functions the reference imports are not needed: the few
that matter (Hamilton product, rotate one vector, axis-angle) are written out here.
"""
import ctypes
import math

import numpy as np

from . import _lib


def vecnorm_NDarray(v, axis=-1):
    """Normalise along `axis`, mapping 0/0 to 0 (transforms3d_supplement.py:40-52)."""
    v = np.asarray(v)
    with np.errstate(all="ignore"):
        if v.ndim > 1:
            return np.nan_to_num(v / np.linalg.norm(v, axis=axis, keepdims=True))
        return np.nan_to_num(v / np.linalg.norm(v))


def quat_invert(q):
    """Conjugate (transforms3d_supplement.py:185-186); float32 input is promoted to float64 like the reference."""
    return np.asarray(q) * [1.0, -1.0, -1.0, -1.0]


def quat_mult_simd(q1, q2):
    """Hamilton product along the last axis (transforms3d_supplement.py:163-183)."""
    q1, q2 = np.asarray(q1), np.asarray(q2)
    w1, v1 = q1[..., 0], q1[..., 1:4]
    w2, v2 = q2[..., 0], q2[..., 1:4]
    out = np.zeros(np.broadcast(q1, q2).shape, dtype=np.result_type(q1, q2))
    out[..., 0] = w1 * w2 - np.sum(v1 * v2, axis=-1)
    out[..., 1:4] = w1[..., None] * v2 + w2[..., None] * v1 + np.cross(v1, v2)
    return out


def quat_reduce_simd(q, qref=(1, 0, 0, 0), axis=-1):
    """Image with q.qref >= 0 (transforms3d_supplement.py:219-234)."""
    q = np.asarray(q)
    if axis == -1:
        s = np.sign(q @ np.asarray(qref, dtype=float))
        s[s == 0] = 1.0
        return q * s[:, None]
    s = np.sign(np.tensordot(np.asarray(qref, dtype=float), q, axes=(0, 0)))
    s[s == 0] = 1.0
    return q * s[None, :]


def rotate_vector_simd(v, q, axis=-1, bNormalised=False):
    """qs.rotate_vector_simd (transforms3d_supplement.py:270-296).

    float32 (..., 3) arrays with a single quaternion -- the trajectory case of
    calculate-Ct-from-traj.py:567 -- go through sr_rotate_vectors_f32_f64 and come back float64,
    bit-identical to NumPy.  Anything else (float64 input, per-vector quaternions, axis=0) is small-scale
    use in the reference and is evaluated with the same formula in NumPy.
    """
    v = np.asarray(v)
    q = np.asarray(q, dtype=np.float64) if np.ndim(q) == 1 else np.asarray(q)
    if not bNormalised:
        q = vecnorm_NDarray(q)
    if axis == -1 and q.ndim == 1 and v.dtype == np.float32 and v.ndim >= 2 and v.shape[-1] == 3 and v.size >= 3 * 4096:
        torch = _lib.require_cuda()
        lib = _lib.load()
        vc = np.ascontiguousarray(v)
        vd = torch.from_numpy(vc).cuda()
        out = torch.empty(vc.shape, dtype=torch.float64, device=vd.device)
        qc = (ctypes.c_double * 4)(*[float(x) for x in q])
        _lib.check(lib.sr_rotate_vectors_f32_f64(vd.data_ptr(), vc.size // 3, qc, out.data_ptr(),
                                                 _lib.current_stream_ptr()), "sr_rotate_vectors_f32_f64")
        return out.cpu().numpy()
    if axis == -1:
        qw, qv = q[..., 0], q[..., 1:4]
        a = np.cross(qv, v) + (qw[..., None] if np.ndim(qw) else qw) * v
        b = np.cross(qv, a)
        return b + b + v
    if axis == 0:
        qw, qv = q[0, ...], q[1:4, ...]
        a = np.cross(qv, v, axisa=0, axisb=0, axisc=0) + qw[None, ...] * v
        b = np.cross(qv, a, axisa=0, axisb=0, axisc=0)
        return b + b + v
    raise ValueError("rotate_vector_simd: axis must be -1 or 0")


# ---- single-quaternion algebra used once per lag ----------------------------------------------------
def _qmult(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def _rotate_one(v, q):
    qc = np.array(q) * np.array([1.0, -1.0, -1.0, -1.0])
    return _qmult(q, _qmult(np.concatenate(([0.0], v)), qc))[1:]


def quat_v1v2(v1, v2):
    """Minimum-angle quaternion rotating v1 onto v2 (transforms3d_supplement.py:71-83)."""
    th = math.acos(np.dot(v1, v2))
    ax = np.cross(v1, v2)
    if all(np.isnan(ax)):
        return np.array([1.0, 0.0, 0.0, 0.0])
    ax = np.asarray(ax, dtype=float)
    ax = ax / math.sqrt(float(np.dot(ax, ax)))
    return np.concatenate(([math.cos(th / 2.0)], ax * math.sin(th / 2.0)))


def quat_frame_transform_min(axes):
    """transforms3d_supplement.py:137-149: rotate axes[2] onto +-Z then axes[0] onto +-X, larger q_w wins."""
    c1 = (quat_v1v2(axes[2], (0, 0, 1)), quat_v1v2(axes[2], (0, 0, -1)))
    q1 = c1[0] if c1[0][0] > c1[1][0] else c1[1]
    x_rot = _rotate_one(axes[0], q1)
    c2 = (quat_v1v2(x_rot, (1, 0, 0)), quat_v1v2(x_rot, (-1, 0, 0)))
    q2 = c2[0] if c2[0][0] > c2[1][0] else c2[1]
    return _qmult(q2, q1)


# ---- the same algebra over a stack of frames: one call for a whole lag list (calculate-dq-distribution.py:554-650
# evaluates it once per lag; with 1e5 lag windows a Python loop costs more than the GPU reduction it follows) ------
def _qmult_batch(a, b):
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2), axis=-1)


def _quat_v1v2_batch(v1, v2):
    """quat_v1v2 for v1 (n, 3) against one fixed unit vector v2: same operations element by element."""
    v2 = np.asarray(v2, dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        th = np.arccos(v1[:, 0] * v2[0] + v1[:, 1] * v2[1] + v1[:, 2] * v2[2])
        ax = np.stack((v1[:, 1] * v2[2] - v1[:, 2] * v2[1], v1[:, 2] * v2[0] - v1[:, 0] * v2[2],
                       v1[:, 0] * v2[1] - v1[:, 1] * v2[0]), axis=1)
        dead = np.all(np.isnan(ax), axis=1)
        ax = ax / np.sqrt(ax[:, 0] * ax[:, 0] + ax[:, 1] * ax[:, 1] + ax[:, 2] * ax[:, 2])[:, None]
        out = np.concatenate((np.cos(th / 2.0)[:, None], ax * np.sin(th / 2.0)[:, None]), axis=1)
    out[dead] = (1.0, 0.0, 0.0, 0.0)
    return out


def quat_frame_transform_min_batch(axes):
    """quat_frame_transform_min for axes (n, 3, 3) (rows = the three axes of every frame): (n, 4)."""
    axes = np.asarray(axes, dtype=float)
    a, b = _quat_v1v2_batch(axes[:, 2], (0, 0, 1)), _quat_v1v2_batch(axes[:, 2], (0, 0, -1))
    q1 = np.where((a[:, 0] > b[:, 0])[:, None], a, b)
    v = np.concatenate((np.zeros((len(axes), 1)), axes[:, 0]), axis=1)
    x_rot = _qmult_batch(q1, _qmult_batch(v, q1 * np.array([1.0, -1.0, -1.0, -1.0])))[:, 1:]
    a, b = _quat_v1v2_batch(x_rot, (1, 0, 0)), _quat_v1v2_batch(x_rot, (-1, 0, 0))
    q2 = np.where((a[:, 0] > b[:, 0])[:, None], a, b)
    return _qmult_batch(q2, q1)


def nearly_identity(q, rtol=1e-5, atol=1e-8):
    """transforms3d.quaternions.nearly_equivalent(q, (1,0,0,0)) as used at calculate-dq-distribution.py:124."""
    q = np.asarray(q, dtype=float)
    e = np.array([1.0, 0.0, 0.0, 0.0])
    return bool(np.allclose(q, e, rtol, atol) or np.allclose(-q, e, rtol, atol))


def rotation_matrix(q):
    """3x3 matrix R with R v == rotate_vector_simd(v, q) for the normalised q."""
    q = vecnorm_NDarray(np.asarray(q, dtype=float))
    return np.array([_rotate_one(e, q) for e in np.eye(3)]).T
