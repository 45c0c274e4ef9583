"""Minimal `transforms3d.quaternions` (published algorithm, w-x-y-z convention) for the call sites
SpinRelax uses: calculate-dq-distribution.py:124,403-405,566,619; transforms3d_supplement.py:80-83,
124-149,155,251.  Test infrastructure only."""
import math

import numpy as np


def qeye(dtype=np.float64):
    return np.array([1.0, 0.0, 0.0, 0.0], dtype=dtype)


def qmult(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def qconjugate(q):
    return np.array(q) * np.array([1.0, -1.0, -1.0, -1.0])


def qnorm(q):
    return math.sqrt(float(np.dot(q, q)))


def qisunit(q):
    return bool(np.allclose(qnorm(q), 1.0))


def qinverse(q):
    return qconjugate(q) / float(np.dot(q, q))


def rotate_vector(v, q):
    varr = np.zeros((4,))
    varr[1:] = v
    return qmult(q, qmult(varr, qconjugate(q)))[1:]


def nearly_equivalent(q1, q2, rtol=1e-5, atol=1e-8):
    q1 = np.array(q1)
    q2 = np.array(q2)
    if np.allclose(q1, q2, rtol, atol):
        return True
    return bool(np.allclose(q1 * -1, q2, rtol, atol))


def axangle2quat(vector, theta, is_normalized=False):
    vector = np.array(vector, dtype=float)
    if not is_normalized:
        vector = vector / math.sqrt(float(np.dot(vector, vector)))
    t2 = theta / 2.0
    st2 = math.sin(t2)
    return np.concatenate(([math.cos(t2)], vector * st2))


def mat2quat(M):
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, dtype=float).flat
    K = np.array([[Qxx - Qyy - Qzz, 0, 0, 0],
                  [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                  [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                  [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = q * -1
    return q
