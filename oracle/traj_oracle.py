"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the trajectory front end of calculate-Ct-from-traj.py.

* xh_vectors follows obtain_XHvecs (calculate-Ct-from-traj.py:83-84) and qs.vecnorm_NDarray
  (transforms3d_supplement.py:40-52) line by line; it is pinned by tests/golden/traj.npz, which
  tests/golden/make_golden.py generates by calling the REAL obtain_XHvecs on a stand-in trajectory object.
* superposition: the reference delegates to the third-party mdtraj (`trj.center_coordinates();
  trj.superpose(ref, frame=0, atom_indices=fit_indices)`, :466-467; requirements.txt pins no version and mdtraj
  is absent from this image), whose published algorithm is the least-squares superposition of the fit atoms
  (Theobald's QCP, a proper rotation).  PARITY UNPINNED at that boundary: the restatement below is the textbook
  Kabsch solution in float64 (SVD with the reflection fix), which has the same optimum.
"""
import numpy as np


def vecnorm_ndarray(v, axis=-1):
    """transforms3d_supplement.py:40-52"""
    sh = list(v.shape)
    sh[axis] = 1
    with np.errstate(all="ignore"):
        return np.nan_to_num(v / np.linalg.norm(v, axis=axis).reshape(sh))


def xh_vectors(xyz, index_h, index_x):
    """calculate-Ct-from-traj.py:83-84"""
    vec = np.take(xyz, index_h, axis=1) - np.take(xyz, index_x, axis=1)
    return vecnorm_ndarray(vec, axis=2)


def kabsch_rotations(xyz, ref_xyz, fit_indices):
    """Per frame the proper rotation R minimising sum_i |R (x_i - <x>) - (y_i - <y>)|^2 over the fit atoms."""
    x = np.asarray(xyz, dtype=np.float64)[:, fit_indices]
    y = np.asarray(ref_xyz, dtype=np.float64)[fit_indices]
    x = x - x.mean(axis=1, keepdims=True)
    y = y - y.mean(axis=0, keepdims=True)
    rots = np.empty((len(x), 3, 3))
    for f in range(len(x)):
        h = x[f].T @ y                       # sum_i x_i y_i^T
        u, _, vt = np.linalg.svd(h)
        d = np.sign(np.linalg.det(vt.T @ u.T))
        rots[f] = vt.T @ np.diag([1.0, 1.0, d]) @ u.T
    return rots


def xh_vectors_superposed(xyz, ref_xyz, fit_indices, index_h, index_x):
    """obtain_XHvecs after center_coordinates + superpose (:466-469): rotate every frame, then :83-84.
    The rotation is applied to the float32 difference vectors in float64 and rounded once (mdtraj rotates the
    float32 coordinates themselves; the two differ at the 1e-7 level, inside the float32 noise of the input)."""
    rots = kabsch_rotations(xyz, ref_xyz, fit_indices)
    d = (np.take(xyz, index_h, axis=1) - np.take(xyz, index_x, axis=1)).astype(np.float64)
    w = np.einsum("fab,frb->fra", rots, d).astype(np.float32)
    return vecnorm_ndarray(w, axis=2), rots
