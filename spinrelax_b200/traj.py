"""X-H bond vectors straight from Cartesian trajectories (the step in front of the C(t) hot path).

Mirrors obtain_XHvecs (calculate-Ct-from-traj.py:64-86) and the centre + superpose step the reference delegates
to mdtraj (:466-469) on the GPU (csrc/traj.cu).  mdtraj itself -- file formats, atom-selection language -- stays
out of scope; `ArrayTrajectory` is the minimal stand-in (`.xyz`, `.n_frames`, `.timestep`, `.topology.select`)
that lets the reference-shaped functions be called on coordinate arrays with named index selections.
"""
import sys

import numpy as np

from . import _lib


class _ArrayTopology:
    def __init__(self, selections, n_atoms):
        self._sel = {k: np.asarray(v, dtype=np.int64) for k, v in selections.items()}
        self.n_atoms = n_atoms

    def select(self, seltxt):
        """Named selections only: the text is a key of the `selections` dict (unknown text selects nothing,
        which the callers report the way the reference does for an empty mdtraj selection)."""
        return self._sel.get(seltxt, np.zeros(0, dtype=np.int64))


class ArrayTrajectory:
    """Coordinates (frames, atoms, 3) float32 in nm plus named atom-index selections."""

    def __init__(self, xyz, selections, timestep=1.0):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        if self.xyz.ndim != 3 or self.xyz.shape[-1] != 3:
            raise ValueError("ArrayTrajectory: xyz must be (frames, atoms, 3)")
        self.n_frames, self.n_atoms = self.xyz.shape[:2]
        self.timestep = float(timestep)
        self.topology = _ArrayTopology(selections, self.n_atoms)


def _device_indices(torch, idx, n_atoms, what):
    idx = np.asarray(idx, dtype=np.int64).ravel()
    if idx.size and (idx.min() < 0 or idx.max() >= n_atoms):
        raise _lib.SpinRelaxError("%s: atom index outside [0, %d)" % (what, n_atoms))
    return torch.from_numpy(idx.astype(np.int32)).cuda()


def xh_vectors_device(xyz_dev, index_h, index_x):
    """(frames, atoms, 3) float32 CUDA tensor -> (frames, nR, 3) unit vectors (CUDA)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if not (xyz_dev.is_cuda and xyz_dev.dtype == torch.float32 and xyz_dev.is_contiguous() and xyz_dev.dim() == 3):
        raise _lib.SpinRelaxError("xh_vectors_device: need a contiguous float32 CUDA (frames, atoms, 3) tensor")
    nF, nA, _ = xyz_dev.shape
    ih, ix = _device_indices(torch, index_h, nA, "indexH"), _device_indices(torch, index_x, nA, "indexX")
    if ih.numel() != ix.numel():
        raise _lib.SpinRelaxError("xh_vectors_device: %d H atoms but %d X atoms" % (ih.numel(), ix.numel()))
    out = torch.empty((nF, ih.numel(), 3), dtype=torch.float32, device=xyz_dev.device)
    _lib.check(lib.sr_xh_vectors(xyz_dev.data_ptr(), nF, nA, ih.data_ptr(), ix.data_ptr(), ih.numel(), out.data_ptr(),
                                 _lib.current_stream_ptr()), "sr_xh_vectors")
    return out


def xh_vectors_superposed_device(xyz_dev, ref_xyz, fit_indices, index_h, index_x, return_rotations=False):
    """Bond vectors of every frame after least-squares superposition of its fit atoms onto `ref_xyz` (atoms, 3)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    nF, nA, _ = xyz_dev.shape
    fit = np.asarray(fit_indices, dtype=np.int64).ravel()
    ref = np.asarray(ref_xyz, dtype=np.float64)
    if ref.shape != (nA, 3):
        raise _lib.SpinRelaxError("superpose: reference has shape %s, trajectory has %d atoms" % (ref.shape, nA))
    ih, ix = _device_indices(torch, index_h, nA, "indexH"), _device_indices(torch, index_x, nA, "indexX")
    fi = _device_indices(torch, fit, nA, "fit_indices")
    ref_fit = ref[fit] - ref[fit].mean(axis=0, keepdims=True)
    ref_dev = torch.from_numpy(np.ascontiguousarray(ref_fit)).cuda()
    out = torch.empty((nF, ih.numel(), 3), dtype=torch.float32, device=xyz_dev.device)
    rot = torch.empty((nF, 3, 3), dtype=torch.float64, device=xyz_dev.device) if return_rotations else None
    _lib.check(lib.sr_xh_vectors_superposed(xyz_dev.data_ptr(), nF, nA, fi.data_ptr(), ref_dev.data_ptr(), fi.numel(),
                                            ih.data_ptr(), ix.data_ptr(), ih.numel(), out.data_ptr(),
                                            rot.data_ptr() if rot is not None else None, _lib.current_stream_ptr()),
               "sr_xh_vectors_superposed")
    return (out, rot) if return_rotations else out


def obtain_XHvecs(traj, Hseltxt, Xseltxt, bSuppressPrint=False):
    """Drop-in for obtain_XHvecs (calculate-Ct-from-traj.py:64-86): same prints, same exits, NumPy in/out."""
    torch = _lib.require_cuda()
    if not bSuppressPrint:
        print("= = = Obtaining XH-vectors from trajectory...")
    indexX = traj.topology.select(Xseltxt)
    indexH = traj.topology.select(Hseltxt)
    numX, numH = len(indexX), len(indexH)
    if numX == 0 or numH == 0:
        print("= = = ERROR: selection text failed to find atoms!")
        print("     ....debug: N(%s) = %i , N(%s) = %i" % (Xseltxt, numX, Hseltxt, numH))
        sys.exit(1)
    if numH != numX:
        print("= = = ERROR: selection text found different number of atoms!")
        print("     ....debug: N(%s) = %i , N(%s) = %i" % (Xseltxt, numX, Hseltxt, numH))
        sys.exit(1)
    xyz = torch.from_numpy(np.ascontiguousarray(traj.xyz, dtype=np.float32)).cuda()
    return xh_vectors_device(xyz, indexH, indexX).cpu().numpy()


def obtain_XHvecs_fitted(traj, ref, fit_indices, Hseltxt, Xseltxt):
    """What the reference gets from `trj.center_coordinates(); trj.superpose(ref, frame=0, atom_indices=fit_indices);
    obtain_XHvecs(trj, ...)` (:466-469), without rewriting traj.xyz."""
    torch = _lib.require_cuda()
    indexX = traj.topology.select(Xseltxt)
    indexH = traj.topology.select(Hseltxt)
    if len(indexX) == 0 or len(indexH) == 0 or len(indexX) != len(indexH):
        print("= = = ERROR: selection text failed to find atoms!")
        sys.exit(1)
    ref_xyz = ref.xyz[0] if hasattr(ref, "xyz") else np.asarray(ref)
    xyz = torch.from_numpy(np.ascontiguousarray(traj.xyz, dtype=np.float32)).cuda()
    return xh_vectors_superposed_device(xyz, ref_xyz, fit_indices, indexH, indexX).cpu().numpy()
