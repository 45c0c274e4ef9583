"""PAF rotation + Lambert-cylindrical histogram of bond vectors (the `--vecRot q --vecHist` block of
calculate-Ct-from-traj.py:535-630) on the GPU, with bit-exact counts.

The kernel (csrc/hist.cu) bins every sample that is provably farther than a tolerance from all bin
edges.  Samples inside the tolerance band (expected ~6e-10 of the float64/rotated stream, ~1e-4 of the
float32/unrotated stream) have a bin that depends on the last ulp of NumPy's arctan2/arccos/cos; for
those -- and only those -- the host applies the reference's own formula (:567, :588, :613, :618) so that
the counts are identical to np.histogramdd's by construction.
"""
import ctypes

import numpy as np

from . import _lib

# margins: float64 path = error of our FP64 cross-product test + NumPy libm error (~1e-15) with headroom;
# float32 path = a few float32 ulps of phi (<= pi) and cos(theta) (<= 1) as computed by NumPy in float32
TOL_F64 = (1e-11, 1e-11)
TOL_F32 = (4e-6, 2e-6)


def _rotate(v, q):
    """rotate_vector_simd semantics (transforms3d_supplement.py:270-296) for a handful of rows."""
    q = np.asarray(q, dtype=np.float64)
    q = np.nan_to_num(q / np.linalg.norm(q))
    a = np.cross(q[1:4], v) + q[0] * v
    b = np.cross(q[1:4], a)
    return b + b + v


def _reference_bins(v, q_rot, edges_phi, edges_cos):
    """Exact NumPy evaluation for the tie-break samples; returns (flat_bin or -1) per row."""
    with np.errstate(all="ignore"):
        w = _rotate(v, q_rot) if q_rot is not None else v
        r = np.linalg.norm(w, axis=-1)
        phi = np.arctan2(w[..., 1], w[..., 0])
        cth = np.cos(np.arccos(w[..., 2] / r))
    out = np.full(len(v), -1, dtype=np.int64)
    nbx, nby = len(edges_phi) - 1, len(edges_cos) - 1
    # np.histogramdd: searchsorted(side='right'), samples equal to the last edge go to the last bin
    ix = np.searchsorted(edges_phi, phi, side="right")
    ix[phi == edges_phi[-1]] -= 1
    iy = np.searchsorted(edges_cos, cth, side="right")
    iy[cth == edges_cos[-1]] -= 1
    ok = (ix >= 1) & (ix <= nbx) & (iy >= 1) & (iy <= nby) & ~np.isnan(phi) & ~np.isnan(cth)
    out[ok] = (ix[ok] - 1) * nby + (iy[ok] - 1)
    return out


class SphereHistogram:
    """Device-resident accumulator: counts (nR, nbx, nby) uint32 plus the ambiguous-sample list."""

    def __init__(self, nR, nbx=72, device=None, amb_capacity=1 << 22):
        torch = _lib.require_cuda()
        self.torch = torch
        self.lib = _lib.load()
        self.nR, self.nbx, self.nby = nR, int(nbx), int(nbx / 2)
        self.dev = device if device is not None else torch.device("cuda")
        # the edges np.histogramdd builds for range=((-pi,pi),(-1,1))
        self.edges_phi = np.linspace(-np.pi, np.pi, self.nbx + 1)
        self.edges_cos = np.linspace(-1.0, 1.0, self.nby + 1)
        table = np.concatenate((np.stack((np.cos(self.edges_phi), np.sin(self.edges_phi)), axis=1).ravel(),
                                self.edges_cos))
        self.table = torch.from_numpy(table).to(self.dev)
        self.counts = torch.zeros((nR, self.nbx, self.nby), dtype=torch.int32, device=self.dev)
        self.amb_idx = torch.empty(amb_capacity, dtype=torch.int64, device=self.dev)
        self.amb_count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.amb_capacity = amb_capacity

    def accumulate_device(self, v_dev, q_rot, reset=True):
        """v_dev: CUDA float32 (frames, nR, 3). Adds to counts; ambiguous sample ids go to amb_idx.
        With reset=False the counts of several arrays are accumulated: call resolve_pending(v_dev, q_rot) (or
        finish) for each array before launching the next one, because the ambiguous ids index into v_dev."""
        torch = self.torch
        if not (v_dev.is_cuda and v_dev.dtype == torch.float32 and v_dev.is_contiguous() and v_dev.shape[1] == self.nR):
            raise _lib.SpinRelaxError("SphereHistogram: need contiguous float32 CUDA (frames, %d, 3)" % self.nR)
        if reset:
            self.counts.zero_()
            self.amb_count.zero_()
            self._host_extra = None
            self.last_ambiguous = 0
        tol = TOL_F64 if q_rot is not None else TOL_F32
        q = None if q_rot is None else (ctypes.c_double * 4)(*[float(x) for x in q_rot])
        rc = self.lib.sr_sphere_hist(v_dev.data_ptr(), v_dev.shape[0], self.nR, q, self.nbx, self.nby,
                                     self.table.data_ptr(), tol[0], tol[1], self.counts.data_ptr(),
                                     self.amb_idx.data_ptr(), self.amb_capacity, self.amb_count.data_ptr(),
                                     _lib.current_stream_ptr())
        _lib.check(rc, "sr_sphere_hist")

    def resolve_pending(self, v_dev, q_rot):
        """Bin the samples the device left undecided (ids >= 0 in the list) with the reference's own NumPy formula;
        the result is kept on the host and the device list is emptied."""
        n_amb = int(self.amb_count.item())
        if n_amb > self.amb_capacity:
            raise _lib.SpinRelaxError("SphereHistogram: %d retry / ambiguous samples exceed capacity %d"
                                      % (n_amb, self.amb_capacity))
        if getattr(self, "_host_extra", None) is None:
            self._host_extra = np.zeros((self.nR, self.nbx * self.nby), dtype=np.int64)
        if n_amb:
            idx = self.amb_idx[:n_amb]
            idx = idx[idx >= 0]          # -1 = resolved on the device by sphere_hist_resolve_kernel
            self.last_ambiguous = getattr(self, "last_ambiguous", 0) + int(idx.numel())
            if idx.numel():
                rows = v_dev.reshape(-1, 3)[idx].cpu().numpy()
                idx = idx.cpu().numpy()
                bins = _reference_bins(rows, q_rot, self.edges_phi, self.edges_cos)
                ok = bins >= 0
                np.add.at(self._host_extra, (idx[ok] % self.nR, bins[ok]), 1)
            self.amb_count.zero_()

    def finish(self, v_dev, q_rot):
        """Resolve the ambiguous samples of the last array and return int64 counts (NumPy)."""
        self.resolve_pending(v_dev, q_rot)
        counts = self.counts.cpu().numpy().astype(np.int64)
        return counts + self._host_extra.reshape(counts.shape)


def sphere_histogram(frames_vecs, q_rot=None, nbins_phi=72):
    """Drop-in for the histogram block of calculate-Ct-from-traj.py:567-626.

    frames_vecs: (frames, nR, 3) float32 NumPy array (host).  Returns (hist_list, edges) exactly as the
    reference builds them: hist_list (nR, nbx, nby) holding integer counts in float64 (rotated path, :567
    promotes) or float32 (no rotation), edges = [phi edges, cos(theta) edges].
    """
    torch = _lib.require_cuda()
    from . import multigpu, resident
    v = np.asarray(frames_vecs, dtype=np.float32)
    if v.ndim != 3 or v.shape[-1] != 3:
        raise ValueError("sphere_histogram: expected (frames, nR, 3)")

    def work(dev, a, b):        # every vector has its own histogram: shard the vectors, no reduction
        acc = SphereHistogram(b - a, nbins_phi, device=torch.device("cuda", dev))
        v_dev = resident.device_block(v, dev, a, b)
        acc.accumulate_device(v_dev, q_rot)
        return acc.finish(v_dev, q_rot), acc.edges_phi, acc.edges_cos

    res = multigpu.run(multigpu.plan(v.shape[1]), work)
    counts = np.concatenate([r[0] for r in res], axis=0)
    dtype = np.float64 if q_rot is not None else np.float32
    return counts.astype(dtype), [res[0][1], res[0][2]]


def save_vec_histogram(path, names, hist_list, edges):
    """np.savez_compressed(out+'_vecHistogram.npz', ...) of calculate-Ct-from-traj.py:629-630; `edges` is a
    ragged 2-list and must be stored as an object array on NumPy >= 1.24 (reader: spectral_densities.py:285)."""
    e = np.empty(2, dtype=object)
    e[0], e[1] = edges[0], edges[1]
    np.savez_compressed(path, names=names, dataType="LambertCylindrical", bHistogram=True, edges=e,
                        axisLabels=["phi", "cos(theta)"], data=hist_list)
