"""GPU parity of K3 (rotation + spherical histogram): counts must equal np.histogramdd's bit for bit."""
import numpy as np
import pytest

from oracle import ct_oracle

pytestmark = pytest.mark.gpu


def _gpu_hist(v, q, nbx=72):
    from spinrelax_b200 import hist
    return hist.sphere_histogram(v, q, nbx)


def test_hist_golden_rotated(golden):
    g = golden("hist.npz")
    h, e = _gpu_hist(g["vecs_rot"], g["q"])
    assert h.dtype == np.float64 and h.shape == (7, 72, 36)
    assert np.array_equal(h.astype(np.int64), g["hist_rot"])
    assert np.array_equal(e[0], g["edges_phi"]) and np.array_equal(e[1], g["edges_cos"])
    h36, _ = _gpu_hist(g["vecs_rot36"], g["q"], 36)
    assert np.array_equal(h36.astype(np.int64), g["hist_rot36"])


def test_hist_golden_float32_path_and_special(golden):
    g = golden("hist.npz")
    h, _ = _gpu_hist(g["vecs_f32"], None)
    assert h.dtype == np.float32
    assert np.array_equal(h.astype(np.int64), g["hist_f32"])
    hs, _ = _gpu_hist(g["special"], None)       # +-z, +-x, +y, zero vector: on-edge and NaN semantics
    assert np.array_equal(hs.astype(np.int64), g["hist_special"])


@pytest.mark.parametrize("nR,frames,rot", [(1, 1, True), (3, 257, True), (76, 5000, True), (76, 5000, False),
                                            (9, 20000, True), (130, 999, True)])
def test_hist_vs_oracle_same_host(nR, frames, rot):
    """Unfiltered samples, oracle run on the same host (NumPy decides the ties on both sides)."""
    from spinrelax_b200 import synth
    v = synth.nh_vectors(frames, nR, seed=900 + nR)
    q = np.array([0.2, 0.5, -0.7, 0.1]) if rot else None
    h, _ = _gpu_hist(v, q)
    ho, _ = ct_oracle.sphere_histogram(v, q)
    assert np.array_equal(h.astype(np.int64), ho.astype(np.int64))
    assert h.sum() == frames * nR


@pytest.mark.parametrize("nbx", [10, 120, 144])
def test_hist_other_bin_counts(nbx):
    """Non-default --histBin: runtime bin constants, and 8 / 4 vectors per CTA when 16 histograms do not fit."""
    from spinrelax_b200 import hist, synth
    v = synth.nh_vectors(3000, 21, seed=nbx)
    for q in (np.array([0.2, 0.5, -0.7, 0.1]), None):
        with np.errstate(all="ignore"):
            h, e = hist.sphere_histogram(v, q, nbx)
            ho, eo = ct_oracle.sphere_histogram(v, q, nbins_phi=nbx)
        assert np.array_equal(h.astype(np.int64), ho.astype(np.int64)) and h.sum() == 3000 * 21
        assert np.array_equal(e[0], eo[0]) and np.array_equal(e[1], eo[1])


def test_hist_uniform_sphere_and_ties():
    """Uniform vectors touch every bin; axis-aligned and grid-aligned vectors sit exactly on edges."""
    rng = np.random.default_rng(5)
    v = rng.standard_normal((40000, 4, 3))
    v = (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)
    # exact edge directions: phi = k*5deg in the xy plane, cos(theta) = j/18
    k = np.arange(72)
    edge = np.stack((np.cos(np.deg2rad(5.0 * k - 180)), np.sin(np.deg2rad(5.0 * k - 180)), np.zeros(72)), axis=1)
    v[:72, 0] = edge.astype(np.float32)
    j = np.linspace(-1, 1, 37)
    v[100:137, 1] = np.stack((np.sqrt(1 - j * j), np.zeros(37), j), axis=1).astype(np.float32)
    for q in (None, np.array([1.0, 0, 0, 0]), np.array([0.83, -0.31, 0.22, 0.41])):
        with np.errstate(all="ignore"):
            h, _ = _gpu_hist(v, q)
            ho, _ = ct_oracle.sphere_histogram(v, q)
        assert np.array_equal(h.astype(np.int64), ho.astype(np.int64))
        assert (h > 0).all()


def test_hist_full_size_sum():
    """BASELINE config-2 sized stream (1e6 frames x 76): every unit vector is counted exactly once."""
    import torch
    from spinrelax_b200 import hist, synth
    v = synth.nh_vectors(1000000, 76, seed=77)
    acc = hist.SphereHistogram(76)
    vd = torch.from_numpy(v).cuda()
    q = np.array([0.83, -0.31, 0.22, 0.41])
    acc.accumulate_device(vd, q)
    counts = acc.finish(vd, q)
    assert counts.sum() == 76 * 1000000 and (counts.sum(axis=(1, 2)) == 1000000).all()
    assert acc.last_ambiguous < 100
    # a 1/50 slice against the oracle
    sl = v[::50]
    h, _ = hist.sphere_histogram(sl, q)
    ho, _ = ct_oracle.sphere_histogram(sl, q)
    assert np.array_equal(h.astype(np.int64), ho.astype(np.int64))


def test_hist_accumulates_over_arrays():
    """Counts of two trajectories accumulated call by call equal the histogram of their concatenation (the retry list
    of each call is resolved against that call's own array)."""
    import torch
    from spinrelax_b200 import hist, synth
    q = np.array([0.83, -0.31, 0.22, 0.41])
    a, b = synth.nh_vectors(30011, 12, seed=5), synth.nh_vectors(17003, 12, seed=6)
    acc = hist.SphereHistogram(12)
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    acc.accumulate_device(ad, q, reset=True)
    acc.resolve_pending(ad, q)
    acc.accumulate_device(bd, q, reset=False)
    both = acc.finish(bd, q)
    whole, _ = ct_oracle.sphere_histogram(np.concatenate((a, b)), q)
    assert np.array_equal(both, whole.astype(np.int64))
    for rot in (None,):                                   # float32 reference path: host tie-breaks per array
        acc.accumulate_device(ad, rot, reset=True)
        acc.resolve_pending(ad, rot)
        acc.accumulate_device(bd, rot, reset=False)
        both = acc.finish(bd, rot)
        whole, _ = ct_oracle.sphere_histogram(np.concatenate((a, b)), rot)
        assert np.array_equal(both, whole.astype(np.int64))


def test_hist_host_entry_point():
    """sr_sphere_hist_host (C ABI on pageable host buffers): its counts plus its short undecided list, binned with the
    reference formula, equal the oracle; the edges it builds itself are NumPy's."""
    import ctypes
    from spinrelax_b200 import _lib, hist, synth
    lib = _lib.load()
    for q in (np.array([0.83, -0.31, 0.22, 0.41]), None):
        v = synth.nh_vectors(20011, 12, seed=21)
        counts = np.zeros((12, 72, 36), dtype=np.uint32)
        cap = 1 << 16
        amb = np.zeros(cap, dtype=np.int64)
        n_amb = ctypes.c_int(0)
        qc = None if q is None else (ctypes.c_double * 4)(*q)
        rc = lib.sr_sphere_hist_host(v.ctypes.data, v.shape[0], 12, qc, 72, 36, counts.ctypes.data, amb.ctypes.data, cap,
                                     ctypes.byref(n_amb))
        _lib.check(rc, "sr_sphere_hist_host")
        total = counts.astype(np.int64).reshape(12, -1)
        idx = amb[:n_amb.value]
        if len(idx):
            bins = hist._reference_bins(v.reshape(-1, 3)[idx], q, np.linspace(-np.pi, np.pi, 73), np.linspace(-1, 1, 37))
            ok = bins >= 0
            np.add.at(total, (idx[ok] % 12, bins[ok]), 1)
        ho, _ = ct_oracle.sphere_histogram(v, q)
        assert np.array_equal(total.reshape(12, 72, 36), ho.astype(np.int64))
        assert n_amb.value < (50 if q is not None else 2000)
