"""GPU parity of the C(t) path (K2 pack, K1 lag sums, Palmer finalize) through the C ABI.

Tolerances (north star): C(t) within 1e-6 relative of the reference evaluated on float64-upcast inputs;
dC(t) gets the same bound on small cases and an absolute floor tied to C(t)'s 1e-6 where the spread
across chunks is tiny."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ct_oracle

pytestmark = pytest.mark.gpu

RTOL_CT = 1e-6


def _gpu_ct(v4):
    from spinrelax_b200 import ct
    return ct.calculate_Ct_Palmer(v4)


def test_ct_small_vs_golden_and_oracle(golden):
    g = golden("ct_small.npz")
    Ct, dCt = _gpu_ct(g["vecs"])
    assert Ct.dtype == np.float32 and Ct.shape == g["Ct64"].shape
    assert rel_err(Ct, g["Ct64"]) < RTOL_CT
    assert rel_err(dCt, g["dCt64"]) < 1e-5
    oCt, odCt = ct_oracle.ct_palmer(g["vecs"].astype(np.float64))
    assert rel_err(Ct, oCt) < RTOL_CT
    # the float32 reference itself is further from float64 than we are
    assert rel_err(Ct, g["Ct64"]) <= rel_err(g["Ct32"], g["Ct64"]) + 1e-7


def test_ct_config1_vs_golden(golden):
    from spinrelax_b200 import synth
    g = golden("ct_config1.npz")
    v = synth.nh_vectors(10000, 76, seed=int(g["seed"]))
    v4 = v.reshape(10, 1000, 76, 3)
    Ct, dCt = _gpu_ct(v4)
    assert Ct.shape == (500, 76)
    S = ct_oracle.ct_lag_sums_fft(v4)
    oCt, odCt = ct_oracle.ct_from_lag_sums(S, 1000, dtype=np.float64)
    assert rel_err(Ct, oCt) < RTOL_CT
    assert np.max(np.abs(dCt - odCt)) < 1e-6 * np.max(np.abs(oCt))
    if np.array_equal(v[:4], g["input_head"]):
        lags = g["lags"]
        assert rel_err(Ct[lags - 1], g["Ct64"]) < RTOL_CT


def test_ct_lag_sums_raw(golden):
    import torch
    from spinrelax_b200 import ct
    g = golden("ct_small.npz")
    v = g["vecs"]
    S = ct.ct_lag_sums_device(torch.from_numpy(v).cuda()).cpu().numpy()
    assert rel_err(S, ct_oracle.ct_lag_sums_fft(v)) < 2e-7


@pytest.mark.parametrize("shape", [(1, 2, 1), (1, 3, 2), (2, 51, 1), (3, 480, 4), (2, 961, 3), (2, 1921, 2),
                                   (1, 2883, 1), (7, 100, 33), (2, 64, 130)])
def test_ct_ragged_shapes(shape):
    """Tile-boundary shapes: nF around the 960-frame tile and 480-lag tile, odd nF, tiny inputs."""
    from spinrelax_b200 import synth
    nC, nF, nR = shape
    v4 = synth.nh_vectors(nC * nF, nR, seed=100 + nF).reshape(nC, nF, nR, 3)
    with np.errstate(all="ignore"):
        Ct, dCt = _gpu_ct(v4)
        oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
    assert Ct.shape == (nF // 2, nR)
    assert rel_err(Ct, oCt) < RTOL_CT
    if nC > 1:
        assert np.max(np.abs(dCt - odCt)) < 2e-6 * max(1e-3, float(np.max(np.abs(odCt))))
    else:
        assert np.all(np.isnan(dCt))          # G3: 0/0


def test_ct_known_answers(golden):
    g = golden("ct_edge.npz")
    Ct, dCt = _gpu_ct(g["static"])
    assert np.allclose(Ct, 1.0, atol=2e-7) and np.allclose(dCt, 0.0, atol=2e-7)
    # i.i.d. uniform vectors: C(t) -> 0 within a few sigma = 1/sqrt(5 n)
    rng = np.random.default_rng(3)
    v = rng.standard_normal((4, 4000, 3, 3))
    v = (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)
    Ct, _ = _gpu_ct(v)
    assert np.max(np.abs(Ct)) < 6.0 / np.sqrt(5 * 4 * 2000)
    # rotation invariance of C(t): pack with a PAF quaternion gives the same lag sums
    import torch
    from spinrelax_b200 import _lib
    import ctypes
    lib = _lib.load()
    nC, nF, nR = 2, 700, 3
    vt = torch.from_numpy(v[:nC, :nF].copy()).cuda()
    pitch = lib.sr_ct_row_pitch(nF)
    outs = []
    for q in (None, (ctypes.c_double * 4)(0.83, -0.31, 0.22, 0.41)):
        packed = torch.empty((nR, nC, 3, pitch), dtype=torch.float32, device="cuda")
        S = torch.empty((nR, nC, nF // 2), dtype=torch.float64, device="cuda")
        _lib.check(lib.sr_pack_vectors_f32(vt.data_ptr(), nC, nF, nR, q, packed.data_ptr(), pitch, None))
        _lib.check(lib.sr_ct_lag_sums(packed.data_ptr(), pitch, nC, nF, nR, nF // 2, S.data_ptr(), None))
        outs.append(S.cpu().numpy())
    assert rel_err(outs[1], outs[0]) < 5e-6


def test_ct_full_size_properties():
    """BASELINE config-2 sized rows (nF = 2e5, L = 1e5) on 2 vectors: FFT oracle on every lag."""
    import torch
    from spinrelax_b200 import ct, synth
    nC, nF, nR = 2, 200000, 2
    v4 = synth.nh_vectors(nC * nF, nR, seed=77).reshape(nC, nF, nR, 3)
    S = ct.ct_lag_sums_device(torch.from_numpy(v4).cuda()).cpu().numpy()
    So = ct_oracle.ct_lag_sums_fft(v4)
    assert S.shape == (nR, nC, 100000)
    assert rel_err(S, So) < 1e-7
    Ct, dCt = ct.ct_palmer_device(torch.from_numpy(v4).cuda())
    oCt, _ = ct_oracle.ct_from_lag_sums(So, nF, dtype=np.float64)
    assert rel_err(Ct.cpu().numpy(), oCt) < RTOL_CT


def test_error_codes():
    from spinrelax_b200 import _lib
    lib = _lib.load()
    assert lib.sr_ct_lag_sums(None, 100, 1, 10, 1, 5, None, None) == -1
    assert b"null" in lib.sr_last_error()
    import torch
    x = torch.zeros(16, device="cuda")
    assert lib.sr_ct_palmer_device(x.data_ptr(), 1, 100, 1, x.data_ptr(), x.data_ptr(), x.data_ptr(), 8, None) == -3
    # chunk-range entry points: range outside [0, nC), misaligned packed stream
    big = torch.zeros(1 << 16, device="cuda")
    pitch = lib.sr_ct_row_pitch(100)
    assert lib.sr_pack_vectors_f32_chunks(big.data_ptr(), 2, 1, 2, 100, 1, None, big.data_ptr(), pitch, None) == -1
    assert b"chunk range" in lib.sr_last_error()
    assert lib.sr_ct_lag_sums_chunks(big.data_ptr() + 4, pitch, 2, 0, 1, 100, 1, 50, big.data_ptr(), None) == -1
    assert b"aligned" in lib.sr_last_error()
    assert lib.sr_ct_lag_sums_chunks(big.data_ptr(), pitch, 2, 2, 1, 100, 1, 50, big.data_ptr(), None) == -1
    # front end, dq extras
    idx = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert lib.sr_xh_vectors(big.data_ptr(), 10, 5, idx.data_ptr(), idx.data_ptr(), 0, big.data_ptr(), None) == -1
    assert lib.sr_xh_vectors_superposed(big.data_ptr(), 10, 5, idx.data_ptr(), big.data_ptr(), 2, idx.data_ptr(),
                                        idx.data_ptr(), 4, big.data_ptr(), None, None) == -1
    assert b"3 fit atoms" in lib.sr_last_error()
    lags = torch.ones(1, dtype=torch.int64, device="cuda")
    assert lib.sr_dq_moments_pooled(big.data_ptr(), 100, lags.data_ptr(), 1, 1, 1, 3, 3, 0, big.data_ptr(), None) == -1
    assert b"replica" in lib.sr_last_error()
    assert lib.sr_dq_hist3d(big.data_ptr(), 100, 1, big.data_ptr(), 4096, big.data_ptr(), big.data_ptr(), 8, big.data_ptr(),
                            None) == -1
    from spinrelax_b200 import traj
    with pytest.raises(_lib.SpinRelaxError):            # atom index outside the trajectory
        traj.xh_vectors_device(torch.zeros((3, 5, 3), device="cuda"), [7], [0])


def test_s2_and_average_vector(golden):
    """Next-tier rows fused on the same stream: S2 by outer product (:96-145) and --vecAvg (:579-583)."""
    from spinrelax_b200 import ct
    g = golden("s2.npz")
    v = g["vecs"]
    s2 = ct.calculate_S2_by_outerProduct(v, 10.0, 10000.0)
    assert s2.shape == g["s2_blocks"].shape
    assert rel_err(s2[:, 0], g["s2_blocks"][:, 0]) < 1e-6
    assert np.max(np.abs(s2[:, 1] - g["s2_blocks"][:, 1])) < 1e-6 * np.max(g["s2_blocks"][:, 0])
    s2all = ct.calculate_S2_by_outerProduct(v)
    assert rel_err(s2all, ct_oracle.s2_outer_product(v.astype(np.float64))) < 1e-6
    q = np.array([0.83, -0.31, 0.22, 0.41])
    avg = ct.average_vectors(v, q)
    ref = ct_oracle.average_vector(ct_oracle.rotate_vectors(v, q))
    assert np.max(np.abs(avg - ref)) < 1e-6


def test_cli_ct_end_to_end(tmp_path, golden):
    """The CLI mirror on .npy vector trajectories writes the reference's files; contents against the oracle."""
    import contextlib, io
    from spinrelax_b200 import cli_ct, io_formats, synth
    v1 = synth.nh_vectors(2300, 5, seed=41)
    v2 = synth.nh_vectors(1700, 5, seed=42)
    np.save(tmp_path / "a.npy", v1)
    np.save(tmp_path / "b.npy", v2)
    pref = str(tmp_path / "rotdif")
    with contextlib.redirect_stdout(io.StringIO()):
        cli_ct.main(["-s", "ref.pdb", "-f", str(tmp_path / "a.npy"), str(tmp_path / "b.npy"), "--dt", "10", "--tau", "5000",
                     "-o", pref, "--vecRot", "0.8 -0.36 0.48 0.0", "--vecHist", "--binary", "--vecAvg", "--S2", "--Ct"])
    v4 = ct_oracle.reformat_by_tau([v1, v2], 10.0, 5000.0)
    assert v4.shape == (7, 500, 5, 3)
    legs, x, y, dy = io_formats.load_sxydylist(pref + "_Ctint.dat")
    oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
    assert legs == ["1", "2", "3", "4", "5"] and np.array_equal(x[0], ct_oracle.ct_time_axis(10.0, 5000.0))
    assert rel_err(y, oCt.T) < 2e-6 and np.max(np.abs(dy - odCt.T)) < 1e-6
    q = np.array([0.8, -0.36, 0.48, 0.0])
    z = np.load(pref + "_vecHistogram.npz", allow_pickle=True)
    ho, eo = ct_oracle.sphere_histogram(v4.reshape(-1, 5, 3), q)
    assert str(z["dataType"]) == "LambertCylindrical" and bool(z["bHistogram"])
    assert np.array_equal(z["data"], ho) and np.array_equal(z["edges"][0], eo[0])
    # the reference's reader accepts the file (spectral_densities.py:279-306 semantics)
    from spinrelax_b200 import specdens
    rot = specdens.globalRotationalDiffusion_Axisymmetric(D=[2e-5, 1.2])
    rot.import_frame_vectors(pref + "_vecHistogram.npz")
    assert rot.vecXH.shape == (2592, 5, 3) and rot.vecWeights.shape == (2592, 5)
    s2 = np.loadtxt(pref + "_S2.dat", comments="&")
    ref = ct_oracle.s2_outer_product(v4.reshape(-1, 5, 3).astype(np.float64), 10.0, 5000.0) * (1.02 / 1.04) ** 6
    assert np.allclose(s2[:, 1:], ref, rtol=2e-5, atol=1e-7)
    with pytest.raises(SystemExit) as e:
        cli_ct.main(["-f", str(tmp_path / "a.npy"), "--Ct"])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:
        cli_ct.main(["-f", str(tmp_path / "a.npy"), "--vecRot", "1 1 0 0"])
    assert e.value.code == 23


def test_pipeline_host_path_matches_device_path():
    """CtHistStep.run_host (per-chunk H2D pipelined with K2/K1 through the *_chunks entry points) must give
    exactly what the whole-array device path gives, and both must match the oracle."""
    import torch
    from spinrelax_b200 import pipeline, synth
    nC, nF, nR = 3, 9000, 8
    q = (0.83, -0.31, 0.22, 0.41)
    v4 = synth.nh_vectors(nC * nF, nR, seed=77).reshape(nC, nF, nR, 3)
    step = pipeline.CtHistStep(nC, nF, nR, q_rot=q)
    Ct_d, dCt_d, hist_d = step.run_device(torch.from_numpy(v4).cuda())
    Ct_d, dCt_d = Ct_d.cpu().numpy().copy(), dCt_d.cpu().numpy().copy()
    for _ in range(2):                                   # second call reuses the staging buffer and side stream
        Ct_h, dCt_h, hist_h = step.run_host(v4)
    assert np.array_equal(Ct_h, Ct_d) and np.array_equal(dCt_h, dCt_d, equal_nan=True)
    assert np.array_equal(hist_h, hist_d)
    oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
    assert rel_err(Ct_h, oCt) < RTOL_CT
    ohist, _ = ct_oracle.sphere_histogram(v4.reshape(nC * nF, nR, 3), np.array(q))
    assert np.array_equal(hist_h, ohist)


def test_wide_vector_set_config4_shape():
    """BASELINE config 4 is wide (1000 N-H + 1000 C-H vectors): 2000 vectors, short chunks -- every vector group,
    the partial last groups of K2 / K3 and the per-vector rows of K1 against the oracle."""
    import torch
    from spinrelax_b200 import hist, synth
    nC, nF, nR = 2, 300, 2000
    v4 = synth.nh_vectors(nC * nF, nR, seed=2000).reshape(nC, nF, nR, 3)
    Ct, dCt = _gpu_ct(v4)
    oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
    assert rel_err(Ct, oCt) < RTOL_CT
    assert np.max(np.abs(dCt - odCt)) < 2e-6 * max(1e-3, float(np.max(np.abs(odCt))))
    q = np.array([0.83, -0.31, 0.22, 0.41])
    for n_sub in (2000, 1999, 1001):                      # nR % 4 != 0 takes the scalar-load kernels
        sub = np.ascontiguousarray(v4.reshape(nC * nF, nR, 3)[:, :n_sub])
        h, _ = hist.sphere_histogram(sub, q)
        ho, _ = ct_oracle.sphere_histogram(sub, q)
        assert np.array_equal(h.astype(np.int64), ho.astype(np.int64))
    sub = np.ascontiguousarray(v4[:, :, :1001])
    Ct2, _ = _gpu_ct(sub)
    assert np.array_equal(Ct2, Ct[:, :1001])              # rows are independent: same bits whatever the set
