"""GPU parity of the --vecDist (no --vecHist) branch: gm.xyz_to_rtp, the PhiTheta outputs of the Ct CLI mirror."""
import contextlib
import io

import numpy as np
import pytest

from oracle import ct_oracle

pytestmark = pytest.mark.gpu

ULP = 4          # CUDA atan2/acos are <= 2 ulp, glibc/SVML <= 1-4 ulp; r and z/r are bit-identical inputs


def _ulps(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    with np.errstate(all="ignore"):
        return np.max(np.abs(a[ok] - b[ok]) / np.spacing(np.abs(b[ok]))) if ok.any() else 0.0


def test_xyz_to_rtp_vs_reference(golden):
    from spinrelax_b200 import gm, qs
    g = golden("rtp.npz")
    r32 = gm.xyz_to_rtp(g["vecs"])
    assert np.array_equal(r32[..., 0], g["rtp32"][..., 0])          # |v|: bit-identical to np.linalg.norm
    assert _ulps(r32, g["rtp32"]) <= ULP
    w = qs.rotate_vector_simd(g["vecs"], g["q"])
    r64 = gm.xyz_to_rtp(w)
    assert np.array_equal(r64[..., 0], g["rtp64"][..., 0]) and _ulps(r64, g["rtp64"]) <= ULP
    # shipped bUnit form: theta = arccos(z / phi) -- a last-ulp difference in phi is amplified without bound where
    # |z / phi| -> 1, so theta is pinned through the same formula on the returned phi
    v64 = g["vecs"].astype(np.float64)
    u = gm.xyz_to_rtp(v64, bUnit=True)
    assert u.shape == g["unit64"].shape and _ulps(u[..., 0], g["unit64"][..., 0]) <= ULP
    with np.errstate(all="ignore"):
        assert _ulps(u[..., 1], np.arccos(v64[..., 2] / u[..., 0])) <= ULP
    with np.errstate(all="ignore"):
        far = np.abs(v64[..., 2] / g["unit64"][..., 0]) < 0.9
    assert np.allclose(u[..., 1][far], g["unit64"][..., 1][far], rtol=1e-14, atol=0)
    ax0 = gm.xyz_to_rtp(np.ascontiguousarray(np.moveaxis(g["vecs"][:9].astype(np.float64), -1, 0)), vaxis=0)
    assert _ulps(ax0, g["ax0"]) <= ULP
    assert _ulps(gm.xyz_to_rtp(g["vecs"][7, 2].astype(np.float64)), g["one"]) <= ULP
    assert np.allclose(gm.rtp_to_xyz(g["pt"], vaxis=-1, bUnit=True), g["back"], rtol=0, atol=1e-15)
    assert np.allclose(gm.rtp_to_xyz(np.array([1.3, 0.4, 2.1])), g["back1"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 100003])
def test_xyz_to_rtp_ragged_vs_oracle(n):
    from spinrelax_b200 import gm, synth
    v = synth.nh_vectors(n, 1, seed=n)[:, 0] * np.float32(2.5)
    for arr in (v, v.astype(np.float64)):
        got, ref = gm.xyz_to_rtp(arr), ct_oracle.xyz_to_rtp(arr)
        assert np.array_equal(got[..., 0], ref[..., 0]) and _ulps(got, ref) <= ULP


def test_cli_phitheta_outputs(golden, tmp_path):
    """`--vecDist --binary` -> _vecPhiTheta.npz with the reference's keys; text form starts with the one set the
    shipped print_s3d manages to write and continues in the same format."""
    from spinrelax_b200 import cli_ct
    g = golden("rtp.npz")
    names = np.array(["A", "B", "C", "D", "E"])
    np.savez(tmp_path / "v.npz", vecs=g["vecs"], names=names, dt=10.0)
    qtxt = " ".join(repr(float(x)) for x in g["q"])
    pref = str(tmp_path / "o")
    with contextlib.redirect_stdout(io.StringIO()):
        cli_ct.main(["-f", str(tmp_path / "v.npz"), "-o", pref, "--vecDist", "--binary", "--vecRot", qtxt])
    z = np.load(pref + "_vecPhiTheta.npz", allow_pickle=True)
    assert str(z["dataType"]) == "PhiTheta" and not bool(z["bHistogram"]) and list(z["axisLabels"]) == ["phi", "theta"]
    assert list(z["names"]) == list(names)
    assert z["data"].shape == (5, 700, 2) and _ulps(z["data"], np.transpose(g["rtp64"], (1, 0, 2))[..., 1:3]) <= ULP
    with contextlib.redirect_stdout(io.StringIO()):
        cli_ct.main(["-f", str(tmp_path / "v.npz"), "-o", pref + "32", "--vecDist", "--binary"])
    z32 = np.load(pref + "32_vecPhiTheta.npz", allow_pickle=True)
    assert z32["data"].dtype == np.float32
    assert _ulps(z32["data"], np.transpose(g["rtp32"], (1, 0, 2))[..., 1:3]) <= ULP
    with contextlib.redirect_stdout(io.StringIO()):
        cli_ct.main(["-f", str(tmp_path / "v.npz"), "-o", pref, "--vecDist", "--vecRot", qtxt])
    txt = open(pref + "_vecPhiTheta.dat").read()
    head = str(g["s3d_partial"])
    gl, hl = txt.splitlines(), head.splitlines()
    assert len(gl) == 5 * 702 and gl[0] == hl[0] and gl[701] == "&" and gl[702] == '@s1 legend "B"'
    a = np.array([l.split() for l in gl[1:701]], dtype=float)
    b = np.array([l.split() for l in hl[1:701]], dtype=float)
    assert np.allclose(a, b, rtol=2e-6, atol=1e-12)      # "%g" keeps 6 digits
    assert sum(x != y for x, y in zip(gl[:702], hl)) <= 2         # a last-ulp difference may flip a 6th digit
