"""Glue of the C(t) CLI mirror without a GPU: the device stages are replaced by the oracle's NumPy restatements and
`cli_ct.main` must write the files that calculate-Ct-from-traj.py's own functions and writers produce when chained
as its __main__ chains them (tests/golden/ct_cli.npz): unequal trajectories cut into tau-blocks, Ct of the unfitted
and the fitted vectors, rotation into the PAF frame before average vector / histogram / S2, zeta on S2."""
import contextlib
import io

import numpy as np

from oracle import ct_oracle


def _block_moments_stand_in(vecs3, frames_per_block):
    v = np.asarray(vecs3, dtype=np.float64)
    nFr, nR, _ = v.shape
    nB = -(-nFr // frames_per_block)
    out = np.zeros((nB, nR, 9))
    for b in range(nB):
        blk = v[b * frames_per_block: (b + 1) * frames_per_block]
        out[b, :, :3] = blk.sum(axis=0)
        k = 3
        for i in range(3):
            for j in range(i, 3):
                out[b, :, k] = (blk[..., i] * blk[..., j]).sum(axis=0)
                k += 1
    return out


def _numbers(text):
    return np.array([float(t) for t in text.replace("&", " ").split()])


def test_ct_cli_glue_reproduces_reference_files(golden, tmp_path, monkeypatch):
    from spinrelax_b200 import cli_ct, ct, hist
    monkeypatch.setattr(ct, "calculate_Ct_Palmer", lambda v, _verbose=True: ct_oracle.ct_palmer(np.asarray(v)))
    monkeypatch.setattr(ct, "_block_moments", _block_moments_stand_in)
    monkeypatch.setattr(hist, "sphere_histogram", lambda v, q=None, nb=72: ct_oracle.sphere_histogram(v, q, nb))
    g = golden("ct_cli.npz")
    files = []
    for tag in ("A", "B"):
        fn = tmp_path / ("traj%s.npz" % tag)
        np.savez(fn, vecs=g["fit" + tag], vecs_unfitted=g["ext" + tag], names=g["names"], dt=10.0)
        files.append(str(fn))
    pref = str(tmp_path / "o")
    qtxt = " ".join(repr(float(x)) for x in g["q"])
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        cli_ct.main(["-f"] + files + ["-o", pref, "--tau", "600", "--Ct", "--vecAvg", "--S2", "--vecHist", "--binary",
                                      "--vecRot", qtxt])
    # C(t): float32 values printed with str() -- identical text
    assert open(pref + "_Ctext.dat").read() == str(g["Ctext"])
    assert open(pref + "_Ctint.dat").read() == str(g["Ctint"])
    # S2 / average vector: the product sums in float64, the reference in float64 after the rotation too: %g text
    for suf, key in (("_S2.dat", "S2"), ("_avgvec.dat", "avgvec")):
        got, ref = open(pref + suf).read(), str(g[key])
        assert len(got.splitlines()) == len(ref.splitlines())
        assert np.allclose(_numbers(got), _numbers(ref), rtol=2e-5, atol=1e-9), suf
    z = np.load(pref + "_vecHistogram.npz", allow_pickle=True)
    assert str(z["dataType"]) == "LambertCylindrical" and bool(z["bHistogram"]) and list(z["axisLabels"]) == ["phi", "cos(theta)"]
    assert list(z["names"]) == list(g["names"])
    assert np.array_equal(z["data"].astype(np.int64), g["hist"])
    assert np.array_equal(z["edges"][0], g["edges_phi"]) and np.array_equal(z["edges"][1], g["edges_cos"])
