"""Host half of the dq stage without a GPU: the device reduction (`dq.dq_moment_sums`) is replaced by the oracle's
NumPy evaluation of the same raw sums, and the CLI mirror must then reproduce the reference's output files."""
import contextlib
import io

import numpy as np

from oracle import dq_oracle


def _moment_sums_stand_in(q, lags, nchunk=1):
    q = np.asarray(q, dtype=np.float32)
    if q.ndim == 2:
        q = q[None]
    lags = np.asarray(lags, dtype=np.int64)
    nCh = max(1, int(nchunk))
    M = np.zeros((lags.size, nCh, 6))
    n = np.zeros(lags.size, dtype=np.int64)
    counts = np.zeros((lags.size, nCh), dtype=np.int64)
    for k, d in enumerate(lags):
        v = dq_oracle.pooled_vectors(q, int(d))
        n[k] = len(v)
        nb = -(-len(v) // nCh)
        for c in range(nCh):
            blk = v[nb * c: min(len(v), nb * (c + 1))]
            counts[k, c] = len(blk)
            S = np.einsum("ti,tj->ij", blk, blk)
            M[k, c] = [S[0, 0], S[0, 1], S[0, 2], S[1, 1], S[1, 2], S[2, 2]]
    return M, n, counts


def _isnum(t):
    try:
        float(t)
        return True
    except ValueError:
        return False


def test_dq_cli_host_logic_reproduces_reference_files(golden, tmp_path, monkeypatch):
    from spinrelax_b200 import dq
    monkeypatch.setattr(dq, "dq_moment_sums", _moment_sums_stand_in)
    g = golden("dq_cli.npz")
    fn = tmp_path / "colvar-q"
    fn.write_text(str(g["plumed"]))
    pref = str(tmp_path / "rotdif")
    with contextlib.redirect_stdout(io.StringIO()):
        dq.main(["--iso", "--aniso", "-f", str(fn), "-o", pref, "--mindt", "500", "--skip", "500", "--maxdt", "50000",
                 "--num_chunk", "4"])
    for suf, key in (("-aniso2.dat", "aniso2"), ("-aniso_q.dat", "aniso_q"), ("-iso.dat", "iso"), ("-moi.xyz", "moi_xyz")):
        got, ref = open(pref + suf).read().splitlines(), str(g[key]).splitlines()
        assert len(got) == len(ref), suf
        for a, b in zip(got, ref):       # same structure line by line; numbers compared below
            assert [t for t in a.split() if not _isnum(t)] == [t for t in b.split() if not _isnum(t)], (suf, a, b)
        if suf in ("-moi.xyz", "-iso.dat"):
            continue                      # eigenvector signs / the unphysical shipped iso curve (G1): structure only
        na = np.array([float(t) for l in got for t in l.replace("=", " ").split() if _isnum(t)])
        nb = np.array([float(t) for l in ref for t in l.replace("=", " ").split() if _isnum(t)])
        assert na.shape == nb.shape and np.allclose(na, nb, rtol=2e-6, atol=1e-12), suf
    assert open(pref + "-aniso_q.dat").readline() == str(g["aniso_q"]).splitlines()[0] + "\n"


def test_dq_moment_stand_in_is_the_pooled_reference(golden):
    """The stand-in itself against the reference-generated pooled moments (so the test above rests on pinned sums)."""
    from spinrelax_b200 import dq
    g = golden("dq_multi.npz")
    nch = int(g["nchunk"])
    M, n, counts = _moment_sums_stand_in(g["q"], g["lags"], nch)
    for k in range(len(g["lags"])):
        full = M[k].sum(axis=0)
        assert np.isclose(dq._iso_shipped(full), g["iso"][k], rtol=1e-12)
        assert np.allclose(dq._sym3(full) / n[k], g["moi"][k], rtol=1e-12, atol=1e-20)
        for c in range(nch):
            assert np.allclose(dq._sym3(M[k, c]) / counts[k, c], g["chunk_moi"][k][c], rtol=1e-12, atol=1e-20)
