"""dropin/ scripts run as real subprocesses by their reference names on the GPU, with run-all.bash's argument strings
(run-all.bash:379-387, 475-481, 488-491): outputs against the goldens the unmodified reference scripts produced."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from test_dq_host_cpu import _isnum

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "dropin")


def _run(script, argv, cwd):
    res = subprocess.run([sys.executable, os.path.join(DROPIN, script)] + argv, cwd=str(cwd), capture_output=True, text=True)
    assert res.returncode == 0, (script, res.stdout[-2000:], res.stderr[-2000:])
    return res.stdout


def _numbers(text):
    return np.array([float(t) for l in text.splitlines() for t in l.replace("=", " ").replace("&", " ").split() if _isnum(t)])


@pytest.mark.parametrize("kind", ["plumed", "xvg"])
def test_dq_script(golden, tmp_path, kind):
    if kind == "plumed":
        g, fn, t100, tau = golden("dq_cli.npz"), tmp_path / "colvar-q", "500", "50000"
        fn.write_text(str(g["plumed"]))
    else:
        g, fn, t100, tau = golden("dq_xvg.npz"), tmp_path / "rotmat.xvg", "200", "10000"
        fn.write_text(str(g["xvg"]))
    _run("calculate-dq-distribution.py", ["--iso", "--aniso", "-f", str(fn), "-o", "rotdif", "--mindt", t100, "--skip", t100,
                                          "--maxdt", tau, "--num_chunk", "4"], tmp_path)
    for suf, key in (("-aniso2.dat", "aniso2"), ("-aniso_q.dat", "aniso_q")):
        a, b = _numbers((tmp_path / ("rotdif" + suf)).read_text()), _numbers(str(g[key]))
        assert a.shape == b.shape and np.allclose(a, b, rtol=2e-6, atol=1e-12), suf


def test_ct_fit_relax_chain(golden, tmp_path):
    g = golden("ct_cli.npz")
    files = []
    for tag in ("A", "B"):
        fn = tmp_path / ("traj%s.npz" % tag)
        np.savez(fn, vecs=g["fit" + tag], vecs_unfitted=g["ext" + tag], names=g["names"], dt=10.0)
        files.append(str(fn))
    quat = " ".join(repr(float(x)) for x in g["q"])
    _run("calculate-Ct-from-traj.py", ["-s", "reference.pdb", "-f"] + files + ["--tau", "600", "-o", "rotdif", "--vecRot", quat,
                                                                              "--vecHist", "--binary", "--vecAvg", "--S2", "--Ct"],
         tmp_path)
    a, b = _numbers((tmp_path / "rotdif_Ctint.dat").read_text()), _numbers(str(g["Ctint"]))
    assert a.shape == b.shape and np.allclose(a, b, rtol=2e-5, atol=2e-6)       # float32 reference text, see test_ct_gpu
    z = np.load(tmp_path / "rotdif_vecHistogram.npz", allow_pickle=True)
    assert np.array_equal(z["data"].astype(np.int64), g["hist"])
    _run("calculate-fitted-Ct.py", ["-f", "rotdif_Ctint.dat", "-o", "rotdif"], tmp_path)
    assert (tmp_path / "rotdif_fittedCt.dat").exists()
    # relaxation script on the reference's own fitted-Ct text and experiment files (tests/golden/relax_cli.npz)
    gc, r = golden("relax_cli.npz"), golden("relax.npz")
    from spinrelax_b200 import hist
    (tmp_path / "x_fittedCt.dat").write_text(str(gc["fitted"]))
    hist.save_vec_histogram(str(tmp_path / "h_vecHistogram.npz"), np.arange(6), r["hist"].astype(np.float64),
                            [r["edges_phi"], r["edges_cos"]])
    expt = []
    for t in ("R1", "R2", "NOE"):
        (tmp_path / ("e_%s.dat" % t)).write_text(str(gc["expt_" + t]))
        expt.append("e_%s.dat" % t)
    _run("calculate-relaxations-multi-field.py", ["-f", "x_fittedCt.dat", "--distfn", "h_vecHistogram.npz", "-D", "2.1e-5",
                                                  "--aniso", "1.35", "-o", "ours"] + expt, tmp_path)
    for t in ("R1", "R2", "NOE"):
        assert (tmp_path / ("ours_15N1H_600MHz_%s.xvg" % t)).read_text() == str(gc["xvg_" + t]), t
