"""Host half of the relaxation stage without a GPU.  Every device evaluation of spinrelax_b200.specdens goes through
`_gpu_relax` (and `npufunc.Jomega`); here both are replaced by the oracle's NumPy restatement of the reference
arithmetic, and the CLI mirror -- prediction and the three optimisation modes -- must reproduce the files written
by the reference CLI (tests/golden/relax_cli.npz, relax_opt.npz)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import sd_oracle


class _JomegaStandIn:
    """npufunc.Jomega: callable with broadcasting and .outer (Jomega/Jomega.c)."""

    def __call__(self, x, y):
        return sd_oracle.jomega(np.asarray(x, dtype=float), np.asarray(y, dtype=float))

    def outer(self, a, b):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        return sd_oracle.jomega(a.reshape(a.shape + (1,) * b.ndim), b)


def _relax_stand_in(rotdif, models, omega, f_csa, f_dd, gammaA=-27.116e6, gammaB=267.513e6, time_fact=1e-12,
                    csa_per_residue=False, want="R", subset=None):
    """Same contract as specdens._gpu_relax: (nR, nField, nCSA, 6) = R1, R2, NOE and their sigmas over the bins."""
    from spinrelax_b200 import specdens as sd
    iso = isinstance(rotdif, sd.globalRotationalDiffusion_Isotropic)
    nR, nField = len(models), omega.shape[0]
    if want == "J":
        return np.array([[sd_oracle.j_isotropic(omega[f], rotdif.D, m.S2, m.C, m.tau, m.zeta) for f in range(nField)]
                         for m in models])
    nCSA = 1 if csa_per_residue else f_csa.shape[1]
    out = np.zeros((nR, nField, nCSA, 6))
    for i, m in enumerate(models):
        r = i if subset is None else subset[i]
        for f in range(nField):
            if iso:
                J, w = sd_oracle.j_isotropic(omega[f], rotdif.D, m.S2, m.C, m.tau, m.zeta), None
            else:
                A = rotdif.A_J[:, r, :] if rotdif.A_J.ndim == 3 else rotdif.A_J[r]
                w = rotdif.vecWeights[:, r] if rotdif.vecWeights is not None else None
                J = sd_oracle.j_axisymmetric(omega[f], A, rotdif.D_J, m.S2, m.C, m.tau, m.zeta)
            for c in range(nCSA):
                fc = f_csa[f, i] if csa_per_residue else f_csa[f, c]
                R1 = time_fact * (f_dd * (J[..., 2] + 3 * J[..., 1] + 6 * J[..., 4]) + fc * J[..., 1])
                R2 = time_fact * (0.5 * f_dd * (4 * J[..., 0] + J[..., 2] + 3 * J[..., 1] + 6 * J[..., 4] + 6 * J[..., 3])
                                  + 1.0 / 6.0 * fc * (4 * J[..., 0] + 3 * J[..., 1]))
                if w is None:
                    noe = 1.0 + time_fact * gammaB / (gammaA * R1) * f_dd * (6 * J[..., 4] - J[..., 2])
                    out[i, f, c, :3] = R1, R2, noe
                else:
                    r1, e1 = sd_oracle.wavg_std(R1, w)
                    r2, e2 = sd_oracle.wavg_std(R2, w)
                    noe = 1.0 + time_fact * gammaB / (gammaA * r1) * f_dd * (6 * J[..., 4] - J[..., 2])   # averaged R1 (G8)
                    n, en = sd_oracle.wavg_std(noe, w)
                    out[i, f, c] = r1, r2, n, e1, e2, en
    return out


@pytest.fixture
def cpu_relax(monkeypatch):
    from spinrelax_b200 import specdens as sd
    monkeypatch.setattr(sd, "_gpu_relax", _relax_stand_in)
    monkeypatch.setattr(sd.npufunc, "Jomega", _JomegaStandIn())
    return sd


def _write_inputs(golden, tmp_path, fields=(600,), from_opt=False):
    from spinrelax_b200 import hist
    gc, r = golden("relax_cli.npz"), golden("relax.npz")
    (tmp_path / "x_fittedCt.dat").write_text(str(gc["fitted"]))
    hist.save_vec_histogram(str(tmp_path / "h_vecHistogram.npz"), np.arange(6), r["hist"].astype(np.float64),
                            [r["edges_phi"], r["edges_cos"]])
    files = []
    go = golden("relax_opt.npz") if from_opt else None
    for f in fields:
        for t in ("R1", "R2", "NOE"):
            fn = tmp_path / ("e_%s_%d.dat" % (t, f))
            fn.write_text(str(go["expt_%s_%d" % (t, f)]) if from_opt else str(gc["expt_" + t]))
            files.append(str(fn))
    return files


def test_relax_cli_prediction_text_identical(golden, tmp_path, cpu_relax):
    from spinrelax_b200 import cli_relax
    files = _write_inputs(golden, tmp_path)
    g = golden("relax_cli.npz")
    with contextlib.redirect_stdout(io.StringIO()):
        cli_relax.main(["-f", str(tmp_path / "x_fittedCt.dat"), "--distfn", str(tmp_path / "h_vecHistogram.npz"), "-D", "2.1e-5",
                        "--aniso", "1.35", "-o", str(tmp_path / "ours")] + files)
    for t in ("R1", "R2", "NOE"):
        assert (tmp_path / ("ours_15N1H_600MHz_%s.xvg" % t)).read_text() == str(g["xvg_" + t]), t


@pytest.mark.parametrize("mode", ["Diso", "rsCSA", "mixed"])
def test_relax_cli_optimisation_text_identical(golden, tmp_path, cpu_relax, mode):
    """`--localopt powell` replays the reference's sequence of SciPy Powell searches; with the reference arithmetic
    underneath, the written files are the reference's character for character."""
    from spinrelax_b200 import cli_relax
    files = _write_inputs(golden, tmp_path, fields=(600, 800), from_opt=True)
    g = golden("relax_opt.npz")
    opt = {"Diso": "Diso", "rsCSA": "rsCSA", "mixed": "Diso,rsCSA"}[mode]
    extra = ["--cycles", "4"] if mode == "mixed" else []
    pref = str(tmp_path / "o")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cli_relax.main(["-f", str(tmp_path / "x_fittedCt.dat"), "--distfn", str(tmp_path / "h_vecHistogram.npz"), "-D", "2.1e-5",
                        "--aniso", "1.35", "-o", pref, "--opt", opt, "--localopt", "powell"] + extra + files)
    chi = float([l for l in buf.getvalue().splitlines() if "Final chi-value" in l][-1].split(":")[-1])
    assert abs(chi - float(g["chi_" + mode])) <= 1e-6 * float(g["chi_" + mode])
    for f in (600, 800):
        for t in ("R1", "R2", "NOE"):
            assert open("%s_15N1H_%dMHz_%s.xvg" % (pref, f, t)).read() == str(g["xvg_%s_%s_%d" % (mode, t, f)]), (t, f)
    if mode != "Diso":
        assert np.allclose(np.loadtxt(pref + "_CSA_opt.dat"), g["csa_" + mode], rtol=1e-9, atol=0)
