"""GPU parity of K5 (batched C(t) fits), K6 (J(omega) -> R1/R2/NOE) and K7 (Jomega ufunc)."""
import io
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import fit_oracle, sd_oracle

pytestmark = pytest.mark.gpu

RTOL_RATE = 1e-4      # north star: fitted rates within 1e-4 relative


def _models_from_golden(g, fitct, zeta):
    ac = fitct.autoCorrelations()
    for i, row in enumerate(g["params"]):
        nc = int(row[0])
        ac.add_model(str(i), name=i, listC=list(row[2:2 + nc]), listTau=list(row[5:5 + nc]), S2=row[1])
    ac.set_zeta(zeta)
    return ac


def test_jomega_ufunc(golden):
    from spinrelax_b200 import npufunc
    g = golden("relax.npz")
    out = npufunc.Jomega(g["jomega_x"], g["jomega_y"])
    assert out.dtype == np.float64 and np.array_equal(out, g["jomega_out"])
    assert np.array_equal(npufunc.Jomega.outer(g["jomega_x"][:3], g["jomega_y"][:5]), g["jomega_outer"])
    f32 = npufunc.Jomega(g["jomega_x"].astype(np.float32), g["jomega_y"].astype(np.float32))
    assert f32.dtype == np.float32 and np.array_equal(f32, g["jomega_f32"])
    assert npufunc.Jomega(4.0, 0.0) == 0.25 and npufunc.Jomega(3.0, 3.0) == 3.0 / 18.0
    assert np.isnan(npufunc.Jomega(0.0, 0.0))
    assert npufunc.Jomega.nin == 2 and npufunc.Jomega.nout == 1 and 'dd->d' in npufunc.Jomega.types
    big = np.random.default_rng(1).uniform(1e-6, 1, (257, 33))
    assert np.array_equal(npufunc.Jomega(big, big[:1]), sd_oracle.jomega(big, big[:1]))   # broadcasting
    # a real numpy.ufunc object like the reference's (Jomega.c:135-156): out=, where=, integer inputs promoted to 'dd->d'
    assert isinstance(npufunc.Jomega, np.ufunc)
    out = np.full(5, -1.0)
    r = npufunc.Jomega(np.arange(1.0, 6.0), 2.0, out=out, where=np.array([True, False, True, False, True]))
    assert r is out and np.array_equal(out[[1, 3]], [-1.0, -1.0]) and out[0] == 1.0 / 5.0 and out[4] == 5.0 / 29.0
    assert npufunc.Jomega(np.arange(1, 4), 1).dtype == np.float64
    with pytest.raises(TypeError):
        npufunc.Jomega(np.ones(2, dtype=np.complex128), 1.0)


def test_relaxation_classes_vs_golden(golden, tmp_path):
    from spinrelax_b200 import fitct, hist, specdens as sd
    g = golden("relax.npz")
    zeta = float(g["zeta"])
    ac = _models_from_golden(g, fitct, zeta)
    fn = str(tmp_path / "h_vecHistogram.npz")
    hist.save_vec_histogram(fn, np.arange(6), g["hist"].astype(np.float64), [g["edges_phi"], g["edges_cos"]])
    for tag, Dani in (("prolate", 1.35), ("oblate", 0.8)):
        rot = sd.globalRotationalDiffusion_Axisymmetric(D=[float(g["Diso"]), Dani], bConvert=False)
        rot.import_frame_vectors(fn)
        for field in (600.133, 800.0):
            w = sd.angularFrequencies("15N", "1H", field, "MHz", "ps")
            for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
                ex = cls(name, "ps", w, rot, ac)
                ex.eval()
                ref = g["%s_%s_%d" % (tag, name, round(field))]
                assert rel_err(ex.values, ref[0]) < 1e-10, (tag, name, field)
                assert rel_err(ex.errors, ref[1]) < 1e-8, (tag, name, field)
        # J(omega) surface through the GPU ufunc
        w = sd.angularFrequencies("15N", "1H", 600.133, "MHz", "ps")
        J = rot.calc_Jomega(w.omega, ac)
        vec, _ = sd_oracle.hist_to_vectors(g["hist"].astype(np.float64), (g["edges_phi"], g["edges_cos"]))
        A = sd_oracle.a_coefficients(vec, Dani > 1)
        row = g["params"][2]; nc = int(row[0])
        Jo = sd_oracle.j_axisymmetric(w.omega, A, sd_oracle.d_coefficients(float(g["Diso"]), Dani), row[1],
                                      row[2:2 + nc], row[5:5 + nc], zeta)
        assert rel_err(J[:, 2, :], Jo) < 1e-12
    iso = sd.globalRotationalDiffusion_Isotropic(D=float(g["Diso"]))
    w = sd.angularFrequencies("15N", "1H", 600.133, "MHz", "ps")
    assert np.array_equal(w.omega, g["omega_600"]) and w.get_factor_DD() == float(g["f_dd"])
    for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
        ex = cls(name, "ps", w, iso, ac)
        ex.eval()
        assert ex.errors is None and rel_err(ex.values, g["iso_%s_600" % name]) < 1e-12
    # per-residue CSA array, eval(ind=i): the rsCSA inner loop
    rot = sd.globalRotationalDiffusion_Axisymmetric(D=[float(g["Diso"]), 1.35], bConvert=False)
    rot.import_frame_vectors(fn)
    w2 = sd.angularFrequencies("15N", "1H", 600.133, "MHz", "ps")
    w2.initialise_CSA_array(6, g["csa_array"])
    for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
        ex = cls(name, "ps", w2, rot, ac)
        for i in range(6):
            ex.eval(ind=i)
        assert rel_err(ex.values, g["csa_%s_600" % name][0]) < 1e-10
        assert rel_err(ex.errors, g["csa_%s_600" % name][1]) < 1e-8
        ex2 = cls(name, "ps", w2, rot, ac)
        ex2.eval()
        assert rel_err(ex2.values, g["csa_%s_600" % name][0]) < 1e-10


def test_relax_grid_vs_oracle(golden, tmp_path):
    """residue x field x CSA grid in one launch against the bin-by-bin oracle."""
    from spinrelax_b200 import fitct, hist, specdens as sd
    g = golden("relax.npz")
    zeta = float(g["zeta"])
    ac = _models_from_golden(g, fitct, zeta)
    fn = str(tmp_path / "h.npz")
    hist.save_vec_histogram(fn, np.arange(6), g["hist"].astype(np.float64), [g["edges_phi"], g["edges_cos"]])
    rot = sd.globalRotationalDiffusion_Axisymmetric(D=[float(g["Diso"]), 1.35], bConvert=False)
    rot.import_frame_vectors(fn)
    fields = [500.0, 600.133, 700.0, 800.0, 950.0]
    csa = np.linspace(-220e-6, -120e-6, 64)
    res = sd.relax_grid(rot, ac, fields, csa)
    assert res["R1"][0].shape == (6, 5, 64)
    vec, wts = sd_oracle.hist_to_vectors(g["hist"].astype(np.float64), (g["edges_phi"], g["edges_cos"]))
    models = [(r[1], r[2:2 + int(r[0])], r[5:5 + int(r[0])]) for r in g["params"]]
    for fi in (0, 3):
        for ci in (0, 17, 63):
            o = sd_oracle.relax_axisymmetric(fields[fi], float(g["Diso"]), 1.35, vec, wts, models, csa=csa[ci], zeta=zeta)
            for name in ("R1", "R2", "NOE"):
                assert rel_err(res[name][0][:, fi, ci], o[name][0]) < 1e-10
                assert rel_err(res[name][1][:, fi, ci], o[name][1]) < 1e-8


def _rates(fitct, sd, models_ac):
    iso = sd.globalRotationalDiffusion_Isotropic(D=2.1e-5)
    out = []
    for f in (600.133, 800.0):
        r = sd.relax_grid(iso, models_ac, [f])
        out.append(np.stack([r[k][0][:, 0, 0] for k in ("R1", "R2", "NOE")]))
    return np.array(out)


def test_fit_single_rungs_vs_golden(golden):
    """Same p0/bounds as the reference (real fitting_Ct_functions.py on SciPy, tests/golden/fit.npz): chi^2, the
    quality flags and the parameters of every well-posed rung."""
    from spinrelax_b200 import fitct
    g = golden("fit.npz")
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    k = 0
    for i in range(len(Ct)):
        for npar in (2, 3, 5):
            r = g["single"][k]; k += 1
            m = fitct.autoCorrelationModel(name=i)
            m.set_nParams(npar)
            chi, qual = m.conduct_curve_fitting(t, Ct[i], dCt[i], bReInitialise=True, fp=io.StringIO())
            assert np.isfinite(chi) == np.isfinite(r[1]), (i, npar)        # fails exactly where curve_fit raised upstream
            assert [float(b) for b in qual] == list(r[2:5]), (i, npar)
            if not np.isfinite(r[1]):
                continue
            assert rel_err(chi, r[1]) < 1e-6, (i, npar, chi, r[1])
            if npar <= 3:
                nc = npar // 2
                assert rel_err(m.C, r[6:6 + nc]) < 1e-6 and rel_err(m.tau, r[9:9 + nc]) < 1e-6


def test_fit_ladder_rates_vs_reference(golden):
    """Full ladder on the GPU, then isotropic J -> R1/R2/NOE: the rung the reference chose for every residue, and rates
    within 1e-4 of the reference chain."""
    from spinrelax_b200 import fitct, specdens as sd
    g = golden("fit.npz")
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    nres = len(Ct)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(nres)], [t] * nres, Ct, dCt)
    chis = ac.fit_all_residues(fp=io.StringIO())
    ref = fitct.autoCorrelations()
    for i, row in enumerate(g["ladder"]):
        nc = int(row[0]) // 2
        ref.add_model(str(i), name=i, listC=list(row[3:3 + nc]), listTau=list(row[7:7 + nc]), S2=row[2],
                      bS2Fast=(int(row[0]) % 2 == 1))
        assert ac.model[str(i)].nParams == int(row[0]), i
    a, b = _rates(fitct, sd, ac), _rates(fitct, sd, ref)
    for i in range(nres):
        assert rel_err(a[:, :, i], b[:, :, i]) < RTOL_RATE, i
        assert rel_err(chis[i], g["ladder"][i][1]) < 1e-6
    # one model at a time through the reference-shaped API gives the same answer as the batched ladder
    m = fitct.autoCorrelationModel(name=0)
    chi = m.optimised_curve_fitting(t, Ct[0], dCt[0], fp=io.StringIO())
    assert m.nParams == ac.model["0"].nParams and rel_err(chi, chis[0]) < 1e-12


def test_fit_against_scipy_oracle_random():
    """Fresh curves, oracle = SciPy TRF (the reference's solver) run here: same rung, fitted curve and rates agree."""
    from spinrelax_b200 import fitct, specdens as sd
    rng = np.random.default_rng(42)
    t = (np.arange(400) + 1.0) * 5.0
    n = 12
    Y, SG = [], []
    for i in range(n):
        S2 = rng.uniform(0.5, 0.9); C1 = (1 - S2) * rng.uniform(0.3, 0.7); C2 = (1 - S2) - C1
        y = S2 + C1 * np.exp(-t / rng.uniform(20, 80)) + C2 * np.exp(-t / rng.uniform(300, 900))
        sg = np.full_like(t, 0.003)
        Y.append(y + rng.standard_normal(len(t)) * 0.0015); SG.append(sg)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(n)], [t] * n, np.array(Y), np.array(SG))
    ac.fit_all_residues(fp=io.StringIO())
    ref = fitct.autoCorrelations()
    for i in range(n):
        best = fit_oracle.fit_ladder(t, Y[i], SG[i])
        ref.add_model(str(i), name=i, listC=list(best["C"]), listTau=list(best["tau"]), S2=best["S2"],
                      bS2Fast=(best["n_params"] % 2 == 1))
        assert best["n_params"] == ac.model[str(i)].nParams, i
    a, b = _rates(fitct, sd, ac), _rates(fitct, sd, ref)
    assert rel_err(a[:, :2], b[:, :2]) < RTOL_RATE                               # R1, R2
    # NOE crosses zero at 800 MHz for the poorly fitting 2-parameter models: 1e-4 of its O(1) scale
    assert np.allclose(a[:, 2], b[:, 2], rtol=RTOL_RATE, atol=2e-5)


def test_config5_all_1000_residues_vs_scipy_oracle():
    """BASELINE config 5: the 1000 synthetic residues of the fit benchmark, whole 2-3-5-7-9 ladder on the GPU against
    the oracle's ladder on SciPy (about 35 s of CPU).  Gates: the same rung on >= 99 % of the residues (every
    disagreement is printed with both chi^2 so the margin to the 0.5 threshold can be read), and on every agreeing
    residue chi^2 <= 1e-6, parameters <= 1e-4 and R1/R2/NOE at two fields <= 1e-4 of the reference chain.  Also:
    every solve that hit SciPy's evaluation cap (curve_fit raises -> the reference's `failed!` path) is reported as
    failed, and fewer than 0.5 % of the solves the *selected* ladders depend on end that way."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from bench_secondary import synth_curves
    from spinrelax_b200 import fitct, specdens as sd
    n = 1000
    t, Y, SG = synth_curves(n, 500, 77)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(n)], [t] * n, Y, SG)
    log = io.StringIO()
    chis = ac.fit_all_residues(fp=log)
    ref = fitct.autoCorrelations()
    refs, bad = [], []
    for i in range(n):
        best = fit_oracle.fit_ladder(t, Y[i], SG[i])
        refs.append(best)
        ref.add_model(str(i), name=i, listC=list(best["C"]), listTau=list(best["tau"]), S2=best["S2"],
                      bS2Fast=(best["n_params"] % 2 == 1))
        if best["n_params"] != ac.model[str(i)].nParams:
            bad.append((i, ac.model[str(i)].nParams, best["n_params"], float(chis[i]), float(best["chi"])))
    for row in bad:
        print("selection differs: residue %d ours nParams=%d chi=%g, oracle nParams=%d chi=%g" % (row[0], row[1], row[3], row[2], row[4]))
    assert len(bad) <= n // 100, bad
    same = np.array([ac.model[str(i)].nParams == refs[i]["n_params"] for i in range(n)])
    a, b = _rates(fitct, sd, ac), _rates(fitct, sd, ref)
    worst = 0.0
    for i in np.nonzero(same)[0]:
        m, r = ac.model[str(i)], refs[i]
        assert rel_err(chis[i], r["chi"]) < 1e-6, i
        assert rel_err(m.tau, r["tau"]) < 1e-4 and np.max(np.abs(m.C - r["C"])) < 1e-4 and abs(m.S2 - r["S2"]) < 1e-4, i
        assert rel_err(a[:, :2, i], b[:, :2, i]) < RTOL_RATE, i
        assert np.allclose(a[:, 2, i], b[:, 2, i], rtol=RTOL_RATE, atol=2e-5), i
        worst = max(worst, rel_err(a[:, :2, i], b[:, :2, i]))
    print("config 5: %d/%d rungs agree, worst R1/R2 deviation %.3g" % (int(same.sum()), n, worst))
    # the cap: count the reference's own `failed!` lines among all solves of the ladders (about 3000 here)
    n_solves = log.getvalue().count("yield chiSq")
    n_failed = log.getvalue().count("failed!")
    assert n_failed <= 0.02 * n_solves, (n_failed, n_solves)


def test_fit_long_curves_use_the_global_workspace():
    """Curves too long for shared memory (L = 6000 > ~1900 points) take the workspace path of the kernel: same result
    as the oracle on a subsampled-equivalent problem solved by SciPy directly."""
    from spinrelax_b200 import fitct
    rng = np.random.default_rng(5)
    L, n = 6000, 3
    t = (np.arange(L) + 1.0) * 2.0
    Y = np.array([0.8 + 0.15 * np.exp(-t / tau) + rng.standard_normal(L) * 0.002 for tau in (150.0, 700.0, 2500.0)])
    SG = np.full((n, L), 0.004)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(n)], [t] * n, Y, SG)
    chis = ac.fit_all_residues(fp=io.StringIO())
    for i in range(n):
        best = fit_oracle.fit_ladder(t, Y[i], SG[i])
        m = ac.model[str(i)]
        assert m.nParams == best["n_params"], i
        assert rel_err(chis[i], best["chi"]) < 1e-6 and rel_err(m.tau, best["tau"]) < 1e-5


def test_cli_relax_and_fit_files(golden, tmp_path):
    """CLI mirrors: calculate-relaxations-multi-field.py output text identical to the reference's; fittedCt
    writer/reader round trip through calculate-fitted-Ct's mirror."""
    import contextlib
    from spinrelax_b200 import cli_fit, cli_relax, fitct, hist, io_formats
    g, r = golden("relax_cli.npz"), golden("relax.npz")
    (tmp_path / "x_fittedCt.dat").write_text(str(g["fitted"]))
    hist.save_vec_histogram(str(tmp_path / "h_vecHistogram.npz"), np.arange(6), r["hist"].astype(np.float64),
                            [r["edges_phi"], r["edges_cos"]])
    files = []
    for t in ("R1", "R2", "NOE"):
        (tmp_path / ("e_%s.dat" % t)).write_text(str(g["expt_" + t]))
        files.append(str(tmp_path / ("e_%s.dat" % t)))
    with contextlib.redirect_stdout(io.StringIO()):
        cli_relax.main(["-f", str(tmp_path / "x_fittedCt.dat"), "--distfn", str(tmp_path / "h_vecHistogram.npz"), "-D", "2.1e-5",
                        "--aniso", "1.35", "-o", str(tmp_path / "ours")] + files)
    for t in ("R1", "R2", "NOE"):
        assert (tmp_path / ("ours_15N1H_600MHz_%s.xvg" % t)).read_text() == str(g["xvg_" + t]), t
    # fit CLI: Ctint-format input -> fittedCt file that the reference-format reader parses back
    f = golden("fit.npz")
    t, Ct, dCt = f["t"], f["Ct"], f["dCt"]
    io_formats.print_sxylist(str(tmp_path / "c_Ctint.dat"), list(range(1, 10)), t,
                             np.stack((Ct.astype(np.float32), dCt.astype(np.float32)), axis=-1))
    with contextlib.redirect_stdout(io.StringIO()):
        cli_fit.main(["-f", str(tmp_path / "c_Ctint.dat"), "-o", str(tmp_path / "c")])
    back = fitct.read_fittedCt_parameters(str(tmp_path / "c_fittedCt.dat"))
    assert back.nModels == 9
    for i, row in enumerate(f["ladder"]):
        m = back.model[str(i + 1)]
        if m.nParams == int(row[0]):
            assert abs(m.S2 - row[2]) < 2e-4 * max(1.0, abs(row[2]))


def _xvg_rows(text):
    rows = [l.split() for l in text.splitlines() if l and l[0] not in "#@&"]
    return np.array([[float(x) for x in r[1:]] for r in rows])


def _xvg_header(text, key):
    for l in text.splitlines():
        if l.startswith("# ") and (" %s:" % key) in l:
            return l
    return None


@pytest.mark.parametrize("mode", ["Diso", "rsCSA", "mixed"])
def test_cli_relax_optimisation_matches_reference(golden, tmp_path, mode):
    """--opt against the reference CLI run on the same files (tests/golden/relax_opt.npz).

    `--localopt powell` replays the reference's own sequence of SciPy Powell searches on top of the GPU evaluation,
    so it must land on the reference's numbers (predictions agree to 1e-10, the searches to their own tolerance);
    the batched residue-specific solver must reach a chi that is at least as good and the same CSA values within
    Powell's line-search tolerance."""
    import contextlib
    from spinrelax_b200 import cli_relax, hist
    g, gc, r = golden("relax_opt.npz"), golden("relax_cli.npz"), golden("relax.npz")
    (tmp_path / "x_fittedCt.dat").write_text(str(gc["fitted"]))
    hist.save_vec_histogram(str(tmp_path / "h_vecHistogram.npz"), np.arange(6), r["hist"].astype(np.float64),
                            [r["edges_phi"], r["edges_cos"]])
    files = []
    for f in (600, 800):
        for t in ("R1", "R2", "NOE"):
            fn = tmp_path / ("e_%s_%d.dat" % (t, f))
            fn.write_text(str(g["expt_%s_%d" % (t, f)]))
            files.append(str(fn))
    opt = {"Diso": "Diso", "rsCSA": "rsCSA", "mixed": "Diso,rsCSA"}[mode]
    extra = ["--cycles", "4"] if mode == "mixed" else []
    for local in (["powell", "batched"] if mode != "Diso" else ["batched"]):
        pref = str(tmp_path / ("o_" + local))
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            cli_relax.main(["-f", str(tmp_path / "x_fittedCt.dat"), "--distfn", str(tmp_path / "h_vecHistogram.npz"), "-D",
                            "2.1e-5", "--aniso", "1.35", "-o", pref, "--opt", opt, "--localopt", local] + extra + files)
        chi = float([l for l in buf.getvalue().splitlines() if "Final chi-value" in l][-1].split(":")[-1])
        chi_ref = float(g["chi_" + mode])
        if local == "powell" or mode == "Diso":
            assert abs(chi - chi_ref) < 2e-4 * chi_ref
        else:
            assert chi < chi_ref * (1 + 2e-3)
        for f in (600, 800):
            for t in ("R1", "R2", "NOE"):
                ours = open("%s_15N1H_%dMHz_%s.xvg" % (pref, f, t)).read()
                ref = str(g["xvg_%s_%s_%d" % (mode, t, f)])
                a, b = _xvg_rows(ours), _xvg_rows(ref)
                assert a.shape == b.shape
                tol = 1e-4 if (local == "powell" or mode == "Diso") else 2e-3
                assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)) < tol, (t, f, local)
                assert _xvg_header(ours, "Diso").split()[1] == _xvg_header(ref, "Diso").split()[1]   # Optimised / Fixed
        if mode != "Diso":
            csa, csa_ref = np.loadtxt(pref + "_CSA_opt.dat"), g["csa_" + mode]
            assert np.array_equal(csa[:, 0], csa_ref[:, 0])
            assert np.max(np.abs(csa[:, 1] / csa_ref[:, 1] - 1)) < (2e-4 if local == "powell" else 3e-3)


def test_vectorised_ladder_equals_residue_loop(golden):
    """fit_all_residues (array form of the selection ladder) against fit_all_residues_loop (the reference's
    per-residue flow through the model objects): same rung chosen, same parameters, same printed log lines."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from bench_secondary import synth_curves
    from spinrelax_b200 import fitct
    f = golden("fit.npz")
    sets = [synth_curves(60, 500, 77), (f["t"], f["Ct"], f["dCt"])]
    for t, Y, SG in sets:
        nR = len(Y)
        for sig in (SG, None):
            res = []
            for method in ("fit_all_residues", "fit_all_residues_loop"):
                ac = fitct.autoCorrelations()
                ac.import_target_array([str(i) for i in range(nR)], [t] * nR, Y, sig)
                log = io.StringIO()
                chis = getattr(ac, method)(fp=log)
                res.append((ac, chis, sorted(log.getvalue().splitlines())))
            (a, ca, la), (b, cb, lb) = res
            assert np.array_equal(ca, cb)
            assert la == lb
            for k in a.model:
                ma, mb = a.model[k], b.model[k]
                assert (ma.nParams, ma.bS2Fast, ma.bHasFit) == (mb.nParams, mb.bS2Fast, mb.bHasFit), k
                assert np.array_equal(ma.C, mb.C) and np.array_equal(ma.tau, mb.tau) and ma.S2 == mb.S2, k
                if ma.bHasFit:
                    assert np.allclose(ma.dC, mb.dC, rtol=1e-9, atol=0) and np.allclose(ma.dtau, mb.dtau, rtol=1e-9, atol=0)
                    assert ma.chiSq == mb.chiSq
