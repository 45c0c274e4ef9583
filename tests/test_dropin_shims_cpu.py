"""The reference-named entry points under dropin/ (what run-all.bash's $script_loc would point at), executed by their
reference names with run-all.bash's own argument strings (run-all.bash:379-387, 475-481, 488-491).  No GPU here: the
device stage of each script is replaced *in the test* by the oracle's restatement of that stage, exactly as the
tests/test_*_host_cpu.py files do, so what is pinned is the surface (script names, flags, output files) and the glue.
`tests/test_dropin_gpu.py` runs the same scripts as real subprocesses on the device."""
import contextlib
import io
import os
import runpy
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import ct_oracle
from test_abi_and_host import _scipy_stand_in
from test_ct_cli_host_cpu import _block_moments_stand_in
from test_dq_host_cpu import _isnum, _moment_sums_stand_in

DROPIN = os.path.join(ROOT, "dropin")


def _run(script, argv, monkeypatch):
    monkeypatch.setattr(sys, "argv", [os.path.join(DROPIN, script)] + argv)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        try:
            runpy.run_path(os.path.join(DROPIN, script), run_name="__main__")
        except SystemExit as e:
            assert not e.code, (script, e.code)


def _same_numbers(got, ref, rtol):
    got, ref = got.splitlines(), ref.splitlines()
    assert len(got) == len(ref)
    na = np.array([float(t) for l in got for t in l.replace("=", " ").split() if _isnum(t)])
    nb = np.array([float(t) for l in ref for t in l.replace("=", " ").split() if _isnum(t)])
    assert na.shape == nb.shape and np.allclose(na, nb, rtol=rtol, atol=1e-12)


@pytest.mark.parametrize("kind", ["plumed", "xvg"])
def test_step2_dq_distribution_by_its_reference_name(golden, tmp_path, monkeypatch, kind):
    from spinrelax_b200 import dq
    monkeypatch.setattr(dq, "dq_moment_sums", _moment_sums_stand_in)
    if kind == "plumed":
        g, fn, t100, tau = golden("dq_cli.npz"), tmp_path / "colvar-q", "500", "50000"
        fn.write_text(str(g["plumed"]))
    else:
        g, fn, t100, tau = golden("dq_xvg.npz"), tmp_path / "rotmat.xvg", "200", "10000"
        fn.write_text(str(g["xvg"]))
    pref = str(tmp_path / "rotdif")
    _run("calculate-dq-distribution.py", ["--iso", "--aniso", "-f", str(fn), "-o", pref, "--mindt", t100, "--skip", t100,
                                          "--maxdt", tau, "--num_chunk", "4"], monkeypatch)
    for suf, key in (("-aniso2.dat", "aniso2"), ("-aniso_q.dat", "aniso_q")):
        _same_numbers(open(pref + suf).read(), str(g[key]), 2e-6)
    assert os.path.exists(pref + "-iso.dat")


def test_xvg_rotation_matrices_convert_like_the_reference(golden, tmp_path):
    """`.xvg` input (calculate-dq-distribution.py:389-408, 487-492): bit-identical to the real load_xys +
    rotmatrix_to_quaternion(bInvert=True)."""
    from spinrelax_b200 import dq
    g = golden("dq_xvg.npz")
    fn = tmp_path / "r.xvg"
    fn.write_text(str(g["xvg"]))
    fields, data = dq.read_quaternion_input(str(fn))
    assert data.shape == g["data"].shape and np.array_equal(data, g["data"])


def test_step3_ct_then_fit_by_their_reference_names(golden, tmp_path, monkeypatch):
    from spinrelax_b200 import ct, fitct, hist
    monkeypatch.setattr(ct, "calculate_Ct_Palmer", lambda v, _verbose=True: ct_oracle.ct_palmer(np.asarray(v)))
    monkeypatch.setattr(ct, "_block_moments", _block_moments_stand_in)
    monkeypatch.setattr(hist, "sphere_histogram", lambda v, q=None, nb=72: ct_oracle.sphere_histogram(v, q, nb))
    monkeypatch.setattr(fitct, "gpu_curve_fit", _scipy_stand_in)
    g = golden("ct_cli.npz")
    files = []
    for tag in ("A", "B"):
        fn = tmp_path / ("traj%s.npz" % tag)
        np.savez(fn, vecs=g["fit" + tag], vecs_unfitted=g["ext" + tag], names=g["names"], dt=10.0)
        files.append(str(fn))
    pref = str(tmp_path / "rotdif")
    quat = " ".join(repr(float(x)) for x in g["q"])
    # run-all.bash:475-481: -s $refpdb_loc -f $sxtc_list --tau $tau_ps -o ${outpref} --vecRot "$quat" $fittxtstr
    #                       $vecDistArgs --vecAvg --S2 --Ct        (vecDistArgs = --vecHist --binary)
    _run("calculate-Ct-from-traj.py", ["-s", "reference.pdb", "-f"] + files + ["--tau", "600", "-o", pref, "--vecRot", quat,
                                                                              "--fitsel", "name CA", "--vecHist", "--binary",
                                                                              "--vecAvg", "--S2", "--Ct"], monkeypatch)
    assert open(pref + "_Ctint.dat").read() == str(g["Ctint"])
    assert os.path.exists(pref + "_vecHistogram.npz") and os.path.exists(pref + "_S2.dat")
    # run-all.bash:488-491
    _run("calculate-fitted-Ct.py", ["-f", pref + "_Ctint.dat", "-o", pref], monkeypatch)
    back = fitct.read_fittedCt_parameters(pref + "_fittedCt.dat")
    assert back.nModels == len(g["names"])


def test_fit_script_writes_the_reference_file(golden, tmp_path, monkeypatch):
    from spinrelax_b200 import fitct
    monkeypatch.setattr(fitct, "gpu_curve_fit", _scipy_stand_in)
    g = golden("fit_cli.npz")
    (tmp_path / "c_Ctint.dat").write_text(str(g["ctint"]))
    _run("calculate-fitted-Ct.py", ["-f", str(tmp_path / "c_Ctint.dat"), "-o", str(tmp_path / "c")], monkeypatch)
    assert (tmp_path / "c_fittedCt.dat").read_text() == str(g["ladder"])


def test_help_sel_exits_cleanly(monkeypatch, capsys):
    monkeypatch.setattr(sys, "argv", ["calculate-Ct-from-traj.py", "-s", "x", "-f", "y", "--help_sel"])
    with pytest.raises(SystemExit) as e:
        runpy.run_path(os.path.join(DROPIN, "calculate-Ct-from-traj.py"), run_name="__main__")
    assert e.value.code == 0
    assert "MDTraj" in capsys.readouterr().out


def test_reference_module_names_resolve(monkeypatch):
    monkeypatch.syspath_prepend(DROPIN)
    for name in ("fitting_Ct_functions", "spectral_densities", "npufunc", "transforms3d_supplement", "general_maths"):
        sys.modules.pop(name, None)
    import fitting_Ct_functions as fitCt
    import general_maths as gm
    import npufunc
    import spectral_densities as sd
    import transforms3d_supplement as qs
    # the names the reference's stage scripts use from each module
    for mod, names in ((fitCt, ["autoCorrelations", "autoCorrelationModel", "read_fittedCt_parameters", "curvefit_exponential"]),
                       (sd, ["spinRelaxationExperiments", "globalRotationalDiffusion_Axisymmetric", "angularFrequencies",
                             "globalRotationalDiffusion_Isotropic", "spinRelaxationR1", "spinRelaxationNOE"]),
                       (qs, ["rotate_vector_simd", "quat_mult_simd", "quat_invert", "quat_frame_transform_min", "quat_reduce_simd",
                             "vecnorm_NDarray"]),
                       (gm, ["xyz_to_rtp", "rtp_to_xyz"]), (npufunc, ["Jomega"])):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)
    assert npufunc.Jomega.nin == 2
    for name in ("fitting_Ct_functions", "spectral_densities", "npufunc", "transforms3d_supplement", "general_maths"):
        sys.modules.pop(name, None)
