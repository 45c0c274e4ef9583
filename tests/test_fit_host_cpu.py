"""Host half of the fit stage without a GPU: `fitct.gpu_curve_fit` is replaced by scipy.optimize.curve_fit (what the
reference calls) and the CLI mirror must write the `_fittedCt.dat` that calculate-fitted-Ct.py itself wrote
(tests/golden/fit_cli.npz) -- default ladder, fixed component count, and without the fast S2 component."""
import contextlib
import io

import pytest

from test_abi_and_host import _scipy_stand_in


@pytest.mark.parametrize("tag,extra", [("ladder", []), ("nc2", ["--nc", "2"]), ("nofast", ["--nofast"])])
def test_fit_cli_reproduces_reference_file(golden, tmp_path, monkeypatch, tag, extra):
    from spinrelax_b200 import cli_fit, fitct
    monkeypatch.setattr(fitct, "gpu_curve_fit", _scipy_stand_in)
    g = golden("fit_cli.npz")
    (tmp_path / "c_Ctint.dat").write_text(str(g["ctint"]))
    with contextlib.redirect_stdout(io.StringIO()):
        cli_fit.main(["-f", str(tmp_path / "c_Ctint.dat"), "-o", str(tmp_path / tag)] + extra)
    # same solver on the same numbers: the text must be the reference's, character for character
    assert (tmp_path / (tag + "_fittedCt.dat")).read_text() == str(g[tag])
