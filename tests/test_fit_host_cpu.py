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


def test_device_solve_marshals_arguments_as_the_header_declares(monkeypatch):
    """`fitct._device_solve` against a stand-in library whose `sr_ct_fit_trf` takes its parameters in the order parsed
    from include/spinrelax_b200.h, reads the buffers through the raw pointers it is handed and writes results back
    the same way: catches a wrong argument order / dtype / shape in the ctypes call without a GPU."""
    import ctypes
    import os
    import re
    import numpy as np
    import torch
    from conftest import ROOT
    from spinrelax_b200 import _lib, fitct

    hdr = open(os.path.join(ROOT, "include", "spinrelax_b200.h")).read()
    decl = re.search(r"int\s+sr_ct_fit_trf\s*\(([^;]*)\)\s*;", hdr, re.S).group(1)
    names = [p.strip().split()[-1].lstrip("*") for p in decl.replace("\n", " ").split(",")]
    seen = {}

    def view(ptr, shape, ctype=ctypes.c_double, dtype=np.float64):
        ptr = ptr.value if hasattr(ptr, "value") else ptr
        n = int(np.prod(shape))
        return np.frombuffer((ctype * n).from_address(int(ptr)), dtype=dtype).reshape(shape)

    class FakeLib:
        def sr_ct_fit_workspace_bytes(nR, L, nP):
            return 0

        @staticmethod
        def sr_ct_fit_trf(*args):
            a = dict(zip(names, args))
            nR, L, nP = int(a["nR"]), int(a["L"]), int(a["nParams"])
            seen.update(nR=nR, L=L, nP=nP, max_nfev=int(a["max_nfev"]), ftol=float(a["ftol"]), xtol=float(a["xtol"]),
                        gtol=float(a["gtol"]), work=a["d_work"], work_bytes=int(a["work_bytes"]),
                        t=view(a["d_t"], (nR, L)).copy(), y=view(a["d_y"], (nR, L)).copy(),
                        sigma=view(a["d_sigma"], (nR, L)).copy(), p0=view(a["d_p0"], (nR, nP)).copy(),
                        lo=view(a["d_lo"], (nR, nP)).copy(), hi=view(a["d_hi"], (nR, nP)).copy())
            view(a["d_popt"], (nR, nP))[:] = seen["p0"] + 1.0
            view(a["d_R"], (nR, nP, nP))[:] = np.eye(nP)[None] * np.arange(1, nR + 1)[:, None, None]
            view(a["d_cost"], (nR,))[:] = np.arange(nR) + 0.5
            view(a["d_status"], (nR, 2), ctypes.c_int32, np.int32)[:] = [[1, 7]] * nR
            view(a["d_chi"], (nR,))[:] = np.arange(nR) * 0.25
            return 0

    class FakeTorch:
        def __getattr__(self, k):
            return getattr(torch, k)

        @staticmethod
        def device(*_):
            return torch.device("cpu")

    monkeypatch.setattr(_lib, "require_cuda", lambda: FakeTorch())
    monkeypatch.setattr(_lib, "load", lambda: FakeLib)
    monkeypatch.setattr(_lib, "current_stream_ptr", lambda: None)
    from spinrelax_b200 import multigpu
    monkeypatch.setattr(multigpu, "plan", lambda n, min_per_device=1: [(0, 0, n)])       # one device, all residues
    monkeypatch.setattr(multigpu, "run", lambda blocks, fn: [fn(*b) for b in blocks])
    rng = np.random.default_rng(3)
    nR, L, nP = 4, 11, 5
    t, y, sg = np.arange(1.0, L + 1), rng.random((nR, L)), rng.random((nR, L)) + 0.1
    p0, hi = rng.random((nR, nP)), np.full((nR, nP), 9.0)
    popt, Rf, cost, status, chi = fitct._device_solve(np.atleast_2d(t), y, sg, p0, np.zeros_like(p0), hi)
    assert np.array_equal(chi, np.arange(nR) * 0.25)
    assert (seen["nR"], seen["L"], seen["nP"]) == (nR, L, nP)
    assert seen["max_nfev"] == fitct.MAX_NFEV == 0 and (seen["ftol"], seen["xtol"], seen["gtol"]) == (1e-8, 1e-8, 1e-8)
    assert seen["work_bytes"] == 0 and not seen["work"]
    assert np.array_equal(seen["t"], np.broadcast_to(t, (nR, L))) and np.array_equal(seen["y"], y)
    assert np.array_equal(seen["sigma"], sg) and np.array_equal(seen["p0"], p0)
    assert np.array_equal(seen["lo"], np.zeros((nR, nP))) and np.array_equal(seen["hi"], hi)
    assert np.array_equal(popt, p0 + 1.0) and np.array_equal(cost, np.arange(nR) + 0.5)
    assert np.array_equal(Rf[2], 3.0 * np.eye(nP)) and np.array_equal(status, [[1, 7]] * nR)
    # and through the public wrapper: covariance from those R factors (J^T J = R^T R = 4 I for row 1) and costs
    popt2, pcov, cost2, _, _ = fitct.gpu_curve_fit(t, y, sg, p0, np.zeros_like(p0), hi)
    assert np.allclose(pcov[1], np.eye(nP) / 4.0 * (2.0 * 1.5 / (L - nP)))
