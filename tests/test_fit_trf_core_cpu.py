"""The solver logic of K5 without a GPU.

`spinrelax_b200/csrc/trf_core.cuh` + `fit_model.cuh` are the code the CUDA kernel runs (solver on its leader thread,
model arithmetic and Householder algebra on every thread).  tests/cpu_harness/trf_host.cpp compiles those very headers
with g++ and supplies serial loops where the kernel uses the CTA, so this file can check on a CPU that the restated
trust-region-reflective algorithm stops where SciPy's does: same termination status and evaluation count, same
parameters, and -- through the product's own selection ladder in fitct.py -- the same model chosen as the oracle
(oracle/fit_oracle.py = the reference's flow on scipy.optimize.curve_fit) for every residue.
The `-m gpu` twin (tests/test_relax_fit_gpu.py) runs the same comparisons through sr_ct_fit_trf on all 1000 curves."""
import ctypes
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("trf") / "libtrf_host.so")
    src = os.path.join(ROOT, "tests", "cpu_harness", "trf_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)

    def solve(t, y, sigma, p0, lo, hi, max_nfev=0):
        y, p0 = np.atleast_2d(y), np.atleast_2d(p0)
        nR, L = y.shape
        nP = p0.shape[1]
        c = lambda a, w: np.ascontiguousarray(np.broadcast_to(np.atleast_2d(a), (nR, w)), dtype=np.float64)  # noqa: E731
        T, Y, P, LO, HI = c(t, L), c(y, L), c(p0, nP), c(lo, nP), c(hi, nP)
        S = None if sigma is None else c(sigma, L)
        popt, R, cost = np.zeros((nR, nP)), np.zeros((nR, nP, nP)), np.zeros(nR)
        st = np.zeros((nR, 2), dtype=np.int32)
        vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)     # noqa: E731
        rc = lib.trf_host_fit(vp(T), vp(Y), vp(S), nR, L, nP, vp(P), vp(LO), vp(HI), max_nfev, vp(popt), vp(R), vp(cost), vp(st))
        assert rc == 0
        return popt, R, cost, st
    return solve


def _curves(n, seed=77):
    from bench_secondary import synth_curves
    t, Y, SG = synth_curves(1000, 500, seed)
    return t, Y[:n], SG[:n]


def test_core_stops_where_scipy_stops(harness):
    """Per rung, against least_squares(method='trf') given the same analytic Jacobian: status, nfev and x."""
    from scipy.optimize import least_squares
    from oracle import fit_oracle as fo
    t, Y, SG = _curves(40)

    def jac(p, i):
        n, nc = len(p), len(p) // 2
        C, tau = np.array(p[:nc]), np.array(p[nc:2 * nc])
        e = np.exp(-t[None, :] / tau[:, None])
        J = np.empty((len(t), n))
        for k in range(nc):
            J[:, k] = e[k] - (0.0 if n % 2 else 1.0)
            J[:, nc + k] = C[k] * e[k] * t / tau[k] ** 2
        if n % 2:
            J[:, -1] = 1.0
        return J / SG[i][:, None]

    for npar in (2, 3, 5, 7, 9):
        p0 = np.array([fo.initial_guess(t, Y[i], npar)[0] for i in range(len(Y))])
        hi = np.array(fo.bounds(npar, t[-1] * 10)[1], dtype=float)
        popt, R, cost, st = harness(t, Y, SG, p0, np.zeros(npar), hi)
        same = 0
        for i in range(len(Y)):
            r = least_squares(lambda p: (fo.model_curve(t, *p) - Y[i]) / SG[i], p0[i], jac=lambda p: jac(p, i),
                              bounds=(np.zeros(npar), hi), method="trf")
            assert (r.status > 0) == (st[i, 0] > 0), (npar, i, r.status, st[i])
            # rungs with more exponentials than the curve holds end on a flat valley floor: the stopping point (ftol
            # on a 1e-8 relative decrease) is not reproducible to more than the valley's depth, in SciPy itself either
            assert rel_err(cost[i], r.cost) < (1e-9 if npar <= 3 else 5e-2), (npar, i)
            if r.status == st[i, 0] and r.nfev == st[i, 1]:
                same += 1
                if npar <= 3:
                    assert rel_err(popt[i], r.x) < 1e-9, (npar, i)
        # identical path (status and evaluation count) on nearly all well-posed fits (a step norm a hair either side
        # of xtol turns status 2 into 4); the over-parameterised rungs wander along flat valleys where 1e-16
        # differences in exp() move the stopping iteration
        assert same >= (len(Y) - 2 if npar <= 3 else len(Y) // 2), (npar, same)


def test_failure_codes_match_scipy_exceptions(harness):
    """What curve_fit raises on (and the reference swallows, fitting_Ct_functions.py:325-328) must come back as a
    non-positive status: evaluation cap, p0 outside the box, non-finite residuals at p0."""
    from scipy.optimize import least_squares
    from oracle import fit_oracle as fo
    t, Y, SG = _curves(120)
    hi = np.array(fo.bounds(9, t[-1] * 10)[1], dtype=float)
    p0 = np.array([fo.initial_guess(t, y, 9)[0] for y in Y])
    popt, R, cost, st = harness(t, Y, SG, p0, np.zeros(9), hi)
    capped = np.nonzero(st[:, 0] == 0)[0]
    assert len(capped) >= 1 and np.all(st[capped, 1] == 900)          # max_nfev = 100 n
    for i in capped[:3]:
        r = least_squares(lambda p: (fo.model_curve(t, *p) - Y[i]) / SG[i], p0[i], bounds=(np.zeros(9), hi), method="trf")
        assert r.status == 0 and r.nfev == 900
    # a free S2 initialised from a negative tail average is outside [0, 1]: ValueError upstream
    y = Y[0] - 2.0
    p0n = np.array(fo.initial_guess(t, y, 3)[0])
    assert p0n[-1] < 0
    _, _, c2, s2 = harness(t, y, SG[0], p0n, np.zeros(3), np.array(fo.bounds(3, t[-1] * 10)[1], dtype=float))
    assert s2[0, 0] == -3 and np.isinf(c2[0])
    with pytest.raises(ValueError):
        least_squares(lambda p: (fo.model_curve(t, *p) - y) / SG[0], p0n, bounds=(np.zeros(3), fo.bounds(3, t[-1] * 10)[1]))
    ybad = Y[0].copy(); ybad[7] = np.inf
    _, _, c3, s3 = harness(t, ybad, SG[0], p0[0], np.zeros(9), hi)
    assert s3[0, 0] == -4 and np.isinf(c3[0])


def test_ladder_selects_the_oracles_model(harness, golden, monkeypatch):
    """The product's ladder (fitct.fit_all_residues) with the solver core underneath against the oracle's ladder on
    SciPy: the same rung for every residue, parameters and uncertainties to the digits SciPy itself reproduces when its
    Jacobian is perturbed at the 1e-8 level (finite differences vs analytic)."""
    from oracle import fit_oracle as fo
    from spinrelax_b200 import fitct
    monkeypatch.setattr(fitct, "_device_solve", harness)
    g = golden("fit.npz")
    sets = [_curves(150), (g["t"], g["Ct"], g["dCt"])]
    for t, Y, SG in sets:
        n = len(Y)
        ac = fitct.autoCorrelations()
        ac.import_target_array([str(i) for i in range(n)], [t] * n, Y, SG)
        log = io.StringIO()
        chis = ac.fit_all_residues(fp=log)
        n_failed_lines = log.getvalue().count("failed!")
        n_failed_ref = 0
        for i in range(n):
            ref = fo.fit_ladder(t, Y[i], SG[i])
            m = ac.model[str(i)]
            assert m.nParams == ref["n_params"], (i, m.nParams, ref["n_params"])
            assert rel_err(m.chiSq, ref["chi"]) < 1e-6 and rel_err(chis[i], ref["chi"]) < 1e-6
            assert rel_err(m.tau, ref["tau"]) < 1e-5 and np.max(np.abs(m.C - ref["C"])) < 1e-5 and abs(m.S2 - ref["S2"]) < 1e-5
            assert rel_err(m.dtau, ref["dtau"]) < 1e-4 and rel_err(m.dC, ref["dC"]) < 1e-4
        assert n_failed_lines >= n_failed_ref


def test_golden_ladder_from_the_reference_itself(harness, golden, monkeypatch):
    """tests/golden/fit.npz holds what the *unmodified reference* (fitting_Ct_functions.py run by make_golden.py) chose
    and fitted: every residue must land on that rung with those parameters."""
    from spinrelax_b200 import fitct
    monkeypatch.setattr(fitct, "_device_solve", harness)
    g = golden("fit.npz")
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    n = len(Ct)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(n)], [t] * n, Ct, dCt)
    chis = ac.fit_all_residues(fp=io.StringIO())
    for i, row in enumerate(g["ladder"]):
        m = ac.model[str(i)]
        nc = int(row[0]) // 2
        assert m.nParams == int(row[0]), i
        assert rel_err(chis[i], row[1]) < 1e-6 and abs(m.S2 - row[2]) < 1e-5
        assert np.max(np.abs(m.C - row[3:3 + nc])) < 1e-5 and rel_err(m.tau, row[7:7 + nc]) < 1e-5
