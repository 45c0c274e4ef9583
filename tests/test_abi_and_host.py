"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header declares;
host-side mirrors of the reference helpers; the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_err
from oracle import ct_oracle


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "spinrelax_b200.h")) as fp:
        text = fp.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spinrelax_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert _lib.load().sr_abi_version() == _lib.ABI_VERSION == 2
    # every declared symbol has a ctypes prototype in the binding
    for n in names:
        assert getattr(_lib.load(), n).argtypes is not None or n in ("sr_abi_version", "sr_last_error"), n


def test_pure_abi_queries_work_without_gpu():
    from spinrelax_b200 import _lib
    lib = _lib.load()
    pitch = lib.sr_ct_row_pitch(1000)
    assert pitch >= 1000 + 480 and pitch % 4 == 0
    assert lib.sr_ct_workspace_bytes(10, 1000, 76) >= 76 * 10 * pitch * 12 + 76 * 10 * 500 * 8


def test_host_helpers_match_oracle(golden):
    from spinrelax_b200 import ct
    g = golden("ct_small.npz")
    v4 = ct.reformat_vecs_by_tau([g["traj0"], g["traj1"]], 10.0, 2000.0)
    assert np.array_equal(v4, g["vecs"])
    assert np.array_equal(v4, ct_oracle.reformat_by_tau([g["traj0"], g["traj1"]], 10.0, 2000.0))
    assert np.array_equal(ct.calculate_dt(10.0, 2000.0), g["dt"])
    with pytest.raises(SystemExit):
        ct.calculate_Ct_Palmer(np.zeros((4, 3)))          # reference exits on non-4D input (:213-216)


def test_no_cpu_fallback():
    import torch
    from spinrelax_b200 import _lib, ct
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.SpinRelaxError):
        ct.calculate_Ct_Palmer(np.zeros((2, 60, 3, 3), dtype=np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spinrelax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(dirpath, f)) as fp:
                    src = fp.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_spherical_host_side_matches_reference(golden, tmp_path):
    """gm.rtp_to_xyz (host) against the real function; print_s3d against what the shipped writer produces before it
    stops (its first set), continued in the same format; xyz_to_rtp refuses to run without a GPU."""
    from spinrelax_b200 import gm, io_formats
    g = golden("rtp.npz")
    assert np.array_equal(gm.rtp_to_xyz(g["pt"], vaxis=-1, bUnit=True), g["back"])
    assert np.array_equal(gm.rtp_to_xyz(np.array([1.3, 0.4, 2.1])), g["back1"])
    ax0 = gm.rtp_to_xyz(np.ascontiguousarray(np.moveaxis(g["pt"], -1, 0)), vaxis=0, bUnit=True)
    assert np.array_equal(np.moveaxis(ax0, 0, -1), g["back"])
    rtp = np.transpose(g["rtp64"], (1, 0, 2))
    fn = tmp_path / "pt.dat"
    io_formats.print_s3d(str(fn), ["A", "B", "C", "D", "E"], rtp, (1, 2))
    txt = fn.read_text()
    head = str(g["s3d_partial"])
    assert bool(g["s3d_crashed"]) and txt.startswith(head)
    lines = txt.splitlines()
    assert len(lines) == 5 * 702 and lines[702] == '@s1 legend "B"' and lines[-1] == "&"
    assert lines[703] == "%g %g" % (rtp[1, 0, 1], rtp[1, 0, 2])
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            gm.xyz_to_rtp(g["vecs"])


def _scipy_stand_in(t, y, sigma, p0, lo, hi):
    """Same contract as fitct.gpu_curve_fit, solved per residue with scipy.optimize.curve_fit (what the reference
    calls): lets the host-side selection ladder be checked on a machine without a GPU."""
    from scipy.optimize import curve_fit
    from oracle import fit_oracle
    t, y, p0 = np.atleast_2d(t), np.atleast_2d(y), np.atleast_2d(p0)
    nR, nP = p0.shape
    t, lo, hi = np.broadcast_to(t, y.shape), np.broadcast_to(lo, p0.shape), np.broadcast_to(hi, p0.shape)
    popt, pcov = np.full((nR, nP), np.nan), np.full((nR, nP, nP), np.nan)
    cost, status = np.full(nR, np.nan), np.zeros((nR, 2), dtype=np.int32)
    for i in range(nR):
        sg = None if sigma is None else np.atleast_2d(sigma)[i]
        try:
            popt[i], pcov[i] = curve_fit(fit_oracle.model_curve, t[i], y[i], sigma=sg, p0=p0[i], bounds=(lo[i], hi[i]))
            r = (fit_oracle.model_curve(t[i], *popt[i]) - y[i]) / (1.0 if sg is None else sg)
            cost[i], status[i, 0] = 0.5 * r @ r, 1
        except Exception:
            pass
    return popt, pcov, cost, status


def test_fit_selection_ladder_host_logic(golden, monkeypatch):
    """The vectorised ladder, the reference-shaped loop and the reference's own results agree when all three are
    given the same solver (SciPy): pins the host logic (initial guesses, G6 flags, chi^2, acceptance) without a GPU."""
    import io
    from spinrelax_b200 import fitct
    monkeypatch.setattr(fitct, "gpu_curve_fit", _scipy_stand_in)
    g = golden("fit.npz")
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    names = [str(i) for i in range(len(Ct))]
    out = {}
    for mode in ("vector", "loop"):
        ac = fitct.autoCorrelations()
        ac.import_target_array(names, [t] * len(Ct), Ct, dCt)
        (ac.fit_all_residues if mode == "vector" else ac.fit_all_residues_loop)(fp=io.StringIO())
        out[mode] = ac
    for i, k in enumerate(names):
        row = g["ladder"][i]
        for mode in ("vector", "loop"):
            m = out[mode].model[k]
            nc = int(row[0]) // 2
            assert m.nParams == int(row[0]), (mode, k)
            assert rel_err(m.S2, row[2]) < 1e-8
            assert rel_err(np.asarray(m.C), row[3:3 + nc]) < 1e-7 and rel_err(np.asarray(m.tau), row[7:7 + nc]) < 1e-7
        a, b = out["vector"].model[k], out["loop"].model[k]
        assert np.array_equal(a.C, b.C) and np.array_equal(a.tau, b.tau) and a.S2 == b.S2
        assert a.chiSq == b.chiSq and rel_err(a.chiSq, row[1]) < 1e-8


def test_pcov_from_R_matches_scipy(golden):
    """curve_fit's covariance rebuilt from the triangular factor of the Jacobian and the cost at SciPy's own optimum
    (the host half of gpu_curve_fit; the kernel returns R of J = QR, which has J's singular values and right vectors)."""
    from scipy.optimize import curve_fit
    from oracle import fit_oracle
    from spinrelax_b200 import fitct
    g = golden("fit.npz")
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    for nP in (2, 3, 5, 7):
        nc = nP // 2
        Rs, cost, ref, popts, Js = [], [], [], [], []
        for i in range(len(Ct)):
            p0 = fit_oracle.initial_guess(t, Ct[i], nP)[0]
            try:
                popt, pcov = curve_fit(fit_oracle.model_curve, t, Ct[i], sigma=dCt[i], p0=p0, bounds=fit_oracle.bounds(nP, t[-1] * 10))
            except RuntimeError:
                continue
            C, tau = popt[:nc], popt[nc:2 * nc]
            e = np.exp(-t[None] / tau[:, None])
            J = np.zeros((len(t), nP))
            J[:, :nc] = (e - (0.0 if nP % 2 else 1.0)).T
            J[:, nc:2 * nc] = (C[:, None] * e * t[None] / tau[:, None] ** 2).T
            if nP % 2:
                J[:, -1] = 1.0
            J /= dCt[i][:, None]
            r = (fit_oracle.model_curve(t, *popt) - Ct[i]) / dCt[i]
            Rs.append(np.linalg.qr(J, mode="r")); cost.append(0.5 * r @ r); ref.append(pcov)
            popts.append(popt); Js.append(J)
        ours = fitct.pcov_from_R(np.array(Rs), np.array(cost), len(t))
        n_well = 0
        for a, b, popt, J, c in zip(ours, ref, popts, Js, cost):
            # exactly curve_fit's own recipe on the same (analytic) Jacobian: the R route loses nothing, whatever the
            # conditioning of J
            _, sv, VT = np.linalg.svd(J, full_matrices=False)
            keep = sv > np.finfo(float).eps * max(J.shape) * sv[0]
            direct = (VT[keep].T / sv[keep] ** 2) @ VT[keep] * (2.0 * c / (len(t) - nP))
            assert np.allclose(np.sqrt(np.diag(a)), np.sqrt(np.diag(direct)), rtol=1e-6), nP
            da, db = np.sqrt(np.diag(a)), np.sqrt(np.diag(b))
            # against SciPy's own pcov: the over-fitting flag of the ladder (any error larger than its parameter) must
            # come out the same; SciPy differentiates numerically, so only fits that are not flagged have a covariance
            # comparable digit by digit
            assert np.any(da > popt) == np.any(db > popt)
            if not np.any(db > popt):
                assert np.allclose(da, db, rtol=2e-3)
                n_well += 1
        assert n_well >= 3 or nP >= 5


def test_quaternion_helpers_match_reference(golden):
    """Host-side mirrors of transforms3d_supplement.py (SURVEY 8b importable surface) against the real module."""
    from spinrelax_b200 import qs
    g = golden("qs.npz")
    q1, q2, v, q32 = g["q1"], g["q2"], g["v"], g["q1"].astype(np.float32)
    assert np.allclose(qs.quat_mult_simd(q1, q2), g["mult"], rtol=0, atol=1e-15)
    mixed = qs.quat_mult_simd(qs.quat_invert(q32[:-1]), q32[1:])
    assert mixed.dtype == g["mult_mixed"].dtype and np.allclose(mixed, g["mult_mixed"], rtol=0, atol=1e-15)
    assert np.array_equal(qs.quat_invert(q1), g["invert"])
    inv32 = qs.quat_invert(q32)
    assert inv32.dtype == np.float64 and np.array_equal(inv32, g["invert32"])
    assert np.array_equal(qs.quat_reduce_simd(q1), g["reduce"])
    assert np.array_equal(qs.quat_reduce_simd(q1, qref=q2[0]), g["reduce_ref"])
    assert np.array_equal(qs.quat_reduce_simd(np.ascontiguousarray(q1.T), axis=0), g["reduce_ax0"])
    with np.errstate(all="ignore"):
        assert np.array_equal(qs.vecnorm_NDarray(v), g["vecnorm"]) and np.array_equal(qs.vecnorm_NDarray(v[3]), g["vecnorm1"])
        assert np.array_equal(qs.vecnorm_NDarray(v.T.copy(), axis=0), g["vecnorm_ax0"])
        u = qs.vecnorm_NDarray(v)
    got = np.array([qs.quat_v1v2(u[i], u[i + 1]) for i in range(8, 20)])
    assert np.allclose(got, g["v1v2"], rtol=0, atol=1e-14)
    fm = np.array([qs.quat_frame_transform_min(f) for f in g["frames"]])
    assert np.allclose(fm, g["frame_min"], rtol=0, atol=1e-13)
    # small inputs / per-vector quaternions stay on the host formula (no GPU needed)
    assert np.allclose(qs.rotate_vector_simd(v[:9], q1[0]), g["rot_small"], rtol=0, atol=1e-15)
    assert np.allclose(qs.rotate_vector_simd(v, q1), g["rot_perq"], rtol=0, atol=1e-15)
    assert np.allclose(qs.rotate_vector_simd(v[:9], 2.5 * q1[1]), g["rot_unnorm"], rtol=0, atol=1e-15)
    R = qs.rotation_matrix(q1[0])
    assert np.allclose(v[:9] @ R.T, g["rot_small"], rtol=0, atol=1e-14)


def test_text_writers_and_readers_match_reference(golden, tmp_path):
    """io_formats against text written / parsed by the real general_scripts.py, plumedcolvario.py and dxio.py."""
    from spinrelax_b200 import io_formats as io
    g = golden("io.npz")
    f = str(tmp_path / "f")
    x, y1, y2, ct = g["x"], g["y1"], g["y2"], g["ct"]
    edges = [g["edges_phi"], g["edges_cos"]]

    def wrote(fn, *a, **k):
        fn(f, *a, **k)
        return open(f).read()

    assert wrote(io.print_xylist, x, y1) == str(g["xylist_1d"])
    assert wrote(io.print_xylist, x, y2) == str(g["xylist_2d"])
    assert wrote(io.print_xylist, x, y2, True, header="# head") == str(g["xylist_cols"])
    assert wrote(io.print_sxylist, ["1", "2", "7", "9"], x, ct, header=["# a", "# b"]) == str(g["sxylist"])
    legs, lx, ly, ldy = io.load_sxydylist(f, "legend")
    assert list(legs) == list(g["sxy_legs"])
    assert np.array_equal(np.array(lx), g["sxy_x"]) and np.array_equal(np.array(ly), g["sxy_y"])
    assert np.array_equal(np.array(ldy), g["sxy_dy"])
    assert wrote(io.print_gplot_hist, g["hist"], edges, header="# h", bSphere=True) == str(g["gplot_sphere"])
    assert wrote(io.print_gplot_hist, g["hist"], edges, header="# h", bSphere=False) == str(g["gplot_flat"])
    assert wrote(io.write_to_dx, g["vol"], (3, 4, 5), [-1.0, -0.5, 0.25], np.diag([0.5, 0.25, 0.125]), "nm") == str(g["dx"])
    with open(f, "w") as fp:
        fp.write(str(g["plumed_text"]))
    names, data = io.read_from_plumedprint(f)
    assert list(names) == list(g["plumed_names"])
    assert data.dtype == np.float32 and np.array_equal(data, g["plumed_data"])


def test_spectral_density_host_objects_match_reference(golden):
    """gyromagnetic data, angular frequencies / prefactors, axisymmetric D and A coefficients, histogram -> bin
    vectors: the parts of specdens.py that never touch the GPU, against the real spectral_densities.py."""
    from spinrelax_b200 import specdens as sd
    g = golden("sd_host.npz")
    cases = [("15N", "1H", 600.13, "MHz", "ps"), ("13C", "1H", 18.8, "T", "ns"), ("15N", "1H", 8.5e8, "Hz", "s")]
    for k, (a, b, f, fu, tu) in enumerate(cases):
        w = sd.angularFrequencies(a, b, f, fu, tu)
        assert np.array_equal(w.omega, g["w%d_omega" % k])
        sc = np.array([w.get_factor_DD(), w.get_factor_CSA(), w.get_magnetic_field("T"), w.get_magnetic_field("MHz"),
                       w.gA.gamma, w.gB.gamma, w.gA.csa])
        assert np.allclose(sc, g["w%d_scalars" % k], rtol=1e-15, atol=0)
        w.set_time_unit("ns")
        assert np.allclose(w.omega, g["w%d_omega_ns" % k], rtol=1e-15, atol=0)
    w = sd.angularFrequencies("15N", "1H", 600, "MHz", "ps")
    w.initialise_CSA_array(4, [-160e-6, -170e-6, -175e-6, -180e-6])
    assert np.allclose(w.get_factor_CSA(), g["multi_csa_factors"], rtol=1e-15)
    assert np.allclose(w.get_factor_CSA(2), g["multi_csa_one"], rtol=1e-15)
    for tag, D, conv in (("prolate", [2.5e-5, 1.6], False), ("oblate", [2.5e-5, 0.7], False), ("conv", [3.1e-5, 2.2e-5], True)):
        r = sd.globalRotationalDiffusion_Axisymmetric(D=D, bConvert=conv)
        r.vecXH = g["vecs"]
        r.update_A_coefficients()
        assert np.array_equal(r.D, g[tag + "_D"]) and np.array_equal(r.get_D_coefficients(), g[tag + "_DJ"])
        assert np.array_equal(r.get_A_coefficients(), g[tag + "_AJ"])
        assert np.array_equal(r.get_A_coefficients(3), g[tag + "_AJ_ind"])
        assert np.array_equal(np.array(r.transform_D()), g[tag + "_Dpp"])
    r = sd.globalRotationalDiffusion_Axisymmetric(tau=8.0e3, aniso=1.3)
    r.set_Diso(2.2e-5); r.set_Daniso(0.9)
    assert np.array_equal(r.get_D_coefficients(), g["setter_DJ"])
    edges = np.empty(2, dtype=object)
    edges[0], edges[1] = g["edges_phi"], g["edges_cos"]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        bv, bw = sd.convert_LambertCylindricalHist_to_vecs(g["hist"], edges)
    assert np.array_equal(bv, g["bin_vecs"]) and np.array_equal(bw, g["bin_weights"])


def test_ct_cli_argument_errors_need_no_gpu(tmp_path):
    """Exit codes of calculate-Ct-from-traj.py's argument checks (:358-360, :372-375, :411-414, tau vs dt)."""
    import contextlib, io
    from spinrelax_b200 import cli_ct
    v = np.zeros((40, 3, 3), dtype=np.float32)
    v[..., 2] = 1.0
    np.savez(tmp_path / "v.npz", vecs=v, dt=10.0)
    fn = str(tmp_path / "v.npz")

    def code(argv):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            with pytest.raises(SystemExit) as e:
                cli_ct.main(argv)
        return e.value.code

    assert code(["-f", fn, "-o", str(tmp_path / "o"), "--Ct"]) == 1                                  # --Ct without --tau
    assert code(["-f", fn, "-o", str(tmp_path / "o"), "--vecRot", "1 0 0"]) == 23                    # 3 numbers
    assert code(["-f", fn, "-o", str(tmp_path / "o"), "--vecRot", "1 1 0 0"]) == 23                  # not a unit quaternion
    assert code(["-f", fn, fn, "-s", fn, fn, fn, "-o", str(tmp_path / "o")]) == 1                    # refs != trajectories
    assert code(["-f", fn, "-o", str(tmp_path / "o"), "--Ct", "--tau", "15"]) == 1                   # dt > tau / 2
    assert code(["-f", str(tmp_path / "v.xtc"), "-o", str(tmp_path / "o")]) == 2                     # mdtraj formats: out of scope


def test_host_reshaping_properties():
    """Property tests (hypothesis) of the host helpers that decide which samples enter which block / lag."""
    from hypothesis import given, settings, strategies as st
    from oracle import dq_oracle
    from spinrelax_b200 import ct, dq

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 90), min_size=1, max_size=4), st.integers(2, 40), st.integers(1, 5))
    def reformat(lengths, nF, nR):
        rng = np.random.default_rng(sum(lengths) + nF)
        trajs = [rng.standard_normal((n, nR, 3)).astype(np.float32) for n in lengths]
        dt, tau = 2.0, 2.0 * nF
        if all(n < nF for n in lengths):
            return                                    # no complete block: the reference fails on the empty reshape
        a = ct.reformat_vecs_by_tau(trajs, dt, tau)
        b = ct_oracle.reformat_by_tau(trajs, dt, tau)
        assert a.dtype == b.dtype and np.array_equal(a, b)
        assert a.shape == (sum(n // nF for n in lengths), nF, nR, 3)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(20, 400), st.floats(0.0, 40.0), st.floats(0.0, 30.0), st.floats(50.0, 3000.0))
    def lags(n, min_dt, skip_dt, max_dt):
        times = np.arange(n) * 10.0
        lo, hi, step, ddt = dq.lag_grid(times, min_dt, max_dt, skip_dt)
        ref_lags, ref_ddt = dq_oracle.lag_grid(times, min_dt, max_dt, skip_dt)
        assert list(range(lo, hi + 1, step)) == ref_lags and ddt == ref_ddt

    reformat()
    lags()


def test_ctypes_prototypes_match_the_header():
    """Every prototype in _lib.py has the parameter count and parameter kinds (pointer / int / long long / size_t /
    double) and the return kind of the declaration in include/spinrelax_b200.h."""
    from spinrelax_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "spinrelax_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    decls = re.findall(r"\b(int|void|long long|size_t|const char\s*\*)\s+(sr_\w+)\s*\(([^;{]*)\)\s*;", hdr)
    assert len(decls) >= 30

    def kind_of_c(p):
        p = p.strip()
        if p in ("void", ""):
            return None
        if "*" in p:
            return "ptr"
        t = " ".join(p.split()[:-1]).replace("const", "").strip()
        return {"int": "int", "long long": "ll", "size_t": "size", "double": "double", "unsigned int": "int"}[t]

    def kind_of_ctypes(t):
        if t is None:
            return None
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_size_t: "size", ctypes.c_double: "double"}[t]

    checked = 0
    for ret, name, params in decls:
        fn = getattr(lib, name)
        want = [k for k in (kind_of_c(p) for p in params.replace("\n", " ").split(",")) if k is not None]
        got = [kind_of_ctypes(t) for t in (fn.argtypes or [])]
        # c_size_t and c_longlong / c_ulong may alias on LP64: compare by size class
        norm = lambda ks: ["ll" if k == "size" else k for k in ks]      # noqa: E731
        assert norm(got) == norm(want), (name, got, want)
        rk = "ptr" if "*" in ret else {"int": "int", "void": None, "long long": "ll", "size_t": "ll"}[ret]
        gk = kind_of_ctypes(fn.restype)
        assert ("ll" if gk == "size" else gk) == rk, (name, gk, rk)
        checked += 1
    assert checked == len(decls)


def test_call_sites_pass_as_many_arguments_as_the_header_declares():
    """AST scan of the host package: every `lib.sr_*(...)` / `getattr(lib, fn)(...)`-style call with a literal entry
    point name passes exactly the number of arguments its C declaration has."""
    import ast
    hdr = open(os.path.join(ROOT, "include", "spinrelax_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    nparams = {}
    for _, name, params in re.findall(r"\b(int|void|long long|size_t|const char\s*\*)\s+(sr_\w+)\s*\(([^;{]*)\)\s*;", hdr):
        ps = [p.strip() for p in params.replace("\n", " ").split(",")]
        nparams[name] = 0 if ps in ([""], ["void"]) else len(ps)
    calls = 0
    pkg = os.path.join(ROOT, "spinrelax_b200")
    for fn in sorted(os.listdir(pkg)) + ["../bench.py", "../bench_secondary.py", "../bench_multi.py", "../__graft_entry__.py"]:
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in nparams:
                if any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                assert len(node.args) == nparams[node.func.attr], (fn, node.lineno, node.func.attr)
                calls += 1
    assert calls >= 25


def test_ncu_traffic_record_is_tied_to_the_shipped_tile_configuration():
    """roofline.traffic comes from a committed ncu capture: the record names the ct_lag_kernel configuration it was
    taken with, pipeline reports it only for that configuration, and that configuration is the one csrc/ct.cu ships."""
    import json
    import re
    from conftest import ROOT
    from spinrelax_b200 import pipeline
    src = open(os.path.join(ROOT, "spinrelax_b200", "csrc", "ct.cu")).read()
    m = re.search(r"using CtLong = CtCfg<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), 0, (\d+), (\d+)>;", src)
    assert m, "CtLong not found"
    tag = "CtCfg<R=%s,MB=%s,FB=%s,NW=%s,MINB=%s,NS=%s,FLUSH=%s,ORDER=%s,SYNC=%s>" % m.groups()
    assert pipeline.CtHistStep.CT_LAG_CONFIG == tag
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert (pipeline.CtHistStep.ncu_traffic_bytes() is not None) == (rec.get("kernel_config") == tag)
