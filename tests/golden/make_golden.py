"""Generate the committed golden vectors by running the REAL SpinRelax code (build container only).

    python tests/golden/make_golden.py

Needs /root/reference and oracle/_ref/npufunc (cd oracle && make).  Inputs are seeded synthetic data
from spinrelax_b200.synth; inputs small enough are stored next to the reference outputs so the tests
never depend on regenerating them bit-for-bit.  The reference has no tests or fixtures of its own
(SURVEY.md section 4), so these files are the pin for oracle/ and for the CUDA path.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from spinrelax_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print("%-28s %8.1f kB" % (name, os.path.getsize(path) / 1e3))


def golden_ct(refct):
    # small, ragged-friendly case: 2 trajectories with different frame counts -> reformat -> C(t)
    trajs = [synth.nh_vectors(1130, 6, seed=synth.BASE_SEED + 1), synth.nh_vectors(877, 6, seed=synth.BASE_SEED + 2)]
    with quiet():
        v4 = refct.reformat_vecs_by_tau(trajs, 10.0, 2000.0)          # -> (9, 200, 6, 3)
        Ct32, dCt32 = refct.calculate_Ct_Palmer(v4)
        Ct64, dCt64 = refct.calculate_Ct_Palmer(v4.astype(np.float64))
        dt = refct.calculate_dt(10.0, 2000.0)
    save("ct_small.npz", traj0=trajs[0], traj1=trajs[1], vecs=v4, Ct32=Ct32, dCt32=dCt32, Ct64=Ct64, dCt64=dCt64,
         dt=dt)
    # BASELINE config 1 shape (76 vectors, 10 chunks x 1000 frames); input regenerated from the seed,
    # outputs stored for a lag subset to stay small, plus a checksum of the input
    v = synth.nh_vectors(10000, 76, seed=synth.BASE_SEED + 11)
    with quiet():
        v4 = refct.reformat_vecs_by_tau([v], 10.0, 10000.0)
        Ct64, dCt64 = refct.calculate_Ct_Palmer(v4.astype(np.float64))
        Ct32, dCt32 = refct.calculate_Ct_Palmer(v4)
    lags = np.unique(np.concatenate((np.arange(1, 33), np.arange(40, 500, 23), [479, 480, 481, 499, 500])))
    save("ct_config1.npz", seed=synth.BASE_SEED + 11, shape=np.array(v4.shape), lags=lags,
         input_sum=np.float64(v.astype(np.float64).sum()), input_head=v[:4],
         Ct64=Ct64[lags - 1], dCt64=dCt64[lags - 1], Ct32=Ct32[lags - 1], dCt32=dCt32[lags - 1])
    # known-answer inputs: static vectors, single chunk (dCt = 0/0), odd frame count
    stat = np.repeat(synth.nh_vectors(1, 3, seed=5)[None], 60, axis=1).reshape(1, 60, 3, 3).repeat(2, axis=0)
    one = synth.nh_vectors(101, 2, seed=9).reshape(1, 101, 2, 3)
    with quiet(), np.errstate(all="ignore"):
        a = refct.calculate_Ct_Palmer(stat.astype(np.float64))
        b = refct.calculate_Ct_Palmer(one.astype(np.float64))
    save("ct_edge.npz", static=stat, static_Ct=a[0], static_dCt=a[1], one=one, one_Ct=b[0], one_dCt=b[1])


def _away_from_edges(v, q, nbx, margin):
    """Drop samples whose phi or cos(theta) lies within `margin` of a bin edge, so that the committed
    counts do not depend on the last ulp of the host libm (arctan2/arccos/cos differ between CPUs)."""
    qs = ref_loader.module("transforms3d_supplement")
    gm = ref_loader.module("general_maths")
    w = qs.rotate_vector_simd(v, q) if q is not None else v
    rtp = gm.xyz_to_rtp(w.astype(np.float64))
    phi, cth = rtp[..., 1], np.cos(rtp[..., 2])
    ephi = np.linspace(-np.pi, np.pi, nbx + 1)
    ecth = np.linspace(-1, 1, nbx // 2 + 1)
    dphi = np.min(np.abs(phi[..., None] - ephi), axis=-1)
    dcth = np.min(np.abs(cth[..., None] - ecth), axis=-1)
    ok = np.all((dphi > margin) & (dcth > margin), axis=1)
    return v[ok]


class _FakeTopology:
    """Stand-in for mdtraj's Topology.select: named index lists (mdtraj itself is absent from the image)."""

    def __init__(self, sel):
        self.sel = sel

    def select(self, txt):
        return self.sel[txt]


class _FakeTraj:
    def __init__(self, xyz, sel):
        self.xyz, self.topology = xyz, _FakeTopology(sel)


def golden_traj(refct):
    # the REAL obtain_XHvecs (calculate-Ct-from-traj.py:64-86) on a stand-in trajectory object
    xyz, sel, ref = synth.backbone_trajectory(40, 12, seed=synth.BASE_SEED + 21)
    xyz[3, sel["name H"][2]] = xyz[3, sel["name N and not resname PRO"][2]]      # zero vector -> nan_to_num path
    with quiet():
        vec = refct.obtain_XHvecs(_FakeTraj(xyz, sel), "name H", "name N and not resname PRO")
    save("traj.npz", xyz=xyz, ref=ref, indexH=sel["name H"], indexX=sel["name N and not resname PRO"],
         fit=sel["custom occupancy"], vecXH=vec)


def golden_ct_cli(refct=None):
    """The output files of calculate-Ct-from-traj.py for `--tau --Ct --vecAvg --S2 --vecHist --binary --vecRot q`,
    produced by chaining the REAL functions and writers exactly as its __main__ does (:514-646; the script itself
    needs mdtraj to load a trajectory).  Two trajectories of unequal length, fitted and unfitted vectors."""
    import tempfile
    refct = refct or ref_loader.script("calculate-Ct-from-traj.py")
    qs = ref_loader.module("transforms3d_supplement")
    gs = ref_loader.module("general_scripts")
    gm = ref_loader.module("general_maths")
    q = np.array([0.83, -0.31, 0.22, 0.41]); q = q / np.linalg.norm(q)
    dt, tau, zeta = 10.0, 600.0, (1.02 / 1.04) ** 6
    fitA, fitB = synth.nh_vectors(190, 4, seed=synth.BASE_SEED + 71), synth.nh_vectors(131, 4, seed=synth.BASE_SEED + 72)
    tum = synth.quaternion_walk(190, seed=synth.BASE_SEED + 73, sigma=(0.02, 0.02, 0.02)).astype(np.float64)
    extA = np.array([qs.rotate_vector_simd(fitA[i].astype(np.float64), tum[i]) for i in range(190)]).astype(np.float32)
    extB = np.array([qs.rotate_vector_simd(fitB[i].astype(np.float64), tum[i]) for i in range(131)]).astype(np.float32)
    names = [3, 4, 7, 9]
    out = {}
    with tempfile.TemporaryDirectory() as td, quiet(), np.errstate(all="ignore"):
        vecXH = refct.reformat_vecs_by_tau([extA, extB], dt, tau)
        vecXHfit = refct.reformat_vecs_by_tau([fitA, fitB], dt, tau)
        tt = refct.calculate_dt(dt, tau)
        Ct, dCt = refct.calculate_Ct_Palmer(vecXH)
        gs.print_sxylist(td + "/Ctext", names, tt, np.stack((Ct.T, dCt.T), axis=-1)); out["Ctext"] = open(td + "/Ctext").read()
        Ct, dCt = refct.calculate_Ct_Palmer(vecXHfit)
        gs.print_sxylist(td + "/Ctint", names, tt, np.stack((Ct.T, dCt.T), axis=-1)); out["Ctint"] = open(td + "/Ctint").read()
        sh = vecXHfit.shape
        v3 = vecXHfit.reshape((sh[0] * sh[1], sh[-2], sh[-1]))
        v3 = qs.rotate_vector_simd(v3, q)
        avg = gs.normalise_vector_array(np.mean(v3, axis=0))
        gs.print_xylist(td + "/avg", names, np.array(avg).T, True); out["avgvec"] = open(td + "/avg").read()
        rtp = np.transpose(gm.xyz_to_rtp(v3), axes=(1, 0, 2))
        rtp = np.delete(rtp, 0, axis=2)
        rtp[..., 1] = np.cos(rtp[..., 1])
        hl = np.zeros((4, 72, 36), dtype=rtp.dtype)
        for i in range(4):
            hl[i], edges = np.histogramdd(rtp[i], bins=(72, 36), range=((-np.pi, np.pi), (-1, 1)))
        S2 = refct.calculate_S2_by_outerProduct(v3, dt, tau)
        gs.print_xylist(td + "/S2", names, (S2.T) * zeta, True); out["S2"] = open(td + "/S2").read()
    save("ct_cli.npz", q=q, fitA=fitA, fitB=fitB, extA=extA, extB=extB, names=np.array(names), hist=hl.astype(np.int64),
         edges_phi=edges[0], edges_cos=edges[1], **{k: np.array(v) for k, v in out.items()})


def golden_hist(refct):
    qs = ref_loader.module("transforms3d_supplement")
    gm = ref_loader.module("general_maths")
    q = np.array([0.83, -0.31, 0.22, 0.41])
    q = q / np.linalg.norm(q)
    v = synth.nh_vectors(6000, 7, seed=synth.BASE_SEED + 21)
    # add exact on-axis / on-edge samples: +z, -z, +x, -x, +y, and a zero vector (NaN -> dropped)
    special = np.zeros((6, 7, 3), dtype=np.float32)
    special[0, :, 2] = 1; special[1, :, 2] = -1; special[2, :, 0] = 1; special[3, :, 0] = -1; special[4, :, 1] = 1

    def run(vecs, quat, nbx):
        with np.errstate(all="ignore"):
            w = qs.rotate_vector_simd(vecs, quat) if quat is not None else vecs          # :567
            rtp = gm.xyz_to_rtp(w)                                                      # :588
            rtp = np.transpose(rtp, axes=(1, 0, 2))                                     # :600
            rtp = np.delete(rtp, 0, axis=2)                                             # :611
            rtp[..., 1] = np.cos(rtp[..., 1])                                           # :613
            hl = np.zeros((vecs.shape[1], nbx, nbx // 2), dtype=rtp.dtype)
            for i in range(vecs.shape[1]):                                              # :617-626 (no `normed`)
                h, e = np.histogramdd(rtp[i], bins=(nbx, nbx // 2), range=((-np.pi, np.pi), (-1, 1)))
                hl[i] = h
        return hl, e

    vr = _away_from_edges(v, q, 72, 1e-9)
    h_rot, e = run(vr, q, 72)
    v32 = _away_from_edges(v, None, 72, 2e-5)
    h_f32, _ = run(v32, None, 72)
    h_special, _ = run(special, None, 72)          # identity-frame on-edge semantics (float32 path)
    vr36 = _away_from_edges(v, q, 36, 1e-9)
    h_rot36, e36 = run(vr36, q, 36)
    save("hist.npz", q=q, vecs_rot=vr, hist_rot=h_rot.astype(np.int64), edges_phi=e[0], edges_cos=e[1],
         vecs_f32=v32, hist_f32=h_f32.astype(np.int64), special=special, hist_special=h_special.astype(np.int64),
         vecs_rot36=vr36, hist_rot36=h_rot36.astype(np.int64))
    # S2 and average vector of the same stream (next-tier rows, same pass)
    with quiet():
        s2 = refct.calculate_S2_by_outerProduct(vr.astype(np.float64)[:5000], 10.0, 10000.0)
        s2_all = refct.calculate_S2_by_outerProduct(vr.astype(np.float64))
    save("s2.npz", vecs=vr[:5000], s2_blocks=s2, s2_all=s2_all)


def golden_rtp():
    """The REAL gm.xyz_to_rtp / gm.rtp_to_xyz and qs.rotate_vector_simd on the --vecDist (no --vecHist) branch
    (calculate-Ct-from-traj.py:567,588,600-607): float32 without rotation, float64 after it."""
    qs = ref_loader.module("transforms3d_supplement")
    gm = ref_loader.module("general_maths")
    q = np.array([0.83, -0.31, 0.22, 0.41])
    q = q / np.linalg.norm(q)
    v = synth.nh_vectors(700, 5, seed=synth.BASE_SEED + 31)
    v[:3] *= np.float32(1.7)                                   # non-unit rows: r is a real output
    v[5, :, :] = 0
    v[5, 0, 2] = 1; v[5, 1, 2] = -1; v[5, 2, 0] = -1; v[5, 3, 1] = -1       # poles, branch cut, zero vector
    with np.errstate(all="ignore"):
        rtp32 = gm.xyz_to_rtp(v)
        rtp64 = gm.xyz_to_rtp(qs.rotate_vector_simd(v, q))
        unit64 = gm.xyz_to_rtp(v.astype(np.float64), bUnit=True)
        ax0 = gm.xyz_to_rtp(np.ascontiguousarray(np.moveaxis(v[:9].astype(np.float64), -1, 0)), vaxis=0)
        one = gm.xyz_to_rtp(v[7, 2].astype(np.float64))
    pt = np.stack(np.meshgrid(np.linspace(-3, 3, 7), np.linspace(0.1, 3.0, 5), indexing="ij"), axis=-1)
    back = gm.rtp_to_xyz(pt, vaxis=-1, bUnit=True)
    back1 = gm.rtp_to_xyz(np.array([1.3, 0.4, 2.1]))
    # gs.print_s3d stops with a TypeError after its first set (general_scripts.py:303-306): keep what it wrote
    import gc
    import tempfile
    gs = ref_loader.module("general_scripts")
    with tempfile.TemporaryDirectory() as td:
        try:
            gs.print_s3d(td + "/pt.dat", ["A", "B", "C", "D", "E"], np.transpose(rtp64, axes=(1, 0, 2)), (1, 2))
            crashed = False
        except TypeError:
            crashed = True
        gc.collect()
        s3d = open(td + "/pt.dat").read()
    save("rtp.npz", s3d_partial=np.array(s3d), s3d_crashed=np.array(crashed), q=q, vecs=v, rtp32=rtp32, rtp64=rtp64, unit64=unit64, ax0=ax0, one=one, pt=pt, back=back,
         back1=back1)


def golden_qs():
    """The REAL host-side quaternion helpers of transforms3d_supplement.py (with the transforms3d stand-in of
    oracle/_stubs) that SURVEY 8b lists as importable surface."""
    qs = ref_loader.module("transforms3d_supplement")
    rng = np.random.default_rng(synth.BASE_SEED + 41)
    q1 = rng.standard_normal((50, 4)); q1 /= np.linalg.norm(q1, axis=1, keepdims=True)
    q2 = rng.standard_normal((50, 4)); q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    q32 = q1.astype(np.float32)
    v = rng.standard_normal((50, 3))
    v[7] = 0.0                                                     # 0/0 -> 0 in vecnorm_NDarray
    frames = []
    for k in range(12):                                            # right-handed orthonormal frames (eigenvector-like)
        A = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        if np.linalg.det(A) < 0:
            A[:, 2] *= -1
        frames.append(A.T)                                         # rows are the axes
    frames = np.array(frames)
    with np.errstate(all="ignore"):
        out = dict(
            q1=q1, q2=q2, v=v, frames=frames,
            mult=qs.quat_mult_simd(q1, q2), mult_mixed=qs.quat_mult_simd(qs.quat_invert(q32[:-1]), q32[1:]),
            invert=qs.quat_invert(q1), invert32=qs.quat_invert(q32),
            reduce=qs.quat_reduce_simd(q1), reduce_ref=qs.quat_reduce_simd(q1, qref=q2[0]),
            reduce_ax0=qs.quat_reduce_simd(np.ascontiguousarray(q1.T), axis=0),
            vecnorm=qs.vecnorm_NDarray(v), vecnorm1=qs.vecnorm_NDarray(v[3]), vecnorm_ax0=qs.vecnorm_NDarray(v.T.copy(), axis=0),
            v1v2=np.array([qs.quat_v1v2(qs.vecnorm_NDarray(v[i]), qs.vecnorm_NDarray(v[i + 1])) for i in range(8, 20)]),
            frame_min=np.array([qs.quat_frame_transform_min(f) for f in frames]),
            rot_small=qs.rotate_vector_simd(v[:9], q1[0]), rot_perq=qs.rotate_vector_simd(v, q1),
            rot_unnorm=qs.rotate_vector_simd(v[:9], 2.5 * q1[1]),
        )
    save("qs.npz", **out)


def golden_io():
    """Text produced / parsed by the REAL general_scripts.py, plumedcolvario.py and dxio.py writers and readers."""
    import tempfile
    gs = ref_loader.module("general_scripts")
    pl = ref_loader.module("plumedcolvario")
    dx = ref_loader.module("dxio")
    rng = np.random.default_rng(synth.BASE_SEED + 51)
    x = (np.arange(7) + 1) * 10.0
    y1 = rng.standard_normal(7)
    y2 = rng.standard_normal((3, 7))
    ct = np.stack((rng.random((4, 7)).astype(np.float32), (rng.random((4, 7)) * 1e-3).astype(np.float32)), axis=-1)   # (nR, L, 2) float32
    hist = rng.integers(0, 50, (6, 4)).astype(np.float64)
    edges = [np.linspace(-np.pi, np.pi, 7), np.linspace(-1, 1, 5)]
    vol = rng.random((3, 4, 5))
    q = synth.quaternion_walk(12, seed=synth.BASE_SEED + 52)
    out = {}
    with tempfile.TemporaryDirectory() as td, quiet():
        f = td + "/f"
        gs.print_xylist(f, x, y1); out["xylist_1d"] = open(f).read()
        gs.print_xylist(f, x, y2); out["xylist_2d"] = open(f).read()
        gs.print_xylist(f, x, y2, True, header="# head"); out["xylist_cols"] = open(f).read()
        gs.print_sxylist(f, ["1", "2", "7", "9"], x, ct, header=["# a", "# b"]); out["sxylist"] = open(f).read()
        legs, lx, ly, ldy = gs.load_sxydylist(f, "legend")
        out["sxy_legs"], out["sxy_x"], out["sxy_y"], out["sxy_dy"] = np.array(legs), np.array(lx), np.array(ly), np.array(ldy)
        gs.print_gplot_hist(f, hist, edges, header="# h", bSphere=True); out["gplot_sphere"] = open(f).read()
        gs.print_gplot_hist(f, hist, edges, header="# h", bSphere=False); out["gplot_flat"] = open(f).read()
        dx.write_to_dx(f, vol, (3, 4, 5), [-1.0, -0.5, 0.25], np.diag([0.5, 0.25, 0.125]), "nm"); out["dx"] = open(f).read()
        with open(f, "w") as fp:
            fp.write("#! FIELDS time q.w q.x q.y q.z\n#! SET something 1\n")
            for i, r in enumerate(q):
                fp.write(" %f %16g %16g %16g %16g\n" % (i * 0.1234567, *[float(v) for v in r]))
        out["plumed_text"] = open(f).read()
        names, data = pl.read_from_plumedprint(f)
        out["plumed_names"], out["plumed_data"] = np.array(names), np.array(data)
    save("io.npz", x=x, y1=y1, y2=y2, ct=ct, hist=hist, edges_phi=edges[0], edges_cos=edges[1], vol=vol,
         **{k: np.array(v) for k, v in out.items()})


def golden_sd_host():
    """Host-only parts of the REAL spectral_densities.py: gyromagnetic data, angular frequencies and prefactors,
    axisymmetric D / A coefficients, histogram -> bin vectors."""
    sd = ref_loader.module("spectral_densities")
    rng = np.random.default_rng(synth.BASE_SEED + 61)
    out = {}
    with quiet():
        cases = [("15N", "1H", 600.13, "MHz", "ps"), ("13C", "1H", 18.8, "T", "ns"), ("15N", "1H", 8.5e8, "Hz", "s")]
        for k, (a, b, f, fu, tu) in enumerate(cases):
            w = sd.angularFrequencies(a, b, f, fu, tu)
            out["w%d_omega" % k] = np.array(w.omega)
            out["w%d_scalars" % k] = np.array([w.get_factor_DD(), w.get_factor_CSA(), w.get_magnetic_field("T"),
                                               w.get_magnetic_field("MHz"), w.gA.gamma, w.gB.gamma, w.gA.csa])
            w.set_time_unit("ns")
            out["w%d_omega_ns" % k] = np.array(w.omega)
        w = sd.angularFrequencies("15N", "1H", 600, "MHz", "ps")
        w.initialise_CSA_array(4, [-160e-6, -170e-6, -175e-6, -180e-6])
        out["multi_csa_factors"] = np.array(w.get_factor_CSA())
        out["multi_csa_one"] = np.array(w.get_factor_CSA(2))
        vecs = rng.standard_normal((5, 7, 3))
        vecs /= np.linalg.norm(vecs, axis=-1, keepdims=True)
        for tag, D, conv in (("prolate", [2.5e-5, 1.6], False), ("oblate", [2.5e-5, 0.7], False), ("conv", [3.1e-5, 2.2e-5], True)):
            r = sd.globalRotationalDiffusion_Axisymmetric(D=D, bConvert=conv)
            r.vecXH = vecs
            r.update_A_coefficients()
            out[tag + "_D"] = np.array(r.D)
            out[tag + "_DJ"] = np.array(r.get_D_coefficients())
            out[tag + "_AJ"] = np.array(r.get_A_coefficients())
            out[tag + "_AJ_ind"] = np.array(r.get_A_coefficients(3))
            out[tag + "_Dpp"] = np.array(r.transform_D())
        r = sd.globalRotationalDiffusion_Axisymmetric(tau=8.0e3, aniso=1.3)
        r.set_Diso(2.2e-5); r.set_Daniso(0.9)
        out["setter_DJ"] = np.array(r.get_D_coefficients())
        hist = rng.integers(0, 9, (3, 8, 4)).astype(float)
        edges = np.empty(2, dtype=object)
        edges[0], edges[1] = np.linspace(-np.pi, np.pi, 9), np.linspace(-1, 1, 5)
        bv, bw = sd.convert_LambertCylindricalHist_to_vecs(hist, edges)
        out["bin_vecs"], out["bin_weights"] = np.array(bv), np.array(bw)
    save("sd_host.npz", vecs=vecs, hist=hist, edges_phi=edges[0], edges_cos=edges[1], **out)


def golden_dq(refdq):
    q = synth.quaternion_walk(6000, seed=synth.BASE_SEED + 31, sigma=(0.01, 0.015, 0.03))     # float32 (G2)
    lags = [5, 10, 40, 100, 333, 1000, 2999]
    nch = 4
    rec = dict(vxx=[], iso=[], moi=[], chunk_iso=[], chunk_moi=[], moi_rot=[], chunk_moi_rot=[])
    qf = np.array([0.9, 0.1, -0.3, 0.2]); qf = qf / np.linalg.norm(qf)
    for d in lags:
        dq = refdq.obtain_self_dq(q, d)
        v = dq[..., 1:4]
        n = len(v)
        rec["iso"].append(refdq.average_LegendreP1quat(n, v))
        rec["moi"].append(refdq.average_anisotropic_tensor(n, v))
        rec["moi_rot"].append(refdq.average_anisotropic_tensor(n, v, qf))
        rec["chunk_iso"].append(refdq.average_LegendreP1quat_chunk(n, v, nch))
        rec["chunk_moi"].append(refdq.average_anisotropic_tensor_chunk(n, v, nch))
        rec["chunk_moi_rot"].append(refdq.average_anisotropic_tensor_chunk(n, v, nch, qf))
    dq5 = refdq.obtain_self_dq(q, 5)
    save("dq_moments.npz", q=q, lags=np.array(lags), nchunk=nch, qframe=qf, dq_lag5=dq5,
         **{k: np.array(val) for k, val in rec.items() if val})
    # end-to-end CLI on a PLUMED text file (run-all.bash:379-387 flags), outputs stored as text
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        qq = synth.quaternion_walk(20000, seed=synth.BASE_SEED + 32, sigma=(0.004, 0.006, 0.012), dtype=np.float64)
        fn = os.path.join(td, "colvar-q")
        synth.write_plumed_quaternions(fn, qq, dt=10.0)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref_loader.STUBS, ref_loader.REF_BUILD, ref_loader.REFERENCE]))
        cmd = [sys.executable, os.path.join(ref_loader.REFERENCE, "calculate-dq-distribution.py"), "--iso", "--aniso",
               "-f", fn, "-o", os.path.join(td, "rotdif"), "--mindt", "500", "--skip", "500", "--maxdt", "50000",
               "--num_chunk", "4"]
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=td)
        files = {}
        for suf in ("-iso.dat", "-aniso2.dat", "-aniso_q.dat", "-moi.xyz"):
            with open(os.path.join(td, "rotdif" + suf)) as fp:
                files[suf] = fp.read()
        with open(fn) as fp:
            plumed = fp.read()
    save("dq_cli.npz", plumed=np.array(plumed), iso=np.array(files["-iso.dat"]), aniso2=np.array(files["-aniso2.dat"]),
         aniso_q=np.array(files["-aniso_q.dat"]), moi_xyz=np.array(files["-moi.xyz"]))


def _synthetic_curves(n_res, n_pts, seed):
    rng = np.random.default_rng(seed)
    t = (np.arange(n_pts) + 1.0) * 10.0
    Ct, dCt, truth = [], [], []
    for i in range(n_res):
        S2 = rng.uniform(0.45, 0.9)
        nc = 1 + i % 3
        C = rng.dirichlet(np.ones(nc)) * (1 - S2) * rng.uniform(0.85, 1.0)
        tau = np.sort(10 ** rng.uniform(1.2, 3.2, nc))
        y = S2 + np.sum(C[:, None] * np.exp(-t[None] / tau[:, None]), axis=0)
        sig = 0.002 + 0.004 * t / t[-1]
        y = y + rng.standard_normal(n_pts) * sig * 0.5
        Ct.append(y); dCt.append(sig); truth.append((S2, C, tau))
    return t, np.array(Ct), np.array(dCt), truth


def golden_dq_multi(refdq):
    """Replica pooling of calculate-dq-distribution-multi.py:529-540 (that script imports a missing `xvgio` and cannot
    run, so its loop body is replayed here with the REAL helper functions it shares with calculate-dq-distribution.py:
    per replica obtain_self_dq, concatenate, then the pooled averages), and the --hist 3-D histogram (:633-634) with
    `density=True` in place of the removed `normed=True`."""
    reps = np.stack([synth.quaternion_walk(3000, seed=synth.BASE_SEED + 60 + r, sigma=(0.01, 0.015, 0.03)) for r in range(3)])
    lags = [1, 7, 50, 333, 1499]
    nch = 4
    rec = dict(iso=[], moi=[], chunk_iso=[], chunk_moi=[])
    for d in lags:
        v = np.concatenate([refdq.obtain_self_dq(reps[r], d)[..., 1:4] for r in range(3)], axis=0)
        n = len(v)
        rec["iso"].append(refdq.average_LegendreP1quat(n, v))
        rec["moi"].append(refdq.average_anisotropic_tensor(n, v))
        rec["chunk_iso"].append(refdq.average_LegendreP1quat_chunk(n, v, nch))
        rec["chunk_moi"].append(refdq.average_anisotropic_tensor_chunk(n, v, nch))
    hists = {}
    for d, nb in ((50, 21), (333, 101)):
        v = refdq.obtain_self_dq(reps[0], d)[..., 1:4]
        h, e = np.histogramdd(v, range=[(-1, 1)] * 3, bins=(nb, nb, nb), density=True)
        nz = np.nonzero(h)
        hists["hist_%d_idx" % d] = np.stack(nz, axis=1).astype(np.int32)
        hists["hist_%d_val" % d] = h[nz]
        hists["hist_%d_nb" % d] = np.array(nb)
    save("dq_multi.npz", q=reps, lags=np.array(lags), nchunk=nch, **{k: np.array(v) for k, v in rec.items()}, **hists)


def golden_fit():
    fitCt = ref_loader.module("fitting_Ct_functions")
    t, Ct, dCt, _ = _synthetic_curves(9, 250, synth.BASE_SEED + 41)
    rows = []
    single = []
    for i in range(len(Ct)):
        m = fitCt.autoCorrelationModel(name=i)
        with quiet():
            chi = m.optimised_curve_fitting(t, Ct[i], dCt[i], listDoG=[2, 3, 5, 7, 9], chiSqThreshold=0.5,
                                            fp=io.StringIO())
        p = np.full(12, np.nan)
        p[0] = m.nParams; p[1] = chi; p[2] = m.S2
        p[3:3 + m.nComps] = m.C; p[7:7 + m.nComps] = m.tau
        rows.append(p)
        for npar in (2, 3, 5):
            m2 = fitCt.autoCorrelationModel(name=i)
            m2.set_nParams(npar)
            with quiet():
                chi2, qual = m2.conduct_curve_fitting(t, Ct[i], dCt[i], bReInitialise=True, fp=io.StringIO())
            r = np.full(16, np.nan)
            r[0] = npar; r[1] = chi2; r[2:5] = qual; r[5] = m2.S2
            if np.isfinite(chi2):
                r[6:6 + m2.nComps] = m2.C; r[9:9 + m2.nComps] = m2.tau
                r[12:12 + m2.nComps] = m2.dC
            single.append(r)
    save("fit.npz", t=t, Ct=Ct, dCt=dCt, ladder=np.array(rows), single=np.array(single))


def golden_relax():
    sd = ref_loader.module("spectral_densities")
    fitCt = ref_loader.module("fitting_Ct_functions")
    npufunc = ref_loader.module("npufunc")
    rng = np.random.default_rng(synth.BASE_SEED + 51)
    nR = 6
    # histogram weights: counts of a synthetic rotated stream
    q = np.array([0.83, -0.31, 0.22, 0.41]); q = q / np.linalg.norm(q)
    v = synth.nh_vectors(3000, nR, seed=synth.BASE_SEED + 52)
    from oracle import ct_oracle
    hist, edges = ct_oracle.sphere_histogram(v, q)
    import tempfile
    models = fitCt.autoCorrelations()
    pars = []
    for i in range(nR):
        nc = 1 + i % 3
        S2 = rng.uniform(0.5, 0.9)
        C = rng.dirichlet(np.ones(nc)) * (1 - S2)
        tau = np.sort(10 ** rng.uniform(1.0, 3.3, nc))
        models.add_model(str(i), name=i, listC=list(C), listTau=list(tau), S2=S2, bS2Fast=False)
        row = np.full(8, np.nan); row[0] = nc; row[1] = S2; row[2:2 + nc] = C; row[5:5 + nc] = tau
        pars.append(row)
    zeta = (1.02 / 1.04) ** 6
    models.set_zeta(zeta)
    res = {}
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "h_vecHistogram.npz")
        e_obj = np.empty(2, dtype=object); e_obj[0] = edges[0]; e_obj[1] = edges[1]
        np.savez_compressed(fn, names=np.arange(nR), dataType="LambertCylindrical", bHistogram=True, edges=e_obj,
                            axisLabels=["phi", "cos(theta)"], data=hist)
        for tag, Dani in (("prolate", 1.35), ("oblate", 0.8)):
            with quiet():
                rot = sd.globalRotationalDiffusion_Axisymmetric(D=[2.1e-5, Dani], bConvert=False)
                rot.import_frame_vectors(fn)
            for field in (600.133, 800.0):
                with quiet():
                    w = sd.angularFrequencies("15N", "1H", field, "MHz", "ps")
                    for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
                        ex = cls(name, "ps", w, rot, models)
                        ex.eval()
                        res["%s_%s_%d" % (tag, name, round(field))] = np.stack((ex.values, ex.errors))
        with quiet():
            iso = sd.globalRotationalDiffusion_Isotropic(D=2.1e-5)
            w = sd.angularFrequencies("15N", "1H", 600.133, "MHz", "ps")
            for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
                ex = cls(name, "ps", w, iso, models)
                ex.eval()
                res["iso_%s_600" % name] = np.array(ex.values)
            # per-residue CSA array path (the rsCSA grid workload): eval(ind=i) with a CSA array
            w2 = sd.angularFrequencies("15N", "1H", 600.133, "MHz", "ps")
            csa = -170e-6 + 10e-6 * np.arange(nR)
            w2.initialise_CSA_array(nR, csa)
            rotp = sd.globalRotationalDiffusion_Axisymmetric(D=[2.1e-5, 1.35], bConvert=False)
            rotp.import_frame_vectors(fn)
            for name, cls in (("R1", sd.spinRelaxationR1), ("R2", sd.spinRelaxationR2), ("NOE", sd.spinRelaxationNOE)):
                ex = cls(name, "ps", w2, rotp, models)
                for i in range(nR):
                    ex.eval(ind=i)
                res["csa_%s_600" % name] = np.stack((ex.values, ex.errors))
            om = w.omega.copy(); fdd = w.get_factor_DD(); fcsa = w.get_factor_CSA()
    x = rng.uniform(1e-6, 1e-2, 64); y = rng.uniform(0, 1e-2, 64)
    save("relax.npz", hist=hist.astype(np.int64), edges_phi=edges[0], edges_cos=edges[1], params=np.array(pars),
         zeta=zeta, Diso=2.1e-5, omega_600=om, f_dd=fdd, f_csa_600=fcsa, csa_array=csa,
         jomega_x=x, jomega_y=y, jomega_out=npufunc.Jomega(x, y), jomega_outer=npufunc.Jomega.outer(x[:3], y[:5]),
         jomega_f32=npufunc.Jomega(x.astype(np.float32), y.astype(np.float32)), **res)


def golden_relax_cli():
    """calculate-relaxations-multi-field.py (prediction mode) run unmodified on a fittedCt file + histogram."""
    import subprocess
    import tempfile
    fitCt = ref_loader.module("fitting_Ct_functions")
    g = np.load(os.path.join(OUT, "relax.npz"))
    with tempfile.TemporaryDirectory() as td:
        e = np.empty(2, dtype=object); e[0] = g["edges_phi"]; e[1] = g["edges_cos"]
        np.savez_compressed(td + "/h_vecHistogram.npz", names=np.arange(6), dataType="LambertCylindrical", bHistogram=True,
                            edges=e, axisLabels=["phi", "cos(theta)"], data=g["hist"].astype(float))
        ac = fitCt.autoCorrelations()
        for i, row in enumerate(g["params"]):
            nc = int(row[0])
            m = ac.add_model(str(i), name=i, listC=list(row[2:2 + nc]), listTau=list(row[5:5 + nc]), S2=row[1])
            m.bHasFit = True; m.chiSq = 0.1; m.dC = np.zeros(nc); m.dtau = np.zeros(nc); m.dS2 = 0.0
        ac.import_target_array([str(i) for i in range(6)], [np.arange(1, 11.) * 10] * 6, [np.ones(10)] * 6)
        ac.export(td + "/x_fittedCt.dat")
        expt = {}
        for t in ("R1", "R2", "NOE"):
            expt[t] = "# Type %s\n# NucleiA 15N\n# NucleiB 1H\n# Frequency 600.133\n" % t + \
                "".join("%d 1.0 0.1\n" % i for i in range(6))
            with open(td + "/e_%s.dat" % t, "w") as fp:
                fp.write(expt[t])
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref_loader.STUBS, ref_loader.REF_BUILD, ref_loader.REFERENCE]))
        cmd = [sys.executable, ref_loader.REFERENCE + "/calculate-relaxations-multi-field.py", "-f", td + "/x_fittedCt.dat",
               "--distfn", td + "/h_vecHistogram.npz", "-D", "2.1e-5", "--aniso", "1.35", "-o", td + "/ref",
               td + "/e_R1.dat", td + "/e_R2.dat", td + "/e_NOE.dat"]
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out = {t: open(td + "/ref_15N1H_600MHz_%s.xvg" % t).read() for t in ("R1", "R2", "NOE")}
        fitted = open(td + "/x_fittedCt.dat").read()
    save("relax_cli.npz", fitted=np.array(fitted), **{"xvg_" + k: np.array(v) for k, v in out.items()},
         **{"expt_" + k: np.array(v) for k, v in expt.items()})


def golden_fit_cli():
    """calculate-fitted-Ct.py run unmodified on a `_Ctint.dat` file: default ladder, --nc 2, --nofast."""
    import subprocess
    import tempfile
    gs = ref_loader.module("general_scripts")
    g = np.load(os.path.join(OUT, "fit.npz"))
    t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        with quiet():
            gs.print_sxylist(td + "/c_Ctint.dat", list(range(1, len(Ct) + 1)), t,
                             np.stack((Ct.astype(np.float32), dCt.astype(np.float32)), axis=-1))
        out["ctint"] = open(td + "/c_Ctint.dat").read()
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref_loader.STUBS, ref_loader.REF_BUILD, ref_loader.REFERENCE]))
        for tag, extra in (("ladder", []), ("nc2", ["--nc", "2"]), ("nofast", ["--nofast"])):
            cmd = [sys.executable, ref_loader.REFERENCE + "/calculate-fitted-Ct.py", "-f", td + "/c_Ctint.dat", "-o", td + "/" + tag] + extra
            subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[tag] = open(td + "/%s_fittedCt.dat" % tag).read()
    save("fit_cli.npz", **{k: np.array(v) for k, v in out.items()})


def golden_relax_opt():
    """calculate-relaxations-multi-field.py --opt ... (Powell optimisation against experiment, spectral_densities.py
    :1302-1447) run unmodified.  The "experiments" are the reference's own predictions at perturbed parameters
    (Diso x 1.08, residue-specific CSA), two fields, so the optimum is known to exist and is well conditioned."""
    import subprocess
    import tempfile
    g = np.load(os.path.join(OUT, "relax.npz"))
    gc = np.load(os.path.join(OUT, "relax_cli.npz"))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref_loader.STUBS, ref_loader.REF_BUILD, ref_loader.REFERENCE]))
    script = ref_loader.REFERENCE + "/calculate-relaxations-multi-field.py"
    rng = np.random.default_rng(synth.BASE_SEED + 55)
    csa_true = -170e-6 * (1.0 + 0.08 * rng.standard_normal(6))
    res = {}
    with tempfile.TemporaryDirectory() as td:
        e = np.empty(2, dtype=object); e[0] = g["edges_phi"]; e[1] = g["edges_cos"]
        np.savez_compressed(td + "/h_vecHistogram.npz", names=np.arange(6), dataType="LambertCylindrical", bHistogram=True,
                            edges=e, axisLabels=["phi", "cos(theta)"], data=g["hist"].astype(float))
        with open(td + "/x_fittedCt.dat", "w") as fp:
            fp.write(str(gc["fitted"]))
        with open(td + "/csa_true.dat", "w") as fp:
            for i, c in enumerate(csa_true):
                fp.write("%d %.9g\n" % (i, c))
        # 1. "experimental" data = prediction at the true parameters, residues 0..5 (one peak missing in one file)
        dummy = []
        for f in (600.133, 800.25):
            for t in ("R1", "R2", "NOE"):
                fn = td + "/d_%s_%d.dat" % (t, f)
                with open(fn, "w") as fp:
                    fp.write("# Type %s\n# NucleiA 15N\n# NucleiB 1H\n# Frequency %g\n" % (t, f))
                    fp.write("".join("%d 1.0 0.1\n" % i for i in range(6)))
                dummy.append(fn)
        cmd = [sys.executable, script, "-f", td + "/x_fittedCt.dat", "--distfn", td + "/h_vecHistogram.npz", "-D", "2.268e-5",
               "--aniso", "1.35", "--csa", td + "/csa_true.dat", "-o", td + "/true"] + dummy
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        expt = {}
        files = []
        for f in (600, 800):
            for t in ("R1", "R2", "NOE"):
                rows = [l.split() for l in open(td + "/true_15N1H_%dMHz_%s.xvg" % (f, t)) if l[0] not in "#@&\n"]
                txt = "# Type %s\n# NucleiA 15N\n# NucleiB 1H\n# Frequency %s\n" % (t, "600.133" if f == 600 else "800.25")
                for r in rows:
                    if t == "NOE" and f == 800 and r[0] == "4":
                        continue                          # unresolved peak: exercises the name maps
                    y = float(r[1])
                    txt += "%s %.6g %.3g\n" % (r[0], y, 0.02 * abs(y))
                expt["%s_%d" % (t, f)] = txt
                fn = td + "/e_%s_%d.dat" % (t, f)
                with open(fn, "w") as fp:
                    fp.write(txt)
                files.append(fn)
        # 2. the optimisation modes
        for tag, opt, extra in (("Diso", "Diso", []), ("rsCSA", "rsCSA", []), ("mixed", "Diso,rsCSA", ["--cycles", "4"])):
            cmd = [sys.executable, script, "-f", td + "/x_fittedCt.dat", "--distfn", td + "/h_vecHistogram.npz", "-D", "2.1e-5",
                   "--aniso", "1.35", "-o", td + "/" + tag, "--opt", opt] + extra + files
            out = subprocess.run(cmd, env=env, check=True, capture_output=True, text=True).stdout
            chi = [l for l in out.splitlines() if "Final chi-value" in l][-1]
            res["chi_" + tag] = np.array(float(chi.split(":")[-1]))
            for f in (600, 800):
                for t in ("R1", "R2", "NOE"):
                    res["xvg_%s_%s_%d" % (tag, t, f)] = np.array(open(td + "/%s_15N1H_%dMHz_%s.xvg" % (tag, f, t)).read())
            if "rsCSA" in opt:
                res["csa_" + tag] = np.loadtxt(td + "/%s_CSA_opt.dat" % tag)
            print(tag, chi)
    save("relax_opt.npz", csa_true=csa_true, **{"expt_" + k: np.array(v) for k, v in expt.items()}, **res)


def golden_dq_xvg():
    """`gmx rotmat` input of calculate-dq-distribution.py (:389-408, 487-492): the real load_xys +
    rotmatrix_to_quaternion(bInvert=True) on an .xvg text of rotation matrices, and the real CLI run on that file."""
    import subprocess
    import tempfile
    refdq = ref_loader.script("calculate-dq-distribution.py")
    gs = ref_loader.module("general_scripts")
    qq = synth.quaternion_walk(4000, seed=synth.BASE_SEED + 41, sigma=(0.004, 0.006, 0.012), dtype=np.float64)
    w, x, y, z = qq.T
    # rotation matrix of the *inverse* rotation is what gmx rotmat reports relative to the reference; any proper
    # rotation matrices will do for the parser and the converter
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1)
    lines = ["# This file was created by a test generator in the format of gmx rotmat", "@    title \"Fit matrix\"",
             "@    xaxis  label \"Time (ps)\"", "@TYPE xy", "@ s0 legend \"xx\""]
    for i in range(len(R)):
        lines.append("%12.3f" % (i * 10.0) + "".join(" %10.6f" % v for v in R[i]))
    lines.append("&")
    text = "\n".join(lines) + "\n"
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "rotmat.xvg")
        with open(fn, "w") as fp:
            fp.write(text)
        t, m = gs.load_xys(fn)
        with quiet():
            data = refdq.rotmatrix_to_quaternion(t, m, bInvert=True)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref_loader.STUBS, ref_loader.REF_BUILD, ref_loader.REFERENCE]))
        cmd = [sys.executable, os.path.join(ref_loader.REFERENCE, "calculate-dq-distribution.py"), "--iso", "--aniso",
               "-f", fn, "-o", os.path.join(td, "rotdif"), "--mindt", "200", "--skip", "200", "--maxdt", "10000",
               "--num_chunk", "4"]
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=td)
        files = {}
        for suf in ("-iso.dat", "-aniso2.dat", "-aniso_q.dat"):
            with open(os.path.join(td, "rotdif" + suf)) as fp:
                files[suf] = fp.read()
    save("dq_xvg.npz", xvg=np.array(text), data=data, iso=np.array(files["-iso.dat"]), aniso2=np.array(files["-aniso2.dat"]),
         aniso_q=np.array(files["-aniso_q.dat"]))


def main():
    if not ref_loader.available():
        sys.exit("reference tree not found at %s" % ref_loader.REFERENCE)
    if len(sys.argv) > 1:                       # regenerate selected files only: make_golden.py golden_rtp ...
        for name in sys.argv[1:]:
            globals()[name]()
        return
    refct = ref_loader.script("calculate-Ct-from-traj.py")
    refdq = ref_loader.script("calculate-dq-distribution.py")
    golden_ct(refct)
    golden_traj(refct)
    golden_ct_cli(refct)
    golden_hist(refct)
    golden_rtp()
    golden_qs()
    golden_io()
    golden_sd_host()
    golden_dq(refdq)
    golden_dq_multi(refdq)
    golden_dq_xvg()
    golden_fit()
    golden_relax()
    golden_relax_cli()
    golden_fit_cli()
    golden_relax_opt()


if __name__ == "__main__":
    main()
