"""GPU parity of the trajectory front end: X-H vector extraction and superposition."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ct_oracle, traj_oracle

pytestmark = pytest.mark.gpu


def test_xh_vectors_bit_exact(golden):
    import torch
    from spinrelax_b200 import traj
    g = golden("traj.npz")
    v = traj.xh_vectors_device(torch.from_numpy(g["xyz"]).cuda(), g["indexH"], g["indexX"]).cpu().numpy()
    assert np.array_equal(v, g["vecXH"])                 # incl. the zero vector -> (0, 0, 0)
    t = traj.ArrayTrajectory(g["xyz"], {"name H": g["indexH"], "name N and not resname PRO": g["indexX"]})
    v2 = traj.obtain_XHvecs(t, "name H", "name N and not resname PRO", bSuppressPrint=True)
    assert np.array_equal(v2, g["vecXH"])
    with pytest.raises(SystemExit):                      # empty selection exits like the reference (:71-74)
        traj.obtain_XHvecs(t, "name HN", "name N and not resname PRO", bSuppressPrint=True)


def test_superposed_vectors_match_oracle(golden):
    import torch
    from spinrelax_b200 import traj
    g = golden("traj.npz")
    xyz = torch.from_numpy(g["xyz"]).cuda()
    v, R = traj.xh_vectors_superposed_device(xyz, g["ref"], g["fit"], g["indexH"], g["indexX"], return_rotations=True)
    ov, oR = traj_oracle.xh_vectors_superposed(g["xyz"], g["ref"], g["fit"], g["indexH"], g["indexX"])
    assert np.max(np.abs(R.cpu().numpy() - oR)) < 1e-9
    assert np.max(np.abs(v.cpu().numpy() - ov)) < 2e-7     # one float32 ulp of a unit-vector component
    assert np.array_equal(v.cpu().numpy()[3, 2], np.zeros(3, dtype=np.float32))


@pytest.mark.parametrize("shape", [(1, 1, 3), (33, 7, 3), (257, 40, 100)])
def test_superposition_shapes_and_ct(shape):
    """Ragged sizes (fewer bonds / fit atoms than a warp, one frame); the fitted vectors feed C(t) unchanged."""
    import torch
    from spinrelax_b200 import ct, synth, traj
    nF, nRes, _ = shape
    xyz, sel, ref = synth.backbone_trajectory(nF, nRes, seed=900 + nF)
    fit = sel["custom occupancy"][:max(3, shape[2])] if nRes * 3 >= 3 else sel["custom occupancy"]
    ih, ix = sel["name H"], sel["name N and not resname PRO"]
    v = traj.xh_vectors_superposed_device(torch.from_numpy(xyz).cuda(), ref, fit, ih, ix)
    ov, _ = traj_oracle.xh_vectors_superposed(xyz, ref, fit, ih, ix)
    assert v.shape == (nF, nRes, 3)
    assert np.max(np.abs(v.cpu().numpy() - ov)) < 3e-7
    if nF >= 200:
        v4 = v[:256].reshape(2, 128, nRes, 3)
        Ct, dCt = ct.ct_palmer_device(v4.contiguous())
        oCt, _ = ct_oracle.ct_palmer(ov[:256].reshape(2, 128, nRes, 3).astype(np.float64))
        assert rel_err(Ct.cpu().numpy(), oCt) < 5e-6       # inputs differ by a float32 ulp


def test_superposition_removes_tumbling():
    """Property at scale: after the fit, C(t) of the tumbling trajectory equals C(t) of the internal motion."""
    import torch
    from spinrelax_b200 import synth, traj
    nF, nRes = 20000, 16
    xyz, sel, ref = synth.backbone_trajectory(nF, nRes, seed=4242, tumbling_sigma=0.05, noise=0.0)
    ih, ix = sel["name H"], sel["name N and not resname PRO"]
    fit = np.concatenate((sel["name N and not resname PRO"], sel["name CA"]))      # rigid atoms only
    xd = torch.from_numpy(xyz).cuda()
    raw = traj.xh_vectors_device(xd, ih, ix).cpu().numpy().astype(np.float64)
    fitv = traj.xh_vectors_superposed_device(xd, ref, fit, ih, ix).cpu().numpy().astype(np.float64)
    internal = synth.nh_vectors(nF, nRes, seed=4242 + 1).astype(np.float64)
    # the fitted vectors are the internal ones up to ONE global rotation (reference frame of `ref`)
    gram_fit = np.einsum("fra,fsa->frs", fitv[:50], fitv[:50])
    gram_int = np.einsum("fra,fsa->frs", internal[:50], internal[:50])
    assert np.max(np.abs(gram_fit - gram_int)) < 5e-4      # float32 coordinates, 0.102 nm bonds
    assert np.max(np.abs(np.einsum("fra,fsa->frs", raw[:50], raw[:50]) - gram_int)) < 5e-4
    drift_raw = np.abs((raw[0] * raw[-1]).sum(-1) - (internal[0] * internal[-1]).sum(-1)).max()
    drift_fit = np.abs((fitv[0] * fitv[-1]).sum(-1) - (internal[0] * internal[-1]).sum(-1)).max()
    assert drift_fit < 5e-4 < drift_raw


def test_cli_ct_from_coordinates(tmp_path):
    """`calculate-Ct-from-traj`-style run from Cartesian coordinates: Ctext from the raw vectors, Ctint from the
    superposed ones, both against the oracle chain (traj_oracle -> ct_oracle)."""
    import contextlib, io
    from spinrelax_b200 import cli_ct, io_formats, synth
    nF, nRes = 1200, 6
    xyz, sel, ref = synth.backbone_trajectory(nF, nRes, seed=31)
    ih, ix, fit = sel["name H"], sel["name N and not resname PRO"], sel["custom occupancy"]
    np.savez(tmp_path / "trj.npz", xyz=xyz, indexH=ih, indexX=ix, fit=fit, names=np.arange(2, 2 + nRes), dt=10.0)
    np.save(tmp_path / "ref.npy", ref)
    pref = str(tmp_path / "out")
    with contextlib.redirect_stdout(io.StringIO()):
        cli_ct.main(["-s", str(tmp_path / "ref.npy"), "-f", str(tmp_path / "trj.npz"), "--tau", "3000", "-o", pref, "--Ct"])
    ofit, _ = traj_oracle.xh_vectors_superposed(xyz, ref, fit, ih, ix)
    oraw = traj_oracle.xh_vectors(xyz, ih, ix)
    for suffix, vecs in (("_Ctint.dat", ofit), ("_Ctext.dat", oraw)):
        legs, x, y, dy = io_formats.load_sxydylist(pref + suffix)
        v4 = ct_oracle.reformat_by_tau([vecs], 10.0, 3000.0)
        oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
        assert legs == [str(i) for i in range(2, 2 + nRes)]
        assert rel_err(y, oCt.T) < 5e-6 and np.max(np.abs(dy - odCt.T)) < 2e-6
