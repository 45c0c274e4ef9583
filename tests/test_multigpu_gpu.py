"""Multi-GPU product paths (BASELINE configs 4 / 5).  Needs >= 2 CUDA devices (`gpurun --gpus 2`); on a
one-GPU box these tests skip.

 * single process, several devices (spinrelax_b200.multigpu): the public functions shard bond vectors / residues / lags
   over the selected GPUs -- C(t), histograms, fits and relaxation must be bit-identical to the one-GPU run (uneven shares included), FP64
   moment sums equal to rounding;
 * one process per GPU (torch.distributed, NCCL): pipeline.CtHistStep at world 2 gathers C(t), dC(t) and the histogram
   of the whole vector set on rank 0 -- compared with the oracle."""
import io
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu


def _need2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")


@pytest.fixture
def two_gpus():
    _need2()
    from spinrelax_b200 import multigpu
    yield multigpu
    multigpu.set_devices(None)


def test_ct_hist_s2_sharded_by_vector_equal_single_gpu(two_gpus):
    from oracle import ct_oracle
    from spinrelax_b200 import ct, hist, synth
    nC, nF, nR = 3, 9000, 7                       # 7 vectors over 2 GPUs: blocks of 4 and 3
    v4 = synth.nh_vectors(nC * nF, nR, seed=404).reshape(nC, nF, nR, 3)
    q = np.array([0.83, -0.31, 0.22, 0.41])
    two_gpus.set_devices([0])
    one = (ct.calculate_Ct_Palmer_quiet(v4), hist.sphere_histogram(v4.reshape(-1, nR, 3), q),
           ct.calculate_S2_by_outerProduct(v4.reshape(-1, nR, 3), 10.0, 10.0 * nF), ct.average_vectors(v4.reshape(-1, nR, 3), q))
    two_gpus.set_devices([0, 1])
    assert [b[1:] for b in two_gpus.plan(nR)] == [(0, 4), (4, 7)]
    two = (ct.calculate_Ct_Palmer_quiet(v4), hist.sphere_histogram(v4.reshape(-1, nR, 3), q),
           ct.calculate_S2_by_outerProduct(v4.reshape(-1, nR, 3), 10.0, 10.0 * nF), ct.average_vectors(v4.reshape(-1, nR, 3), q))
    assert np.array_equal(one[0][0], two[0][0]) and np.array_equal(one[0][1], two[0][1])
    assert np.array_equal(one[1][0], two[1][0])
    # the moment kernel tiles (frames x vectors) per launch, so a vector's FP64 sums are added in another order when the
    # launch holds 4 instead of 7 vectors: equal to rounding, not bit for bit
    assert np.allclose(one[2], two[2], rtol=1e-12, atol=0) and np.allclose(one[3], two[3], rtol=1e-12, atol=1e-15)
    oCt, _ = ct_oracle.ct_palmer(v4.astype(np.float64))
    assert rel_err(two[0][0], oCt) < 1e-6
    ho, _ = ct_oracle.sphere_histogram(v4.reshape(-1, nR, 3), q)
    assert np.array_equal(two[1][0].astype(np.int64), ho.astype(np.int64))


def test_fits_relaxation_dq_sharded_equal_single_gpu(two_gpus):
    sys.path.insert(0, ROOT)
    from bench_secondary import synth_curves
    from spinrelax_b200 import dq, fitct, specdens as sd, synth
    n = 300
    t, Y, SG = synth_curves(n, 500, 77)
    res = []
    for devs in ([0], [0, 1]):
        two_gpus.set_devices(devs)
        ac = fitct.autoCorrelations()
        ac.import_target_array([str(i) for i in range(n)], [t] * n, Y, SG)
        chis = ac.fit_all_residues(fp=io.StringIO())
        iso = sd.globalRotationalDiffusion_Isotropic(D=2.1e-5)
        grid = sd.relax_grid(iso, ac, [600.133, 800.0], np.linspace(-200e-6, -140e-6, 8))
        q = synth.quaternion_walk(60000, seed=11, sigma=(0.004, 0.006, 0.012))
        M, cnt, counts = dq.dq_moment_sums(q, np.arange(10, 2010, 10), 4)
        res.append((chis, [(m.nParams, m.S2, tuple(m.C), tuple(m.tau)) for m in ac.model.values()], grid, M))
    (c1, m1, g1, M1), (c2, m2, g2, M2) = res
    assert np.array_equal(c1, c2) and m1 == m2
    for k in ("R1", "R2", "NOE"):
        assert np.array_equal(g1[k][0], g2[k][0])
    # the moment sums are FP64 atomics over tiles whose grouping depends on the block of the lag list a device gets
    assert np.max(np.abs(M1 - M2) / np.max(np.abs(M1), axis=(1, 2), keepdims=True)) < 1e-13


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from spinrelax_b200 import pipeline, shard, synth
        nC, nF, n_total = 2, 9000, 9                       # 9 vectors over 2 ranks: 5 + 4
        v4 = synth.nh_vectors(nC * nF, n_total, seed=505).reshape(nC, nF, n_total, 3)
        mine = np.ascontiguousarray(shard.shard_vectors(v4, world, rank))
        q = (0.83, -0.31, 0.22, 0.41)
        step = pipeline.CtHistStep(nC, nF, mine.shape[2], q_rot=q, device=torch.device("cuda", rank), world=world, rank=rank,
                                   n_total=n_total)
        step.run_device(torch.from_numpy(mine).cuda())
        dev_res = None if rank else (step.gathered.cpu().numpy(), step.gathered_hist.cpu().numpy())
        step.run_host(mine)                                  # the host-buffer path gathers the same things
        if rank == 0:
            np.savez(os.path.join(out_dir, "g.npz"), both=dev_res[0], hist=dev_res[1], both_host=step.gathered.cpu().numpy(),
                     hist_host=step.gathered_hist.cpu().numpy(), v4=v4)
        else:
            assert step.gathered is None and step.gathered_hist is None
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu_gathers_ct_and_histogram(tmp_path):
    _need2()
    import torch.multiprocessing as mp
    from oracle import ct_oracle
    mp.spawn(_rank_main, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "g.npz")
    v4 = g["v4"]
    L = v4.shape[1] // 2
    oCt, odCt = ct_oracle.ct_palmer(v4.astype(np.float64))
    for key in ("both", "both_host"):
        assert g[key].shape == (2 * L, 9)
        assert rel_err(g[key][:L], oCt) < 1e-6 and np.max(np.abs(g[key][L:] - odCt)) < 1e-6
    ho, _ = ct_oracle.sphere_histogram(v4.reshape(-1, 9, 3), np.array([0.83, -0.31, 0.22, 0.41]))
    assert np.array_equal(g["hist"].astype(np.int64), ho.astype(np.int64))
    assert np.array_equal(g["hist_host"].astype(np.int64), ho.astype(np.int64))
