"""N > 1 host logic on CPU: world_size-2 gloo process group exercising the vector / lag sharding helpers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spinrelax_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nR, nLags, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # vectors: each rank "computes" a (L, n_local) block whose entries encode (lag row, global vector id)
        a, b = shard.split_range(nR, world, rank)
        L = 7
        local = torch.arange(L, dtype=torch.float32)[:, None] * 1000 + torch.arange(a, b, dtype=torch.float32)[None, :]
        full = shard.gather_columns(local, nR, dst=0)
        # lags: round-robin shard, rows encode the lag value
        lags = np.arange(3, 3 + nLags)
        mine = shard.shard_lags(lags, world, rank)
        rows = torch.tensor(mine, dtype=torch.float64)[:, None].repeat(1, 6)
        merged = shard.merge_lag_results(rows, lags, world, dst=0)
        # frame-sharded histogram: all-reduce of integer counts
        h = torch.full((3, 4), rank + 1, dtype=torch.int64)
        shard.allreduce_sum_(h)
        # vector-sharded histogram: per-vector (nbx, nby) count blocks gathered along axis 0 (uneven shares included)
        hv = torch.arange(a, b, dtype=torch.int32)[:, None, None] * 100 + torch.arange(6, dtype=torch.int32).reshape(1, 3, 2)
        hall = shard.gather_rows(hv, nR, dst=0)
        # a block of the wrong width must be refused, not silently truncated / misplaced (uneven partitions)
        bad = False
        if nR % world:
            try:
                shard.gather_columns(torch.zeros((L, (b - a) + 1)), nR, dst=0)
            except ValueError:
                bad = True
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), full=full.numpy(), merged=merged.numpy(), h=h.numpy(), hall=hall.numpy(),
                     refused=bad or not nR % world)
        else:
            assert full is None and merged is None and hall is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nR,nLags", [(76, 100), (5, 3), (2, 1), (77, 10)])
def test_gloo_world2(tmp_path, nR, nLags):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nR, nLags, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r0.npz")
    L = 7
    expect = np.arange(L)[:, None] * 1000 + np.arange(nR)[None, :]
    assert np.array_equal(r["full"], expect.astype(np.float32))
    assert np.array_equal(r["merged"][:, 0], np.arange(3, 3 + nLags))
    assert (r["h"] == 3).all()
    assert np.array_equal(r["hall"], np.arange(nR)[:, None, None] * 100 + np.arange(6).reshape(1, 3, 2))
    assert bool(r["refused"])


def test_partition_properties():
    for n in (1, 2, 76, 1000, 2001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.split_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard.split_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            lags = np.arange(n)
            got = np.sort(np.concatenate([shard.shard_lags(lags, world, r) for r in range(world)]))
            assert np.array_equal(got, lags)
    v = np.zeros((2, 10, 7, 3))
    assert shard.shard_vectors(v, 2, 0).shape == (2, 10, 4, 3) and shard.shard_vectors(v, 2, 1).shape == (2, 10, 3, 3)


def test_multigpu_plan_partitions_units_over_selected_devices(monkeypatch):
    """Single-process multi-device sharding (spinrelax_b200.multigpu): contiguous balanced blocks, a minimum block size,
    and never more blocks than devices -- checked without a GPU by faking the device list."""
    from spinrelax_b200 import multigpu
    monkeypatch.setattr(multigpu, "devices", lambda: [0, 1, 2, 3, 4, 5, 6, 7])
    blocks = multigpu.plan(2000)
    assert [d for d, _, _ in blocks] == list(range(8)) and blocks[0][1] == 0 and blocks[-1][2] == 2000
    assert all(blocks[i][2] == blocks[i + 1][1] for i in range(7)) and {b - a for _, a, b in blocks} == {250}
    assert [(a, b) for _, a, b in multigpu.plan(7)] == [(0, 1)] * 0 + [(i, i + 1) for i in range(7)]      # fewer units than devices
    assert len(multigpu.plan(100, min_per_device=32)) == 3                                              # 34 + 33 + 33
    assert multigpu.plan(10, min_per_device=32) == [(0, 0, 10)]
    monkeypatch.setattr(multigpu, "devices", lambda: [3])
    assert multigpu.plan(76) == [(3, 0, 76)]


def _oracle_worker(rank, world, port, out_dir):
    """Each rank runs the oracle on its share (the stand-in for the device stage) and the product's shard helpers move
    the results: what rank 0 ends up with must be what the unsharded oracle produces."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ct_oracle, dq_oracle
    from spinrelax_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nC, nF, nR = 3, 200, 7                       # 7 vectors over 2 ranks: shares of 4 and 3
        v4 = synth.nh_vectors(nC * nF, nR, seed=11).reshape(nC, nF, nR, 3)
        q_rot = np.array([0.83, -0.31, 0.22, 0.41])
        a, b = shard.split_range(nR, world, rank)
        Ct, dCt = ct_oracle.ct_palmer(v4[:, :, a:b])
        both = torch.from_numpy(np.concatenate((Ct, dCt), axis=0))
        full = shard.gather_columns(both, nR, dst=0)
        h, _ = ct_oracle.sphere_histogram(v4[:, :, a:b].reshape(nC * nF, b - a, 3), q_rot)
        hall = shard.gather_rows(torch.from_numpy(h.astype(np.int64)), nR, dst=0)
        # the same histogram sharded by frames instead: one all-reduce of the counts
        fa, fb = shard.split_range(nC * nF, world, rank)
        hf, _ = ct_oracle.sphere_histogram(v4.reshape(nC * nF, nR, 3)[fa:fb], q_rot)
        hf = torch.from_numpy(hf.astype(np.int64))
        shard.allreduce_sum_(hf)
        # dq second moments: lags round-robin over the ranks (every rank holds the trajectory) ...
        q = synth.quaternion_walk(600, seed=5, sigma=(0.01, 0.015, 0.03))
        lags = np.arange(2, 40, 3)
        mine = shard.shard_lags(lags, world, rank)
        rows = np.array([np.concatenate(([dq_oracle.iso_moment_shipped(dq_oracle.self_dq(q, int(d))[..., 1:4])],
                                        dq_oracle.aniso_tensor(dq_oracle.self_dq(q, int(d))[..., 1:4]).ravel())) for d in mine])
        merged = shard.merge_lag_results(torch.from_numpy(rows), lags, world, dst=0)
        # ... and replica pooling (calculate-dq-distribution-multi.py:529-540): one trajectory per rank, raw sums all-reduced
        qr = synth.quaternion_walk(600, seed=20 + rank, sigma=(0.01, 0.015, 0.03))
        vq = dq_oracle.self_dq(qr, 5)[..., 1:4]
        sums = torch.from_numpy(np.concatenate(((vq[:, :, None] * vq[:, None, :]).sum(axis=0).ravel(), [len(vq)])))
        shard.allreduce_sum_(sums)
        if rank == 0:
            np.savez(os.path.join(out_dir, "o0.npz"), full=full.numpy(), hall=hall.numpy(), hf=hf.numpy(),
                     merged=merged.numpy(), pooled=sums.numpy())
    finally:
        dist.destroy_process_group()


def test_gloo_sharded_stages_equal_the_unsharded_oracle(tmp_path):
    """Bond vectors are independent (calculate-Ct-from-traj.py:222-228), histograms add over frames, lags are independent
    and replicas pool their samples: the sharded data flow of every stage reproduces the one-process result."""
    from oracle import ct_oracle, dq_oracle
    from spinrelax_b200 import synth
    world = 2
    mp.spawn(_oracle_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "o0.npz")
    nC, nF, nR = 3, 200, 7
    v4 = synth.nh_vectors(nC * nF, nR, seed=11).reshape(nC, nF, nR, 3)
    q_rot = np.array([0.83, -0.31, 0.22, 0.41])
    Ct, dCt = ct_oracle.ct_palmer(v4)
    assert np.array_equal(r["full"], np.concatenate((Ct, dCt), axis=0))
    h, _ = ct_oracle.sphere_histogram(v4.reshape(nC * nF, nR, 3), q_rot)
    assert np.array_equal(r["hall"], h.astype(np.int64)) and np.array_equal(r["hf"], h.astype(np.int64))
    q = synth.quaternion_walk(600, seed=5, sigma=(0.01, 0.015, 0.03))
    lags = np.arange(2, 40, 3)
    for row, d in zip(r["merged"], lags):
        vq = dq_oracle.self_dq(q, int(d))[..., 1:4]
        assert row[0] == dq_oracle.iso_moment_shipped(vq) and np.array_equal(row[1:], dq_oracle.aniso_tensor(vq).ravel())
    pooled = dq_oracle.pooled_vectors([synth.quaternion_walk(600, seed=20 + k, sigma=(0.01, 0.015, 0.03)) for k in range(world)], 5)
    assert r["pooled"][-1] == len(pooled)
    np.testing.assert_allclose(r["pooled"][:9] / r["pooled"][-1], dq_oracle.aniso_tensor(pooled).ravel(), rtol=1e-13, atol=1e-18)
