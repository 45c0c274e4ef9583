#!/usr/bin/env python
"""Multi-GPU secondary benchmark: the two stages of the path that end in a collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_multi.py

  * dq moments over replica trajectories (calculate-dq-distribution-multi.py:529-540): every rank holds one replica of
    10^6 quaternions, runs sr_dq_moments_pooled(replica = rank, nReplicas = world) over all windows (lags 1..10^5)
    and ONE NCCL all-reduce (sum, float64, 19 MB) yields the pooled moments on every rank.
  * PAF rotation + histogram sharded by frames: every rank bins its own 10^6-frame block of the 76 vectors and ONE NCCL
    all-reduce (sum, int32, 0.8 MB) yields the counts of the whole trajectory.
Both are weak scaling; times are CUDA events, barrier on both sides, max over ranks.  One JSON line each (rank 0).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run(emit=True):
    """Both stages on the already initialised process group (bench.py calls this after its own timed region);
    returns the two JSON lines on rank 0, [] elsewhere."""
    import torch
    import torch.distributed as dist
    from spinrelax_b200 import _lib, hist, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    lib = _lib.load()
    lines = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=3):
        fn()
        barrier()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(); fn(); b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t[0]))
        return best

    # ---- dq moments, one replica per rank ----------------------------------------------------------------
    N, nCh = 1000000, 4
    q = synth.quaternion_walk(N, seed=synth.BASE_SEED + 300 + rank, sigma=(0.004, 0.006, 0.012))
    qd = torch.from_numpy(q).to(dev)
    lags = np.arange(1, 100001, dtype=np.int64)
    ld = torch.from_numpy(lags).to(dev)
    M = torch.empty((len(lags), nCh, 6), dtype=torch.float64, device=dev)

    def dq_step():
        _lib.check(lib.sr_dq_moments_pooled(qd.data_ptr(), N, ld.data_ptr(), len(lags), 1, nCh, rank, world, 0, M.data_ptr(),
                                            _lib.current_stream_ptr()), "sr_dq_moments_pooled")
        if world > 1:
            dist.all_reduce(M, op=dist.ReduceOp.SUM)

    ms = timed(dq_step)
    pairs = float(np.sum(N - lags)) * world
    # self-check: lag 1000 of the pooled moments against float64 NumPy on the gathered replicas (rank 0)
    k = 999
    v_loc = None
    from spinrelax_b200 import dq as dqmod
    v_loc = torch.from_numpy(dqmod.obtain_self_dq(q, int(lags[k]))[:, 1:4].copy()).to(dev)
    outer = torch.einsum("ti,tj->ij", v_loc, v_loc)
    if world > 1:
        dist.all_reduce(outer, op=dist.ReduceOp.SUM)
    got = M[k].sum(dim=0)
    ref = torch.stack((outer[0, 0], outer[0, 1], outer[0, 2], outer[1, 1], outer[1, 2], outer[2, 2]))
    err = float(torch.max(torch.abs(got - ref) / torch.abs(ref)))
    if rank == 0:
        lines.append(({"metric": "dq_pairs_per_s", "value": pairs / ms * 1e3, "unit": "frame*lag pairs/s", "n_gpus": world,
                          "ms_per_step": ms, "scaling": "weak", "collective": "1 x NCCL all-reduce sum f64 (%d bytes)"
                          % (M.numel() * 8), "config": {"workload": "c3 all windows, one 1e6-frame replica per rank, pooled "
                          "moments (calculate-dq-distribution-multi semantics), 4 sub-chunks"}, "data": "synthetic",
                          "pooled_moment_rel_err_lag1000": err}))
    assert err < 1e-10, err

    # ---- histogram, frames sharded -------------------------------------------------------------------------
    nR, F = 76, 1000000
    v = torch.from_numpy(synth.nh_vectors(F, nR, seed=synth.BASE_SEED + 400 + rank)).to(dev)
    acc = hist.SphereHistogram(nR, device=dev)
    qrot = np.array([0.83, -0.31, 0.22, 0.41])

    def hist_step():
        acc.accumulate_device(v, qrot, reset=True)
        if world > 1:
            dist.all_reduce(acc.counts, op=dist.ReduceOp.SUM)

    ms = timed(hist_step)
    # counts of the local tie-break samples are added on the host by every rank for its own block
    acc.accumulate_device(v, qrot, reset=True)
    local_counts = acc.finish(v, qrot)
    tot = torch.from_numpy(local_counts).to(dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ok = bool((tot.sum(dim=(1, 2)) == F * world).all())
    if rank == 0:
        lines.append(({"metric": "hist_samples_per_s", "value": F * nR * world / ms * 1e3, "unit": "vector*frame samples/s",
                          "n_gpus": world, "ms_per_step": ms, "scaling": "weak",
                          "collective": "1 x NCCL all-reduce sum i32 (%d bytes)" % (acc.counts.numel() * 4),
                          "config": {"workload": "PAF rotation + 72x36 histogram, 76 vectors, 1e6 frames per rank"},
                          "data": "synthetic", "every_sample_counted_once": ok}))
    assert ok
    if emit:
        for l in lines:
            print(json.dumps(l))
    return lines


def main():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    run(emit=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
