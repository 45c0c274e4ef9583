#!/usr/bin/env python
"""Headline benchmark: C(t) bond-vector x frame x lag pairs/s on BASELINE config 2
(76 N-H vectors, 10^6 frames = 5 chunks x 2e5, max lag 1e5, + PAF rotation and spherical histogram).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over the batch: K2 pack -> K1 lag sums -> Palmer finalize
(-> K3 rotation + histogram).  `value` is timed with inputs resident in HBM, `e2e` goes through the
public host-buffer call (pinned host input, H2D, kernels, D2H).

N = 1: BASELINE config 2 (the configuration the metric is quoted on), plus -- outside the timed region, in the same JSON
line under `secondary` -- the other BASELINE metrics: config 3 dq moments (run-all lag set and all windows), config 5
fits and the R1/R2/NOE grid, each with its own roofline / CPU baseline and the clocks sampled while they ran.
N > 1: one process per GPU (torchrun), BASELINE config 4: 1000 N-H + 1000 Calpha-Halpha vectors x 10^6 frames,
sharded by bond vector over the ranks (strong scaling: the job is the same at every N > 1; no data-path collective,
the (2L, nR) result columns and the per-vector histograms are gathered to rank 0 inside the timed region).  After the
timed region every rank checks four of its vectors against the FFT oracle and the line carries the worst deviation;
`secondary` then holds the two collective-terminated stages (pooled dq moments, frame-sharded histogram).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload = BASELINE.json configs[1]
N_VEC, N_CHUNK, N_FRAMES_PER_CHUNK = 76, 5, 200000
Q_PAF = (0.83, -0.31, 0.22, 0.41)
METRIC = "ct_pairs_per_s"
UNIT = "bond-vector*frame*lag pairs/s"


C4_N_VEC = 2000        # BASELINE configs[3]: 1000 N-H + 1000 Calpha-Halpha


def n_pairs(nR, nC, nF, L):
    return nR * nC * (L * nF - L * (L + 1) // 2)


def workload_config(n_gpus):
    if n_gpus > 1 or os.environ.get("SPINRELAX_BENCH_C4"):
        from spinrelax_b200 import shard
        sizes = shard.split_sizes(C4_N_VEC, n_gpus)
        return {"workload": "BASELINE config 4: %d bond vectors (1000 N-H + 1000 Calpha-Halpha) x 1e6 frames (%d chunks x %d), "
                            "max lag %d, C(t) Palmer + PAF rotation + 72x36 vecHistogram, sharded by bond vector over %d GPUs"
                            % (C4_N_VEC, N_CHUNK, N_FRAMES_PER_CHUNK, N_FRAMES_PER_CHUNK // 2, n_gpus),
                "n_vectors": C4_N_VEC, "n_vectors_per_gpu": sizes, "n_chunks": N_CHUNK, "frames_per_chunk": N_FRAMES_PER_CHUNK,
                "max_lag": N_FRAMES_PER_CHUNK // 2, "sharding": "bond vectors, contiguous balanced blocks (shard.split_range)",
                "l2": "inputs (%.1f GB AoS per GPU) exceed the 126 MB L2; no flush needed" % (max(sizes) * 12e6 / 1e9),
                "n1_workload": "bench.py --gpus 1 runs BASELINE config 2 (76 vectors); per-GPU work here is %d vectors" % max(sizes),
                "pairs_per_step": n_pairs(C4_N_VEC, N_CHUNK, N_FRAMES_PER_CHUNK, N_FRAMES_PER_CHUNK // 2)}
    return {"workload": "BASELINE config 2: %d N-H vectors x 1e6 frames (%d chunks x %d), max lag %d, "
                        "C(t) Palmer + PAF rotation + 72x36 vecHistogram" % (N_VEC, N_CHUNK, N_FRAMES_PER_CHUNK,
                                                                             N_FRAMES_PER_CHUNK // 2),
            "n_vectors_per_gpu": N_VEC, "n_chunks": N_CHUNK, "frames_per_chunk": N_FRAMES_PER_CHUNK,
            "max_lag": N_FRAMES_PER_CHUNK // 2, "sharding": "bond vectors, %d per rank" % N_VEC,
            "l2": "inputs (0.9 GB AoS, 0.9 GB packed SoA) exceed the 126 MB L2; no flush needed",
            "pairs_per_step_per_gpu": n_pairs(N_VEC, N_CHUNK, N_FRAMES_PER_CHUNK, N_FRAMES_PER_CHUNK // 2)}


def make_input(rank):
    from spinrelax_b200 import synth
    v = synth.nh_vectors(N_CHUNK * N_FRAMES_PER_CHUNK, N_VEC, seed=synth.BASE_SEED + 2 + 1000 * rank)
    return v.reshape(N_CHUNK, N_FRAMES_PER_CHUNK, N_VEC, 3)


def make_input_c4_device(rank, world, dev):
    """This rank's block of the config-4 vector set, built on the device: every vector is one of the 76 seeded
    config-2 trajectories rolled along the frame axis of each chunk and turned by its own fixed rotation (same
    autocorrelation, different data), so no rank ever holds more than its own share."""
    import torch
    from spinrelax_b200 import shard, synth
    base = torch.from_numpy(make_input(0)).to(dev)                       # (nC, nF, 76, 3)
    rng = np.random.default_rng(synth.BASE_SEED + 4040)
    quat = rng.standard_normal((C4_N_VEC, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    shift = rng.integers(0, N_FRAMES_PER_CHUNK, C4_N_VEC)
    a, b = shard.split_range(C4_N_VEC, world, rank)
    out = torch.empty((N_CHUNK, N_FRAMES_PER_CHUNK, b - a, 3), dtype=torch.float32, device=dev)
    for j in range(a, b):
        w, x, y, z = quat[j]
        R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float32, device=dev)
        out[:, :, j - a] = torch.roll(base[:, :, j % N_VEC], int(shift[j]), dims=1) @ R.T
    del base
    return out


def spot_check(step, v_host4, cols):
    """After the timed region: C(t) of a few of this rank's vectors against the FFT oracle (float64, a different
    algorithm), and every histogram must have counted each frame once.  Returns the worst relative deviation."""
    from oracle import ct_oracle
    S = ct_oracle.ct_lag_sums_fft(v_host4[:, :, cols])
    oCt, _ = ct_oracle.ct_from_lag_sums(S, v_host4.shape[1], dtype=np.float64)
    ours = step.Ct[:, cols].cpu().numpy().astype(np.float64)
    return float(np.max(np.abs(ours - oCt) / np.abs(oCt)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        busy = [x for x in sm if x > 0.5 * max(mx)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    p = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}
    if os.path.exists(path):
        with open(path) as fp:
            m = json.load(fp)
        p = {"hbm_gbs": m.get("hbm_gbs", 6650.0), "sm_max_mhz": m.get("sm_max_mhz", 1965.0),
             "source": "MEASURED_PEAKS.json"}
    # FP32 CUDA-core FMA peak is not in MEASURED_PEAKS.json: 148 SM x 128 lanes x 2 flop x measured max SM clock
    p["fp32_tflops"] = 148 * 128 * 2 * p["sm_max_mhz"] * 1e6 / 1e12
    return p


# =====================================================================================================
# reference arm: the reference's own CPU implementation of the path (NumPy einsum per lag,
# calculate-Ct-from-traj.py:222-228) via the oracle port, on all host cores, bounded lag sample per step
# =====================================================================================================
_REF_V = None


def _ref_task(args):
    lag, r0, r1 = args
    from oracle import ct_oracle
    ct_oracle.ct_lag_body(_REF_V[:, :, r0:r1], lag)
    nC, nF = _REF_V.shape[0], _REF_V.shape[1]
    return nC * (nF - lag) * (r1 - r0)


def sample_lags(n):
    L = N_FRAMES_PER_CHUNK // 2
    return sorted(set(int(x) for x in np.unique(np.round(np.linspace(1, L, n)))))


def run_reference(args):
    global _REF_V
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(os.cpu_count() or 1, 32)
    _REF_V = make_input(0)
    lags = sample_lags(max(2, cores // 2))
    blocks = [(0, 19), (19, 38), (38, 57), (57, 76)]
    tasks = [(lag, a, b) for lag in lags for (a, b) in blocks]
    ctx = mp.get_context("fork")
    times, pairs = [], 0
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pairs = sum(pool.map(_ref_task, tasks, chunksize=1))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = pairs * len(times) / total
    sample = ("oracle port of the per-lag body (calculate-Ct-from-traj.py:223-228) on the full config-2 array, "
              "%d lags spread over 1..1e5 x 4 vector blocks per step, %d worker processes" % (len(lags), cores))
    if args.gpus > 1:
        sample += ("; the config-4 job is the same loop body over 2000 instead of 76 vectors (cost per pair identical), "
                   "sampled here on the 76 seeded base trajectories the config-4 set is built from")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline_single(v4):
    """Reference path, single thread as shipped: exact per-lag body on a bounded lag sample (~10-20 s)."""
    from oracle import ct_oracle
    lags = sample_lags(12)
    t0 = time.perf_counter()
    pairs = 0
    for lag in lags:
        ct_oracle.ct_lag_body(v4, lag)
        pairs += v4.shape[0] * (v4.shape[1] - lag) * v4.shape[2]
        if time.perf_counter() - t0 > 25.0:
            break
    dt = time.perf_counter() - t0
    return {"value": pairs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle.ct_oracle.ct_lag_body (calculate-Ct-from-traj.py:223-228) on the full config-2 "
                      "float32 array for lags %s; %.1f s" % (lags[:len(lags)], dt),
            "host_cores": os.cpu_count()}


# =====================================================================================================
# our arm
# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from spinrelax_b200 import _lib, ct, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner out of the one-JSON-line stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _lib.load()

    nC, nF = N_CHUNK, N_FRAMES_PER_CHUNK
    L = nF // 2
    force_c4 = bool(os.environ.get("SPINRELAX_BENCH_C4"))      # capture config 4 on one GPU too (profiles/, not the default)
    if world == 1 and not force_c4:
        v_host = torch.from_numpy(make_input(rank)).pin_memory()
        nR = n_total = N_VEC
        v_dev = v_host.to(dev, non_blocking=True)
    else:
        v_dev = make_input_c4_device(rank, world, dev)
        nR, n_total = v_dev.shape[2], C4_N_VEC
        v_host = torch.empty(v_dev.shape, dtype=torch.float32).pin_memory()
        v_host.copy_(v_dev)
    torch.cuda.synchronize()
    step = pipeline.CtHistStep(nC, nF, nR, q_rot=Q_PAF, device=dev, world=world, rank=rank, n_total=n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step.run_device(v_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step.reset_kernel_timers()
    e0.record()
    for _ in range(args.steps):
        step.run_device(v_dev, time_kernels=True)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    kt = step.kernel_times_ms()           # per-kernel average launch duration (CUDA events on the launch stream)

    # ---- end to end through the public host API ------------------------------------------------------
    v_np = v_host.numpy()
    for _ in range(min(args.warmup, 1)):
        step.run_host(v_np)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        out = step.run_host(v_np)
    barrier()
    e2e_s = (time.perf_counter() - t0)

    # the reference-signature drop-in itself: calculate_Ct_Palmer(vecs) on an ordinary (pageable) NumPy array through
    # the single C-ABI host call sr_ct_palmer_host (C(t) only: the histogram is a separate call in the reference too)
    dropin_s = None
    if world == 1 and not force_c4:
        v_page = np.array(v_np)                      # pageable copy
        ct.calculate_Ct_Palmer_quiet(v_page)                # warm-up: sizes the library's scratch cache
        t0 = time.perf_counter()
        ct.calculate_Ct_Palmer_quiet(v_page)
        dropin_s = time.perf_counter() - t0
        del v_page

    # ---- after the timed regions: result check on every rank, worst deviation to rank 0 -------------------------------
    step.run_device(v_dev)
    torch.cuda.synchronize()
    cols = sorted(set([0, nR // 3, (2 * nR) // 3, nR - 1]))
    dev_ct = spot_check(step, v_np, cols)
    hist_ok = 1.0
    if step.has_hist:
        _, _, h = step.run_device(v_dev)
        hist_ok = float(np.all(h.reshape(nR, -1).sum(axis=1) == nC * nF))
    t = torch.tensor([ms_total, e2e_s, dev_ct, -hist_ok], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, dev_ct, hist_ok = float(t[0]), float(t[1]), float(t[2]), -float(t[3])
    gathered_ok = None
    if world > 1 and rank == 0:
        # the gathered columns of this rank's own block must be the rank's own result, bit for bit
        a0, b0 = 0, nR
        gathered_ok = bool(torch.equal(step.gathered[:L, a0:b0], step.Ct) and step.gathered.shape == (2 * L, n_total)
                           and step.gathered_hist.shape[0] == n_total)

    secondary = None
    try:
        if not force_c4:
            secondary = run_secondary(world, rank, local, dev)
    except Exception as exc:                                   # the headline line must survive a secondary failure
        secondary = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        pairs_gpu = n_pairs(nR, nC, nF, L)                    # rank 0 holds a largest block of the partition
        pairs_job = n_pairs(n_total, nC, nF, L)
        value = pairs_job * args.steps / (ms_total * 1e-3)
        pk = peaks()
        lag_ms = kt["ct_lag_kernel"]
        achieved = pairs_gpu * 7 / (lag_ms * 1e-3) / 1e12
        roof = {"kernel": "ct_lag_kernel", "bound": "fp32", "achieved": achieved, "peak": pk["fp32_tflops"],
                "unit": "TFLOP/s", "frac": achieved / pk["fp32_tflops"], "traffic": step.ncu_traffic_bytes(),
                "flop_per_pair": 7, "peak_source": "148 SM x 128 FP32 lanes x 2 x sm_max_mhz (%s); FP32 CUDA-core "
                "peak is not in MEASURED_PEAKS.json, measured FFMA microbenchmark = 72.5 TFLOP/s "
                "(profiles/r01_microbench_b200.jsonl)" % pk["source"],
                "share_of_step": lag_ms / (ms_total / args.steps), "kernel_ms": kt}
        if "sphere_hist_kernel" in kt:
            samples = nC * nF * nR
            gbs = samples * 12 / (kt["sphere_hist_kernel"] * 1e-3) / 1e9
            roof["streaming"] = {"kernel": "sphere_hist_kernel", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "bytes_per_sample": 12}
        cpu = cpu_baseline_single(v_np) if (world == 1 and not force_c4) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak" if (world == 1 and not force_c4) else "strong", "vs_baseline": None, "dtype": "f32 products, f64 accumulation",
                "data": "synthetic", "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": pairs_job * e2e_steps / e2e_s, "unit": UNIT,
                        "h2d_bytes_per_step": int(v_np.nbytes), "d2h_bytes_per_step": int(step.d2h_bytes()),
                        "steps": e2e_steps, "api": "spinrelax_b200.pipeline.CtHistStep.run_host: pinned NumPy in -> per-chunk H2D pipelined with "
                        "sr_pack_vectors_f32_chunks, sr_ct_lag_sums_chunks -> sr_ct_palmer_finalize, sr_sphere_hist (C ABI) -> D2H"},
                "gpu_launches": step.launches_per_step() * args.steps, "roofline": roof, "cpu_baseline": cpu,
                "check": {"ct_max_rel_dev_vs_fft_oracle": dev_ct, "vectors_checked_per_rank": len(cols),
                          "every_frame_binned_once": bool(hist_ok), "gathered_block_equals_local": gathered_ok,
                          "when": "after the timed regions, every rank, max over ranks"},
                "secondary": secondary}
        if dropin_s is not None:
            line["e2e_dropin"] = {"value": pairs_gpu / dropin_s, "unit": UNIT, "seconds": dropin_s,
                                  "api": "spinrelax_b200.ct.calculate_Ct_Palmer(vecs) on a pageable NumPy array -> "
                                         "sr_ct_palmer_host (C ABI, pinned staging + chunk pipelining inside)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_secondary(world, rank, local, dev):
    """The other BASELINE metrics, measured right after the headline with their own clock samples (rank 0 reports).
    N = 1: config 3 dq moments, config 5 fits and relaxation grid (bench_secondary.py).  N > 1: the two stages of the
    path that end in a collective (bench_multi.py)."""
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world == 1:
        import bench_secondary
        lines = bench_secondary.bench_dq(False, emit=False) + bench_secondary.bench_fit_relax(False, emit=False)
    else:
        import bench_multi
        lines = bench_multi.run(emit=False)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    return {"clocks": clocks, "lines": lines}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
