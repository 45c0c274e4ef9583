#!/usr/bin/env python
"""Secondary benchmarks for the other BASELINE configs (the headline config-2 line lives in bench.py):

  c3  calculate-dq-distribution on a 1e6-frame quaternion trajectory: run-all lag set (100 windows, 4 chunks)
      and all windows (lags 1..1e5) -> frame*lag pairs/s
  c5  batched multi-exponential C(t) fits for 1000 residues -> residues/s, then J(omega) -> R1/R2/NOE for
      1000 residues x 5 fields x 64-point CSA grid -> evaluations/s

One JSON line per measurement, each with a `cpu_baseline` from the oracle port on a bounded sample.
    python bench_secondary.py [--quick]
"""
import argparse
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ev_ms(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def bench_dq(quick, emit=True):
    import torch
    from oracle import dq_oracle
    from spinrelax_b200 import _lib, dq, synth
    lib = _lib.load()
    N = 1000000
    q = synth.quaternion_walk(N, seed=synth.BASE_SEED + 3, sigma=(0.004, 0.006, 0.012))
    qd = torch.from_numpy(q).cuda()
    out = []
    sets = [("c3 run-all lag set: 100 windows (1000..100000 step 1000), 4 chunks", np.arange(1000, 100001, 1000))]
    if not quick:
        sets.append(("c3 all windows: lags 1..100000, 4 chunks", np.arange(1, 100001)))
    for name, lags in sets:
        ld = torch.from_numpy(lags.astype(np.int64)).cuda()
        M = torch.empty((len(lags), 4, 6), dtype=torch.float64, device="cuda")
        ms = ev_ms(torch, lambda: _lib.check(lib.sr_dq_moments(qd.data_ptr(), N, ld.data_ptr(), len(lags), int(lags.min()), 4,
                                                               M.data_ptr(), None)))
        pairs = float(np.sum(N - lags))
        # end to end like the CLI (calculate-dq-distribution.py --iso --aniso --num_chunk 4): H2D, the reduction over ALL lags
        # of the list, the per-lag eigen-frames / rotated tensors, and the 20 Powell fits (iso, 4 chunk-iso, 3 aniso, 12
        # chunk-aniso) that turn the curves into D
        import contextlib
        t0 = time.perf_counter()
        res = dq.dq_curves(q, lags, 10.0, nchunk=4)
        t_curves = time.perf_counter() - t0
        with contextlib.redirect_stdout(io.StringIO()):
            taus = [dq.conduct_exponential_fit(res["dt"], res["iso"], 1.5, -0.5)]
            taus += [dq.conduct_exponential_fit(res["dt"], res["chunk_iso"][i], 1.5, -0.5) for i in range(4)]
            taus += [dq.conduct_exponential_fit(res["dt"], res["aniso2"][i], 0.5, 0.5) for i in range(3)]
            taus += [dq.conduct_exponential_fit(res["dt"], res["chunk_aniso2"][i][j], 0.5, 0.5) for i in range(4) for j in range(3)]
        e2e_s = time.perf_counter() - t0
        # 40 FP64 flop per pair: 28 for the quaternion product + 12 for the six moments
        line = {"metric": "dq_pairs_per_s", "value": pairs / ms * 1e3, "unit": "frame*lag pairs/s", "n_gpus": 1,
                "ms_per_step": ms, "config": {"workload": name, "n_frames": N, "n_lags": int(len(lags))},
                "dtype": "f64 on f32-rounded input", "data": "synthetic",
                "roofline": {"kernel": "dq_moments_kernel", "bound": "fp64", "achieved": pairs * 40 / ms * 1e-9,
                             "unit": "TFLOP/s", "peak": 148 * 64 * 2 * 1.965e9 / 1e12,
                             "peak_source": "148 SM x 64 FP64 lanes x 2 x 1965 MHz (measured DFMA microbenchmark 34.2)",
                             "frac": pairs * 40 / ms * 1e-9 / (148 * 64 * 2 * 1.965e9 / 1e12)},
                "e2e": {"seconds": e2e_s, "curves_seconds": t_curves, "value": pairs / e2e_s, "unit": "frame*lag pairs/s",
                        "what": "dq.dq_curves over the whole lag list (H2D, kernels, D2H, stacked eigh / frame quaternions) + 20 "
                                "SciPy Powell fits, as the CLI runs them", "D_iso_from_fit": 0.5e12 / taus[0]}}
        out.append(line)
    # CPU baseline: the reference's per-lag body (obtain_self_dq + iso + tensor + chunks) via the oracle port
    t0 = time.perf_counter()
    pairs = 0
    for d in (1000, 50000, 100000):
        v = dq_oracle.self_dq(q, d)[..., 1:4]
        dq_oracle.iso_moment_shipped(v); dq_oracle.aniso_tensor(v); dq_oracle.iso_moment_chunks(v, 4)
        dq_oracle.aniso_tensor_chunks(v, 4)
        pairs += len(v)
    cpu = {"value": pairs / (time.perf_counter() - t0), "unit": "frame*lag pairs/s", "cores": 1, "kind": "port",
           "sample": "oracle per-lag body (calculate-dq-distribution.py:560-625) for lags 1000, 50000, 100000 on the full trajectory"}
    for line in out:
        line["cpu_baseline"] = cpu
        if emit:
            print(json.dumps(line))
    return out


def synth_curves(n_res, n_pts, seed):
    rng = np.random.default_rng(seed)
    t = (np.arange(n_pts) + 1.0) * 10.0
    Y, SG = np.zeros((n_res, n_pts)), np.zeros((n_res, n_pts))
    for i in range(n_res):
        S2 = rng.uniform(0.45, 0.9)
        nc = 1 + i % 3
        C = rng.dirichlet(np.ones(nc)) * (1 - S2) * rng.uniform(0.85, 1.0)
        tau = np.sort(10 ** rng.uniform(1.2, 3.2, nc))
        sig = 0.002 + 0.004 * t / t[-1]
        Y[i] = S2 + np.sum(C[:, None] * np.exp(-t[None] / tau[:, None]), axis=0) + rng.standard_normal(n_pts) * sig * 0.5
        SG[i] = sig
    return t, Y, SG


def bench_fit_relax(quick, emit=True):
    import torch
    from oracle import fit_oracle, sd_oracle                    # cpu_baseline legs only
    from spinrelax_b200 import fitct, hist as sr_hist, specdens as sd, synth
    nR = 200 if quick else 1000
    t, Y, SG = synth_curves(nR, 500, 77)
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(nR)], [t] * nR, Y, SG)
    ac.fit_all_residues(fp=io.StringIO())                       # warm-up (module load, allocations)
    fitct.KERNEL_EVENTS = []
    t0 = time.perf_counter()
    ac.fit_all_residues(fp=io.StringIO())
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    kern_ms = sum(a.elapsed_time(b) for a, b in fitct.KERNEL_EVENTS)
    fitct.KERNEL_EVENTS = None
    t0 = time.perf_counter()
    nref = 8
    for i in range(nref):
        fit_oracle.fit_ladder(t, Y[i], SG[i])
    cpu_fit = nref / (time.perf_counter() - t0)
    lines = []
    lines.append(({"metric": "ct_fit_residues_per_s", "value": nR / fit_s, "unit": "residues/s (full 2-3-5-7-9 ladder, "
                      "500-point curves, host selection logic included)", "n_gpus": 1, "ms_per_step": fit_s * 1e3,
                      "config": {"workload": "c5 fits: %d residues x 500-point C(t)" % nR}, "dtype": "f64",
                      "kernel_ms": kern_ms, "kernel_only_residues_per_s": nR / (kern_ms * 1e-3),
                      "note": "ct_fit_trf_kernel time summed over the five rungs; the rest of ms_per_step is the reference's "
                              "per-residue selection ladder on the host (fitting_Ct_functions.py:278-304)",
                      "data": "synthetic", "roofline": None,
                      "cpu_baseline": {"value": cpu_fit, "unit": "residues/s", "cores": 1, "kind": "port",
                                       "sample": "oracle fit_ladder (SciPy curve_fit TRF, fitting_Ct_functions.py:278-345) on %d residues" % nref}}))
    # relaxation grid: histogram weights from a synthetic rotated stream, 5 fields x 64 CSA values
    q = np.array([0.8, -0.36, 0.48, 0.0])
    v = synth.nh_vectors(2000, nR, seed=5)
    hist, edges = sr_hist.sphere_histogram(v, q)                # input preparation (not timed): the product's K3
    vecs, w = sd.convert_LambertCylindricalHist_to_vecs(hist, edges)
    rot = sd.globalRotationalDiffusion_Axisymmetric(D=[2.1e-5, 1.35])
    rot.set_frame_vectors(np.arange(nR), vecs, w)
    ac.set_zeta(0.890023)
    fields = [500.0, 600.133, 700.0, 800.0, 950.0]
    csa = np.linspace(-220e-6, -120e-6, 64)
    sd.relax_grid(rot, ac, fields, csa)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rot._amom = None                                            # include the A-moment pass over the histogram
    res = sd.relax_grid(rot, ac, fields, csa)
    torch.cuda.synchronize()
    rel_s = time.perf_counter() - t0
    n_eval = nR * len(fields) * len(csa)
    models = [(m.S2, m.C, m.tau) for m in ac.model.values()]
    vec, wts = sd_oracle.hist_to_vectors(hist.astype(np.float64), edges)
    nref = 20
    t0 = time.perf_counter()
    sd_oracle.relax_axisymmetric(600.133, 2.1e-5, 1.35, vec, wts[:nref], models[:nref], zeta=0.890023)
    cpu_rel = nref / (time.perf_counter() - t0)
    lines.append(({"metric": "relax_evaluations_per_s", "value": n_eval / rel_s, "unit": "residue*field*CSA evaluations/s "
                      "(each = R1, R2, NOE with mean and sigma over 2592 bin vectors; host packing + D2H included)",
                      "n_gpus": 1, "ms_per_step": rel_s * 1e3,
                      "config": {"workload": "c5 relaxation: %d residues x 5 fields x 64 CSA, 72x36 histogram" % nR},
                      "residues_per_s": nR / rel_s, "dtype": "f64", "data": "synthetic", "roofline": None,
                      "cpu_baseline": {"value": cpu_rel, "unit": "residue*field*CSA evaluations/s", "cores": 1, "kind": "port",
                                       "sample": "oracle relax_axisymmetric (spectral_densities.py:552-557,751-763,824-907) "
                                                 "on %d residues, 1 field, 1 CSA" % nref}}))
    if emit:
        for l in lines:
            print(json.dumps(l))
    return lines


def bench_rtp(emit=True):
    """--vecDist spherical conversion on the config-2 stream (25.6M float32 vectors): HBM-bound, 24 B/vector."""
    import torch
    from oracle import ct_oracle
    from spinrelax_b200 import gm, synth
    nF, nR = 100000, 256
    v = synth.nh_vectors(nF, nR, seed=synth.BASE_SEED + 2)
    vd = torch.from_numpy(v).cuda()
    ms = ev_ms(torch, lambda: gm.xyz_to_rtp_device(vd), reps=5)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n = nF * nR
    sub = v[:4000]
    t0 = time.perf_counter()
    ct_oracle.xyz_to_rtp(sub)
    cpu = sub.shape[0] * nR / (time.perf_counter() - t0)
    line = ({"metric": "rtp_vectors_per_s", "value": n / ms * 1e3, "unit": "vectors/s", "n_gpus": 1,
                      "ms_per_step": ms, "config": {"workload": "c2 stream: gm.xyz_to_rtp on (100000, 256, 3) float32"},
                      "dtype": "f32", "data": "synthetic",
                      "roofline": {"kernel": "xyz_to_rtp_vec4_kernel<float>", "bound": "hbm", "achieved": n * 24 / ms * 1e-6,
                                   "unit": "GB/s", "peak": peak, "frac": n * 24 / ms * 1e-6 / peak},
                      "cpu_baseline": {"value": cpu, "unit": "vectors/s", "cores": 1, "kind": "port",
                                       "sample": "oracle xyz_to_rtp (general_maths.py:143-158) on 4000 frames x 256"}})
    if emit:
        print(json.dumps(line))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="dq | fit | rtp")
    args = ap.parse_args()
    if args.only in ("", "dq"):
        bench_dq(args.quick)
    if args.only in ("", "fit"):
        bench_fit_relax(args.quick)
    if args.only in ("", "rtp"):
        bench_rtp()


if __name__ == "__main__":
    main()
