"""Scratch GPU check for the C(t) path: correctness against a float64 numpy evaluation on small
shapes, then timing of the lag kernel on the BASELINE config-2 shape.  Writes gpurun_out/ct_check.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import _lib, ct, synth  # noqa: E402


def ref_lag_sums(v):
    """float64 direct evaluation: S[r, c, d-1] = sum_t (u_t . u_{t+d})^2"""
    v = v.astype(np.float64)
    nC, nF, nR, _ = v.shape
    L = nF // 2
    S = np.zeros((nR, nC, L))
    for d in range(1, L + 1):
        dots = np.einsum("ijkl,ijkl->ijk", v[:, :-d], v[:, d:])
        S[:, :, d - 1] = np.einsum("ijk->ki", dots * dots)
    return S


def main():
    out = {}
    lib = _lib.load()
    for shape in [(2, 100, 3), (3, 999, 5), (10, 1000, 76), (2, 4001, 2)]:
        nC, nF, nR = shape
        v = synth.nh_vectors(nC * nF, nR, seed=5 + nF).reshape(nC, nF, nR, 3)
        S_ref = ref_lag_sums(v)
        vt = torch.from_numpy(v).cuda()
        S = ct.ct_lag_sums_device(vt).cpu().numpy()
        rel = np.max(np.abs(S - S_ref) / np.abs(S_ref))
        Ct, dCt = ct.ct_palmer_device(vt)
        out["S_rel_err_%dx%dx%d" % shape] = float(rel)
        print(shape, "max rel err S", rel, "Ct[0,:3]", Ct[0, :3].cpu().numpy())
    # timing on config 2: 76 vectors, 5 chunks x 200000 frames, L = 100000
    nC, nF, nR = 5, 200000, 76
    g = torch.Generator(device="cuda").manual_seed(1)
    vt = torch.randn((nC, nF, nR, 3), device="cuda", generator=g)
    vt = (vt / vt.norm(dim=-1, keepdim=True)).contiguous()
    L = nF // 2
    pitch = lib.sr_ct_row_pitch(nF)
    packed = torch.empty((nR, nC, 3, pitch), dtype=torch.float32, device="cuda")
    S = torch.empty((nR, nC, L), dtype=torch.float64, device="cuda")
    st = _lib.current_stream_ptr()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(2):
        ev[0].record()
        _lib.check(lib.sr_pack_vectors_f32(vt.data_ptr(), nC, nF, nR, None, packed.data_ptr(), pitch, st))
        ev[1].record()
        _lib.check(lib.sr_ct_lag_sums(packed.data_ptr(), pitch, nC, nF, nR, L, S.data_ptr(), st))
        ev[2].record()
        torch.cuda.synchronize()
        pairs = nR * nC * (L * nF - L * (L + 1) // 2)
        ms_pack, ms_lag = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        print("iter", it, "pack ms", ms_pack, "lag ms", ms_lag, "pairs/s %.4g" % (pairs / ms_lag * 1e3),
              "TFLOP/s(7/pair) %.2f" % (pairs * 7 / ms_lag * 1e-9))
        out["c2_pack_ms"], out["c2_lag_ms"], out["c2_pairs_per_s"] = ms_pack, ms_lag, pairs / ms_lag * 1e3
    # spot-check a few lags of config 2 against float64 on the host
    vh = vt[:, :, :2].cpu().numpy().astype(np.float64)
    errs = []
    for d in (1, 2, 479, 480, 481, 50000, 99999, 100000):
        dots = np.einsum("ijkl,ijkl->ijk", vh[:, :-d], vh[:, d:])
        ref = np.einsum("ijk->ki", dots * dots)
        got = S[:2, :, d - 1].cpu().numpy()
        errs.append(float(np.max(np.abs(got - ref) / np.abs(ref))))
    out["c2_spot_rel_err"] = errs
    print("c2 spot rel errs", errs)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ct_check.json", "w") as fp:
        json.dump(out, fp, indent=1)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("total s", time.time() - t0)
