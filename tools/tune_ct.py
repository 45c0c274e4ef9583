"""Time the experimental K1 tile configurations (sr_ct_lag_sums_variant of the -DSR_TUNING build, tools/build_tune.py)
on a config-2-shaped slice and check accuracy against the FFT oracle.  Scratch tool, not product."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ct_oracle  # noqa: E402  (scratch tool, not product)
from spinrelax_b200 import synth  # noqa: E402

lib = ctypes.CDLL(os.path.join(HERE, "libct_tune.so"))
ll, i, vp = ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p
lib.sr_ct_row_pitch.restype = ll
lib.sr_ct_row_pitch.argtypes = [ll]
lib.sr_pack_vectors_f32.argtypes = [vp, i, ll, i, vp, vp, ll, vp]
lib.sr_ct_lag_sums_variant.argtypes = [vp, ll, i, ll, i, ll, vp, i, vp]
lib.sr_last_error.restype = ctypes.c_char_p


def check(rc):
    if rc:
        raise RuntimeError(lib.sr_last_error().decode())


nC, nF, nR = 5, 200000, int(os.environ.get("TUNE_NR", "8"))
L = nF // 2
v = synth.nh_vectors(nC * nF, nR, seed=17).reshape(nC, nF, nR, 3)
vt = torch.from_numpy(v).cuda()
pitch = lib.sr_ct_row_pitch(nF)
packed = torch.empty((nR, nC, 3, pitch), dtype=torch.float32, device="cuda")
S = torch.empty((nR, nC, L), dtype=torch.float64, device="cuda")
check(lib.sr_pack_vectors_f32(vt.data_ptr(), nC, nF, nR, None, packed.data_ptr(), pitch, None))
So = ct_oracle.ct_lag_sums_fft(v[:, :, :2])
pairs = nR * nC * (L * nF - L * (L + 1) // 2)
res = {}
variants = [int(x) for x in os.environ.get("TUNE_VARIANTS", ",".join(str(k) for k in range(20))).split(",")]
for var in variants:
    best = 1e30
    for it in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib.sr_ct_lag_sums_variant(packed.data_ptr(), pitch, nC, nF, nR, L, S.data_ptr(), var, None))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    err = float(np.max(np.abs(S[:2].cpu().numpy() - So) / So))
    res[var] = {"ms": best, "pairs_per_s": pairs / best * 1e3, "tflops7": pairs * 7 / best * 1e-9,
                "frac_of_74.45": pairs * 7 / best * 1e-9 / 74.45, "max_rel_err_S": err}
    print(var, res[var], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/%s.json" % os.environ.get("TUNE_TAG", "tune_ct"), "w"), indent=1)
