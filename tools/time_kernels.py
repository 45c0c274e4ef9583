"""Scratch timing of the secondary kernels (K3 histogram, K4 dq moments, K5 fits, K6 relax) at BASELINE sizes."""
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import _lib, dq, hist, synth  # noqa: E402

out = {}


def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


# K3: 1e6 frames x 76 vectors
g = torch.Generator(device="cuda").manual_seed(1)
v = torch.randn((1000000, 76, 3), device="cuda", generator=g)
v = (v / v.norm(dim=-1, keepdim=True)).contiguous()
acc = hist.SphereHistogram(76)
q = np.array([0.83, -0.31, 0.22, 0.41])
ms = ev_time(lambda: acc.accumulate_device(v, q))
out["hist_rot_ms"] = ms; out["hist_rot_GBps_alg12"] = 76e6 * 12 / ms / 1e6
ms = ev_time(lambda: acc.accumulate_device(v, None))
out["hist_f32_ms"] = ms; out["hist_f32_ambiguous"] = int(acc.amb_count.item())
vs = torch.from_numpy(synth.nh_vectors(200000, 76, seed=3)).cuda()
acc2 = hist.SphereHistogram(76)
ms = ev_time(lambda: acc2.accumulate_device(vs, q))
out["hist_rot_synth200k_ms"] = ms; out["hist_rot_synth200k_GBps"] = 200000 * 76 * 12 / ms / 1e6
del v
# K4: 1e6 quaternions; run-all lag set (100 lags) and a dense block of 2000 lags
lib = _lib.load()
qq = torch.from_numpy(synth.quaternion_walk(1000000, seed=5)).cuda()
for name, lags in (("runall100", np.arange(1000, 100001, 1000)), ("dense2000", np.arange(1, 2001)),
                   ("dense2000_far", np.arange(98001, 100001))):
    ld = torch.from_numpy(lags.astype(np.int64)).cuda()
    M = torch.empty((len(lags), 4, 6), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.check(lib.sr_dq_moments(qq.data_ptr(), 1000000, ld.data_ptr(), len(lags), int(lags.min()), 4,
                                              M.data_ptr(), None))
    ms = ev_time(fn, 3)
    pairs = float(np.sum(1000000 - lags))
    out["dq_%s_ms" % name] = ms; out["dq_%s_pairs_per_s" % name] = pairs / ms * 1e3
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_kernels.json", "w"), indent=1)
