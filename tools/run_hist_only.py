"""Scratch: launch only the K3 histogram kernel at the config-2 size (for ncu captures)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import hist, synth  # noqa: E402

n = int(os.environ.get("HIST_FRAMES", "1000000"))
nR = int(os.environ.get("HIST_NR", "76"))
v = torch.from_numpy(synth.nh_vectors(n, nR, seed=3)).cuda()
acc = hist.SphereHistogram(nR)
q = np.array([0.83, -0.31, 0.22, 0.41])
for _ in range(3):
    acc.accumulate_device(v, q)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); acc.accumulate_device(v, q); b.record(); torch.cuda.synchronize()
print("hist ms", a.elapsed_time(b), "GB/s(12B)", n * nR * 12 / a.elapsed_time(b) / 1e6, "ambiguous", int(acc.amb_count.item()))
