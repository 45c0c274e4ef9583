"""Launch only the K4 reduction over all lag windows (for ncu captures).  Scratch tool."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import _lib, synth  # noqa: E402

lib = _lib.load()
N = int(os.environ.get("DQ_N", "1000000"))
nl = int(os.environ.get("DQ_LAGS", "100000"))
q = synth.quaternion_walk(N, seed=synth.BASE_SEED + 3, sigma=(0.004, 0.006, 0.012))
qd = torch.from_numpy(q).cuda()
lags = np.arange(1, nl + 1, dtype=np.int64)
ld = torch.from_numpy(lags).cuda()
M = torch.empty((nl, 4, 6), dtype=torch.float64, device="cuda")
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.sr_dq_moments(qd.data_ptr(), N, ld.data_ptr(), nl, 1, 4, M.data_ptr(), None))
    b.record()
    torch.cuda.synchronize()
    print("ms", a.elapsed_time(b))
