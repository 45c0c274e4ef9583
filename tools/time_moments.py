"""Scratch: time sr_vec_block_moments (S2 / vecAvg reductions) at config-2 size."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import _lib, synth
lib = _lib.load()
nF, nR = 1000000, 76
v = torch.from_numpy(synth.nh_vectors(nF, nR, seed=3)).cuda()
for per in (nF, 200000, 500):
    nB = -(-nF // per)
    out = torch.empty((nB, nR, 9), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.check(lib.sr_vec_block_moments(v.data_ptr(), nF, nR, per, out.data_ptr(), None))
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    ref = v[:per].double()
    chk = torch.stack((ref[..., 0].sum(0), (ref[..., 0] * ref[..., 1]).sum(0)), -1)
    err = float((out[0, :, [0, 4]] - chk).abs().max())
    print("framesPerBlock", per, "ms", best, "GB/s", nF * nR * 12 / best / 1e6, "err", err)
