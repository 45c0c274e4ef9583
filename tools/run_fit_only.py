"""Launch the K5 ladder once on the config-5 curves (for ncu captures / timing).  Scratch tool."""
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench_secondary import synth_curves  # noqa: E402
from spinrelax_b200 import fitct  # noqa: E402

n = int(os.environ.get("FIT_N", "1000"))
t, Y, SG = synth_curves(n, 500, 77)
for rep in range(2):
    ac = fitct.autoCorrelations()
    ac.import_target_array([str(i) for i in range(n)], [t] * n, Y, SG)
    fitct.KERNEL_EVENTS = []
    t0 = time.perf_counter()
    ac.fit_all_residues(fp=io.StringIO())
    torch.cuda.synchronize()
    print("e2e ms", 1e3 * (time.perf_counter() - t0), "kernels ms", [round(a.elapsed_time(b), 2) for a, b in fitct.KERNEL_EVENTS])
