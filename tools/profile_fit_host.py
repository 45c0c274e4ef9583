"""Scratch: time the host half of fit_all_residues at config-5 size with the device solve replaced by cached
solutions of the NumPy emulation (tools/lm_emulate.py) -- shows where the non-kernel time of the fits goes."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_secondary import synth_curves  # noqa: E402
from lm_emulate import evaluate, lm  # noqa: E402
from spinrelax_b200 import fitct  # noqa: E402

nR = int(os.environ.get("N", "1000"))
t, Y, SG = synth_curves(nR, 500, 20260105)
cache = {}
solve_s = [0.0]


def fake_solve(tt, y, sigma, p0, lo, hi):
    t0 = time.perf_counter()
    n, nP = p0.shape
    lo, hi = np.broadcast_to(lo, p0.shape), np.broadcast_to(hi, p0.shape)
    popt, JtJ, cost = np.zeros((n, nP)), np.zeros((n, nP, nP)), np.zeros(n)
    for i in range(n):
        key = (nP, y[i].tobytes()[:64])
        if key not in cache:
            with np.errstate(all="ignore"):
                p, it, ev, st = lm(tt[i], y[i], sigma[i], p0[i], lo[i].copy(), hi[i].copy(), max_iter=400)
                A, g, c = evaluate(tt[i], y[i], sigma[i], p, nP)
            cache[key] = (p, A, c)
        popt[i], JtJ[i], cost[i] = cache[key]
    solve_s[0] += time.perf_counter() - t0
    return popt, JtJ, cost, np.ones((n, 2), dtype=np.int32)


fitct._device_solve = fake_solve
names = [str(i) for i in range(nR)]


def run():
    ac = fitct.autoCorrelations()
    ac.import_target_array(names, [t] * nR, Y, SG)
    ac.fit_all_residues(fp=io.StringIO())
    return ac


run()                       # fills the cache
for rep in range(2):
    solve_s[0] = 0.0
    t0 = time.perf_counter()
    ac = run()
    tot = time.perf_counter() - t0
    print("host ladder %.1f ms (total %.1f ms - cached solve lookups %.1f ms)" % ((tot - solve_s[0]) * 1e3, tot * 1e3, solve_s[0] * 1e3))
pr = cProfile.Profile()
pr.enable(); run(); pr.disable()
st = pstats.Stats(pr, stream=sys.stdout).sort_stats("cumulative")
st.print_stats(28)
