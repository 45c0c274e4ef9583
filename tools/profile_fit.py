"""Scratch: where does fit_all_residues spend its time (host vs GPU)?"""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench_secondary import synth_curves
from spinrelax_b200 import fitct
nR = 1000
t, Y, SG = synth_curves(nR, 500, 77)
ac = fitct.autoCorrelations()
ac.import_target_array([str(i) for i in range(nR)], [t] * nR, Y, SG)
ac.fit_all_residues(fp=io.StringIO())
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
ac.fit_all_residues(fp=io.StringIO())
torch.cuda.synchronize()
print("total s", time.perf_counter() - t0)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:3500])
