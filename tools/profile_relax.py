"""Scratch: where does relax_grid spend its time?"""
import cProfile, io, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench_secondary import synth_curves
from oracle import ct_oracle
from spinrelax_b200 import fitct, specdens as sd, synth
nR = 1000
t, Y, SG = synth_curves(nR, 500, 77)
ac = fitct.autoCorrelations()
ac.import_target_array([str(i) for i in range(nR)], [t] * nR, Y, SG)
ac.fit_all_residues(fp=io.StringIO())
q = np.array([0.8, -0.36, 0.48, 0.0])
v = synth.nh_vectors(2000, nR, seed=5)
hist, edges = ct_oracle.sphere_histogram(v, q)
vecs, w = sd.convert_LambertCylindricalHist_to_vecs(hist, edges)
rot = sd.globalRotationalDiffusion_Axisymmetric(D=[2.1e-5, 1.35])
rot.set_frame_vectors(np.arange(nR), vecs, w)
ac.set_zeta(0.890023)
fields = [500.0, 600.133, 700.0, 800.0, 950.0]
csa = np.linspace(-220e-6, -120e-6, 64)
sd.relax_grid(rot, ac, fields, csa)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
rot._amom = None
sd.relax_grid(rot, ac, fields, csa)
torch.cuda.synchronize()
print("total s", time.perf_counter() - t0)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(16); print(s.getvalue()[:3000])
