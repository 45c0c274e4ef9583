import io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import fitct
g = np.load("tests/golden/fit.npz")
t, Ct, dCt = g["t"], g["Ct"], g["dCt"]
k = 0
for i in range(len(Ct)):
    for npar in (2, 3, 5):
        r = g["single"][k]; k += 1
        m = fitct.autoCorrelationModel(name=i)
        m.set_nParams(npar)
        m.initialise_for_fit_advanced(t, Ct[i])
        lo, hi = m.get_bounds_as_list(tauMax=t[-1] * 10)
        p0 = np.array(m.get_params_as_list(), dtype=float)
        popt, pcov, cost, status = fitct.gpu_curve_fit(t, Ct[i], dCt[i], p0[None], np.zeros_like(p0)[None], np.array(hi, float)[None])
        nc = npar // 2
        C, tau = popt[0][:nc], popt[0][nc:2 * nc]
        S2 = popt[0][-1] if npar % 2 else 1 - C.sum()
        model = S2 + np.sum(C[:, None] * np.exp(-t[None] / tau[:, None]), axis=0)
        chi = np.mean((model - Ct[i]) ** 2 / dCt[i])
        print(i, npar, "chi ours %.6e ref %.6e ratio %.5f" % (chi, r[1], chi / r[1]), "cost", cost[0], "status", status[0],
              "popt", np.array2string(popt[0], precision=5), "ref C", r[6:6 + nc], "tau", r[9:9 + nc])
