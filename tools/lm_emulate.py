"""Scratch: NumPy emulation of ct_fit_lm_kernel's iteration (csrc/fit.cu) to count LM iterations per residue on the
bench_secondary curves -- tells whether the launch time is set by stragglers or by the typical residue."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_secondary import synth_curves  # noqa: E402
from oracle import fit_oracle  # noqa: E402


def evaluate(t, y, sg, q, nP):
    nc = nP // 2
    free = nP % 2 == 1
    C, tau = q[:nc], q[nc:2 * nc]
    S2 = q[-1] if free else 1.0 - C.sum()
    w = 1.0 / sg
    e = np.exp(-t[None] / tau[:, None])
    f = S2 + (C[:, None] * e).sum(0)
    g = np.zeros((nP, len(t)))
    g[:nc] = (e - (0.0 if free else 1.0)) * w
    g[nc:2 * nc] = C[:, None] * e * t[None] / (tau[:, None] ** 2) * w
    if free:
        g[-1] = w
    r = (f - y) * w
    return g @ g.T, g @ r, 0.5 * r @ r


def lm(t, y, sg, p0, lo, hi, max_iter=2000, ftol=1e-14):
    nP = len(p0)
    nc = nP // 2
    lo, hi = lo.copy(), hi.copy()
    lo[nc:2 * nc] = np.maximum(lo[nc:2 * nc], 1e-12 * hi[nc:2 * nc])
    eps = 1e-10 * (hi - lo)
    p = np.minimum(np.maximum(p0, lo + eps), hi - eps)
    A, g, c = evaluate(t, y, sg, p, nP)
    lam, small, evals = 1e-3, 0, 1
    for it in range(max_iter):
        span = hi - lo
        fixed = ((p - lo <= 1e-12 * span) & (g > 0)) | ((hi - p <= 1e-12 * span) & (g < 0))
        idx = np.where(~fixed)[0]
        d = np.zeros(nP)
        ok = True
        if len(idx):
            M = A[np.ix_(idx, idx)].copy()
            dg = np.diag(M).copy()
            M[np.diag_indices_from(M)] += lam * np.where(dg > 0, dg, 1.0)
            try:
                Lc = np.linalg.cholesky(M)
                d[idx] = -np.linalg.solve(Lc.T, np.linalg.solve(Lc, g[idx]))
            except np.linalg.LinAlgError:
                ok = False
        if not ok:
            lam *= 10
            if lam > 1e20:
                return p, it, evals, 3
            continue
        q = p + d
        trunc = False
        below, above = q < lo, q > hi
        if below.any() or above.any():
            trunc = True
            q = np.where(below, p - 0.995 * (p - lo), np.where(above, p + 0.995 * (hi - p), q))
        smax = np.max(np.abs(q - p) / (np.abs(p) + 1e-300))
        if smax < 1e-15 and not trunc:
            return p, it, evals, 2
        A1, g1, c1 = evaluate(t, y, sg, q, nP)
        evals += 1
        if c1 <= c and c1 == c1:
            dec = c - c1
            p, A, g, c0, c = q, A1, g1, c, c1
            lam = max(lam * 0.3, 1e-12)
            small = small + 1 if (dec <= ftol * c0 and not trunc) else 0
            if small >= 3 and lam <= 1e-7:
                return p, it + 1, evals, 1
        else:
            lam *= 4
            if lam > 1e20:
                return p, it, evals, 3
    return p, max_iter, evals, 0


if __name__ == "__main__":
    n = int(os.environ.get("N", "150"))
    t, Y, SG = synth_curves(1000, 500, 20260105)
    for nP in (2, 3, 5, 7, 9):
        its, sts = [], []
        for i in range(n):
            p0 = np.array(fit_oracle.initial_guess(t, Y[i], nP)[0])
            lo = np.zeros(nP)
            hi = np.array(fit_oracle.bounds(nP, t[-1] * 10)[1], dtype=float)
            with np.errstate(all="ignore"):
                p, it, ev, st = lm(t, Y[i], SG[i], p0, lo, hi)
            its.append(ev); sts.append(st)
        its = np.array(its)
        print("nP", nP, "evals mean %.0f median %.0f p90 %.0f max %d" % (its.mean(), np.median(its), np.percentile(its, 90), its.max()),
              "status counts", {s: sts.count(s) for s in set(sts)}, flush=True)
