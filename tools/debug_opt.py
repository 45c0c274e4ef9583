"""Scratch: run the --opt modes and print CSA / chi against tests/golden/relax_opt.npz."""
import contextlib, io, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spinrelax_b200 import cli_relax, hist
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
g, gc, r = np.load(G + "/relax_opt.npz"), np.load(G + "/relax_cli.npz"), np.load(G + "/relax.npz")
td = tempfile.mkdtemp()
open(td + "/x_fittedCt.dat", "w").write(str(gc["fitted"]))
hist.save_vec_histogram(td + "/h_vecHistogram.npz", np.arange(6), r["hist"].astype(np.float64), [r["edges_phi"], r["edges_cos"]])
files = []
for f in (600, 800):
    for t in ("R1", "R2", "NOE"):
        fn = td + "/e_%s_%d.dat" % (t, f)
        open(fn, "w").write(str(g["expt_%s_%d" % (t, f)]))
        files.append(fn)
for mode, opt, extra in (("rsCSA", "rsCSA", []), ("mixed", "Diso,rsCSA", ["--cycles", "4"])):
    for local in ("powell", "batched"):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            cli_relax.main(["-f", td + "/x_fittedCt.dat", "--distfn", td + "/h_vecHistogram.npz", "-D", "2.1e-5", "--aniso", "1.35",
                            "-o", td + "/o", "--opt", opt, "--localopt", local] + extra + files)
        chi = [l for l in buf.getvalue().splitlines() if "Final chi-value" in l][-1]
        csa = np.loadtxt(td + "/o_CSA_opt.dat")[:, 1]
        print(mode, local, chi, "ref chi", float(g["chi_" + mode]))
        print("   csa ours", csa)
        print("   csa ref ", g["csa_" + mode][:, 1])
        print("   csa true", g["csa_true"])
        for f in (600, 800):
            for t in ("R1", "R2", "NOE"):
                rows = lambda txt: np.array([[float(x) for x in l.split()[1:]] for l in txt.splitlines() if l and l[0] not in "#@&"])
                a, b = rows(open(td + "/o_15N1H_%dMHz_%s.xvg" % (f, t)).read()), rows(str(g["xvg_%s_%s_%d" % (mode, t, f)]))
                print("   ", t, f, "max rel val", np.max(np.abs(a[:6, 0] - b[:6, 0]) / np.abs(b[:6, 0])), "max abs sig", np.max(np.abs(a[:6, 1] - b[:6, 1])), "sig", b[:2, 1])
