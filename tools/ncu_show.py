"""Print selected metrics of an .ncu-rep (first kernel) -- usage: python tools/ncu_show.py REP [substr ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "issue_active.avg.pct", "pipe_fma", "issue_stalled", "warps_active.avg.pct",
                        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct", "registers_per_thread ",
                        "inst_executed.sum ", "pipe_lsu.avg.pct", "pipe_xu.avg.pct", "pipe_fp64.avg.pct", "pipe_alu.avg.pct"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, r):
        if any(w.strip() in h and (not w.endswith(" ") or h.endswith(w.strip())) for w in want):
            print("  %-90s %-12s %s" % (h, u, v))
