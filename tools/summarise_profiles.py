"""Condense .ncu-rep captures (gpurun_out/) into a small JSON for profiles/ (the .ncu-rep files are scratch).

    python tools/summarise_profiles.py OUT.json name=path.ncu-rep [name=path.ncu-rep ...]
"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__inst_executed_op_shared_atom.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:160]}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                try:
                    d[h] = [float(v), u]
                except ValueError:
                    d[h] = [v, u]
        kernels.append(d)
    return kernels


if __name__ == "__main__":
    res = {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        res[name] = summarise(path)
    with open(sys.argv[1], "w") as fp:
        json.dump(res, fp, indent=1)
    print("wrote", sys.argv[1], {k: len(v) for k, v in res.items()})
