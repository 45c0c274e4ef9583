#!/bin/bash
# Round-2 GPU profiling session (run on the GPU box via gpurun): plain bench first, then the ncu launch list of the same
# command, then --set full captures of the kernels the round worked on.  Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ct_lag -s 3 -c 1 -f -o gpurun_out/prof_ctlag_full_$TAG \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_ctlag_$TAG.log 2>&1
echo "ncu full ct_lag (config-2 size) rc=$?"
DQ_LAGS=20000 ncu --set full --clock-control none --import-source on -k regex:dq_moments_consec -c 1 -f -o gpurun_out/prof_dq_consec_$TAG \
    python tools/run_dq_only.py > gpurun_out/ncu_full_dq_$TAG.log 2>&1
echo "ncu full dq rc=$?"
FIT_N=300 ncu --set full --clock-control none --import-source on -k regex:ct_fit_trf -s 4 -c 1 -f -o gpurun_out/prof_fit9_$TAG \
    python tools/run_fit_only.py > gpurun_out/ncu_full_fit_$TAG.log 2>&1
echo "ncu full fit rc=$?"
ls -la gpurun_out | grep $TAG
