#!/bin/bash
# Run on the GPU box: tests, plain bench, ncu launch list of the same command, full captures of the small kernels.
# (The full capture of ct_lag_kernel at config-2 size costs ~12 GPU-minutes; it is taken separately on a slice.)
set -u
mkdir -p gpurun_out
TAG=${1:-r01c}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench_secondary.py > gpurun_out/bench_secondary_$TAG.jsonl 2> gpurun_out/bench_secondary_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sphere_hist_kernel -s 3 -c 1 -f -o gpurun_out/prof_hist_$TAG \
    python tools/run_hist_only.py > gpurun_out/ncu_full_hist_$TAG.log 2>&1
echo "ncu full hist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 8 -c 1 -f -o gpurun_out/prof_pack_$TAG \
    python tools/gpu_check_ct.py > gpurun_out/ncu_full_pack_$TAG.log 2>&1
echo "ncu full pack rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dq_moments -s 1 -c 1 -f -o gpurun_out/prof_dq_$TAG \
    python bench_secondary.py --quick > gpurun_out/ncu_full_dq_$TAG.log 2>&1
echo "ncu full dq rc=$?"
TUNE_VARIANTS=17 TUNE_NR=8 ncu --set full --clock-control none --import-source on -k regex:ct_lag -c 1 -f -o gpurun_out/prof_ctlag_slice_$TAG \
    python tools/tune_ct.py > gpurun_out/ncu_full_ctlag_slice_$TAG.log 2>&1
echo "ncu full ct_lag slice rc=$?"
ls gpurun_out | grep $TAG
