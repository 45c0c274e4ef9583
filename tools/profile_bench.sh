#!/bin/bash
# Run on the GPU box: plain bench first, then the ncu launch list and one full capture of the top kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_plain_$TAG.err; exit 1; }
cat gpurun_out/bench_plain_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ct_lag -s 3 -c 1 -f -o gpurun_out/prof_ctlag_$TAG \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full ct_lag rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sphere_hist -s 3 -c 1 -f -o gpurun_out/prof_hist_$TAG \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_hist_$TAG.log 2>&1
echo "ncu full hist rc=$?"
ls -la gpurun_out
