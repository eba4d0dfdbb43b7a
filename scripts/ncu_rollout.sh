#!/bin/bash
# One `ncu --set full` capture of rware_rollout_kernel at the headline shape (under gpurun, ONE GPU).
# The bench command runs plain first; the .ncu-rep lands in gpurun_out/ and is read with
#   ncu -i gpurun_out/prof_rollout.ncu-rep --page source --csv --print-source sass
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:rware_rollout_kernel -s 6 -c 1 \
    -o gpurun_out/prof_rollout -f $B > gpurun_out/ncu_rollout.log 2>&1
ls -la gpurun_out/prof_rollout.ncu-rep
