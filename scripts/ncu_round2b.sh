#!/bin/bash
# Round-2 evidence run, final build (under gpurun, ONE GPU): launch list of the headline bench and
# `ncu --set full` captures of the kernels that changed since scripts/ncu_round2.sh ran.  Every
# command runs plain first (exit 0) and only then under ncu.  Outputs land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline"
FULL="ncu --set full --clock-control none --import-source on"
$B > gpurun_out/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_r2b.csv $B > gpurun_out/ncu_launches.log 2>&1
for k in ppo_fused_kernel rware_rollout_kernel reduce_clip_adam_kernel; do
  $FULL -k regex:$k -s 6 -c 1 -o gpurun_out/prof_r2b_$k -f $B > gpurun_out/ncu_$k.log 2>&1
done
S="python sweep_rollout.py --scenario small-4ag --min-log2 20 --max-log2 20 --steps 40 --warmup 10"
$S > gpurun_out/plain_sweep.log 2>&1 && \
  $FULL -k regex:rware_step_kernel -s 20 -c 1 -o gpurun_out/prof_r2b_rware_step_kernel -f $S > gpurun_out/ncu_rware_step.log 2>&1
ls -la gpurun_out/*r2b*.ncu-rep
