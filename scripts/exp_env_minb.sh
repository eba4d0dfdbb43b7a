#!/bin/bash
# experiment: step-kernel register budget (resident CTAs per SM) vs throughput at 2^20 envs
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
for m in "$@"; do
  nvcc $FLAGS -DMAVA_RWARE_MINB=$m -Xptxas -v -c mava_b200/csrc/env_rware.cu -o mava_b200/build/env_rware.o 2>&1 | grep -A2 "step_kernelILi4ELi1" | grep -E "registers|spill" | tr '\n' ' '
  nvcc -shared -o mava_b200/libmava_b200.so mava_b200/build/*.o -lcudart
  echo "MINB=$m"
  python sweep_rollout.py --scenario small-4ag --min-log2 20 --max-log2 20 --steps 200 --warmup 20
  python sweep_rollout.py --scenario tiny-4ag --min-log2 20 --max-log2 20 --steps 200 --warmup 20
done
