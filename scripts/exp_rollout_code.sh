#!/bin/bash
# development aid (run on the GPU box): the headline bench with the cold code of the rollout kernel's
# step loop emitted in different ways (MAVA_ROLL_GEN x MAVA_ROLL_COLD, see csrc/rollout_tc.cu)
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
if [ $# -eq 0 ]; then set -- "0 1" "1 0" "1 1" "3 1" "2 1"; fi
for cfg in "$@"; do
  set -- $cfg
  nvcc $FLAGS -DMAVA_ROLL_GEN=$1 -DMAVA_ROLL_COLD=$2 $MAVA_EXTRA_FLAGS -c mava_b200/csrc/rollout_tc.cu -o mava_b200/build/rollout_tc.o || exit 1
  nvcc -shared -o mava_b200/libmava_b200.so mava_b200/build/*.o -lcudart || exit 1
  for i in 1 2; do
    python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GEN=$1 COLD=$2', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],4), 'ms')"
  done
done
