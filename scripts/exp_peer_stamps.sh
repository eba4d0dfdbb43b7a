#!/bin/bash
# development aid (GPU box, N ranks): phase times of reduce_clip_adam_kernel (globaltimer stamps)
cd "$(dirname "$0")/.."
N=${1:-2}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
nvcc $FLAGS -DMAVA_PEER_STAMPS -c mava_b200/csrc/peer.cu -o mava_b200/build/peer.o || exit 1
nvcc -shared -o mava_b200/libmava_b200.so mava_b200/build/*.o -lcudart || exit 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_reduce_adam.py
