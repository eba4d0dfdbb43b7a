#!/bin/bash
# development aid: per-phase clock64() stamps of ppo_fused_kernel (CTA 0, tiles 2..17)
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
nvcc $FLAGS -DMAVA_PROFILE_PHASES $MAVA_STAMP_FLAGS -c mava_b200/csrc/ppo_tc.cu -o mava_b200/build/ppo_tc.o
nvcc $FLAGS -DMAVA_PROFILE_PHASES -c mava_b200/csrc/rollout_tc.cu -o mava_b200/build/rollout_tc.o
nvcc -shared -o mava_b200/libmava_b200.so mava_b200/build/*.o -lcudart
python - <<'PY'
import ctypes, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mava_b200 import prng, _lib
from mava_b200.config import compose
from mava_b200.systems.ppo import ff_mappo
from mava_b200.utils import make_env
import bench
cfg = compose(ff_mappo.CONFIG_NAME, bench.OVERRIDES + ["+arch.use_cuda_graph=False"])
env, _ = make_env.make(cfg, add_global_state=True)
key, _, ak, ck = prng.split(prng.PRNGKey(42), 4)
learn, _, state = ff_mappo.learner_setup(env, (key, ak, ck), cfg)
cfg.system.num_updates_per_eval = 1
learn(state); learn(state)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 256)()
lib.mava_debug_phases.restype = ctypes.c_int
assert lib.mava_debug_phases(buf) == 0
a = np.array(buf[:]).reshape(16, 16)
names = ["start", "x_built", "gemm1_issued", "gemm1_done", "epi1", "epi1_sync", "epi2_sync(gemm2+epi2)",
         "gemm3_done", "loss_sync", "dh2_done", "dz2_sync", "dh1_done", "dz1+dw1_issued"]
d = np.diff(a[:, :13], axis=1)
print("cycles per phase, mean over 16 tiles (actor CTA 0):")
for n, v in zip(names[1:], d.mean(0)):
    print(f"  {n:28s} {v:9.0f}")
print("tile total (start->start):", np.diff(a[:, 0]).mean())
print("x_built -> gemm1 issued:", (a[:, 13] - a[:, 1]).mean(), " -> prefetch issued:", (a[:, 14] - a[:, 13]).mean(),
      " -> loss inputs requested:", (a[:, 2] - a[:, 14]).mean())
buf2 = (ctypes.c_longlong * 160)()
lib.mava_debug_phases2.restype = ctypes.c_int
assert lib.mava_debug_phases2(buf2) == 0
b = np.array(buf2[:128]).reshape(16, 8)
print("idle-warp branch of the loss phase (warp 4): wait+bar %.0f, expand %.0f, fence+bar %.0f, issue copies %.0f" %
      tuple(np.diff(b[:, :5], axis=1).mean(0)))
print("critic build (with -DMAVA_STAMP_CRITIC): wait dW2 -> %s" % np.diff(np.concatenate([a[:, :1], b[:, 3:8]], axis=1), axis=1).mean(0))
c = np.array(buf2[128:144])
print("critic tile (last CTA), start->start: %.0f cycles" % np.diff(c[2:]).mean())
print("gemm3_done -> dz3 stored:", (a[:, 15] - a[:, 7]).mean(), " -> db3 reduced + sync:", (a[:, 8] - a[:, 15]).mean())
wb = (ctypes.c_longlong * 96)()
lib.mava_debug_wg1.restype = ctypes.c_int
assert lib.mava_debug_wg1(wb) == 0
w = np.array(wb[:]).reshape(8, 12)
wn = ["wait rows", "tail+sync", "expand", "fence+sync", "request rows", "wait dZ1", "sync+MMA issue", "wait MMA"]
print("ppo_wgrad1_kernel, cycles per phase (CTA 0, tiles 0..6):")
for i in range(7):
    print("  tile %d: " % i + ", ".join("%s %d" % (n_, v) for n_, v in zip(wn, np.diff(w[i, :9]))) + ", total %d" % (w[i, 8] - w[i, 0]))
print("  flush (TMEM -> atomics): %d cycles" % (w[7, 11] - w[7, 10]))
rb = (ctypes.c_longlong * 256)()
lib.mava_debug_rollout_phases.restype = ctypes.c_int
assert lib.mava_debug_rollout_phases(rb) == 0
ra = np.array(rb[:]).reshape(16, 16)
rn = ["gemm1", "epi1+sync", "gemm2", "epi2+sync", "gemm3", "head", "env step", "reset + rows + X row", "sync"]
print("rollout kernel, cycles per phase of a step (CTA 0, steps 16..31):")
for n_, v in zip(rn, np.diff(ra[:, :10], axis=1).mean(0)):
    print(f"  {n_:22s} {v:8.0f}")
print(f"  {'regeneration + store':22s} {(ra[:, 12] - ra[:, 9]).mean():8.0f}")
print("  inside 'reset + rows + X row': reset %.0f, wait for the quadrant's env steps %.0f, quarter row %.0f, wait for the other quarters %.0f, expand %.0f" %
      ((ra[:, 13] - ra[:, 7]).mean(), (ra[:, 14] - ra[:, 13]).mean(), (ra[:, 15] - ra[:, 14]).mean(),
       (ra[:, 11] - ra[:, 15]).mean(), (ra[:, 8] - ra[:, 11]).mean()))
print("  per step: total", np.diff(ra[:, 0]), " regeneration + store", ra[:, 12] - ra[:, 9])
print("  step total (start->start):", np.diff(ra[:, 0]).mean())
PY
