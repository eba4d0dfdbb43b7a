"""Micro-benchmark helper (not a test): time the bf16 minibatch kernels at the headline size.

    MAVA_TC_DEBUG=<stage> python tests/tc_bench.py      # stage 0 = full kernel

Stages cut the fused kernel's tile loop short (see ppo_tc.cu), so successive stages give a
cumulative time profile of the loop without a profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mava_b200 import native
from mava_b200._lib import PpoHyper

DEV = "cuda:0"
A, FR, N, U, E, T = 4, 66, 5, 2, 1024, 128
mb = T * E // 2
rng = np.random.default_rng(0)
S = T * U * E
view = torch.from_numpy(rng.integers(0, 12, size=(S, A, FR)).astype(np.int8)).to(DEV)
mask = torch.full((S, A), 31, dtype=torch.uint8, device=DEV)
action = torch.zeros(S, A, dtype=torch.int8, device=DEV)
actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
ap = (torch.randn(na, device=DEV) * 0.05)
cp = (torch.randn(nc, device=DEV) * 0.05)
ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
native.mlp_pack_bf16(actor, ap, ai)
native.mlp_pack_bf16(critic, cp, ci)
z = lambda *s: torch.randn(*s, device=DEV)
rows = torch.randperm(T * U * E, device=DEV)[: U * mb].to(torch.int32)
g = torch.zeros(na + nc + 8, device=DEV)
ws = torch.zeros(native.ppo_workspace_bytes_bf16(actor, critic, U * mb), dtype=torch.uint8, device=DEV)
args = (actor, ap, ai, critic, cp, ci, PpoHyper(0.2, 0.01, 0.5), view, mask, action, z(S, A), z(S, A),
        z(S, A), z(S, A), rows, U, mb, g, ws)
for _ in range(3):
    native.ppo_loss_grad_bf16(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    native.ppo_loss_grad_bf16(*args)
e1.record()
torch.cuda.synchronize()
print(f"stage {os.environ.get('MAVA_TC_DEBUG', '0')}: {e0.elapsed_time(e1) / n:.3f} ms per minibatch")
