"""Debug helper (not a test): run the fused tensor-core minibatch kernel up to a given stage.
   MAVA_TC_DEBUG=<stage> python tests/tc_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mava_b200 import native
from mava_b200._lib import PpoHyper
from tests.test_mlp_gpu import flat, make_params, random_batch
DEV = "cuda:0"
A, FR, N, U, mb, T, E = 4, 66, 5, 2, 96, 12, 40
rng = np.random.default_rng(1)
S = T * U * E
view, mask_bool, mask = random_batch(rng, S, A, FR, N)
actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
ap = torch.from_numpy(flat(make_params(rng, actor.in_dim, 128, 128, N, scale_out=0.05))).to(DEV)
cp = torch.from_numpy(flat(make_params(rng, critic.in_dim, 128, 128, 1, scale_out=0.05))).to(DEV)
action = np.zeros((S, A), np.int8)
z = lambda *s: torch.zeros(*s, device=DEV)
rows = torch.arange(U * mb, dtype=torch.int32, device=DEV)
ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
native.mlp_pack_bf16(actor, ap, ai); native.mlp_pack_bf16(critic, cp, ci)
na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
g = z(na + nc + 8)
ws = torch.zeros(native.ppo_workspace_bytes_bf16(actor, critic, U * mb), dtype=torch.uint8, device=DEV)
dev = lambda x: torch.from_numpy(x).to(DEV)
native.ppo_loss_grad_bf16(actor, ap, ai, critic, cp, ci, PpoHyper(0.2, 0.01, 0.5), dev(view), dev(mask),
                          dev(action), z(S, A), z(S, A), z(S, A), z(S, A), rows, U, mb, g, ws)
torch.cuda.synchronize()
print("stage", os.environ.get("MAVA_TC_DEBUG"), "OK", float(g.abs().sum()))
