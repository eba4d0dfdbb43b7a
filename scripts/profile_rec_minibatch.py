#!/usr/bin/env python
"""One recurrent PPO minibatch (mava_rec_ppo_loss_grad) at the rec_mappo_smax bench shapes on
random buffers -- the unit to run under `ncu --metrics gpu__time_duration.sum` for a launch list."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mava_b200 import native  # noqa: E402
from mava_b200._lib import PpoHyper  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
E, U, A, F, G, N, H = 2048, 2, 8, 205, 168, 13, 128
NE, chunk, nc, mbc = U * E, T, 1, E // 2
ad = native.rnn_desc(native.IN_DENSE, False, A, 1, H, H, N, F, A, prec)
cd = native.rnn_desc(native.IN_DENSE, False, A, 1, H, H, 1, G, 1, prec)
na, ncr = native.rnn_param_count(ad), native.rnn_param_count(cd)
ap = torch.randn(na, device=dev) * 0.05
cp = torch.randn(ncr, device=dev) * 0.05
oa = torch.rand(T, NE, A, F, device=dev)
oc = torch.rand(T, NE, 1, G, device=dev)
mask = torch.full((T, NE, A), 0x1FFF, dtype=torch.uint16, device=dev)
action = torch.randint(0, N, (T, NE, A), device=dev, dtype=torch.int8)
old_logp = -torch.rand(T, NE, A, device=dev) - 1.0
old_value, adv, targets = (torch.randn(T, NE, A, device=dev) for _ in range(3))
done = (torch.rand(T, NE, device=dev) < 0.01).to(torch.uint8)
hs_a = torch.zeros(nc, NE * A, H, device=dev)
hs_c = torch.zeros(nc, NE, H, device=dev)
cols = torch.randperm(E * nc, device=dev)[:mbc].to(torch.int32)
grad = torch.zeros(na + ncr + 8, device=dev)
ws = torch.zeros(native.rec_ppo_workspace_bytes(ad, cd, U * mbc, chunk), dtype=torch.uint8, device=dev)
hyper = PpoHyper(0.2, 0.01, 0.5)
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    native.rec_ppo_loss_grad(ad, ap, cd, cp, hyper, None, oa, oc, mask, action, old_logp, old_value,
                             adv, targets, done, hs_a, hs_c, cols, U, E, mbc, chunk, nc, grad, ws)
    torch.cuda.synchronize()
    print(f"iter {it}: {1e3 * (time.perf_counter() - t0):.2f} ms, loss5 {grad[-8:-3].tolist()}")
