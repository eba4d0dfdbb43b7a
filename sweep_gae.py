#!/usr/bin/env python
"""GAE reverse-scan kernel against the HBM roofline.

    python sweep_gae.py [--T 128] [--min-log2 17] [--max-log2 23]

For NE*A = 2^k env-agents and T time steps: times `mava_gae` (CUDA events on the launching
stream, L2 flushed by the working set itself: every size here is larger than the 126 MB L2) and
reports achieved GB/s = 17 B per env-agent-step (read reward 4, value 4, done 1 per env; write
advantage 4, target 4) x elements / time against MEASURED_PEAKS.json `hbm_gbs`.  Both flavours
(ff: done of the transition; rec: next_done carry) are timed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main() -> None:
    import torch

    from mava_b200 import native

    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--A", type=int, default=4)
    ap.add_argument("--min-log2", type=int, default=17)
    ap.add_argument("--max-log2", type=int, default=23)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = 6548.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    T, A = args.T, args.A
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = 1 << lg  # env-agents
        NE = n // A
        reward = torch.randn(T, NE, A, device=dev)
        value = torch.randn(T, NE, A, device=dev)
        done = (torch.rand(T, NE, device=dev) < 0.01).to(torch.uint8)
        last_val = torch.randn(NE, A, device=dev)
        last_done = torch.zeros(NE, dtype=torch.uint8, device=dev)
        adv, tgt = torch.empty_like(reward), torch.empty_like(reward)
        for rec in (False, True):
            kw = dict(last_done=last_done) if rec else {}
            for _ in range(3):
                native.gae(reward, value, done, last_val, 0.99, 0.95, T, NE, A, adv, tgt, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                native.gae(reward, value, done, last_val, 0.99, 0.95, T, NE, A, adv, tgt, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            elems = T * n
            algo = 16 * elems + T * NE + 4 * n
            gbs = algo / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": "gae_kernel<rec>" if rec else "gae_kernel<ff>", "T": T,
                              "env_agents": n, "elements": elems, "ms": ms, "achieved_gbs": gbs,
                              "hbm_peak_gbs": peak, "frac": gbs / peak,
                              "algo_bytes": algo}), flush=True)
        del reward, value, done, adv, tgt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
