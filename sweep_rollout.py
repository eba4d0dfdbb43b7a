#!/usr/bin/env python
"""Rollout-only sweep (BASELINE.json configs[4]): the fused env-step kernel against the HBM roofline.

    python sweep_rollout.py [--scenario small-4ag] [--min-log2 10] [--max-log2 20] [--steps 1000]

For E = 2^10 .. 2^20 envs: reset, 100 warm-up steps, then `steps` timed env steps (CUDA events on the
launching stream).  Actions come from a fixed device buffer of uniform random actions (seed 0);
illegal ones become no-ops inside the kernel, exactly like the reference's `get_valid_actions`.
Reports ns/step, env-steps/s and achieved GB/s = algorithmic bytes per env-step (SURVEY.md 8d,
`mava_env_dims.algo_bytes_per_step`) x E / time, against MEASURED_PEAKS.json `hbm_gbs`.
Below ~2^19 envs the state fits the 126 MB L2, so those rows are L2-resident / latency-bound.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENARIOS = {
    "tiny-2ag": dict(kind="rware", num_agents=2, request_queue_size=2, shelf_rows=1),
    "tiny-4ag": dict(kind="rware", num_agents=4, request_queue_size=4, shelf_rows=1),
    "small-4ag": dict(kind="rware", num_agents=4, request_queue_size=4, shelf_rows=2),
    "lbf-8x8-2p-2f-coop": dict(kind="lbf", grid_size=8, fov=8, num_agents=2, num_food=2),
}


def main() -> None:
    import numpy as np
    import torch

    from mava_b200 import native, prng

    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="small-4ag", choices=sorted(SCENARIOS))
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=20)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sc = dict(SCENARIOS[args.scenario])
    kind = sc.pop("kind")
    env = native.Env.rware(**sc) if kind == "rware" else native.Env.lbf(**sc)
    A, FR, N = env.num_agents, env.view_dim, env.num_actions
    algo = env.dims.algo_bytes_per_step
    peak = 6548.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        E = 1 << lg
        keys = torch.from_numpy(prng.split(prng.PRNGKey(0), E).copy()).to(dev)
        state = env.alloc_state(E, dev)
        view = torch.zeros(E, A, FR, dtype=torch.int8, device=dev)
        mask = torch.zeros(E, A, dtype=torch.uint8, device=dev)
        reward = torch.zeros(E, A, device=dev)
        done = torch.zeros(E, dtype=torch.uint8, device=dev)
        ep_ret = torch.zeros(E, device=dev)
        ep_len = torch.zeros(E, dtype=torch.int32, device=dev)
        env.reset(keys, state, view, mask, E)
        g = torch.Generator(device=dev).manual_seed(0)
        nbuf = 8  # a few action buffers so consecutive steps differ
        acts = [torch.randint(0, N, (E, A), generator=g, device=dev, dtype=torch.int8)
                for _ in range(nbuf)]
        for i in range(args.warmup):
            env.step(state, acts[i % nbuf], view, mask, reward, done, ep_ret, ep_len, E, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            env.step(state, acts[i % nbuf], view, mask, reward, done, ep_ret, ep_len, E, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ns_step = ms * 1e6 / args.steps
        sps = E * args.steps / (ms * 1e-3)
        gbs = sps * algo / 1e9
        row = {"scenario": args.scenario, "envs": E, "ns_per_step": ns_step, "env_steps_per_s": sps,
               "achieved_gbs": gbs, "hbm_peak_gbs": peak, "frac": gbs / peak,
               "algo_bytes_per_env_step": algo,
               "regime": "HBM" if E * env.state_stride > 126e6 else "L2-resident / latency-bound"}
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()
