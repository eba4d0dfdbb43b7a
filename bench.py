#!/usr/bin/env python
"""Headline benchmark: RWARE ff_mappo env-steps/sec (rollout + GAE + PPO) per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): ff_mappo on RWARE tiny-4ag, 2048 envs per GPU
(update_batch_size 2 x num_envs 1024), 128-step rollouts, 4 PPO epochs x 2 minibatches.  A "step" is
one full update = 262 144 env-steps per GPU.  N > 1 is launched by torchrun, one rank per GPU
(weak scaling: per-GPU work fixed); the timed region is K graph replays bracketed by barrier +
synchronize, timed with CUDA events, max over ranks.  Prints ONE JSON line on rank 0.

`e2e` is the same metric through the public `learn(state)` call with host buffers: every step copies
the learner inputs (parameters, Adam moments, key: one allocation, one copy) from pinned host memory,
runs `learn` -- which ends with the device -> host copy of what the run loop consumes (sort-overflow
flag, finished-episode statistics reduced on the device, minibatch losses: one 256-byte block) and
its only host wait -- and copies the updated parameters / moments / key back.
`roofline` times the PPO minibatch kernels (fused forward + backward, critic first-layer gradient)
with CUDA events around each call (median over the timed launches); `cpu_baseline` / `--impl reference` time the oracle
port on the host cores (the reference itself cannot run in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RWARE ff_mappo env-steps/sec (rollout+GAE+PPO)"
UNIT = "env-steps/s"
TASK = dict(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
            request_queue_size=4)
OVERRIDES = ["env=rware", "env/scenario=tiny-4ag", "arch.num_envs=1024",
             "system.update_batch_size=2", "system.rollout_length=128", "system.ppo_epochs=4",
             "system.num_minibatches=2", "logger.use_console=False"]


_COMMON = ["system.rollout_length=128", "system.ppo_epochs=4", "system.num_minibatches=2",
           "system.update_batch_size=2", "logger.use_console=False"]
# The default is the headline (BASELINE.json configs[1]); the others are the remaining single-GPU
# configurations of BASELINE.json, selectable with --workload (their lines go to profiles/).
WORKLOADS = {
    "ff_mappo_rware": dict(
        system="ff_mappo", metric=METRIC, envs_per_gpu=2048,
        overrides=OVERRIDES,
        text="ff_mappo RWARE tiny-4ag, 2048 envs/GPU (update_batch_size 2 x num_envs 1024), "
             "rollout 128, 4 epochs x 2 minibatches (BASELINE.json configs[1])"),
    "ff_ippo_lbf": dict(
        system="ff_ippo", metric="LBF ff_ippo env-steps/sec (rollout+GAE+PPO)", envs_per_gpu=65536,
        overrides=["env=lbf", "env/scenario=8x8-2p-2f-coop", "arch.num_envs=32768"] + _COMMON,
        text="ff_ippo Level-Based Foraging 8x8-2p-2f-coop, 65536 envs/GPU (update_batch_size 2 x "
             "num_envs 32768), rollout 128, 4 epochs x 2 minibatches (BASELINE.json configs[2])"),
    "rec_mappo_smax": dict(
        system="rec_mappo", metric="SMAX-shaped rec_mappo env-steps/sec (rollout+GAE+PPO)",
        envs_per_gpu=4096,
        overrides=["env=smax_synthetic", "arch.num_envs=2048"] + _COMMON,
        text="rec_mappo (GRU 128) on synthetic SMAX 3s5z-shaped tensors (8 agents, obs 205, state "
             "168, 13 actions), 4096 envs/GPU (update_batch_size 2 x num_envs 2048), rollout 128, "
             "4 epochs x 2 minibatches (BASELINE.json configs[3]; no SMAX dynamics, see DESIGN.md)"),
    "rec_mappo_rware": dict(
        system="rec_mappo", metric="RWARE rec_mappo env-steps/sec (rollout+GAE+PPO)",
        envs_per_gpu=2048,
        overrides=["env=rware", "env/scenario=tiny-4ag", "arch.num_envs=1024"] + _COMMON,
        text="rec_mappo (GRU 128) RWARE tiny-4ag, 2048 envs/GPU (update_batch_size 2 x num_envs "
             "1024), rollout 128, 4 epochs x 2 minibatches"),
}


def workload_config(n_gpus: int, name: str = "ff_mappo_rware") -> dict:
    w = WORKLOADS[name]
    return {"workload": w["text"], "envs_per_gpu": w["envs_per_gpu"],
            "global_envs": w["envs_per_gpu"] * n_gpus, "rollout_length": 128,
            "parallelism": f"dp{n_gpus}", "l2": "inputs are rewritten by every update; "
            "a 256 MiB buffer is written between timed steps to flush L2"}


# ------------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    """The reference's CPU implementation of the path, timed on host cores.

    The reference itself (JAX_PLATFORMS=cpu) cannot run in this image: jax/jumanji are not
    installed and there is no network, so this times the oracle port
    (oracle/cpu_baseline.py) with every host core, as the tier contract prescribes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb

    sample_envs = WORKLOADS["ff_mappo_rware"]["envs_per_gpu"]  # one GPU's full share: 2048 envs
    res = cb.run(TASK, num_envs=sample_envs, updates=max(1, args.steps), warmup=min(args.warmup, 1))
    value = res["env_steps_per_s"]
    sample = (f"{max(1, args.steps)} update(s) of {sample_envs} envs x 128 steps, 4 epochs x 2 "
              "minibatches = one GPU's whole share of the workload per step (oracle port: C/OpenMP "
              "env + torch-CPU networks, every host core)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * res["seconds"] / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event reasons of one GPU, sampled through NVML every `period_ms` (a thread
    in this process: nvidia-smi's own loop cannot go below ~100 ms, which is longer than the timed
    region of the headline).  start() before the warm-up, mark() at the start of the timed region,
    stop() after it: `sm_mhz` is the median over the timed region (`samples` of them), reasons are
    collected over warm-up + timed region."""

    def __init__(self, gpu_index: int, period_ms: float = 4.0):
        self.gpu, self.period = gpu_index, period_ms * 1e-3
        self.rows, self._stop, self._t_mark, self._thread, self._h = [], False, None, None, None
        self._nv = None

    def _handle(self):
        import pynvml as nv

        nv.nvmlInit()
        self._nv = nv
        try:
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() \
                else self.gpu
            return nv.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            self._h = self._handle()
            self._max = float(self._nv.nvmlDeviceGetMaxClockInfo(self._h, self._nv.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self._h, self._err = None, repr(e)
            return self
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def _loop(self):
        nv, h = self._nv, self._h
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.rows.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        self._t_mark = time.perf_counter()

    def stop(self) -> dict:
        if self._h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0,
                    "reasons": ["nvml unavailable: " + getattr(self, "_err", "?")]}
        time.sleep(2 * self.period)
        self._stop = True
        self._thread.join(timeout=1.0)
        nv = self._nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        timed = [r for r in self.rows if self._t_mark is None or r[0] >= self._t_mark]
        use = timed or self.rows
        sm = sorted(r[1] for r in use)
        reasons = sorted({n for r in self.rows for bit, n in names.items() if r[2] & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                "sm_max_mhz": self._max, "reasons": reasons, "samples": len(timed),
                "samples_incl_warmup": len(self.rows), "period_ms": self.period * 1e3,
                "source": "nvml"}


def loss_grad_flops(L) -> float:
    """Executed FLOPs of one ppo_loss_grad call (forward + backward of both networks)."""
    if hasattr(L, "chunk"):  # recurrent: pre, x-gates, h-gates, post, head; backward = 2 x forward
        def rnet(d, rows):
            H, Q = d.hidden, d.post
            return 3.0 * rows * 2.0 * (d.in_dim * H + 2 * H * 3 * H + H * Q + Q * d.out_dim)
        S = L.U * L.mbc * L.chunk
        return rnet(L.actor_desc, S * L.A) + rnet(L.critic_desc, S * L.critic_desc.rows_per_env)

    def net(d, rows):
        fwd = 2.0 * (d.in_dim * d.h1 + d.h1 * d.h2 + d.h2 * d.out_dim)
        bwd = 2.0 * (d.h2 * d.out_dim + d.h1 * d.h2)          # dZ2, dZ1
        wg = 2.0 * (d.in_dim * d.h1 + d.h1 * d.h2 + d.h2 * d.out_dim)
        return rows * (fwd + bwd + wg)
    R = L.U * L.mb
    crows = R if L.critic_desc.input_mode == 1 else R * L.A
    return net(L.actor_desc, R * L.A) + net(L.critic_desc, crows)


def _finish(world: int, device) -> None:
    """Leave a multi-rank run: the update graphs hold no NCCL collective (the gradient mean runs
    inside the optimiser kernel over peer-mapped buffers), so the group tears down normally."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize(device)
    sys.stdout.flush()
    sys.stderr.flush()
    dist.barrier()
    dist.destroy_process_group()


def _peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


COLLECTIVE = "peer"  # --collective: peer (fused NVLink all-reduce, default) | nccl (the checker)


def _setup(name: str, precision: str, device):
    """Learner of one BASELINE.json workload on this rank's GPU (synthetic data, seeded init)."""
    import importlib

    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.utils import make_env

    wl = WORKLOADS[name]
    system = importlib.import_module(f"mava_b200.systems.ppo.{wl['system']}")
    cfg = compose(system.CONFIG_NAME, wl["overrides"] + [f"+arch.precision={precision}",
                                                         f"+arch.collective={COLLECTIVE}"])
    env, _ = make_env.make(cfg, add_global_state=system.CENTRALISED_CRITIC, device=device)
    key, _, ak, ck = prng.split(prng.PRNGKey(cfg.system.seed), 4)
    learn, _, state = system.learner_setup(env, (key, ak, ck), cfg)
    cfg.system.num_updates_per_eval = 1
    return learn, state, learn.learner


def _time_updates(L, steps: int, warmup: int, flush, barrier, learn=None, state=None):
    """W untimed + K timed updates (graph replays), CUDA events per update; returns total ms."""
    import torch

    for _ in range(warmup):
        learn(state) if learn is not None else L.learn(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    barrier()
    for k in range(steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed steps
        ev[k][0].record()
        L.learn(1)
        ev[k][1].record()
    barrier()
    return sum(a.elapsed_time(b) for a, b in ev)


def other_workloads(args, device, world, flush, barrier) -> dict:
    """The other single-GPU-share configurations of BASELINE.json (configs[2], configs[3]) in the same
    run: value (whole job, env-steps/s) and ms per update, device-timed like the headline."""
    import gc

    import torch
    import torch.distributed as dist

    out = {}
    for name in ("ff_ippo_lbf", "rec_mappo_smax"):
        learn, state, L = _setup(name, args.precision, device)
        steps = max(2, min(args.steps, 5))
        ms = _time_updates(L, steps, 3, flush, barrier, learn, state)
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[name] = {"metric": WORKLOADS[name]["metric"],
                     "value": world * L.T * L.NE * steps / (ms * 1e-3), "unit": UNIT,
                     "ms_per_step": ms / steps, "steps": steps, "warmup": 3, "n_gpus": world,
                     "dtype": L.compute_dtype, "envs_per_gpu": L.NE,
                     "workload": WORKLOADS[name]["text"]}
        if hasattr(L, "release"):
            L.release()
        del learn, state, L
        gc.collect()
        torch.cuda.empty_cache()
    return out


def roofline_extra(device) -> list:
    """north_star's HBM-bound kernels against the measured copy bandwidth, each with its own NVML
    clock record over warm-up + timed region: the env-step kernels at 2^20 envs (state + outputs =
    0.6-0.9 GB per step, far beyond the 126 MB L2, so every step streams from HBM) and GAE at 64 M
    elements (1.1 GB).  achieved = algorithmic bytes / CUDA-event time."""
    import torch

    from mava_b200 import native, prng
    from sweep_rollout import SCENARIOS

    peak = float(_peaks().get("hbm_gbs", 6548.8))
    dev_index = torch.cuda.current_device()
    rows = []
    E = 1 << 20
    for scen, kernel in (("small-4ag", "rware_step_kernel"), ("tiny-4ag", "rware_step_kernel"),
                         ("lbf-8x8-2p-2f-coop", "lbf_step_kernel")):
        sc = dict(SCENARIOS[scen])
        kind = sc.pop("kind")
        env = native.Env.rware(**sc) if kind == "rware" else native.Env.lbf(**sc)
        A, FR, N = env.num_agents, env.view_dim, env.num_actions
        keys = torch.from_numpy(prng.split(prng.PRNGKey(0), E).copy()).to(device)
        state = env.alloc_state(E, device)
        view = torch.zeros(E, A, FR, dtype=torch.int8, device=device)
        mask = torch.zeros(E, A, dtype=torch.uint8, device=device)
        reward = torch.zeros(E, A, device=device)
        done = torch.zeros(E, dtype=torch.uint8, device=device)
        ep_ret = torch.zeros(E, device=device)
        ep_len = torch.zeros(E, dtype=torch.int32, device=device)
        env.reset(keys, state, view, mask, E)
        g = torch.Generator(device=device).manual_seed(0)
        acts = [torch.randint(0, N, (E, A), generator=g, device=device, dtype=torch.int8)
                for _ in range(8)]
        sampler = ClockSampler(dev_index).start()
        for i in range(50):
            env.step(state, acts[i % 8], view, mask, reward, done, ep_ret, ep_len, E, True)
        torch.cuda.synchronize(device)
        steps = 300
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.mark()
        e0.record()
        for i in range(steps):
            env.step(state, acts[i % 8], view, mask, reward, done, ep_ret, ep_len, E, True)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / steps
        algo = int(env.dims.algo_bytes_per_step) * E
        gbs = algo / (ms * 1e-3) / 1e9
        rows.append({"kernel": kernel, "workload": f"{scen}, 2^20 envs, random actions, auto-reset",
                     "bound": "hbm", "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
                     "algo_bytes": algo, "algo_bytes_per_env_step": int(env.dims.algo_bytes_per_step),
                     "ms": ms, "env_steps_per_s": E / (ms * 1e-3), "launches_timed": steps,
                     "clocks": sampler.stop()})
        del state, view, mask, reward, done, ep_ret, ep_len, acts, keys
        torch.cuda.empty_cache()
    # GAE: 64 Mi elements (T = 128 x 2^19 env-agents), 17 B per element
    T, A, n = 128, 4, 1 << 19
    NE = n // A
    reward = torch.randn(T, NE, A, device=device)
    value = torch.randn(T, NE, A, device=device)
    done = (torch.rand(T, NE, device=device) < 0.01).to(torch.uint8)
    last_val = torch.randn(NE, A, device=device)
    adv, tgt = torch.empty_like(reward), torch.empty_like(reward)
    sampler = ClockSampler(dev_index).start()
    for _ in range(5):
        native.gae(reward, value, done, last_val, 0.99, 0.95, T, NE, A, adv, tgt)
    torch.cuda.synchronize(device)
    iters = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    e0.record()
    for _ in range(iters):
        native.gae(reward, value, done, last_val, 0.99, 0.95, T, NE, A, adv, tgt)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / iters
    algo = 16 * T * n + T * NE + 4 * n
    gbs = algo / (ms * 1e-3) / 1e9
    rows.append({"kernel": "gae_kernel", "workload": "T=128 x 2^19 env-agents = 64 Mi elements",
                 "bound": "hbm", "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
                 "algo_bytes": algo, "ms": ms, "launches_timed": iters, "clocks": sampler.stop()})
    return rows


def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from mava_b200.systems.ppo import _runner
    from mava_b200.systems.ppo.anakin import episode_summary

    device = _runner.init_distributed()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    wl = WORKLOADS[args.workload]
    learn, state, L = _setup(args.workload, args.precision, device)
    steps_per_update = L.T * L.NE  # per GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    sampler = ClockSampler(torch.cuda.current_device()).start()  # covers warm-up + timed region
    for _ in range(max(args.warmup, 3)):
        learn(state)
    torch.cuda.synchronize(device)
    launches_per_update = L.launches_per_update

    # ---- device-timed region: K updates, inputs resident in HBM -------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    sampler.mark()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed steps
        ev[k][0].record()
        L.learn(1)
        ev[k][1].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop()

    # ---- end to end through the public API: learn(state) + host copies each step ---------------
    blob = getattr(L, "state_blob", None)  # params | Adam moments | key in one allocation
    if blob is not None:
        host_blob = torch.empty_like(blob, device="cpu").pin_memory()
        host_blob.copy_(blob)
        h2d = blob.numel() * 4
        d2h = L.report.numel()  # episode statistics + sort flag + minibatch losses, read by learn()

        def e2e_step():
            # host -> device: the replicated learner inputs (params, optimiser moments, key)
            blob.copy_(host_blob, non_blocking=True)
            out = learn(state)  # ends with the device -> host copy of the report block + a sync
            assert out.train_metrics["entropy"].shape[-2:] == L.loss_host.shape[:2]
            # device -> host: the learner state back; what run_experiment logs is already there
            host_blob.copy_(blob, non_blocking=True)
            episode_summary(L)  # from the report block of this call: no further copy
            torch.cuda.synchronize(device)
    else:
        host_params = torch.empty_like(L.params, device="cpu").pin_memory()
        host_opt = torch.empty(2 * L.params.numel(), dtype=torch.float32).pin_memory()
        host_key = torch.empty(2, dtype=torch.uint32).pin_memory()
        host_params.copy_(L.params)
        host_opt[: L.params.numel()].copy_(L.mu)
        host_opt[L.params.numel():].copy_(L.nu)
        host_key.copy_(L.key)
        pin_loss = torch.empty(4, 1, L.epochs, L.nmb).pin_memory()
        h2d = (host_params.numel() + host_opt.numel()) * 4 + 8
        d2h = 10 * 8 + pin_loss.numel() * 4  # finished-episode statistics (device reduction) + losses

        def e2e_step():
            # host -> device: the replicated learner inputs (params, optimiser moments, key)
            L.params.copy_(host_params, non_blocking=True)
            L.mu.copy_(host_opt[: L.params.numel()], non_blocking=True)
            L.nu.copy_(host_opt[L.params.numel():], non_blocking=True)
            L.key.copy_(host_key, non_blocking=True)
            out = learn(state)
            # device -> host: what run_experiment consumes (episode statistics + train metrics), then
            # the parameters back
            for i, k2 in enumerate(("total_loss", "value_loss", "actor_loss", "entropy")):
                pin_loss[i].copy_(out.train_metrics[k2], non_blocking=True)
            host_params.copy_(L.params, non_blocking=True)
            host_opt[: L.params.numel()].copy_(L.mu, non_blocking=True)
            host_opt[L.params.numel():].copy_(L.nu, non_blocking=True)
            host_key.copy_(L.key, non_blocking=True)
            episode_summary(L)  # 80 bytes, synchronises the stream
            torch.cuda.synchronize(device)

    for _ in range(3):  # warm-up of the host path (first pinned copies, allocator, metric code)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    d2h += h2d

    # ---- dominant kernel: ppo_loss_grad, CUDA events around each call over K eager updates ------
    L.use_graph_saved, L._graph_saved = L.use_graph, L._graph
    L.use_graph, L._graph = False, None
    L.time_loss_grad = []
    L.time_reduce_apply = []
    roof_steps = min(args.steps, 3)
    for _ in range(roof_steps):
        L.learn(1)
    torch.cuda.synchronize(device)
    lg_ms = [a.elapsed_time(b) for a, b in L.time_loss_grad]
    ra_ms = [a.elapsed_time(b) for a, b in L.time_reduce_apply]
    L.time_loss_grad = None
    L.time_reduce_apply = None
    L.use_graph, L._graph = L.use_graph_saved, L._graph_saved

    # ---- reduce over ranks (max time) -----------------------------------------------------------
    # (median over the timed launches: an eager launch that waits for the side stream's row lists,
    #  or at N > 1 for a peer whose host is behind, is not the kernel's duration)
    t = torch.tensor([dev_ms, e2e_s * 1000.0, float(np.median(lg_ms)),
                      float(np.median(ra_ms)) if ra_ms else 0.0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, lg_mean_ms, ra_mean_ms = (float(x) for x in t.tolist())

    extras = not args.no_extras and args.workload == "ff_mappo_rware"
    flops = loss_grad_flops(L)
    dominant, dtype_name = L.dominant_kernel, L.compute_dtype
    collective = {"peer": "one-shot all-reduce over peer-mapped buffers (NVLink) inside the "
                          "optimiser kernel (csrc/peer.cu)",
                  "nccl": "nccl all_reduce + optimiser kernel", "none": "none (1 rank)"}[L.collective]
    import gc

    L.release()  # drop the graph and the peer mappings before the next learner / the teardown
    del learn, state, L
    gc.collect()
    torch.cuda.empty_cache()
    workloads = other_workloads(args, device, world, flush, barrier) if extras else None
    if rank != 0:
        _finish(world, device)
        return
    extra = roofline_extra(device) if extras and world == 1 else None

    peaks = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[args.workload]
        traffic = int(tr["dram_read_bytes"]) + int(tr["dram_write_bytes"])
    except Exception:
        pass
    achieved = flops / (lg_mean_ms * 1e-3) / 1e12
    value = world * steps_per_update * args.steps / (dev_ms * 1e-3)
    e2e_value = world * steps_per_update * args.steps / (e2e_ms * 1e-3)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "ff_mappo_rware":  # N = 1 only
        from oracle import cpu_baseline as cb

        envs = wl["envs_per_gpu"]
        res = cb.run(TASK, num_envs=envs, updates=1, warmup=0)
        cpu_baseline = {"value": res["env_steps_per_s"], "unit": UNIT, "cores": res["cores"],
                        "kind": "port",
                        "sample": f"1 update of {envs} envs x 128 steps, 4 epochs x 2 minibatches "
                                  "(one GPU's whole share), oracle port on all host cores"}

    line = {"metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype_name, "data": "synthetic",
            "config": workload_config(world, args.workload),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_per_update * args.steps),
            "roofline": {"kernel": dominant, "bound": "tensor", "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "peak_source": peak_src,
                         "flops_per_launch": flops, "ms_per_launch": lg_mean_ms,
                         "launches_timed": len(lg_ms), "statistic": "median over the timed launches"},
            "collective": {"kind": collective, "reduce_clip_adam_us": ra_mean_ms * 1e3,
                           "per_update": len(ra_ms) // max(1, roof_steps),
                           "note": "gradient mean over ranks + clip + Adam + bf16 repack per "
                                   "minibatch, CUDA events, eager launches, median over the launches, max "
                                   "over ranks"},
            "cpu_baseline": cpu_baseline}
    if extra is not None:
        line["roofline_extra"] = extra
    if workloads is not None:
        line["workloads"] = workloads
    print(json.dumps(line), flush=True)
    _finish(world, device)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip roofline_extra and the other BASELINE.json workloads")
    ap.add_argument("--workload", default="ff_mappo_rware", choices=sorted(WORKLOADS))
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="gradient mean over ranks: fused peer-memory all-reduce, or NCCL (checker)")
    args = ap.parse_args()
    global COLLECTIVE
    COLLECTIVE = args.collective
    if os.environ.get("MAVA_BENCH_DEBUG"):  # dump every thread's stack if the run gets stuck
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["MAVA_BENCH_DEBUG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
