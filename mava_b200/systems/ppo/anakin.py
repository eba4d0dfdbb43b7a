"""The Anakin PPO learner shared by ff_ippo / ff_mappo: rollout -> GAE -> PPO epochs.

Mirrors ``get_learner_fn`` / ``learner_setup`` of mava/systems/ppo/ff_mappo.py:45-432 (ff_ippo.py is
the same file with a decentralised critic).  The JAX program
``pmap(scan[updates](vmap[update_batch_size](_update_step)))`` becomes, per GPU (one process per GPU):

* one stream of hand-written CUDA kernels per update (include/mava_b200.h), with no host sync,
  captured once into a CUDA graph and replayed ``num_updates_per_eval`` times;
* the ``update_batch_size`` replicas live side by side on the env axis (NE = U * E envs), share
  the parameters and the PRNG key like the reference (ff_mappo.py:417-426), normalise advantages
  per replica and average their gradients inside the loss kernel (pmean "batch", :224-226);
* ``pmean("device")`` (:228-238) is ONE NCCL all-reduce per minibatch of the contiguous buffer
  [actor grads | critic grads | 5 loss scalars]; 1/world_size is folded into the Adam kernel.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ... import native, prng
from ..._lib import PpoHyper
from ...peer import PeerGroup
from ...networks import FeedForwardActor, FeedForwardValueNet
from ...types import (ExperimentOutput, LearnerState, OptStates, Params, StepType, TimeStep)
from ...wrappers import EnvState, NativeMarlEnv


def world() -> Tuple[int, int]:
    """(rank, world_size) of the data-parallel group (1 process per GPU)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _nccl_allreduce(grad: torch.Tensor) -> None:
    """pmean("device") as a sum over ranks (1/world is folded into the optimiser kernel)."""
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)


def _u32(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32)).to(device)


class FFLearner:
    """Device buffers + the kernel schedule of one GPU's share of the Anakin learner."""

    def __init__(self, env: NativeMarlEnv, actor: FeedForwardActor, critic: FeedForwardValueNet,
                 config, centralised_critic: bool, device: torch.device,
                 rank_world: Optional[Tuple[int, int]] = None):
        s = config.system
        self.env, self.config, self.device = env, config, device
        # rank_world overrides the process group: a test can build the learners of several ranks in
        # one process and do pmean("device") by hand (tests/test_multirank_gpu.py)
        self.rank, self.world = rank_world if rank_world is not None else world()
        self.allreduce = _nccl_allreduce if rank_world is None else None
        # pmean("device"): "peer" = inside the optimiser kernel over peer-mapped buffers (csrc/peer.cu),
        # "nccl" = dist.all_reduce before it (kept as the checker of the peer path)
        self.collective = str(config.arch.get("collective", "peer"))
        if self.collective not in ("peer", "nccl"):
            raise ValueError(f"arch.collective must be peer or nccl, got {self.collective}")
        if self.world == 1 or rank_world is not None:
            self.collective = "none"
        self.T, self.U, self.E = int(s.rollout_length), int(s.update_batch_size), int(
            config.arch.num_envs)
        self.NE = self.U * self.E
        self.A, self.FR, self.N = env.num_agents, env.native.view_dim, env.action_dim
        self.epochs, self.nmb = int(s.ppo_epochs), int(s.num_minibatches)
        if (self.T * self.E) % self.nmb != 0:
            raise ValueError("rollout_length * num_envs must be divisible by num_minibatches")
        self.mb = self.T * self.E // self.nmb
        add_id = bool(s.add_agent_id)
        self.actor_desc = actor.desc(self.A, self.FR, add_id, native.IN_AGENT_VIEW)
        self.critic_desc = critic.desc(
            self.A, self.FR, add_id, native.IN_GLOBAL if centralised_critic else native.IN_AGENT_VIEW)
        self.na = native.mlp_param_count(self.actor_desc)
        self.nc = native.mlp_param_count(self.critic_desc)
        self.hyper = PpoHyper(float(s.clip_eps), float(s.ent_coef), float(s.vf_coef))
        self.use_graph = bool(config.arch.get("use_cuda_graph", True))
        self.precision = str(config.arch.get("precision", "auto"))

        dev, T, NE, A = device, self.T, self.NE, self.A
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        # learner state
        # parameters | Adam moments | key live in ONE allocation: a host that keeps the learner state
        # moves it with one copy each way (bench.py's end-to-end leg)
        n_par = self.na + self.nc
        n_pad = (n_par + 3) // 4 * 4  # the optimiser kernel wants 16-byte aligned vectors
        self.state_blob = z(3 * n_pad + 4)
        self.params = self.state_blob[:n_par]
        self.mu = self.state_blob[n_pad:n_pad + n_par]
        self.nu = self.state_blob[2 * n_pad:2 * n_pad + n_par]
        self.key = self.state_blob[3 * n_pad:3 * n_pad + 2].view(torch.uint32)
        self.counts = z(2, dtype=torch.int32)
        self.env_buf = env.native.alloc_state(NE, dev)
        # rollout buffers (slot t holds the observation step t acts on; slot T the bootstrap obs)
        self.view = z(T + 1, NE, A, self.FR, dtype=torch.int8)
        self.mask = z(T + 1, NE, A, dtype=torch.uint8)
        self.action = z(T, NE, A, dtype=torch.int8)
        self.logp, self.reward = z(T, NE, A), z(T, NE, A)
        self.value_all = z(T + 1, NE, A)  # slot T is the bootstrap value
        self.value, self.last_val = self.value_all[:T], self.value_all[T]
        self.done = z(T, NE, dtype=torch.uint8)
        self.ep_ret = z(T, NE)
        self.ep_len = z(T, NE, dtype=torch.int32)
        self.adv, self.targets = z(T, NE, A), z(T, NE, A)
        # finished-episode statistics of the current learn() call, reduced on the device
        # ... in one small block with everything else the host reads after a learn() call (sort
        # overflow flag, the minibatch losses): one device -> host copy, one synchronisation
        n_loss = int(s.ppo_epochs) * int(s.num_minibatches) * 5
        self.report = z(96 + 4 * n_loss, dtype=torch.uint8)
        self.ep_stats = self.report[:80].view(torch.float64)
        self._report_host = None
        self._report_fresh = False
        self._stats_host = None
        # scratch
        self.policy_keys = z(T, 2, dtype=torch.uint32)
        self.key3 = z(3, 2, dtype=torch.uint32)
        self.key3_ep = z(int(s.ppo_epochs), 3, 2, dtype=torch.uint32)
        self._side = torch.cuda.Stream(device=dev)
        n_perm = T * self.E
        self.perm_rounds = int(math.ceil(3 * math.log(max(1, n_perm)) / math.log(2 ** 32 - 1)))
        self.arange_n = torch.arange(n_perm, dtype=torch.int32, device=dev)
        self.perm_buf = z(int(s.ppo_epochs), 2, n_perm, dtype=torch.int32)
        self.key2_r = z(int(s.ppo_epochs), max(1, self.perm_rounds), 2, 2, dtype=torch.uint32)
        self.sort_ws = z(native.sort_workspace_bytes(n_perm), dtype=torch.uint8)
        self.sort_overflow = self.report[80:84].view(torch.int32)
        self.key2 = z(2, 2, dtype=torch.uint32)
        self.bits = z(T * self.E, dtype=torch.uint32)
        self.rows = z(self.U * self.mb, dtype=torch.int32)
        # bf16 path: row lists and advantage statistics of ALL minibatches of an update, prepared off
        # the critical path (side stream) -- they depend on the permutations and on GAE only
        self.rows_all = z(int(s.ppo_epochs), self.nmb, self.U * self.mb, dtype=torch.int32)
        self.adv_stats_all = z(int(s.ppo_epochs), self.nmb, 16, dtype=torch.float64)
        # gradients live in this rank's exchange buffer [actor | critic | 8 loss scalars]; with
        # collective == "peer" the other ranks read it over NVLink inside the optimiser kernel
        n_grad = self.na + self.nc + 8
        self.peer = PeerGroup(n_grad, dev, self.rank, self.world) if self.collective == "peer" \
            else PeerGroup(n_grad, dev, 0, 1)
        self.peer_local = self.peer if self.peer.world == 1 else \
            PeerGroup(n_grad, dev, 0, 1, local_bufs=[self.peer.struct.buf[self.rank]])
        self.grad = self.peer.grad
        self.gsum = z(self.na + self.nc)
        self.loss_buf = self.report[96:].view(torch.float32).view(self.epochs, self.nmb, 5)
        # precision: the bf16 tensor-core kernels need two hidden layers of width 128
        tc_ok = all(d.h1 == 128 and d.h2 == 128 and d.out_dim <= 16
                    for d in (self.actor_desc, self.critic_desc))
        if self.precision not in ("auto", "fp32", "bf16"):
            raise ValueError(f"arch.precision must be auto, fp32 or bf16, got {self.precision}")
        if self.precision == "bf16" and not tc_ok:
            raise ValueError("arch.precision=bf16 needs MLP torsos with layer_sizes [128, 128]")
        self.bf16 = tc_ok and self.precision != "fp32"
        # one persistent kernel for the whole rollout scan where the env kernel supports it
        self.fused_rollout = (self.bf16 and native.ff_rollout_supported(env.native)
                              and bool(config.arch.get("fused_rollout", True)))
        if self.bf16:
            self.actor_img = z(native.mlp_pack_bytes(self.actor_desc), dtype=torch.uint8)
            self.critic_img = z(native.mlp_pack_bytes(self.critic_desc), dtype=torch.uint8)
            ws_bytes = native.ppo_workspace_bytes_bf16(self.actor_desc, self.critic_desc,
                                                       self.U * self.mb)
        else:
            ws_bytes = native.ppo_workspace_bytes(self.actor_desc, self.critic_desc,
                                                  self.U * self.mb)
        self.workspace = z(ws_bytes, dtype=torch.uint8)
        # the loss call adds into the (zeroed) gradient vector and the optimiser kernel clears it
        # again -- legal when nobody else reads the vector: one rank, or the push exchange
        # (a learner built with ``rank_world`` has its gradients summed by its caller: plain pair)
        self.acc = (self.bf16 and self.collective in ("peer", "none") and rank_world is None
                    and bool(config.arch.get("accumulate_grads", True))
                    and not (self.world > 1 and os.environ.get("MAVA_PEER_PULL", "0") == "1"))
        n = T * self.E
        self.perm_rounds = int(math.ceil(3 * math.log(max(1, n)) / math.log(2 ** 32 - 1)))
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_update = 0  # mava_b200 kernels per update (counted on the first run)
        self.time_loss_grad = None  # list of (start, end) CUDA events when profiling (bench.py)
        self.time_reduce_apply = None
        self.compute_dtype = "bf16" if self.bf16 else "f32"
        self.dominant_kernel = ("ppo_fused_kernel + ppo_wgrad1_kernel (tcgen05, bf16)" if self.bf16
                                else "ppo_loss_grad (fp32 mlp_fwd/mlp_bwd/mlp_wgrad kernels)")

    @property
    def lr_decay_updates(self) -> int:
        """num_updates of the linear schedule (mava/utils/training.py:38-47), read when the update
        is issued / captured: run_experiment replaces ``self.config`` after check_total_timesteps
        has rewritten ``system.num_updates``, like the reference's schedule closes over the mutated
        config and is traced on the first learn() call."""
        s = self.config.system
        return int(s.num_updates) if bool(s.decay_learning_rates) else 0

    # -- views of the state ---------------------------------------------------------------------
    @property
    def actor_params(self) -> torch.Tensor:
        return self.params[: self.na]

    @property
    def critic_params(self) -> torch.Tensor:
        return self.params[self.na:]

    def learner_state(self) -> LearnerState:
        na = self.na
        params = Params(self.params[:na], self.params[na:])
        opt = OptStates({"mu": self.mu[:na], "nu": self.nu[:na], "count": self.counts[0:1]},
                        {"mu": self.mu[na:], "nu": self.nu[na:], "count": self.counts[1:2]})
        env_state = EnvState(self.env_buf, self.view[0], self.mask[0])

        def make_timestep() -> TimeStep:
            steps = self.env.step_count(env_state)
            obs = self.env.decode_observation(self.view[0], self.mask[0], steps)
            return TimeStep(
                torch.full((self.NE,), StepType.MID, dtype=torch.int8, device=self.device),
                self.reward[-1], torch.ones(self.NE, self.A, device=self.device), obs, {})

        return LearnerState(params, opt, self.key, env_state, LazyTimeStep(make_timestep))

    # -- kernels --------------------------------------------------------------------------------
    def _rollout(self) -> None:
        """ff_mappo.py:76-106: T acting + env steps, then the bootstrap value (:110)."""
        envn = self.env.native
        if self.bf16:
            self._pack()
        if self.fused_rollout:
            native.ff_rollout_bf16(envn, self.actor_desc, self.actor_params, self.actor_img,
                                   self.env_buf, self.view, self.mask, self.policy_keys, self.E,
                                   self.NE, self.T, self.action, self.logp, self.reward, self.done,
                                   self.ep_ret, self.ep_len)
            # the critic is not needed to act: all T + 1 slots in one batched launch
            native.ff_act_bf16(None, None, None, self.critic_desc, self.critic_params,
                               self.critic_img, self.view, None, None, self.E,
                               (self.T + 1) * self.NE, None, None, self.value_all)
            return
        for t in range(self.T):
            if self.bf16:
                native.ff_act_bf16(self.actor_desc, self.actor_params, self.actor_img,
                                   self.critic_desc, self.critic_params, self.critic_img,
                                   self.view[t], self.mask[t], self.policy_keys[t], self.E, self.NE,
                                   self.action[t], self.logp[t], self.value[t])
            else:
                native.ff_act(self.actor_desc, self.actor_params, self.critic_desc,
                              self.critic_params, self.view[t], self.mask[t], self.policy_keys[t],
                              self.E, self.NE, self.action[t], self.logp[t], self.value[t])
            envn.step(self.env_buf, self.action[t], self.view[t + 1], self.mask[t + 1],
                      self.reward[t], self.done[t], self.ep_ret[t], self.ep_len[t], self.NE, True)
        if self.bf16:
            native.ff_act_bf16(None, None, None, self.critic_desc, self.critic_params,
                               self.critic_img, self.view[self.T], None, None, self.E, self.NE,
                               None, None, self.last_val)
        else:
            native.ff_value(self.critic_desc, self.critic_params, self.view[self.T], self.NE,
                            self.last_val)

    def _pack(self) -> None:
        """fp32 parameters -> bf16 operand images (after every optimiser step)."""
        native.mlp_pack_bf16(self.actor_desc, self.actor_params, self.actor_img)
        native.mlp_pack_bf16(self.critic_desc, self.critic_params, self.critic_img)

    def _permutation(self, shuffle_key: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """jax.random.permutation(shuffle_key, n): rounds of a stable sort of the running permutation
        by fresh threefry bits (ff_mappo.py:273, rec_mappo.py:350-352).  Bits and sort are kernels of
        this library (csrc/env.cu, csrc/sort.cu); `slot` selects the output buffers so that the
        permutations of several epochs can be alive at once."""
        n = self.T * self.E
        src, k = self.arange_n, shuffle_key
        for r in range(self.perm_rounds):
            native.prng_split(k, self.key2_r[slot][r], 2)
            k = self.key2_r[slot][r][0]
            native.prng_random_bits(self.key2_r[slot][r][1], self.bits, n)
            dst = self.perm_buf[slot][r & 1]
            native.sort_by_key(self.bits, src, dst, n, self.sort_ws, self.sort_overflow)
            src = dst
        return src

    def _epoch_permutations(self):
        """The shuffles of all epochs (ff_mappo.py:269-273).  They depend only on the key left by
        the rollout's split chain, not on the rollout or on training, so they run on a side stream
        (a parallel branch of the CUDA graph) next to the rollout kernel, which leaves SMs idle."""
        k = self.key
        perms = []
        for ep in range(self.epochs):
            native.prng_split(k, self.key3_ep[ep], 3)  # key, shuffle_key, entropy_key (:269)
            k = self.key3_ep[ep][0]
            perms.append(self._permutation(self.key3_ep[ep][1], ep))
            if self.bf16:
                for m in range(self.nmb):
                    native.ppo_minibatch_rows(perms[ep], m, self.mb, self.U, self.E,
                                              self.rows_all[ep, m])
        return perms

    def _epochs_begin(self) -> None:
        """Before the first minibatch: the advantage statistics of every minibatch of the update on
        the side stream (GAE is done on `main`); minibatch (ep, m) waits for its own event only."""
        self._stats_done = []
        if not self.bf16:
            return
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            for ep in range(self.epochs):
                evs = []
                for m in range(self.nmb):
                    native.ppo_adv_stats(self.adv, self.rows_all[ep, m], self.U, self.mb, self.A,
                                         self.adv_stats_all[ep, m])
                    ev = torch.cuda.Event()
                    ev.record(self._side)
                    evs.append(ev)
                self._stats_done.append(evs)

    def _minibatch_grad(self, ep: int, m: int, perms) -> None:
        """value_and_grad of both losses on minibatch m of epoch ep, averaged over the replicas
        (ff_mappo.py:150-226) -> self.grad = [actor grads | critic grads | 5 loss scalars]."""
        if self.bf16:
            torch.cuda.current_stream().wait_event(self._stats_done[ep][m])
        else:
            native.ppo_minibatch_rows(perms[ep], m, self.mb, self.U, self.E, self.rows)
        if self.time_loss_grad is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.bf16:
            loss_grad = native.ppo_loss_grad_bf16_acc if self.acc else native.ppo_loss_grad_bf16_stats
            loss_grad(
                self.actor_desc, self.actor_params, self.actor_img, self.critic_desc,
                self.critic_params, self.critic_img, self.hyper, self.view, self.mask,
                self.action, self.logp, self.value, self.adv, self.targets,
                self.rows_all[ep, m], self.U, self.mb, self.adv_stats_all[ep, m], self.grad,
                self.workspace)
        else:
            native.ppo_loss_grad(self.actor_desc, self.actor_params, self.critic_desc,
                                 self.critic_params, self.hyper, self.view, self.mask,
                                 self.action, self.logp, self.value, self.adv, self.targets,
                                 self.rows, self.U, self.mb, self.grad, self.workspace)
        if self.time_loss_grad is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.time_loss_grad.append((e0, e1))

    def _minibatch_apply(self, ep: int, m: int) -> None:
        """pmean("device") of the gradient buffers -> clip_by_global_norm -> adam -> apply_updates
        -> refreshed bf16 operand images -> the minibatch's loss metrics (ff_mappo.py:228-250,
        260-265): ONE launch (csrc/peer.cu).  With collective == "nccl" (or gradients already summed
        by the caller) the kernel runs on the local buffer and only applies 1/world."""
        s = self.config.system
        steps_per_update = self.epochs * self.nmb
        group = self.peer if self.collective == "peer" else self.peer_local
        if self.time_reduce_apply is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.collective == "nccl":
            self.allreduce(self.grad)
        if self.acc:
            # paired with ppo_loss_grad_bf16_acc: loss metrics from the accumulators, gradient
            # vector and accumulators left zero (no memsets, no finalize launch per minibatch)
            native.reduce_clip_adam_pair_acc(
                self.params, self.mu, self.nu, self.counts, group, self.gsum, self.na, self.nc,
                self.actor_desc, self.actor_img, self.critic_desc, self.critic_img,
                1.0 / self.world, float(s.actor_lr), float(s.critic_lr), float(s.max_grad_norm),
                self.lr_decay_updates, steps_per_update, self.loss_buf[ep, m], self.workspace,
                self.hyper, self.U * self.mb * self.A)
        else:
            native.reduce_clip_adam_pair(
                self.params, self.mu, self.nu, self.counts, group, self.gsum, self.na, self.nc,
                self.actor_desc if self.bf16 else None, self.actor_img if self.bf16 else None,
                self.critic_desc if self.bf16 else None, self.critic_img if self.bf16 else None,
                1.0 / self.world, float(s.actor_lr), float(s.critic_lr), float(s.max_grad_norm),
                self.lr_decay_updates, steps_per_update, self.loss_buf[ep, m])
        if self.time_reduce_apply is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.time_reduce_apply.append((e0, e1))

    def _epochs_end(self) -> None:
        self.key.copy_(self.key3_ep[self.epochs - 1][0])

    def _update_epochs(self, perms) -> None:
        """ff_mappo.py:141-295."""
        self._epochs_begin()
        for ep in range(self.epochs):
            for m in range(self.nmb):
                self._minibatch_grad(ep, m, perms)
                self._minibatch_apply(ep, m)
        self._epochs_end()

    def _rollout_and_gae(self):
        """ff_mappo.py:76-139 for all U replicas; returns the epoch permutations (prepared on the
        side stream next to the rollout)."""
        native.prng_split_chain(self.key, self.policy_keys, self.T)  # key, policy_key per step (:81)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            perms = self._epoch_permutations()
        self._rollout()
        # finished-episode statistics: off the critical path, next to GAE (a graph branch)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            native.episode_stats(self.done, self.ep_ret, self.ep_len, self.T * self.NE, False,
                                 self.ep_stats)
        native.gae(self.reward, self.value, self.done, self.last_val, float(self.config.system.gamma),
                   float(self.config.system.gae_lambda), self.T, self.NE, self.A, self.adv,
                   self.targets)
        main.wait_stream(self._side)
        return perms

    def _carry_over(self) -> None:
        """The bootstrap observation becomes slot 0 of the next update."""
        self.view[0].copy_(self.view[self.T])
        self.mask[0].copy_(self.mask[self.T])

    def _update_step(self) -> None:
        """One ``_update_step`` of the reference (ff_mappo.py:56-300) for all U replicas."""
        n0 = native.LAUNCHES
        perms = self._rollout_and_gae()
        self._update_epochs(perms)
        self._carry_over()
        self.launches_per_update = native.LAUNCHES - n0

    # -- CUDA graph -----------------------------------------------------------------------------
    def _state_tensors(self):
        return [self.params, self.mu, self.nu, self.counts, self.key, self.env_buf, self.view,
                self.mask]

    def _capture(self) -> None:
        snap = [t.clone() for t in self._state_tensors()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up run: lazy initialisation, allocator priming
            self._update_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        for t, c in zip(self._state_tensors(), snap):
            t.copy_(c)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._update_step()
        for t, c in zip(self._state_tensors(), snap):  # capture does not execute, but be explicit
            t.copy_(c)
        self._graph = g

    def check_sort(self) -> None:
        """The permutation sort assumes uniform keys; a bucket overflow is raised by the learn()
        call that produced it (the flag is read after that call's work has drained)."""
        if self._report_host is None:
            self._report_host = torch.zeros_like(self.report, device="cpu").pin_memory()
            self.loss_host = self._report_host[96:].view(torch.float32).view(self.epochs, self.nmb, 5)
        # the flag travels with the finished-episode statistics and the losses of the call; with
        # several ranks the statistics of all of them are gathered first (the reference reduces over
        # the pmap output of all devices), so that the host waits once
        gathered = self.world > 1 and self.allreduce is not None
        if gathered:
            if getattr(self, "_stats_all", None) is None:
                self._stats_all = torch.zeros(self.world, 10, dtype=torch.float64, device=self.device)
                self._stats_all_host = torch.zeros(self.world, 10, dtype=torch.float64).pin_memory()
            dist.all_gather_into_tensor(self._stats_all, self.ep_stats)
            self._stats_all_host.copy_(self._stats_all, non_blocking=True)
        self._report_host.copy_(self.report, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._report_fresh = True
        self._gather_fresh = gathered
        if int(self._report_host[80:84].view(torch.int32)[0]) != 0:
            raise RuntimeError("mava_sort_by_key: bucket overflow (non-uniform sort keys); the "
                               "parameters of this learn() call are not to be trusted")
        self.peer.check()  # a peer handshake of the fused all-reduce timed out

    def release(self) -> None:
        """Drop the CUDA graph and the peer mappings (before the process group is destroyed)."""
        self._graph = None
        if self.peer is not None:
            if self.peer_local is not self.peer:
                self.peer_local.release()
            self.peer.release()
            self.peer = self.peer_local = None
            self.grad = None

    # -- public -----------------------------------------------------------------------------------
    def learn(self, num_updates: int) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        dev, T, NE = self.device, self.T, self.NE
        if self.use_graph and self._graph is None:
            self._capture()
        self._report_fresh = self._gather_fresh = False
        native.episode_stats(None, None, None, 0, True, self.ep_stats)  # re-initialise
        single = num_updates == 1  # the rollout buffers themselves are the metrics: no copies
        if not single:
            ep_ret = torch.empty(num_updates, T, NE, device=dev)
            ep_len = torch.empty(num_updates, T, NE, dtype=torch.int32, device=dev)
            term = torch.empty(num_updates, T, NE, dtype=torch.bool, device=dev)
            losses = torch.empty(num_updates, self.epochs, self.nmb, 5, device=dev)
        for u in range(num_updates):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._update_step()
            if not single:
                ep_ret[u].copy_(self.ep_ret)
                ep_len[u].copy_(self.ep_len)
                term[u].copy_(self.done)
                losses[u].copy_(self.loss_buf)
        if single:
            ep_ret, ep_len = self.ep_ret.unsqueeze(0), self.ep_len.unsqueeze(0)
            term, losses = self.done.view(torch.bool).unsqueeze(0), self.loss_buf.unsqueeze(0)
        shape = lambda x: x.reshape(num_updates, T, self.U, self.E).permute(0, 2, 1, 3)
        episode_metrics = {"episode_return": shape(ep_ret), "episode_length": shape(ep_len),
                           "is_terminal_step": shape(term)}
        train_metrics = {"total_loss": losses[..., 0] + losses[..., 3], "value_loss": losses[..., 4],
                         "actor_loss": losses[..., 1], "entropy": losses[..., 2]}
        return episode_metrics, train_metrics


def episode_summary(learner) -> Tuple[Dict[str, Dict[str, float]], bool]:
    """What the run loop logs about finished episodes (get_final_step_metrics + the logger's
    describe(), mava/wrappers/episode_metrics.py:114-132, mava/utils/logger.py:44-58) from the
    80-byte device reduction of the last learn() call: {metric: {mean, std, min, max}}, and whether
    any episode finished."""
    if learner._stats_host is None:
        learner._stats_host = torch.zeros(10, dtype=torch.float64).pin_memory()
    if learner.world > 1 and learner.allreduce is not None:
        # the reference reduces over the pmap output of all devices: ONE all_gather of the ranks'
        # 80-byte reductions, combined on the host (sums add, extrema combine)
        if getattr(learner, "_stats_all", None) is None:
            learner._stats_all = torch.zeros(learner.world, 10, dtype=torch.float64,
                                             device=learner.device)
            learner._stats_all_host = torch.zeros(learner.world, 10, dtype=torch.float64).pin_memory()
        if not getattr(learner, "_gather_fresh", False):  # else: check_sort() of the call did it
            dist.all_gather_into_tensor(learner._stats_all, learner.ep_stats)
            learner._stats_all_host.copy_(learner._stats_all, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        a = learner._stats_all_host
        n, sr, qr = float(a[:, 0].sum()), float(a[:, 1].sum()), float(a[:, 2].sum())
        mnr, mxr = float(a[:, 3].min()), float(a[:, 4].max())
        sl, ql = float(a[:, 5].sum()), float(a[:, 6].sum())
        mnl, mxl = float(a[:, 7].min()), float(a[:, 8].max())
    elif getattr(learner, "_report_fresh", False):
        # check_sort() of the learn() call has already brought them over with the overflow flag
        n, sr, qr, mnr, mxr, sl, ql, mnl, mxl, _ = learner._report_host[:80].view(torch.float64).tolist()
    else:
        learner._stats_host.copy_(learner.ep_stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n, sr, qr, mnr, mxr, sl, ql, mnl, mxl, _ = learner._stats_host.tolist()
    if n == 0:
        zero = {"mean": 0.0, "std": 0.0, "min": 0.0, "max": 0.0}
        return {"episode_return": dict(zero), "episode_length": dict(zero)}, False

    def desc(s1, s2, mn, mx):
        mean = s1 / n
        return {"mean": mean, "std": math.sqrt(max(s2 / n - mean * mean, 0.0)), "min": mn, "max": mx}

    return {"episode_return": desc(sr, qr, mnr, mxr), "episode_length": desc(sl, ql, mnl, mxl),
            "count": n}, True


class LazyTimeStep:
    """The TimeStep of a learner state, decoded from the packed device buffers on first access.
    The run loop only threads the state back into ``learn`` (ff_mappo.py:497-536), so the float32
    observation tensors of the reference layout are normally never built."""

    def __init__(self, make):
        self._make, self._ts = make, None

    def _get(self) -> TimeStep:
        if self._ts is None:
            self._ts = self._make()
        return self._ts

    def __getattr__(self, name):
        return getattr(self._get(), name)

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]

    def last(self):
        return self._get().last()


def get_learner_fn(learner: FFLearner, config):
    """The ``learn`` callable of the reference (ff_mappo.py:302-330): LearnerState -> ExperimentOutput.
    The state's tensors alias the learner's buffers and are advanced in place."""

    def learner_fn(learner_state: LearnerState) -> ExperimentOutput:
        _adopt(learner, learner_state)
        n = int(config.system.get("num_updates_per_eval", 1))
        episode_metrics, train_metrics = learner.learn(n)
        learner.check_sort()
        return ExperimentOutput(learner.learner_state(), episode_metrics, train_metrics)

    return learner_fn


def _adopt(learner: FFLearner, st: LearnerState) -> None:
    """Copy a foreign state (e.g. restored parameters) into the learner's buffers."""
    pairs = [(learner.params[: learner.na], st.params.actor_params),
             (learner.params[learner.na:], st.params.critic_params),
             (learner.key, st.key), (learner.env_buf, st.env_state.buf),
             (learner.view[0], st.env_state.view), (learner.mask[0], st.env_state.mask)]
    for dst, src in pairs:
        if src.data_ptr() != dst.data_ptr():
            dst.copy_(src)


def learner_setup(env: NativeMarlEnv, keys, config, centralised_critic: bool,
                  device: Optional[torch.device] = None,
                  rank_world: Optional[Tuple[int, int]] = None):
    """ff_mappo.py:333-432: networks, optimiser state, env reset, replicated learner state.
    ``rank_world`` builds the learner of one rank of a larger job without a process group."""
    from ...networks import instantiate

    device = device or env.device
    rank, n_devices = rank_world if rank_world is not None else world()
    config.system.num_agents = env.num_agents
    key, actor_net_key, critic_net_key = keys

    actor_torso = instantiate(config.network.actor_network.pre_torso)
    action_head = instantiate(config.network.action_head, action_dim=env.action_dim)
    critic_torso = instantiate(config.network.critic_network.pre_torso)
    actor_network = FeedForwardActor(torso=actor_torso, action_head=action_head)
    critic_network = FeedForwardValueNet(torso=critic_torso, centralised_critic=centralised_critic)

    learner = FFLearner(env, actor_network, critic_network, config, centralised_critic, device,
                        rank_world)
    ap = actor_network.init(actor_net_key, learner.actor_desc.in_dim)
    cp = critic_network.init(critic_net_key, learner.critic_desc.in_dim)
    # Load model from checkpoint if specified (ff_mappo.py:405-414, rec_mappo.py:527-536).
    if config.logger.checkpointing.load_model:
        from ...utils.checkpointing import Checkpointer

        loaded = Checkpointer(model_name=config.logger.system_name,
                              **dict(config.logger.checkpointing.load_args))
        ap, cp = loaded.restore_params(actor_network, critic_network)
    learner.params.copy_(torch.from_numpy(np.concatenate([ap, cp])).to(device))
    learner.networks = (actor_network, critic_network)

    # env keys: one per (device, replica, env), this rank takes its block (ff_mappo.py:392-403)
    per_dev = learner.U * learner.E
    all_keys = prng.split(key, n_devices * per_dev + 1)
    key, env_keys = all_keys[0], all_keys[1 + rank * per_dev: 1 + (rank + 1) * per_dev]
    env.native.reset(_u32(env_keys, device), learner.env_buf, learner.view[0], learner.mask[0],
                     learner.NE)
    # the same step key on every replica and device (ff_mappo.py:417-426)
    key, step_key = prng.split(key)
    learner.key.copy_(_u32(step_key, device))

    learn = get_learner_fn(learner, config)
    learn.learner = learner  # type: ignore[attr-defined]
    return learn, actor_network, learner.learner_state()
