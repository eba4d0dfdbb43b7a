"""The recurrent Anakin PPO learner shared by rec_ippo / rec_mappo.

Mirrors ``get_learner_fn`` / ``learner_setup`` of mava/systems/ppo/rec_mappo.py:59-560 (rec_ippo.py
is the same file with a decentralised critic).  Same process model as ``anakin.py`` (one process
per GPU, replicas side by side on the env axis, one NCCL all-reduce per minibatch); what differs
from the feed-forward learner is exactly what differs in the reference:

* both networks carry a GRU hidden state through the rollout; the flag stored with a transition
  is the done flag ENTERING the step (``last_done``, rec_mappo.py:134-143) and resets the carry;
* GAE bootstraps with ``next_done`` (rec_mappo.py:177-199) -> ``mava_gae(rec=1)``;
* minibatches are whole sequences: the rollout is viewed as (chunk, E * num_chunks) exactly like
  rec_mappo.py:339-349, the permutation runs over those columns, and only the hidden state entering
  the first position of a column is used (:221,254), so only ``num_chunks`` hidden states per env
  are stored instead of one per step.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ... import native, prng
from ..._lib import PpoHyper
from ...peer import PeerGroup
from ...networks import RecurrentActor, RecurrentValueNet
from ...types import (ExperimentOutput, HiddenStates, OptStates, Params, RNNLearnerState, StepType,
                      TimeStep)
from ...wrappers import EnvState, NativeMarlEnv
from .anakin import _nccl_allreduce, _u32, world


class RecLearner:
    """Device buffers + the kernel schedule of one GPU's share of the recurrent learner."""

    def __init__(self, env: NativeMarlEnv, actor: RecurrentActor, critic: RecurrentValueNet,
                 config, centralised_critic: bool, device: torch.device,
                 rank_world: Optional[Tuple[int, int]] = None):
        s = config.system
        self.env, self.config, self.device = env, config, device
        self.rank, self.world = rank_world if rank_world is not None else world()
        self.allreduce = _nccl_allreduce if rank_world is None else None
        self.collective = str(config.arch.get("collective", "peer"))  # see FFLearner
        if self.collective not in ("peer", "nccl"):
            raise ValueError(f"arch.collective must be peer or nccl, got {self.collective}")
        if self.world == 1 or rank_world is not None:
            self.collective = "none"
        self.T, self.U, self.E = int(s.rollout_length), int(s.update_batch_size), int(
            config.arch.num_envs)
        self.NE = self.U * self.E
        self.dense = bool(getattr(env, "dense", False))  # f32 observation rows (synthetic SMAX)
        self.A, self.N = env.num_agents, env.action_dim
        self.FR = 1 if self.dense else env.native.view_dim
        self.epochs, self.nmb = int(s.ppo_epochs), int(s.num_minibatches)
        chunk = s.get("recurrent_chunk_size", None)
        self.chunk = int(chunk) if chunk else self.T
        if self.T % self.chunk != 0:
            raise ValueError("Rollout length must be divisible by recurrent chunk size.")
        self.nc = self.T // self.chunk
        if (self.E * self.nc) % self.nmb != 0:
            raise ValueError("num_envs * num_chunks must be divisible by num_minibatches")
        self.mbc = self.E * self.nc // self.nmb
        add_id = bool(s.add_agent_id)
        # precision: bf16 tensor-core contractions unless fp32 is asked for (as the ff learner)
        self.precision = str(config.arch.get("precision", "auto"))
        if self.precision not in ("auto", "fp32", "bf16"):
            raise ValueError(f"arch.precision must be auto, fp32 or bf16, got {self.precision}")
        self.bf16 = self.precision != "fp32"
        pr = 1 if self.bf16 else 0
        if self.dense:
            # AgentIDWrapper appends nothing here: the synthetic rows already have their width
            self.actor_desc = actor.desc(self.A, 1, False, native.IN_DENSE, env.obs_dim, self.A, pr)
            self.critic_desc = (critic.desc(self.A, 1, False, native.IN_DENSE, env.state_dim, 1, pr)
                                if centralised_critic else
                                critic.desc(self.A, 1, False, native.IN_DENSE, env.obs_dim, self.A, pr))
        else:
            self.actor_desc = actor.desc(self.A, self.FR, add_id, native.IN_AGENT_VIEW, precision=pr)
            self.critic_desc = critic.desc(
                self.A, self.FR, add_id,
                native.IN_GLOBAL if centralised_critic else native.IN_AGENT_VIEW, precision=pr)
        self.centralised_critic = centralised_critic
        self.H = self.actor_desc.hidden
        self.rpc = self.critic_desc.rows_per_env
        self.na = native.rnn_param_count(self.actor_desc)
        self.nc_params = native.rnn_param_count(self.critic_desc)
        self.hyper = PpoHyper(float(s.clip_eps), float(s.ent_coef), float(s.vf_coef))
        self.use_graph = bool(config.arch.get("use_cuda_graph", True))

        dev, T, NE, A = device, self.T, self.NE, self.A
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        n_all = self.na + self.nc_params
        self.params, self.mu, self.nu = z(n_all), z(n_all), z(n_all)
        self.counts = z(2, dtype=torch.int32)
        self.key = z(2, dtype=torch.uint32)
        self.env_buf = env.native.alloc_state(NE, dev)
        self.h_actor = z(NE * A, self.H)
        self.h_critic = z(NE * self.rpc, self.H)
        # rollout buffers; done_in[t] is the flag entering step t, slot T is last_done
        if self.dense:
            self.view = None
            self.obs_a = z(T + 1, NE, A, env.obs_dim)
            self.obs_c = (z(T + 1, NE, 1, env.state_dim) if centralised_critic else self.obs_a)
            self.mask = z(T + 1, NE, A, dtype=torch.uint16 if self.N > 8 else torch.uint8)
        else:
            self.view = z(T + 1, NE, A, self.FR, dtype=torch.int8)
            self.obs_a = self.obs_c = None
            self.mask = z(T + 1, NE, A, dtype=torch.uint8)
        self.done_in = z(T + 1, NE, dtype=torch.uint8)
        self.action = z(T, NE, A, dtype=torch.int8)
        self.logp, self.value, self.reward = z(T, NE, A), z(T, NE, A), z(T, NE, A)
        self.ep_ret = z(T, NE)
        self.ep_len = z(T, NE, dtype=torch.int32)
        self.ep_stats = z(10, dtype=torch.float64)  # finished-episode statistics (device reduction)
        self._stats_host = None
        self.last_val = z(NE, A)
        self.adv, self.targets = z(T, NE, A), z(T, NE, A)
        self.hs_actor = z(self.nc, NE * A, self.H)
        self.hs_critic = z(self.nc, NE * self.rpc, self.H)
        # scratch
        self.policy_keys = z(T, 2, dtype=torch.uint32)
        self.key3 = z(3, 2, dtype=torch.uint32)
        self.key2 = z(2, 2, dtype=torch.uint32)
        self.ncols = self.E * self.nc
        self.bits = z(self.ncols, dtype=torch.uint32)
        self.peer = PeerGroup(n_all + 8, dev, self.rank, self.world) if self.collective == "peer" \
            else PeerGroup(n_all + 8, dev, 0, 1)
        self.peer_local = self.peer if self.peer.world == 1 else \
            PeerGroup(n_all + 8, dev, 0, 1, local_bufs=[self.peer.struct.buf[self.rank]])
        self.grad = self.peer.grad  # this rank's exchange buffer (csrc/peer.cu)
        self.gsum = z(n_all)
        self.loss_buf = z(self.epochs, self.nmb, 5)
        self.act_ws = z(native.rec_act_workspace_bytes(self.actor_desc, self.critic_desc, NE),
                        dtype=torch.uint8)
        self.workspace = z(native.rec_ppo_workspace_bytes(self.actor_desc, self.critic_desc,
                                                          self.U * self.mbc, self.chunk),
                           dtype=torch.uint8)
        self._scratch_state = z(NE, env.state_dim) if self.dense and not centralised_critic else None
        self.perm_rounds = int(math.ceil(3 * math.log(max(1, self.ncols)) / math.log(2 ** 32 - 1)))
        self.arange_n = torch.arange(self.ncols, dtype=torch.int32, device=dev)
        self.perm_buf = z(1, 2, self.ncols, dtype=torch.int32)
        self.key2_r = z(1, max(1, self.perm_rounds), 2, 2, dtype=torch.uint32)
        self.sort_ws = z(native.sort_workspace_bytes(self.ncols), dtype=torch.uint8)
        self.sort_overflow = z(1, dtype=torch.int32)
        self._ovf_host = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_update = 0
        self.time_loss_grad = None
        self.time_reduce_apply = None
        self.compute_dtype = "bf16" if self.bf16 else "f32"
        self.dominant_kernel = ("tc_gemm_kernel (tcgen05 bf16 GRU scan + dense layers)" if self.bf16
                                else "sgemm_kernel (fp32 GRU scan + dense layers)")

    @property
    def lr_decay_updates(self) -> int:
        """Read when the update is issued / captured (see FFLearner.lr_decay_updates)."""
        s = self.config.system
        return int(s.num_updates) if bool(s.decay_learning_rates) else 0

    # -- views of the state ---------------------------------------------------------------------
    @property
    def actor_params(self) -> torch.Tensor:
        return self.params[: self.na]

    @property
    def critic_params(self) -> torch.Tensor:
        return self.params[self.na:]

    def learner_state(self) -> RNNLearnerState:
        na = self.na
        params = Params(self.params[:na], self.params[na:])
        opt = OptStates({"mu": self.mu[:na], "nu": self.nu[:na], "count": self.counts[0:1]},
                        {"mu": self.mu[na:], "nu": self.nu[na:], "count": self.counts[1:2]})
        if self.dense:
            env_state = EnvState(self.env_buf, self.obs_a[0], self.mask[0])
            obs = {"agents_view": self.obs_a[0], "action_mask": self.mask[0],
                   "global_state": self.obs_c[0]}
        else:
            env_state = EnvState(self.env_buf, self.view[0], self.mask[0])
            steps = self.env.step_count(env_state)
            obs = self.env.decode_observation(self.view[0], self.mask[0], steps)
        ts = TimeStep(torch.full((self.NE,), StepType.MID, dtype=torch.int8, device=self.device),
                      self.reward[-1], torch.ones(self.NE, self.A, device=self.device), obs, {})
        dones = self.done_in[0].bool().unsqueeze(-1).expand(self.NE, self.A)
        hstates = HiddenStates(self.h_actor.view(self.NE, self.A, self.H),
                               self.h_critic.view(self.NE, self.rpc, self.H))
        return RNNLearnerState(params, opt, self.key, env_state, ts, dones, hstates)

    # -- kernels --------------------------------------------------------------------------------
    def _rollout(self) -> None:
        """rec_mappo.py:91-172: T acting + env steps, then the bootstrap value."""
        envn = self.env.native
        native.prng_split_chain(self.key, self.policy_keys, self.T)
        for t in range(self.T):
            if t < self.nc:  # the hidden state entering a column's first position (:221,254)
                self.hs_actor[t].copy_(self.h_actor)
                self.hs_critic[t].copy_(self.h_critic)
            view_t = None if self.dense else self.view[t]
            oa_t = self.obs_a[t] if self.dense else None
            oc_t = self.obs_c[t] if self.dense else None
            native.rec_act(self.actor_desc, self.actor_params, self.critic_desc, self.critic_params,
                           view_t, oa_t, oc_t, self.mask[t], self.done_in[t], self.h_actor,
                           self.h_actor, self.h_critic, self.h_critic, self.policy_keys[t], self.E,
                           self.NE, self.action[t], self.logp[t], self.value[t], self.act_ws)
            if self.dense:
                oc_next = self.obs_c[t + 1] if self.centralised_critic else self._scratch_state
                envn.step(self.policy_keys[t], self.env_buf, self.action[t], self.obs_a[t + 1],
                          oc_next, self.mask[t + 1], self.reward[t], self.done_in[t + 1],
                          self.ep_ret[t], self.ep_len[t], self.NE)
            else:
                envn.step(self.env_buf, self.action[t], self.view[t + 1], self.mask[t + 1],
                          self.reward[t], self.done_in[t + 1], self.ep_ret[t], self.ep_len[t],
                          self.NE, True)
        # bootstrap value; the advanced critic state is discarded (rec_mappo.py:165)
        native.rec_act(None, None, self.critic_desc, self.critic_params,
                       None if self.dense else self.view[self.T], None,
                       self.obs_c[self.T] if self.dense else None, None, self.done_in[self.T], None,
                       None, self.h_critic, None, None, self.E, self.NE, None, None, self.last_val,
                       self.act_ws)

    def _permutation(self, shuffle_key: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """jax.random.permutation(shuffle_key, n): rounds of a stable sort of the running permutation
        by fresh threefry bits (ff_mappo.py:273, rec_mappo.py:350-352).  Bits and sort are kernels of
        this library (csrc/env.cu, csrc/sort.cu); `slot` selects the output buffers so that the
        permutations of several epochs can be alive at once."""
        n = self.ncols
        src, k = self.arange_n, shuffle_key
        for r in range(self.perm_rounds):
            native.prng_split(k, self.key2_r[slot][r], 2)
            k = self.key2_r[slot][r][0]
            native.prng_random_bits(self.key2_r[slot][r][1], self.bits, n)
            dst = self.perm_buf[slot][r & 1]
            native.sort_by_key(self.bits, src, dst, n, self.sort_ws, self.sort_overflow)
            src = dst
        return src

    def _update_epochs(self) -> None:
        """rec_mappo.py:201-383."""
        s = self.config.system
        na, nc = self.na, self.nc_params
        scale = 1.0 / self.world
        steps_per_update = self.epochs * self.nmb
        for ep in range(self.epochs):
            native.prng_split(self.key, self.key3, 3)  # key, shuffle_key, entropy_key (:332)
            self.key.copy_(self.key3[0])
            perm = self._permutation(self.key3[1], 0)
            for m in range(self.nmb):
                cols = perm[m * self.mbc:(m + 1) * self.mbc]
                if self.time_loss_grad is not None:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                native.rec_ppo_loss_grad(
                    self.actor_desc, self.actor_params, self.critic_desc, self.critic_params,
                    self.hyper, self.view, self.obs_a, self.obs_c, self.mask, self.action, self.logp,
                    self.value, self.adv, self.targets, self.done_in, self.hs_actor,
                    self.hs_critic, cols, self.U, self.E, self.mbc, self.chunk, self.nc, self.grad,
                    self.workspace)
                if self.time_loss_grad is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    self.time_loss_grad.append((e0, e1))
                # pmean("device") + clip + Adam + loss metrics in one launch (rec_mappo.py:283-305)
                if self.time_reduce_apply is not None:
                    r0 = torch.cuda.Event(enable_timing=True)
                    r0.record()
                if self.collective == "nccl":
                    self.allreduce(self.grad)
                native.reduce_clip_adam_pair(
                    self.params, self.mu, self.nu, self.counts,
                    self.peer if self.collective == "peer" else self.peer_local, self.gsum, na, nc,
                    None, None, None, None, scale, float(s.actor_lr), float(s.critic_lr),
                    float(s.max_grad_norm), self.lr_decay_updates, steps_per_update,
                    self.loss_buf[ep, m])
                if self.time_reduce_apply is not None:
                    r1 = torch.cuda.Event(enable_timing=True)
                    r1.record()
                    self.time_reduce_apply.append((r0, r1))

    def _update_step(self) -> None:
        """One ``_update_step`` of the reference (rec_mappo.py:68-402) for all U replicas."""
        n0 = native.LAUNCHES
        self._rollout()
        # is_terminal_step of transition t is the flag entering step t + 1
        native.episode_stats(self.done_in[1:], self.ep_ret, self.ep_len, self.T * self.NE, False,
                             self.ep_stats)
        native.gae(self.reward, self.value, self.done_in, self.last_val,
                   float(self.config.system.gamma), float(self.config.system.gae_lambda), self.T,
                   self.NE, self.A, self.adv, self.targets, last_done=self.done_in[self.T])
        self._update_epochs()
        if self.dense:
            self.obs_a[0].copy_(self.obs_a[self.T])
            if self.centralised_critic:
                self.obs_c[0].copy_(self.obs_c[self.T])
        else:
            self.view[0].copy_(self.view[self.T])
        self.mask[0].copy_(self.mask[self.T])
        self.done_in[0].copy_(self.done_in[self.T])
        self.launches_per_update = native.LAUNCHES - n0

    # -- CUDA graph -----------------------------------------------------------------------------
    def _state_tensors(self):
        obs = [self.obs_a] + ([self.obs_c] if self.centralised_critic else []) if self.dense \
            else [self.view]
        return [self.params, self.mu, self.nu, self.counts, self.key, self.env_buf, self.mask,
                self.done_in, self.h_actor, self.h_critic] + obs

    def _capture(self) -> None:
        snap = [t.clone() for t in self._state_tensors()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._update_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        for t, c in zip(self._state_tensors(), snap):
            t.copy_(c)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._update_step()
        for t, c in zip(self._state_tensors(), snap):
            t.copy_(c)
        self._graph = g

    def check_sort(self) -> None:
        """Bucket overflow of the permutation sort, raised by the learn() call that produced it."""
        if self._ovf_host is None:
            self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._ovf_host.copy_(self.sort_overflow, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if int(self._ovf_host[0]) != 0:
            raise RuntimeError("mava_sort_by_key: bucket overflow (non-uniform sort keys); the "
                               "parameters of this learn() call are not to be trusted")
        self.peer.check()

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
        ep_ret = torch.empty(num_updates, T, NE, device=dev)
        ep_len = torch.empty(num_updates, T, NE, dtype=torch.int32, device=dev)
        term = torch.empty(num_updates, T, NE, dtype=torch.bool, device=dev)
        losses = torch.empty(num_updates, self.epochs, self.nmb, 5, device=dev)
        if self.use_graph and self._graph is None:
            self._capture()
        native.episode_stats(None, None, None, 0, True, self.ep_stats)  # re-initialise
        for u in range(num_updates):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._update_step()
            ep_ret[u].copy_(self.ep_ret)
            ep_len[u].copy_(self.ep_len)
            # is_terminal_step of transition t is the flag entering step t + 1; slot 0 was
            # overwritten with slot T at the end of the update, slots 1..T are intact
            term[u].copy_(self.done_in[1:])
            losses[u].copy_(self.loss_buf)
        shape = lambda x: x.reshape(num_updates, T, self.U, self.E).permute(0, 2, 1, 3)
        episode_metrics = {"episode_return": shape(ep_ret), "episode_length": shape(ep_len),
                           "is_terminal_step": shape(term)}
        train_metrics = {"total_loss": losses[..., 0] + losses[..., 3], "value_loss": losses[..., 4],
                         "actor_loss": losses[..., 1], "entropy": losses[..., 2]}
        return episode_metrics, train_metrics


def get_learner_fn(learner: RecLearner, config):
    """The ``learn`` callable (rec_mappo.py:404-432): RNNLearnerState -> ExperimentOutput."""

    def learner_fn(learner_state: RNNLearnerState) -> ExperimentOutput:
        _adopt(learner, learner_state)
        n = int(config.system.get("num_updates_per_eval", 1))
        episode_metrics, train_metrics = learner.learn(n)
        learner.check_sort()
        return ExperimentOutput(learner.learner_state(), episode_metrics, train_metrics)

    return learner_fn


def _adopt(learner: RecLearner, st: RNNLearnerState) -> None:
    pairs = [(learner.params[: learner.na], st.params.actor_params),
             (learner.params[learner.na:], st.params.critic_params),
             (learner.key, st.key), (learner.env_buf, st.env_state.buf),
             (learner.obs_a[0] if learner.dense else learner.view[0], st.env_state.view),
             (learner.mask[0], st.env_state.mask)]
    if st.hstates is not None:
        pairs += [(learner.h_actor, st.hstates.policy_hidden_state.reshape(learner.h_actor.shape)),
                  (learner.h_critic, st.hstates.critic_hidden_state.reshape(learner.h_critic.shape))]
    for dst, src in pairs:
        if src.data_ptr() != dst.data_ptr():
            dst.copy_(src)


def learner_setup(env: NativeMarlEnv, keys, config, centralised_critic: bool,
                  device: Optional[torch.device] = None,
                  rank_world: Optional[Tuple[int, int]] = None):
    """rec_mappo.py:435-560: networks, optimiser state, env reset, replicated learner state."""
    from ...networks import instantiate

    device = device or env.device
    rank, n_devices = rank_world if rank_world is not None else world()
    config.system.num_agents = env.num_agents
    key, actor_net_key, critic_net_key = keys
    if config.system.get("recurrent_chunk_size", None) is None:
        config.system.recurrent_chunk_size = config.system.rollout_length

    hsd = int(config.network.hidden_state_dim)
    actor_network = RecurrentActor(
        pre_torso=instantiate(config.network.actor_network.pre_torso),
        post_torso=instantiate(config.network.actor_network.post_torso),
        action_head=instantiate(config.network.action_head, action_dim=env.action_dim),
        hidden_state_dim=hsd)
    critic_network = RecurrentValueNet(
        pre_torso=instantiate(config.network.critic_network.pre_torso),
        post_torso=instantiate(config.network.critic_network.post_torso),
        centralised_critic=centralised_critic, hidden_state_dim=hsd)

    learner = RecLearner(env, actor_network, critic_network, config, centralised_critic, device,
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

    per_dev = learner.U * learner.E
    all_keys = prng.split(key, n_devices * per_dev + 1)
    key, env_keys = all_keys[0], all_keys[1 + rank * per_dev: 1 + (rank + 1) * per_dev]
    if learner.dense:
        oc0 = learner.obs_c[0] if centralised_critic else learner._scratch_state
        env.native.reset(_u32(env_keys[0], device), learner.env_buf, learner.obs_a[0], oc0,
                         learner.mask[0], learner.NE)
    else:
        env.native.reset(_u32(env_keys, device), learner.env_buf, learner.view[0], learner.mask[0],
                         learner.NE)
    key, step_key = prng.split(key)
    learner.key.copy_(_u32(step_key, device))

    learn = get_learner_fn(learner, config)
    learn.learner = learner  # type: ignore[attr-defined]
    return learn, actor_network, learner.learner_state()
