"""Run under torchrun with N >= 2 GPUs (tests/test_peer_gpu.py::test_peer_two_gpus, or by hand:
`torchrun --nproc-per-node N tests/peer_multi_gpu.py`).  Each rank builds the ff_mappo learner twice
-- arch.collective=peer (the fused NVLink all-reduce inside the optimiser kernel) and
arch.collective=nccl (dist.all_reduce + the same kernel on the local buffer) -- from the same seed,
runs the same updates with CUDA graphs and compares parameters and keys; before that the collective
itself is checked on identical gradients.  World 2: bit-identical (a two-term sum has one order).
World > 2: NCCL's reduction order differs from rank order, so parameters agree to fp32 rounding."""
import os
import sys

ROOT = os.environ.get("MAVA_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def exact_collective_check(device, rank, world):
    """Same gradients through both paths: the fused peer all-reduce + optimiser against
    dist.all_reduce + the same kernel on the local buffer.  World 2: bit for bit."""
    from mava_b200 import native
    from mava_b200.peer import PeerGroup

    A, FR, N = 4, 66, 5
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
    n = na + nc
    g = torch.Generator(device="cpu").manual_seed(1234)
    p0 = (torch.randn(n, generator=g) * 0.1).to(device)
    gr = torch.Generator(device="cpu").manual_seed(100 + rank)
    grads = [torch.cat([torch.randn(n, generator=gr) * 0.05, torch.randn(8, generator=gr)]).to(device)
             for _ in range(6)]
    peer = PeerGroup(n + 8, device, rank, world)
    local = PeerGroup(n + 8, device)
    z = lambda k, dt=torch.float32: torch.zeros(k, dtype=dt, device=device)
    st = {k: dict(p=p0.clone(), mu=z(n), nu=z(n), c=z(2, torch.int32), gsum=z(n), loss=torch.zeros(6, 5, device=device))
          for k in ("peer", "nccl")}
    for k, grad in enumerate(grads):
        peer.grad.copy_(grad)
        s = st["peer"]
        native.reduce_clip_adam_pair(s["p"], s["mu"], s["nu"], s["c"], peer, s["gsum"], na, nc, None,
                                     None, None, None, 1.0 / world, 2.5e-4, 2.5e-4, 0.5, 0, 1,
                                     s["loss"][k])
        local.grad.copy_(grad)
        dist.all_reduce(local.grad)
        s = st["nccl"]
        native.reduce_clip_adam_pair(s["p"], s["mu"], s["nu"], s["c"], local, s["gsum"], na, nc, None,
                                     None, None, None, 1.0 / world, 2.5e-4, 2.5e-4, 0.5, 0, 1,
                                     s["loss"][k])
    torch.cuda.synchronize(device)
    seq, err = peer.status()
    assert err == 0 and seq == len(grads), (seq, err)
    a, b = st["peer"], st["nccl"]
    if world == 2:
        for name in ("p", "mu", "nu", "c", "loss"):
            assert torch.equal(a[name], b[name]), (rank, name)
    else:
        torch.testing.assert_close(a["p"], b["p"], rtol=1e-5, atol=1e-7)
    ref = a["p"].clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(a["p"], ref), "peer path: parameters differ between ranks"
    peer.release()
    local.release()


def main():
    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.systems.ppo import _runner, ff_mappo
    from mava_b200.utils import make_env

    device = _runner.init_distributed()
    rank, world = dist.get_rank(), dist.get_world_size()
    exact_collective_check(device, rank, world)
    learners = {}
    for coll in ("peer", "nccl"):
        cfg = compose(ff_mappo.CONFIG_NAME, [
            "env/scenario=tiny-4ag", "arch.num_envs=64", "system.update_batch_size=2",
            "system.rollout_length=32", "system.ppo_epochs=2", "system.num_minibatches=2",
            f"+arch.collective={coll}", "logger.use_console=False"])
        env, _ = make_env.make(cfg, add_global_state=True, device=device)
        key, _, ak, ck = prng.split(prng.PRNGKey(11), 4)
        learn, _, state = ff_mappo.learner_setup(env, (key, ak, ck), cfg)
        cfg.system.num_updates_per_eval = 1
        learners[coll] = (learn, state, learn.learner.params.clone())
    for step in range(2):
        for coll in ("peer", "nccl"):
            learn, state, _ = learners[coll]
            learn(state)
    torch.cuda.synchronize(device)
    a, b = learners["peer"][0].learner, learners["nccl"][0].learner
    assert a.collective == "peer" and b.collective == "nccl"
    seq, err = a.peer.status()
    assert err == 0 and seq > 0, (seq, err)
    # The two learners agree to fp32 rounding of the gradient sums only: the loss kernels flush their
    # per-CTA partial gradients with atomics, whose order differs from launch to launch.
    p0 = learners["peer"][2]
    moved = float((a.params - p0).abs().max())
    diff = (a.params - b.params).abs()
    # (a rounding difference can flip a sampled action in a later rollout, after which the two runs
    # see different data: the bars are those of "the same training run", not of rounding)
    assert moved > 1e-4 and float(diff.max()) < 0.5 * moved, (moved, float(diff.max()))
    assert float((diff < 5e-2 * moved).float().mean()) > 0.9
    assert torch.equal(a.key, b.key) and torch.equal(a.counts, b.counts)
    # replicated state is identical on every rank
    mine = a.params.clone()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), "parameters diverged between ranks"
    for coll in ("peer", "nccl"):
        learners[coll][0].learner.release()
    dist.barrier()
    if rank == 0:
        print("PEER_OK world", world, "calls", seq, flush=True)
    # graphs that captured NCCL were dropped by release(); the group can be torn down normally
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
