"""Run under torchrun with N >= 2 GPUs (tests/test_peer_gpu.py::test_peer_two_gpus, or by hand:
`torchrun --nproc-per-node N tests/peer_multi_gpu.py`).  Each rank builds the ff_mappo learner twice
-- arch.collective=peer (the fused NVLink all-reduce inside the optimiser kernel) and
arch.collective=nccl (dist.all_reduce + the same kernel on the local buffer) -- from the same seed,
runs the same updates with CUDA graphs and compares parameters, optimiser state, losses and rollouts.
World 2: bit-identical (a two-term sum has one order).  World > 2: NCCL's reduction order differs
from rank order, so parameters agree to fp32 rounding of the sums."""
import os
import sys

ROOT = os.environ.get("MAVA_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.systems.ppo import _runner, ff_mappo
    from mava_b200.utils import make_env

    device = _runner.init_distributed()
    rank, world = dist.get_rank(), dist.get_world_size()
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
        learners[coll] = (learn, state)
    for step in range(3):
        for coll in ("peer", "nccl"):
            learn, state = learners[coll]
            learn(state)
    torch.cuda.synchronize(device)
    a, b = learners["peer"][0].learner, learners["nccl"][0].learner
    assert a.collective == "peer" and b.collective == "nccl"
    seq, err = a.peer.status()
    assert err == 0 and seq > 0, (seq, err)
    if world == 2:
        for name in ("params", "mu", "nu", "counts", "key", "loss_buf", "action", "view"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (rank, name)
    else:
        moved = float((a.params - b.params).abs().max())
        assert moved < 2e-5, moved
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
