"""pmean("device") (ff_mappo.py:224-238,392-403): an N-rank update equals the 1-rank update.

Two learners are built in ONE process as rank 0 and rank 1 of a world of 2 (`rank_world`): each takes
its block of the env keys exactly like the reference's pmap split (`ff_mappo.py:392-403`), both hold
the same parameters and the same step key (`:417-426`).  They are driven in lockstep and their
gradient buffers are summed by hand where the job would all-reduce (no NCCL needed, so this runs in
the 1-GPU `pytest -m gpu`).  The reference computes `pmean(pmean(grads, "batch"), "device")`, i.e.
the mean over U x D replicas -- the same as ONE learner with `update_batch_size = 2U` on all the
env keys, whose replicas sit side by side on the env axis.  Checked:

* both ranks end with bit-identical parameters, optimiser state and key (same summed gradients);
* their rollouts are, bit for bit, the two halves of the single learner's rollout;
* their parameters equal the single learner's to fp32 rounding (the gradient sums associate
  differently: per-CTA TMEM partial sums + atomics vs. the hand sum of two buffers).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(system, overrides, rank_world, seed=5):
    import importlib

    from mava_b200 import prng
    from mava_b200.config import compose
    from mava_b200.systems.ppo import anakin
    from mava_b200.utils import make_env

    mod = importlib.import_module(f"mava_b200.systems.ppo.{system}")
    cfg = compose(mod.CONFIG_NAME, overrides + ["+arch.use_cuda_graph=False",
                                                "logger.use_console=False"])
    env, _ = make_env.make(cfg, add_global_state=mod.CENTRALISED_CRITIC)
    key, _, ak, ck = prng.split(prng.PRNGKey(seed), 4)
    learn, _, _ = anakin.learner_setup(env, (key, ak, ck), cfg, mod.CENTRALISED_CRITIC,
                                       rank_world=rank_world)
    return learn.learner


@pytest.mark.parametrize("system,precision,E,T", [
    ("ff_mappo", "fp32", 16, 16), ("ff_mappo", "bf16", 256, 32), ("ff_ippo", "bf16", 64, 16)])
def test_two_rank_update_equals_one_rank_update(lib_built, system, precision, E, T):
    torch.cuda.set_device(0)
    U, updates = 1, 1  # after one update the parameters agree to rounding only, not bit for bit
    base = ["env/scenario=tiny-4ag", f"arch.num_envs={E}", f"system.rollout_length={T}",
            "system.ppo_epochs=2", "system.num_minibatches=2", "env.kwargs.time_limit=20",
            f"+arch.precision={precision}"]
    ranks = [_make(system, base + [f"system.update_batch_size={U}"], (r, 2)) for r in range(2)]
    single = _make(system, base + [f"system.update_batch_size={2 * U}"], None)
    assert single.world == 1 and all(L.world == 2 for L in ranks)
    # same parameters, same step key; rank r holds env-key block r
    for L in ranks:
        assert torch.equal(L.params, single.params) and torch.equal(L.key, single.key)
    NEr = ranks[0].NE
    for r, L in enumerate(ranks):
        assert torch.equal(L.view[0], single.view[0][r * NEr:(r + 1) * NEr])
    p0 = single.params.clone()

    for _ in range(updates):
        single._update_step()
        perms = [L._rollout_and_gae() for L in ranks]
        for L in ranks:
            L._epochs_begin()
        for ep in range(single.epochs):
            for m in range(single.nmb):
                for L, pm in zip(ranks, perms):
                    L._minibatch_grad(ep, m, pm)
                total = ranks[0].grad + ranks[1].grad            # the all-reduce, by hand
                for L in ranks:
                    L.grad.copy_(total)
                    L._minibatch_apply(ep, m)
        for L in ranks:
            L._epochs_end()
            L._carry_over()
        torch.cuda.synchronize()
        # rollouts: the two halves of the single learner's rollout (replica u == rank u)
        for r, L in enumerate(ranks):
            sl = slice(r * NEr, (r + 1) * NEr)
            assert torch.equal(L.action, single.action[:, sl])
            assert torch.equal(L.view, single.view[:, sl])
            assert torch.equal(L.reward, single.reward[:, sl])
            assert torch.equal(L.done, single.done[:, sl])
            assert torch.equal(L.logp, single.logp[:, sl])
            assert torch.equal(L.adv, single.adv[:, sl])

    a, b = ranks
    assert torch.equal(a.params, b.params) and torch.equal(a.mu, b.mu) and torch.equal(a.nu, b.nu)
    assert torch.equal(a.key, b.key) and torch.equal(a.counts, b.counts)
    assert torch.equal(a.key, single.key) and torch.equal(a.counts, single.counts)
    got, want = a.params.cpu().numpy(), single.params.cpu().numpy()
    moved = np.abs(want - p0.cpu().numpy()).max()
    assert moved > 1e-4
    # fp32 rounding in near-zero gradient elements can flip an Adam step (|step| ~ lr): compare on
    # the scale of the total movement, like the oracle comparison in test_learner_gpu.py
    np.testing.assert_allclose(got, want, rtol=0, atol=0.05 * moved)
    assert np.mean(np.abs(got - want) < 1e-3 * moved) > 0.99
    # loss metrics: mean over the two ranks == the single learner's mean over its two replicas
    torch.testing.assert_close(a.loss_buf, single.loss_buf, rtol=1e-4, atol=1e-6)
    assert torch.equal(a.loss_buf, b.loss_buf)
