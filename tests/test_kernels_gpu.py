"""GPU parity for PRNG, GAE, minibatch rows and the fused clip+Adam step (via the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo
from oracle import threefry as tf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _u32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32)).to(DEV)


@pytest.mark.parametrize("n", [1, 2, 5, 110, 257, 4097])
def test_random_bits_bit_exact(lib_built, n):
    from mava_b200 import native

    key = tf.prng_key(1234)
    out = torch.zeros(n, dtype=torch.uint32, device=DEV)
    native.prng_random_bits(_u32(key), out, n)
    np.testing.assert_array_equal(out.cpu().numpy(), tf.random_bits(key, (n,)))


def test_split_and_chain_bit_exact(lib_built):
    from mava_b200 import native

    key = tf.prng_key(42)
    out = torch.zeros(33, 2, dtype=torch.uint32, device=DEV)
    native.prng_split(_u32(key), out, 33)
    np.testing.assert_array_equal(out.cpu().numpy(), tf.split(key, 33))
    kio = _u32(key)
    subs = torch.zeros(17, 2, dtype=torch.uint32, device=DEV)
    native.prng_split_chain(kio, subs, 17)
    k = key
    for i in range(17):
        k, s = tf.split(k)
        np.testing.assert_array_equal(subs[i].cpu().numpy(), s)
    np.testing.assert_array_equal(kio.cpu().numpy(), k)


@pytest.mark.parametrize("rec", [False, True])
# (small problems take the staged kernel -- 128-step chunks of 32 env-agents through shared memory --,
#  more than 2^16 env-agents the one-thread-per-env-agent kernel)
@pytest.mark.parametrize("T,NE,A", [(128, 33, 4), (5, 7, 2), (37, 129, 3), (300, 50, 3),
                                    (128, 2048, 4), (16, 40000, 2)])
def test_gae_matches_oracle(lib_built, rec, T, NE, A):
    from mava_b200 import native

    rng = np.random.default_rng(0)
    reward = rng.normal(size=(T, NE, A)).astype(np.float32)
    value = rng.normal(size=(T, NE, A)).astype(np.float32)
    done = (rng.random((T, NE)) < 0.1)
    last_val = rng.normal(size=(NE, A)).astype(np.float32)
    last_done = rng.random(NE) < 0.2
    done_a = np.repeat(done[:, :, None], A, 2)
    if rec:
        oadv, otgt = oppo.gae_rec(reward, value, done_a, last_val, np.repeat(last_done[:, None], A, 1),
                                  0.99, 0.95)
    else:
        oadv, otgt = oppo.gae_ff(reward, value, done_a, last_val, 0.99, 0.95)
    adv = torch.zeros(T, NE, A, device=DEV)
    tgt = torch.zeros(T, NE, A, device=DEV)
    native.gae(torch.from_numpy(reward).to(DEV), torch.from_numpy(value).to(DEV),
               torch.from_numpy(done.astype(np.uint8)).to(DEV), torch.from_numpy(last_val).to(DEV),
               0.99, 0.95, T, NE, A, adv, tgt,
               last_done=torch.from_numpy(last_done.astype(np.uint8)).to(DEV) if rec else None)
    # tolerance: rtol 1e-5 fp32 (BASELINE.json north_star), atol for cancellations near zero
    np.testing.assert_allclose(adv.cpu().numpy(), oadv, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(tgt.cpu().numpy(), otgt, rtol=1e-5, atol=2e-5)


def test_minibatch_rows(lib_built):
    from mava_b200 import native

    T, E, U, nmb = 8, 6, 2, 2
    perm = np.random.default_rng(1).permutation(T * E).astype(np.int32)
    mb = T * E // nmb
    for m in range(nmb):
        rows = torch.zeros(U * mb, dtype=torch.int32, device=DEV)
        native.ppo_minibatch_rows(torch.from_numpy(perm).to(DEV), m, mb, U, E, rows)
        exp = []
        for u in range(U):
            for j in range(mb):
                t, e = divmod(int(perm[m * mb + j]), E)
                exp.append(t * U * E + u * E + e)
        np.testing.assert_array_equal(rows.cpu().numpy(), np.array(exp, np.int32))


@pytest.mark.parametrize("gscale,decay", [(1.0, 0), (0.25, 0), (1.0, 10)])
def test_clip_adam_matches_oracle(lib_built, gscale, decay):
    from mava_b200 import native

    rng = np.random.default_rng(5)
    n = 26245
    p = rng.normal(size=n).astype(np.float32)
    mu = np.zeros(n, np.float32)
    nu = np.zeros(n, np.float32)
    tp, tmu, tnu = (torch.from_numpy(x.copy()).to(DEV) for x in (p, mu, nu))
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    for step in range(12):
        scale = 0.001 if step % 3 == 0 else 1.0  # below and above the clip threshold
        g = (rng.normal(size=n) * scale).astype(np.float32)
        lr = 2.5e-4
        if decay:
            lr = oppo.linear_lr(2.5e-4, step, 2, 2, decay)
        p, mu, nu = oppo.clip_adam(p, g * np.float32(gscale), mu, nu, step, lr, 0.5)
        native.clip_adam(tp, tmu, tnu, cnt, torch.from_numpy(g).to(DEV), n, gscale, 2.5e-4, 0.5,
                         lr_decay_num_updates=decay, steps_per_update=4)
    assert int(cnt.item()) == 12
    np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tmu.cpu().numpy(), mu, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(tnu.cpu().numpy(), nu, rtol=1e-5, atol=1e-10)


def test_clip_adam_pair_matches_single(lib_built):
    """The two-network multi-CTA Adam launch equals two single-network steps."""
    from mava_b200 import native

    rng = np.random.default_rng(9)
    na, nc = 26245, 50561
    p = torch.from_numpy(rng.normal(size=na + nc).astype(np.float32)).to(DEV)
    p2 = p.clone()
    mu, nu = torch.zeros_like(p), torch.zeros_like(p)
    mu2, nu2 = torch.zeros_like(p), torch.zeros_like(p)
    cnt = torch.zeros(2, dtype=torch.int32, device=DEV)
    cnt2 = torch.zeros(2, dtype=torch.int32, device=DEV)
    for step in range(5):
        g = torch.from_numpy((rng.normal(size=na + nc) * (0.001 if step % 2 else 1.0)).astype(np.float32)).to(DEV)
        native.clip_adam_pair(p, mu, nu, cnt, g, na, nc, 0.5, 2.5e-4, 1e-4, 0.5, 7, 4)
        native.clip_adam(p2[:na], mu2[:na], nu2[:na], cnt2[0:1], g[:na], na, 0.5, 2.5e-4, 0.5, 7, 4)
        native.clip_adam(p2[na:], mu2[na:], nu2[na:], cnt2[1:2], g[na:], nc, 0.5, 1e-4, 0.5, 7, 4)
    torch.cuda.synchronize()
    assert cnt.tolist() == [5, 5]
    torch.testing.assert_close(p, p2, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(mu, mu2, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(nu, nu2, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("n", [1, 37, 1000, 16384, 131072, 1 << 20, 4194304])
def test_sort_by_key_matches_stable_sort(lib_built, n):
    """One round of jax.random.permutation's sort_key_val: stable on ties, any size."""
    from mava_b200 import native

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(n)
    keys = torch.randint(0, 2 ** 32, (n,), generator=g, device=dev, dtype=torch.int64)
    if n > 8:  # equal keys in distant positions: stability decides
        idx = torch.randint(0, n, (min(n // 4, 5000),), generator=g, device=dev)
        keys[idx] = keys[(idx * 7 + 3) % n]
    vals = torch.randperm(n, generator=g, device=dev).to(torch.int32)
    out = torch.empty_like(vals)
    ws = torch.zeros(native.sort_workspace_bytes(n), dtype=torch.uint8, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    native.sort_by_key(keys.to(torch.uint32), vals, out, n, ws, ovf)
    order = torch.sort(keys, stable=True).indices
    assert int(ovf.item()) == 0
    assert torch.equal(out, vals[order])


@pytest.mark.gpu
@pytest.mark.parametrize("n,p_done", [(1, 1.0), (1000, 0.0), (4096, 0.05), (262144, 0.3)])
def test_episode_stats_matches_get_final_step_metrics(lib_built, n, p_done):
    """mava_episode_stats against get_final_step_metrics + describe() on the same arrays
    (mava/wrappers/episode_metrics.py:114-132, mava/utils/logger.py:44-58): counts and min / max
    exact, sums to fp64 round-off; empty selections leave the initial values."""
    import numpy as np

    from mava_b200 import native

    rng = np.random.default_rng(n)
    done = (rng.random(n) < p_done).astype(np.uint8)
    ret = rng.normal(size=n).astype(np.float32) * 3
    length = rng.integers(1, 500, size=n).astype(np.int32)
    stats = torch.zeros(10, dtype=torch.float64, device="cuda")
    dev = lambda x: torch.from_numpy(x).cuda()
    native.episode_stats(None, None, None, 0, True, stats)
    half = n // 2  # two accumulating calls, like two updates of one evaluation interval
    d, r, l = dev(done), dev(ret), dev(length)
    if half:
        native.episode_stats(d[:half].contiguous(), r[:half].contiguous(), l[:half].contiguous(), half,
                             False, stats)
    native.episode_stats(d[half:].contiguous(), r[half:].contiguous(), l[half:].contiguous(), n - half,
                         False, stats)
    got = stats.cpu().numpy()
    sel = done.astype(bool)
    assert got[0] == sel.sum()
    if sel.any():
        rs, ls = ret[sel].astype(np.float64), length[sel].astype(np.float64)
        np.testing.assert_allclose(got[1:3], [rs.sum(), (rs * rs).sum()], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(got[5:7], [ls.sum(), (ls * ls).sum()], rtol=1e-12)
        assert got[3] == rs.min() and got[4] == rs.max() and got[7] == ls.min() and got[8] == ls.max()
    else:
        assert got[1] == 0 and np.isposinf(got[3]) and np.isneginf(got[4])
