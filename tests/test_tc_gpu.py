"""tcgen05 operand-convention self-test: each GEMM arrangement against torch on bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,K", [(128, 128), (16, 128), (128, 80), (128, 272), (128, 16), (64, 32)])
def test_tc_gemm_conventions(lib_built, mode, N, K):
    from mava_b200 import native

    g = torch.Generator(device="cpu").manual_seed(mode * 100 + N + K)
    if mode == 0:
        A, B = torch.randn(128, K, generator=g), torch.randn(K, N, generator=g)
        ref = A.bfloat16().float() @ B.bfloat16().float()
    elif mode == 1:
        A, B = torch.randn(128, K, generator=g), torch.randn(N, K, generator=g)
        ref = A.bfloat16().float() @ B.bfloat16().float().T
    else:
        A, B = torch.randn(K, 128, generator=g), torch.randn(K, N, generator=g)
        ref = A.bfloat16().float().T @ B.bfloat16().float()
    D = torch.zeros(128, N, device=DEV)
    native.tc_selftest(mode, A.to(DEV).contiguous(), B.to(DEV).contiguous(), D, N, K)
    torch.cuda.synchronize()
    # bf16 products are exact in fp32; only the accumulation order differs
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("A,FR,N,critic_mode,NE", [
    (4, 66, 5, "global", 300), (2, 66, 5, "agent", 129), (2, 12, 6, "agent", 64)])
def test_act_bf16_matches_fp32_kernel(lib_built, A, FR, N, critic_mode, NE):
    """bf16 tensor-core acting step against the fp32 kernels on the same inputs: values and
    log-probs within the bf16 tolerance (2e-2), replayed actions identical."""
    import numpy as np

    from mava_b200 import native
    from tests.test_mlp_gpu import flat, make_params, random_batch

    rng = np.random.default_rng(0)
    view, mask_bool, mask = random_batch(rng, NE, A, FR, N)
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    cmode = native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW
    critic = native.mlp_desc(cmode, True, A, FR, 128, 128, 1)
    ap = torch.from_numpy(flat(make_params(rng, actor.in_dim, 128, 128, N))).to(DEV)
    cp = torch.from_numpy(flat(make_params(rng, critic.in_dim, 128, 128, 1))).to(DEV)
    tv, tm = torch.from_numpy(view).to(DEV), torch.from_numpy(mask).to(DEV)
    key = torch.from_numpy(np.array([7, 9], np.uint32)).to(DEV)

    a32 = torch.zeros(NE, A, dtype=torch.int8, device=DEV)
    l32, v32 = torch.zeros(NE, A, device=DEV), torch.zeros(NE, A, device=DEV)
    native.ff_act(actor, ap, critic, cp, tv, tm, key, NE, NE, a32, l32, v32)

    ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
    ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
    native.mlp_pack_bf16(actor, ap, ai)
    native.mlp_pack_bf16(critic, cp, ci)
    a16 = torch.zeros(NE, A, dtype=torch.int8, device=DEV)
    l16, v16 = torch.zeros(NE, A, device=DEV), torch.zeros(NE, A, device=DEV)
    # replay the fp32 path's actions so log-probs are comparable row by row
    native.ff_act_bf16(actor, ap, ai, critic, cp, ci, tv, tm, None, NE, NE, a16, l16, v16,
                       actions_in=a32)
    torch.cuda.synchronize()
    assert torch.equal(a16, a32)
    scale = v32.abs().max().item()
    torch.testing.assert_close(v16, v32, rtol=2e-2, atol=2e-2 * scale)
    # log-probs are differences of logits: 2e-2 relative to the logit scale of this random net
    lscale = max(1.0, l32.abs().max().item())
    torch.testing.assert_close(l16, l32, rtol=2e-2, atol=2e-2 * lscale)
    # own sampling: legal actions, and mostly the same as the fp32 path (same noise, close logits)
    native.ff_act_bf16(actor, ap, ai, critic, cp, ci, tv, tm, key, NE, NE, a16, l16, v16)
    torch.cuda.synchronize()
    legal = np.take_along_axis(mask_bool, a16.cpu().numpy()[..., None].astype(np.int64), -1)
    assert legal.all()
    assert (a16 == a32).float().mean().item() > 0.97


@pytest.mark.parametrize("A,FR,N,critic_mode,U,mb,T,E", [
    (4, 66, 5, "global", 2, 96, 12, 40),
    (2, 66, 5, "agent", 1, 200, 12, 40),
    (4, 66, 5, "global", 2, 2048, 32, 256),
])
def test_ppo_loss_grad_bf16_matches_fp32_kernel(lib_built, A, FR, N, critic_mode, U, mb, T, E):
    """Fused tensor-core fwd+bwd against the fp32 kernels on the same minibatch: losses and every
    gradient block within the bf16 tolerance (2e-2 of the block's scale)."""
    import numpy as np

    from mava_b200 import native
    from mava_b200._lib import PpoHyper
    from tests.test_mlp_gpu import flat, make_params, random_batch

    rng = np.random.default_rng(1)
    NE = U * E
    S = T * NE
    view, mask_bool, mask = random_batch(rng, S, A, FR, N)
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    cmode = native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW
    critic = native.mlp_desc(cmode, True, A, FR, 128, 128, 1)
    ap = torch.from_numpy(flat(make_params(rng, actor.in_dim, 128, 128, N, scale_out=0.05))).to(DEV)
    cp = torch.from_numpy(flat(make_params(rng, critic.in_dim, 128, 128, 1, scale_out=0.05))).to(DEV)
    legal = mask_bool.reshape(-1, N)
    action = np.array([rng.choice(np.flatnonzero(r)) for r in legal], np.int8).reshape(S, A)
    old_logp = (-rng.random((S, A)) * 2.0).astype(np.float32)
    old_value = rng.normal(size=(S, A)).astype(np.float32)
    adv = rng.normal(size=(S, A)).astype(np.float32)
    targets = (old_value + rng.normal(size=(S, A)) * 0.5).astype(np.float32)
    perm = torch.from_numpy(rng.permutation(T * E).astype(np.int32)).to(DEV)
    hyper = PpoHyper(0.2, 0.01, 0.5)
    rows = torch.zeros(U * mb, dtype=torch.int32, device=DEV)
    native.ppo_minibatch_rows(perm, 0, mb, U, E, rows)
    dev = lambda x: torch.from_numpy(x).to(DEV)
    args = (dev(view), dev(mask), dev(action), dev(old_logp), dev(old_value), dev(adv), dev(targets))
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)

    g32 = torch.zeros(na + nc + 8, device=DEV)
    ws32 = torch.zeros(native.ppo_workspace_bytes(actor, critic, U * mb), dtype=torch.uint8, device=DEV)
    native.ppo_loss_grad(actor, ap, critic, cp, hyper, *args, rows, U, mb, g32, ws32)

    ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
    ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
    native.mlp_pack_bf16(actor, ap, ai)
    native.mlp_pack_bf16(critic, cp, ci)
    g16 = torch.zeros(na + nc + 8, device=DEV)
    ws16 = torch.zeros(native.ppo_workspace_bytes_bf16(actor, critic, U * mb), dtype=torch.uint8,
                       device=DEV)
    native.ppo_loss_grad_bf16(actor, ap, ai, critic, cp, ci, hyper, *args, rows, U, mb, g16, ws16)
    torch.cuda.synchronize()
    g32, g16 = g32.cpu().numpy(), g16.cpu().numpy()
    np.testing.assert_allclose(g16[na + nc:na + nc + 5], g32[na + nc:na + nc + 5], rtol=2e-2,
                               atol=2e-3)

    def blocks(d, off):
        sizes = [d.in_dim * 128, 128, 128 * 128, 128, 128 * d.out_dim, d.out_dim]
        names = ["w1", "b1", "w2", "b2", "w3", "b3"]
        for n_, s_ in zip(names, sizes):
            yield n_, slice(off, off + s_)
            off += s_

    worst = [0.0, 0.0]
    for net, d, off in (("actor", actor, 0), ("critic", critic, na)):
        for name, sl in blocks(d, off):
            ref, got = g32[sl], g16[sl]
            # tolerance: BASELINE.json allows 2e-2 per bf16 GEMM.  Against the float64 oracle at the
            # headline minibatch size every block is within 1.1e-2 (tests/test_bf16_path_gpu.py,
            # profiles/bf16_grad_errors_r2.json).  Here the reference is the fp32 KERNEL and the
            # minibatches are small, so single clip decisions that differ between the two paths show:
            # measured worst block 0.034 / 0.041 Frobenius (0.051 / 0.094 max) at mb = 96 / 200 and
            # 0.013 (0.015 max) at mb = 2048; bars at <= 1.5x of that.
            scale = np.abs(ref).max() + 1e-12
            err_max = np.abs(got - ref).max() / scale
            err_fro = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12)
            print(f"{net}.{name}: fro {err_fro:.4f} max {err_max:.4f}")
            worst[0], worst[1] = max(worst[0], err_fro), max(worst[1], err_max)
            bar_fro, bar_max = (2e-2, 2.3e-2) if mb >= 2048 else (6e-2, 1.4e-1)
            assert err_fro < bar_fro and err_max < bar_max, (
                f"{net}.{name}: fro err {err_fro:.4f}, max err {err_max:.4f} of scale {scale:.3e}")
    print(f"WORST mb={mb} U={U}: fro {worst[0]:.4f} max {worst[1]:.4f}")


@pytest.mark.gpu
@pytest.mark.parametrize("A,FR,N,critic_mode", [(4, 66, 5, "global"), (2, 12, 6, "agent")])
def test_clip_adam_pair_pack_refreshes_images(lib_built, A, FR, N, critic_mode):
    """The optimiser step that also refreshes the packed bf16 operand images: parameters and
    moments identical to mava_clip_adam_pair, images identical to packing the new parameters."""
    from mava_b200 import native

    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    cmode = native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW
    critic = native.mlp_desc(cmode, True, A, FR, 128, 128, 1)
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
    g = torch.Generator(device="cpu").manual_seed(3)
    p0 = (torch.randn(na + nc, generator=g) * 0.1).to(DEV)
    grad = torch.cat([torch.randn(na + nc, generator=g) * 0.3, torch.zeros(8)]).to(DEV)

    def fresh():
        return (p0.clone(), torch.zeros(na + nc, device=DEV), torch.zeros(na + nc, device=DEV),
                torch.zeros(2, dtype=torch.int32, device=DEV))

    pa, mua, nua, ca = fresh()
    pb, mub, nub, cb = fresh()
    ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
    ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
    native.mlp_pack_bf16(actor, pb[:na], ai)
    native.mlp_pack_bf16(critic, pb[na:], ci)
    for _ in range(3):
        native.clip_adam_pair(pa, mua, nua, ca, grad, na, nc, 0.5, 2.5e-4, 2.5e-4, 0.5, 10, 4)
        native.clip_adam_pair_pack(pb, mub, nub, cb, grad, actor, ai, critic, ci, 0.5, 2.5e-4, 2.5e-4,
                                   0.5, 10, 4)
    torch.cuda.synchronize()
    assert torch.equal(pa, pb) and torch.equal(mua, mub) and torch.equal(nua, nub)
    assert torch.equal(ca, cb)
    ai_ref, ci_ref = torch.zeros_like(ai), torch.zeros_like(ci)
    native.mlp_pack_bf16(actor, pa[:na], ai_ref)
    native.mlp_pack_bf16(critic, pa[na:], ci_ref)
    assert torch.equal(ai, ai_ref) and torch.equal(ci, ci_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("NE", [129, 300, 301, 4133])
def test_value_batch_matches_acting_kernel(lib_built, NE):
    """The persistent batched critic pass (bulk-copied joint observations, weights loaded once per
    CTA) against the critic half of the acting kernel on the same rows: same operands, same MMA
    order -- identical values.  NE covers full tiles, a last tile that arrives by bulk copy and one
    that does not (odd row count: byte count not a multiple of 16)."""
    import numpy as np

    from mava_b200 import native
    from tests.test_mlp_gpu import flat, make_params, random_batch

    A, FR, N = 4, 66, 5
    rng = np.random.default_rng(NE)
    view, mask_bool, mask = random_batch(rng, NE, A, FR, N)
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
    ap = torch.from_numpy(flat(make_params(rng, actor.in_dim, 128, 128, N))).to(DEV)
    cp = torch.from_numpy(flat(make_params(rng, critic.in_dim, 128, 128, 1))).to(DEV)
    ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
    ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
    native.mlp_pack_bf16(actor, ap, ai)
    native.mlp_pack_bf16(critic, cp, ci)
    tv, tm = torch.from_numpy(view).to(DEV), torch.from_numpy(mask).to(DEV)
    key = torch.tensor([1, 2], dtype=torch.uint32, device=DEV)
    act = torch.zeros(NE, A, dtype=torch.int8, device=DEV)
    logp, v_act = torch.zeros(NE, A, device=DEV), torch.zeros(NE, A, device=DEV)
    native.ff_act_bf16(actor, ap, ai, critic, cp, ci, tv, tm, key, NE, NE, act, logp, v_act)
    v_batch = torch.full((NE, A), float("nan"), device=DEV)
    native.ff_act_bf16(None, None, None, critic, cp, ci, tv, None, None, NE, NE, None, None, v_batch)
    torch.cuda.synchronize()
    assert torch.equal(v_batch, v_act)
