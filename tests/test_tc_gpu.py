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
