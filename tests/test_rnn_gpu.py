"""Recurrent (GRU) kernels against the torch-autograd oracle (oracle/ppo.py): one acting step and
the gradients of one rec_mappo / rec_ippo minibatch (float64 oracle, rtol stated per check)."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo

pytestmark = pytest.mark.gpu

F32_MIN = float(np.finfo(np.float32).min)


def _bits(mask_bool):
    return (mask_bool.astype(np.int64) << np.arange(mask_bool.shape[-1])).sum(-1).astype(np.uint8)


def _inputs(mode, view, dense, A, add_id):
    """Network input rows (steps..., rows, in_dim) the way the wrappers build them."""
    from mava_b200 import native

    if mode == native.IN_DENSE:
        return dense.double()
    v = view.double()
    if mode == native.IN_GLOBAL:
        return v.reshape(*v.shape[:-2], 1, -1)
    if add_id:
        eye = torch.eye(A, dtype=torch.float64).expand(*v.shape[:-2], A, A)
        v = torch.cat([eye, v], -1)
    return v


def _make(rng, mode, A, FR, H, Q, out, dense_dim=0, rpe=None, precision=0):
    from mava_b200 import native

    d = native.rnn_desc(mode, True, A, FR, H, Q, out, dense_dim, rpe, precision)
    n = native.rnn_param_count(d)
    # weight scale ~ 1 / sqrt(fan-in) relative to the H = 16 cases (0.3): keeps the pre-activations of
    # the wide nets out of saturation, where a bf16 rounding of h flips gates
    flat = (rng.standard_normal(n) * 0.3 * min(1.0, (16.0 / H) ** 0.5)).astype(np.float32)
    return d, flat


@pytest.mark.parametrize("critic_mode", ["global", "agent", "dense"])
def test_rec_act_matches_oracle(lib_built, critic_mode):
    from mava_b200 import native

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    NE, A, FR, H, Q, N, U = 37, 3, 11, 32, 24, 5, 1
    dense_dim = 9
    if critic_mode == "dense":
        ad, ap = _make(rng, native.IN_DENSE, A, FR, H, Q, N, dense_dim, A)
        cd, cp = _make(rng, native.IN_DENSE, A, FR, H, Q, 1, dense_dim + 4, 1)
    else:
        ad, ap = _make(rng, native.IN_AGENT_VIEW, A, FR, H, Q, N)
        cd, cp = _make(rng, native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW,
                       A, FR, H, Q, 1)
    view = torch.from_numpy(rng.integers(-3, 4, (NE, A, FR)).astype(np.int8))
    oa = torch.from_numpy(rng.standard_normal((NE, A, dense_dim)).astype(np.float32))
    oc = torch.from_numpy(rng.standard_normal((NE, 1, dense_dim + 4)).astype(np.float32))
    mask_b = rng.random((NE, A, N)) < 0.7
    mask_b[..., 0] = True
    done = rng.random(NE) < 0.3
    ha = torch.from_numpy(rng.standard_normal((NE * A, H)).astype(np.float32))
    hc = torch.from_numpy(rng.standard_normal((NE * cd.rows_per_env, H)).astype(np.float32))
    acts = torch.from_numpy(rng.integers(0, N, (NE, A)).astype(np.int8))
    acts = torch.where(torch.from_numpy(np.take_along_axis(mask_b, acts.numpy()[..., None].astype(np.int64), -1)[..., 0]),
                       acts, torch.zeros_like(acts))

    g = lambda t: t.to(dev)
    ws = torch.zeros(native.rec_act_workspace_bytes(ad, cd, NE), dtype=torch.uint8, device=dev)
    action = torch.zeros(NE, A, dtype=torch.int8, device=dev)
    logp = torch.zeros(NE, A, device=dev)
    value = torch.zeros(NE, A, device=dev)
    ha_out, hc_out = torch.zeros_like(ha, device=dev), torch.zeros_like(hc, device=dev)
    key = torch.tensor([1, 2], dtype=torch.uint32, device=dev)
    native.rec_act(ad, g(torch.from_numpy(ap)), cd, g(torch.from_numpy(cp)), g(view), g(oa), g(oc),
                   g(torch.from_numpy(_bits(mask_b))), g(torch.from_numpy(done.astype(np.uint8))),
                   g(ha), ha_out, g(hc), hc_out, key, NE // U, NE, action, logp, value, ws,
                   actions_in=g(acts))
    torch.cuda.synchronize()

    pa = oppo.rnn_unflatten(torch.from_numpy(ap).double(), ad.in_dim, H, Q, N)
    pc = oppo.rnn_unflatten(torch.from_numpy(cp).double(), cd.in_dim, H, Q, 1)
    xa = _inputs(ad.input_mode, view, oa, A, True).reshape(1, NE * A, -1)
    xc = _inputs(cd.input_mode, view, oc, A, True).reshape(1, NE * cd.rows_per_env, -1)
    ra = torch.from_numpy(np.repeat(done, A)).reshape(1, -1)
    rc = torch.from_numpy(np.repeat(done, cd.rows_per_env)).reshape(1, -1)
    h1a, logits = oppo.rec_net(pa, ha.double(), xa, ra)
    h1c, val = oppo.rec_net(pc, hc.double(), xc, rc)
    logits = torch.where(torch.from_numpy(mask_b).reshape(1, NE * A, N), logits,
                         torch.full_like(logits, F32_MIN))
    lp = oppo.categorical_log_prob(logits, acts.reshape(1, -1))
    np.testing.assert_allclose(ha_out.cpu().numpy(), h1a.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(hc_out.cpu().numpy(), h1c.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(action.cpu().numpy(), acts.numpy())
    np.testing.assert_allclose(logp.cpu().numpy().reshape(-1), lp.numpy().reshape(-1), rtol=1e-4,
                               atol=1e-5)
    v = val.reshape(NE, cd.rows_per_env).numpy()
    v = np.repeat(v, A, 1) if cd.rows_per_env == 1 else v
    np.testing.assert_allclose(value.cpu().numpy(), v, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("critic_mode,chunk,precision,H", [
    ("global", 8, 0, 16), ("agent", 4, 0, 16), ("dense", 2, 0, 16), ("global", 8, 1, 16),
    ("dense", 4, 1, 16),
    # hidden width 128 + bf16: the persistent GRU scan kernels (csrc/gru_scan.cu) carry the time loop
    ("global", 8, 1, 128), ("dense", 4, 1, 128), ("agent", 2, 1, 128)])
def test_rec_ppo_loss_grad_matches_autograd(lib_built, critic_mode, chunk, precision, H):
    from mava_b200 import native
    from mava_b200._lib import PpoHyper

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    T, E, U, A, FR, Q, N, nmb = 8, 6, 2, 3, 7, 24, 5, 2
    NE, nc = U * E, T // chunk
    dense_dim = 10
    if critic_mode == "dense":
        ad, ap = _make(rng, native.IN_DENSE, A, FR, H, Q, N, dense_dim, A, precision)
        cd, cp = _make(rng, native.IN_DENSE, A, FR, H, Q, 1, dense_dim + 3, 1, precision)
    else:
        ad, ap = _make(rng, native.IN_AGENT_VIEW, A, FR, H, Q, N, precision=precision)
        cd, cp = _make(rng, native.IN_GLOBAL if critic_mode == "global" else native.IN_AGENT_VIEW,
                       A, FR, H, Q, 1, precision=precision)
    rpc = cd.rows_per_env
    f32 = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    view = torch.from_numpy(rng.integers(-2, 3, (T, NE, A, FR)).astype(np.int8))
    oa, oc = f32(T, NE, A, dense_dim), f32(T, NE, rpc, dense_dim + 3)
    mask_b = rng.random((T, NE, A, N)) < 0.7
    mask_b[..., 0] = True
    action = torch.from_numpy(rng.integers(0, N, (T, NE, A)).astype(np.int8))
    action = torch.where(torch.from_numpy(np.take_along_axis(mask_b, action.numpy()[..., None].astype(np.int64), -1)[..., 0]),
                         action, torch.zeros_like(action))
    # (own seeded generator: the global torch generator differs from process to process, and with it
    # which ratios sit next to a clip boundary -- the bf16 cases then flipped a few clip decisions
    # in some runs and not in others)
    old_logp = -torch.rand(T, NE, A, generator=torch.Generator().manual_seed(7)) * 2 - 0.3
    old_value, adv, targets = f32(T, NE, A), f32(T, NE, A), f32(T, NE, A)
    done = rng.random((T, NE)) < 0.25
    hs_a, hs_c = f32(T, NE * A, H), f32(T, NE * rpc, H)  # hidden entering every step (oracle view)
    mb_cols = E * nc // nmb
    perm = torch.from_numpy(rng.permutation(E * nc).astype(np.int32))
    cols = perm[mb_cols:2 * mb_cols].contiguous()
    hyper = PpoHyper(0.2, 0.01, 0.5)

    g = lambda t: t.to(dev)
    na, ncr = native.rnn_param_count(ad), native.rnn_param_count(cd)
    grad = torch.zeros(na + ncr + 8, device=dev)
    ws = torch.zeros(native.rec_ppo_workspace_bytes(ad, cd, U * mb_cols, chunk), dtype=torch.uint8,
                     device=dev)
    native.rec_ppo_loss_grad(
        ad, g(torch.from_numpy(ap)), cd, g(torch.from_numpy(cp)), hyper, g(view), g(oa), g(oc),
        g(torch.from_numpy(_bits(mask_b))), g(action), g(old_logp), g(old_value), g(adv), g(targets),
        g(torch.from_numpy(done.astype(np.uint8))), g(hs_a[:nc].contiguous()),
        g(hs_c[:nc].contiguous()), g(cols), U, E, mb_cols, chunk, nc, grad, ws)
    torch.cuda.synchronize()
    grad = grad.cpu().numpy()

    # ---- oracle: per replica exactly what _update_minibatch computes, then pmean over "batch"
    pa_flat = torch.from_numpy(ap).double().requires_grad_(True)
    pc_flat = torch.from_numpy(cp).double().requires_grad_(True)
    pa = oppo.rnn_unflatten(pa_flat, ad.in_dim, H, Q, N)
    pc = oppo.rnn_unflatten(pc_flat, cd.in_dim, H, Q, 1)
    tot_a = tot_c = 0.0
    la = ent = vl = 0.0
    xa_all = _inputs(ad.input_mode, view, oa, A, True)     # (T, NE, A, in)
    xc_all = _inputs(cd.input_mode, view, oc, A, True)     # (T, NE, rpc, in)
    for u in range(U):
        sl = slice(u * E, (u + 1) * E)
        cb = lambda x: oppo.rec_chunk_batch(x[:, sl], chunk, cols)
        xa, xc = cb(xa_all), cb(xc_all)                      # (chunk, mb, rows, in)
        dn = cb(torch.from_numpy(done))                      # (chunk, mb)
        mb = xa.shape[1]
        h0a = cb(hs_a.reshape(T, NE, A, H).double())[0]
        h0c = cb(hs_c.reshape(T, NE, rpc, H).double())[0]
        _, logits = oppo.rec_net(pa, h0a.reshape(mb * A, H), xa.reshape(chunk, mb * A, -1),
                                 dn[:, :, None].expand(chunk, mb, A).reshape(chunk, -1))
        _, val = oppo.rec_net(pc, h0c.reshape(mb * rpc, H), xc.reshape(chunk, mb * rpc, -1),
                              dn[:, :, None].expand(chunk, mb, rpc).reshape(chunk, -1))
        logits = logits.reshape(chunk, mb, A, N)
        logits = torch.where(cb(torch.from_numpy(mask_b)), logits, torch.full_like(logits, F32_MIN))
        val = val.reshape(chunk, mb, rpc)
        val = val.expand(chunk, mb, A) if rpc == 1 else val
        t_a, l_a, e_a = oppo.actor_loss(logits, cb(action), cb(old_logp).double(), cb(adv).double(),
                                        hyper.clip_eps, hyper.ent_coef)
        t_c, v_l = oppo.critic_loss(val, cb(old_value).double(), cb(targets).double(),
                                    hyper.clip_eps, hyper.vf_coef)
        tot_a, tot_c = tot_a + t_a / U, tot_c + t_c / U
        la, ent, vl = la + l_a.item() / U, ent + e_a.item() / U, vl + v_l.item() / U
    ga, = torch.autograd.grad(tot_a, pa_flat)
    gc, = torch.autograd.grad(tot_c, pc_flat)
    scale_a, scale_c = float(ga.abs().max()), float(gc.abs().max())
    if precision == 1:
        # bf16 tensor-core contractions (fp32 accumulation): 2e-2 tolerance of BASELINE.json on the
        # losses, gradient blocks within 5e-2 in Frobenius norm (chains of 3 to 5 bf16 GEMMs).  At
        # hidden width 128 every recurrent contraction sums 128 bf16 products per step and the chain
        # runs through up to 8 steps: measured 0.054 ... 0.085 on these random networks (the scan
        # kernels and the per-step schedule give bit-identical numbers), bar 0.13.
        bar = 5e-2 if H <= 16 else 0.13
        for got, want in ((grad[:na], ga.numpy()), (grad[na:na + ncr], gc.numpy())):
            err = np.linalg.norm(got - want) / np.linalg.norm(want)
            print(f"rec bf16 grad error H={H} {critic_mode}: {err:.4f}")
            assert err < bar, err
        np.testing.assert_allclose(grad[na + ncr:na + ncr + 5],
                                   [tot_a.item(), la, ent, tot_c.item(), vl], rtol=2e-2, atol=2e-3)
        return
    np.testing.assert_allclose(grad[:na], ga.numpy(), rtol=2e-4, atol=2e-5 * scale_a)
    np.testing.assert_allclose(grad[na:na + ncr], gc.numpy(), rtol=2e-4, atol=2e-5 * scale_c)
    np.testing.assert_allclose(grad[na + ncr:na + ncr + 5],
                               [tot_a.item(), la, ent, tot_c.item(), vl], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("E,A,chunk,T", [(44, 3, 8, 16), (64, 4, 16, 16), (20, 8, 4, 8)])
def test_gru_scan_matches_per_step_schedule(lib_built, E, A, chunk, T, monkeypatch):
    """The persistent scan kernels against the per-time-step launches they replace (same bf16
    contractions, same fp32 gate arithmetic): losses and every gradient agree to rounding.  Sizes
    cover several CTAs and a last tile of sequences that is only partly filled; a quarter of the
    steps reset the hidden state."""
    from mava_b200 import native
    from mava_b200._lib import PpoHyper

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(E + A)
    U, FR, Hd, Q, N, nmb, dense_dim = 2, 9, 128, 128, 6, 2, 21
    NE, nc = U * E, T // chunk
    ad, ap = _make(rng, native.IN_DENSE, A, FR, Hd, Q, N, dense_dim, A, 1)
    cd, cp = _make(rng, native.IN_DENSE, A, FR, Hd, Q, 1, dense_dim + 3, 1, 1)
    f32 = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(dev)
    oa, oc = f32(T, NE, A, dense_dim), f32(T, NE, 1, dense_dim + 3)
    mask = torch.full((T, NE, A), (1 << N) - 1, dtype=torch.uint8, device=dev)
    action = torch.from_numpy(rng.integers(0, N, (T, NE, A)).astype(np.int8)).to(dev)
    old_logp = (-torch.rand(T, NE, A, generator=torch.Generator().manual_seed(3)) * 2 - 0.3).to(dev)
    old_value, adv, targets = f32(T, NE, A), f32(T, NE, A), f32(T, NE, A)
    done = torch.from_numpy((rng.random((T, NE)) < 0.25).astype(np.uint8)).to(dev)
    hs_a, hs_c = f32(nc, NE * A, Hd), f32(nc, NE, Hd)
    mb_cols = E * nc // nmb
    cols = torch.from_numpy(rng.permutation(E * nc).astype(np.int32))[:mb_cols].contiguous().to(dev)
    hyper = PpoHyper(0.2, 0.01, 0.5)
    na, ncr = native.rnn_param_count(ad), native.rnn_param_count(cd)
    tap, tcp = torch.from_numpy(ap).to(dev), torch.from_numpy(cp).to(dev)

    def run():
        grad = torch.zeros(na + ncr + 8, device=dev)
        ws = torch.zeros(native.rec_ppo_workspace_bytes(ad, cd, U * mb_cols, chunk),
                         dtype=torch.uint8, device=dev)
        native.rec_ppo_loss_grad(ad, tap, cd, tcp, hyper, None, oa, oc, mask, action, old_logp,
                                 old_value, adv, targets, done, hs_a, hs_c, cols, U, E, mb_cols, chunk,
                                 nc, grad, ws)
        torch.cuda.synchronize()
        return grad.cpu().numpy()

    monkeypatch.delenv("MAVA_NO_GRU_SCAN", raising=False)
    scan = run()
    monkeypatch.setenv("MAVA_NO_GRU_SCAN", "1")
    steps = run()
    np.testing.assert_allclose(scan[na + ncr:na + ncr + 5], steps[na + ncr:na + ncr + 5], rtol=1e-5)
    for sl in (slice(0, na), slice(na, na + ncr)):
        scale = np.abs(steps[sl]).max()
        np.testing.assert_allclose(scan[sl], steps[sl], rtol=1e-3, atol=1e-5 * scale)
