"""The fused pmean("device") + clip + Adam kernel (csrc/peer.cu, ff_mappo.py:224-250).

* world = 1: against the oracle's optax restatement (oracle/ppo.py::clip_adam) and against the
  two-launch kernels it replaces; the refreshed bf16 images equal a fresh packing;
* world = 2 / 4 / 8 inside ONE process: W exchange buffers on one GPU, W launches on W streams that
  handshake with each other through the flag blocks exactly like W ranks over NVLink -- all "ranks"
  end with bit-identical parameters equal to the oracle step on the rank-ordered sum, several calls
  in a row with the buffers overwritten right after each call (the read-done handshake);
* 2 real GPUs (skipped on a 1-GPU box): tests/peer_multi_gpu.py under torchrun, peer path against
  the NCCL path bit for bit.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ppo as oppo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nets(A=4, FR=66, N=5):
    from mava_b200 import native

    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
    return actor, critic, native.mlp_param_count(actor), native.mlp_param_count(critic)


def _oracle_steps(p0, grads_per_call, na, scale, lrs, max_norm, decay, spu):
    """optax.chain(clip_by_global_norm, adam) per network on the rank-ordered sums."""
    p, mu, nu = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for c, gs in enumerate(grads_per_call):
        g = gs[0].copy()
        for x in gs[1:]:
            g = g + x  # float32, rank order
        g = (g * np.float32(scale)).astype(np.float32)
        for sl, lr in ((slice(0, na), lrs[0]), (slice(na, None), lrs[1])):
            step_lr = oppo.linear_lr(lr, c, spu, 1, decay) if decay else lr
            p[sl], mu[sl], nu[sl] = oppo.clip_adam(p[sl], g[sl], mu[sl], nu[sl], c, step_lr, max_norm)
    return p, mu, nu


@pytest.mark.parametrize("big_grads", [False, True])
def test_reduce_clip_adam_world1(lib_built, big_grads):
    from mava_b200 import native
    from mava_b200.peer import PeerGroup

    actor, critic, na, nc = _nets()
    n = na + nc
    g = torch.Generator(device="cpu").manual_seed(7)
    p0 = (torch.randn(n, generator=g) * 0.1)
    # big_grads: global norms above max_norm (the clip branch); else below it
    gscale = 0.3 if big_grads else 1e-4
    grads = [torch.cat([torch.randn(n, generator=g) * gscale, torch.randn(8, generator=g)])
             for _ in range(3)]
    grp = PeerGroup(n + 8, DEV)
    params, mu, nu = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    counts = torch.zeros(2, dtype=torch.int32, device=DEV)
    gsum = torch.zeros(n, device=DEV)
    ai = torch.zeros(native.mlp_pack_bytes(actor), dtype=torch.uint8, device=DEV)
    ci = torch.zeros(native.mlp_pack_bytes(critic), dtype=torch.uint8, device=DEV)
    native.mlp_pack_bf16(actor, params[:na], ai)
    native.mlp_pack_bf16(critic, params[na:], ci)
    loss = torch.zeros(3, 5, device=DEV)
    # the kernels it replaces
    pb, mub, nub = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    cb = torch.zeros(2, dtype=torch.int32, device=DEV)
    for k, gr in enumerate(grads):
        grp.grad.copy_(gr.to(DEV))
        native.reduce_clip_adam_pair(params, mu, nu, counts, grp, gsum, na, nc, actor, ai, critic, ci,
                                     0.5, 2.5e-4, 3e-4, 0.5, 10, 2, loss[k])
        native.clip_adam_pair(pb, mub, nub, cb, gr.to(DEV), na, nc, 0.5, 2.5e-4, 3e-4, 0.5, 10, 2)
    torch.cuda.synchronize()
    want, wmu, wnu = _oracle_steps(p0.numpy(), [[x.numpy()[:n]] for x in grads], na, 0.5,
                                   (2.5e-4, 3e-4), 0.5, 10, 2)
    np.testing.assert_allclose(params.cpu().numpy(), want, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mu.cpu().numpy(), wmu, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(nu.cpu().numpy(), wnu, rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(params, pb, rtol=1e-6, atol=1e-8)
    assert counts.tolist() == [3, 3] and cb.tolist() == [3, 3]
    for k, gr in enumerate(grads):  # loss scalars: the 5 floats behind the gradients x 1/world
        torch.testing.assert_close(loss[k].cpu(), gr[n:n + 5] * 0.5, rtol=1e-6, atol=0)
    ai_ref, ci_ref = torch.zeros_like(ai), torch.zeros_like(ci)
    native.mlp_pack_bf16(actor, params[:na], ai_ref)
    native.mlp_pack_bf16(critic, params[na:], ci_ref)
    assert torch.equal(ai, ai_ref) and torch.equal(ci, ci_ref)
    grp.release()


@pytest.mark.parametrize("pull", [0, 1], ids=["push", "pull"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_reduce_clip_adam_ranks_as_streams(lib_built, world, pull, monkeypatch):
    from mava_b200 import native
    from mava_b200.peer import PeerGroup

    # both exchange protocols of csrc/peer.cu: lines pushed into the peers' receive areas (default)
    # and the flag handshake + peer reads it is checked against
    monkeypatch.setenv("MAVA_PEER_PULL", str(pull))

    # world 8: eight cooperative grids must be co-resident on ONE GPU here (register file: 592 CTAs
    # of this kernel), so the eight-rank case runs on half-size vectors (2 agents, 32 features)
    actor, critic, na, nc = _nets() if world < 8 else _nets(A=2, FR=32)
    n = na + nc
    calls = 6
    g = torch.Generator(device="cpu").manual_seed(world)
    p0 = (torch.randn(n, generator=g) * 0.1)
    grads = [[torch.cat([torch.randn(n, generator=g) * 0.05, torch.randn(8, generator=g)])
              for _ in range(world)] for _ in range(calls)]
    dgrads = [[x.to(DEV) for x in per_call] for per_call in grads]
    groups = PeerGroup.local_group(n + 8, DEV, world)
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    st = []
    for r in range(world):
        st.append(dict(p=p0.to(DEV), mu=torch.zeros(n, device=DEV), nu=torch.zeros(n, device=DEV),
                       c=torch.zeros(2, dtype=torch.int32, device=DEV),
                       gsum=torch.zeros(n, device=DEV), loss=torch.zeros(calls, 5, device=DEV)))
    torch.cuda.synchronize()
    for k in range(calls):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                # the rank overwrites its buffer right after the previous call on ITS stream: legal
                # because only the rank itself reads it (push), or because that call ended after
                # every peer had read it (pull)
                groups[r].grad.copy_(dgrads[k][r], non_blocking=True)
                native.reduce_clip_adam_pair(st[r]["p"], st[r]["mu"], st[r]["nu"], st[r]["c"],
                                             groups[r], st[r]["gsum"], na, nc, None, None, None,
                                             None, 1.0 / world, 2.5e-4, 2.5e-4, 0.5, 0, 1,
                                             st[r]["loss"][k])
    torch.cuda.synchronize()
    for r in range(world):
        seq, err = groups[r].status()
        assert err == 0 and seq == calls, (r, seq, err)
    for r in range(1, world):  # replicated parameters stay replicated, bit for bit
        assert torch.equal(st[r]["p"], st[0]["p"]) and torch.equal(st[r]["nu"], st[0]["nu"])
        assert torch.equal(st[r]["loss"], st[0]["loss"])
    want, _, _ = _oracle_steps(p0.numpy(), [[x.numpy()[:n] for x in pc] for pc in grads], na,
                               1.0 / world, (2.5e-4, 2.5e-4), 0.5, 0, 1)
    np.testing.assert_allclose(st[0]["p"].cpu().numpy(), want, rtol=1e-5, atol=1e-7)
    for k in range(calls):
        s = grads[k][0][n:n + 5].clone()
        for x in grads[k][1:]:
            s = s + x[n:n + 5]
        torch.testing.assert_close(st[0]["loss"][k].cpu(), s * (1.0 / world), rtol=1e-6, atol=1e-7)
    for grp in groups[1:] + groups[:1]:
        grp.release()


def test_peer_two_gpus(lib_built):
    """2 ranks on 2 GPUs under torchrun: the peer path equals the NCCL path bit for bit (a sum of two
    terms does not depend on the order) through whole learner updates."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, MAVA_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "peer_multi_gpu.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_OK" in r.stdout
