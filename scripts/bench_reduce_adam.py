"""Micro-benchmark: the fused reduce + clip + Adam kernel (csrc/peer.cu) against the two-launch
optimiser it replaces, both replayed from CUDA graphs of 64 back-to-back calls (headline networks).
Under torchrun with N ranks the peer path is also timed against NCCL all_reduce + the same kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mava_b200 import native
from mava_b200.peer import PeerGroup
from mava_b200.systems.ppo import _runner


def graph_time(fn, calls=64, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        if dist.is_initialized():
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / calls)
    return best


def main():
    dev = _runner.init_distributed()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    A, FR, N = 4, 66, 5
    actor = native.mlp_desc(native.IN_AGENT_VIEW, True, A, FR, 128, 128, N)
    critic = native.mlp_desc(native.IN_GLOBAL, True, A, FR, 128, 128, 1)
    na, nc = native.mlp_param_count(actor), native.mlp_param_count(critic)
    n = na + nc
    z = lambda k, dt=torch.float32: torch.zeros(k, dtype=dt, device=dev)
    params, mu, nu, counts, gsum = torch.randn(n, device=dev) * 0.1, z(n), z(n), z(2, torch.int32), z(n)
    ai, ci = z(native.mlp_pack_bytes(actor), torch.uint8), z(native.mlp_pack_bytes(critic), torch.uint8)
    loss = z(5)
    local = PeerGroup(n + 8, dev)
    local.grad.copy_(torch.randn(n + 8, device=dev) * 0.01)
    out = {}
    out["old_two_launches_us"] = graph_time(lambda: native.clip_adam_pair_pack(
        params, mu, nu, counts, local.grad, actor, ai, critic, ci, 1.0, 2.5e-4, 2.5e-4, 0.5))
    out["fused_world1_us"] = graph_time(lambda: native.reduce_clip_adam_pair(
        params, mu, nu, counts, local, gsum, na, nc, actor, ai, critic, ci, 1.0, 2.5e-4, 2.5e-4, 0.5,
        0, 1, loss))
    if world > 1:
        grp = PeerGroup(n + 8, dev, rank, world)
        grp.grad.copy_(torch.randn(n + 8, device=dev) * 0.01)
        out["fused_peer_us"] = graph_time(lambda: native.reduce_clip_adam_pair(
            params, mu, nu, counts, grp, gsum, na, nc, actor, ai, critic, ci, 1.0 / world, 2.5e-4,
            2.5e-4, 0.5, 0, 1, loss))
        seq, err = grp.status()
        assert err == 0, (seq, err)
        import ctypes
        from mava_b200 import _lib as _l
        raw = ctypes.CDLL(str(_l.LIB_PATH))
        if hasattr(raw, "mava_debug_peer_stamps"):  # built with -DMAVA_PEER_STAMPS
            buf = (ctypes.c_ulonglong * 8)()
            raw.mava_debug_peer_stamps(buf, 1)
            torch.cuda.synchronize()
            fn = lambda: native.reduce_clip_adam_pair(
                params, mu, nu, counts, grp, gsum, na, nc, actor, ai, critic, ci, 1.0 / world, 2.5e-4,
                2.5e-4, 0.5, 0, 1, loss)
            dist.barrier()
            for _ in range(256):
                fn()
            torch.cuda.synchronize()
            raw.mava_debug_peer_stamps(buf, 0)
            nn = max(int(buf[7]), 1)
            names = ["ready handshake", "peer reads + sum", "norm + grid barrier", "adam", "done handshake"]
            print("PEER_STAMPS rank", rank, "calls", nn,
                  {k: round(buf[i] / nn / 1e3, 2) for i, k in enumerate(names)}, flush=True)

        def nccl_path():
            dist.all_reduce(local.grad)
            native.reduce_clip_adam_pair(params, mu, nu, counts, local, gsum, na, nc, actor, ai,
                                         critic, ci, 1.0 / world, 2.5e-4, 2.5e-4, 0.5, 0, 1, loss)
        out["nccl_plus_fused_us"] = graph_time(nccl_path)
        t = torch.tensor([out[k] for k in sorted(out)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = dict(zip(sorted(out), t.tolist()))
        grp.release()
    if rank == 0:
        print("REDUCE_ADAM", world, {k: round(v, 2) for k, v in out.items()}, flush=True)
    if world > 1:
        import gc
        gc.collect()
        os._exit(0)  # graphs that captured NCCL are alive


if __name__ == "__main__":
    main()
