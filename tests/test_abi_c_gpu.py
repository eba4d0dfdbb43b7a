"""The C header is the contract: a plain C99 program (tests/c/abi_smoke.c) that includes
include/mava_b200.h and links libmava_b200.so runs reset -> env steps -> GAE; the same sequence
through the ctypes binding must give byte-identical outputs (FNV-1a checksums)."""
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def fnv(buf: bytes) -> str:
    h = 1469598103934665603
    for b in buf:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_c_client_matches_ctypes_binding(lib_built, tmp_path):
    from mava_b200 import native

    exe = tmp_path / "abi_smoke"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT}/include", f"-I{cuda}/include",
           f"{ROOT}/tests/c/abi_smoke.c", f"-L{ROOT}/mava_b200", "-lmava_b200", f"-L{cuda}/lib64",
           "-lcudart", f"-Wl,-rpath,{ROOT}/mava_b200", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    E, T = 64, 24
    r = subprocess.run([str(exe), str(E), str(T)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = dict(line.split() for line in r.stdout.splitlines() if not line.startswith("ABI_SMOKE_OK"))
    assert "ABI_SMOKE_OK" in r.stdout

    env = native.Env.rware(column_height=8, shelf_rows=1, shelf_columns=3, num_agents=2,
                           sensor_range=1, request_queue_size=2, time_limit=12)
    A, FR = env.num_agents, env.view_dim
    e = np.arange(E, dtype=np.uint64)
    keys = np.stack([(0x9E3779B9 * (e + 1)) & 0xFFFFFFFF, 0x85EBCA6B ^ e], 1).astype(np.uint32)
    i = np.arange(T * E * A)
    act = ((i * 7 + i // 5) % 5).astype(np.int8).reshape(T, E, A)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=DEV)
    state = env.alloc_state(E, DEV)
    view, mask = z(T + 1, E, A, FR, dt=torch.int8), z(T + 1, E, A, dt=torch.uint8)
    reward, done = z(T, E, A), z(T, E, dt=torch.uint8)
    er, el = z(T, E), z(T, E, dt=torch.int32)
    tact = torch.from_numpy(act).to(DEV)
    env.reset(torch.from_numpy(keys).to(DEV), state, view[0], mask[0], E)
    for t in range(T):
        env.step(state, tact[t], view[t + 1], mask[t + 1], reward[t], done[t], er[t], el[t], E, True)
    value = (0.25 * (torch.arange(T, device=DEV) % 4).float()).view(T, 1, 1).expand(T, E, A).contiguous()
    last_val = torch.full((E, A), 0.5, device=DEV)
    adv, tgt = z(T, E, A), z(T, E, A)
    native.gae(reward, value, done, last_val, 0.99, 0.95, T, E, A, adv, tgt)
    torch.cuda.synchronize()
    want = {"state": state, "view": view, "mask": mask, "reward": reward, "done": done,
            "ep_return": er, "ep_length": el, "adv": adv, "targets": tgt}
    for name, ten in want.items():
        assert got[name] == fnv(ten.cpu().numpy().tobytes()), name
    assert int(done.sum()) > 0  # time limit 12 < T: auto-resets happened
