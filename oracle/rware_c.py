"""ctypes front end of the C port of the RWARE oracle (oracle/c/rware_oracle.c).

TEST INFRASTRUCTURE ONLY.  Batched over envs with OpenMP; used where the per-env numpy oracle is
too slow (large parity runs, the CPU baseline of bench.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "librware_oracle.so"


class _Spec(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("H", "W", "A", "Q", "R", "n", "FR", "time_limit",
                                       "column_height")]


def build() -> Path:
    src = _DIR / "c" / "rware_oracle.c"
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.rw_state_words.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RwareC:
    def __init__(self, column_height=8, shelf_rows=1, shelf_columns=3, num_agents=4, sensor_range=1,
                 request_queue_size=4, time_limit=500):
        lib = _load()
        self.spec = _Spec()
        lib.rw_make_spec(column_height, shelf_rows, shelf_columns, num_agents, sensor_range,
                         request_queue_size, time_limit, C.byref(self.spec))
        self.A, self.FR = self.spec.A, self.spec.FR
        self.words = lib.rw_state_words(C.byref(self.spec))

    def reset(self, keys: np.ndarray):
        n = keys.shape[0]
        keys = np.ascontiguousarray(keys, np.uint32)
        state = np.zeros((n, self.words), np.int32)
        view = np.zeros((n, self.A, self.FR), np.int8)
        mask = np.zeros((n, self.A), np.uint8)
        _load().rw_reset(C.byref(self.spec), _ptr(keys), _ptr(state), _ptr(view), _ptr(mask), n)
        return state, view, mask

    def step(self, state, action, auto_reset=True):
        n = state.shape[0]
        action = np.ascontiguousarray(action, np.int8)
        view = np.zeros((n, self.A, self.FR), np.int8)
        mask = np.zeros((n, self.A), np.uint8)
        reward = np.zeros((n, self.A), np.float32)
        done = np.zeros(n, np.uint8)
        ep_ret = np.zeros(n, np.float32)
        ep_len = np.zeros(n, np.int32)
        _load().rw_step(C.byref(self.spec), _ptr(state), _ptr(action), _ptr(view), _ptr(mask),
                        _ptr(reward), _ptr(done), _ptr(ep_ret), _ptr(ep_len), n, int(auto_reset))
        return view, mask, reward, done, ep_ret, ep_len
