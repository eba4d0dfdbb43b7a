"""Host-side key derivation with the reference's PRNG conventions (jax.random, threefry2x32).

Used for the handful of keys derived once at start-up (run_experiment / learner_setup,
mava/systems/ppo/ff_mappo.py:392-394,417,445-447).  Per-step randomness is generated on the
device (csrc/prng.cuh).  This is product code and does not depend on ``oracle/``.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _block(k0, k1, x0, x1):
    with np.errstate(over="ignore"):
        k0, k1 = _U32(k0), _U32(k1)
        x0 = np.asarray(x0, _U32).copy()
        x1 = np.asarray(x1, _U32).copy()
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 += ks[0]
        x1 += ks[1]
        for r in range(5):
            for rot in _ROT[r % 2]:
                x0 += x1
                x1 = (x1 << _U32(rot)) | (x1 >> _U32(32 - rot))
                x1 ^= x0
            x0 += ks[(r + 1) % 3]
            x1 += ks[(r + 2) % 3] + _U32(r + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], _U32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split -> uint32[num, 2]."""
    counts = np.arange(2 * num, dtype=_U32)
    y0, y1 = _block(key[0], key[1], counts[:num], counts[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)
