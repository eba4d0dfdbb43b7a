"""Oracle: JAX's default PRNG (threefry2x32) restated in numpy.  TEST INFRASTRUCTURE ONLY.

The reference draws every random number through ``jax.random`` with the default
``threefry2x32`` implementation (``jax==0.4.30``, requirements/requirements.txt:9-10,
``jax_threefry_partitionable=False``).  Call sites on the hot path:
``mava/systems/ppo/ff_mappo.py:81,204,269,273,392,417,445``,
``mava/wrappers/auto_reset_wrapper.py:74``, ``mava/wrappers/episode_metrics.py:61``.

jax itself is NOT under /root/reference (third-party, pinned 0.4.30).  This restates its
published algorithm (jax/_src/prng.py, jax/_src/random.py) and is pinned by

* the Random123 threefry2x32 known-answer vectors (also used by jax's own test-suite), and
* the ``jax.random.split`` outputs printed in the JAX documentation,

both held in ``tests/golden/threefry_kat.json``.
"""
from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """20-round Threefry-2x32 block cipher on arrays of counters (x0, x1) with key (k0, k1)."""
    with np.errstate(over="ignore"):
        k0 = U32(k0)
        k1 = U32(k1)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, U32(k0 ^ k1 ^ _PARITY))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for r in range(5):
            for rot in _ROT[r % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, rot)
                x1 = x1 ^ x0
            x0 = x0 + ks[(r + 1) % 3]
            x1 = x1 + ks[(r + 2) % 3] + U32(r + 1)
    return x0, x1


def threefry_2x32(key, count: np.ndarray) -> np.ndarray:
    """jax._src.prng.threefry_2x32: hash a flat counter array, pairing first/second halves."""
    count = np.asarray(count, dtype=U32)
    flat = count.ravel()
    odd = flat.size % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, U32)])
    half = flat.size // 2
    y0, y1 = threefry2x32(key[0], key[1], flat[:half], flat[half:])
    out = np.concatenate([y0, y1])
    if odd:
        out = out[:-1]
    return out.reshape(count.shape)


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) -> uint32[2] = (high word, low word)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split (original, non-partitionable threefry): uint32[num, 2]."""
    counts = np.arange(num * 2, dtype=U32)
    return threefry_2x32(key, counts).reshape(num, 2)


def random_bits(key, shape) -> np.ndarray:
    """jax._src.prng._threefry_random_bits_original for bit_width=32."""
    size = int(np.prod(shape, dtype=np.int64)) if len(tuple(shape)) else 1
    bits = threefry_2x32(key, np.arange(size, dtype=U32))
    return bits.reshape(shape)


def uniform(key, shape, minval=0.0, maxval=1.0) -> np.ndarray:
    """jax.random.uniform for float32."""
    bits = random_bits(key, shape)
    fbits = (bits >> U32(9)) | U32(0x3F800000)
    floats = fbits.view(np.float32) - np.float32(1.0)
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, floats * (maxval - minval) + minval).astype(np.float32)


def gumbel(key, shape) -> np.ndarray:
    """jax.random.gumbel float32: -log(-log(U(tiny, 1)))."""
    u = uniform(key, shape, minval=np.finfo(np.float32).tiny, maxval=1.0)
    return (-np.log(-np.log(u))).astype(np.float32)


def categorical(key, logits: np.ndarray) -> np.ndarray:
    """jax.random.categorical over the last axis (Gumbel arg-max)."""
    g = gumbel(key, logits.shape)
    return np.argmax(g + logits.astype(np.float32), axis=-1).astype(np.int32)


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """jax.random.randint for int32 with 0 <= span < 2**31."""
    k1, k2 = split(key, 2)
    hi = random_bits(k1, shape).astype(np.uint64)
    lo = random_bits(k2, shape).astype(np.uint64)
    span = np.uint64(max(1, maxval - minval))
    mult = np.uint64((2 ** 16) % int(span))
    mult = np.uint64((int(mult) * int(mult)) % int(span))
    off = ((hi % span) * mult) & np.uint64(0xFFFFFFFF)
    off = (off + (lo % span)) & np.uint64(0xFFFFFFFF)
    off = off % span
    return (minval + off.astype(np.int64)).astype(np.int32)


def shuffle(key, x: np.ndarray) -> np.ndarray:
    """jax._src.random._shuffle along axis 0: rounds of stable sort by fresh 32-bit keys."""
    x = np.asarray(x)
    n = x.shape[0]
    rounds = int(math.ceil(3 * math.log(max(1, x.size)) / math.log(2 ** 32 - 1)))
    for _ in range(rounds):
        key, sub = split(key, 2)
        sort_keys = random_bits(sub, (n,))
        order = np.argsort(sort_keys, kind="stable")
        x = x[order]
    return x


def permutation(key, n: int) -> np.ndarray:
    """jax.random.permutation(key, n)."""
    return shuffle(key, np.arange(n, dtype=np.int32))


def choice_no_replace(key, a: np.ndarray, num: int) -> np.ndarray:
    """jax.random.choice(key, a, (num,), replace=False) == permutation(key, a)[:num]."""
    return shuffle(key, np.asarray(a))[:num]
