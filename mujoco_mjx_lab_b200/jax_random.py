"""Host-side restatement of the jax.random pieces the reference trainers use to make env keys
(train_ppo.py:117-118,150-151: `rng, k = random.split(rng); keys = random.split(k, num_envs)`), so that the
`keys` handed to v_reset are the raw uint32[2] threefry key data JAX would produce (threefry2x32, 20 rounds,
jax_threefry_partitionable=True as in jax 0.7.2).  numpy only, vectorised."""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(key, c0, c1):
    """key: uint32[2]; c0, c1: uint32 arrays -> (x0, x1) uint32 arrays."""
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = (np.asarray(c0, dtype=np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(c1, dtype=np.uint32) + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = ((x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))).astype(np.uint32)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    idx = np.arange(num, dtype=np.uint32)
    x0, x1 = threefry2x32(key, np.zeros(num, dtype=np.uint32), idx)
    return np.stack([x0, x1], axis=1)


def bits(key, n: int) -> np.ndarray:
    x0, x1 = threefry2x32(key, np.zeros(n, dtype=np.uint32), np.arange(n, dtype=np.uint32))
    return x0 ^ x1


def uniform(key, n: int, minval: float = 0.0, maxval: float = 1.0) -> np.ndarray:
    b = (bits(key, n) >> np.uint32(9)) | np.uint32(0x3F800000)
    f = b.view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(minval), f * np.float32(maxval - minval) + np.float32(minval))
