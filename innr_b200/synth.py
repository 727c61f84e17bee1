"""Stateless synthetic generators on the host (numpy), identical to the device generators in csrc/common.cuh
(and to the oracle's): G-hash = splitmix64(salt + index) -> 24-bit uniform f32 in [-1, 1) / raw u64 words.
Used to make queries for bench.py and tests; corpora are generated on the device."""
from __future__ import annotations

import numpy as np

SALT_CORPUS = 0x5EED0000
SALT_QUERY = 0x5EED0001
SALT_CODES = 0x5EED0002

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def ghash_u64(salt: int, first_idx: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + np.uint64((salt + first_idx) & 0xFFFFFFFFFFFFFFFF)
    return splitmix64(idx)


def ghash_f32(salt: int, first_idx: int, count: int) -> np.ndarray:
    u = ghash_u64(salt, first_idx, count)
    return ((u >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)).astype(np.float32)
