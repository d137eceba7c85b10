"""Deterministic synthetic key generators (numpy), shared by tests/ and bench.py.

The reference draws keys from ``std::default_random_engine`` seeded with 0 (Utils.h:24-33), which is
implementation-defined and therefore not reproducible across hosts.  These
generators use an explicit counter-based hash (splitmix64 finaliser) instead, so the same
(kind, n, seed) gives the same keys everywhere.

Distributions are BASELINE.json's configs 2 and 4 (definitions from SURVEY 8(d)).
"""
from __future__ import annotations

import numpy as np

KINDS = ("uniform", "all_equal", "entropy4_table", "low_nibble", "sorted", "reverse")


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def uniform_u32(n: int, seed: int = 0, offset: int = 0) -> np.ndarray:
    """n uniform 32-bit words: hash(seed, offset + i)."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        salt = _splitmix64(np.array([seed], dtype=np.uint64))[0]
        return (_splitmix64(idx ^ salt) >> np.uint64(32)).astype(np.uint32)


def make_keys(kind: str, n: int, seed: int = 0) -> np.ndarray:
    if kind == "uniform":
        return uniform_u32(n, seed)
    if kind == "all_equal":  # every pass has one full bucket: all passes skippable
        return np.full(n, 0xDEADBEEF, dtype=np.uint32)
    if kind == "entropy4_table":  # 16 distinct random 32-bit values: <=16 live buckets per digit, nothing skippable
        table = uniform_u32(16, seed + 12345)
        return table[(uniform_u32(n, seed) & np.uint32(15)).astype(np.int64)]
    if kind == "low_nibble":  # only the low 4 bits vary: passes above digit 0 are skippable for r >= 4
        return uniform_u32(n, seed) & np.uint32(0xF)
    if kind == "sorted":  # strictly ascending, max < 2^32 for n <= 2^28
        step = max(1, min(16, (1 << 32) // max(n, 1)))
        return (np.arange(n, dtype=np.uint64) * np.uint64(step)).astype(np.uint32)
    if kind == "reverse":
        step = max(1, min(16, (1 << 32) // max(n, 1)))
        return (np.arange(n - 1, -1, -1, dtype=np.int64).astype(np.uint64) * np.uint64(step)).astype(np.uint32)
    raise ValueError(f"unknown key kind {kind!r}")
