"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the keyed quartet sampler (tq_sample_quartets, tq_pack.cu).

Quartet i of a replicate's random sample is unrank(pi_key(i)), pi_key a pseudo-random permutation of [0, C(n,4)):
a 6-round Feistel network over ``2 * half`` bits (``half = ceil(ceil(log2 C(n,4)) / 2)``) with round function
``fmix32(fmix32(R ^ k0 ^ r * 0x9E3779B9) + k1) & mask`` (fmix32 = MurmurHash3's finaliser), cycle-walked until the image
falls below C(n,4).  It replaces the reference's host draw ``rng.choice(C(n,4), Q, replace=False)``
(tetrad/src/combinations.py:109-114) in distribution, not in stream.
"""
from math import comb

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def fmix32(h):
    h = h & M32
    h ^= h >> np.uint64(16); h = (h * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13); h = (h * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    return h


def permute(idx, nsamples: int, key: int):
    """pi_key(idx) for an array of indices < C(nsamples, 4)."""
    total = comb(nsamples, 4)
    bits = 2
    while (1 << bits) < total:
        bits += 1
    bits += bits & 1
    half = np.uint64(bits // 2)
    mask = np.uint64((1 << (bits // 2)) - 1)
    k0, k1 = np.uint64(key & 0xFFFFFFFF), np.uint64((key >> 32) & 0xFFFFFFFF)
    x = np.asarray(idx, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        L, R = x[todo] >> half, x[todo] & mask
        for r in range(6):
            f = fmix32((fmix32(R ^ k0 ^ np.uint64((r * 0x9E3779B9) & 0xFFFFFFFF)) + k1) & M32) & mask
            L, R = R, L ^ f
        x[todo] = (L << half) | R
        todo = x >= np.uint64(total)
    return x.astype(np.int64)
