"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the device-side bootstrap sampler.

The reference builds a bootstrap replicate with ``resample_tmp_database`` + ``jit_resample`` +
``jit_resolve_ambigs`` (tetrad/src/run_inference.py:99-143, tetrad/jit/resample.py:20-64,
tetrad/jit/resolve_ambigs.py:12-36): loci drawn with replacement, the columns of every drawn
locus randomly permuted, every IUPAC ambiguity cell resolved to one of its two bases with p = 1/2,
A/C/G/T -> 0..3.  Its permutations and coins come from numba's private generator, which no other
implementation can replay, so the CUDA path (``tq_boot_pack``) defines its own
counter-based streams (Philox4x32-10) and THIS module restates exactly those streams in numpy:
given the same ``lidxs`` and seeds it materialises the very ``tmparr``/``tmpmap`` whose bit-planes
the device packs, which makes the device sampler checkable bit for bit (tests/test_gpu_boot.py);
its distribution is checked against the reference's semantics in tests/test_boot_oracle.py.
"""
from __future__ import annotations

import numpy as np

GETCONS = np.array([[82, 71, 65], [75, 71, 84], [83, 71, 67], [89, 84, 67], [87, 84, 65], [77, 67, 65]], dtype=np.uint8)
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int) -> np.ndarray:
    """First output word of Philox4x32-10 for counter (c0,c1,c2,c3) and key (k0,k1), vectorised."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & _MASK, p0 & _MASK, n0 & _MASK, n2 & _MASK
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0.astype(np.uint32)


def instance_columns(spans: np.ndarray, lidxs: np.ndarray, seed_shuffle: int):
    """(cols i64[S'], inst i64[S'], rank i64[S']): source column, instance and within-instance
    rank of every column of the replicate, instance by instance in rank order."""
    lens = (spans[lidxs, 1] - spans[lidxs, 0]).astype(np.int64)
    total = int(lens.sum())
    inst = np.repeat(np.arange(len(lidxs), dtype=np.int64), lens)
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])
    j = np.arange(total) - np.repeat(first, lens)                      # column index inside the locus
    key = philox4x32_10(inst, j, 0, 0, seed_shuffle, 0x5EED).astype(np.int64)
    order = np.lexsort((j, key, inst))                                  # by instance, then (key, j)
    cols = np.repeat(spans[lidxs, 0], lens) + j[order]
    return cols, inst, j                                                # inst and rank are already in order


def bootstrap_arrays(seqarr: np.ndarray, spans: np.ndarray, lidxs: np.ndarray, seed_shuffle: int, seed_ambig: int):
    """The replicate (tmparr u8[n,S'] with codes 0..3 / 78, tmpmap u32[S',2]) that
    ``tq_boot_pack`` packs for these arguments."""
    cols, inst, rank = instance_columns(spans, np.asarray(lidxs), seed_shuffle)
    tmparr = np.ascontiguousarray(seqarr[:, cols])
    ntips = tmparr.shape[0]
    tips = np.arange(ntips)[:, None]
    coin = (philox4x32_10(inst[None, :], rank[None, :], tips, 1, seed_ambig, 0xA3B1) & 1).astype(bool)
    out = np.full(tmparr.shape, 78, dtype=np.uint8)
    for code, idx in ((65, 0), (67, 1), (71, 2), (84, 3), (0, 0), (1, 1), (2, 2), (3, 3)):
        out[tmparr == code] = idx
    lut = {65: 0, 67: 1, 71: 2, 84: 3}
    for ambig, res1, res2 in GETCONS:
        where = tmparr == ambig
        out[where & coin] = lut[int(res1)]                              # coin set -> first base
        out[where & ~coin] = lut[int(res2)]
    tmpmap = np.zeros((len(cols), 2), dtype=np.uint32)
    tmpmap[:, 0] = inst
    tmpmap[:, 1] = np.arange(len(cols))
    return out, tmpmap
