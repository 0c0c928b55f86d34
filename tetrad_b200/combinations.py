"""Quartet index sets: the caller side of the hot path (reference ``tetrad/src/combinations.py``).

* ``all_quartets(n)``            -- lexicographic ``itertools.combinations(range(n), 4)``
  (reference ``iter_chunks_full`` :82-87), as one u32[m,4] array.
* ``unrank_quartets(idx, n)``    -- vectorised ``_index_to_combination`` (:94-106): lexicographic
  rank -> 4 ascending tip ids.
* ``random_quartets(n, size, rng)`` -- ``random_combination_sample_via_index`` (:109-114):
  ``rng.choice(C(n,4), size, replace=False)`` then unrank, so the same numpy Generator state
  yields the same quartets as the reference.
"""
from __future__ import annotations

from itertools import combinations
from math import comb

import numpy as np


def all_quartets(nsamples: int) -> np.ndarray:
    if nsamples < 4:
        return np.zeros((0, 4), dtype=np.uint32)
    return np.array(list(combinations(range(nsamples), 4)), dtype=np.uint32)


def unrank_quartets(index, nsamples: int) -> np.ndarray:
    """Lexicographic rank(s) -> quartets u32[m,4]; identical to the reference's scalar loop."""
    n = int(nsamples)
    r = np.atleast_1d(np.asarray(index, dtype=np.int64))
    if n < 4 or (r.size and (r.min() < 0 or r.max() >= comb(n, 4))):
        raise ValueError("quartet rank out of range")
    c2 = np.array([comb(m, 2) for m in range(n + 1)], dtype=np.int64)
    c3 = np.array([comb(m, 3) for m in range(n + 1)], dtype=np.int64)
    c4 = np.array([comb(m, 4) for m in range(n + 1)], dtype=np.int64)
    # first id a: ranks below cum1[a] start with an id < a
    cum1 = c4[n] - c4[n - np.arange(n + 1)]
    a = np.searchsorted(cum1, r, side="right") - 1
    r1 = r - cum1[a]
    x = c3[n - a - 1] - r1                        # >= 1
    b = n - np.searchsorted(c3, x, side="left")
    r2 = r1 - (c3[n - a - 1] - c3[n - b])
    y = c2[n - b - 1] - r2
    c = n - np.searchsorted(c2, y, side="left")
    r3 = r2 - (c2[n - b - 1] - c2[n - c])
    d = c + 1 + r3
    return np.stack([a, b, c, d], axis=1).astype(np.uint32)


def random_quartets(nsamples: int, size: int, rng) -> np.ndarray:
    """``size`` distinct random quartets (reference combinations.py:109-114)."""
    rng = np.random.default_rng(rng)
    total = comb(nsamples, 4)
    sampled = rng.choice(total, size=min(size, total), replace=False)
    return unrank_quartets(sampled, nsamples)


def iter_chunks_full(nsamples: int, max_size: int):
    """Chunks of all quartets (reference :82-87) as u32 arrays."""
    quartets = all_quartets(nsamples)
    for lo in range(0, len(quartets), max_size):
        yield quartets[lo: lo + max_size]


def iter_chunks_random(nsamples: int, size: int, max_size: int, rng):
    """Chunks of a random quartet sample (reference :117-121) as u32 arrays."""
    quartets = random_quartets(nsamples, size, rng)
    for lo in range(0, len(quartets), max_size):
        yield quartets[lo: lo + max_size]


def device_quartets(nsamples: int, device: int = 0, ranks=None, start: int = 0, count: int | None = None):
    """Quartets unranked on the GPU (``tq_unrank``): an int32 CUDA tensor [m,4].  ``ranks`` (numpy
    or CUDA int64) selects a sample; otherwise the contiguous rank range ``[start, start+count)``
    (default: all C(n,4) quartets) is produced without any host->device copy."""
    import torch
    from . import _lib
    lib = _lib.load()
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if ranks is not None:
        if isinstance(ranks, torch.Tensor):
            r = ranks.to(dev, dtype=torch.int64).contiguous()
            bad = bool(r.numel()) and (int(r.min()) < 0 or int(r.max()) >= comb(nsamples, 4))    # device tensor: one sync
        else:
            host = np.ascontiguousarray(ranks, dtype=np.int64)
            bad = bool(host.size) and (int(host.min()) < 0 or int(host.max()) >= comb(nsamples, 4))   # checked on the host
            from .resolve_quartets import upload
            r = upload(host, dev)
        if bad:
            raise ValueError("quartet rank out of range")
        out = torch.empty((r.numel(), 4), dtype=torch.int32, device=dev)
        _lib.check(lib.tq_unrank(device, r.data_ptr(), 0, r.numel(), nsamples, out.data_ptr(), stream))
        return out
    m = comb(nsamples, 4) - start if count is None else count
    out = torch.empty((m, 4), dtype=torch.int32, device=dev)
    _lib.check(lib.tq_unrank(device, None, start, m, nsamples, out.data_ptr(), stream))
    return out


def device_sample(nsamples: int, key: int, count: int, device: int = 0, start: int = 0):
    """``count`` quartets of the keyed random sample (``tq_sample_quartets``): quartet i = unrank(pi_key(i)), pi_key a
    pseudo-random permutation of all C(n,4) ranks, so the quartets are distinct and any index range can be produced
    on any GPU with no host draw and no upload.  int32 CUDA tensor [count, 4]."""
    import torch
    from . import _lib
    lib = _lib.load()
    dev = torch.device("cuda", device)
    out = torch.empty((count, 4), dtype=torch.int32, device=dev)
    _lib.check(lib.tq_sample_quartets(device, start, count, nsamples, int(key) & (2**64 - 1), out.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream))
    return out
