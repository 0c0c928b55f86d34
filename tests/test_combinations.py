"""CPU: quartet enumeration / unranking against the reference's golden vectors."""
from itertools import combinations
from math import comb

import numpy as np

from conftest import load_golden
from tetrad_b200 import combinations as cb


def test_unrank_matches_reference_vectors():
    g = load_golden("unrank")
    assert (cb.unrank_quartets(g["idx12"], 12) == g["comb12"]).all()
    assert (cb.unrank_quartets(g["idx200"], 200) == g["comb200"]).all()


def test_unrank_is_lexicographic_enumeration():
    for n in (4, 5, 9, 13):
        want = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
        assert (cb.unrank_quartets(np.arange(comb(n, 4)), n) == want).all()
        assert (cb.all_quartets(n) == want).all()


def test_random_quartets_follow_generator_stream():
    rng = np.random.default_rng(5)
    idx = np.random.default_rng(5).choice(comb(30, 4), size=500, replace=False)
    q = cb.random_quartets(30, 500, rng)
    assert (q == cb.unrank_quartets(idx, 30)).all()
    assert len({tuple(r) for r in q.tolist()}) == 500
    assert (np.diff(q.astype(int), axis=1) > 0).all()
    chunks = list(cb.iter_chunks_random(30, 500, 128, 5))
    assert sum(len(c) for c in chunks) == 500 and (np.concatenate(chunks) == q).all()
    assert sum(len(c) for c in cb.iter_chunks_full(10, 64)) == comb(10, 4)


def test_keyed_sampler_is_a_permutation_and_matches_its_oracle(lib):
    """tq_sample_ranks_host (the host twin of the device sampler's Feistel permutation) against the numpy restatement;
    a bijection on [0, C(n,4)); its first Q images look like a uniform sample (chi-square over 64 equal buckets)."""
    import ctypes
    from math import comb
    from oracle import sample_oracle
    from tetrad_b200 import _lib
    for n, key in ((12, 1), (13, 0xDEADBEEFCAFEF00D), (50, 42)):
        total = comb(n, 4)
        out = np.empty(total, dtype=np.int64)
        _lib.check(lib.tq_sample_ranks_host(0, total, n, key, out.ctypes.data))
        assert (np.sort(out) == np.arange(total)).all(), "not a permutation"
        assert (out == sample_oracle.permute(np.arange(total), n, key)).all()
    n, q = 200, 400_000
    out = np.empty(q, dtype=np.int64)
    _lib.check(lib.tq_sample_ranks_host(0, q, n, 7, out.ctypes.data))
    assert len(np.unique(out)) == q and out.max() < comb(n, 4)
    assert (out[:5000] == sample_oracle.permute(np.arange(5000), n, 7)).all()
    hist = np.bincount((out * 64 // comb(n, 4)).astype(np.int64), minlength=64)
    chi2 = ((hist - q / 64) ** 2 / (q / 64)).sum()
    assert chi2 < 120, chi2                                   # 63 degrees of freedom: P(chi2 > 120) ~ 1e-5
    other = np.empty(1000, dtype=np.int64)
    _lib.check(lib.tq_sample_ranks_host(0, 1000, n, 8, other.ctypes.data))
    assert (other != out[:1000]).mean() > 0.99                # another key, another sample
    assert lib.tq_sample_ranks_host(comb(12, 4) - 3, 4, 12, 1, other.ctypes.data) == _lib.TQ_EINVAL
