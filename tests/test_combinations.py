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
