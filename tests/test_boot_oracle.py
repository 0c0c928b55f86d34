"""CPU: the oracle-side restatement of the device bootstrap sampler has the reference's semantics
(tetrad/jit/resample.py:20-64, tetrad/jit/resolve_ambigs.py:12-36): every instance is a uniform
random permutation of its source locus, every ambiguity cell is one of its two bases with p = 1/2."""
import numpy as np

from oracle import boot_oracle
from tetrad_b200 import synth


def test_philox_known_answers():
    # Random123 known-answer vectors for Philox4x32-10 (first output word)
    assert int(boot_oracle.philox4x32_10(0, 0, 0, 0, 0, 0)) == 0x6627E8D5
    assert int(boot_oracle.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)) == 0x408F276D
    assert int(boot_oracle.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)) == 0xD16CFE09


def test_instances_are_permutations_with_instance_ids():
    d = synth.generate(5, 60, 260, seed=2, p_ambig=0.0)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=1)
    lidxs = np.random.default_rng(3).choice(60, 60, replace=True)
    tmparr, tmpmap = boot_oracle.bootstrap_arrays(db["seqarr"], db["spans"], lidxs, 123, 456)
    lens = db["spans"][lidxs, 1] - db["spans"][lidxs, 0]
    assert tmparr.shape[1] == lens.sum() and (np.bincount(tmpmap[:, 0], minlength=60) == lens).all()
    recoded = synth.recode_acgt(db["seqarr"].copy())
    off = 0
    for l in lidxs:
        a, b = db["spans"][l]
        assert sorted(map(bytes, recoded[:, a:b].T)) == sorted(map(bytes, tmparr[:, off: off + b - a].T))
        off += b - a
    assert set(np.unique(tmparr)) <= {0, 1, 2, 3, 78}
    again, _ = boot_oracle.bootstrap_arrays(db["seqarr"], db["spans"], lidxs, 123, 456)
    other, _ = boot_oracle.bootstrap_arrays(db["seqarr"], db["spans"], lidxs, 124, 456)
    assert (again == tmparr).all() and (other != tmparr).any()


def test_permutations_are_uniform_and_coins_fair():
    # one locus of 4 distinguishable columns drawn 6000 times: each of the 24 orders ~ 250 times
    seq = np.array([[65, 67, 71, 84], [65, 65, 65, 65]], dtype=np.uint8)
    spans = np.array([[0, 4], [0, 4]], dtype=np.int64)
    lidxs = np.zeros(6000, dtype=np.int64)
    tmparr, _ = boot_oracle.bootstrap_arrays(seq, spans, lidxs, 99, 7)
    orders = tmparr[0].reshape(6000, 4)
    codes = orders @ np.array([64, 16, 4, 1])
    counts = np.unique(codes, return_counts=True)[1]
    assert len(counts) == 24
    chi2 = ((counts - 250.0) ** 2 / 250.0).sum()
    assert chi2 < 60.0                              # 23 dof: P(chi2 > 60) < 1e-4
    # ambiguity R (G|A): 20000 cells, half G (2) half A (0)
    amb = np.full((4, 5000), 82, dtype=np.uint8)
    spans = np.array([[i, i + 1] for i in range(5000)], dtype=np.int64)
    out, _ = boot_oracle.bootstrap_arrays(amb, spans, np.arange(5000), 1, 2)
    assert set(np.unique(out)) == {0, 2}
    frac = (out == 2).mean()
    assert abs(frac - 0.5) < 0.015
    # all six codes resolve to exactly their two bases (tetrad/src/utils.py:14-21)
    for code, pair in ((82, {2, 0}), (75, {2, 3}), (83, {2, 1}), (89, {3, 1}), (87, {3, 0}), (77, {1, 0})):
        out, _ = boot_oracle.bootstrap_arrays(np.full((3, 400), code, np.uint8), spans[:400], np.arange(400), 5, 6)
        assert set(np.unique(out)) == pair
