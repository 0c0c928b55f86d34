"""Seeded tree-structured synthetic SNP matrices in ipyrad ``snps``/``snpsmap`` form.

This is the input generator SURVEY.md §8(d) specifies for every config of
BASELINE.json (the reference ships no sample data).  It emits exactly what
``init_database`` reads (reference ``tetrad/src/write_database.py:132-168``):

* ``snps``    u8[n, S]   ASCII codes (A C G T N - and IUPAC R K S Y W M)
* ``snpsmap`` i64[S, 5]  col0 = 1-indexed locus id (sorted), col1 = snp index in locus
* ``names``   list[str]

``prepare_database`` then restates ``init_database``'s transformation into the
working arrays (``seqarr``, ``spans``, ``tmparr``, ``tmpmap``) the hot path consumes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# IUPAC two-allele ambiguity codes; rows follow tetrad/src/utils.py:14-21 (GETCONS)
GETCONS = np.array(
    [[82, 71, 65], [75, 71, 84], [83, 71, 67], [89, 84, 67], [87, 84, 65], [77, 67, 65]],
    dtype=np.uint8,
)
_BASES = np.array([65, 67, 71, 84], dtype=np.uint8)  # A C G T

# named shapes of BASELINE.json:configs  (n tips, L loci, S snps, Q quartets, boots)
CONFIGS = {
    "C1": dict(ntips=12, nloci=5_000, nsnps=20_000, nquartets=0, nboots=0),
    "C2": dict(ntips=50, nloci=50_000, nsnps=200_000, nquartets=0, nboots=0),
    "C3": dict(ntips=100, nloci=100_000, nsnps=500_000, nquartets=1_000_000, nboots=20),
    "C4": dict(ntips=200, nloci=200_000, nsnps=1_000_000, nquartets=1_000_000, nboots=100),
    "C5": dict(ntips=500, nloci=500_000, nsnps=2_000_000, nquartets=10_000_000, nboots=10),
}


@dataclass
class SynthData:
    snps: np.ndarray      # u8[n, S] ascii
    snpsmap: np.ndarray   # i64[S, 5]
    names: list
    clades: np.ndarray    # bool[2n-2, n] descendant sets of every branch (the planted tree)


def _random_tree_clades(ntips: int, rng: np.random.Generator):
    """Random binary tree by random joining; returns (clades bool[2n-2, n], branch lengths)."""
    active = [np.eye(ntips, dtype=bool)[i] for i in range(ntips)]
    clades = list(active)
    while len(active) > 2:
        i, j = sorted(rng.choice(len(active), 2, replace=False))
        merged = active[i] | active[j]
        active.pop(j)
        active.pop(i)
        active.append(merged)
        clades.append(merged)
    clades = np.array(clades[: 2 * ntips - 2] if len(clades) >= 2 * ntips - 2 else clades)
    # internal branches a little shorter than tips, as in a radiating species tree
    sizes = clades.sum(1)
    blen = rng.gamma(2.0, 1.0, size=clades.shape[0]) * np.where(sizes == 1, 1.0, 0.6)
    return clades, blen


def generate(ntips: int, nloci: int, nsnps: int, seed: int = 1234, homoplasy: float = 0.15,
             p_missing: float = 0.10, p_gap: float = 0.02, p_ambig: float = 0.005) -> SynthData:
    """Generate ``snps``/``snpsmap`` of the requested shape."""
    if nloci < 2 or nsnps < nloci:
        raise ValueError("need nloci >= 2 and nsnps >= nloci")
    rng = np.random.default_rng(seed)
    clades, blen = _random_tree_clades(ntips, rng)
    prob = blen / blen.sum()

    # locus sizes: every locus >= 1 SNP, the remainder multinomial
    sizes = 1 + rng.multinomial(nsnps - nloci, np.full(nloci, 1.0 / nloci))
    locus = np.repeat(np.arange(1, nloci + 1, dtype=np.int64), sizes)
    snpsmap = np.zeros((nsnps, 5), dtype=np.int64)
    snpsmap[:, 0] = locus
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    snpsmap[:, 1] = np.arange(nsnps) - np.repeat(starts, sizes)
    snpsmap[:, 2] = snpsmap[:, 1] * 7 + 1
    snpsmap[:, 4] = np.arange(1, nsnps + 1)

    snps = np.empty((ntips, nsnps), dtype=np.uint8)
    step = max(1, min(nsnps, (1 << 24) // max(1, ntips)))
    for lo in range(0, nsnps, step):
        hi = min(nsnps, lo + step)
        w = hi - lo
        anc = rng.integers(0, 4, size=w)
        der = (anc + rng.integers(1, 4, size=w)) % 4
        br = rng.choice(clades.shape[0], size=w, p=prob)
        idx = np.where(clades[br].T, der[None, :], anc[None, :])  # [n, w]
        # second, independent mutation on a fraction of columns (homoplasy / third allele)
        second = rng.random(w) < homoplasy
        br2 = rng.choice(clades.shape[0], size=w, p=prob)
        der2 = rng.integers(0, 4, size=w)
        hit = clades[br2].T & second[None, :]
        idx = np.where(hit, der2[None, :], idx)
        block = _BASES[idx]
        # IUPAC ambiguity of the column's two alleles on a few cells
        amb = rng.random((ntips, w)) < p_ambig
        if amb.any():
            lut = np.zeros((4, 4), dtype=np.uint8)
            for code, r1, r2 in GETCONS:
                a, b = int(np.where(_BASES == r1)[0][0]), int(np.where(_BASES == r2)[0][0])
                lut[a, b] = lut[b, a] = code
            acode = lut[anc, der]  # anc != der always, so every entry is a valid IUPAC code
            block = np.where(amb, acode[None, :], block)
        r = rng.random((ntips, w))
        block = np.where(r < p_missing, np.uint8(78), block)
        block = np.where((r >= p_missing) & (r < p_missing + p_gap), np.uint8(45), block)
        snps[:, lo:hi] = block
    names = [f"t{i}" for i in range(ntips)]
    return SynthData(snps=snps, snpsmap=snpsmap, names=names, clades=clades)


def get_spans(locus_ids: np.ndarray) -> np.ndarray:
    """(start, end) column span of every locus run; restates jit_get_spans
    (reference tetrad/jit/get_spans.py:11-48) for sorted contiguous locus ids."""
    locus_ids = np.asarray(locus_ids)
    change = np.flatnonzero(locus_ids[1:] != locus_ids[:-1]) + 1
    starts = np.concatenate([[0], change])
    ends = np.concatenate([change, [locus_ids.shape[0]]])
    return np.stack([starts, ends], axis=1).astype(np.int64)


def resolve_ambigs(seq: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Coin-flip every IUPAC R/K/S/Y/W/M cell to one of its two bases, in place
    (semantics of jit_resolve_ambigs, reference tetrad/jit/resolve_ambigs.py:12-36;
    the numba RNG stream itself is not reproducible)."""
    for ambig, res1, res2 in GETCONS:
        where = np.flatnonzero(seq.ravel() == ambig)
        if where.size:
            coin = rng.random(where.size) < 0.5
            seq.ravel()[where] = np.where(coin, res1, res2)
    return seq


def recode_acgt(seq: np.ndarray) -> np.ndarray:
    """A/C/G/T (65/67/71/84) -> 0/1/2/3 in place (reference write_database.py:164-167)."""
    lut = np.arange(256, dtype=np.uint8)
    lut[[65, 67, 71, 84]] = [0, 1, 2, 3]
    np.take(lut, seq, out=seq)
    return seq


def prepare_database(snps: np.ndarray, snpsmap: np.ndarray, seed: int = 0) -> dict:
    """Arrays of ``{name}.database.hdf5`` as init_database builds them
    (reference tetrad/src/write_database.py:136-168)."""
    rng = np.random.default_rng(seed)
    nsnps = snps.shape[1]
    tmpmap = np.zeros((nsnps, 2), dtype=np.uint32)
    tmpmap[:, 0] = snpsmap[:, 0] - 1
    tmpmap[:, 1] = np.arange(nsnps)
    spans = get_spans(tmpmap[:, 0])
    seqarr = snps.astype(np.uint8).copy()
    seqarr[seqarr == 45] = 78
    tmparr = recode_acgt(resolve_ambigs(seqarr.copy(), rng))
    return dict(seqarr=seqarr, spans=spans, tmparr=tmparr, tmpmap=tmpmap)


def bootstrap_resample(seqarr: np.ndarray, spans: np.ndarray, rng: np.random.Generator):
    """One bootstrap replicate of (tmparr, tmpmap): loci drawn with replacement, columns of
    every drawn locus randomly permuted, ambiguities re-resolved, ACGT->0..3
    (semantics of resample_tmp_database + jit_resample, reference
    tetrad/src/run_inference.py:99-143, tetrad/jit/resample.py:20-64)."""
    nloci = spans.shape[0]
    lidxs = rng.choice(nloci, nloci, replace=True)
    lens = (spans[lidxs, 1] - spans[lidxs, 0]).astype(np.int64)
    total = int(lens.sum())
    inst = np.repeat(np.arange(nloci, dtype=np.int64), lens)
    # random order inside every instance: sort by (instance, random key)
    order = np.lexsort((rng.random(total), inst))
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])
    within = order - np.repeat(first, lens)
    cols = np.repeat(spans[lidxs, 0], lens) + within
    tmparr = seqarr[:, cols]
    tmpmap = np.zeros((total, 2), dtype=np.uint32)
    tmpmap[:, 0] = inst
    tmpmap[:, 1] = np.arange(total)
    tmparr = recode_acgt(resolve_ambigs(np.ascontiguousarray(tmparr), rng))
    return tmparr, tmpmap


def config_arrays(name: str, seed: int = 1234) -> dict:
    """Generate config ``name`` (C1..C5) and return its working arrays plus metadata."""
    cfg = CONFIGS[name]
    data = generate(cfg["ntips"], cfg["nloci"], cfg["nsnps"], seed=seed)
    out = prepare_database(data.snps, data.snpsmap, seed=seed)
    out.update(names=data.names, clades=data.clades, config=dict(cfg, name=name))
    return out
