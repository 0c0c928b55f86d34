#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (eaton-lab/tetrad).

Run in the build container only (``python oracle/gen_golden.py``): it imports the
reference's own functions from /root/reference (pure Python + numba; ``h5py`` is absent
from the image and only needed by the HDF5 wrapper, so it is stubbed) and freezes their
inputs and outputs.  /root/reference does not exist on the GPU box, so tests read only the
committed ``.npz`` files.  TEST INFRASTRUCTURE ONLY.

Reference functions executed (paths relative to /root/reference):
  tetrad/src/resolve_quartets.py:191  new_infer_resolved_quartets  -> rstat, rscor
  tetrad/src/resolve_quartets.py:42   subsample_chunk_to_matrices  -> mats (subsample mode)
  tetrad/src/resolve_quartets.py:76   full_chunk_to_matrices       -> mats (full mode)
  tetrad/jit/resample.py:20           jit_resample                 -> bootstrap tmparr/tmpmap
  tetrad/jit/resolve_ambigs.py:12     jit_resolve_ambigs
  tetrad/jit/get_spans.py:11          jit_get_spans
  tetrad/src/combinations.py:94       _index_to_combination
"""
from __future__ import annotations

import sys
import types
from itertools import combinations
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("h5py", types.ModuleType("h5py"))

from tetrad.src import resolve_quartets as ref            # noqa: E402
from tetrad.jit.resample import jit_resample              # noqa: E402
from tetrad.jit.resolve_ambigs import jit_resolve_ambigs  # noqa: E402
from tetrad.jit.get_spans import jit_get_spans            # noqa: E402
from tetrad.src.combinations import _index_to_combination  # noqa: E402

from tetrad_b200 import synth                             # noqa: E402
from oracle import oracle                                 # noqa: E402

GOLD = ROOT / "tests" / "golden"


def ref_mats(tmparr, tmpmap, quartets, subsample):
    """Reference count matrices, built the way new_infer_resolved_quartets does (:210-223)."""
    out = np.zeros((len(quartets), 3, 16, 16), dtype=np.uint32)
    for i, sidx in enumerate(quartets):
        seqs = tmparr[sidx, :]
        nmask0 = np.sum(seqs >= 78, axis=0)
        nmask1 = np.sum(seqs == seqs[0], axis=0) == 4
        fn = ref.subsample_chunk_to_matrices if subsample else ref.full_chunk_to_matrices
        out[i] = fn(seqs, tmpmap[:, 0], nmask0 + nmask1)
    return out


def freeze(name, tmparr, tmpmap, quartets, extra=None):
    quartets = np.asarray(quartets, dtype=np.uint32)
    rec = dict(tmparr=tmparr.astype(np.uint8), tmpmap=tmpmap.astype(np.uint32), quartets=quartets)
    for subsample in (True, False):
        tag = "sub" if subsample else "full"
        np.random.seed(7)  # the reference's zero-SNP branch draws from the global RNG
        _, rstat, rscor = ref.new_infer_resolved_quartets(tmparr, tmpmap, quartets, subsample)
        mats = ref_mats(tmparr, tmpmap, quartets, subsample)
        # the reference exposes no selection mask: restate it, then prove it by recounting
        omats, sel = oracle.counts(tmparr, tmpmap, quartets, subsample, want_mask=True)
        assert (omats == mats).all(), f"{name}/{tag}: oracle counts != reference counts"
        for q in range(min(len(quartets), 64)):
            cols = tmparr[quartets[q]][:, sel[q].astype(bool)].astype(np.int64)
            idx = ((cols[0] * 4 + cols[1]) * 4 + cols[2]) * 4 + cols[3]
            assert (np.bincount(idx, minlength=256).reshape(16, 16) == mats[q, 0]).all()
        ranks = np.array([[np.linalg.matrix_rank(mats[q, k].astype(np.float64)) for k in range(3)]
                          for q in range(len(quartets))], dtype=np.int64)
        rec[f"{tag}_rstat"] = rstat
        rec[f"{tag}_rscor"] = rscor
        rec[f"{tag}_mats0"] = mats[:, 0].copy()       # mats[1], mats[2] are transposes of it
        rec[f"{tag}_ranks"] = ranks
        rec[f"{tag}_selmask"] = np.packbits(sel, axis=1, bitorder="little")
        assert (mats[:, 1] == mats[:, 0].reshape(-1, 4, 4, 4, 4).transpose(0, 1, 3, 2, 4).reshape(-1, 16, 16)).all()
        assert (mats[:, 2] == mats[:, 0].reshape(-1, 4, 4, 4, 4).transpose(0, 1, 4, 2, 3).reshape(-1, 16, 16)).all()
        print(f"{name:10s} {tag:4s} q={len(quartets):4d} S={tmparr.shape[1]:6d} "
              f"nsnps[min,max]=({rstat[:, 1].min()},{rstat[:, 1].max()}) "
              f"minrank={sorted(set(np.minimum(10, ranks.min(1)).tolist()))}")
    if extra:
        rec.update(extra)
    np.savez_compressed(GOLD / f"{name}.npz", **rec)


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    import numba
    versions = dict(numpy=np.__version__, numba=numba.__version__)

    # (1) noisy tree-like data, all 495 quartets of 12 tips: rank-16 regime, minrank = 10
    d = synth.generate(12, 600, 2400, seed=1234)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=1)
    allq = np.array(list(combinations(range(12), 4)), dtype=np.uint32)
    freeze("noisy12", db["tmparr"], db["tmpmap"], allq)

    # (2) clean data (no homoplasy, no missing): true split has rank < 10  -> minrank < 10
    d = synth.generate(10, 400, 1600, seed=5, homoplasy=0.0, p_missing=0.0, p_gap=0.0, p_ambig=0.0)
    db2 = synth.prepare_database(d.snps, d.snpsmap, seed=2)
    q10 = np.array(list(combinations(range(10), 4)), dtype=np.uint32)
    freeze("clean10", db2["tmparr"], db2["tmpmap"], q10)

    # (3) tiny data: 50 loci -> ranks 6..15
    d = synth.generate(8, 50, 170, seed=9)
    db3 = synth.prepare_database(d.snps, d.snpsmap, seed=3)
    q8 = np.array(list(combinations(range(8), 4)), dtype=np.uint32)
    freeze("tiny8", db3["tmparr"], db3["tmpmap"], q8)

    # (4) one all-missing tip and one all-invariant pair -> zero-SNP quartets (scores 0.001)
    d = synth.generate(7, 80, 300, seed=11)
    db4 = synth.prepare_database(d.snps, d.snpsmap, seed=4)
    db4["tmparr"][3, :] = 78
    q7 = np.array(list(combinations(range(7), 4)), dtype=np.uint32)
    freeze("missing7", db4["tmparr"], db4["tmpmap"], q7)

    # (5) a bootstrap replicate built by the REFERENCE's resampler (within-locus shuffles,
    #     repeated loci, tmpmap[:,0] = instance index), run_inference.py:117-136
    rng = np.random.default_rng(42)
    seqarr, spans = db["seqarr"], db["spans"]
    assert (jit_get_spans(db["tmpmap"].astype(np.int64)) == spans).all()
    lidxs = rng.choice(spans.shape[0], spans.shape[0], replace=True)
    tmparr, tmpmap = jit_resample(seqarr, spans, lidxs, seed=int(rng.integers(2**31)))
    tmparr = jit_resolve_ambigs(tmparr, seed=int(rng.integers(2**31)))
    for code, idx in ((65, 0), (67, 1), (71, 2), (84, 3)):
        tmparr[tmparr == code] = idx
    freeze("boot12", tmparr, tmpmap, allq[::5],
           extra=dict(seqarr=seqarr, spans=spans, lidxs=lidxs))

    # (6) hand-checkable micro case: 4 tips x 8 columns, 3 loci, one missing, one invariant
    #     cols:        0  1  2  3  4  5  6  7
    micro = np.array([[0, 0, 78, 1, 2, 3, 3, 0],
                      [0, 1, 1, 1, 2, 3, 0, 0],
                      [0, 0, 2, 1, 3, 3, 0, 1],
                      [0, 1, 3, 1, 3, 2, 3, 1]], dtype=np.uint8)
    mmap = np.zeros((8, 2), dtype=np.uint32)
    mmap[:, 0] = [0, 0, 0, 1, 1, 1, 2, 2]
    mmap[:, 1] = np.arange(8)
    freeze("micro4", micro, mmap, np.array([[0, 1, 2, 3]], dtype=np.uint32))

    # (7) quartet unranking, combinations.py:94-106
    idx = np.array([0, 1, 2, 494, 100, 333, 17, 250], dtype=np.int64)
    unr12 = np.array([_index_to_combination(int(i), 12) for i in idx], dtype=np.uint32)
    big = np.array([0, 1, 64684949, 12345678, 40000000, 999], dtype=np.int64)
    unr200 = np.array([_index_to_combination(int(i), 200) for i in big], dtype=np.uint32)
    np.savez_compressed(GOLD / "unrank.npz", idx12=idx, comb12=unr12, idx200=big, comb200=unr200)

    # (8) result sink + wQMC formatting: the reference's own pandas line (run_inference.py:232-234)
    #     and iter_qmc_formatted (:254-305), executed on the noisy12 results
    import types as _types
    for mod in ("ipyparallel", "toytree"):
        sys.modules.setdefault(mod, _types.ModuleType(mod))
    sys.modules["ipyparallel"].Client = object
    clus = _types.ModuleType("tetrad.src.cluster"); clus.Cluster = object
    sys.modules["tetrad.src.cluster"] = clus
    import pandas as pd
    from tetrad.src import run_inference as ref_run
    import tempfile
    g = np.load(GOLD / "noisy12.npz")
    rqrts, rstat, rscor = g["quartets"], g["sub_rstat"], g["sub_rscor"]
    tabular = pd.concat([pd.DataFrame(i) for i in (rqrts, rscor, rstat)], axis=1)
    tsv = tabular.to_csv(sep="\t", float_format='%.6f', index=False, header=False)
    qmc = {}
    with tempfile.TemporaryDirectory() as tmp:
        f = Path(tmp) / "q.tsv"
        f.write_text(tsv)
        for weights in (0, 1, 2, 3):
            qmc[f"w{weights}"] = "\n".join(ref_run.iter_qmc_formatted(f, weights))
        qmc["w1_minsnps400"] = "\n".join(ref_run.iter_qmc_formatted(f, 1, min_snps=400))
        qmc["w2_minratio5"] = "\n".join(ref_run.iter_qmc_formatted(f, 2, min_ratio=5.0))
    np.savez_compressed(GOLD / "sink.npz", tsv=np.array(tsv), **{k: np.array(v) for k, v in qmc.items()})
    print("sink:", len(tsv.splitlines()), "tsv lines;", {k: len(v.splitlines()) for k, v in qmc.items()})

    (GOLD / "VERSIONS.txt").write_text(
        "golden vectors generated by oracle/gen_golden.py executing /root/reference with\n"
        + "\n".join(f"{k}={v}" for k, v in versions.items()) + "\n")
    print("wrote", sorted(p.name for p in GOLD.glob("*")))


if __name__ == "__main__":
    main()
