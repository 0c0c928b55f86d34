"""TEST INFRASTRUCTURE ONLY -- CPU oracle of tetrad's per-quartet resolve loop.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``tetrad_b200/`` does.

What it restates (paths relative to /root/reference):

* ``new_infer_resolved_quartets``  tetrad/src/resolve_quartets.py:191-265
  - integer half (gather, masks, selection, counts, flattenings) in C:
    ``oracle/quartet_oracle.c`` (``counts``), with a pure-numpy twin (``counts_numpy``)
    used to cross-check the C on small cases;
  - floating half exactly as the reference calls it: ``np.linalg.svd(M)[1]`` (:242) and
    ``np.linalg.matrix_rank(M)`` (:243), i.e. LAPACK dgesdd from the numpy in this image
    (numpy 2.3.5 / scipy-openblas 0.3.30 -- the reference leaves numpy unpinned,
    setup.py:45), ``minrank = min(10, min rank)`` (:246), tail 2-norm (:248),
    first ``argmin`` (:251).
* zero-SNP quartets: the reference draws ``np.random.randint(3)`` from the unseeded global
  RNG (:230-232).  That cannot be reproduced, so both this oracle and the CUDA path draw
  the split from a counter-based hash of (seed, tip ids) -- ``tiebreak_split``.

Parity status: **pinned against the executed reference** -- ``oracle/gen_golden.py`` runs the
reference's own functions (imported from /root/reference with an ``h5py`` stub) on seeded
inputs and freezes inputs + outputs under ``tests/golden``; ``tests/test_oracle.py`` holds
this oracle to those vectors bit-for-bit (counts, rstat) and to 1e-12 (scores).  The
reference ships no tests or golden vectors of its own.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None
_MASK64 = (1 << 64) - 1


def build(force: bool = False) -> Path:
    """Compile oracle/quartet_oracle.c -> oracle/liboracle.so (gcc)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "quartet_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(
            ["gcc", "-O3", "-fPIC", "-std=c11", "-shared", "-o", str(so), str(src)])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(str(build()))
        lib.tq_oracle_counts.restype = ctypes.c_int
        lib.tq_oracle_counts.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def counts(tmparr, tmpmap, quartets, subsample_snps: bool, want_mask: bool = False):
    """u32[m,3,16,16] site-pattern matrices (+ optional u8[m,S] selected-column mask)."""
    tmparr = np.ascontiguousarray(tmparr, dtype=np.uint8)
    tmpmap = np.ascontiguousarray(tmpmap, dtype=np.uint32)
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    n, ncols = tmparr.shape
    if tmpmap.ndim == 1:
        stride = 1
    else:
        stride = tmpmap.shape[1]
    if tmpmap.shape[0] != ncols:
        raise ValueError("tmpmap and tmparr disagree on the number of columns")
    m = quartets.shape[0]
    mats = np.zeros((m, 3, 16, 16), dtype=np.uint32)
    sel = np.zeros((m, ncols), dtype=np.uint8) if want_mask else None
    rc = _lib().tq_oracle_counts(
        tmparr.ctypes.data, n, ncols, tmparr.strides[0],
        tmpmap.ctypes.data, stride,
        quartets.ctypes.data, m, int(bool(subsample_snps)),
        mats.ctypes.data, sel.ctypes.data if want_mask else None)
    if rc != 0:
        raise IndexError("quartet tip index out of range")
    return (mats, sel) if want_mask else mats


def counts_numpy(tmparr, tmpmap, quartets, subsample_snps: bool, want_mask: bool = False):
    """Pure-numpy twin of ``counts`` (slow; cross-check for the C on small cases)."""
    tmparr = np.asarray(tmparr, dtype=np.uint8)
    locus = np.asarray(tmpmap)[:, 0] if np.asarray(tmpmap).ndim == 2 else np.asarray(tmpmap)
    quartets = np.asarray(quartets, dtype=np.int64).reshape(-1, 4)
    m, ncols = quartets.shape[0], tmparr.shape[1]
    mats = np.zeros((m, 3, 16, 16), dtype=np.uint32)
    sel = np.zeros((m, ncols), dtype=np.uint8)
    for q in range(m):
        seqs = tmparr[quartets[q]]                                   # :212
        valid = (seqs <= 3).all(axis=0)                              # :216 (+ header note)
        invariant = (seqs == seqs[0]).all(axis=0)                    # :218
        cand = valid & ~invariant
        if subsample_snps:
            # state machine of :58-64: keep c iff cand and locus[c] != locus[last kept]
            keep = np.zeros(ncols, dtype=bool)
            last = np.uint32(0xFFFFFFFF)
            for c in np.flatnonzero(cand):
                if locus[c] != last:
                    keep[c] = True
                    last = locus[c]
        else:
            keep = cand
        cols = seqs[:, keep].astype(np.int64)
        idx = ((cols[0] * 4 + cols[1]) * 4 + cols[2]) * 4 + cols[3]
        tens = np.bincount(idx, minlength=256).astype(np.uint32).reshape(4, 4, 4, 4)
        mats[q, 0] = tens.reshape(16, 16)                            # (ab|cd)
        mats[q, 1] = tens.transpose(0, 2, 1, 3).reshape(16, 16)      # (ac|bd)  :67-72
        mats[q, 2] = tens.transpose(0, 3, 1, 2).reshape(16, 16)      # (ad|bc)
        sel[q] = keep
    return (mats, sel) if want_mask else mats


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK64
    return x ^ (x >> 31)


def tiebreak_split(quartet, seed: int = 0) -> int:
    """Split index in {0,1,2} for a quartet with no informative SNP (replaces the
    reference's unseeded np.random.randint(3), resolve_quartets.py:231)."""
    h = seed & _MASK64
    for t in quartet:
        h = _splitmix64(h ^ int(t))
    return int(h % 3)


def scores_from_mats(mats: np.ndarray):
    """(svals f64[m,3,16], ranks i64[m,3], minrank i64[m], scores f64[m,3]) as :241-248."""
    m = mats.shape[0]
    svals = np.zeros((m, 3, 16))
    ranks = np.zeros((m, 3), dtype=np.int64)
    for q in range(m):
        for k in range(3):
            mat = mats[q, k].astype(np.float64)
            svals[q, k] = np.linalg.svd(mat)[1]                      # :242
            ranks[q, k] = np.linalg.matrix_rank(mat)                 # :243
    minrank = np.minimum(10, ranks.min(axis=1))                      # :246
    scores = np.zeros((m, 3))
    for q in range(m):
        scores[q] = np.sqrt(np.sum(svals[q, :, minrank[q]:] ** 2, axis=1))   # :248
    return svals, ranks, minrank, scores


def new_infer_resolved_quartets(tmparr, tmpmap, quartets, subsample_snps: bool, seed: int = 0,
                                return_details: bool = False):
    """Oracle of reference ``new_infer_resolved_quartets`` (resolve_quartets.py:191-265):
    returns (quartets u32[m,4], rstat u32[m,2] = (best, nsnps), rscor f64[m,3])."""
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    mats = counts(tmparr, tmpmap, quartets, subsample_snps)
    m = quartets.shape[0]
    rscor = np.zeros((m, 3), dtype=np.float64)
    rstat = np.zeros((m, 2), dtype=np.uint32)
    nsnps = mats[:, 0].reshape(m, -1).sum(axis=1, dtype=np.uint64)   # :226
    svals, ranks, minrank, scores = scores_from_mats(mats)
    for q in range(m):
        if nsnps[q] == 0:                                            # :230-232
            rstat[q, 0] = tiebreak_split(quartets[q], seed)
            rscor[q] = 0.001
        else:
            rscor[q] = scores[q]
            rstat[q, 0] = np.argmin(rscor[q])                        # :251
        rstat[q, 1] = nsnps[q]                                       # :264
    if return_details:
        return quartets, rstat, rscor, dict(mats=mats, svals=svals, ranks=ranks, minrank=minrank)
    return quartets, rstat, rscor


def _pool_worker(args):
    lo, hi = args
    g = _POOL_STATE
    return new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], g["quartets"][lo:hi],
                                       g["subsample"], g["seed"])


_POOL_STATE: dict = {}


def resolve_parallel(tmparr, tmpmap, quartets, subsample_snps: bool, seed: int = 0,
                     nproc: int | None = None):
    """``new_infer_resolved_quartets`` over contiguous quartet chunks in a fork pool
    (the reference's engine farm without ipyparallel; one BLAS thread per worker as
    tetrad/deprecated/threadfix.py intends).  Used by bench.py's CPU legs."""
    import multiprocessing as mp
    nproc = nproc or os.cpu_count() or 1
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    m = quartets.shape[0]
    if nproc == 1 or m < 2 * nproc:
        return new_infer_resolved_quartets(tmparr, tmpmap, quartets, subsample_snps, seed)
    _POOL_STATE.update(tmparr=tmparr, tmpmap=tmpmap, quartets=quartets,
                       subsample=subsample_snps, seed=seed)
    edges = np.linspace(0, m, nproc + 1).astype(int)
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(var, "1")
    with mp.get_context("fork").Pool(nproc) as pool:
        parts = pool.map(_pool_worker, list(zip(edges[:-1], edges[1:])))
    _POOL_STATE.clear()
    return (quartets,
            np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))
