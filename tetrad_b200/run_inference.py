"""Replicate loop around the hot path: the caller side of ``infer_resolved_quartets``.

Mirrors the reference's tetrad/src/run_inference.py: ``run_inference`` (:360-407),
``resample_tmp_database`` (:99-143), the TSV sink of ``distributor`` (:232-247),
``iter_qmc_formatted`` / ``write_qmc_format`` (:254-327), ``run_qmc`` (:146-166),
``relabel_tree`` (:169-181), ``infer_supertree`` (:330-357).  The ipyparallel process farm is
replaced by the GPU engine (``QuartetResolver``), optionally sharded over ranks
(``tetrad_b200.sharding``); wQMC itself (a prebuilt binary the reference ships in ``bin/``) is
invoked unchanged when it can be found, and skipped with a warning otherwise.

The numpy ``Generator`` stream is consumed in the reference's order, so for a given
``random_seed`` every replicate resolves the same quartet set and the same resampled loci as the
reference would; only the within-locus shuffles and ambiguity coin-flips, which the reference
draws from numba's private RNG, come from a different stream.
"""
from __future__ import annotations

import logging
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np

from . import combinations, database, synth
from .schema import Project, RNGStateModel

logger = logging.getLogger("tetrad_b200")

# split index -> tip order "a,b|c,d" (reference run_inference.py:41-45, 264-270)
TIDXS = np.array([[0, 1, 2, 3], [0, 2, 1, 3], [0, 3, 1, 2]], dtype=np.uint8)


# ----------------------------------------------------------------------------------------------
# bootstrap resampling (reference :99-143)
# ----------------------------------------------------------------------------------------------
def resample_arrays(seqarr: np.ndarray, spans: np.ndarray, rng: np.random.Generator):
    """One bootstrap replicate of (tmparr, tmpmap).  Draws from ``rng`` exactly what
    ``resample_tmp_database`` draws (:117 loci with replacement, :120 and :123 two 31-bit
    seeds); the two seeds feed the within-locus shuffle and the ambiguity resolution."""
    nloci = spans.shape[0]
    lidxs = rng.choice(nloci, nloci, replace=True)
    seed_shuffle = int(rng.integers(2**31))
    seed_ambig = int(rng.integers(2**31))
    lens = (spans[lidxs, 1] - spans[lidxs, 0]).astype(np.int64)
    total = int(lens.sum())
    inst = np.repeat(np.arange(nloci, dtype=np.int64), lens)
    sub = np.random.default_rng(seed_shuffle)
    order = np.lexsort((sub.random(total), inst))          # random order inside every instance
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])
    cols = np.repeat(spans[lidxs, 0], lens) + (order - np.repeat(first, lens))
    tmparr = np.ascontiguousarray(seqarr[:, cols])
    tmpmap = np.zeros((total, 2), dtype=np.uint32)
    tmpmap[:, 0] = inst                                     # resample.py:58
    tmpmap[:, 1] = np.arange(total)
    tmparr = synth.recode_acgt(synth.resolve_ambigs(tmparr, np.random.default_rng(seed_ambig)))
    return tmparr, tmpmap


# ----------------------------------------------------------------------------------------------
# result sink (reference :232-247)
# ----------------------------------------------------------------------------------------------
_SINK_CHUNK = 1 << 16          # rows formatted per C call
_sink_pool = None


def _format_parallel(call, nrows: int, row_bytes: int):
    """Yield, in row order, the bytes ``call(lo, hi, buf)`` formats for consecutive row ranges.  The C formatters
    run without the GIL (ctypes), so the ranges are formatted on a small thread pool: at C3/C4 sizes the host
    sink, not the GPU, was what a replicate waited for."""
    import ctypes
    import os
    from concurrent.futures import ThreadPoolExecutor
    global _sink_pool
    if _sink_pool is None:
        _sink_pool = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) - 1)),
                                        thread_name_prefix="tetrad-sink")

    def work(lo):
        hi = min(nrows, lo + _SINK_CHUNK)
        buf = ctypes.create_string_buffer(row_bytes * (hi - lo))
        n = call(lo, hi, buf)
        if n == -4:                                        # TQ_ENOMEM: a row holds a non-score (inf, 1e300, ...) whose
            buf = ctypes.create_string_buffer(1200 * (hi - lo))   # full decimal text is longer than the row budget
            n = call(lo, hi, buf)
        if n < 0:
            from . import _lib
            _lib.check(int(n))
        return ctypes.string_at(buf, n)

    starts = range(0, nrows, _SINK_CHUNK)
    if nrows <= _SINK_CHUNK:
        yield from (work(lo) for lo in starts)
    else:
        yield from _sink_pool.map(work, starts)


def write_quartets_tsv(path: Path, quartets, rstat, rscor, append: bool = False) -> None:
    """``t0 t1 t2 t3 s0 s1 s2 best nsnps``, tab-separated, scores ``%.6f``, no header: the line
    format pandas produces at run_inference.py:232-234.  Formatted in C (``tq_format_tsv``)."""
    from . import _lib
    lib = _lib.load()
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    rstat = np.ascontiguousarray(rstat, dtype=np.uint32).reshape(-1, 2)
    rscor = np.ascontiguousarray(rscor, dtype=np.float64).reshape(-1, 3)
    qp, sp, cp = quartets.ctypes.data, rstat.ctypes.data, rscor.ctypes.data

    def call(lo, hi, buf):
        return lib.tq_format_tsv(qp + 16 * lo, sp + 8 * lo, cp + 24 * lo, hi - lo, buf, len(buf))

    with open(path, "ab" if append else "wb") as out:
        for piece in _format_parallel(call, len(quartets), 160):
            out.write(piece)


def qmc_weights(quartets, rstat, rscor, weights: int, min_snps: int = 0, min_ratio: float = 1.0):
    """In-memory twin of ``iter_qmc_formatted``: (tips u32[n,4] in split order, weight f64[n]) of the
    quartets that pass the filters.  Scores are first rounded through the TSV's ``%.6f`` text
    (``tq_round_scores``) because the reference computes its weights from the parsed file."""
    from . import _lib
    if weights not in (0, 1, 2, 3):
        raise ValueError(f"no weight strategy {weights}")
    lib = _lib.load()
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    rstat = np.ascontiguousarray(rstat, dtype=np.uint32).reshape(-1, 2)
    rscor = np.ascontiguousarray(rscor, dtype=np.float64).reshape(-1, 3)
    # one pass in C (tq_qmc_weights); scores that are not finite non-negative numbers take the numpy path below
    tips_out = np.empty((len(quartets), 4), dtype=np.uint32)
    weight_out = np.empty(len(quartets), dtype=np.float64)
    n = lib.tq_qmc_weights(quartets.ctypes.data, rstat.ctypes.data, rscor.ctypes.data, len(quartets), int(weights),
                           int(min_snps), float(min_ratio), tips_out.ctypes.data, weight_out.ctypes.data)
    if n >= 0:
        return tips_out[:n], weight_out[:n]
    scores = np.empty_like(rscor)
    _lib.check(lib.tq_round_scores(rscor.ctypes.data, rscor.size, scores.ctypes.data))
    min_snps = max(1, min_snps)
    best = np.minimum(rstat[:, 0], 2)
    # tips in split order a,b|c,d = quartets[:, TIDXS[best]] (column selects instead of a fancy-index gather)
    t1, t2, t3 = quartets[:, 1], quartets[:, 2], quartets[:, 3]
    is0, is1 = best == 0, best == 1
    tips = np.empty_like(quartets)
    tips[:, 0] = quartets[:, 0]
    tips[:, 1] = np.where(is0, t1, np.where(is1, t2, t3))
    tips[:, 2] = np.where(is0, t2, t1)
    tips[:, 3] = np.where(is0 | is1, t3, t2)
    if not weights:
        weight = np.ones(len(quartets))
        ratio = np.ones(len(quartets))
    else:
        # the sorted scores (reference :280) without a generic sort: min, median and max of three, exactly
        a, b, c = scores[:, 0], scores[:, 1], scores[:, 2]
        lo_ab, hi_ab = np.minimum(a, b), np.maximum(a, b)
        smin = np.minimum(lo_ab, c)
        smax = np.maximum(hi_ab, c)
        smid = np.maximum(lo_ab, np.minimum(hi_ab, c))
        smean = (smid + smax) / 2.0                      # == srt[:, 1:].mean(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(smin == 0, 1.0, smean / np.where(smin == 0, 1.0, smin))
            third = 1.0 - smin / ((smin + smid) + smax)  # == srt.sum(axis=1), added in sorted order
        weight = {1: smean, 2: ratio, 3: third}[weights]
    keep = (rstat[:, 1] >= min_snps) & ~(ratio < min_ratio)
    return np.ascontiguousarray(tips[keep], dtype=np.uint32), np.ascontiguousarray(weight[keep], dtype=np.float64)


def format_qmc_lines(tips: np.ndarray, weight: np.ndarray, rows=None) -> bytes:
    """``a,b|c,d:%.5f`` lines (run_inference.py:305), formatted in C (``tq_format_qmc``); ``rows`` (an index
    array, e.g. the shuffle of :308-327) selects and orders the lines without copying the arrays."""
    from . import _lib
    lib = _lib.load()
    tips = np.ascontiguousarray(tips, dtype=np.uint32).reshape(-1, 4)
    weight = np.ascontiguousarray(weight, dtype=np.float64)
    tp, wp = tips.ctypes.data, weight.ctypes.data
    if rows is None:
        def call(lo, hi, buf):
            return lib.tq_format_qmc(tp + 16 * lo, wp + 8 * lo, hi - lo, buf, len(buf))
        n = len(tips)
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        if len(rows) and (rows.min() < 0 or rows.max() >= len(tips)):
            raise IndexError("format_qmc_lines: row index out of range")
        rp = rows.ctypes.data

        def call(lo, hi, buf):
            return lib.tq_format_qmc_rows(tp, wp, rp + 8 * lo, hi - lo, buf, len(buf))
        n = len(rows)
    return b"".join(_format_parallel(call, n, 128))


def read_quartets_tsv(path: Path):
    """(quartets i64[m,4], scores f64[m,3], best i64[m], nsnps i64[m]) parsed from the TSV."""
    if os.path.getsize(path) == 0:
        z = np.zeros((0, 9))
    else:
        z = np.loadtxt(path, delimiter="\t", ndmin=2)
    return z[:, 0:4].astype(np.int64), z[:, 4:7], z[:, 7].astype(np.int64), z[:, 8].astype(np.int64)


def iter_qmc_formatted(qrts_file: Path, weights: int, min_snps: int = 0, min_ratio: float = 1.0):
    """Yield resolved quartets in wQMC format ``a,b|c,d:weight`` (reference :254-305), computed
    from the TSV text like the reference does, but vectorised."""
    if weights not in (0, 1, 2, 3):
        raise ValueError(f"no weight strategy {weights}")
    min_snps = max(1, min_snps)
    quartets, scores, best, nsnps = read_quartets_tsv(qrts_file)
    order = TIDXS[best]                                          # :264-270
    tips = np.take_along_axis(quartets, order.astype(np.int64), axis=1)
    srt = np.sort(scores, axis=1)                                # :281
    if not weights:
        weight = np.ones(len(quartets))
        ratio = np.ones(len(quartets))
    else:
        smean = srt[:, 1:].mean(axis=1)                          # :282
        smin = srt[:, 0]                                         # :283
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(smin == 0, 1.0, smean / np.where(smin == 0, 1.0, smin))   # :284
            third = 1.0 - smin / srt.sum(axis=1)                 # :292
        weight = {1: smean, 2: ratio, 3: third}[weights]
    keep = (nsnps >= min_snps) & ~(ratio < min_ratio)            # :275, :300
    for (a, b, c, d), w in zip(tips[keep].tolist(), weight[keep].tolist()):
        yield "{},{}|{},{}:{:.5f}".format(a, b, c, d, w)


def write_qmc_format(qrts_file: Path, qmc_in_file: Path, weights: int = 0, min_snps: int = 0,
                     min_ratio: float = 1.0, rng=None, results=None) -> int:
    """Write the weighted quartets in random order (reference :308-327; the reference shells out
    to ``shuf``, which cannot be seeded -- here the order comes from ``rng``).  ``results`` =
    (quartets, rstat, rscor) still in memory skips re-reading the TSV; the output is identical."""
    if results is None:
        quartets, scores, best, nsnps = read_quartets_tsv(qrts_file)
        results = (quartets, np.stack([best, nsnps], axis=1), scores)
    tips, weight = qmc_weights(*results, weights, min_snps, min_ratio)
    perm = np.random.default_rng(rng).permutation(len(tips))
    with open(qmc_in_file, "wb") as out:
        out.write(format_qmc_lines(tips, weight, rows=perm))
    return len(tips)


# ----------------------------------------------------------------------------------------------
# supertree (reference :48-70, :146-181, :330-357)
# ----------------------------------------------------------------------------------------------
def get_qmc_bin():
    """Path of the wQMC ``max-cut-tree`` binary, or None.  Search order: $TETRAD_QMC_BIN,
    ``sys.prefix/bin`` (conda install, reference :56), ``PATH``."""
    cands = [os.environ.get("TETRAD_QMC_BIN"), str(Path(sys.prefix) / "bin" / "max-cut-tree"),
             shutil.which("max-cut-tree")]
    for cand in cands:
        if cand and Path(cand).exists():
            return Path(cand)
    return None


def run_qmc(qmc_in_file: Path, qmc_out_file: Path, use_weights: bool) -> None:
    """Run wQMC (reference :146-166)."""
    qmc = get_qmc_bin()
    if qmc is None:
        raise IOError("No wQMC binary found (set TETRAD_QMC_BIN)")
    cmd = [str(qmc), f"qrtt={qmc_in_file}", f"otre={qmc_out_file}", f"weights={'on' if use_weights else 'off'}"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if proc.returncode:
        raise RuntimeError(f"error in wQMC: {proc.stdout!r}")


_TIP = re.compile(r"(?<=[(,])\s*(\d+)\s*(?=[:,)])")


def relabel_tree(newick: str, samples: dict) -> str:
    """Replace numeric tip labels by sample names (reference :169-181, without toytree)."""
    newick = newick.strip()
    return _TIP.sub(lambda m: str(samples[int(m.group(1))]), newick)


def infer_supertree(proj: Project, idx: int, weights: int, min_snps: int = 0, min_ratio: float = 1.0,
                    rng=None, results=None):
    """Quartet supertree of replicate ``idx`` as newick, or None without a wQMC binary.  Unlike the
    reference (:337-351) the tree that is relabelled is the one wQMC wrote for THIS replicate."""
    qrts_file = proj.workdir / f"{proj.name}.quartets_{idx}.tsv"
    if idx:
        qmc_in = proj.qmc_in_file.parent / f"tmp_in{idx}"
        qmc_out = proj.qmc_out_file.parent / f"tmp_out{idx}"
    else:
        qmc_in, qmc_out = proj.qmc_in_file, proj.qmc_out_file
    write_qmc_format(qrts_file, qmc_in, weights, min_snps, min_ratio, rng=rng, results=results)
    if get_qmc_bin() is None:
        logger.warning("wQMC binary not found: quartets written, supertree skipped")
        if idx:
            qmc_in.unlink()          # bootstrap inputs are temporary in the reference too (:353-356)
        return None
    run_qmc(qmc_in, qmc_out, bool(weights))
    nwk = relabel_tree(Path(qmc_out).read_text(), proj.samples)
    if idx:
        qmc_in.unlink()
        qmc_out.unlink()
    return nwk


# ----------------------------------------------------------------------------------------------
# replicate loop (reference :360-407)
# ----------------------------------------------------------------------------------------------
def draw_bootstrap(spans: np.ndarray, rng: np.random.Generator):
    """The three draws ``resample_tmp_database`` takes from the project's generator
    (run_inference.py:117, :120, :123): (lidxs, seed_shuffle, seed_ambig)."""
    from .replicate import draw_bootstrap as _draw
    return _draw(spans, rng)


def agreed_seed(seed):
    """The seed every rank builds the project's generator from.  A seeded project: its seed.  An unseeded one (a JSON
    written by the reference, or by an older ``init``): fresh OS entropy -- drawn on rank 0 and broadcast when a
    process group is up, because every rank replays the same draws (loci, Philox seeds, quartet ranks) from its own
    copy of the generator and rank 0 writes the gathered shards under ITS quartet list."""
    import torch.distributed as dist
    if seed is not None:
        return int(seed)
    box = [int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0] >> np.uint64(1))]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def run_inference(proj: Project, ncores: int = 0, nboots: int = 0, device: int = 0, resolver=None,
                  device_bootstrap: bool = True, overlap_sink: bool = True, save_invariants: bool = False,
                  quartet_sampler: str | None = None) -> None:
    """Resolve replicate 0 (original data) and ``nboots`` bootstrap replicates, checkpointing
    after each one exactly like the reference (:405-407).  ``ncores`` is accepted for CLI
    compatibility and ignored: the work runs on the GPU(s).  ``resolver`` (tests) overrides the
    engine: any object with ``load(tmparr, tmpmap)`` and ``resolve(quartets, subsample)``.
    With ``device_bootstrap`` (default) and an engine that has ``load_bootstrap``, replicates >= 1
    are shuffled, ambiguity-resolved and packed on the GPU from the resident ``seqarr``; otherwise
    the host builds ``tmparr``/``tmpmap`` (``resample_arrays``) and the engine packs them.
    ``save_invariants`` keeps every replicate's site-pattern count matrices, the deprecated API's option
    (deprecated/distributor.py:231-257 stores them as ``invariants/boot{k}`` u16[nquartets,16,16] in its HDF5):
    here ``{name}.invariants_{rep}.npy``, u32[nquartets,16,16] (= ``mats[0]``; u32 because counts pass 65 535).

    ``quartet_sampler`` ("device" | "numpy", default: the project's, else "device"): how a random quartet sample
    (``nquartets < C(n,4)``) is drawn, see ``replicate.py``.

    The GPU side of a replicate is ``replicate.ReplicatePipeline`` (draw -> prepare -> resolve -> fetch, two engines
    alternating): replicate r + 1 is drawn, planned and queued on the device while replicate r resolves, and the
    sink of replicate r (TSV, wQMC input, wQMC, checkpoint) runs on worker threads beside both."""
    import time
    import torch.distributed as dist
    from .replicate import ReplicatePipeline
    if save_invariants and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        raise NotImplementedError("save_invariants is a single-process option (1 KB per quartet is not gathered)")

    arrays = database.read_arrays(proj.database_file, ["seqarr", "spans", "tmparr", "tmpmap"])
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    if proj.bootstrap_rng is None:
        logger.info("starting quartet tree inference")
        rng = np.random.default_rng(agreed_seed(proj.random_seed))
    else:
        logger.info("continuing quartet tree inference")
        rng = proj.bootstrap_rng.to_rng()

    pipe = ReplicatePipeline(arrays, proj.nsamples, proj.nqrts, proj.nqrts_total, proj.subsample_snps, device=device,
                             seed=proj.random_seed or 0, resolver=resolver, device_bootstrap=device_bootstrap,
                             sampler=quartet_sampler or getattr(proj, "quartet_sampler", None) or "device")

    from concurrent.futures import ThreadPoolExecutor
    # Two sink workers: the heavy part of a replicate's sink (TSV, wQMC input, wQMC itself) is independent of its
    # neighbours', so two replicates' sinks run side by side; each waits for its predecessor before the ordered tail
    # (tree files and the checkpoint), so replicates still become durable in order (reference :396-407).
    pool = ThreadPoolExecutor(max_workers=2) if overlap_sink else None
    pending = []

    draws = prep = out = None
    if proj.bootstrap_idx <= nboots:
        draws = pipe.draw(proj.bootstrap_idx, rng)
        prep = pipe.prepare(draws)
        out = pipe.resolve(prep, want_counts=save_invariants)          # asynchronous on the GPU
    while proj.bootstrap_idx <= nboots:
        t_start = time.perf_counter()
        idx = proj.bootstrap_idx
        rng_state = draws["state"]                                     # generator state after this replicate's draws
        # the next replicate is drawn, planned, packed and resolved behind this one before this one's results are read
        next_draws = next_prep = next_out = None
        if idx + 1 <= nboots:
            next_draws = pipe.draw(idx + 1, rng)
            next_prep = pipe.prepare(next_draws)
            next_out = pipe.resolve(next_prep, want_counts=save_invariants)
        t_next = time.perf_counter()
        got = pipe.fetch(prep, out)                                    # rank 0: this replicate's results on the host
        t_gpu = time.perf_counter()
        while len(pending) >= 2:
            pending.pop(0).result()                        # at most two replicates' sinks in flight
        nq = 0
        if rank == 0:
            quartets, rstat, rscor = got
            nq = len(quartets)
            if save_invariants:
                np.save(proj.workdir / f"{proj.name}.invariants_{idx}.npy",
                        out[2].cpu().numpy().view(np.uint32).reshape(-1, 16, 16))
            job = (idx, quartets, rstat, rscor, rng_state)
            if pool is not None:
                pending.append(pool.submit(_sink_replicate, proj, *job, after=pending[-1] if pending else None))
            else:
                _sink_replicate(proj, *job)
        logger.info(f"finished rep {idx}: {nq} quartets, host part of the next replicate {t_next - t_start:.3f} s, "
                    f"replicate wall {t_gpu - t_start:.3f} s ({nq / max(t_gpu - t_start, 1e-9):.3g} quartets/s), "
                    f"hand-off to sink {time.perf_counter() - t_gpu:.3f} s")
        proj.bootstrap_idx += 1
        proj.bootstrap_rng = rng_state
        draws, prep, out = next_draws, next_prep, next_out
    for fut in pending:
        fut.result()
    if pool is not None:
        pool.shutdown()
    pipe.close()


def _sink_replicate(proj: Project, idx: int, quartets, rstat, rscor, rng_state, after=None) -> None:
    """Everything that happens to a replicate's results on the host: TSV, wQMC input, supertree,
    tree files, and the JSON checkpoint that makes the replicate durable (reference :396-407).
    ``after``: the future of the previous replicate's sink; the ordered tail waits for it."""
    snap = proj.model_copy(update=dict(bootstrap_idx=idx))
    write_quartets_tsv(snap.qrts_file, quartets, rstat, rscor)
    nwk = infer_supertree(snap, idx, proj.weights, rng=idx, results=(quartets, rstat, rscor))
    if after is not None:
        after.result()                                     # tree files and checkpoints in replicate order
    if nwk is not None:
        if idx == 0:
            proj.best_file.write_text(nwk + "\n")
        else:
            with open(proj.boots_file, "a") as outfh:
                outfh.write(nwk + "\n")
    done = proj.model_copy(update=dict(bootstrap_idx=idx + 1, bootstrap_rng=rng_state))
    done.save_json()
