"""GPU (-m gpu): the CUDA path, called through the C ABI, against the golden vectors of the
executed reference and against the oracle on seeded inputs.  Nothing here reads /root/reference."""
import ctypes

import numpy as np
import pytest

from conftest import GOLDEN_CASES, assert_best_matches, assert_scores_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tetrad_b200
    return tetrad_b200


@pytest.fixture(params=["fused", "mma"])
def kernel(request, tb):
    """Every resolve-kernel variant behind launch_resolve (tq_resolve.cu): the fused one-warp-per-quartet POPC kernel
    and the tensor-core (IMMA) count kernel.  The default rule picks by launch size, so small test launches would
    otherwise only ever see the fused kernel.  (Full mode has a fourth: the tip-pair GEMM, test_pairgemm_* below.)"""
    from tetrad_b200 import _lib
    _lib.set_kernel(request.param)
    yield request.param
    _lib.set_kernel("auto")
    _lib.check(_lib.load().tq_set_option(b"mma_flush_every", 0))


def last_kernel():
    from tetrad_b200 import _lib
    return _lib.load().tq_last_kernel().decode()


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("mode", ["sub", "full"])
def test_golden_host_entry(tb, kernel, case, mode):
    """new_infer_resolved_quartets (tq_resolve_host) == reference outputs."""
    g = load_golden(case)
    sub = mode == "sub"
    q, rstat, rscor, extra = tb.new_infer_resolved_quartets(
        g["tmparr"], g["tmpmap"], g["quartets"], sub, return_counts=True)
    assert (q == g["quartets"]).all()
    assert (extra["counts"] == g[f"{mode}_mats0"]).all(), "counts not bit-exact"
    ref_stat, ref_scor = g[f"{mode}_rstat"], g[f"{mode}_rscor"]
    assert (rstat[:, 1] == ref_stat[:, 1]).all()
    assert_scores_close(rscor, ref_scor, ref_stat[:, 1])
    assert_best_matches(rstat[:, 0], ref_scor, ref_stat)
    has = ref_stat[:, 1] > 0
    ref_minrank = np.minimum(10, g[f"{mode}_ranks"].min(axis=1))
    assert (extra["minrank"][has] == ref_minrank[has]).all(), "minrank (matrix_rank) mismatch"


@pytest.mark.parametrize("case", ["noisy12", "boot12"])
def test_golden_device_engine(tb, case):
    """QuartetResolver (tq_plan_create + tq_pack + tq_resolve on torch memory)."""
    g = load_golden(case)
    eng = tb.QuartetResolver(device=0)
    eng.load(g["tmparr"], g["tmpmap"])
    for mode, sub in (("sub", True), ("full", False)):
        rstat, rscor, counts, minrank = eng.resolve(g["quartets"], sub, want_counts=True)
        eng.check_status()
        assert (counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16) == g[f"{mode}_mats0"]).all()
        ref_stat, ref_scor = g[f"{mode}_rstat"], g[f"{mode}_rscor"]
        assert_scores_close(rscor.cpu().numpy(), ref_scor, ref_stat[:, 1])
        assert_best_matches(rstat[:, 0].cpu().numpy(), ref_scor, ref_stat)
    eng.close()


def test_zero_snp_quartets(tb, oracle, kernel):
    g = load_golden("missing7")
    for seed in (0, 5):
        _, rstat, rscor = tb.new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], g["quartets"], True, seed=seed)
        _, ostat, oscor = oracle.new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], g["quartets"], True, seed=seed)
        zero = ostat[:, 1] == 0
        assert zero.any() and (rscor[zero] == 0.001).all()
        assert (rstat[zero] == ostat[zero]).all()          # same counter-based draw as the oracle


@pytest.mark.parametrize("shape", [(9, 300, 1100), (16, 2500, 9000), (6, 2, 5), (5, 40, 40)])
@pytest.mark.parametrize("sub", [True, False])
def test_seeded_vs_oracle(tb, oracle, kernel, shape, sub):
    from itertools import combinations
    from tetrad_b200 import synth
    n, nloci, nsnps = shape
    d = synth.generate(n, nloci, nsnps, seed=100 + n)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=n)
    quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
    if len(quartets) > 400:
        quartets = quartets[np.random.default_rng(1).choice(len(quartets), 400, replace=False)]
    _, rstat, rscor, extra = tb.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, sub, return_counts=True)
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, sub, return_details=True)
    assert (extra["counts"] == det["mats"][:, 0]).all()
    assert (rstat[:, 1] == ostat[:, 1]).all()
    assert_scores_close(rscor, oscor, ostat[:, 1])
    assert_best_matches(rstat[:, 0], oscor, ostat)
    has = ostat[:, 1] > 0
    assert (extra["minrank"][has] == det["minrank"][has]).all()
    assert {"fused": "resolve_kernel", "mma": "resolve_mma_kernel"}[kernel] in last_kernel()


def test_bootstrap_replicate_vs_oracle(tb, oracle):
    from tetrad_b200 import synth
    d = synth.generate(10, 1500, 6000, seed=77)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=3)
    rng = np.random.default_rng(9)
    tmparr, tmpmap = synth.bootstrap_resample(db["seqarr"], db["spans"], rng)
    quartets = np.array([[0, 1, 2, 3], [0, 4, 7, 9], [2, 3, 5, 8], [1, 5, 6, 9]], dtype=np.uint32)
    _, rstat, rscor, extra = tb.new_infer_resolved_quartets(tmparr, tmpmap, quartets, True, return_counts=True)
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(tmparr, tmpmap, quartets, True, return_details=True)
    assert (extra["counts"] == det["mats"][:, 0]).all()
    assert_scores_close(rscor, oscor, ostat[:, 1])


def test_edge_cases(tb, kernel):
    g = load_golden("tiny8")
    # empty quartet list
    q, rstat, rscor = tb.new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], np.zeros((0, 4), np.uint32), True)
    assert rstat.shape == (0, 2) and rscor.shape == (0, 3)
    # tip index out of range is an error, not a crash
    from tetrad_b200._lib import TetradB200Error
    with pytest.raises(TetradB200Error):
        tb.new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], np.array([[0, 1, 2, 99]], np.uint32), True)
    # locus ids split in two runs
    bad = g["tmpmap"].copy()
    bad[-1, 0] = bad[0, 0]
    with pytest.raises(TetradB200Error):
        tb.new_infer_resolved_quartets(g["tmparr"], bad, g["quartets"][:2], True)
    # unresolved byte (neither 0..3 nor >= 78) counts as missing
    arr = g["tmparr"].copy()
    arr[0, :] = 66
    _, rstat, _ = tb.new_infer_resolved_quartets(arr, g["tmpmap"], np.array([[0, 1, 2, 3]], np.uint32), True)
    assert rstat[0, 1] == 0
    # device engine flags bad tips without crashing
    eng = tb.QuartetResolver(device=0).load(g["tmparr"], g["tmpmap"])
    eng.resolve(np.array([[0, 1, 2, 200]], np.uint32))
    with pytest.raises(IndexError):
        eng.check_status()


def test_many_tiles_and_long_loci(tb, oracle, kernel):
    """> 1 tile (1024 loci each), one very long locus, loci of length 1."""
    rng = np.random.default_rng(4)
    lens = np.concatenate([[70], rng.integers(1, 7, size=2600), np.ones(300, dtype=np.int64)])
    rng.shuffle(lens)
    locus = np.repeat(np.arange(lens.size), lens)
    ncols = locus.size
    tmparr = rng.integers(0, 4, size=(6, ncols)).astype(np.uint8)
    tmparr[rng.random(tmparr.shape) < 0.15] = 78
    tmpmap = np.stack([locus, np.arange(ncols)], 1).astype(np.uint32)
    quartets = np.array([[0, 1, 2, 3], [1, 2, 4, 5], [0, 3, 4, 5]], dtype=np.uint32)
    for sub in (True, False):
        _, rstat, rscor, extra = tb.new_infer_resolved_quartets(tmparr, tmpmap, quartets, sub, return_counts=True)
        _, ostat, oscor, det = oracle.new_infer_resolved_quartets(tmparr, tmpmap, quartets, sub, return_details=True)
        assert (extra["counts"] == det["mats"][:, 0]).all()
        assert_scores_close(rscor, oscor, ostat[:, 1])


def test_size_independent_properties_at_bench_shape(tb):
    """BASELINE config C2 shape (50 tips, 50k loci, 200k SNPs): properties that need no oracle."""
    from tetrad_b200 import synth
    db = synth.config_arrays("C2")
    rng = np.random.default_rng(0)
    quartets = np.sort(np.stack([rng.choice(50, 4, replace=False) for _ in range(3000)]), axis=1).astype(np.uint32)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    rstat, rscor, counts, minrank = eng.resolve(quartets, True, want_counts=True)
    frstat, frscor, fcounts, _ = eng.resolve(quartets, False, want_counts=True)
    c = counts.cpu().numpy().view(np.uint32).astype(np.int64); fc = fcounts.cpu().numpy().view(np.uint32).astype(np.int64)
    nloci = eng.info["nloci"]
    assert (c.sum(1) == rstat[:, 1].cpu().numpy()).all() and (c.sum(1) <= nloci).all()
    assert (fc >= c).all() and (fc.sum(1) <= 200_000).all()
    for a in range(4):
        assert (c[:, 85 * a] == 0).all() and (fc[:, 85 * a] == 0).all()
    # relabelling tips permutes the count tensor: swap t0<->t1 and t2<->t3 => T[b,a,d,c]
    perm = quartets[:, [1, 0, 3, 2]].copy()
    _, _, pc, _ = eng.resolve(perm, True, want_counts=True)
    t = c.reshape(-1, 4, 4, 4, 4)
    assert (pc.cpu().numpy().view(np.uint32).reshape(-1, 4, 4, 4, 4) == t.transpose(0, 2, 1, 4, 3)).all()
    # scores: finite, non-negative, Frobenius identity  sum_k sigma^2 = ||T||_F^2  bounds the tail
    s = rscor.cpu().numpy()
    assert np.isfinite(s).all() and (s >= 0).all()
    assert (s.max(1) ** 2 <= (c.astype(np.float64) ** 2).sum(1) * (1 + 1e-12)).all()
    # idempotence
    rstat2, rscor2 = eng.resolve(quartets, True)
    assert (rstat2 == rstat).all() and (rscor2 == rscor).all()


def test_tetrad_facade_and_cli_end_to_end(tb, oracle, tmp_path):
    """Tetrad(name, data, nquartets, nboots).run() and `tetrad init` + `tetrad run`: TSV rows equal
    the oracle's resolution of the same replicate arrays."""
    from tetrad_b200 import cli, database, run_inference, synth
    from tetrad_b200.schema import Project
    d = synth.generate(8, 300, 1200, seed=21)
    database.write_snps_input(tmp_path / "in.npz", d.snps, d.snpsmap, d.names)
    tet = tb.Tetrad("fac", str(tmp_path / "in.npz"), workdir=str(tmp_path / "w"), nquartets=0, nboots=1, seed=7)
    tet.run()
    proj = Project.load_json(tmp_path / "w" / "fac.json")
    assert proj.bootstrap_idx == 2 and proj.nqrts == 70
    q, scores, best, nsnps = run_inference.read_quartets_tsv(tmp_path / "w" / "fac.quartets_0.tsv")
    arrays = database.read_arrays(proj.database_file, ["tmparr", "tmpmap"])
    _, ostat, oscor = oracle.new_infer_resolved_quartets(arrays["tmparr"], arrays["tmpmap"], q.astype(np.uint32), True)
    assert (nsnps == ostat[:, 1]).all() and (best == ostat[:, 0]).all()
    np.testing.assert_allclose(scores, oscor, rtol=1e-5, atol=1e-6)      # TSV carries 6 decimals
    assert (tmp_path / "w" / "fac.quartets_1.tsv").exists() and (tmp_path / "w" / "fac.qmc_in.txt").exists()
    # save_invariants (deprecated API option): the count matrices of every replicate
    tet2 = tb.Tetrad("inv", str(tmp_path / "in.npz"), workdir=str(tmp_path / "w2"), nquartets=0, nboots=0, seed=7,
                     save_invariants=True)
    tet2.run()
    inv = np.load(tmp_path / "w2" / "inv.invariants_0.npy")
    _, _, _, det = oracle.new_infer_resolved_quartets(arrays["tmparr"], arrays["tmpmap"], q.astype(np.uint32), True,
                                                      return_details=True)
    assert inv.dtype == np.uint32 and (inv == det["mats"][:, 0]).all()
    # CLI: init + run (full mode: the live CLI's default is subsample_snps=False)
    assert cli.main(["init", str(tmp_path / "in.npz"), "-n", "c", "-w", str(tmp_path / "c"), "-r", "3"]) == 0
    assert cli.main(["run", str(tmp_path / "c" / "c.json"), "-c", "4", "-b", "0"]) == 0
    q, scores, best, nsnps = run_inference.read_quartets_tsv(tmp_path / "c" / "c.quartets_0.tsv")
    cproj = Project.load_json(tmp_path / "c" / "c.json")
    arrays = database.read_arrays(cproj.database_file, ["tmparr", "tmpmap"])
    _, ostat, _ = oracle.new_infer_resolved_quartets(arrays["tmparr"], arrays["tmpmap"], q.astype(np.uint32), False)
    assert cproj.subsample_snps is False and (nsnps == ostat[:, 1]).all() and (best == ostat[:, 0]).all()
    # engine entry point with the reference's signature
    rq, rstat, rscor = tb.infer_resolved_quartets(cproj.database_file, 8, [(0, 1, 2, 3), (1, 3, 5, 7)], False)
    assert rstat.shape == (2, 2) and rscor.shape == (2, 3) and rq.shape == (2, 4)


def _matrix_from_counts(T):
    """4 x S code matrix (one column per counted SNP, every column its own locus) realising the
    count tensor T[a,b,c,d] exactly."""
    pats = np.repeat(np.arange(256), T.reshape(-1))
    arr = np.stack([(pats >> 6) & 3, (pats >> 4) & 3, (pats >> 2) & 3, pats & 3]).astype(np.uint8)
    tmpmap = np.stack([np.arange(arr.shape[1]), np.arange(arr.shape[1])], 1).astype(np.uint32)
    return arr, tmpmap


def test_crafted_count_tensors_exercise_the_svd(tb, oracle):
    """Singular-value epilogue on count tensors chosen for their spectra: rank 1..5, repeated
    singular values, seven orders of magnitude of dynamic range, dense random, near-duplicates."""
    rng = np.random.default_rng(12)
    tensors = []
    for rank in (1, 2, 3, 5):                                     # few non-zero patterns
        T = np.zeros(256, np.int64)
        T[rng.choice(256, rank, replace=False)] = rng.integers(1, 400, rank)
        tensors.append(T)
    T = np.zeros((4, 4, 4, 4), np.int64); T[0, 1, 0, 1] = T[1, 0, 1, 0] = T[2, 3, 2, 3] = T[3, 2, 3, 2] = 100
    tensors.append(T.reshape(-1))                                 # four equal singular values
    T = rng.integers(0, 3, 256); T[17] = 60000; T[200] = 9000
    tensors.append(T)                                             # large dynamic range
    tensors.append(rng.integers(0, 60, 256))                      # dense, full rank
    T = np.zeros((16, 16), np.int64); T[:, :] = rng.integers(1, 30, 16)[None, :]; T[5, 5] += 1
    tensors.append(T.reshape(-1))                                 # rank 1 plus one perturbed entry
    tensors.append(np.outer(rng.integers(1, 9, 16), rng.integers(1, 9, 16)).reshape(-1))   # exact rank 1, dense
    quartet = np.array([[0, 1, 2, 3]], dtype=np.uint32)
    for T in tensors:
        T = np.asarray(T, np.int64).copy()
        T[[0, 85, 170, 255]] = 0                                  # invariant patterns are never counted
        arr, tmpmap = _matrix_from_counts(T)
        for sub in (True, False):
            _, rstat, rscor, extra = tb.new_infer_resolved_quartets(arr, tmpmap, quartet, sub, return_counts=True)
            _, ostat, oscor, det = oracle.new_infer_resolved_quartets(arr, tmpmap, quartet, sub, return_details=True)
            assert (extra["counts"].reshape(-1) == T).all()
            assert extra["minrank"][0] == det["minrank"][0], (T.nonzero()[0][:8], extra["minrank"], det["ranks"])
            assert_scores_close(rscor, oscor, ostat[:, 1])
            assert_best_matches(rstat[:, 0], oscor, ostat)


def test_device_unranking_matches_itertools_and_reference_vectors(tb):
    from itertools import combinations
    from math import comb
    from tetrad_b200 import combinations as cb
    from tetrad_b200._lib import TetradB200Error
    for n in (4, 5, 12, 37):
        want = np.array(list(combinations(range(n), 4)), dtype=np.int32)
        assert (cb.device_quartets(n).cpu().numpy() == want).all()
    g = load_golden("unrank")
    assert (cb.device_quartets(12, ranks=g["idx12"]).cpu().numpy().view(np.uint32) == g["comb12"]).all()
    assert (cb.device_quartets(200, ranks=g["idx200"]).cpu().numpy().view(np.uint32) == g["comb200"]).all()
    rng = np.random.default_rng(3)
    ranks = rng.choice(comb(500, 4), size=20000, replace=False)          # C5-sized tip count
    assert (cb.device_quartets(500, ranks=ranks).cpu().numpy().view(np.uint32) == cb.unrank_quartets(ranks, 500)).all()
    part = cb.device_quartets(50, start=1000, count=77).cpu().numpy()
    assert (part == np.array(list(combinations(range(50), 4)), dtype=np.int32)[1000:1077]).all()
    with pytest.raises(TetradB200Error):
        cb.device_quartets(10, start=200, count=100)
    with pytest.raises(ValueError):
        cb.device_quartets(10, ranks=np.array([comb(10, 4)]))


@pytest.mark.parametrize("sub", [True, False])
def test_more_than_63_tiles_exercises_the_counter_flush(tb, oracle, kernel, sub):
    """> 64 512 loci = more than 63 tiles of 1024: the packed 16-bit counters of the POPC kernels are flushed
    inside the tile loop (as at the C3-C5 shapes); the tensor-core kernel's scaled s32 accumulators are flushed
    every 16 384 tile products, lowered here to 5 so that the path runs.  Exact counts against the oracle."""
    from tetrad_b200 import _lib, synth
    _lib.check(_lib.load().tq_set_option(b"mma_flush_every", 5))
    d = synth.generate(6, 70_000, 150_000, seed=17)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=17)
    quartets = np.array([[0, 1, 2, 3], [0, 2, 4, 5], [1, 3, 4, 5], [0, 1, 4, 5], [2, 3, 4, 5]], dtype=np.uint32)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    assert eng.info["ntiles"] > 63
    rstat, rscor, counts, minrank = eng.resolve(quartets, sub, want_counts=True)
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, sub, return_details=True)
    assert (counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16) == det["mats"][:, 0]).all()
    assert (rstat.cpu().numpy().view(np.uint32) == ostat).all()
    assert_scores_close(rscor.cpu().numpy(), oscor, ostat[:, 1])
    assert (minrank.cpu().numpy() == det["minrank"]).all()


@pytest.mark.parametrize("count_warps", [8, 12])
def test_stacked_tiles_count_every_locus_once(tb, oracle, count_warps):
    """Subsample mode stacks the merged planes of consecutive sparse tiles before one tile product; loci whose bit slot
    is taken are counted by shared-memory atomics.  Whatever the stacking limits (never / only without collisions / the
    defaults / everything, i.e. hundreds of loci through the atomics) and whatever the count / score warp mix, the counts
    are the oracle's and all results are the same bits; the diagnostics in the status block show the paths were taken."""
    from tetrad_b200 import _lib, synth
    lib = _lib.load()
    d = synth.generate(9, 30_000, 90_000, seed=23)                  # ~30 tiles, a third of the loci selected
    db = synth.prepare_database(d.snps, d.snpsmap, seed=23)
    from itertools import combinations
    quartets = np.array(list(combinations(range(9), 4)), dtype=np.uint32)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets[::6], True, return_details=True)
    outs, stats = [], []
    try:
        _lib.set_kernel("mma")
        _lib.check(lib.tq_set_option(b"count_warps", count_warps))
        _lib.check(lib.tq_set_option(b"stack_stats", 1))
        for fill, ovf in ((0, 0), (85, 0), (-1, 8), (100, 32), (40, 3)):
            _lib.check(lib.tq_set_option(b"stack_fill", fill))
            _lib.check(lib.tq_set_option(b"stack_ovf", ovf))
            rstat, rscor, counts, minrank = eng.resolve(quartets, True, want_counts=True)
            outs.append((rstat.cpu().numpy().copy(), rscor.cpu().numpy().copy(), counts.cpu().numpy().copy(), minrank.cpu().numpy().copy()))
            stats.append(eng._last_status.cpu().numpy().astype(np.int64) & 0xFFFFFFFF)
            assert lib.tq_last_count_warps() == count_warps
    finally:
        _lib.set_kernel("auto")
        for name, val in ((b"count_warps", 0), (b"stack_fill", -1), (b"stack_ovf", 8), (b"stack_stats", 0)):
            _lib.check(lib.tq_set_option(name, val))
    got = outs[0][2].view(np.uint32).reshape(-1, 16, 16)
    assert (got[::6] == det["mats"][:, 0]).all() and (outs[0][0].view(np.uint32)[::6] == ostat).all()
    assert_scores_close(outs[0][1][::6], oscor, ostat[:, 1])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert a.tobytes() == b.tobytes()
    ntiles, m = eng.info["ntiles"], len(quartets)
    assert stats[0][1] == ntiles * m and stats[0][2] == 0 and stats[0][3] == 0        # never stacked: one product per tile
    assert stats[1][3] == 0 and stats[1][1] + stats[1][2] == ntiles * m               # stacked only without collisions
    assert stats[3][1] < stats[2][1] < stats[0][1]                                     # more stacking, fewer products
    assert stats[3][3] > stats[2][3] > 0                                               # ... and more loci through the atomics
    with pytest.raises(_lib.TetradB200Error):
        _lib.check(lib.tq_set_option(b"count_warps", 10))
    with pytest.raises(_lib.TetradB200Error):
        _lib.check(lib.tq_set_option(b"stack_ovf", 33))


def test_c3_shape_random_quartets_sampled_parity(tb, oracle):
    """BASELINE config C3 shape (100 tips, 100 000 loci, 500 000 SNPs): a sample of random quartets,
    unranked on the device from the reference's rng.choice draw, exact against the oracle."""
    from math import comb
    from tetrad_b200 import combinations as cb, synth
    db = synth.config_arrays("C3")
    ranks = np.random.default_rng(42).choice(comb(100, 4), size=1_000_000, replace=False)[:3000]
    q_dev = cb.device_quartets(100, ranks=ranks)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    rstat, rscor, counts, minrank = eng.resolve(q_dev, True, want_counts=True)
    quartets = q_dev.cpu().numpy().view(np.uint32)
    assert (quartets == cb.unrank_quartets(ranks, 100)).all()
    pick = np.arange(0, 3000, 25)
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets[pick], True, return_details=True)
    assert (counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16)[pick] == det["mats"][:, 0]).all()
    assert (rstat.cpu().numpy().view(np.uint32)[pick] == ostat).all()
    assert_scores_close(rscor.cpu().numpy()[pick], oscor, ostat[:, 1])
    assert (minrank.cpu().numpy()[pick] == det["minrank"]).all()


@pytest.mark.parametrize("sub", [True, False])
def test_kernel_variants_agree_bit_for_bit(tb, sub):
    """The three resolve kernels differ in how they count (POPC vs tensor cores) and in how many flattenings share a
    Jacobi pass, but every flattening goes through the same arithmetic and stops after its own last sweep: counts,
    nsnps, best split and the f64 scores are identical bits, also between two runs of the same kernel."""
    from tetrad_b200 import _lib, synth
    d = synth.generate(14, 4000, 15000, seed=5)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=5)
    from itertools import combinations
    quartets = np.array(list(combinations(range(14), 4)), dtype=np.uint32)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    outs = []
    try:
        for name in ("fused", "mma", "mma") + (() if sub else ("pairgemm",)):
            _lib.set_kernel(name)
            rstat, rscor, counts, minrank = eng.resolve(quartets, sub, want_counts=True)
            outs.append((rstat.cpu().numpy().copy(), rscor.cpu().numpy().copy(), counts.cpu().numpy().copy(),
                         minrank.cpu().numpy().copy()))
    finally:
        _lib.set_kernel("auto")
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert a.tobytes() == b.tobytes()


def test_results_do_not_depend_on_order_or_batching(tb):
    """C2 shape, 20 000 quartets through the tensor-core kernel: a quartet's result is the same bits whatever its
    position in the list and whatever the launch it is part of (the score warps batch four quartets per Jacobi
    pass, picked by scheduling; every flattening stops after its own last sweep, so the batch does not matter)."""
    from tetrad_b200 import synth
    db = synth.config_arrays("C2")
    rng = np.random.default_rng(3)
    quartets = np.sort(np.stack([rng.choice(50, 4, replace=False) for _ in range(20000)]), axis=1).astype(np.uint32)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    rstat, rscor = eng.resolve(quartets, True)
    rstat, rscor = rstat.cpu().numpy(), rscor.cpu().numpy()
    perm = rng.permutation(len(quartets))
    pstat, pscor = eng.resolve(np.ascontiguousarray(quartets[perm]), True)
    assert (pstat.cpu().numpy() == rstat[perm]).all() and (pscor.cpu().numpy() == rscor[perm]).all()
    for lo, hi in ((0, 7001), (7001, 20000)):                     # two launches of other sizes
        hstat, hscor = eng.resolve(np.ascontiguousarray(quartets[lo:hi]), True)
        assert (hstat.cpu().numpy() == rstat[lo:hi]).all() and (hscor.cpu().numpy() == rscor[lo:hi]).all()


def test_pairgemm_full_mode_small_and_ragged(tb, oracle):
    """The tip-pair GEMM (tcgen05.mma kind::i8) forced on small shapes: fewer pairs than one tile holds, a tile grid with
    ragged edges, an out-of-range quartet, a duplicate quartet and an empty chunk of the list -- exact counts, nsnps,
    minrank and the same score bits as the per-quartet kernel."""
    from itertools import combinations
    from tetrad_b200 import _lib, synth
    for n, nloci, nsnps in ((5, 40, 160), (9, 700, 2600), (21, 1500, 5000)):
        d = synth.generate(n, nloci, nsnps, seed=200 + n)
        db = synth.prepare_database(d.snps, d.snpsmap, seed=n)
        quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
        if len(quartets) > 1500:
            quartets = quartets[np.sort(np.random.default_rng(2).choice(len(quartets), 1500, replace=False))]
        quartets = np.vstack([quartets, quartets[:1]])                 # a duplicate
        eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
        try:
            _lib.set_kernel("pairgemm")
            rstat, rscor, counts, minrank = eng.resolve(quartets, False, want_counts=True)
            assert "pair_gemm_kernel" in last_kernel()
            eng.check_status()
            _lib.set_kernel("mma")
            mstat, mscor, mcounts, mminrank = eng.resolve(quartets, False, want_counts=True)
        finally:
            _lib.set_kernel("auto")
        _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, False, return_details=True)
        got = counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16)
        assert (got == det["mats"][:, 0]).all(), "tip-pair GEMM counts differ from the oracle"
        assert (rstat.cpu().numpy().view(np.uint32) == ostat).all()
        assert (counts == mcounts).all() and (rstat == mstat).all() and (minrank == mminrank).all()
        assert rscor.cpu().numpy().tobytes() == mscor.cpu().numpy().tobytes()
        assert_scores_close(rscor.cpu().numpy(), oscor, ostat[:, 1], counts=got)
        # a quartet with a tip out of range is flagged, the others are unaffected
        nb = min(40, len(quartets))
        bad = quartets[:nb].copy()
        bad[nb // 2, 3] = n + 3
        try:
            _lib.set_kernel("pairgemm")
            bstat, bscor = eng.resolve(bad, False)
            with pytest.raises(IndexError):
                eng.check_status()
        finally:
            _lib.set_kernel("auto")
        keep = np.arange(nb) != nb // 2
        assert (bstat.cpu().numpy().view(np.uint32)[keep] == ostat[:nb][keep]).all()
        assert bstat.cpu().numpy().view(np.uint32)[nb // 2, 0] == 0xFFFFFFFF
        eng.close()


def test_full_mode_is_routed_on_the_device(tb, oracle):
    """Full mode, default kernel rule: an exhaustive list (and a contiguous eighth of it, one rank's share of an 8-GPU
    run) goes to the tip-pair GEMM, a sparse random sample and a list with a quartet whose tips are not ascending go to
    the per-quartet kernel; the decision is taken by the planner on the device and every route gives the same bits."""
    import torch
    from itertools import combinations
    from tetrad_b200 import _lib, synth
    n = 40
    d = synth.generate(n, 800, 3000, seed=77)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=7)
    allq = np.array(list(combinations(range(n), 4)), dtype=np.uint32)           # 91 390 quartets
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])

    def routed(q):
        torch.cuda.synchronize()
        before = _lib.pairgemm_stats()
        out = eng.resolve(q, False, want_counts=True)
        torch.cuda.synchronize()
        after = _lib.pairgemm_stats()
        return (after[0] - before[0], after[1] - before[1]), out

    route, full = routed(allq)
    assert route == (1, 0), route
    share = allq[len(allq) // 8 * 3: len(allq) // 8 * 4]
    route, part = routed(share)
    assert route == (1, 0), route
    lo = len(allq) // 8 * 3
    for a, b in zip(part, full):
        assert torch.equal(a, b[lo: lo + len(share)])
    pick = np.sort(np.random.default_rng(3).choice(len(allq), 4000, replace=False))
    route, sparse = routed(allq[pick])
    assert route == (0, 1), route
    idx = torch.from_numpy(pick).to(full[0].device)
    for a, b in zip(sparse, full):
        assert torch.equal(a, b[idx])
    mixed = allq.copy()
    mixed[5] = mixed[5][[1, 0, 2, 3]]                                            # valid tips, not ascending
    route, res = routed(mixed)
    assert route == (0, 1), route
    eng.check_status()
    keep = np.arange(len(allq)) != 5
    assert torch.equal(res[0][torch.from_numpy(keep).to(res[0].device)], full[0][torch.from_numpy(keep).to(res[0].device)])
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], allq[::97], False, return_details=True)
    assert (full[2].cpu().numpy().view(np.uint32).reshape(-1, 16, 16)[::97] == det["mats"][:, 0]).all()
    eng.close()
