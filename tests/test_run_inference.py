"""CPU: the caller side of the hot path -- result sink, wQMC formatting, resampling, project
files, replicate loop -- against golden text produced by the executed reference."""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import load_golden
from tetrad_b200 import combinations, database, run_inference, synth, write_database
from tetrad_b200.schema import Project, RNGStateModel


def test_tsv_matches_reference_pandas_output(tmp_path):
    g, sink = load_golden("noisy12"), load_golden("sink")
    f = tmp_path / "q.tsv"
    run_inference.write_quartets_tsv(f, g["quartets"], g["sub_rstat"], g["sub_rscor"])
    assert f.read_text() == str(sink["tsv"])


@pytest.mark.parametrize("key,kwargs", [("w0", dict(weights=0)), ("w1", dict(weights=1)), ("w2", dict(weights=2)),
                                        ("w3", dict(weights=3)), ("w1_minsnps400", dict(weights=1, min_snps=400)),
                                        ("w2_minratio5", dict(weights=2, min_ratio=5.0))])
def test_qmc_format_matches_reference(tmp_path, key, kwargs):
    sink = load_golden("sink")
    f = tmp_path / "q.tsv"
    f.write_text(str(sink["tsv"]))
    assert "\n".join(run_inference.iter_qmc_formatted(f, **kwargs)) == str(sink[key])


def test_write_qmc_format_is_a_seeded_permutation(tmp_path):
    sink = load_golden("sink")
    f = tmp_path / "q.tsv"
    f.write_text(str(sink["tsv"]))
    n = run_inference.write_qmc_format(f, tmp_path / "a.txt", 1, rng=3)
    run_inference.write_qmc_format(f, tmp_path / "b.txt", 1, rng=3)
    a = (tmp_path / "a.txt").read_text()
    assert a == (tmp_path / "b.txt").read_text() and n == 495
    assert sorted(a.splitlines()) == sorted(str(sink["w1"]).splitlines())
    assert a.splitlines() != str(sink["w1"]).splitlines()


def test_in_memory_sink_equals_file_based_sink(tmp_path):
    """The C sink fed from memory (what the replicate loop does) writes the same wQMC input as the
    reference-style path that re-reads the TSV, for every weighting and filter."""
    g = load_golden("noisy12")
    f = tmp_path / "q.tsv"
    run_inference.write_quartets_tsv(f, g["quartets"], g["sub_rstat"], g["sub_rscor"])
    res = (g["quartets"], g["sub_rstat"], g["sub_rscor"])
    for kw in (dict(weights=0), dict(weights=1), dict(weights=2), dict(weights=3), dict(weights=1, min_snps=400),
               dict(weights=2, min_ratio=5.0)):
        tips, weight = run_inference.qmc_weights(*res, **kw)
        text = run_inference.format_qmc_lines(tips, weight).decode()
        assert text == "".join(line + "\n" for line in run_inference.iter_qmc_formatted(f, **kw))
        run_inference.write_qmc_format(f, tmp_path / "a.txt", rng=5, **kw)
        run_inference.write_qmc_format(f, tmp_path / "b.txt", rng=5, results=res, **kw)
        assert (tmp_path / "a.txt").read_bytes() == (tmp_path / "b.txt").read_bytes()


def test_c_formatter_rounds_like_printf(tmp_path):
    """%.6f / %.5f ties, carries, tiny, huge and non-finite values: byte-identical to Python's
    (= the C library's) formatting, and the read-back value equals float(text)."""
    import ctypes
    from tetrad_b200 import _lib
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.random(20000) * 10.0 ** rng.integers(-8, 7, 20000), (rng.integers(0, 10**9, 20000) + 0.5) / 1e6,
                           np.arange(2000) / 64.0 + 0.0000005, [0.0, 0.001, 0.9999995, 0.9999994999, 1e-7, 5e-7, 123456789.123456789,
                                                               2.5e-6, 3.9e9, 4.1e9, 1e15]])
    vals = vals[: len(vals) // 3 * 3].reshape(-1, 3)
    m = len(vals)
    q = np.zeros((m, 4), np.uint32); st = np.zeros((m, 2), np.uint32)
    f = tmp_path / "v.tsv"
    run_inference.write_quartets_tsv(f, q, st, vals)
    got = [line.split("\t")[4:7] for line in f.read_text().splitlines()]
    want = [["%.6f" % x for x in row] for row in vals]
    assert got == want
    back = np.empty_like(vals)
    assert _lib.load().tq_round_scores(vals.ctypes.data, vals.size, back.ctypes.data) == 0
    assert (back == np.array([[float(x) for x in row] for row in want])).all()
    w = np.ascontiguousarray(vals[:, 0])
    lines = run_inference.format_qmc_lines(np.zeros((m, 4), np.uint32), w).decode().splitlines()
    assert lines == ["0,0|0,0:%.5f" % x for x in w]


def test_relabel_tree():
    names = {0: "a", 1: "b_1", 2: "c", 3: "d", 10: "k"}
    assert run_inference.relabel_tree("((0,1),(2,(3,10)));\n", names) == "((a,b_1),(c,(d,k)));"
    assert run_inference.relabel_tree("((0:1,1:1):2,(2,3));", names) == "((a:1,b_1:1):2,(c,d));"


def test_resample_arrays_semantics_and_rng_stream():
    d = synth.generate(6, 40, 150, seed=3, p_ambig=0.0)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=1)
    rng = np.random.default_rng(11)
    tmparr, tmpmap = run_inference.resample_arrays(db["seqarr"], db["spans"], rng)
    # the main stream advanced by exactly the reference's three draws (run_inference.py:117-123)
    ref = np.random.default_rng(11)
    lidxs = ref.choice(40, 40, replace=True); ref.integers(2**31); ref.integers(2**31)
    assert rng.integers(1 << 30) == ref.integers(1 << 30)
    lens = db["spans"][lidxs, 1] - db["spans"][lidxs, 0]
    assert tmparr.shape[1] == lens.sum() and (np.bincount(tmpmap[:, 0], minlength=40) == lens).all()
    assert (tmpmap[:, 1] == np.arange(tmparr.shape[1])).all()
    assert set(np.unique(tmparr)) <= {0, 1, 2, 3, 78}
    # every instance is a column permutation of its source locus (no ambiguity codes in this data set)
    recoded = synth.recode_acgt(db["seqarr"].copy())
    off, moved = 0, 0
    for l in lidxs:
        a, b = db["spans"][l]
        src, got = recoded[:, a:b], tmparr[:, off: off + (b - a)]
        off += b - a
        assert sorted(map(bytes, src.T)) == sorted(map(bytes, got.T))
        moved += int((src != got).any())
    assert moved > 0                      # and the order really is shuffled


def test_project_json_roundtrip_and_names(tmp_path):
    proj = Project(version="0.1.0", data=tmp_path / "in.npz", name="T", workdir=tmp_path / "w", subsample_snps=True,
                   weights=1, nquartets=0, random_seed=5)
    assert proj.qrts_file.name == "T.quartets_0.tsv" and proj.database_file.name == "T.database.hdf5"
    assert proj.best_file.name == "T.best_tree.nwk" and proj.boots_file.name == "T.boots.nwk"
    assert proj.qmc_in_file.name == "T.qmc_in.txt" and proj.json_file.name == "T.json"
    rng = np.random.default_rng(1); rng.random(3)
    proj.bootstrap_rng = RNGStateModel.from_rng(rng)
    proj.samples = {0: "a", 1: "b"}
    proj.save_json()
    back = Project.load_json(proj.json_file)
    assert back.samples == {0: "a", 1: "b"} and back.bootstrap_rng.to_rng().random() == rng.random()
    fields = json.loads(proj.json_file.read_text())
    for key in ("version", "data", "name", "workdir", "subsample_snps", "weights", "nquartets", "random_seed",
                "bootstrap_idx", "bootstrap_rng", "nsnps", "nsamples", "nqrts", "nqrts_total", "samples"):
        assert key in fields


def test_get_nquartets_follows_reference_rules():
    assert write_database.get_nquartets(12, 0) == (495, 495)
    assert write_database.get_nquartets(12, 10**6) == (495, 495)
    assert write_database.get_nquartets(50, 100) == (100, 230300)          # low: warn, keep
    assert write_database.get_nquartets(50, 100000) == (100000, 230300)


class _OracleEngine:
    """stands in for QuartetResolver on the CPU box (tests only)"""
    def load(self, tmparr, tmpmap):
        self.tmparr, self.tmpmap = tmparr, tmpmap
    def resolve(self, quartets, subsample):
        import torch
        from oracle import oracle
        _, rstat, rscor = oracle.new_infer_resolved_quartets(self.tmparr, self.tmpmap, quartets.numpy().view(np.uint32), subsample)
        return torch.from_numpy(rstat.view(np.int32)), torch.from_numpy(rscor)


def test_replicate_loop_checkpoints_and_resumes(tmp_path, oracle):
    d = synth.generate(7, 60, 240, seed=8)
    database.write_snps_input(tmp_path / "in.npz", d.snps, d.snpsmap, d.names)
    proj = Project(version="0.1.0", data=tmp_path / "in.npz", name="R", workdir=tmp_path, subsample_snps=True,
                   weights=1, nquartets=0, random_seed=42)
    write_database.write_database(proj)
    assert (proj.nsamples, proj.nsnps, proj.nqrts, proj.nqrts_total) == (7, 240, 35, 35)
    proj.save_json()
    run_inference.run_inference(proj, nboots=1, resolver=_OracleEngine())
    assert proj.bootstrap_idx == 2
    t0 = (tmp_path / "R.quartets_0.tsv").read_text().splitlines()
    t1 = (tmp_path / "R.quartets_1.tsv").read_text().splitlines()
    assert len(t0) == 35 and len(t1) == 35 and t0 != t1 and len(t0[0].split("\t")) == 9
    assert (tmp_path / "R.qmc_in.txt").exists()
    # resume from the JSON checkpoint: replicate 2 must equal a straight 0..2 run
    again = Project.load_json(proj.json_file)
    run_inference.run_inference(again, nboots=2, resolver=_OracleEngine())
    fresh = Project(version="0.1.0", data=tmp_path / "in.npz", name="S", workdir=tmp_path, subsample_snps=True,
                    weights=1, nquartets=0, random_seed=42)
    write_database.write_database(fresh)
    run_inference.run_inference(fresh, nboots=2, resolver=_OracleEngine())
    assert (tmp_path / "R.quartets_2.tsv").read_text() == (tmp_path / "S.quartets_2.tsv").read_text()


@pytest.mark.parametrize("weights", [0, 1, 2, 3])
def test_qmc_weights_c_pass_equals_numpy_formulation(weights):
    """tq_qmc_weights (one C pass: scores through the TSV text, filters, tips in split order, weights) against the
    numpy formulation of the same reference lines (run_inference.py:258-300), which a NaN score forces."""
    from tetrad_b200 import run_inference as ri
    rng = np.random.default_rng(weights)
    m = 20000
    q = np.sort(rng.integers(0, 60, (m, 4)), axis=1).astype(np.uint32)
    rstat = np.stack([rng.integers(0, 3, m), rng.integers(0, 3000, m)], axis=1).astype(np.uint32)
    rscor = rng.random((m, 3)) * rng.choice([1e-3, 1.0, 300.0], size=(m, 1))
    rscor[::97, 0] = 0.0                                   # smallest score exactly zero: ratio defined as 1
    rscor[::89] = rscor[::89, :1]                          # three equal scores
    kw = dict(min_snps=100, min_ratio=1.1 if weights else 1.0)
    tips, wt = ri.qmc_weights(q, rstat, rscor, weights, **kw)
    q2 = np.vstack([q, [[0, 1, 2, 3]]]).astype(np.uint32)
    r2 = np.vstack([rstat, [[0, 0]]]).astype(np.uint32)    # nsnps 0: filtered out again
    s2 = np.vstack([rscor, [[np.nan, 1.0, 1.0]]])
    tips2, wt2 = ri.qmc_weights(q2, r2, s2, weights, **kw)
    assert 0 < len(tips) < m and len(tips2) == len(tips)
    assert (tips == tips2).all() and wt.tobytes() == wt2.tobytes()


def test_threaded_sink_equals_one_call(tmp_path):
    """The sink formats ranges of 65 536 rows on a thread pool; the bytes must be those of one C call over all rows
    (and of Python's own %-formatting on a sample), also for the shuffled wQMC lines addressed by an index array."""
    import ctypes
    from tetrad_b200 import _lib, run_inference as ri
    lib = _lib.load()
    rng = np.random.default_rng(11)
    m = 150_001
    q = np.sort(rng.integers(0, 300, (m, 4)), axis=1).astype(np.uint32)
    rstat = np.stack([rng.integers(0, 3, m), rng.integers(0, 90000, m)], axis=1).astype(np.uint32)
    rscor = rng.random((m, 3)) * rng.choice([1e-4, 1.0, 5000.0], size=(m, 1))
    ri.write_quartets_tsv(tmp_path / "a.tsv", q, rstat, rscor)
    buf = ctypes.create_string_buffer(160 * m)
    n = lib.tq_format_tsv(q.ctypes.data, rstat.ctypes.data, rscor.ctypes.data, m, buf, len(buf))
    data = (tmp_path / "a.tsv").read_bytes()
    assert n > 0 and data == buf.raw[:n]
    lines = data.split(b"\n")
    for i in (0, 65535, 65536, 131071, 131072, m - 1):
        want = "%d\t%d\t%d\t%d\t%.6f\t%.6f\t%.6f\t%d\t%d" % (*q[i], *rscor[i], *rstat[i])
        assert lines[i].decode() == want
    tips, w = ri.qmc_weights(q, rstat, rscor, 1)
    perm = rng.permutation(len(tips))
    assert ri.format_qmc_lines(tips, w, rows=perm) == ri.format_qmc_lines(tips[perm], w[perm])
    one = ctypes.create_string_buffer(128 * len(tips))
    tp, wp = np.ascontiguousarray(tips[perm]), np.ascontiguousarray(w[perm])       # keep them alive across the call
    k = lib.tq_format_qmc(tp.ctypes.data, wp.ctypes.data, len(tips), one, len(one))
    assert ri.format_qmc_lines(tips, w, rows=perm) == one.raw[:k]
    with pytest.raises(IndexError):
        ri.format_qmc_lines(tips, w, rows=np.array([len(tips)]))


def test_sink_formats_non_scores_without_overrunning(tmp_path):
    """Values the kernel never produces (inf, NaN, negative, 1e300) are written as printf would write them, through a
    bounded temporary: the row budget is checked, nothing is written past it, and a short buffer is an error with a
    message (the advance used to be snprintf's would-be length)."""
    import ctypes
    from tetrad_b200 import _lib, run_inference as ri
    lib = _lib.load()
    q = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [1, 2, 3, 4]], dtype=np.uint32)
    rstat = np.array([[0, 10], [1, 20], [2, 30]], dtype=np.uint32)
    rscor = np.array([[1e300, 2.5, -1.0], [np.inf, np.nan, 3.0], [0.125, 0.25, 0.5]])
    ri.write_quartets_tsv(tmp_path / "odd.tsv", q, rstat, rscor)
    lines = (tmp_path / "odd.tsv").read_text().split("\n")
    assert lines[0] == "0\t1\t2\t3\t%.6f\t2.500000\t-1.000000\t0\t10" % 1e300
    assert lines[1] == "4\t5\t6\t7\tinf\tnan\t3.000000\t1\t20"
    assert lines[2] == "1\t2\t3\t4\t0.125000\t0.250000\t0.500000\t2\t30"
    small = ctypes.create_string_buffer(3 * 160)
    n = lib.tq_format_tsv(q.ctypes.data, rstat.ctypes.data, rscor.ctypes.data, 3, small, len(small))
    assert n == _lib.TQ_ENOMEM and b"buffer too small" in lib.tq_last_error()
    tips = np.ascontiguousarray(q)
    w = np.array([1e300, 1.5, np.inf])
    text = ri.format_qmc_lines(tips, w).decode().split("\n")
    assert text[0] == "0,1|2,3:%.5f" % 1e300 and text[1] == "4,5|6,7:1.50000" and text[2] == "1,2|3,4:inf"
