"""CPU: the C-ABI library loads, exports every symbol the header declares, and its host logic
(locus plan) is right.  No kernel runs here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, load_golden
from tetrad_b200 import _lib


def _plan(lib, locus, stride=1):
    locus = np.ascontiguousarray(locus, dtype=np.uint32)
    h = ctypes.c_void_p()
    rc = lib.tq_plan_create(locus.ctypes.data, locus.shape[0], stride, -1, None, ctypes.byref(h))
    return rc, h


def _tables(lib, h):
    info = (ctypes.c_int64 * 8)()
    assert lib.tq_plan_info(h, info) == 0
    ntiles = info[2]
    off = np.zeros(ntiles + 1, np.uint32); k = np.zeros(ntiles, np.uint32)
    start = np.zeros(ntiles * 1024, np.uint32); length = np.zeros(ntiles * 1024, np.uint32)
    run = np.zeros(ntiles * 1024, np.uint32)
    assert lib.tq_plan_tables(h, off.ctypes.data, k.ctypes.data, start.ctypes.data, length.ctypes.data, run.ctypes.data) == 0
    _tables.last_run = run
    return list(info), off, k, start, length


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "tetrad_b200.h").read_text()
    declared = set(re.findall(r"\b(tq_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.tq_version() == 100


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under tetrad_b200/ may import, load or execute it."""
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|quartet_oracle|importlib[^\n]*oracle", re.M)
    for path in (ROOT / "tetrad_b200").rglob("*.py"):
        assert not bad.search(path.read_text()), path
    for path in (ROOT / "tetrad_b200" / "csrc").glob("*"):
        assert not re.search(r'#include\s+[<"][^>"]*oracle|dlopen[^\n]*oracle', path.read_text()), path


def test_plan_layout(lib):
    # loci of lengths 3,1,2,5 -> sorted (stable, longest first): 5,3,2,1
    locus = np.repeat([7, 8, 9, 10], [3, 1, 2, 5])
    rc, h = _plan(lib, locus)
    assert rc == 0
    info, off, k, start, length = _tables(lib, h)
    assert info[0] == 11 and info[1] == 4 and info[2] == 1 and info[4] == 5
    assert k.tolist() == [5] and off.tolist() == [0, 5 * 32]
    assert info[3] == 5 * 32 * 4 and info[5] == 5 * 1024
    assert length[:5].tolist() == [5, 3, 2, 1, 0]
    assert start[:4].tolist() == [6, 0, 4, 3]
    assert _tables.last_run[:4].tolist() == [3, 0, 2, 1]
    lib.tq_plan_destroy(h)


def test_plan_strided_and_unsorted_ids(lib):
    tmpmap = np.zeros((6, 2), np.uint32)
    tmpmap[:, 0] = [5, 5, 2, 2, 2, 9]          # ids need not increase, only be contiguous
    rc, h = _plan(lib, tmpmap, stride=2)
    assert rc == 0
    info, off, k, start, length = _tables(lib, h)
    assert info[1] == 3 and length[:3].tolist() == [3, 2, 1] and start[:3].tolist() == [2, 0, 5]
    lib.tq_plan_destroy(h)


def test_plan_from_run_lengths_equals_plan_from_ids(lib):
    lens = np.array([3, 1, 2, 5, 2, 2], np.uint32)
    h = ctypes.c_void_p()
    assert lib.tq_plan_create_runs(lens.ctypes.data, len(lens), -1, None, ctypes.byref(h)) == 0
    a = _tables(lib, h); run_a = _tables.last_run.copy()
    rc, h2 = _plan(lib, np.repeat(np.arange(6), lens))
    b = _tables(lib, h2); run_b = _tables.last_run.copy()
    assert a[0] == b[0] and all((x == y).all() for x, y in zip(a[1:], b[1:])) and (run_a == run_b).all()
    lib.tq_plan_destroy(h); lib.tq_plan_destroy(h2)
    bad = np.array([3, 0, 2], np.uint32)
    assert lib.tq_plan_create_runs(bad.ctypes.data, 3, -1, None, ctypes.byref(h)) == _lib.TQ_EINVAL


def test_plan_rejects_split_locus(lib):
    rc, h = _plan(lib, np.array([1, 1, 2, 1], np.uint32))
    assert rc == _lib.TQ_ELOCUS
    assert b"more than one run" in lib.tq_last_error()


def test_plan_empty_and_many_tiles(lib):
    rc, h = _plan(lib, np.zeros(0, np.uint32))
    assert rc == 0
    info = (ctypes.c_int64 * 8)()
    lib.tq_plan_info(h, info)
    assert info[1] == 0 and info[2] == 0 and info[3] == 0
    lib.tq_plan_destroy(h)
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 9, size=3000)
    locus = np.repeat(np.arange(3000), lens)
    rc, h = _plan(lib, locus)
    info, off, k, start, length = _tables(lib, h)
    assert info[2] == 3 and (np.diff(length[:3000].astype(int)) <= 0).all()
    assert k.tolist() == [length[0], length[1024], length[2048]]
    assert info[5] >= locus.size and off[-1] * 4 == info[3]
    # every column appears exactly once
    cols = np.concatenate([np.arange(s, s + l) for s, l in zip(start[:3000], length[:3000])])
    assert np.array_equal(np.sort(cols), np.arange(locus.size))
    lib.tq_plan_destroy(h)


def _emulate_rank_major(tmparr, quartet, start, length, subsample):
    """numpy model of what the kernel computes from the plan: for every sorted locus, the
    lowest rank whose column is complete and segregating (subsample) / all of them (full)."""
    rows = tmparr[quartet].astype(np.int64)
    counts = np.zeros(256, np.uint32)
    for s, l in zip(start, length):
        for kk in range(l):
            col = rows[:, s + kk]
            if (col > 3).any() or (col == col[0]).all():
                continue
            counts[((col[0] * 4 + col[1]) * 4 + col[2]) * 4 + col[3]] += 1
            if subsample:
                break
    return counts


@pytest.mark.parametrize("case", ["tiny8", "boot12", "micro4"])
def test_rank_major_selection_equals_reference(lib, case):
    """The layout trick (first usable rank per sorted locus) reproduces the reference's
    last_loc state machine exactly -- checked against golden counts from the reference."""
    g = load_golden(case)
    rc, h = _plan(lib, g["tmpmap"], stride=2)
    assert rc == 0
    info, off, k, start, length = _tables(lib, h)
    n = info[1]
    for qi in range(0, len(g["quartets"]), 7):
        for mode, sub in (("sub", True), ("full", False)):
            got = _emulate_rank_major(g["tmparr"], g["quartets"][qi], start[:n], length[:n], sub)
            assert (got == g[f"{mode}_mats0"][qi].ravel()).all()
    lib.tq_plan_destroy(h)


def test_device_calls_fail_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.tq_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and lib.tq_last_error()
    from tetrad_b200 import QuartetResolver
    with pytest.raises(RuntimeError):
        QuartetResolver(device=0)
    rc, hp = _plan(lib, np.array([0, 0, 1], np.uint32))
    assert lib.tq_pack(hp, None, 4, 3, None, None) == _lib.TQ_EINVAL     # host-only plan
    lib.tq_plan_destroy(hp)


@pytest.mark.parametrize("shuffle_ids", [False, True])
def test_plan_tables_equal_a_numpy_restatement(lib, shuffle_ids):
    """Every table of a plan against a numpy restatement on random run lengths (ids increasing, or any distinct ids):
    stable sort by length (longest first), tiles of 1024 loci, K = the tile's longest locus, zero padding."""
    rng = np.random.default_rng(11 + shuffle_ids)
    nl = 2500
    lens = rng.integers(1, 12, size=nl).astype(np.int64)
    ids = rng.permutation(10 * nl)[:nl] if shuffle_ids else np.sort(rng.permutation(10 * nl)[:nl])
    tmpmap = np.zeros((int(lens.sum()), 3), np.uint32)
    tmpmap[:, 0] = np.repeat(ids, lens)
    rc, h = _plan(lib, tmpmap, stride=3)
    assert rc == 0
    info, off, k, start, length = _tables(lib, h)
    run = _tables.last_run
    order = np.argsort(-lens, kind="stable")
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ntiles = (nl + 1023) // 1024
    assert info[0] == lens.sum() and info[1] == nl and info[2] == ntiles and info[4] == lens.max()
    assert (length[:nl] == lens[order]).all() and (start[:nl] == starts[order]).all() and (run[:nl] == order).all()
    assert not length[nl:].any() and not start[nl:].any() and not run[nl:].any()                 # padding is zero
    want_k = np.array([lens[order][t * 1024] for t in range(ntiles)])
    assert (k == want_k).all() and (off == np.concatenate([[0], np.cumsum(want_k * 32)])).all()
    assert info[3] == off[-1] * 4 and info[5] == (want_k * 1024).sum()
    lib.tq_plan_destroy(h)

