import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
GOLDEN_CASES = ["noisy12", "clean10", "tiny8", "missing7", "boot12", "micro4"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(GOLDEN / f"{name}.npz") as g:
        return {k: g[k] for k in g.files}


def mats_from_mats0(m0):
    """All three flattenings from mats[0] (reference resolve_quartets.py:67-72)."""
    t = m0.reshape(-1, 4, 4, 4, 4)
    return np.stack([m0.reshape(-1, 16, 16),
                     t.transpose(0, 1, 3, 2, 4).reshape(-1, 16, 16),
                     t.transpose(0, 1, 4, 2, 3).reshape(-1, 16, 16)], axis=1)


def assert_scores_close(got, want, nsnps, rel=1e-6):
    """north_star tolerance: 1e-6 relative per split.  Scores that are exact zeros in exact
    arithmetic (rank-deficient flattenings) come out of any SVD as rounding noise of size
    ~eps*sigma_max <= eps*nsnps, so differences below 1e-9*max(1,nsnps) are also accepted."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    floor = 1e-9 * np.maximum(1.0, np.asarray(nsnps, dtype=np.float64))[:, None]
    err = np.abs(got - want)
    ok = (err <= rel * np.maximum(np.abs(got), np.abs(want))) | (err <= floor)
    assert ok.all(), f"score mismatch: worst abs err {err[~ok].max()} at {np.argwhere(~ok)[:5]}"


def assert_best_matches(best, rscor_ref, rstat_ref):
    """Chosen split identical except at exact score ties and zero-SNP quartets (north_star)."""
    best = np.asarray(best).astype(np.int64)
    ref_best = rstat_ref[:, 0].astype(np.int64)
    srt = np.sort(rscor_ref, axis=1)
    tie = (srt[:, 1] - srt[:, 0]) <= 1e-9 * np.maximum(1.0, srt[:, 1])
    zero = rstat_ref[:, 1] == 0
    bad = (best != ref_best) & ~tie & ~zero
    assert not bad.any(), f"{bad.sum()} topology mismatches, first at {np.flatnonzero(bad)[:5]}"


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def lib():
    from tetrad_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.load()
