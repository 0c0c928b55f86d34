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


EPS = float(np.finfo(np.float64).eps)
PARITY_LOG = {}          # test label -> worst errors seen (written to gpurun_out/ at session end)


def assert_scores_close(got, want, nsnps, rel=1e-6, counts=None, label=None):
    """north_star tolerance: 1e-6 relative per split.  Every singular value any backward-stable SVD returns
    carries an absolute error of a few eps * sigma_max, so a score that is an exact zero in exact arithmetic
    (rank-deficient flattenings) comes out of LAPACK as noise of that size while the Jacobi returns 0: differences
    below 64 * eps * sigma_max are accepted as well.  sigma_max <= ||T||_F, taken from the counts when the caller
    has them and bounded by nsnps (non-negative integers that sum to nsnps) otherwise.  Returns the worst relative
    error among scores above the floor and the worst absolute error in units of eps * ||T||_F; both are logged
    under `label`."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if counts is not None:
        fro = np.sqrt((np.asarray(counts, dtype=np.float64).reshape(len(got), -1) ** 2).sum(1))
    else:
        fro = np.asarray(nsnps, dtype=np.float64)
    fro = np.maximum(fro, 1.0)[:, None]
    floor = 64.0 * EPS * fro
    err = np.abs(got - want)
    mag = np.maximum(np.abs(got), np.abs(want))
    ok = (err <= rel * mag) | (err <= floor)
    above = mag > floor
    worst_rel = float((err[above] / mag[above]).max()) if above.any() else 0.0
    worst_abs = float((err / (EPS * fro)).max()) if err.size else 0.0
    label = label or os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0].split("::")[-1]
    if label:
        cur = PARITY_LOG.setdefault(label, dict(n=0, worst_rel=0.0, worst_abs_in_eps_fro=0.0))
        cur["n"] += int(got.shape[0])
        cur["worst_rel"] = max(cur["worst_rel"], worst_rel)
        cur["worst_abs_in_eps_fro"] = max(cur["worst_abs_in_eps_fro"], worst_abs)
    assert ok.all(), f"score mismatch: worst abs err {err[~ok].max()} (rel {worst_rel:.3g}) at {np.argwhere(~ok)[:5]}"
    return worst_rel, worst_abs


def pytest_sessionfinish(session, exitstatus):
    """Observed score errors of this session -> gpurun_out/parity_errors.json (copied to profiles/ per round)."""
    if not PARITY_LOG:
        return
    import json
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / "parity_errors.json").write_text(json.dumps(PARITY_LOG, indent=1, sort_keys=True) + "\n")
    except OSError:
        pass


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
