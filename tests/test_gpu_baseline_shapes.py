"""GPU (-m gpu): the CUDA path against the oracle at every BASELINE.json shape, both modes.

C1 exhaustive (all 495 quartets, every resolve kernel), C2 2 000 sampled quartets, C3 / C4 / C5 a few hundred sampled
quartets each -- what the oracle finishes in seconds.  Counts, nsnps and minrank must be bit-exact, scores within 1e-6
relative (floor 64 eps ||T||_F, conftest.assert_scores_close), best split identical off exact ties.  The worst errors
seen are written to gpurun_out/parity_errors.json at session end.  Nothing here reads /root/reference.
"""
from math import comb

import numpy as np
import pytest

from conftest import assert_best_matches, assert_scores_close

pytestmark = pytest.mark.gpu

_ARRAYS = {}


def config(name):
    """Working arrays of BASELINE config `name`, generated once per session (C5 is 1 GB)."""
    from tetrad_b200 import synth
    if name not in _ARRAYS:
        _ARRAYS[name] = synth.config_arrays(name)
    return _ARRAYS[name]


@pytest.fixture(scope="module")
def tb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tetrad_b200
    return tetrad_b200


def sampled_quartets(name, count, seed=42):
    """`count` quartets of the config's replicate-0 quartet set (exhaustive order for C1/C2, the reference's
    rng.choice draw for C3-C5), spread over the whole list."""
    from tetrad_b200 import combinations as cb, synth
    cfg = synth.CONFIGS[name]
    n = cfg["ntips"]
    if cfg["nquartets"] == 0:
        ranks = np.arange(comb(n, 4), dtype=np.int64)
    else:
        ranks = np.random.default_rng(seed).choice(comb(n, 4), size=cfg["nquartets"], replace=False)
    if count < len(ranks):
        ranks = ranks[np.linspace(0, len(ranks) - 1, count).astype(np.int64)]
    return cb.unrank_quartets(ranks, n)


def check_against_oracle(eng, oracle, db, quartets, sub, label):
    rstat, rscor, counts, minrank = eng.resolve(quartets, sub, want_counts=True)
    eng.check_status()
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, sub,
                                                              return_details=True)
    got_counts = counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16)
    assert (got_counts == det["mats"][:, 0]).all(), "site-pattern counts not bit-exact"
    assert (rstat[:, 1].cpu().numpy().view(np.uint32) == ostat[:, 1]).all()
    has = ostat[:, 1] > 0
    assert (minrank.cpu().numpy().view(np.uint32)[has] == det["minrank"][has]).all(), "minrank differs"
    worst = assert_scores_close(rscor.cpu().numpy(), oscor, ostat[:, 1], counts=got_counts, label=label)
    assert_best_matches(rstat[:, 0].cpu().numpy(), oscor, ostat)
    return worst


@pytest.mark.parametrize("kernel_name", ["fused", "mma", "pairgemm"])
@pytest.mark.parametrize("sub", [True, False])
def test_c1_exhaustive(tb, oracle, kernel_name, sub):
    """C1: 12 tips, 5 000 loci, 20 000 SNPs, all 495 quartets -- the reference's own CPU-runnable case."""
    from tetrad_b200 import _lib
    db = config("C1")
    quartets = sampled_quartets("C1", 495)
    assert len(quartets) == 495
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    _lib.set_kernel(kernel_name)
    try:
        check_against_oracle(eng, oracle, db, quartets, sub, f"C1_{'sub' if sub else 'full'}_{kernel_name}")
    finally:
        _lib.set_kernel("auto")
        eng.close()


@pytest.mark.parametrize("sub", [True, False])
def test_c2_sampled(tb, oracle, sub):
    """C2 (the single-GPU benchmark shape): 2 000 of the 230 300 quartets through the kernel large launches run."""
    from tetrad_b200 import _lib
    db = config("C2")
    quartets = sampled_quartets("C2", 2000)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    try:
        _lib.set_kernel("mma")
        check_against_oracle(eng, oracle, db, quartets, sub, f"C2_{'sub' if sub else 'full'}_mma")
    finally:
        _lib.set_kernel("auto")
        eng.close()


def test_c2_full_mode_whole_list_default_kernel(tb, oracle):
    """C2, full mode, the WHOLE exhaustive list in one launch through whatever kernel the size rule picks (the
    tip-pair GEMM for exhaustive lists): a sample of the results against the oracle, every count sum against nsnps."""
    from tetrad_b200 import combinations as cb
    db = config("C2")
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    quartets = cb.device_quartets(50)
    rstat, rscor, counts, minrank = eng.resolve(quartets, False, want_counts=True)
    from tetrad_b200 import _lib
    assert "pair_gemm_kernel" in _lib.load().tq_last_kernel().decode()      # an exhaustive list: the tcgen05 tip-pair GEMM
    pick = np.linspace(0, len(quartets) - 1, 600).astype(np.int64)
    qh = quartets.cpu().numpy().view(np.uint32)[pick]
    _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], qh, False, return_details=True)
    got_counts = counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16)
    assert (got_counts[pick] == det["mats"][:, 0]).all()
    assert (got_counts.reshape(len(got_counts), -1).sum(1) == rstat[:, 1].cpu().numpy().view(np.uint32)).all()
    assert (rstat.cpu().numpy().view(np.uint32)[pick] == ostat).all()
    assert (minrank.cpu().numpy().view(np.uint32)[pick] == det["minrank"]).all()
    assert_scores_close(rscor.cpu().numpy()[pick], oscor, ostat[:, 1], counts=got_counts[pick], label="C2_full_whole_list")
    eng.close()


@pytest.mark.parametrize("name,count,modes", [("C3", 300, (True, False)), ("C4", 200, (True, False)), ("C5", 200, (True,))])
def test_large_shapes_sampled(tb, oracle, name, count, modes):
    """C3 (100 tips, 500 k SNPs), C4 (200 tips, 1 M SNPs), C5 (500 tips, 2 M SNPs): sampled quartets of the
    replicate-0 draw, default kernel rule forced to the tensor-core kernel (what million-quartet launches run)."""
    from tetrad_b200 import _lib
    db = config(name)
    quartets = sampled_quartets(name, count)
    eng = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    _lib.set_kernel("mma")
    try:
        for sub in modes:
            check_against_oracle(eng, oracle, db, quartets, sub, f"{name}_{'sub' if sub else 'full'}_mma")
    finally:
        _lib.set_kernel("auto")
        eng.close()
        _ARRAYS.pop(name, None)          # C4 / C5 are hundreds of MB: one test each, then free
