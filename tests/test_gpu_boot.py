"""GPU (-m gpu): the device-side bootstrap (tq_plan_create_runs + tq_boot_pack) packs exactly the
replicate that the oracle-side restatement of its Philox streams materialises."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tetrad_b200
    return tetrad_b200


@pytest.mark.parametrize("shape,seed", [((9, 700, 2600), 1), ((6, 1500, 4000), 2), ((5, 3, 9), 3)])
def test_device_bootstrap_equals_oracle_sampler(oracle, shape, seed):
    import tetrad_b200
    from oracle import boot_oracle
    from tetrad_b200 import run_inference, synth
    from itertools import combinations
    n, nloci, nsnps = shape
    d = synth.generate(n, nloci, nsnps, seed=50 + seed, p_ambig=0.03)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=seed)
    rng = np.random.default_rng(seed)
    lidxs, s_shuf, s_amb = run_inference.draw_bootstrap(db["spans"], rng)
    quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)[:150]
    eng = tetrad_b200.QuartetResolver(device=0)
    eng.load_bootstrap(db["seqarr"], db["spans"], lidxs, s_shuf, s_amb)
    tmparr, tmpmap = boot_oracle.bootstrap_arrays(db["seqarr"], db["spans"], lidxs, s_shuf, s_amb)
    assert eng.info["ncols"] == tmparr.shape[1] and eng.info["nloci"] == nloci
    for sub in (True, False):
        rstat, rscor, counts, _ = eng.resolve(quartets, sub, want_counts=True)
        want = oracle.counts(tmparr, tmpmap, quartets, sub)[:, 0]
        assert (counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16) == want).all()
    # a second replicate on the same engine reuses the resident seqarr
    lidxs2, a2, b2 = run_inference.draw_bootstrap(db["spans"], rng)
    eng.load_bootstrap(eng.seqarr_dev, db["spans"], lidxs2, a2, b2)
    t2, m2 = boot_oracle.bootstrap_arrays(db["seqarr"], db["spans"], lidxs2, a2, b2)
    _, _, counts, _ = eng.resolve(quartets, True, want_counts=True)
    assert (counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16) == oracle.counts(t2, m2, quartets, True)[:, 0]).all()


def test_replicate_loop_uses_device_bootstrap(oracle, tmp_path):
    """run_inference with the GPU engine: replicate 1's TSV equals the oracle's resolution of the
    replicate that the same Generator stream + Philox seeds define."""
    import tetrad_b200
    from oracle import boot_oracle
    from tetrad_b200 import combinations as cb, database, run_inference, synth
    from tetrad_b200.schema import Project
    d = synth.generate(7, 200, 800, seed=31)
    database.write_snps_input(tmp_path / "in.npz", d.snps, d.snpsmap, d.names)
    proj = Project(version="0", data=tmp_path / "in.npz", name="B", workdir=tmp_path, subsample_snps=True,
                   weights=0, nquartets=0, random_seed=9)
    from tetrad_b200.write_database import write_database
    write_database(proj)
    run_inference.run_inference(proj, nboots=1)
    arrays = database.read_arrays(proj.database_file, ["seqarr", "spans"])
    rng = np.random.default_rng(9)
    lidxs, a, b = run_inference.draw_bootstrap(arrays["spans"], rng)
    tmparr, tmpmap = boot_oracle.bootstrap_arrays(arrays["seqarr"], arrays["spans"], lidxs, a, b)
    q, scores, best, nsnps = run_inference.read_quartets_tsv(tmp_path / "B.quartets_1.tsv")
    _, ostat, oscor = oracle.new_infer_resolved_quartets(tmparr, tmpmap, q.astype(np.uint32), True)
    assert (nsnps == ostat[:, 1]).all() and (best == ostat[:, 0]).all()
    np.testing.assert_allclose(scores, oscor, rtol=1e-5, atol=1e-6)


def test_tip_row_ranges_pack_the_same_planes(tb):
    """tq_pack_rows / tq_boot_pack_rows over any split of the tip rows write exactly the planes of the unsharded call
    (the ambiguity coins are keyed by the global tip index): what lets every rank of a multi-GPU run pack only its share
    before the all-gather."""
    import torch
    from tetrad_b200 import synth
    d = synth.generate(11, 900, 3600, seed=31, p_ambig=0.05)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=4)
    rng = np.random.default_rng(5)
    lidxs = rng.choice(len(db["spans"]), len(db["spans"]), replace=True)
    whole = tb.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    ref0 = whole.planes.clone()
    whole.load_bootstrap(db["seqarr"], db["spans"], lidxs, 123, 456)
    ref1 = whole.planes.clone()
    part = tb.QuartetResolver(device=0)
    for cuts in ((0, 4, 11), (0, 1, 2, 9, 11)):
        part.load(db["tmparr"], db["tmpmap"], tip_range=(0, 0))
        part.planes.zero_()
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            part.pack((lo, hi))
        assert torch.equal(part.planes, ref0)
        for i, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
            if i == 0:
                part.load_bootstrap(db["seqarr"], db["spans"], lidxs, 123, 456, tip_range=(lo, hi))
                rows = {(lo, hi): part.planes[lo:hi].clone()}
            else:
                part.load_bootstrap(db["seqarr"], db["spans"], lidxs, 123, 456, tip_range=(lo, hi))
                rows[(lo, hi)] = part.planes[lo:hi].clone()
        for (lo, hi), got in rows.items():
            assert torch.equal(got, ref1[lo:hi])
    whole.close(); part.close()


def test_device_sampler_equals_its_oracle_and_is_distinct(tb):
    """tq_sample_quartets: quartet i = unrank(pi_key(i)); bit-exact against the numpy restatement of the permutation,
    distinct, ascending tips, any index range on its own."""
    from math import comb
    from oracle import sample_oracle
    from tetrad_b200 import combinations as cb
    for n, key, count in ((12, 3, 495), (200, 0x1234567890ABCDEF, 60000), (500, 99, 20000)):
        got = cb.device_sample(n, key, count).cpu().numpy().view(np.uint32)
        want = cb.unrank_quartets(sample_oracle.permute(np.arange(count), n, key), n)
        assert (got == want).all()
        assert len(np.unique(got, axis=0)) == count and (np.diff(got.astype(np.int64), axis=1) > 0).all() and got.max() < n
        part = cb.device_sample(n, key, 1000 if count > 1000 else count // 2, start=count // 3).cpu().numpy().view(np.uint32)
        assert (part == want[count // 3: count // 3 + len(part)]).all()
    from tetrad_b200._lib import TetradB200Error
    with pytest.raises(TetradB200Error):
        cb.device_sample(12, 1, 10, start=comb(12, 4) - 5)
