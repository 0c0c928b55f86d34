"""GPU (-m gpu): the device-side bootstrap (tq_plan_create_runs + tq_boot_pack) packs exactly the
replicate that the oracle-side restatement of its Philox streams materialises."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
