"""CPU: the N > 1 path (contiguous quartet shards + gather to rank 0) under gloo, world_size 2/3,
with the CPU oracle standing in for the GPU resolver."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, load_golden
from tetrad_b200 import sharding


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 495, 230300):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
    assert [sharding.replicate_owner(r, 4) for r in range(6)] == [0, 1, 2, 3, 0, 1]


def test_pack_roundtrip_is_bit_exact():
    rng = np.random.default_rng(0)
    rstat = torch.from_numpy(rng.integers(0, 2**31 - 1, size=(50, 2)).astype(np.int32))
    rscor = torch.from_numpy(rng.random((50, 3)))
    a, b = sharding.unpack_results(sharding.pack_results(rstat, rscor))
    assert torch.equal(a, rstat) and torch.equal(b, rscor)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    g = np.load(GOLDEN / "noisy12.npz")
    quartets = torch.from_numpy(g["quartets"].view(np.int32).copy())

    def resolve_fn(q):
        _, rstat, rscor = oracle.new_infer_resolved_quartets(
            g["tmparr"], g["tmpmap"], q.numpy().view(np.uint32), True)
        return torch.from_numpy(rstat.view(np.int32)), torch.from_numpy(rscor)

    out = sharding.resolve_sharded(quartets, resolve_fn)
    q2 = sharding.broadcast_quartets(g["quartets"][:10] if rank == 0 else None, src=0)
    assert q2.shape == (10, 4) and (q2.numpy().view(np.uint32) == g["quartets"][:10]).all()
    if rank == 0:
        rstat, rscor = out
        np.savez(os.path.join(outdir, "out.npz"), rstat=rstat.numpy(), rscor=rscor.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_resolve_matches_single_process(world, oracle):
    g = load_golden("noisy12")
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, _free_port(), tmp), nprocs=world, join=True)
        with np.load(os.path.join(tmp, "out.npz")) as out:
            rstat, rscor = out["rstat"].view(np.uint32), out["rscor"]
    assert (rstat == g["sub_rstat"]).all()                       # order preserved, bit-exact
    np.testing.assert_allclose(rscor, g["sub_rscor"], rtol=1e-12, atol=1e-12)


class _OracleEngine:
    """CPU stand-in for QuartetResolver (tests only)."""
    def load(self, tmparr, tmpmap):
        self.tmparr, self.tmpmap = tmparr, tmpmap

    def resolve(self, quartets, subsample):
        from oracle import oracle
        _, rstat, rscor = oracle.new_infer_resolved_quartets(self.tmparr, self.tmpmap, quartets.numpy().view(np.uint32), subsample)
        return torch.from_numpy(rstat.view(np.int32)), torch.from_numpy(rscor)


def _unseeded_worker(rank, world, port, workdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tetrad_b200 import run_inference
    from tetrad_b200.schema import Project
    proj = Project.load_json(os.path.join(workdir, "U.json"))
    assert proj.random_seed is None
    run_inference.run_inference(proj, nboots=1, resolver=_OracleEngine(), overlap_sink=False)
    dist.barrier()
    dist.destroy_process_group()


def test_unseeded_project_on_two_ranks_resolves_one_consistent_replicate(tmp_path, oracle):
    """An unseeded project (a JSON of the reference, or of an `init` without -r before the seed was persisted) under
    two ranks: every rank draws the replicate (bootstrap loci, random quartet ranks) from its own generator, so the
    seed must be agreed on first (run_inference.agreed_seed).  Every TSV row rank 0 writes must be the oracle's
    resolution of THAT row's quartet on the replicate-0 arrays -- with per-rank seeds rank 1's shard belongs to another
    quartet list."""
    from tetrad_b200 import database, run_inference, synth, write_database
    from tetrad_b200.schema import Project
    d = synth.generate(9, 80, 320, seed=5)
    database.write_snps_input(tmp_path / "in.npz", d.snps, d.snpsmap, d.names)
    proj = Project(version="0.1.0", data=tmp_path / "in.npz", name="U", workdir=tmp_path, subsample_snps=True,
                   weights=1, nquartets=60, random_seed=None)
    write_database.write_database(proj)
    assert proj.random_seed is not None                # init persists a concrete seed ...
    proj.random_seed = None                            # ... an older / foreign project file has none
    proj.save_json()
    mp.spawn(_unseeded_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    q, scores, best, nsnps = run_inference.read_quartets_tsv(tmp_path / "U.quartets_0.tsv")
    assert len(q) == 60
    arrays = database.read_arrays(proj.database_file, ["tmparr", "tmpmap"])
    _, ostat, oscor = oracle.new_infer_resolved_quartets(arrays["tmparr"], arrays["tmpmap"], q.astype(np.uint32), True)
    assert (nsnps == ostat[:, 1]).all()
    np.testing.assert_allclose(scores, oscor, rtol=1e-5, atol=2e-6)
    assert (tmp_path / "U.quartets_1.tsv").exists()
