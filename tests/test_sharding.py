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
