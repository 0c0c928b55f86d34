"""Multi-GPU sharding of the hot path: one process per GPU, contiguous quartet ranges, results
gathered to rank 0.

The reference farms quartet chunks out to ipyparallel engines and collects pickled results over
ZeroMQ (``distributor``, tetrad/src/run_inference.py:184-251).  Here every rank owns the
contiguous index range ``shard_range(Q, rank, world)`` of the replicate's quartet list, resolves
it on its own GPU against its own copy of the packed matrix (there is no exchange step inside
the path), and the 32 bytes per quartet of result travel to rank 0 in one ``gather`` -- NCCL
over NVLink on the GPU box, gloo in the CPU tests.  Rank 0 then owns ``(rstat, rscor)`` in
quartet order, exactly what ``distributor`` writes to ``{name}.quartets_{rep}.tsv``.

Nothing here touches the device directly: ``resolve_fn`` is any callable
``quartets[m,4] -> (rstat[m,2], rscor[m,3])`` on torch tensors (``QuartetResolver.resolve`` in
production; the tests inject the CPU oracle to exercise the world_size > 1 logic under gloo).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous share; the first ``n_items % world`` ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def replicate_owner(rep: int, world: int) -> int:
    """Rank that resolves replicate ``rep`` when replicates (not quartets) are dealt out."""
    return rep % world


def pack_results(rstat, rscor):
    """(rstat int32[m,2], rscor f64[m,3]) -> f64[m,4] (32 B / quartet; the ints ride bit-exact)."""
    import torch
    stat64 = rstat.contiguous().view(torch.int32).reshape(-1, 2).contiguous().view(torch.float64)
    return torch.cat([rscor.reshape(-1, 3), stat64.reshape(-1, 1)], dim=1).contiguous()


def unpack_results(packed):
    """Inverse of ``pack_results``."""
    import torch
    rscor = packed[:, :3].contiguous()
    rstat = packed[:, 3:4].contiguous().view(torch.int32).reshape(-1, 2)
    return rstat, rscor


def resolve_sharded(quartets, resolve_fn: Callable, group=None, dst: int = 0):
    """Resolve this rank's share of ``quartets`` (identical on every rank) and gather to ``dst``.

    Returns ``(rstat, rscor)`` for ALL quartets, in order, on rank ``dst``; ``None`` elsewhere.
    Without an initialised process group it degenerates to a plain ``resolve_fn`` call.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return resolve_fn(quartets)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = int(quartets.shape[0])
    lo, hi = shard_range(total, rank, world)
    rstat, rscor = resolve_fn(quartets[lo:hi])
    packed = pack_results(rstat, rscor)
    # equal-size gather: pad every shard to the largest one (differs by at most one row)
    width = shard_range(total, 0, world)[1]
    if packed.shape[0] < width:
        pad = torch.zeros((width - packed.shape[0], 4), dtype=packed.dtype, device=packed.device)
        packed = torch.cat([packed, pad], dim=0)
    bufs = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(total, r, world)
        parts.append(bufs[r][: rhi - rlo])
    return unpack_results(torch.cat(parts, dim=0))


def broadcast_quartets(quartets: Optional[np.ndarray], src: int = 0, device=None, group=None):
    """Give every rank rank ``src``'s quartet list (only needed when the list is not derived
    from a shared seed).  Returns an int32 tensor [m,4] on ``device``."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    n = torch.tensor([0 if quartets is None else len(quartets)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    if rank == src:
        q = torch.from_numpy(np.ascontiguousarray(quartets, dtype=np.uint32).view(np.int32)).to(device)
    else:
        q = torch.empty((int(n.item()), 4), dtype=torch.int32, device=device)
    dist.broadcast(q, src=src, group=group)
    return q
