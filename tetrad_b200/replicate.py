"""GPU side of the replicate loop: draw -> prepare -> resolve -> fetch, double-buffered.

The reference's ``run_inference`` (tetrad/src/run_inference.py:360-407) does, per replicate: resample the database on
the host and rewrite the HDF5 (:99-143, replicates >= 1), draw the quartet set (:381-384), farm the quartets out to
ipyparallel engines that each re-read the database (:184-251), collect.  Here one process per GPU does, per replicate:

``draw``     the reference's draws from the project's generator, in its order (host, microseconds to milliseconds);
``prepare``  locus plan (host) + bit-planes packed on the device from the resident ASCII ``seqarr`` (bootstrap:
             ``tq_boot_pack``; replicate 0: ``tq_pack``) + quartets unranked on the device (``tq_unrank``).  Under
             torchrun every rank packs only its share of the TIP ROWS and the planes are all-gathered over NVLink, so
             the per-replicate prepare shrinks with the number of GPUs like the resolve does;
``resolve``  this rank's contiguous quartet range (``sharding.shard_range``) + one gather of the 32 B/quartet results
             to rank 0;
``fetch``    rank 0 only: results (and the quartet ids the sink prints) to the host -- copied on a side stream
             behind the replicate's own completion event, so the next replicate's kernels are not waited for.

Two engines alternate, so ``prepare(r + 1)`` and ``resolve(r + 1)`` are issued while replicate r is still running: the
host part of the next replicate (draws, plan) overlaps the kernel, and the device goes from one replicate's resolve
straight into the next one's pack and resolve without waiting for the host.  ``run_inference`` and ``bench.py`` both
drive this class.

Random quartet samples (``nquartets < C(n,4)``): ``sampler="device"`` (default) takes ONE 63-bit key per replicate from
the project's generator and lets the device produce the sample (``tq_sample_quartets``: a keyed permutation of all
ranks, nothing drawn or uploaded per quartet); ``sampler="numpy"`` takes the reference's own draw,
``rng.choice(C(n,4), Q, replace=False)`` (tetrad/src/combinations.py:109-114) -- the same sample as the reference for
the same seed, but 0.13 s of serial host time per million quartets, which an 8-GPU loop would spend waiting.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import comb

import numpy as np

from . import combinations, sharding
from .schema import RNGStateModel


def draw_bootstrap(spans: np.ndarray, rng: np.random.Generator):
    """The three draws ``resample_tmp_database`` takes from the project's generator
    (run_inference.py:117, :120, :123): (lidxs, seed_shuffle, seed_ambig)."""
    nloci = spans.shape[0]
    lidxs = rng.choice(nloci, nloci, replace=True)
    return lidxs, int(rng.integers(2**31)), int(rng.integers(2**31))


@dataclass
class Prepared:
    """One replicate ready to resolve."""
    idx: int
    engine: object
    q_dev: object = None            # CUDA int32 [Q, 4] (device engines)
    quartets: np.ndarray = None     # host u32 [Q, 4] (exhaustive list, CPU stand-ins, or filled by fetch)
    q_host_pinned: object = None    # pinned staging + event of the async D2H copy of q_dev (rank 0)
    q_event: object = None
    done: object = None             # event after this replicate's resolve + gather on the compute stream
    fetched: object = None          # event after the D2H copies of its results on the copy stream (rank 0)
    host: dict = field(default_factory=dict)
    events: dict = field(default_factory=dict)


class ReplicatePipeline:
    def __init__(self, arrays: dict, nsamples: int, nqrts: int, nqrts_total: int, subsample_snps: bool,
                 device: int = 0, seed: int = 0, resolver=None, device_bootstrap: bool = True,
                 shard_pack: bool = True, timing: bool = False, sampler: str = "device"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.arrays = arrays
        self.nsamples, self.nqrts, self.subsample = int(nsamples), int(nqrts), bool(subsample_snps)
        self.exhaustive = int(nqrts) == int(nqrts_total)
        self.timing = timing
        if sampler not in ("device", "numpy"):
            raise ValueError(f"no quartet sampler {sampler!r}")
        self.sampler = sampler
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if resolver is None:
            from .resolve_quartets import QuartetResolver
            self.engines = [QuartetResolver(device=device, seed=seed), QuartetResolver(device=device, seed=seed)]
        else:
            self.engines = [resolver]           # CPU stand-ins (tests) resolve synchronously: one is enough
        eng = self.engines[0]
        self.device = getattr(eng, "device", None)
        self.gpu_index = getattr(self.device, "index", None)
        self.on_device = device_bootstrap and hasattr(eng, "load_bootstrap")
        self.shard_pack = shard_pack and self.world > 1 and self.gpu_index is not None
        self.full = self.full_dev = None
        if self.exhaustive and self.gpu_index is not None:
            self.full_dev = combinations.device_quartets(self.nsamples, self.gpu_index)   # unranked on the GPU, no upload
            self.full = self.full_dev.cpu().numpy().view(np.uint32)
        elif self.exhaustive:
            self.full = combinations.all_quartets(self.nsamples)
            self.full_dev = torch.from_numpy(self.full.view(np.int32))
        self.copy_stream = torch.cuda.Stream(self.device) if self.gpu_index is not None else None
        self._seq_dev = None
        self._pinned = {}                   # (what, engine slot) -> pinned staging buffers of rank 0's D2H copies

    # -- host draws ----------------------------------------------------------------------
    def draw(self, idx: int, rng: np.random.Generator) -> dict:
        """Every draw replicate ``idx`` takes from the project's generator, in the reference's order (its quartet
        generator is lazy: the quartet draw happens after the resampling draws), and the generator state after them."""
        from .run_inference import resample_arrays
        boot = host_arrays = ranks = drawn = key = None
        if idx and self.on_device:
            boot = draw_bootstrap(self.arrays["spans"], rng)
        elif idx:
            host_arrays = resample_arrays(self.arrays["seqarr"], self.arrays["spans"], rng)
        if self.full is None and self.gpu_index is not None and self.sampler == "device":
            key = int(rng.integers(2**63))                      # one draw; the device derives the sample from it
        elif self.full is None and self.gpu_index is not None:  # same draw as the reference, unranked on the GPU later
            ranks = np.random.default_rng(rng).choice(comb(self.nsamples, 4), size=self.nqrts, replace=False)
        elif self.full is None:
            drawn = combinations.random_quartets(self.nsamples, self.nqrts, rng)
        return dict(idx=idx, boot=boot, host_arrays=host_arrays, ranks=ranks, drawn=drawn, key=key,
                    state=RNGStateModel.from_rng(rng))

    # -- device prepare ------------------------------------------------------------------
    def _event(self):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream(self.device))
        return ev

    def prepare(self, draws: dict) -> Prepared:
        torch = self.torch
        idx = draws["idx"]
        eng = self.engines[idx % len(self.engines)]
        prep = Prepared(idx=idx, engine=eng)
        if self.timing and self.gpu_index is not None:
            prep.events["prepare_start"] = self._event()
        tips = sharding.shard_range(self.nsamples, self.rank, self.world) if self.shard_pack else None
        if draws["boot"] is not None:
            lidxs, seed_shuffle, seed_ambig = draws["boot"]
            if self._seq_dev is None:
                self._seq_dev = torch.from_numpy(np.ascontiguousarray(self.arrays["seqarr"], dtype=np.uint8)).to(self.device)
            eng.load_bootstrap(self._seq_dev, self.arrays["spans"], lidxs, seed_shuffle, seed_ambig, tip_range=tips)
        elif draws["host_arrays"] is not None:
            eng.load(*draws["host_arrays"])
            tips = None
        elif tips is not None:
            eng.load(self.arrays["tmparr"], self.arrays["tmpmap"], tip_range=tips)
        else:
            eng.load(self.arrays["tmparr"], self.arrays["tmpmap"])
        if tips is not None:
            self._allgather_planes(eng)
        if self.full is not None:
            prep.quartets, prep.q_dev = self.full, self.full_dev
        elif draws["ranks"] is not None or draws["key"] is not None:
            if draws["key"] is not None:
                prep.q_dev = combinations.device_sample(self.nsamples, draws["key"], self.nqrts, self.gpu_index)
            else:
                prep.q_dev = combinations.device_quartets(self.nsamples, self.gpu_index, ranks=draws["ranks"])
            if self.rank == 0:                  # the sink prints the tip ids: copy them back beside the resolve
                done = self._event()
                slot = idx % len(self.engines)
                key = ("q", slot)
                if key not in self._pinned or self._pinned[key].shape != prep.q_dev.shape:
                    self._pinned[key] = torch.empty(prep.q_dev.shape, dtype=torch.int32, pin_memory=True)
                prep.q_host_pinned = self._pinned[key]
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(done)
                    prep.q_host_pinned.copy_(prep.q_dev, non_blocking=True)
                    prep.q_event = torch.cuda.Event()
                    prep.q_event.record(self.copy_stream)
        else:
            prep.quartets = draws["drawn"]
            prep.q_dev = torch.from_numpy(prep.quartets.view(np.int32))
            if self.device is not None:
                prep.q_dev = prep.q_dev.to(self.device)
        if self.timing and self.gpu_index is not None:
            prep.events["prepare_end"] = self._event()
        return prep

    def _allgather_planes(self, eng):
        """Every rank packed tip rows [lo, hi) of ``eng.planes``; exchange them so that every rank holds all rows.
        Equal-size all-gather over the padded row blocks (the shares differ by at most one row)."""
        torch, dist = self.torch, self.dist
        width = sharding.shard_range(self.nsamples, 0, self.world)[1]          # rows of the largest share
        lo, hi = sharding.shard_range(self.nsamples, self.rank, self.world)
        row_words = eng.planes.shape[1]
        if width * self.world == self.nsamples:
            dist.all_gather_into_tensor(eng.planes, eng.planes[lo:hi])
            return
        send = torch.zeros((width, row_words), dtype=eng.planes.dtype, device=self.device)
        send[: hi - lo] = eng.planes[lo:hi]
        recv = torch.empty((self.world * width, row_words), dtype=eng.planes.dtype, device=self.device)
        dist.all_gather_into_tensor(recv, send)
        for r in range(self.world):
            rlo, rhi = sharding.shard_range(self.nsamples, r, self.world)
            if r != self.rank:
                eng.planes[rlo:rhi] = recv[r * width: r * width + (rhi - rlo)]

    # -- resolve + gather ----------------------------------------------------------------
    def resolve(self, prep: Prepared, want_counts: bool = False):
        """Launch the sharded resolve of a prepared replicate (asynchronous on device engines).  Returns
        ``(rstat, rscor)`` for all quartets on rank 0 (device tensors), ``None`` elsewhere; with ``want_counts``
        (single process) ``(rstat, rscor, counts)``."""
        eng = prep.engine
        if self.timing and self.gpu_index is not None:
            prep.events["resolve_start"] = self._event()
        if want_counts:
            rstat, rscor, counts, _ = eng.resolve(prep.q_dev, self.subsample, want_counts=True)
            out = (rstat, rscor, counts)
        else:
            def fn(q):
                res = eng.resolve(q, self.subsample)
                if self.timing and self.gpu_index is not None:
                    prep.events["resolve_end"] = self._event()
                return res
            out = sharding.resolve_sharded(prep.q_dev, fn)
        if self.timing and self.gpu_index is not None:
            prep.events.setdefault("resolve_end", self._event())
            prep.events["gather_end"] = self._event()
        if self.gpu_index is not None:
            prep.done = self.torch.cuda.Event()
            prep.done.record(self.torch.cuda.current_stream(self.device))
            if out is not None and not want_counts:
                self._start_fetch(prep, out)
        return out

    def _start_fetch(self, prep: Prepared, out):
        """Rank 0: results -> pinned host buffers on the copy stream, behind this replicate's completion event only."""
        torch = self.torch
        slot = prep.idx % len(self.engines)
        rstat, rscor = out[0], out[1]
        for name, t in (("rstat", rstat), ("rscor", rscor)):
            key = (name, slot)
            if key not in self._pinned or self._pinned[key].shape != t.shape or self._pinned[key].dtype != t.dtype:
                self._pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(prep.done)
            for name, t in (("rstat", rstat), ("rscor", rscor)):
                self._pinned[(name, slot)].copy_(t, non_blocking=True)
                prep.host[name] = self._pinned[(name, slot)]
            prep.fetched = torch.cuda.Event()
            prep.fetched.record(self.copy_stream)

    def wait(self, prep: Prepared):
        """Block the host until replicate ``prep`` (not the ones queued behind it) has finished on the device."""
        if prep.done is not None:
            prep.done.synchronize()

    def fetch(self, prep: Prepared, out):
        """Rank 0: (quartets u32[Q,4], rstat u32[Q,2], rscor f64[Q,3]) on the host.  Waits for THIS replicate's
        results only; replicates already queued behind it keep the device busy meanwhile."""
        if out is None:
            self.wait(prep)
            return None
        if prep.fetched is not None:
            prep.fetched.synchronize()
            rstat = prep.host["rstat"].numpy().view(np.uint32).copy()        # the staging buffers are reused
            rscor = prep.host["rscor"].numpy().copy()
        else:
            self.wait(prep)
            rstat, rscor = out[0].cpu().numpy().view(np.uint32), out[1].cpu().numpy()
        if prep.quartets is None:
            if prep.q_event is not None:
                prep.q_event.synchronize()
                prep.quartets = prep.q_host_pinned.numpy().view(np.uint32).copy()
            else:
                prep.quartets = prep.q_dev.cpu().numpy().view(np.uint32)
        return prep.quartets, rstat, rscor

    def close(self):
        for eng in self.engines:
            if hasattr(eng, "close"):
                eng.close()
