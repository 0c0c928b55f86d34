"""Host side of the hot path: the reference's ``resolve_quartets`` interface over the C ABI.

Mirrors ``tetrad/src/resolve_quartets.py`` of the reference (same function names, argument
meaning and return contract) so that ``run_inference.distributor`` -- or a user -- can call it
unchanged:

* ``infer_resolved_quartets(database, nsamples, qrts, subsample_snps=True)``   (ref :17-39)
* ``new_infer_resolved_quartets(tmparr, tmpmap, quartets, subsample_snps)``    (ref :191-265)

plus the device-resident engine the replicate loop and ``bench.py`` use:

* ``QuartetResolver`` -- packs one replicate's ``tmparr``/``tmpmap`` into HBM bit-planes once
  (``tq_plan_create`` + ``tq_pack``) and resolves any number of quartet batches against it
  (``tq_resolve``), all on torch-owned device memory and the current torch stream.

Every path runs the hand-written CUDA kernels; there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes
import itertools
from pathlib import Path
from typing import Iterable

import numpy as np

from . import _lib

_CTX: dict[int, int] = {}
_EPOCH = itertools.count(1)


def _ctx(device: int) -> int:
    if device not in _CTX:
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.tq_ctx_create(device, ctypes.byref(h)))
        _CTX[device] = h.value
    return _CTX[device]


_UPLOAD_STREAMS: dict = {}


def upload(array: np.ndarray, device):
    """Host array -> device tensor without tying the host to the compute stream.  The copy runs on a side stream that
    is otherwise idle (so a copy from pageable memory blocks the host only for its own duration, not until a 40 ms
    resolve kernel ahead of it on the compute stream has finished) and the current stream waits for its event.  This
    is what lets the replicate loop prepare replicate r + 1 while replicate r resolves."""
    import torch
    device = torch.device(device)
    side = _UPLOAD_STREAMS.get(device)
    if side is None:
        side = _UPLOAD_STREAMS[device] = torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    with torch.cuda.stream(side):
        t = torch.from_numpy(array).to(device)
        done = torch.cuda.Event()
        done.record(side)
    cur.wait_event(done)
    t.record_stream(cur)
    return t


_PINNED_POOL: list = []          # [pinned uint8 tensor, weak reference to the ndarray handed out from it]


def _pinned_array(shape, dtype):
    """Result array backed by page-locked memory (device -> host copies into it are asynchronous DMA, not staged
    through a driver buffer).  The backing tensors are pooled: one is reused once every array handed out from it has
    been dropped by the caller (all views keep the tensor's base ndarray alive; a weak reference to it tells)."""
    import weakref
    import torch
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    entry = None
    for cand in _PINNED_POOL:
        if nbytes <= cand[0].numel() <= 2 * nbytes + 4096 and cand[1]() is None:
            entry = cand
            break
    if entry is None:
        entry = [torch.empty(max(nbytes, 16), dtype=torch.uint8, pin_memory=True), None]
        _PINNED_POOL.append(entry)
        if len(_PINNED_POOL) > 16:
            _PINNED_POOL.pop(0)
    base = entry[0].numpy()
    entry[1] = weakref.ref(base)
    return base[:nbytes].view(dtype).reshape(shape)


def _as_locus(tmpmap: np.ndarray):
    """(array kept alive, pointer, element stride) of the locus-id column tmpmap[:, 0]."""
    tmpmap = np.asarray(tmpmap)
    if tmpmap.dtype != np.uint32 or not tmpmap.flags.c_contiguous:
        tmpmap = np.ascontiguousarray(tmpmap, dtype=np.uint32)
    if tmpmap.ndim == 1:
        return tmpmap, tmpmap.ctypes.data, 1
    if tmpmap.ndim != 2 or tmpmap.shape[1] < 1:
        raise ValueError("tmpmap must be u32[S] or u32[S, >=1]")
    return tmpmap, tmpmap.ctypes.data, tmpmap.shape[1]


def new_infer_resolved_quartets(tmparr: np.ndarray, tmpmap: np.ndarray, quartets: np.ndarray,
                                subsample_snps: bool, *, seed: int = 0, device: int = 0,
                                return_counts: bool = False, epoch: int | None = None):
    """Resolve ``quartets`` on the GPU.  Same contract as the reference function
    (tetrad/src/resolve_quartets.py:191-265): returns ``(quartets u32[m,4], rstat u32[m,2],
    rscor f64[m,3])`` with ``rstat[:,0]`` the best split (0: 01|23, 1: 02|13, 2: 03|12) and
    ``rstat[:,1]`` the number of SNPs counted.

    ``return_counts=True`` appends a dict with ``counts`` u32[m,16,16] (= the reference's
    ``mats[0]``) and ``minrank`` u32[m].  ``epoch`` (any non-zero int) lets repeated calls that
    pass the *same* tmparr/tmpmap objects reuse the packed matrix already on the device -- the
    reference's engines re-read the database for every chunk (:33-35).
    """
    lib = _lib.load()
    tmparr = np.asarray(tmparr)
    if tmparr.dtype != np.uint8 or tmparr.ndim != 2 or tmparr.strides[1] != 1:
        tmparr = np.ascontiguousarray(tmparr, dtype=np.uint8)
        if tmparr.ndim != 2:
            raise ValueError("tmparr must be u8[ntips, ncols]")
    locus_arr, locus_ptr, locus_stride = _as_locus(tmpmap)
    ntips, ncols = tmparr.shape
    if locus_arr.shape[0] != ncols:
        raise ValueError("tmpmap and tmparr disagree on the number of columns")
    quartets_in = quartets
    quartets = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
    m = quartets.shape[0]
    rscor = _pinned_array((m, 3), np.float64)           # every element is written by the call (or it raises)
    rstat = _pinned_array((m, 2), np.uint32)
    counts = np.empty((m, 16, 16), dtype=np.uint32) if return_counts else None
    minrank = np.empty(m, dtype=np.uint32) if return_counts else None
    _lib.check(lib.tq_resolve_host(
        _ctx(device), tmparr.ctypes.data, ntips, ncols, tmparr.strides[0],
        locus_ptr, locus_stride, quartets.ctypes.data, m, int(bool(subsample_snps)),
        int(seed) & (2**64 - 1), int(epoch or 0),
        rscor.ctypes.data, rstat.ctypes.data,
        counts.ctypes.data if return_counts else None,
        minrank.ctypes.data if return_counts else None))
    out_q = quartets_in if isinstance(quartets_in, np.ndarray) and quartets_in.dtype == np.uint32 \
        and quartets_in.ndim == 2 else quartets
    if return_counts:
        return out_q, rstat, rscor, dict(counts=counts, minrank=minrank)
    return out_q, rstat, rscor


def last_traffic(device: int = 0) -> tuple[int, int]:
    """(host->device, device->host) bytes of the last ``new_infer_resolved_quartets`` call."""
    out = (ctypes.c_int64 * 2)()
    _lib.check(_lib.load().tq_ctx_last_traffic(_ctx(device), out))
    return int(out[0]), int(out[1])


def load_database(database: Path):
    """``tmparr``, ``tmpmap`` of a working database (reference :33-35); see ``database.py`` for
    the container formats (numpy archive, or HDF5 via h5py / the bundled subset reader)."""
    from . import database as dbio
    arrays = dbio.read_arrays(Path(database), ["tmparr", "tmpmap"])
    return arrays["tmparr"], arrays["tmpmap"]


def infer_resolved_quartets(database: Path, nsamples: int, qrts: Iterable[tuple[int, int, int, int]],
                            subsample_snps: bool = True):
    """Return a chunk of resolved quartets; signature of the reference's engine entry point
    (tetrad/src/resolve_quartets.py:17-39)."""
    qrts = np.array(list(qrts), dtype=np.uint32).reshape(-1, 4)
    tmparr, tmpmap = load_database(database)
    return new_infer_resolved_quartets(tmparr, tmpmap, qrts, subsample_snps)


class QuartetResolver:
    """One replicate's SNP matrix resident in HBM as bit-planes; resolves quartet batches.

    >>> eng = QuartetResolver(device=0)
    >>> eng.load(tmparr, tmpmap)                    # H2D + tq_plan_create + tq_pack
    >>> rstat, rscor = eng.resolve(quartets_u32)    # torch tensors on the device
    """

    def __init__(self, device: int = 0, seed: int = 0):
        import torch
        self.plan = None
        if not torch.cuda.is_available():
            raise RuntimeError("tetrad_b200 needs a CUDA device (B200); no CPU path exists")
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.seed = int(seed) & (2**64 - 1)
        self.plan = None
        self.info = None
        self.planes = None
        self.ntips = 0
        self.tmparr_dev = None
        self.seqarr_dev = None
        self._plan_stream = None             # torch stream the plan's tables belong to
        self._last_use = None                # event after the last resolve issued under another stream
        self._plane_store = None
        self._srccol = None

    # -- replicate state -----------------------------------------------------------------
    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        """Destroy the plan.  Its device tables are stream-ordered memory of the stream the plan was created on
        (``tq_plan_destroy`` frees them in that stream's order): a ``resolve`` issued under another torch stream is
        ordered before the free through the event it left behind."""
        if self.plan is not None:
            if self._last_use is not None and self._plan_stream is not None:
                self._plan_stream.wait_event(self._last_use)
            self.lib.tq_plan_destroy(self.plan)
            self.plan = None
            self._last_use = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _new_plan(self, create) -> None:
        torch = self.torch
        self.close()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._plan_stream = torch.cuda.current_stream(self.device)
            _lib.check(create(self._plan_stream.cuda_stream, ctypes.byref(h)))
        self.plan = h.value
        info = (ctypes.c_int64 * 8)()
        _lib.check(self.lib.tq_plan_info(self.plan, info))
        self.info = dict(ncols=info[0], nloci=info[1], ntiles=info[2], row_words=info[3],
                         max_len=info[4], padded_cols=info[5])

    def _planes_for(self, ntips: int):
        """Plane buffer of the replicate; reused between replicates while it is large enough."""
        torch = self.torch
        words = max(4, self.info["row_words"])
        need = ntips * words
        if self._plane_store is None or self._plane_store.numel() < need:
            self._plane_store = torch.empty(need + need // 16, dtype=torch.int32, device=self.device)
        return self._plane_store[:need].view(ntips, words)

    def load(self, tmparr, tmpmap, tip_range=None) -> "QuartetResolver":
        """Upload ``tmparr`` (numpy u8[n,S] or a CUDA uint8 tensor) and pack it.  ``tip_range=(lo, hi)`` packs only
        those tip rows (the caller all-gathers the planes: ``replicate.ReplicatePipeline``)."""
        torch = self.torch
        locus_arr, locus_ptr, locus_stride = _as_locus(tmpmap)
        self._new_plan(lambda stream, out: self.lib.tq_plan_create(locus_ptr, locus_arr.shape[0], locus_stride,
                                                                   self.device.index, stream, out))
        if isinstance(tmparr, torch.Tensor):
            dev = tmparr.to(self.device, dtype=torch.uint8).contiguous()
        else:
            dev = upload(np.ascontiguousarray(tmparr, dtype=np.uint8), self.device)
        if dev.ndim != 2 or dev.shape[1] != self.info["ncols"]:
            raise ValueError("tmparr must be u8[ntips, ncols] with ncols == len(tmpmap)")
        self.tmparr_dev = dev
        self.ntips = int(dev.shape[0])
        self.planes = self._planes_for(self.ntips)
        self.pack(tip_range)
        return self

    def load_bootstrap(self, seqarr, spans, lidxs, seed_shuffle: int, seed_ambig: int, tip_range=None) -> "QuartetResolver":
        """Build a bootstrap replicate on the device (``tq_plan_create_runs`` + ``tq_boot_pack``).

        ``seqarr`` (numpy u8[n,S0] or a CUDA uint8 tensor already resident from an earlier call):
        the database's ASCII matrix (``-`` already mapped to ``N``; ambiguity codes unresolved);
        ``spans`` i64[L,2]; ``lidxs``: the loci drawn with replacement for this replicate
        (``rng.choice(L, L)`` of the reference, run_inference.py:117); the two seeds are the two
        31-bit integers the reference draws next (:120, :123).  ``tip_range`` as in ``load``."""
        torch = self.torch
        spans = np.asarray(spans, dtype=np.int64)
        lidxs = np.asarray(lidxs, dtype=np.int64)
        lens = np.ascontiguousarray(spans[lidxs, 1] - spans[lidxs, 0], dtype=np.uint32)
        starts = np.ascontiguousarray(spans[lidxs, 0], dtype=np.uint32)
        self._new_plan(lambda stream, out: self.lib.tq_plan_create_runs(lens.ctypes.data, len(lens), self.device.index,
                                                                        stream, out))
        if isinstance(seqarr, torch.Tensor):
            dev = seqarr if seqarr.device == self.device and seqarr.dtype == torch.uint8 and seqarr.is_contiguous() \
                else seqarr.to(self.device, dtype=torch.uint8).contiguous()
        else:
            dev = torch.from_numpy(np.ascontiguousarray(seqarr, dtype=np.uint8)).to(self.device)
        self.seqarr_dev = dev
        self.ntips = int(dev.shape[0])
        src = upload(starts.view(np.int32), self.device)
        need = max(1, self.info["padded_cols"])
        if self._srccol is None or self._srccol.numel() < need:
            self._srccol = torch.empty(need + need // 16, dtype=torch.int32, device=self.device)
        self.planes = self._planes_for(self.ntips)
        lo, hi = tip_range if tip_range is not None else (0, self.ntips)
        _lib.check(self.lib.tq_boot_pack_rows(self.plan, dev.data_ptr(), self.ntips, dev.stride(0), lo, hi, src.data_ptr(),
                                              int(seed_shuffle) & 0xFFFFFFFF, int(seed_ambig) & 0xFFFFFFFF,
                                              self._srccol.data_ptr(), self.planes.data_ptr(), self._stream()))
        self.tmparr_dev = None
        self._keep = src                     # alive until the kernels have run
        return self

    def pack(self, tip_range=None):
        """(Re)run K1 on the resident ``tmparr`` (exposed so bench.py can time it)."""
        lo, hi = tip_range if tip_range is not None else (0, self.ntips)
        _lib.check(self.lib.tq_pack_rows(self.plan, self.tmparr_dev.data_ptr(), self.ntips,
                                         self.tmparr_dev.stride(0), lo, hi, self.planes.data_ptr(), self._stream()))

    # -- resolve -------------------------------------------------------------------------
    def resolve(self, quartets, subsample_snps: bool = True, want_counts: bool = False):
        """Resolve ``quartets`` (numpy u32[m,4] or CUDA int32/uint32 tensor [m,4]).  Returns
        device tensors ``(rstat int32[m,2], rscor float64[m,3])`` (+ ``counts int32[m,256]``,
        ``minrank int32[m]`` when ``want_counts``); asynchronous on the current stream."""
        torch = self.torch
        if self.plan is None:
            raise RuntimeError("QuartetResolver.load() has not been called")
        if isinstance(quartets, torch.Tensor):
            q = quartets.to(self.device).contiguous()
            if q.dtype not in (torch.int32, torch.uint32):
                q = q.to(torch.int32)
        else:
            qh = np.ascontiguousarray(quartets, dtype=np.uint32).reshape(-1, 4)
            q = torch.from_numpy(qh.view(np.int32)).to(self.device)
        m = int(q.shape[0])
        rscor = torch.empty((m, 3), dtype=torch.float64, device=self.device)
        rstat = torch.empty((m, 2), dtype=torch.int32, device=self.device)
        status = torch.zeros(4, dtype=torch.int32, device=self.device)
        counts = torch.empty((m, 256), dtype=torch.int32, device=self.device) if want_counts else None
        minrank = torch.empty(m, dtype=torch.int32, device=self.device) if want_counts else None
        _lib.check(self.lib.tq_resolve(
            self.plan, self.planes.data_ptr(), self.ntips, q.data_ptr(), m,
            int(bool(subsample_snps)), self.seed, rscor.data_ptr(), rstat.data_ptr(),
            counts.data_ptr() if want_counts else None,
            minrank.data_ptr() if want_counts else None, status.data_ptr(), self._stream()))
        self._last_status = status
        cur = torch.cuda.current_stream(self.device)
        if self._plan_stream is not None and cur != self._plan_stream:
            self._last_use = torch.cuda.Event()
            self._last_use.record(cur)
        if want_counts:
            return rstat, rscor, counts, minrank
        return rstat, rscor

    def check_status(self):
        """Raise if the last ``resolve`` saw a tip index >= ntips (synchronises)."""
        if int(self._last_status[0].item()) != 0:
            raise IndexError("a quartet referenced a tip index >= ntips")
