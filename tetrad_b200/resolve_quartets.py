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
    rscor = np.empty((m, 3), dtype=np.float64)          # every element is written by the call (or it raises)
    rstat = np.empty((m, 2), dtype=np.uint32)
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

    # -- replicate state -----------------------------------------------------------------
    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if self.plan is not None:
            self.lib.tq_plan_destroy(self.plan)
            self.plan = None

    __del__ = close

    def load(self, tmparr, tmpmap) -> "QuartetResolver":
        """Upload ``tmparr`` (numpy u8[n,S] or a CUDA uint8 tensor) and pack it."""
        torch = self.torch
        locus_arr, locus_ptr, locus_stride = _as_locus(tmpmap)
        self.close()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tq_plan_create(locus_ptr, locus_arr.shape[0], locus_stride,
                                               self.device.index, self._stream(), ctypes.byref(h)))
        self.plan = h.value
        info = (ctypes.c_int64 * 8)()
        _lib.check(self.lib.tq_plan_info(self.plan, info))
        self.info = dict(ncols=info[0], nloci=info[1], ntiles=info[2], row_words=info[3],
                         max_len=info[4], padded_cols=info[5])
        if isinstance(tmparr, torch.Tensor):
            dev = tmparr.to(self.device, dtype=torch.uint8).contiguous()
        else:
            host = np.ascontiguousarray(tmparr, dtype=np.uint8)
            dev = torch.from_numpy(host).to(self.device, non_blocking=False)
        if dev.ndim != 2 or dev.shape[1] != self.info["ncols"]:
            raise ValueError("tmparr must be u8[ntips, ncols] with ncols == len(tmpmap)")
        self.tmparr_dev = dev
        self.ntips = int(dev.shape[0])
        self.planes = torch.empty((self.ntips, max(4, self.info["row_words"])), dtype=torch.int32,
                                  device=self.device)
        self.pack()
        return self

    def load_bootstrap(self, seqarr, spans, lidxs, seed_shuffle: int, seed_ambig: int) -> "QuartetResolver":
        """Build a bootstrap replicate on the device (``tq_plan_create_runs`` + ``tq_boot_pack``).

        ``seqarr`` (numpy u8[n,S0] or a CUDA uint8 tensor already resident from an earlier call):
        the database's ASCII matrix (``-`` already mapped to ``N``; ambiguity codes unresolved);
        ``spans`` i64[L,2]; ``lidxs``: the loci drawn with replacement for this replicate
        (``rng.choice(L, L)`` of the reference, run_inference.py:117); the two seeds are the two
        31-bit integers the reference draws next (:120, :123)."""
        torch = self.torch
        spans = np.asarray(spans, dtype=np.int64)
        lidxs = np.asarray(lidxs, dtype=np.int64)
        lens = np.ascontiguousarray(spans[lidxs, 1] - spans[lidxs, 0], dtype=np.uint32)
        starts = np.ascontiguousarray(spans[lidxs, 0], dtype=np.uint32)
        self.close()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tq_plan_create_runs(lens.ctypes.data, len(lens), self.device.index,
                                                    self._stream(), ctypes.byref(h)))
        self.plan = h.value
        info = (ctypes.c_int64 * 8)()
        _lib.check(self.lib.tq_plan_info(self.plan, info))
        self.info = dict(ncols=info[0], nloci=info[1], ntiles=info[2], row_words=info[3],
                         max_len=info[4], padded_cols=info[5])
        if isinstance(seqarr, torch.Tensor):
            dev = seqarr if seqarr.device == self.device and seqarr.dtype == torch.uint8 and seqarr.is_contiguous() \
                else seqarr.to(self.device, dtype=torch.uint8).contiguous()
        else:
            dev = torch.from_numpy(np.ascontiguousarray(seqarr, dtype=np.uint8)).to(self.device)
        self.seqarr_dev = dev
        self.ntips = int(dev.shape[0])
        src = torch.from_numpy(starts.view(np.int32)).to(self.device)
        scratch = torch.empty(max(1, self.info["padded_cols"]), dtype=torch.int32, device=self.device)
        self.planes = torch.empty((self.ntips, max(4, self.info["row_words"])), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.tq_boot_pack(self.plan, dev.data_ptr(), self.ntips, dev.stride(0), src.data_ptr(),
                                         int(seed_shuffle) & 0xFFFFFFFF, int(seed_ambig) & 0xFFFFFFFF,
                                         scratch.data_ptr(), self.planes.data_ptr(), self._stream()))
        self.tmparr_dev = None
        self._keep = (src, scratch)          # alive until the kernels have run
        return self

    def pack(self):
        """(Re)run K1 on the resident ``tmparr`` (exposed so bench.py can time it)."""
        _lib.check(self.lib.tq_pack(self.plan, self.tmparr_dev.data_ptr(), self.ntips,
                                    self.tmparr_dev.stride(0), self.planes.data_ptr(), self._stream()))

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
        if want_counts:
            return rstat, rscor, counts, minrank
        return rstat, rscor

    def check_status(self):
        """Raise if the last ``resolve`` saw a tip index >= ntips (synchronises)."""
        if int(self._last_status[0].item()) != 0:
            raise IndexError("a quartet referenced a tip index >= ntips")
