#!/usr/bin/env python
"""Benchmark of the quartet-resolution hot path (BASELINE.json metric: quartets resolved/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl reference]

Default workload: BASELINE config C4, the one `north_star` quotes its target on (200 tips, 200 000 loci, 1 M SNPs,
1 M random quartets per replicate, bootstrap replicates; it fits one GPU: 200 MB of seqarr, 105 MB packed).  One "step" = ONE BOOTSTRAP
REPLICATE through the replicate loop the CLI runs (`tetrad_b200.replicate.ReplicatePipeline`, what
`run_inference` drives): bootstrap planes packed on the device from the resident `seqarr` (every rank packs its share
of the tip rows, planes all-gathered over NVLink), quartet ranks unranked on the device, this rank's contiguous
quartet range resolved (one SNP per locus -> 256 site-pattern counts -> 3 SVD scores -> best split), results gathered
to rank 0 with NCCL.  STRONG scaling: a replicate has the same 1 M quartets whatever N is.

* `value`: Q * K / device time of K replicates (one CUDA-event pair around the whole pipelined region on the launching
  stream, max over ranks; barrier + synchronize on both sides).  The replicate's draws (bootstrap loci, quartet ranks)
  are taken before the region starts and `seqarr` is resident; every replicate is a different draw packed from the
  200 MB `seqarr` into a new 105 MB plane matrix (two engines alternate: > L2 together with the results).
* `e2e`: the same loop through the same public API by wall clock, now with the host draws of every replicate, their
  host->device copies, and the device->host read of the replicate's results (and quartet ids) on rank 0 inside.
* `--config C1|C2` (no bootstraps, exhaustive list): a step re-resolves replicate 0 against the resident matrix
  (L2 flushed between steps); e2e = `tetrad_b200.new_infer_resolved_quartets` on pinned host arrays.
* N = 1 adds `cpu_baseline` (oracle port on the host cores, bounded sample), and for the default config the blocks
  `c2` (the single-GPU shape of BASELINE configs[1]) and `full_mode` (subsample_snps=False, the CLI's default mode).
* `--impl reference`: the reference's CPU implementation of the path.  The reference is pure Python (numba/numpy) and
  /root/reference does not exist on the GPU box, so this arm times the oracle port (oracle/: C count loops + the same
  numpy.linalg svd / matrix_rank calls) on all host cores over a bounded random sample of the same workload per step;
  `port_vs_reference` carries the calibration against the executed reference (profiles/cpu_port_vs_reference.json).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time
from math import comb
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from tetrad_b200 import synth  # noqa: E402

print_result = print
METRIC = "quartets_resolved_per_sec"
UNIT = "quartets/s"
DEFAULT_CONFIG = "C4"


# ----------------------------------------------------------------------------------------------
def describe(name: str, subsample: bool = True) -> dict:
    """`config` of the JSON line: the same dict in both arms."""
    cfg = synth.CONFIGS[name]
    n = cfg["ntips"]
    nq = cfg["nquartets"] or comb(n, 4)
    kind = "random quartets per replicate (redrawn every replicate)" if cfg["nquartets"] else "quartets (exhaustive)"
    reps = "bootstrap replicates (per-replicate SNP resampling)" if cfg["nboots"] else "replicate 0 only"
    return {"workload": f"{name}: {n} tips, {cfg['nloci']} loci, {cfg['nsnps']} SNPs, {nq} {kind}, "
                        f"subsample_snps={subsample}, {reps}"}


def replicate0_quartets(name: str, seed: int = 42) -> np.ndarray:
    from tetrad_b200 import combinations as cb
    cfg = synth.CONFIGS[name]
    n = cfg["ntips"]
    if cfg["nquartets"] == 0:
        return cb.all_quartets(n)
    return cb.random_quartets(n, cfg["nquartets"], np.random.default_rng(seed))


def algorithmic_bytes_per_quartet(ncols: int, subsample: bool = True) -> float:
    """4 tips x 4 one-hot planes x S/8 (+ S/8 locus plane in subsample
    mode) + 16 B quartet ids + 32 B result."""
    return (2.125 if subsample else 2.0) * ncols + 48.0


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, label in names.items():
                    if mask & bit:
                        self.reasons.add(label)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_port_rate(tmparr, tmpmap, quartets, sample: int, seed: int = 0, subsample: bool = True):
    """quartets/s of the oracle port on all host cores over a random sample of the workload."""
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(quartets), size=min(sample, len(quartets)), replace=False))
    sub = np.ascontiguousarray(quartets[pick])
    oracle.resolve_parallel(tmparr, tmpmap, sub[: 4 * cores], subsample, nproc=cores)   # pool + BLAS warm-up
    t0 = time.perf_counter()
    oracle.resolve_parallel(tmparr, tmpmap, sub, subsample, nproc=cores)
    dt = time.perf_counter() - t0
    return len(sub) / dt, cores, len(sub), dt


def port_calibration():
    path = ROOT / "profiles" / "cpu_port_vs_reference.json"
    if not path.exists():
        return None
    cal = json.loads(path.read_text())
    return {"port_over_reference_subsample": cal.get("port_over_reference_subsample"),
            "port_over_reference_full": cal.get("port_over_reference_full"),
            "note": "quartets/s per core of the oracle port divided by that of the executed reference "
                    "(numba path, JIT warmed) on the same inputs, measured in the build container: "
                    "profiles/cpu_port_vs_reference.json; > 1 means the port is the faster of the two"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    db = synth.config_arrays(args.config, seed=1234)
    tmparr, tmpmap = np.ascontiguousarray(db["tmparr"]), np.ascontiguousarray(db["tmpmap"])
    quartets = replicate0_quartets(args.config)
    # size the per-step sample so the whole run stays within a few minutes
    probe_rate, cores, _, _ = cpu_port_rate(tmparr, tmpmap, quartets, sample=16 * (os.cpu_count() or 1))
    budget_s = 100.0 / max(1, args.steps + args.warmup)
    sample = int(max(4 * cores, min(len(quartets), probe_rate * min(15.0, budget_s))))
    times, n_done = [], 0
    for step in range(args.warmup + args.steps):
        rate, cores, n, dt = cpu_port_rate(tmparr, tmpmap, quartets, sample, seed=step)
        if step >= args.warmup:
            times.append(dt)
            n_done += n
    value = n_done / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 counts + f64 svd (numpy / LAPACK)", "data": "synthetic",
        "config": describe(args.config),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} random quartets of the replicate-0 quartet set per step (the replicate's "
                                   f"matrix has the same shape whatever the bootstrap draw), fork pool of {cores} "
                                   "workers, OPENBLAS_NUM_THREADS=1 (the reference is pure Python and cannot travel to "
                                   "the GPU box; the oracle port is timed)"},
        "port_vs_reference": port_calibration(),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_result(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def setup_distributed():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"           # keep NCCL's banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def barrier(world, dev):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def max_over_ranks(x: float, world, dev) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measured_peaks(lib, local):
    from tetrad_b200 import _lib
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak_gbs, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak_gbs, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    out = {}
    for key, kind in (("popc", 0), ("lop3", 1), ("imma", 4)):
        v = ctypes.c_double()
        _lib.check(lib.tq_microbench(local, kind, ctypes.byref(v)))
        out[key] = v.value
    return peak_gbs, peak_src, out


def ncu_record(config: str, subsample: bool):
    """What the committed ncu capture of this configuration's resolve launch says (profiles/resolve_ncu_<cfg>.json):
    DRAM bytes per launch for `roofline.traffic` and the counters of the `binding` block.  None when no capture of
    this shape exists -- never another shape's numbers."""
    path = ROOT / "profiles" / f"resolve_ncu_{config}{'' if subsample else '_full'}.json"
    return json.loads(path.read_text()) if path.exists() else None


def stacking_stats(lib, engine, run_once, nq):
    """One extra, untimed launch with the kernel's stacking diagnostics on (tq_set_option("stack_stats")): tile products
    taken, tiles stacked on earlier ones and loci counted one by one (busiest lane), per quartet of this rank."""
    from tetrad_b200 import _lib
    import torch
    _lib.check(lib.tq_set_option(b"stack_stats", 1))
    try:
        run_once()
        torch.cuda.synchronize()
        st = engine._last_status.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    finally:
        _lib.check(lib.tq_set_option(b"stack_stats", 0))
    if nq <= 0 or st[1] == 0:
        return None
    return {"tile_products_per_quartet": float(st[1]) / nq, "tiles_stacked_per_quartet": float(st[2]) / nq,
            "overflow_walk_per_quartet": float(st[3]) / nq, "count_warps": int(lib.tq_last_count_warps()),
            "note": "subsample mode stacks the merged planes of consecutive sparse tiles before ONE 16 x 16 x 1024 tile product; "
                    "loci whose bit slot is taken are counted by shared-memory atomics (walk = the busiest lane's share); "
                    "status words 1..3 of an extra untimed launch over at most 2^20 quartets of the same list"}


def roofline_block(config, ncols, quartets_per_launch, kern_ms, info, lib, local, subsample=True, kernel=None, stacking=None):
    peak_gbs, peak_src, pipes = measured_peaks(lib, local)
    kern_s = kern_ms * 1e-3
    alg = algorithmic_bytes_per_quartet(ncols, subsample)
    achieved = alg * quartets_per_launch / kern_s / 1e9
    rec = ncu_record(config, subsample)
    ntiles, chunks = info["ntiles"], info["row_words"] / 128.0
    products = ntiles if subsample else chunks               # 16 x 16 x 1024 tile products per quartet
    if subsample and stacking:
        products = stacking["tile_products_per_quartet"]     # stacked tiles: fewer products than tiles
    imma_per_quartet = 64.0 * products
    alu_per_quartet = 32.0 * (320.0 * products + (24.0 * chunks if subsample else 0.0))   # stage 1 on code planes: 18 LOP3 + loop arithmetic
    block = {
        "bound": "hbm", "kernel": kernel or lib.tq_last_kernel().decode(), "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": rec.get("dram_bytes_per_launch") if rec else None, "peak_source": peak_src,
        "algorithmic_bytes_per_quartet": alg, "quartets_per_launch": quartets_per_launch, "kernel_ms": kern_ms,
        "note": "SURVEY 8(d) accounting: every quartet counts its four tip rows once; the rows come from L2 (and, past "
                "126 MB of planes, partly from HBM), so frac > 1 is possible and `traffic` (ncu DRAM bytes of the same "
                "launch shape, when a capture is committed) is far below the algorithmic bytes -- see `binding` for "
                "the resource that limits the kernel",
        "popc": {"frac": 256.0 * 32.0 * ntiles * quartets_per_launch / kern_s / pipes["popc"], "peak_popc_per_s": pipes["popc"],
                 "note": "rate a 256-POPC-per-32-loci count would need; the tensor-core count issues no POPC"},
        "imma": {"frac": imma_per_quartet * quartets_per_launch / kern_s / pipes["imma"], "peak_per_s": pipes["imma"],
                 "note": "warp-level IMMA.16832.U8.U8 of the count stage against tq_microbench kind 4"},
        "alu": {"frac": alu_per_quartet * quartets_per_launch / kern_s / (2.0 * pipes["lop3"]), "peak_per_s": 2.0 * pipes["lop3"],
                "note": "LOP3/SHF thread-ops of the bit-plane unpack and of the per-locus selection against tq_microbench kind 1"},
    }
    if stacking:
        block["stacking"] = stacking
    binding = None
    if rec:
        binding = {k: rec.get(k) for k in ("issue_active_pct", "alu_pipe_pct", "fp64_pipe_pct", "lsu_pipe_pct",
                                            "l2_to_sm_bytes_per_launch", "l2_hit_rate_pct", "l2_stream_peak_bytes_per_s",
                                            "source")}
        if rec.get("l2_to_sm_bytes_per_launch") and rec.get("l2_stream_peak_bytes_per_s"):
            binding["l2_to_sm_frac_of_stream_peak"] = rec["l2_to_sm_bytes_per_launch"] / kern_s / rec["l2_stream_peak_bytes_per_s"]
        binding["tensor_pipe_pct"] = rec.get("tensor_pipe_pct")
        binding["note"] = ("ncu counters of the committed capture of this launch shape.  The legacy IMMA and the ALU pipe (LOP3 / SHF "
                           "of the bit-plane unpack and the per-locus selection) share an issue port: alu_pipe_pct + tensor_pipe_pct "
                           "is the port's load.  issue_active_pct includes the mbarrier poll loops of idle score warps (a fifth of "
                           "the executed instructions at C4, harmless for the run time: profiles/r02_warp_mix.md section 7); "
                           "the count warps are bound by the latency of their own dependent instruction streams")
    return block, binding


def pairgemm_tiles(quartets: np.ndarray, ntips: int) -> int:
    """Tiles (8 row pairs x 16 column pairs) of the tip-pair GEMM that hold at least one quartet of the list
    (the device planner's count, recomputed on the host for the roofline arithmetic; tq_pairgemm.cu)."""
    q = quartets.astype(np.int64)
    n = ntips
    cp = q[:, 2] * (2 * n - 1 - q[:, 2]) // 2 + q[:, 3] - q[:, 2] - 1
    npairs = n * (n - 1) // 2
    ntc = (npairs + 15) // 16
    # the planner counts the tiles under both row-pair numberings (lexicographic / by (t1, t0)) and takes the fewer
    rows = (q[:, 0] * (2 * n - 1 - q[:, 0]) // 2 + q[:, 1] - q[:, 0] - 1, q[:, 1] * (q[:, 1] - 1) // 2 + q[:, 0])
    return min(int(len(np.unique((rp // 8) * ntc + cp // 16))) for rp in rows)


def tensor_block(quartets, ntips, padded_cols, kern_ms):
    """Tensor-core roofline of the full-mode GEMM: u8 MACs the launch executed (non-empty tiles x 128 x 256 x K)
    against twice the measured dense bf16 peak (int8 runs at twice the bf16 rate; MEASURED_PEAKS.json)."""
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    bf16 = json.loads(peaks_path.read_text())["bf16_tflops"] if peaks_path.exists() else 1590.0
    tiles = pairgemm_tiles(quartets, ntips)
    executed = 2.0 * tiles * 128 * 256 * padded_cols
    useful = 2.0 * len(quartets) * 256 * padded_cols
    return {"bound": "tensor", "achieved": executed / (kern_ms * 1e-3) / 1e12, "peak": 2.0 * bf16, "unit": "TOP/s",
            "frac": executed / (kern_ms * 1e-3) / 1e12 / (2.0 * bf16), "tiles": tiles, "useful_fraction_of_executed": useful / executed,
            "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (u8 tcgen05.mma runs at twice the bf16 rate)" if peaks_path.exists()
                           else "2 x fallback of B200_PROFILING.md",
            "note": "the launch's duration also covers the planner and the FP64 scores that run beside the GEMM on a second stream"}


def bench_static(config, world, rank, local, dev, lib, steps, warmup, subsample=True, e2e=True):
    """Replicate 0 of an exhaustive config resolved `steps` times against the resident packed matrix; quartet ranges
    sharded over the ranks, results gathered to rank 0 inside the step."""
    import torch
    import tetrad_b200
    from tetrad_b200 import sharding
    db = synth.config_arrays(config, seed=1234)
    tmparr, tmpmap = np.ascontiguousarray(db["tmparr"]), np.ascontiguousarray(db["tmpmap"])
    quartets = replicate0_quartets(config)
    Q = len(quartets)
    eng = tetrad_b200.QuartetResolver(device=local, seed=0).load(tmparr, tmpmap)
    qd = torch.from_numpy(quartets.view(np.int32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    lo, hi = sharding.shard_range(Q, rank, world)

    def step():
        flush.fill_(rank + 1)                                   # L2 flush, outside the events
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        out = {}

        def fn(q):
            res = eng.resolve(q, subsample)
            ev[1].record(stream)
            out["local"] = res
            return res
        out["all"] = sharding.resolve_sharded(qd, fn)
        ev[2].record(stream)
        return ev, out

    for _ in range(max(3, warmup)):
        step()
    barrier(world, dev)
    launches0 = lib.tq_launch_count()
    events = []
    for _ in range(steps):
        ev, out = step()
        events.append(ev)
    barrier(world, dev)
    launches = lib.tq_launch_count() - launches0
    total_ms = max_over_ranks(sum(e[0].elapsed_time(e[2]) for e in events), world, dev)
    kern_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in events)
    res = dict(Q=Q, value=Q * steps / (total_ms * 1e-3), ms_per_step=total_ms / steps, kern_ms=kern_ms,
               kernel=lib.tq_last_kernel().decode(),
               quartets_per_launch=hi - lo, launches=int(launches), ncols=int(tmparr.shape[1]), info=dict(eng.info),
               tmparr=tmparr, tmpmap=tmpmap, quartets=quartets)
    if subsample:
        ns = min(hi - lo, 1 << 20)                              # the u32 diagnostics would wrap on longer lists
        res["stacking"] = stacking_stats(lib, eng, lambda: eng.resolve(qd[lo:lo + ns], subsample), ns)
    if e2e:
        # end to end through the reference-facing host API (pinned host buffers; plan + H2D + pack + resolve + D2H);
        # under N ranks every rank resolves its quartet range of the same matrix
        pin_arr = torch.from_numpy(tmparr).pin_memory()
        pin_map = torch.from_numpy(tmpmap.view(np.int32)).pin_memory()
        pin_q = torch.from_numpy(np.ascontiguousarray(quartets[lo:hi]).view(np.int32)).pin_memory()
        h_arr, h_map, h_q = pin_arr.numpy(), pin_map.numpy().view(np.uint32), pin_q.numpy().view(np.uint32)
        for _ in range(2):
            tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, subsample, device=local)
        barrier(world, dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            _, h_rstat, h_rscor = tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, subsample, device=local)
        dt = max_over_ranks(time.perf_counter() - t0, world, dev)
        h2d, d2h = tetrad_b200.resolve_quartets.last_traffic(local)
        rstat, rscor = out["local"]                             # the two entries must agree bit for bit
        assert (rstat.cpu().numpy().view(np.uint32) == h_rstat).all() and (rscor.cpu().numpy() == h_rscor).all()
        res["e2e"] = {"value": Q * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                      "api": "tetrad_b200.new_infer_resolved_quartets (tq_resolve_host), pinned host arrays", "steps": steps}
    eng.close()
    return res


def bench_replicates(config, world, rank, local, dev, lib, steps, warmup, sampler="device"):
    """Bootstrap replicates of `config` through ReplicatePipeline (strong scaling)."""
    import torch
    from tetrad_b200.replicate import ReplicatePipeline
    cfg = synth.CONFIGS[config]
    db = synth.config_arrays(config, seed=1234)
    arrays = {k: db[k] for k in ("seqarr", "spans", "tmparr", "tmpmap")}
    n, Q = cfg["ntips"], cfg["nquartets"]
    pipe = ReplicatePipeline(arrays, n, Q, comb(n, 4), True, device=local, seed=0, timing=True, sampler=sampler)
    stream = torch.cuda.current_stream(dev)
    rng = np.random.default_rng(42)                      # every rank replays the same generator

    def run(k, first_idx, fetch, draws_ready=None):
        """k replicates, pipelined as run_inference does (replicate r + 1 is prepared and launched before the results
        of replicate r are read); returns the Prepared records (events inside) and the last fetched results."""
        def get_draw(i):
            return draws_ready[i] if draws_ready else pipe.draw(first_idx + i, rng)
        preps, got = [], None
        prep = pipe.prepare(get_draw(0))
        out = pipe.resolve(prep)
        for i in range(k):
            nprep = nout = None
            if i + 1 < k:
                nprep = pipe.prepare(get_draw(i + 1))
                nout = pipe.resolve(nprep)
            if fetch:
                got = pipe.fetch(prep, out)
            preps.append(prep)
            prep, out = nprep, nout
        return preps, got

    run(max(3, warmup), 1, fetch=False)
    # ---- value: device time of K replicates, draws taken before the region --------------------------------------
    ready = [pipe.draw(1000 + i, rng) for i in range(steps)]
    barrier(world, dev)
    clock_sampler = ClockSampler(local)
    clock_sampler.start()
    launches0 = lib.tq_launch_count()
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    preps, _ = run(steps, 1000, fetch=False, draws_ready=ready)
    e1.record(stream)
    barrier(world, dev)
    wall = time.perf_counter() - wall0
    launches = lib.tq_launch_count() - launches0
    clocks = clock_sampler.result()
    total_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    ev = [p.events for p in preps]
    prepare_ms = statistics.mean(e["prepare_start"].elapsed_time(e["prepare_end"]) for e in ev[1:]) if steps > 1 else \
        ev[0]["prepare_start"].elapsed_time(ev[0]["prepare_end"])
    kern_ms = statistics.mean(e["resolve_start"].elapsed_time(e["resolve_end"]) for e in ev)
    gather_ms = statistics.mean(e["resolve_end"].elapsed_time(e["gather_end"]) for e in ev)
    info = dict(preps[-1].engine.info)
    ns = min(-(-Q // world), 1 << 20)                           # the u32 diagnostics would wrap on longer lists
    stacking = stacking_stats(lib, preps[-1].engine, lambda: preps[-1].engine.resolve(preps[-1].q_dev[:ns], True), ns)
    # ---- e2e: the same loop by wall clock, host draws + H2D + D2H of the results on rank 0 inside ----------------
    barrier(world, dev)
    t0 = time.perf_counter()
    _, got = run(steps, 2000, fetch=True)
    dt = max_over_ranks(time.perf_counter() - t0, world, dev)
    if rank == 0:
        quartets, rstat, rscor = got
        assert quartets.shape == (Q, 4) and rstat.shape == (Q, 2) and rscor.shape == (Q, 3) and np.isfinite(rscor).all()
    lpad = info["ntiles"] * 1024
    h2d = (4 * lpad + 2 * info["ntiles"] + 2) * 4 + cfg["nloci"] * 4 + (Q * 8 if sampler == "numpy" else 0)   # plan tables, drawn loci (, quartet ranks)
    d2h = Q * 32 + Q * 16                                                          # results + quartet ids, rank 0
    res = dict(Q=Q, value=Q * steps / (total_ms * 1e-3), ms_per_step=total_ms / steps, kern_ms=kern_ms,
               quartets_per_launch=-(-Q // world), launches=int(launches), ncols=int(info["ncols"]), info=info,
               clocks=clocks, wall=wall, kernel=lib.tq_last_kernel().decode(), sampler_name=sampler, stacking=stacking,
               replicate={"prepare_ms": prepare_ms, "resolve_ms": kern_ms, "gather_ms": gather_ms,
                          "note": "device time of one rank's share of a replicate (CUDA events, rank 0): bootstrap column "
                                  "orders + planes of this rank's tip rows + all-gather + unranking | resolve of Q/N "
                                  "quartets | NCCL gather to rank 0"},
               e2e={"value": Q * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h,
                    "api": "tetrad_b200.replicate.ReplicatePipeline draw/prepare/resolve/fetch (the loop run_inference "
                           "drives), wall clock", "steps": steps},
               tmparr=db["tmparr"], tmpmap=db["tmpmap"])
    pipe.close()
    return res


def run_ours(args):
    import torch.distributed as dist
    from tetrad_b200 import _lib
    world, rank, local, dev = setup_distributed()
    lib = _lib.load()
    cfg = synth.CONFIGS[args.config]
    replicate_mode = cfg["nboots"] > 0
    sampler = None
    if replicate_mode:
        res = bench_replicates(args.config, world, rank, local, dev, lib, args.steps, args.warmup, sampler=args.sampler)
        clocks, wall = res["clocks"], res["wall"]
    else:
        sampler = ClockSampler(local)
        sampler.start()
        wall0 = time.perf_counter()
        res = bench_static(args.config, world, rank, local, dev, lib, args.steps, args.warmup)
        wall = time.perf_counter() - wall0
        clocks = sampler.result()
    extra = {}
    if world == 1 and args.config == DEFAULT_CONFIG and not args.quick:
        c2 = bench_static("C2", world, rank, local, dev, lib, 10, 3)
        roof, bind = roofline_block("C2", c2["ncols"], c2["quartets_per_launch"], c2["kern_ms"], c2["info"], lib, local,
                                    kernel=c2["kernel"], stacking=c2.get("stacking"))
        extra["c2"] = {"config": describe("C2"), "value": c2["value"], "ms_per_step": c2["ms_per_step"], "e2e": c2["e2e"],
                       "roofline": roof, "binding": bind,
                       "note": "BASELINE configs[1] (the 1-B200 shape): replicate 0 re-resolved against the resident matrix, "
                               "L2 flushed between steps"}
        fm = bench_static("C2", world, rank, local, dev, lib, 5, 3, subsample=False, e2e=False)
        roof, bind = roofline_block("C2", fm["ncols"], fm["quartets_per_launch"], fm["kern_ms"], fm["info"], lib, local,
                                    subsample=False, kernel=fm["kernel"])
        if "pair_gemm" in fm["kernel"]:
            roof["tensor"] = tensor_block(fm["quartets"], synth.CONFIGS["C2"]["ntips"], fm["info"]["padded_cols"], fm["kern_ms"])
        extra["full_mode"] = {"config": describe("C2", subsample=False), "value": fm["value"], "ms_per_step": fm["ms_per_step"],
                              "roofline": roof, "binding": bind,
                              "note": "subsample_snps=False (full_chunk_to_matrices, the CLI's default mode) at the C2 shape"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roofline, binding = roofline_block(args.config, res["ncols"], res["quartets_per_launch"], res["kern_ms"], res["info"],
                                       lib, local, kernel=res["kernel"], stacking=res.get("stacking"))
    cpu = None
    if world == 1 and not args.no_cpu:
        quartets = res.get("quartets")
        if quartets is None:
            quartets = replicate0_quartets(args.config)
        probe, cores, _, _ = cpu_port_rate(res["tmparr"], res["tmpmap"], quartets, sample=16 * (os.cpu_count() or 1))
        sample = int(max(4 * cores, min(len(quartets), probe * 15.0)))
        rate, cores, nq, dt = cpu_port_rate(res["tmparr"], res["tmpmap"], quartets, sample)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nq} random quartets of the replicate-0 quartet set in {dt:.1f} s, fork pool of {cores} workers, "
                         "OPENBLAS_NUM_THREADS=1", "port_vs_reference": port_calibration()}
    planes_mb = res["info"]["row_words"] * 4.0 * cfg["ntips"] / 1e6
    seq_mb = cfg["ntips"] * float(res["info"]["ncols"]) / 1e6
    setup = {"l2": (f"inputs larger than L2: every replicate is a different bootstrap draw, packed on the device from the resident "
                    f"{seq_mb:.0f} MB seqarr into a new {planes_mb:.0f} MB plane matrix (two engines alternate, so two of them are live)")
                   if replicate_mode else "flushed between steps (256 MiB write outside the per-step CUDA events)",
             "parallelism": f"quartet ranges x{world} (contiguous shard per rank), NCCL gather to rank 0" +
                            (", bootstrap planes packed by tip-row shares + NCCL all-gather" if replicate_mode and world > 1 else ""),
             "quartets_per_step": res["Q"]}
    if replicate_mode:
        setup["quartet_sampler"] = res["sampler_name"] + (" (keyed permutation on the device, one 63-bit draw per replicate)"
                                                           if res["sampler_name"] == "device" else " (the reference's rng.choice on the host)")
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 x u8 -> s32 tensor-core counts (bit-exact integers) + f64 svd", "data": "synthetic",
        "config": describe(args.config), "setup": setup,
        "e2e": res["e2e"], "gpu_launches": res["launches"], "roofline": roofline, "binding": binding,
        "replicate": res.get("replicate"), "cpu_baseline": cpu, "clocks": clocks, "wall_s_timed_region": wall,
    }
    line.update(extra)
    print_result(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # Only the JSON line may reach stdout (NCCL and torchrun print banners there): route fd 1 to
    # stderr for the whole run and write the result to the saved descriptor at the end.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    global print_result
    print_result = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sampler", default="device", choices=["device", "numpy"],
                    help="random quartet samples of the replicate loop: device (default of the CLI) or the reference's host draw")
    ap.add_argument("--quick", action="store_true", help="skip the extra c2 / full_mode blocks of the default config")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
