#!/usr/bin/env python
"""Benchmark of the quartet-resolution hot path (BASELINE.json metric: quartets resolved/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

One "step" = one replicate: every quartet of the workload resolved against one packed SNP matrix
(subsample mode: one SNP per locus -> 256 site-pattern counts -> 3 SVD scores -> best split).

* default arm (ours): `value` = quartets/s with the packed matrix and the quartet list already in
  HBM (timed per step with CUDA events on the launching stream; L2 flushed between steps);
  `e2e` = the same metric through the reference-facing call
  `tetrad_b200.new_infer_resolved_quartets(tmparr, tmpmap, quartets, subsample)` with pinned HOST
  buffers: locus plan + H2D + pack + resolve + D2H inside the timed region.
* N > 1 (torchrun): weak scaling -- every rank resolves one full replicate of its own
  (bootstrap-resampled) matrix and the 32 B/quartet results are gathered to rank 0 with NCCL
  inside the timed step; `value` = N * Q / max-over-ranks time.
* `--impl reference`: the reference's CPU implementation of the path.  The reference is pure
  Python (numba/numpy) and /root/reference does not exist on the GPU box, so this arm times the
  oracle port (oracle/: C count loops + the same numpy.linalg svd/matrix_rank calls) on all host
  cores over a bounded random sample of the same workload per step.
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
from itertools import combinations
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from tetrad_b200 import synth  # noqa: E402

print_result = print
METRIC = "quartets_resolved_per_sec"
UNIT = "quartets/s"


# ----------------------------------------------------------------------------------------------
def workload(name: str, rank: int = 0):
    """(tmparr, tmpmap, quartets, meta) of BASELINE config `name` for replicate `rank`."""
    cfg = synth.CONFIGS[name]
    db = synth.config_arrays(name, seed=1234)
    tmparr, tmpmap = db["tmparr"], db["tmpmap"]
    if rank > 0:  # every other rank resolves a bootstrap replicate of the same data set
        tmparr, tmpmap = synth.bootstrap_resample(db["seqarr"], db["spans"], np.random.default_rng(42 + rank))
    n = cfg["ntips"]
    if cfg["nquartets"] == 0:
        quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
    else:
        from tetrad_b200.combinations import random_quartets
        quartets = random_quartets(n, cfg["nquartets"], np.random.default_rng(42 + rank))
    meta = dict(cfg, name=name, ncols=int(tmparr.shape[1]))
    return np.ascontiguousarray(tmparr), np.ascontiguousarray(tmpmap), quartets, meta


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
def cpu_port_rate(tmparr, tmpmap, quartets, sample: int, seed: int = 0):
    """quartets/s of the oracle port on all host cores over a random sample of the workload."""
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(quartets), size=min(sample, len(quartets)), replace=False))
    sub = np.ascontiguousarray(quartets[pick])
    oracle.resolve_parallel(tmparr, tmpmap, sub[: 4 * cores], True, nproc=cores)   # pool + BLAS warm-up
    t0 = time.perf_counter()
    oracle.resolve_parallel(tmparr, tmpmap, sub, True, nproc=cores)
    dt = time.perf_counter() - t0
    return len(sub) / dt, cores, len(sub), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tmparr, tmpmap, quartets, meta = workload(args.config)
    # size the per-step sample so the whole run stays within a few minutes
    probe_rate, cores, _, _ = cpu_port_rate(tmparr, tmpmap, quartets, sample=64 * (os.cpu_count() or 1))
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(max(256, min(len(quartets), probe_rate * min(20.0, budget_s))))
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 counts + f64 svd (numpy / LAPACK)",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {meta['ntips']} tips, {meta['nloci']} loci, {meta['ncols']} SNPs, "
                               f"{len(quartets)} quartets, subsample_snps=True, 0 bootstraps",
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} random quartets of the workload per step, fork pool of {cores} "
                                   "workers, OPENBLAS_NUM_THREADS=1 (reference is pure Python; oracle port timed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_result(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tetrad_b200
    from tetrad_b200 import _lib, sharding

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
    lib = _lib.load()

    tmparr, tmpmap, quartets, meta = workload(args.config, rank)
    Q = len(quartets)
    eng = tetrad_b200.QuartetResolver(device=local, seed=0)
    eng.load(tmparr, tmpmap)
    qd = torch.from_numpy(quartets.view(np.int32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gather_buf = [torch.empty((Q, 4), dtype=torch.float64, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None
    stream = torch.cuda.current_stream(dev)

    def step(timed):
        """resolve one replicate (+ gather to rank 0); returns (step event pair, kernel event pair)"""
        flush.fill_(rank + 1)                                   # L2 flush, outside the events
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        ev[2].record(stream)
        rstat, rscor = eng.resolve(qd, True)
        ev[3].record(stream)
        if world > 1:
            packed = sharding.pack_results(rstat, rscor)                     # 32 B / quartet
            dist.gather(packed, gather_buf, dst=0)
        ev[1].record(stream)
        return ev, (rstat, rscor)

    for _ in range(max(3, args.warmup)):
        step(False)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.tq_launch_count()
    wall0 = time.perf_counter()
    events = []
    for _ in range(args.steps):
        ev, out = step(True)
        events.append(ev)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    launches = lib.tq_launch_count() - launches0
    clocks = sampler.result()
    step_ms = [e[0].elapsed_time(e[1]) for e in events]
    kern_ms = [e[2].elapsed_time(e[3]) for e in events]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * Q * args.steps / (total_ms * 1e-3)

    # ---- end to end through the reference-facing host API (pinned host buffers) ---------------
    pin_arr = torch.from_numpy(tmparr).pin_memory()
    pin_map = torch.from_numpy(tmpmap.view(np.int32)).pin_memory()
    pin_q = torch.from_numpy(quartets.view(np.int32)).pin_memory()
    h_arr, h_map, h_q = pin_arr.numpy(), pin_map.numpy().view(np.uint32), pin_q.numpy().view(np.uint32)
    e2e_steps = max(1, args.steps)                     # the same K steps as the device-timed loop
    for _ in range(2):
        tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, True, device=local)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, h_rstat, h_rscor = tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, True, device=local)
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * Q * e2e_steps / float(e2e_dt.item())
    h2d, d2h = tetrad_b200.resolve_quartets.last_traffic(local)
    # the two paths must agree bit for bit
    rstat, rscor = out
    assert (rstat.cpu().numpy().view(np.uint32) == h_rstat).all() and (rscor.cpu().numpy() == h_rscor).all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (resolve_kernel) -------------------------------------
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak_gbs, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak_gbs, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    kern_s = statistics.mean(kern_ms) * 1e-3
    alg_bytes = algorithmic_bytes_per_quartet(meta["ncols"]) * Q
    achieved = alg_bytes / kern_s / 1e9
    popc = ctypes.c_double()
    _lib.check(lib.tq_microbench(local, 0, ctypes.byref(popc)))
    popc_per_quartet = 256.0 * 32.0 * eng.info["ntiles"]          # merged-plane algorithm: 256 POPC / 32 loci
    traffic = None
    prof = ROOT / "profiles" / "resolve_traffic.json"
    if prof.exists():
        traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
    kernel_name = lib.tq_last_kernel().decode()                 # which variant launch_resolve chose (tq_kernels.cu)
    imma = ctypes.c_double()
    _lib.check(lib.tq_microbench(local, 4, ctypes.byref(imma)))
    lop3 = ctypes.c_double()
    _lib.check(lib.tq_microbench(local, 1, ctypes.byref(lop3)))
    imma_per_quartet = 64.0 * eng.info["ntiles"]                  # 16 x 16 x 1024 bit-products per tile / (16 x 8 x 32 per IMMA)
    alu_per_quartet = 32.0 * (320.0 * eng.info["ntiles"] + 36.0 * eng.info["row_words"] / 128.0)   # unpack + stage-1 logic, thread-ops
    roofline = {
        "bound": "hbm", "kernel": kernel_name,
        "achieved": achieved, "peak": peak_gbs,
        "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_quartet": algorithmic_bytes_per_quartet(meta["ncols"]),
        "kernel_ms": statistics.mean(kern_ms),
        "popc": {"achieved_popc_per_s": popc_per_quartet * Q / kern_s, "peak_popc_per_s": popc.value,
                 "frac": popc_per_quartet * Q / kern_s / popc.value,
                 "note": "rate a 256-POPC-per-32-loci implementation would need; the tensor-core count path issues no "
                         "POPC, so frac > 1 is possible; peak measured by tq_microbench on this GPU"},
        "imma": {"achieved_per_s": imma_per_quartet * Q / kern_s, "peak_per_s": imma.value,
                 "frac": imma_per_quartet * Q / kern_s / imma.value,
                 "note": "warp-level IMMA.16832.U8.U8 (mma.sync m16n8k32 u8) of stage 2; peak measured by tq_microbench kind 4"},
        "alu": {"achieved_per_s": alu_per_quartet * Q / kern_s, "peak_per_s": 2.0 * lop3.value,
                "frac": alu_per_quartet * Q / kern_s / (2.0 * lop3.value),
                "note": "LOP3/SHF thread-ops of the bit-plane unpack (320 per tile and lane) and of stage 1 (36 per rank "
                        "chunk and lane), the pipe that binds the count warps; peak measured by tq_microbench kind 1 (LOP3 + SHF pairs, counted as two ops)"},
    }

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample ------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        probe, cores, _, _ = cpu_port_rate(tmparr, tmpmap, quartets, sample=64 * (os.cpu_count() or 1))
        sample = int(max(512, min(Q, probe * 15.0)))
        rate, cores, n, dt = cpu_port_rate(tmparr, tmpmap, quartets, sample)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} random quartets of the same workload in {dt:.1f} s, fork pool of {cores} workers, "
                         "OPENBLAS_NUM_THREADS=1"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 tensor-core counts (bit-exact integers) + f64 svd", "data": "synthetic",
        "config": {"workload": f"{args.config}: {meta['ntips']} tips, {meta['nloci']} loci, {meta['ncols']} SNPs, "
                               f"{Q} quartets per replicate, subsample_snps=True, one replicate per GPU per step",
                   "l2": "flushed between steps (256 MiB write outside the per-step CUDA events)",
                   "quartets_per_step": world * Q, "parallelism": f"replicates x{world}, NCCL gather to rank 0"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "tetrad_b200.new_infer_resolved_quartets (tq_resolve_host)", "steps": e2e_steps},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "wall_s_timed_region": wall,
    }
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
    ap.add_argument("--config", default="C2", choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
