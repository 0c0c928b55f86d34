#!/usr/bin/env python
"""Round-2 bring-up helper (GPU box): one phase per process, so that a trapped kernel cannot poison the next phase.

    python profiles/tools/r2_kernel_check.py pairgemm <desc_mode>     small full-mode case through the tip-pair GEMM vs the oracle
    python profiles/tools/r2_kernel_check.py time <config> <kernel> <sub|full> [reps]   one launch shape timed with CUDA events
"""
import sys
import time
from itertools import combinations
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def pairgemm(desc_mode: int):
    import torch
    import tetrad_b200
    from tetrad_b200 import _lib, synth
    from oracle import oracle
    lib = _lib.load()
    pass
    for n, nloci, nsnps in ((6, 50, 200), (14, 4000, 15000)):
        d = synth.generate(n, nloci, nsnps, seed=5)
        db = synth.prepare_database(d.snps, d.snpsmap, seed=5)
        quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
        eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
        _lib.set_kernel("pairgemm")
        rstat, rscor, counts, minrank = eng.resolve(quartets, False, want_counts=True)
        torch.cuda.synchronize()
        _, ostat, oscor, det = oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets, False, return_details=True)
        got = counts.cpu().numpy().view(np.uint32).reshape(-1, 16, 16)
        want = det["mats"][:, 0]
        ok = (got == want).all()
        print(f"n={n}: counts {'MATCH' if ok else 'DIFFER'}; kernel {lib.tq_last_kernel().decode()}")
        if not ok:
            bad = np.argwhere(got != want)
            print("  first mismatches (quartet, ab, cd):", bad[:6].tolist())
            q = bad[0][0]
            print("  got row sums", got[q].sum(1).tolist(), "\n  want row sums", want[q].sum(1).tolist())
            print("  got total", int(got[q].sum()), "want total", int(want[q].sum()))
        else:
            print("  nsnps equal:", bool((rstat.cpu().numpy().view(np.uint32) == ostat).all()),
                  " max score err:", float(np.abs(rscor.cpu().numpy() - oscor).max()))


def timed(config: str, kernel: str, mode: str, reps: int = 5):
    import torch
    import tetrad_b200
    from tetrad_b200 import _lib, synth, combinations as cb
    lib = _lib.load()
    db = synth.config_arrays(config)
    cfg = synth.CONFIGS[config]
    n = cfg["ntips"]
    if cfg["nquartets"] == 0:
        q = cb.device_quartets(n)
    else:
        q = cb.device_sample(n, 7, cfg["nquartets"])
    eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    _lib.set_kernel(kernel)
    sub = mode == "sub"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for i in range(reps + 2):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rstat, rscor = eng.resolve(q, sub)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    print(f"{config} {mode} {kernel}: {ms:.3f} ms  {len(q) / ms / 1e3:.3f} M quartets/s  [{lib.tq_last_kernel().decode()}]  "
          f"checksum {int(rstat.to(torch.int64).sum())} {float(rscor.sum()):.6f}")


def mixes(config: str, mode: str, nq_cap: int = 1_000_000, reps: int = 3):
    """Count / score warp mixes of the tensor-core kernel on one shape (tq_set_option("count_warps"))."""
    import torch
    import tetrad_b200
    from tetrad_b200 import _lib, synth, combinations as cb
    lib = _lib.load()
    db = synth.config_arrays(config)
    cfg = synth.CONFIGS[config]
    n = cfg["ntips"]
    q = cb.device_quartets(n) if cfg["nquartets"] == 0 else cb.device_sample(n, 7, min(cfg["nquartets"], nq_cap))
    eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    _lib.set_kernel("mma")
    sub = mode == "sub"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    import os
    fills = [int(x) for x in os.environ.get("TQ_FILLS", "-1").split(",")]
    ncs = [int(x) for x in os.environ.get("TQ_NCS", "0,8,12").split(",")]
    _lib.check(lib.tq_set_option(b"stack_stats", int(os.environ.get("TQ_STATS", "1"))))
    ovfs = [int(x) for x in os.environ.get("TQ_OVFS", "8").split(",")]
    for nc, fill, ovf in [(a, b, c) for b in fills for c in ovfs for a in ncs]:
        _lib.check(lib.tq_set_option(b"count_warps", nc))
        _lib.check(lib.tq_set_option(b"stack_fill", fill))
        _lib.check(lib.tq_set_option(b"stack_ovf", ovf))
        times = []
        for i in range(reps + 1):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rstat, rscor = eng.resolve(q, sub)
            e1.record()
            torch.cuda.synchronize()
            if i >= 1:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        chk = (int(rstat.to(torch.int64).sum()), float(rscor.sum()))
        st = eng._last_status.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        per = [float(x) / len(q) for x in st[1:]]
        ref = ref or chk
        print(f"{config} {mode} stack_fill={fill} stack_ovf={ovf} count_warps={nc} (ran {lib.tq_last_count_warps()}+{16 - lib.tq_last_count_warps()}): {ms:.3f} ms  "
              f"{len(q) / ms / 1e3:.3f} M quartets/s  same_results={chk == ref}  per quartet: products {per[0]:.1f} stacked {per[1]:.1f} walk {per[2]:.1f}", flush=True)
    _lib.check(lib.tq_set_option(b"count_warps", 0))
    _lib.check(lib.tq_set_option(b"stack_fill", -1))
    _lib.check(lib.tq_set_option(b"stack_ovf", 8))


if __name__ == "__main__":
    if sys.argv[1] == "pairgemm":
        pairgemm(int(sys.argv[2]))
    elif sys.argv[1] == "mixes":
        mixes(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000)
    else:
        timed(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 5)
