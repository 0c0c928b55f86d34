#!/usr/bin/env python
"""Calibrate bench.py's CPU arm: the oracle port against the EXECUTED reference, one core each.

Run in the build container only (it imports the reference's numba path from /root/reference, which does not exist on
the GPU box): `python profiles/tools/calibrate_cpu_port.py > profiles/cpu_port_vs_reference.json`.

bench.py's `--impl reference` arm and `cpu_baseline` leg time the oracle port (oracle/: C count loops + the same
numpy.linalg svd / matrix_rank calls) because the reference cannot travel.  This script times the reference's own
`new_infer_resolved_quartets` (tetrad/src/resolve_quartets.py:191-265; numba kernels JIT-warmed first) and the port on
the same inputs -- C1 in full (all 495 quartets, both modes) and a sample of C2 -- and records quartets/s per core and
the factor port / reference, so that the CPU numbers in the bench line can be read as reference numbers.
"""
import json
import os
import sys
import time
import types
from pathlib import Path

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
from tetrad.src import resolve_quartets as ref    # noqa: E402

from oracle import oracle                         # noqa: E402
from tetrad_b200 import combinations as cb, synth  # noqa: E402


def rate(fn, *args, repeat=1):
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        fn(*args)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def main():
    oracle.build()
    out = {"numpy": np.__version__, "host": os.uname().nodename, "cores_used": 1, "cases": []}
    import numba
    out["numba"] = numba.__version__
    for name, nq in (("C1", 495), ("C2", 400)):
        db = synth.config_arrays(name)
        n = synth.CONFIGS[name]["ntips"]
        quartets = cb.all_quartets(n)
        if nq < len(quartets):
            quartets = quartets[np.linspace(0, len(quartets) - 1, nq).astype(np.int64)]
        for sub in (True, False):
            ref.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets[:4], sub)        # numba JIT
            oracle.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], quartets[:4], sub)
            t_ref = rate(ref.new_infer_resolved_quartets, db["tmparr"], db["tmpmap"], quartets, sub, repeat=2)
            t_port = rate(oracle.new_infer_resolved_quartets, db["tmparr"], db["tmpmap"], quartets, sub, repeat=2)
            out["cases"].append({"config": name, "quartets": int(len(quartets)), "subsample_snps": sub,
                                 "reference_q_per_s_per_core": len(quartets) / t_ref,
                                 "port_q_per_s_per_core": len(quartets) / t_port,
                                 "port_over_reference": t_ref / t_port})
    sub_cases = [c for c in out["cases"] if c["subsample_snps"]]
    out["port_over_reference_subsample"] = float(np.mean([c["port_over_reference"] for c in sub_cases]))
    out["port_over_reference_full"] = float(np.mean([c["port_over_reference"] for c in out["cases"] if not c["subsample_snps"]]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
