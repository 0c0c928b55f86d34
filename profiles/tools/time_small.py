"""Which resolve kernel wins at which launch size?  C1 matrix (all 495 quartets) and C2 matrix (random subsets).
usage: python profiles/tools/time_small.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
from itertools import combinations
import tetrad_b200
from tetrad_b200 import _lib, synth

def bench(eng, qd, reps=6):
    best = 1e9
    for rep in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.resolve(qd, True); e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    return best

for cfg, sizes in (("C1", [495]), ("C2", [500, 1184, 2368, 4736, 9472, 18944, 37888])):
    db = synth.config_arrays(cfg)
    n = db["tmparr"].shape[0]
    allq = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
    eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    rng = np.random.default_rng(1)
    for m in sizes:
        q = allq if m >= len(allq) else allq[np.sort(rng.choice(len(allq), m, replace=False))]
        qd = torch.from_numpy(q.view(np.int32)).cuda()
        row = []
        for k in ("fused", "ws", "mma"):
            _lib.set_kernel(k)
            ms = bench(eng, qd)
            row.append("%s %.3f ms (%.3g q/s)" % (k, ms, len(q) / ms * 1e3))
        _lib.set_kernel("auto")
        print(cfg, len(q), " | ".join(row), flush=True)
