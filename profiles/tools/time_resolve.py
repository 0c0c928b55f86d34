"""Time one resolve launch of a BASELINE shape on the device (CUDA events, best of 3 after a warm-up).
usage: [TQ_LIB=path/to/variant.so] [TQ_KERNEL=mma|ws|fused] python profiles/tools/time_resolve.py [C2|C3] [sub|full|both]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from tetrad_b200 import _lib
if os.environ.get("TQ_LIB"):
    _lib.LIB_PATH = Path(os.environ["TQ_LIB"]).resolve()
import tetrad_b200
from tetrad_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
modes = {"sub": (True,), "full": (False,), "both": (True, False)}[sys.argv[2] if len(sys.argv) > 2 else "sub"]
db = synth.config_arrays(cfg)
n = db["tmparr"].shape[0]
if cfg == "C2":
    from itertools import combinations
    quartets = np.array(list(combinations(range(n), 4)), dtype=np.uint32)
else:
    rng = np.random.default_rng(42)
    quartets = np.sort(np.stack([rng.permutation(n)[:4] for _ in range(300000)]), axis=1).astype(np.uint32)
eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
qd = torch.from_numpy(quartets.view(np.int32)).cuda()
for sub in modes:
    best = 1e9
    for rep in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); rstat, rscor = eng.resolve(qd, sub); e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    print("%-28s %s %-4s %8.3f ms  %.4g quartets/s  (checksum %d %.6f)" % (
        os.path.basename(os.environ.get("TQ_LIB", "default")), cfg, "sub" if sub else "full", best, len(quartets) / best * 1e3,
        int(rstat[:, 0].sum()), float(rscor.sum())), flush=True)
