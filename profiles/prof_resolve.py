"""Small fixed workload for ncu: C2-shaped matrix, a few quartets per resident warp.
usage: python profiles/prof_resolve.py [nquartets] [subsample 0/1]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import tetrad_b200
from tetrad_b200 import synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8 * 8
sub = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
db = synth.config_arrays("C2")
rng = np.random.default_rng(0)
quartets = np.sort(np.stack([rng.choice(50, 4, replace=False) for _ in range(nq)]), axis=1).astype(np.uint32)
eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
qd = torch.from_numpy(quartets.view(np.int32)).cuda()
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); rstat, rscor = eng.resolve(qd, sub); e1.record(); torch.cuda.synchronize()
    print("resolve ms", e0.elapsed_time(e1), "q/s %.4g" % (nq / e0.elapsed_time(e1) * 1e3))
print("checksum", int(rstat.sum().item()), float(rscor.sum().item()))
