"""Does the processing order of a random quartet list matter?  Same quartets, as drawn vs lexicographically sorted.
usage: python profiles/tools/time_sorted.py [C3|C4]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import tetrad_b200, bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
tmparr, tmpmap, quartets, meta = bench.workload(cfg, 0)
eng = tetrad_b200.QuartetResolver(device=0).load(tmparr, tmpmap)
order = np.lexsort((quartets[:, 3], quartets[:, 2], quartets[:, 1], quartets[:, 0]))
for name, q in (("as drawn", quartets), ("sorted", quartets[order])):
    qd = torch.from_numpy(np.ascontiguousarray(q).view(np.int32)).cuda()
    best = 1e9
    for rep in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); rstat, rscor = eng.resolve(qd, True); e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    print("%s %-9s %9.3f ms  %.4g quartets/s" % (cfg, name, best, len(q) / best * 1e3), flush=True)
