import sys, time, ctypes
from pathlib import Path; R = str(Path(__file__).resolve().parent.parent.parent); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import numpy as np, torch
import tetrad_b200
from tetrad_b200 import _lib, synth
from conftest import load_golden
lib = _lib.load()
for kind, name in ((0, "popc"), (1, "lop3"), (2, "dfma"), (3, "copy")):
    r = ctypes.c_double(); rc = lib.tq_microbench(0, kind, ctypes.byref(r)); print(name, rc, "%.4g" % r.value)
g = load_golden("noisy12")
q, rstat, rscor, ex = tetrad_b200.new_infer_resolved_quartets(g["tmparr"], g["tmpmap"], g["quartets"], True, return_counts=True)
print("counts exact:", (ex["counts"] == g["sub_mats0"]).all(), "score maxrel", np.abs(rscor - g["sub_rscor"]).max())
print("best equal:", (rstat[:, 0] == g["sub_rstat"][:, 0]).mean())
db = synth.config_arrays("C2")
from itertools import combinations
quartets = np.array(list(combinations(range(50), 4)), dtype=np.uint32)
eng = tetrad_b200.QuartetResolver(device=0); t = time.time(); eng.load(db["tmparr"], db["tmpmap"]); torch.cuda.synchronize(); print("load", time.time() - t, eng.info)
qd = torch.from_numpy(quartets.view(np.int32)).cuda()
for sub in (True, False):
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); rstat, rscor = eng.resolve(qd, sub); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); print("sub" if sub else "full", "C2 resolve ms", ms, "q/s %.4g" % (len(quartets) / ms * 1e3))
