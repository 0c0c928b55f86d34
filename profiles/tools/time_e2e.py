"""Per-call wall time of the host entry new_infer_resolved_quartets (tq_resolve_host) at the C2 shape, pinned host arrays as
bench.py's e2e leg uses them.  With TQ_HOST_TIMING=1 a library built with the timing prints reports its phases.
usage: python profiles/tools/time_e2e.py [C2]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import tetrad_b200
import bench
from tetrad_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
db = synth.config_arrays(cfg, seed=1234)
tmparr, tmpmap = np.ascontiguousarray(db["tmparr"]), np.ascontiguousarray(db["tmpmap"])
quartets = bench.replicate0_quartets(cfg)
pin_arr = torch.from_numpy(tmparr).pin_memory(); pin_map = torch.from_numpy(tmpmap.view(np.int32)).pin_memory()
pin_q = torch.from_numpy(quartets.view(np.int32)).pin_memory()
h_arr, h_map, h_q = pin_arr.numpy(), pin_map.numpy().view(np.uint32), pin_q.numpy().view(np.uint32)
for i in range(8):
    t0 = time.perf_counter()
    tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, True, device=0)
    print("host entry call %d: %.2f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
