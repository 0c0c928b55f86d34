"""Per-call wall time of the host entry new_infer_resolved_quartets (tq_resolve_host) at a BASELINE shape.
usage: python profiles/tools/time_e2e.py [C2|C3]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import tetrad_b200
from tetrad_b200 import synth
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
tmparr, tmpmap, quartets, meta = bench.workload(cfg, 0)
pin_arr = torch.from_numpy(tmparr).pin_memory(); pin_map = torch.from_numpy(tmpmap.view(np.int32)).pin_memory()
pin_q = torch.from_numpy(quartets.view(np.int32)).pin_memory()
h_arr, h_map, h_q = pin_arr.numpy(), pin_map.numpy().view(np.uint32), pin_q.numpy().view(np.uint32)
for i in range(8):
    t0 = time.perf_counter()
    tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, True, device=0)
    print("call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
