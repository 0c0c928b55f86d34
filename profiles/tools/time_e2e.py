"""Per-call wall time of the host entry new_infer_resolved_quartets (tq_resolve_host) at a BASELINE shape, and of
its parts through the device-resident engine (plan + H2D + pack, resolve, D2H).
usage: python profiles/tools/time_e2e.py [C2|C3|C4]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import tetrad_b200
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
tmparr, tmpmap, quartets, meta = bench.workload(cfg, 0)
pin_arr = torch.from_numpy(tmparr).pin_memory(); pin_map = torch.from_numpy(tmpmap.view(np.int32)).pin_memory()
pin_q = torch.from_numpy(quartets.view(np.int32)).pin_memory()
h_arr, h_map, h_q = pin_arr.numpy(), pin_map.numpy().view(np.uint32), pin_q.numpy().view(np.uint32)
for i in range(8):
    t0 = time.perf_counter()
    tetrad_b200.new_infer_resolved_quartets(h_arr, h_map, h_q, True, device=0)
    print("host entry call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
eng = tetrad_b200.QuartetResolver(device=0)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.load(h_arr, h_map); torch.cuda.synchronize(); t1 = time.perf_counter()
    qd = pin_q.cuda(non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    rstat, rscor = eng.resolve(qd, True); torch.cuda.synchronize(); t3 = time.perf_counter()
    a = rstat.cpu(); b = rscor.cpu(); t4 = time.perf_counter()
    print("engine %d: load (plan + H2D + pack) %.1f ms, quartets H2D %.1f, resolve %.1f, D2H %.1f" % (
        i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
