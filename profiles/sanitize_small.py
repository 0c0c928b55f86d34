"""Small end-to-end workload for compute-sanitizer: host entry (plan, pack, resolve both modes),
device bootstrap, device unranking.  usage: compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import tetrad_b200
from tetrad_b200 import combinations as cb, run_inference, synth

d = synth.generate(8, 2200, 7000, seed=5, p_ambig=0.02)
db = synth.prepare_database(d.snps, d.snpsmap, seed=5)
q = cb.device_quartets(8)
eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
for sub in (True, False):
    rstat, rscor, counts, minrank = eng.resolve(q, sub, want_counts=True)
lidxs, a, b = run_inference.draw_bootstrap(db["spans"], np.random.default_rng(1))
eng.load_bootstrap(db["seqarr"], db["spans"], lidxs, a, b)
rstat, rscor = eng.resolve(q, True)
_, hstat, hscor = tetrad_b200.new_infer_resolved_quartets(db["tmparr"], db["tmpmap"], q.cpu().numpy().view(np.uint32)[:20], True)
torch.cuda.synchronize()
print("ok", int(rstat.sum().item()), float(rscor.sum().item()), int(hstat.sum()))
