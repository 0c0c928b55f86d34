"""Where does 12 + 4 overtake 8 + 8?  Times both warp mixes of the tensor-core kernel on synthetic shapes between C2 (49 tiles)
and C3 (98 tiles).  usage: python profiles/tools/time_mix_threshold.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, torch
import tetrad_b200
from tetrad_b200 import _lib, synth, combinations as cb

lib = _lib.load()
_lib.set_kernel("mma")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ntips, nloci, nsnps in ((60, 61_000, 250_000), (70, 72_000, 300_000), (80, 84_000, 360_000)):
    d = synth.generate(ntips, nloci, nsnps, seed=7)
    db = synth.prepare_database(d.snps, d.snpsmap, seed=7)
    eng = tetrad_b200.QuartetResolver(device=0).load(db["tmparr"], db["tmpmap"])
    q = cb.device_sample(ntips, 7, 300_000)
    out = []
    for nc in (8, 12, 0):
        _lib.check(lib.tq_set_option(b"count_warps", nc))
        ts = []
        for i in range(4):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.resolve(q, True); e1.record(); torch.cuda.synchronize()
            if i: ts.append(e0.elapsed_time(e1))
        out.append(f"{nc or 'auto'}:{lib.tq_last_count_warps()}+{16 - lib.tq_last_count_warps()} {np.median(ts):.2f} ms")
    print(f"{ntips} tips, {eng.info['ntiles']} tiles, {eng.info['row_words'] // 128} chunks: " + "  ".join(out), flush=True)
    eng.close()
_lib.check(lib.tq_set_option(b"count_warps", 0))
