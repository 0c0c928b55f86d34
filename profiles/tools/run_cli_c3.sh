#!/bin/bash
# The replicate loop end to end at the C3 shape through the CLI (DESIGN.md "The replicate loop at scale"):
# writes a synthetic c3.snps.hdf5, then `init` and `run -b 3` with the timing log on stderr.
set -e
W=${1:-/tmp/c3run}; mkdir -p $W
python - <<PY
import sys, time
sys.path.insert(0, ".")
from tetrad_b200 import synth, database
from pathlib import Path
c = synth.CONFIGS["C3"]
t = time.time(); d = synth.generate(c["ntips"], c["nloci"], c["nsnps"], seed=1234)
database.write_snps_input(Path("$W/c3.snps.hdf5"), d.snps, d.snpsmap, d.names); print("input written in %.1f s" % (time.time() - t))
PY
python -m tetrad_b200 init $W/c3.snps.hdf5 -n c3 -w $W -q 1000000 -r 42 -s 2>&1 | tail -5
python -m tetrad_b200 run $W/c3.json -b 3 2>&1 | tail -30
ls -la $W | head -20
