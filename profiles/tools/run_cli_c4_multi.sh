#!/bin/bash
# BASELINE's target workload shape through the CLI under torchrun: 200 tips, 200 k loci, 1 M SNPs, 1 M random quartets per
# replicate, NBOOTS bootstraps, quartet ranges sharded over the GPUs.  usage: run_cli_c4_multi.sh NGPUS NBOOTS [workdir]
set -e
N=${1:-8}; B=${2:-6}; W=${3:-/tmp/c4multi}; rm -rf $W; mkdir -p $W
python - <<PY
import sys, time
sys.path.insert(0, ".")
from tetrad_b200 import synth, database
from pathlib import Path
c = synth.CONFIGS["C4"]
t = time.time(); d = synth.generate(c["ntips"], c["nloci"], c["nsnps"], seed=1234)
database.write_snps_input(Path("$W/c4.snps.hdf5"), d.snps, d.snpsmap, d.names); print("input written in %.1f s" % (time.time() - t))
PY
python -m tetrad_b200 init $W/c4.snps.hdf5 -n c4 -w $W -q 1000000 -r 42 -s 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 -m tetrad_b200 run $W/c4.json -b $B 2>&1 | grep -E "finished rep|starting|rror"
ls $W | head -20
