#!/bin/bash
# The CLI replicate loop under torchrun (one process per GPU, quartet ranges sharded, NCCL gather to rank 0) must write the
# same files as the single-process run.  usage: run_cli_c3_multi.sh NGPUS [workdir]
set -e
N=${1:-2}; W=${2:-/tmp/c3multi}; rm -rf $W; mkdir -p $W/one $W/many
python - <<PY
import sys
sys.path.insert(0, ".")
from tetrad_b200 import synth, database
from pathlib import Path
c = synth.CONFIGS["C3"]
d = synth.generate(c["ntips"], c["nloci"], c["nsnps"], seed=1234)
database.write_snps_input(Path("$W/c3.snps.hdf5"), d.snps, d.snpsmap, d.names)
PY
for k in one many; do python -m tetrad_b200 init $W/c3.snps.hdf5 -n c3 -w $W/$k -q 1000000 -r 42 -s 2>&1 | tail -1; done
python -m tetrad_b200 run $W/one/c3.json -b 2 2>&1 | grep -E "finished rep|rror"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 -m tetrad_b200 run $W/many/c3.json -b 2 2>&1 | grep -E "finished rep|rror"
for f in c3.quartets_0.tsv c3.quartets_1.tsv c3.quartets_2.tsv c3.qmc_in.txt; do cmp $W/one/$f $W/many/$f && echo "$f identical"; done
