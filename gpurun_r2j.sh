#!/bin/bash
# warp-mix experiment: count/score warp split of resolve_mma_kernel<true> at C4, C3 and C2
mkdir -p gpurun_out
L=gpurun_out/r2j.log; : > $L
for C in C4 C2; do
 for MIX in 0 1 2 3; do
  echo "mix $MIX" >> $L
  TQ_WARP_MIX=$MIX timeout 600 python profiles/tools/r2_kernel_check.py time $C mma sub 3 >> $L 2>&1
 done
done
grep -v "^$" $L
