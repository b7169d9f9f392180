#!/bin/bash
# shared-memory plan matrix of the ResBlock-pair kernel (staging buffers x X ring depth) per shape
mkdir -p gpurun_out
OUT=gpurun_out/r2_pair_plans.txt
: > $OUT
for cfg in "64 3 1 81920" "32 3 1 163840" "32 7 3 163840" "32 11 5 163840" "64 7 3 81920"; do
  set -- $cfg
  for st in 2 1 0; do for x in 3 2 1; do
    CTTA_RBP_STAGING=$st CTTA_RBP_X=$x CTTA_DEBUG=1 timeout 120 python tools/run_one_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 64 --seconds 0.3 2>&1 | grep -E "fused |ctta_resblock_pair" | sort -u | tr '\n' ' ' >> $OUT; echo >> $OUT
  done; done
done
