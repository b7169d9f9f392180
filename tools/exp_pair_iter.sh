#!/bin/bash
# one iteration of the ResBlock-pair tuning loop: determinism stress, per-shape timing, phase trace
mkdir -p gpurun_out
OUT=gpurun_out/r2_pair_iter.txt
: > $OUT
REPS=${REPS:-60} timeout 300 python tools/stress_pair.py 2>&1 | tail -9 >> $OUT
for cfg in "64 3 1 81920" "64 7 3 81920" "32 3 1 163840" "32 7 3 163840"; do
  set -- $cfg
  CTTA_DEBUG=1 timeout 120 python tools/run_one_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 64 --seconds 0.5 2>&1 | grep -E "^pair|ctta_resblock_pair" | sort -u >> $OUT
done
for cfg in "64 3 1 81920" "32 3 1 163840"; do
  set -- $cfg
  CTTA_LIB=$PWD/consistencytta_b200/libctta_trace.so timeout 120 python tools/trace_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 64 2>&1 | grep -v "^ *[0-9][0-9] " >> $OUT
done
