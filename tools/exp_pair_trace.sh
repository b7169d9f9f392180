#!/bin/bash
# Phase timeline of the fused ResBlock-pair kernel; the trace library is built HERE (CPU box) before gpurun:
#   nvcc ... -DRBP_TRACE=1 -c csrc/resblock_pair.cu -o /tmp/rbp_trace.o && nvcc -shared -o consistencytta_b200/libctta_trace.so <other .o> /tmp/rbp_trace.o
mkdir -p gpurun_out
for cfg in "64 3 1 81920" "32 3 1 163840"; do
  set -- $cfg
  CTTA_LIB=$PWD/consistencytta_b200/libctta_trace.so timeout 120 python tools/trace_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 64
done > gpurun_out/r2_pair_trace.txt 2>&1
