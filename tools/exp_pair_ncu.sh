#!/bin/bash
# ncu --set full of the fused ResBlock-pair kernel at the two k = 3 shapes
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for cfg in "64 3 1 81920 c64k3" "32 3 1 163840 c32k3"; do
  set -- $cfg
  $NCU -k regex:resblock_pair -s 3 -c 1 -f -o gpurun_out/r2_pair_$5 python tools/run_one_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 16 --iters 1 > /dev/null 2>&1
  { python tools/ncu_summary.py gpurun_out/r2_pair_$5.ncu-rep; python tools/ncu_hot.py gpurun_out/r2_pair_$5.ncu-rep 45; } > gpurun_out/r2_ncu_pair_$5.txt 2>&1
done
