#!/bin/bash
# Narrow-channel HiFi-GAN convolutions in isolation (20 back-to-back launches each) + two ncu captures.
mkdir -p gpurun_out
{
for c in 32 64 128 256; do
  rows=$((163872*32/c))
  for taps in 3 7 11; do
    for kind in c1 c2h; do
      python tools/run_one_gemm.py conv1d --c $c --taps $taps --dil 1 --rows $rows --batch 64 --kind $kind --iters 20
    done
  done
done
} 2>&1 | tee gpurun_out/exp_narrow.txt
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/n_c32_t3 python tools/run_one_gemm.py conv1d --c 32 --taps 3 --rows 163872 --batch 16 --kind c2h > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/n_c128_t11 python tools/run_one_gemm.py conv1d --c 128 --taps 11 --rows 40968 --batch 16 --kind c2h > /dev/null 2>&1
for f in n_c32_t3 n_c128_t11; do { python tools/ncu_summary.py gpurun_out/$f.ncu-rep; python tools/ncu_hot.py gpurun_out/$f.ncu-rep 24; } > gpurun_out/$f.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
