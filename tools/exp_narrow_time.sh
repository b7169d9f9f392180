#!/bin/bash
# Narrow-channel HiFi-GAN convolutions in isolation (20 back-to-back launches each).
mkdir -p gpurun_out
{
for c in 32 64 128 256; do
  rows=$((163872*32/c))
  for taps in 3 11; do
    for kind in c1 c2h; do
      python tools/run_one_gemm.py conv1d --c $c --taps $taps --dil 1 --rows $rows --batch 64 --kind $kind --iters 20
    done
  done
done
python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res --iters 10
python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res --iters 10
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16 --iters 10
} 2>&1 | tee gpurun_out/exp_narrow.txt
