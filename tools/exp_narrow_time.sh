#!/bin/bash
# Narrow-channel HiFi-GAN convolutions in isolation at sustained clocks, with 64-byte (legacy) and 128-byte epilogue rows.
mkdir -p gpurun_out
{
for legacy in 1 0; do
  if [ $legacy = 1 ]; then export CTTA_NO_WIDE16=1; else unset CTTA_NO_WIDE16; fi
  echo "== 64-byte 16-bit epilogue rows (legacy): $legacy"
  for c in 32 64 128 256; do
    rows=$((163872*32/c))
    for taps in 3 7 11; do
      for kind in c1 c2h; do
        python tools/run_one_gemm.py conv1d --c $c --taps $taps --dil 1 --rows $rows --batch 64 --kind $kind --iters 20 --seconds 1.0
      done
    done
  done
done
} 2>&1 | tee gpurun_out/exp_narrow.txt
