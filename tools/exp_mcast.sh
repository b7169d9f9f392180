#!/bin/bash
# CTA pairs with multicast weight chunks (default) vs independent CTAs (CTTA_NO_MCAST=1), sustained clocks, batch 64.
mkdir -p gpurun_out
{
for off in 1 0; do
  if [ $off = 1 ]; then export CTTA_NO_MCAST=1; else unset CTTA_NO_MCAST; fi
  echo "== CTTA_NO_MCAST=$off"
  python tools/run_one_gemm.py conv1d --c 512 --taps 11 --rows 5121 --batch 64 --kind c1 --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 512 --taps 11 --rows 5121 --batch 64 --kind c2h --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 512 --taps 3 --rows 5121 --batch 64 --kind c2h --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 256 --taps 11 --rows 20484 --batch 64 --kind c1 --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 256 --taps 7 --rows 20484 --batch 64 --kind c2h --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 128 --taps 11 --rows 40968 --batch 64 --kind c1 --iters 20 --seconds 1
  python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res --iters 10 --seconds 1
  python tools/run_one_gemm.py conv2d --c 256 --h 512 --w 32 --batch 64 --kind f32res --iters 10 --seconds 1
  python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res --iters 10 --seconds 1
done
} 2>&1 | tee gpurun_out/exp_mcast.txt
