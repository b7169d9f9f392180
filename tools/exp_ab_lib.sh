#!/bin/bash
# Same-box A/B of two builds of libctta (CTTA_LIB): previous commit vs working tree.
for lib in consistencytta_b200/libctta_prev.so consistencytta_b200/libctta.so; do
  export CTTA_LIB=$PWD/$lib
  echo "== $lib"
  for c in 32 64 128; do rows=$((163872*32/c)); for taps in 3 11; do for kind in c1 c2h; do
    python tools/run_one_gemm.py conv1d --c $c --taps $taps --rows $rows --batch 64 --kind $kind --iters 20 --seconds 1
  done; done; done
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
done
export CTTA_LIB=$PWD/consistencytta_b200/libctta_prev.so; python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
export CTTA_LIB=$PWD/consistencytta_b200/libctta.so; python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
