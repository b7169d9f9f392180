#!/bin/bash
# Multicast clusters of 2 (default) vs 4 CTAs, one box.
CTTA_MCAST_CS=4 CTTA_DEBUG=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | grep -E "passed|failed|co-resident" | sort | uniq -c | tail -4
for cs in 2 4; do
  export CTTA_MCAST_CS=$cs
  echo "== CTTA_MCAST_CS=$cs"
  python tools/run_one_gemm.py conv1d --c 512 --taps 11 --rows 5121 --batch 64 --kind c1 --iters 20 --seconds 1
  python tools/run_one_gemm.py conv1d --c 256 --taps 11 --rows 20484 --batch 64 --kind c2h --iters 20 --seconds 1
  python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res --iters 10 --seconds 1
  python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind geglu --iters 20 --seconds 1
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
done
CTTA_MCAST_CS=2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
CTTA_MCAST_CS=4 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-120
