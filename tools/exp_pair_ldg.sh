#!/bin/bash
# ResBlock-pair kernel (residual rows from global memory, E1 / E2 on separate warp sets): determinism stress with two and
# one slots, kernel tests, per-shape timing against the two-launch path, and a same-box pipeline A/B.
mkdir -p gpurun_out
OUT=gpurun_out/r2_pair_split2.txt
: > $OUT
for s in 2; do
  echo "== stress CTTA_RBP_SLOTS=$s" | tee -a $OUT
  CTTA_RBP_SLOTS=$s REPS=${REPS:-150} timeout 400 python tools/stress_pair.py 2>&1 | tail -10 | tee -a $OUT
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "resblock_pair" 2>&1 | tail -2 | tee -a $OUT
for s in 2; do
  echo "== timing CTTA_RBP_SLOTS=$s (sustained)" | tee -a $OUT
  for cfg in "64 3 1 81920" "64 7 3 81920" "32 3 1 163840" "32 7 3 163840" "32 11 5 163840" "32 11 1 163840"; do
    set -- $cfg
    CTTA_RBP_SLOTS=$s timeout 120 python tools/run_one_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 64 --seconds 0.5 2>&1 | tail -2 | tee -a $OUT
  done
done
for rep in 1 2; do
  for mode in 0 1; do
    r=$(CTTA_FUSE_PAIRS=$mode timeout 400 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.2f clips/s %.2f ms parity %s' % (d['value'], d['ms_per_step'], d.get('parity',{}).get('ok')))")
    echo "pipeline b64 fuse=$mode rep $rep: $r" | tee -a $OUT
  done
done
