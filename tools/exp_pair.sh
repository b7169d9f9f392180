#!/bin/bash
# fused ResBlock pair vs the two-launch path, every narrow-stage shape of HIFIGAN_16K_64 at B = 64, sustained clocks
for c in 32 64; do
  t=$((163872*32/c))
  for taps in 3 7 11; do for dil in 1 3 5; do
    python tools/run_one_pair.py --c $c --taps $taps --dil $dil --t $t --batch 64 --iters 10 --seconds 0.5
  done; done
done
for rep in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 unfused: %.2f clips/s  %.2f ms' % (d['value'], d['ms_per_step']))"
  CTTA_FUSE_PAIRS=1   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 fused:   %.2f clips/s  %.2f ms' % (d['value'], d['ms_per_step']))"
done
