#!/bin/bash
# 16-bit residual stream at the two high-resolution VAE levels (default) vs fp32 everywhere (CTTA_VAE_STREAM16=0), same box
for v in 0 1; do
  echo "== CTTA_VAE_STREAM16=$v"
  CTTA_VAE_STREAM16=$v python -m pytest tests/test_parity_gpu.py -x -q -s -k "vae_decode or end_to_end or b64" 2>&1 | grep -E "vae rel|pipeline:|B=64|passed|failed"
done
for rep in 1 2; do for v in 0 1; do
  CTTA_VAE_STREAM16=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 stream16=$v: %.2f clips/s  %.2f ms' % (d['value'], d['ms_per_step']))"
done; done
