#!/bin/bash
# Same-box pipeline A/B of the fused ResBlock-pair kernel (default on) against the two-launch path (CTTA_FUSE_PAIRS=0),
# after the GPU test suite with the default configuration.
mkdir -p gpurun_out
OUT=gpurun_out/r2_pair_ab.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT
for rep in 1 2; do
  for mode in 0 1; do
    r=$(CTTA_FUSE_PAIRS=$mode timeout 400 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.2f clips/s %.2f ms parity %s launches %s' % (d['value'], d['ms_per_step'], d.get('parity',{}).get('ok'), d.get('gpu_launches')))")
    echo "pipeline b64 CTTA_FUSE_PAIRS=$mode rep $rep: $r" | tee -a $OUT
  done
done
