#!/bin/bash
# time-folded dilation-1 convs of the narrow HiFi-GAN stages (CTTA_FOLD=1) vs the plain form (default), same box
python -m pytest tests/test_kernels_gpu.py -x -q -k "time_folded" 2>&1 | tail -2
python -m pytest tests/test_parity_gpu.py -x -q -k "vocoder or end_to_end" 2>&1 | tail -2
for rep in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 plain : %.2f clips/s  %.2f ms' % (d['value'], d['ms_per_step']))"
  CTTA_FOLD=1   python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 folded: %.2f clips/s  %.2f ms' % (d['value'], d['ms_per_step']))"
done
python tools/profile_layers.py --batch 64 --out /tmp/l0.json | grep -E "vocoder:gemm|total eager"
CTTA_FOLD=1 python tools/profile_layers.py --batch 64 --out /tmp/l1.json | grep -E "vocoder:gemm|total eager|vocoder mode1 rows=(2621952|5243904|10487808)"
