#!/bin/bash
# Same-box A/B of the flash-attention kernel: previous build (consistencytta_b200/libctta_prev.so) vs this tree, kernel
# alone (CTTA_ATTN_POLY = eighths of the exponentials on the FMA pipe) and inside the UNet (graph replay, B = 32).
L=$PWD/consistencytta_b200
python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -2
for cfg in "--b 64 --heads 5 --lq 4096 --lk 4096" "--b 64 --heads 10 --lq 1024 --lk 1024" "--b 64 --heads 20 --lq 256 --lk 256"; do
  CTTA_LIB=$L/libctta_prev.so python tools/run_one_op.py attention $cfg --iters 20 | sed "s/^/prev        /"
  for poly in 0 2 3 4; do
    CTTA_ATTN_POLY=$poly python tools/run_one_op.py attention $cfg --iters 20 | sed "s/^/new poly=$poly\/8  /"
  done
done
CTTA_ATTN_DEBUG=1 python tools/attn_phases.py
for rep in 1 2; do for lib in libctta_prev.so libctta.so; do
  CTTA_LIB=$L/$lib python bench.py --unet-only --batch 32 --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unet b32 $lib: %.3f ms' % d['value'])"
  CTTA_LIB=$L/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline b64 $lib: %.2f clips/s' % d['value'])"
done; done
