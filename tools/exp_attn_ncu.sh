#!/bin/bash
# ncu --set full of the flash-attention kernel (one launch) + phase timestamps of softmax warp 0
NCU="ncu --set full --clock-control none --import-source on"
for poly in 1; do
  CTTA_ATTN_POLY=$poly $NCU -k regex:flash_attn_tc -c 1 -o /tmp/attn_p$poly -f python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096 --iters 1 > /dev/null 2>&1
  { python tools/ncu_summary.py /tmp/attn_p$poly.ncu-rep; python tools/ncu_hot.py /tmp/attn_p$poly.ncu-rep 40; } > gpurun_out/r2b_ncu_attn_poly$poly.txt 2>&1
  cp /tmp/attn_p$poly.ncu-rep gpurun_out/r2b_attn_p$poly.ncu-rep
done
CTTA_ATTN_DEBUG=1 python tools/attn_phases.py > gpurun_out/r2b_attn_phases.txt 2>&1
