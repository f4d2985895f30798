#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_attn_sweep.log
: > $L
for V in 1 3; do
  echo "=== UVLT_ATTN_V=$V" >> $L
  UVLT_ATTN_V=$V UVLT_ATTN_POLY=3 timeout 200 python tools/kernel_sweep.py attn 4 8 16 32 >> $L 2>&1
  UVLT_ATTN_V=$V UVLT_ATTN_POLY=3 SWEEP_NS=1193,1153 SWEEP_H=16 timeout 200 python tools/kernel_sweep.py attn 2 4 8 >> $L 2>&1
done
cat $L
