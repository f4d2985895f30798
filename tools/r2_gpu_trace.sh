#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 4; do UVLT_ATTN_POLY=$v timeout 120 python tools/attn3_trace.py 32 553 > gpurun_out/r2_attn3_trace_b32_var$v.txt 2>&1; done
UVLT_ATTN_POLY=3 SWEEP_NS=553 timeout 120 python tools/kernel_sweep.py attn 32
