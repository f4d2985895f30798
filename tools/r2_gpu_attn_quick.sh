#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_attn_quick.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k attention -p no:cacheprovider
run python tools/kernel_sweep.py attn 1 2 8 32
run env UVLT_ATTN_POLY=0 python tools/kernel_sweep.py attn 1 32
python tools/attn_trace.py 32 553 > gpurun_out/r2_trace_b32.txt 2>&1
python tools/attn_trace.py 1 513 > gpurun_out/r2_trace_b1.txt 2>&1
grep -E "^===|rc=|attn |passed|failed|Error" $L
