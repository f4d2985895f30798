#!/bin/bash
# iteration pass for the third-generation attention kernel: parity (default grid, capped grid), timing, timeline
mkdir -p gpurun_out
L=gpurun_out/r2_iter.log
: > $L
run() { echo "=== $*" >> $L; timeout 200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run env UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and not second and not third"
run env UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 UVLT_ATTN_GRID=5 UVLT_ATTN_POLY=3 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and not second and not third"
for v in 2 3; do run env UVLT_ATTN_V=3 UVLT_ATTN_POLY=$v SWEEP_NS=553,513 python tools/kernel_sweep.py attn 32 8; done
run env UVLT_ATTN_V=3 SWEEP_NS=1193 SWEEP_H=16 python tools/kernel_sweep.py attn 8
UVLT_ATTN_POLY=2 python tools/attn3_trace.py 32 553 > gpurun_out/r2_attn3_trace_b32.txt 2>&1
grep -E "^===|rc=|passed|failed|Error|attn " $L | cut -c1-200
