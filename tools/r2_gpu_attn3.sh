#!/bin/bash
# third-generation attention kernel: parity (all shapes, default grid and a capped grid), then timing against v1
mkdir -p gpurun_out
L=gpurun_out/r2_attn3.log
: > $L
run() { echo "=== $*" >> $L; timeout 240 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run env UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and not second and not third"
run env UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 UVLT_ATTN_GRID=5 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and not second and not third"
run env UVLT_ATTN_V=1 python tools/kernel_sweep.py attn 32 8
run env UVLT_ATTN_V=3 python tools/kernel_sweep.py attn 32 8
for v in 2 3 4 5; do run env UVLT_ATTN_V=3 UVLT_ATTN_POLY=$v SWEEP_NS=553 python tools/kernel_sweep.py attn 32; done
run env UVLT_ATTN_V=3 UVLT_ATTN_POLY=3 UVLT_ATTN_SPLIT=0 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and not second and not third"
run env UVLT_ATTN_V=1 SWEEP_NS=1193 SWEEP_H=16 python tools/kernel_sweep.py attn 8
run env UVLT_ATTN_V=3 SWEEP_NS=1193 SWEEP_H=16 python tools/kernel_sweep.py attn 8
run env UVLT_ATTN_V=3 UVLT_ATTN_POLY=3 SWEEP_NS=1193 SWEEP_H=16 python tools/kernel_sweep.py attn 8
python tools/attn3_trace.py 32 553 > gpurun_out/r2_attn3_trace_b32.txt 2>&1
grep -E "^===|rc=|passed|failed|Error|attn " $L | cut -c1-200
