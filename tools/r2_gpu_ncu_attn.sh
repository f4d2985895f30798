#!/bin/bash
# ncu --set full capture (with source-level stall sampling) of the attention kernel at B=32 n=553 and B=1 n=513
mkdir -p gpurun_out
L=gpurun_out/r2_ncu_attn.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k attention -p no:cacheprovider
run python tools/kernel_sweep.py attn 1 32
run ncu --set full --import-source on --clock-control none -k regex:attention2 -s 2 -c 1 -f -o gpurun_out/r2_attn2_b32 python tools/attn_one.py 32 553 4
run ncu --set full --import-source on --clock-control none -k regex:attention2 -s 2 -c 1 -f -o gpurun_out/r2_attn2_b1 python tools/attn_one.py 1 513 4
tail -c 3000 $L
