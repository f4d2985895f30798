#!/bin/bash
# full GPU suite + kernel sweeps + benches (both arms); logs under gpurun_out/
mkdir -p gpurun_out
L=gpurun_out/r2_full.log
: > $L
run() { echo "=== $*" >> $L; timeout 1200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -q -m gpu -p no:cacheprovider -s
run env SWEEP_BNS=0 python tools/kernel_sweep.py gemm 1 32
run python bench.py --steps 20 --warmup 5
run python bench.py --impl reference --steps 20 --warmup 5
grep -E "^===|rc=|passed|failed|Error|M= |rel_l2|max_abs|decisive" $L | cut -c1-260
grep -E '^\{"' $L | cut -c1-3000
