#!/bin/bash
# full GPU suite + kernel sweeps + short benches; logs under gpurun_out/
mkdir -p gpurun_out
L=gpurun_out/r2_full.log
: > $L
run() { echo "=== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -x -q -m gpu -p no:cacheprovider
run env SWEEP_BNS=0 python tools/kernel_sweep.py gemm 1 8 32
run python tools/kernel_sweep.py attn 1 8 32
run env UVLT_ATTN_V=1 python tools/kernel_sweep.py attn 1 8 32
run python bench.py --steps 100 --warmup 10 --no-cpu-baseline
run python bench.py --batch 32 --mode NLBBOX --steps 20 --warmup 5 --no-cpu-baseline
run env UVLT_ATTN_V=1 python bench.py --batch 32 --mode NLBBOX --steps 20 --warmup 5 --no-cpu-baseline
grep -E "^===|rc=|attn |passed|failed|Error|M= " $L | cut -c1-220
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $L
