#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_full2.log
: > $L
run() { echo "=== $*" >> $L; timeout 1200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -q -m gpu -p no:cacheprovider -x
run env SWEEP_BNS=0 python tools/kernel_sweep.py gemm 1
run env SWEEP_BNS=0 UVLT_MULTICAST=1 python tools/kernel_sweep.py gemm 1
run python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run env UVLT_HEAD_CONV=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs
run compute-sanitizer --print-limit 5 --tool racecheck python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "test_attention and (129 or 40 or 81) and not second"
grep -E "^===|rc=|passed|failed|Error|M= |RACECHECK SUMMARY" $L | cut -c1-260
