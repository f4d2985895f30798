#!/bin/bash
# Round-2 evidence pass on one B200: GPU suite, driver-style bench lines (both arms), launch lists and full ncu captures.
mkdir -p gpurun_out
L=gpurun_out/r2b_all.log
: > $L
run() { echo "=== $*" >> $L; timeout 1200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -q -m gpu -p no:cacheprovider -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_default.json 2>gpurun_out/r2b_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_bench_reference.json 2>gpurun_out/r2b_bench_reference.err
timeout 600 python bench.py --steps 500 --warmup 20 --no-configs --no-cpu-baseline > gpurun_out/r2b_bench_b1_500.json 2>/dev/null
bash tools/profile_step.sh r2b_b1 --no-configs >> $L 2>&1
bash tools/profile_step.sh r2b_b32 --no-configs --batch 32 --mode NLBBOX >> $L 2>&1
grep -E "^===|rc=|passed|failed|Error" $L | cut -c1-200
tail -c 600 gpurun_out/r2b_bench_default.json | head -c 300; echo
cat gpurun_out/r2b_bench_reference.json | cut -c1-400
