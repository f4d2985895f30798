#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_iter2.log
: > $L
run() { echo "=== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py tests/test_forward_gpu.py -q -m gpu -x -p no:cacheprovider
run env SWEEP_BNS=0 python tools/kernel_sweep.py gemm 32
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 32 --mode NLBBOX --no-configs > gpurun_out/r2_iter2_bench.json 2>gpurun_out/r2_iter2_bench.err
grep -E "^===|rc=|passed|failed|Error|M= " $L | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_iter2_bench.json').read().strip().splitlines()[-1])
print('B32', d['value'], d['ms_per_step'], d['e2e']['value'], {k: v['us'] for k, v in d['roofline']['per_shape'].items()})
PY
