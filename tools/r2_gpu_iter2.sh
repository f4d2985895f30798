#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_iter2.log
: > $L
run() { echo "=== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py tests/test_forward_gpu.py -q -m gpu -x -p no:cacheprovider
run env SWEEP_BNS=0 python tools/kernel_sweep.py gemm 1 8 32
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_iter2_bench.json 2>gpurun_out/r2_iter2_bench.err
grep -E "^===|rc=|passed|failed|Error|M= " $L | cut -c1-200
python - <<'PY'
import json
for f in ('gpurun_out/r2_iter2_bench.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print('B1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: v['us'] for k, v in d['roofline']['per_shape'].items()})
    for k,v in d.get('configs',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['e2e']['ms_per_step'], {k2: v2['us'] for k2, v2 in v['roofline']['per_shape'].items()})
PY
