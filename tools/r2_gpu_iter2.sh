#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_iter2.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_preprocess_gpu.py tests/test_tracker_gpu.py tests/test_evaluation.py -q -m gpu -x -p no:cacheprovider
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_iter2_bench.json 2>gpurun_out/r2_iter2_bench.err
timeout 600 python bench.py --steps 500 --warmup 20 --no-cpu-baseline --no-configs > gpurun_out/r2_iter2_bench500.json 2>/dev/null
grep -E "^===|rc=|passed|failed|Error" $L | cut -c1-200
python - <<'PY'
import json
for f in ('gpurun_out/r2_iter2_bench.json','gpurun_out/r2_iter2_bench500.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print('B1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_phases_ms_per_step'])
    for k,v in d.get('configs',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['e2e']['ms_per_step'], v['e2e_phases_ms_per_step'])
PY
