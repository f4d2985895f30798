#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_iter2.log
: > $L
run() { echo "=== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "layernorm or split_k"
run python -m pytest tests/test_forward_gpu.py tests/test_tracker_gpu.py -q -m gpu -x -p no:cacheprovider
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-configs > gpurun_out/r2_iter2_b1.json 2>/dev/null
grep -E "^===|rc=|passed|failed|Error|M= " $L | cut -c1-200
python - <<'PY'
import json
for f in ('gpurun_out/r2_iter2_b1.json',):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d['config']['sequences_per_gpu'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: v['us'] for k, v in d['roofline']['per_shape'].items()}, d['roofline_attention']['avg_launch_us'])
PY
