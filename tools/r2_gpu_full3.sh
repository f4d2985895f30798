#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_full3.log
: > $L
run() { echo "=== $*" >> $L; timeout 1200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -q -m gpu -p no:cacheprovider -x
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_full3_bench.json 2>gpurun_out/r2_full3_bench.err
grep -E "^===|rc=|passed|failed|Error" $L | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_full3_bench.json').read().strip().splitlines()[-1])
print('B1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_phases_ms_per_step'], d['roofline_attention']['avg_launch_us'], d['roofline']['per_shape'])
for k,v in d['configs'].items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['roofline_attention']['avg_launch_us'], v['roofline_attention']['frac'], v['roofline']['frac'], v['roofline']['per_shape'])
PY
bash tools/r2_gpu_sanitize3.sh
