#!/bin/bash
# last check of the tree as it will be judged: the whole GPU suite, then the driver's two bench commands
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
: > $L
run() { echo "=== $*" >> $L; timeout 1200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests -q -m gpu -p no:cacheprovider -x
run python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2>gpurun_out/r2_final_bench.err
timeout 600 python bench.py --steps 500 --warmup 20 --no-configs --no-cpu-baseline > gpurun_out/r2_final_bench500.json 2>/dev/null
grep -E "^===|rc=|passed|failed|Error|smoke" $L | cut -c1-200
python - <<'PY'
import json
for f in ('gpurun_out/r2_final_bench.json','gpurun_out/r2_final_bench500.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print('B1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_attention']['frac'], d['roofline']['traffic'], d['launches_per_step'], d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'))
    for k,v in d.get('configs',{}).items(): print(k, v['value'], v['ms_per_step'], v['e2e']['value'], v['roofline']['frac'], v['roofline_attention']['frac'], v['roofline_attention']['avg_launch_us'])
PY
