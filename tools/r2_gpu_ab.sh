#!/bin/bash
# A/B of library builds on ONE box: AB_TAGS = space-separated tags; "new" = the in-tree library, "r2start" = the build of
# the round's first commit (uvltrack_b200/libuvlt_sm100_r2start.so), other tags = uvltrack_b200/libuvlt_ab_<tag>.so.
# (This is how the silently flipped UVLT_MULTICAST default was found: same GEMM SASS, 6 % slower batch-1 step.)
mkdir -p gpurun_out
for rep in 1 2; do
  for tag in ${AB_TAGS:-r2start new}; do
    lib=$PWD/uvltrack_b200/libuvlt_ab_$tag.so
    [ "$tag" = new ] && lib=""
    [ "$tag" = r2start ] && lib=$PWD/uvltrack_b200/libuvlt_sm100_r2start.so
    UVLT_LIB=$lib timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-configs > gpurun_out/r2_ab_${tag}_$rep.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2_ab_${tag}_$rep.json').read().strip().splitlines()[-1])
print('$tag', $rep, d['value'], d['ms_per_step'], d['e2e']['value'], {k: v['us'] for k, v in d['roofline']['per_shape'].items()}, d['roofline_attention']['avg_launch_us'])
PY
  done
done
