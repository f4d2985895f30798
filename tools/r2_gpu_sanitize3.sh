#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck of the third-generation attention kernel (persistent work loop,
# named-barrier turns, shared-memory exchanges): every parity shape with a capped grid (many items per CTA)
mkdir -p gpurun_out
S="compute-sanitizer --print-limit 20"
T="python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k"
for tool in racecheck synccheck memcheck; do
  O=gpurun_out/r2_attn3_${tool}.txt
  echo "== $tool: attention parity shapes, UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 UVLT_ATTN_GRID=5 (B <= 3)" > $O
  UVLT_ATTN_V=3 UVLT_ATTN_SPLIT=0 UVLT_ATTN_GRID=5 timeout 1200 $S --tool $tool $T "test_attention and not second and not third and not 32-361" >> $O 2>&1; echo "rc=$?" >> $O
  grep -E "^==|rc=|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O
done
