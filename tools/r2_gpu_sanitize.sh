#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck over the operator tests and the smoke forward
mkdir -p gpurun_out
S="compute-sanitizer --print-limit 20"
T="python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k"
SEL="(test_attention or test_gemm or test_layernorm) and not multicast and not second_generation and not 17696"
for tool in racecheck synccheck memcheck; do
  echo "== $tool: operator tests (default kernels)" > gpurun_out/r2_${tool}.txt
  timeout 1500 $S --tool $tool $T "$SEL" >> gpurun_out/r2_${tool}.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_${tool}.txt
  echo "== $tool: attention tests, second-generation kernel (UVLT_ATTN_V=2)" >> gpurun_out/r2_${tool}.txt
  UVLT_ATTN_V=2 UVLT_ATTN_POLY=1 timeout 900 $S --tool $tool $T "test_attention and not second_generation" >> gpurun_out/r2_${tool}.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_${tool}.txt
  echo "== $tool: smoke() (engine forward, batch 3, mixed flags)" >> gpurun_out/r2_${tool}.txt
  timeout 900 $S --tool $tool python __graft_entry__.py smoke >> gpurun_out/r2_${tool}.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_${tool}.txt
  grep -E "^==|rc=|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|smoke ok" gpurun_out/r2_${tool}.txt
done
