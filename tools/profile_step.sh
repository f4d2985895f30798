#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu captures of the dominant kernels of one bench step.
#   tools/profile_step.sh <tag> [bench args...]
# Outputs under gpurun_out/: <tag>_launches.csv, <tag>_gemm.ncu-rep, <tag>_attn.ncu-rep
set -u
TAG=$1; shift
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
# the four layer GEMMs of a middle layer (skip the first ~300 GEMM launches: init + warm-up)
timeout 600 $NCU --set full --import-source on -k regex:gemm_bf16 -s 300 -c 4 -f -o gpurun_out/${TAG}_gemm \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:attention -s 60 -c 2 -f -o gpurun_out/${TAG}_attn \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_attn.log 2>&1
echo "attn capture rc=$?"
