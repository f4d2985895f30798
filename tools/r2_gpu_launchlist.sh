#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for cfg in "b1 --batch 1" "b32 --batch 32 --mode NLBBOX"; do
  set -- $cfg; tag=$1; shift
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 1400 --csv --log-file gpurun_out/r2_${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs "$@" > gpurun_out/r2_${tag}_launches.log 2>&1
  echo "$tag rc=$?"
done
