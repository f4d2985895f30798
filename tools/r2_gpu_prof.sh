#!/bin/bash
mkdir -p gpurun_out
bash tools/profile_step.sh r2c_b1 --no-configs
bash tools/profile_step.sh r2c_b32 --no-configs --batch 32 --mode NLBBOX
ls -la gpurun_out/r2c_*
