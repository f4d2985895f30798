#!/bin/bash
# 8-GPU (or N-GPU) scaling line exactly as the driver launches it, preceded by the tracker tests on one GPU
mkdir -p gpurun_out
L=gpurun_out/r2_n8.log
: > $L
N=${1:-8}
echo "=== tracker / preprocess tests (1 GPU)" >> $L
timeout 600 python -m pytest tests/test_tracker_gpu.py tests/test_preprocess_gpu.py tests/test_evaluation.py -q -m gpu -x -p no:cacheprovider >> $L 2>&1
echo "rc=$?" >> $L
echo "=== torchrun N=$N bench" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 >> $L 2>&1
echo "rc=$?" >> $L
nproc >> $L; free -g | head -2 >> $L
grep -E "^===|rc=|Error|error|passed|failed" $L | head; grep -E '^\{"' $L | cut -c1-400
