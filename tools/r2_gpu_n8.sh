#!/bin/bash
# N-GPU scaling line exactly as the driver launches it
mkdir -p gpurun_out
N=${1:-8}
L=gpurun_out/r2_n$N.log
: > $L
echo "=== torchrun N=$N bench" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline >> $L 2>&1
echo "rc=$?" >> $L
nproc >> $L
grep -E "^===|rc=|Error|error" $L | head; grep -E '^\{"' $L | cut -c1-300
