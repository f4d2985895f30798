#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_n2.log
: > $L
echo "=== torchrun N=2 bench" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 >> $L 2>&1
echo "rc=$?" >> $L
echo "=== torchrun N=2 reference arm" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 >> $L 2>&1
echo "rc=$?" >> $L
grep -E "^===|rc=|Error|error" $L | head; grep -E '^\{"' $L | cut -c1-600
