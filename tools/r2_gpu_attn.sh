#!/bin/bash
# Round-2 attention bring-up on the GPU box: parity of attention2 (both exp variants), A/B timing vs attention v1,
# then the whole GPU suite and a short B=32 bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
L=gpurun_out/r2_attn.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k attention -p no:cacheprovider
UVLT_ATTN_POLY=0 run env UVLT_ATTN_POLY=0 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k attention -p no:cacheprovider
run env UVLT_ATTN_V=1 python tools/kernel_sweep.py attn 1 32
run env UVLT_ATTN_POLY=0 python tools/kernel_sweep.py attn 1 32
run python tools/kernel_sweep.py attn 1 2 8 32
run python -m pytest tests -x -q -m gpu -p no:cacheprovider
run python bench.py --batch 32 --mode NLBBOX --steps 20 --warmup 5 --no-cpu-baseline
run python bench.py --steps 100 --warmup 10 --no-cpu-baseline
tail -c 6000 $L
