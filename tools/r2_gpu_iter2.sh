#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_iter2.log
: > $L
run() { echo "=== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python -m pytest tests/test_preproc_golden.py tests/test_tracker_gpu.py -q -m gpu -x -p no:cacheprovider
grep -E "^===|rc=|passed|failed|Error|assert" $L | cut -c1-200
