#!/bin/bash
# Kernel groups (one process each), whole-path parity, bench.
mkdir -p gpurun_out
bash tests/run_gpu_groups.sh 2>&1 | grep -E "===|passed|failed|FAILED|rel-L2|Error" | head -60
timeout 1500 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -s -p no:cacheprovider 2>&1 | grep -E "parity\]|passed|failed|FAILED|Error|error" | grep -v "call 1\|call 2\|run 1" | tee gpurun_out/unet_gpu.log | tail -40
timeout 1200 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS} 2>&1 | tail -4 | tee gpurun_out/bench_cfg2.log
