#!/bin/bash
# First end-to-end GPU pass: whole-UNet parity, smoke, bench (tiny then the headline workload).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -s -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/unet_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --workload tiny --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/bench_tiny.log
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -8 | tee gpurun_out/bench_cfg2.log
