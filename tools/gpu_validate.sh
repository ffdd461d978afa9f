#!/bin/bash
# Final-tree validation: driver-style GPU suite, smoke, headline bench.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_final.json
