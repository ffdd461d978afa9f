#!/bin/bash
# round 2, GPU call 1: full GPU suite on the refactored tree + headline bench + clips-per-gpu 2 + config 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/r2c1_gpu.txt
( time timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --deselect "tests/test_unet_gpu.py::test_sampler_trace_sd15_50_steps[pndm]" -s ) > gpurun_out/r2c1_pytest.log 2>&1
tail -5 gpurun_out/r2c1_pytest.log
grep "\[parity\]\|\[pipeline\]" gpurun_out/r2c1_pytest.log | tail -60
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2c1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2c1_bench.json
timeout 600 python bench.py --steps 20 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2c1_bench_clips2.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2c1_bench_cfg4.json
