#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "temporal or layernorm" 2>&1 | tail -8
timeout 600 python tools/temporal_probe.py > gpurun_out/temporal_probe.md 2>&1; cat gpurun_out/temporal_probe.md
timeout 600 python tools/norm_probe.py > gpurun_out/norm_probe.md 2>&1; tail -6 gpurun_out/norm_probe.md
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; tail -c 700 gpurun_out/bench_t.json
