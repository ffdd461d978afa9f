#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm" 2>&1 | tail -4
ASVA_LIB=$PWD/asva_b200/lib/libasva_b200_dbg.so timeout 600 python tools/norm_probe.py --forms > gpurun_out/gn_forms.md 2>&1; cat gpurun_out/gn_forms.md
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_gn.json 2> gpurun_out/bench_gn.err; tail -c 600 gpurun_out/bench_gn.json
