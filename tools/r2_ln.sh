#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm or row_stats or ln_fold or gemm_linear or geglu" 2>&1 | tail -8
python tools/lnfold_probe.py > gpurun_out/lnfold_probe_new2.md 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ln2.json 2> gpurun_out/bench_ln2.err; tail -c 900 gpurun_out/bench_ln2.json
