#!/bin/bash
# LayerNorm-fold check: kernel tests, whole-UNet parity, A/B bench
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "row_stats or ln_fold or gemm_linear or geglu" 2>&1 | tail -15
python -m pytest tests/test_unet_gpu.py -x -q -s 2>&1 | grep -E "rel-L2|passed|failed|Error|error" | tail -30
ASVA_LN_FOLD=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_nofold.json 2> gpurun_out/bench_nofold.err; tail -c 1500 gpurun_out/bench_nofold.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; tail -c 1500 gpurun_out/bench_fold.json
