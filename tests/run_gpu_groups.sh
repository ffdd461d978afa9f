#!/bin/bash
# Runs the GPU kernel tests one group per process (a trapped kernel poisons its CUDA context, not the next group's).
mkdir -p gpurun_out
for grp in "small_linear or timestep or cfg_steps or conv_in" "layernorm or groupnorm or temporal" "gemm_linear" "gemm_two or geglu or split_k or block_n or strided" "conv3x3" "tconv" "attention"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== group: $grp"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider 2>&1 | tail -40 | tee "gpurun_out/kt_${name}.log" | tail -25
done
