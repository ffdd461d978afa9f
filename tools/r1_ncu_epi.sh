#!/bin/bash
# ncu --set full of one GEMM launch per epilogue form on the conv_temp GEMM of level 2 (1536 x 1280 x 3840, bias + addend + 2 residuals)
mkdir -p gpurun_out
for epi in 1 2; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/ncu_gemm_tconv2_epi$epi \
      python tools/gemm_probe.py --shapes tconv2 --single 1,128,1,$epi > gpurun_out/ncu_gemm_tconv2_epi$epi.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
