#!/bin/bash
# ncu --set full of single GEMM launches (shape, plan) for the epilogue study
mkdir -p gpurun_out
run() { # name shape plan
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2_ncu_$1 \
      python tools/gemm_probe.py --shapes $2 --single $3 > gpurun_out/r2_ncu_$1.log 2>&1
}
run lin0_e1 lin0 1,160,1,1
run lin0_e2 lin0 1,160,1,2
run lin1_e2 lin1 1,128,1,2
run geglu0 geglu0 1,128,1,1
ls -la gpurun_out/*.ncu-rep | tail
