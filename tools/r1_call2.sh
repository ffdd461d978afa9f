#!/bin/bash
# Per-warp GEMM epilogue + attention exp-pass experiments: parity, A/B timing, bench with the tuner choosing per shape.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_v9.log
timeout 400 python tools/gemm_probe.py --shapes lin0,lin1,lin2,lin3,ff2_0,ff2_2,qkv0,tconv0,tconv1,tconv2,tconv3,conv0,conv1,conv2 --splits 1 --out gpurun_out/gemm_probe_v9.md > gpurun_out/gemm_probe_v9.log 2>&1
grep "^## " gpurun_out/gemm_probe_v9.md
for v in "ASVA_ATTN_PIPE=0" "ASVA_ATTN_POLY=3" "ASVA_ATTN_POLY=2" "ASVA_ATTN_POLY=1" "ASVA_ATTN_POLY=0" "ASVA_ATTN_POLY=4"; do
  echo "== $v"; env $v timeout 120 python tools/attn_probe.py --shapes spatial0,text0,audio0 2>&1 | grep -E "^\| (spatial|text|audio)"
done 2>&1 | tee gpurun_out/attn_variants_v9.log
ASVA_PLAN_CACHE=gpurun_out/plans_v9.txt timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_v9.log
ASVA_PLAN_CACHE=gpurun_out/plans_v9.txt timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v9.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ASVA_PLAN_CACHE=gpurun_out/plans_v9.txt timeout 400 python tools/gemm_census.py --out gpurun_out/gemm_census_v9.md > gpurun_out/gemm_census_v9.log 2>&1
tail -3 gpurun_out/gemm_census_v9.log
ls -la gpurun_out | tail -20
