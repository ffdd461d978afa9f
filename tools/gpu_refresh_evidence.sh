#!/bin/bash
# Evidence refresh on HEAD (gpurun -- bash tools/gpu_refresh_evidence.sh): driver-style GPU suite, smoke, bench, launch list, ncu --set full of the attention kernels.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_refresh.log
# (a profiler run must not re-tune: ASVA_PLAN_CACHE=<file written by a previous bench run> skips the tuning pass)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_refresh.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for s in spatial0 text0 audio0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 3 -f \
      -o gpurun_out/ncu_attn_$s python tools/attn_probe.py --single $s > gpurun_out/ncu_attn_$s.log 2>&1
done
ls -la gpurun_out | tail -20
