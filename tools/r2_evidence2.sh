#!/bin/bash
# Round-2 closing evidence on HEAD (one gpurun call): bench lines (clips 2, cfg4), ncu launch list of one step, ncu --set
# full of the kernels added late in the round, GEMM census, norm / attention sweeps, compute-sanitizer on the new tests.
mkdir -p gpurun_out
export ASVA_PLAN_CACHE=gpurun_out/r2f_plans.txt
timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2f_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1
unset ASVA_PLAN_CACHE
timeout 600 python bench.py --steps 20 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_clips2.json
timeout 900 python bench.py --steps 10 --warmup 3 --workload cfg4 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_cfg4.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:temporal_mma -s 2 -c 1 -f -o gpurun_out/r2f_ncu_temporal \
    python tools/temporal_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm_rows -s 2 -c 1 -f -o gpurun_out/r2f_ncu_layernorm \
    python tools/norm_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn2_stats -s 1 -c 1 -f -o gpurun_out/r2f_ncu_gn2_stats \
    python tools/norm_probe.py --single > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn2_apply -s 1 -c 1 -f -o gpurun_out/r2f_ncu_gn2_apply \
    python tools/norm_probe.py --single > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_cluster -s 2 -c 1 -f -o gpurun_out/r2f_ncu_gn_cluster \
    python tools/norm_probe.py --single > /dev/null 2>&1
timeout 900 python tools/gemm_census.py --out gpurun_out/r2f_gemm_census.md > gpurun_out/r2f_gemm_census.log 2>&1
timeout 600 python tools/norm_probe.py > gpurun_out/r2f_norm_probe.md 2>&1
timeout 600 python tools/attn_probe.py --sweep --out gpurun_out/r2f_attn_sweep_timed.md > gpurun_out/r2f_attn_sweep_timed.log 2>&1
( time timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
    -k "(temporal or layernorm or groupnorm or row_stats or ln_fold or small_key or mma_all) and not 24576 and not 12288 and not 1024-8-40" ) > gpurun_out/r2f_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r2f_sanitizer_memcheck.log
ls -la gpurun_out/r2f_* | awk '{print $5, $9}'
