#!/bin/bash
# Round-2 evidence on HEAD (one gpurun call): bench + plan cache, ncu launch list of one step, ncu --set full of the top
# GEMM shapes and the attention kernels, the config-5 attention sweep with tensor-pipe %, compute-sanitizer on the
# kernel tests.  Everything lands in gpurun_out/r2e_*; summaries are made locally (tools/r2_evidence_summarize.py).
mkdir -p gpurun_out
export ASVA_PLAN_CACHE=gpurun_out/r2e_plans.txt
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/r2e_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_under_ncu.log 2>&1
unset ASVA_PLAN_CACHE
gemm() { # name shape plan
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/r2e_ncu_gemm_$1 \
      python tools/gemm_probe.py --shapes $2 --single $3 > gpurun_out/r2e_ncu_gemm_$1.log 2>&1
}
gemm conv0 conv0 2,160,1,1
gemm conv1 conv1 2,256,1,1
gemm conv2 conv2 2,128,1,1
gemm lin0 lin0 1,160,1,3
gemm geglu0 geglu0 2,128,1,3
for s in spatial0 text0 audio0 spatial1 spatial0_hr; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 2 -c 1 -f \
      -o gpurun_out/r2e_ncu_attn_$s python tools/attn_probe.py --single $s > gpurun_out/r2e_ncu_attn_$s.log 2>&1
done
timeout 600 python tools/attn_probe.py --sweep --out gpurun_out/r2e_attn_sweep_timed.md > gpurun_out/r2e_attn_sweep_timed.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:attn_tc_kernel --csv --log-file gpurun_out/r2e_attn_sweep_ncu.csv \
    python tools/attn_probe.py --sweep --once --out gpurun_out/r2e_attn_sweep_labels.md > gpurun_out/r2e_attn_sweep_once.log 2>&1
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
    -k "not 24576 and not config4 and not 1024-1024 and not 12288" ) > gpurun_out/r2e_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r2e_sanitizer_memcheck.log
ls -la gpurun_out/r2e_* | awk '{print $5, $9}'
