#!/bin/bash
# Final validation + bench lines on the round-2 HEAD: GPU suite, smoke, cfg2 / two clips / cfg4 lines, launch list.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
export ASVA_PLAN_CACHE=gpurun_out/r2g_plans.txt
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/r2g_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
unset ASVA_PLAN_CACHE
timeout 600 python bench.py --steps 20 --warmup 3 --clips-per-gpu 2 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2g_bench_clips2.json
timeout 900 python bench.py --steps 10 --warmup 3 --workload cfg4 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2g_bench_cfg4.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 600 python tools/norm_probe.py > gpurun_out/r2g_norm_probe.md 2>&1
python - <<'PY'
import json
for f in ("r2g_bench","r2g_bench_clips2","r2g_bench_cfg4"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "frac", round(r["frac"],3), "whole", round(r["whole_step_frac"],3), r["family_ms"], d["clocks"]["reasons"])
PY
