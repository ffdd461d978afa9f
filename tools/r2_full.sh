#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
export ASVA_PLAN_CACHE=gpurun_out/plans_full.txt
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1200 gpurun_out/bench_full.json
grep -c ", 4\]" gpurun_out/plans_full.txt; grep ", 4\]" gpurun_out/plans_full.txt | cut -c1-200 | head
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
