#!/bin/bash
# GPU suite only (no -x: report every failure)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -s $PYTEST_EXTRA ) > gpurun_out/r2_pytest.log 2>&1
tail -15 gpurun_out/r2_pytest.log
grep "\[parity\]\|\[pipeline\]" gpurun_out/r2_pytest.log | grep -v "call 1\|call 2\|run 1" | tail -70
