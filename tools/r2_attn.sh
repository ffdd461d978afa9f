#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "small_key or mma_all" 2>&1 | tail -4
S=text0,audio0,text1,audio1,spatial2,text2,audio2,spatial3
timeout 300 python tools/attn_probe.py --shapes $S --form 2 > gpurun_out/attn_form2.md 2>&1; cat gpurun_out/attn_form2.md
