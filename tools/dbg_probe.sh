for dbg in 0 1 2 3; do for plan in "1,160,1" "1,64,1" "1,256,1" "2,256,1"; do echo "dbg=$dbg plan=$plan"; ASVA_GEMM_DBG=$dbg python tools/gemm_probe.py --shapes conv0 --timeplan $plan; done; done
