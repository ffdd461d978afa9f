import dataclasses, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch, gemm_probe
from asva_b200 import ops
be = ops.backend()
for name, plans in (("conv2", [(1, 64, 1, 1), (1, 128, 1, 1)]), ("tconv2", [(1, 64, 1, 2), (1, 128, 1, 2)])):
    spec = gemm_probe.SHAPES[name]()
    for (cg, bn, sp, epi) in plans:
        for cap in ((2, 3, 4, 6, 8) if bn == 64 else (2, 3)):
            os.environ["ASVA_GEMM_STAGES"] = str(cap)
            s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn, split_k=sp, cta_group=cg, epilogue=epi)
            st = be.gemm_plan(s)[3]
            us = []
            for dbg in (0, 3):
                os.environ["ASVA_GEMM_DBG"] = str(dbg)
                us.append(gemm_probe.time_spec(be, s))
            os.environ["ASVA_GEMM_DBG"] = "0"
            print(f"{name} bn={bn} cg={cg} epi={epi} stages cap {cap} -> {st}: normal {us[0]:.1f} us, skeleton {us[1]:.1f} us", flush=True)
