"""Ring depth vs launch time: fixed GEMM plans with the stage count capped by ASVA_GEMM_STAGES, normal and with
ASVA_GEMM_DBG=3 (neither TMA loads nor MMAs).  If the protocol-only time does not move with the depth, the ring is
paced by a single thread's loop iteration, not by a barrier round trip.

    python tools/gemm_stages_sweep.py"""
import dataclasses, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
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
