"""Does asva_gemm_tune pick the plan a CUDA-graph replay ranks first?  For each probe shape: the tuner's choice and
its own timing, then graph-replay timings of that plan and of a few fixed candidates."""
import ctypes as C
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import gemm_probe as gp  # noqa: E402
from asva_b200 import _lib, ops  # noqa: E402


def main():
    be = ops.backend()
    shapes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["conv2", "conv3", "tconv2", "lin2", "lin0", "conv1", "tconv0"]
    cands = [(1, 128, 1, 1), (1, 128, 1, 2), (2, 128, 1, 1), (2, 128, 1, 2), (2, 160, 1, 1), (2, 256, 1, 1)]
    for name in shapes:
        spec = gp.SHAPES[name]()
        for reps in (3, 20):
            d = be._gemm_desc(dataclasses.replace(spec, out=torch.zeros_like(spec.out)))
            bn, sp, cg, ep, us = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_float(0.0)
            _lib.check(be.lib.asva_gemm_tune(d, be._stream(), reps, C.byref(bn), C.byref(sp), C.byref(cg), C.byref(ep),
                                             C.byref(us)), "tune")
            s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn.value, split_k=sp.value,
                                    cta_group=cg.value, epilogue=ep.value)
            print(f"{name}: tune(reps={reps}) -> cg={cg.value} bn={bn.value} split={sp.value} epi={ep.value} "
                  f"tuner {us.value:.1f} us, graph {gp.time_spec(be, s):.1f} us", flush=True)
        row = []
        for cg_, bn_, sp_, ep_ in cands:
            s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn_, split_k=sp_, cta_group=cg_, epilogue=ep_)
            try:
                row.append(f"{cg_}/{bn_}/{sp_}/{ep_}: {gp.time_spec(be, s):.1f}")
            except _lib.AsvaError:
                row.append(f"{cg_}/{bn_}/{sp_}/{ep_}: -")
        print("   graph: " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
