"""Where a GEMM launch spends its time: times fixed plans with ASVA_GEMM_DBG = 0 (normal), 1 (producers skip every
TMA load), 2 (issuer skips every MMA), 3 (both) - the library reads the variable at every call, so one process does
the whole sweep.  Results under dbg != 0 are garbage by construction; only the durations matter.

    python tools/gemm_dbg_sweep.py [--out gpurun_out/gemm_dbg_sweep.md]"""
import argparse
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import gemm_probe  # noqa: E402
from asva_b200 import ops  # noqa: E402

# shape -> (cta_group, block_n, split_k, epilogue): the plans the tuner picks for them in the headline workload
PLANS = {"conv0": (2, 160, 1, 1), "conv1": (2, 256, 1, 1), "conv2": (1, 128, 1, 1), "tconv2": (1, 128, 1, 2),
         "ff2_2": (1, 128, 1, 2), "lin2": (1, 128, 1, 2), "lin1": (1, 128, 1, 2), "qkv0": (1, 160, 1, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    be = ops.backend()
    lines = ["| shape | M | N | K | plan cg/bn/split/epi | K blocks per CTA | normal us | no loads us | no MMAs us | neither us |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    for name, (cg, bn, sp, epi) in PLANS.items():
        spec = gemm_probe.SHAPES[name]()
        s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn, split_k=sp, cta_group=cg, epilogue=epi)
        us = []
        for dbg in (0, 1, 2, 3):
            os.environ["ASVA_GEMM_DBG"] = str(dbg)
            us.append(gemm_probe.time_spec(be, s))
        os.environ["ASVA_GEMM_DBG"] = "0"
        m_tiles = (spec.M + 127) // 128
        tiles = ((m_tiles + cg - 1) // cg) * cg * ((spec.N + bn - 1) // bn)
        per_cta = -(-tiles // min(tiles, 148)) * (spec.K // 64)
        lines.append(f"| {name} | {spec.M} | {spec.N} | {spec.K} | {cg}/{bn}/{sp}/{epi} | {per_cta} | " +
                     " | ".join(f"{u:.1f}" for u in us) + " |")
        print(lines[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
