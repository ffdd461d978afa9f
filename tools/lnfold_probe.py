"""A/B timing of the row-statistics / LayerNorm-fold epilogue options on the headline workload's K = C GEMM shapes.
Every configuration is replayed from a CUDA graph (10 launches).  ASVA_LIB selects the library under test.

    python tools/lnfold_probe.py [--old]      (--old: the library has no stats / fold support: base timings only)"""
import argparse
import dataclasses
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from asva_b200 import ops  # noqa: E402
from gemm_probe import rnd, time_spec  # noqa: E402

DEV = "cuda"

# (name, M, C, N, geglu, (bn, cg))
PRODUCERS = [("o0", 24576, 320, 320, (160, 1)), ("o1", 6144, 640, 640, (128, 2)), ("o2", 1536, 1280, 1280, (128, 1)),
             ("o3", 384, 1280, 1280, (64, 1))]
CONSUMERS = [("q0", 24576, 320, 320, False, (160, 1)), ("q1", 6144, 640, 640, False, (128, 2)),
             ("q2", 1536, 1280, 1280, False, (128, 1)), ("q3", 384, 1280, 1280, False, (64, 1)),
             ("g0", 24576, 320, 2560, True, (128, 1)), ("g1", 6144, 640, 5120, True, (128, 1)),
             ("g2", 1536, 1280, 10240, True, (128, 2)), ("g3", 384, 1280, 10240, True, (128, 1))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--old", action="store_true")
    args = ap.parse_args()
    be = ops.backend()
    print(f"library: {os.environ.get('ASVA_LIB', 'default')}")
    print("| shape | M | C | N | bn/cg | epi | base us | +stats/fold us |")
    print("|---|---|---|---|---|---|---|---|")
    for name, M, C, N, (bn, cg) in PRODUCERS:
        x, w = rnd((M, C), 1), rnd((N, C), 2, 1 / math.sqrt(C))
        b, r = rnd((N,), 3, dtype=torch.float32), rnd((M, N), 4)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        st = torch.zeros(N // 32, M, 2, dtype=torch.float32, device=DEV)
        for epi in (1, 2, 3):
            sp = ops.spec_linear(x, w, out, bias=b, res0=r)
            sp.block_n, sp.cta_group, sp.split_k, sp.epilogue = bn, cg, 1, epi
            t0 = time_spec(be, sp)
            t1 = float("nan")
            if not args.old and epi != 2:
                t1 = time_spec(be, dataclasses.replace(sp, stats_out=st))
            print(f"| {name} +b+r | {M} | {C} | {N} | {bn}/{cg} | {epi} | {t0:.1f} | {t1:.1f} |", flush=True)
    for name, M, C, N, geglu, (bn, cg) in CONSUMERS:
        x, w = rnd((M, C), 5), rnd((N, C), 6, 1 / math.sqrt(C))
        b = rnd((N,), 7, dtype=torch.float32)
        out = torch.empty(M, N // 2 if geglu else N, dtype=torch.bfloat16, device=DEV)
        v = x.float().view(M, C // 32, 32)
        st = torch.stack([v.sum(-1), (v * v).sum(-1)], dim=-1).permute(1, 0, 2).contiguous()
        wsum = w.float().sum(1).contiguous()
        for epi in (1, 3):
            sp = ops.spec_linear(x, w, out, bias=b if geglu else None, geglu=geglu)
            sp.block_n, sp.cta_group, sp.split_k, sp.epilogue = bn, cg, 1, epi
            t0 = time_spec(be, sp)
            t1 = float("nan")
            if not args.old:
                f = dataclasses.replace(sp, bias=b)
                f.ln = ops.LnFold(stats=st, wsum=wsum, cols=C, eps=1e-5)
                t1 = time_spec(be, f)
            print(f"| {name}{' geglu' if geglu else ''} | {M} | {C} | {N} | {bn}/{cg} | {epi} | {t0:.1f} | {t1:.1f} |", flush=True)


if __name__ == "__main__":
    main()
