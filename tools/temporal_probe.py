"""Temporal attention, both forms (memory-bound default / tcgen05 by name), on the headline shapes and the config-5
frame counts: CUDA-graph replay of 10 calls over rotating buffers; GB/s on q|k|v read + out written."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from asva_b200 import ops  # noqa: E402
from norm_probe import timeit  # noqa: E402

DEV = "cuda"


def main():
    be = ops.backend()
    print("| B | F | N | C | d | warp-MMA us | GB/s | thread/query us | GB/s | tcgen05 us | GB/s |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    H = 8
    for F in (8, 12, 16, 24):
        for N, C in ((1024, 320), (256, 640), (64, 1280), (16, 1280)):
            B, d = 2, C // H
            NB = 4
            qkv = [torch.randn(B, F, N, 3 * C, device=DEV).bfloat16() for _ in range(NB)]
            out = [torch.empty(B, F, N, C, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
            by = B * F * N * 4 * C * 2
            t = []
            for form in (3, 2, 1):
                t.append(timeit(lambda i: be.temporal_attention(qkv[i % NB], out[i % NB], B, F, N, H, d,
                                                                1 / math.sqrt(d), form=form)))
            print(f"| {B} | {F} | {N} | {C} | {d} | " + " | ".join(f"{x:.1f} | {by / x / 1e3:.0f}" for x in t) + " |",
                  flush=True)


if __name__ == "__main__":
    main()
