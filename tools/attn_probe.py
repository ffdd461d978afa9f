"""Times asva_attention on the attention shapes of the headline workload (and checks them against the torch spec
interpreter); `--single NAME` launches one shape 3 times eagerly for ncu.

    python tools/attn_probe.py [--out gpurun_out/attn_probe.md] [--single spatial0]"""
import argparse
import dataclasses
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from asva_b200 import ops  # noqa: E402
from sim_backend import SimBackend  # noqa: E402

DEV = "cuda"
# name: (G, R, Nk, d)   heads = 8
SHAPES = {
    "spatial0": (2, 12288, 1024, 40), "text0": (2, 12288, 77, 40), "audio0": (24, 1024, 25, 40),
    "spatial1": (2, 3072, 256, 80), "text1": (2, 3072, 77, 80), "audio1": (24, 256, 25, 80),
    "spatial2": (2, 768, 64, 160), "text2": (2, 768, 77, 160), "audio2": (24, 64, 25, 160),
    "spatial3": (2, 192, 16, 160),
    "spatial0_hr": (2, 98304, 4096, 40),
}


def rnd(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16).to(DEV)


def make(name):
    G, R, Nk, d = SHAPES[name]
    H = 8
    C = H * d
    dpad = ((d + 63) // 64) * 64
    q, kv = rnd((G * R, C), 1), rnd((G * Nk, 2 * C), 2)
    out = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
    return ops.AttnSpec(q=q, kv=kv, out=out, G=G, heads=H, R=R, Nk=Nk, d=d, dpad=dpad, ldq=C, ldkv=2 * C, ldo=C,
                        kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d))


def sweep(out_path, once=False):
    """BASELINE.json config 5: spatial 32^2..64^2 tokens x head dims, temporal 8..24 frames, cross-attention 1..256 keys.
    Core FLOPs = 4 Nq Nk d per (group, head); bytes = Q + O + K + V in bf16.
    once=True: every shape is launched exactly twice, eagerly (warm-up + one), for an `ncu --metrics ...` pass whose
    launch list is then joined with the labels written here (tools/join_attn_sweep.py)."""
    import math as _m
    be = ops.backend()
    H = 8
    lines = ["| kind | tokens / frames / keys | C | d | us | core TFLOP/s | Q+O+KV GB/s |", "|---|---|---|---|---|---|---|"]

    def timeit(fn):
        if once:
            fn()
            fn()
            torch.cuda.synchronize()
            return float("nan")
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(5):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 200

    B, F = 2, 12
    for side in (32, 48, 64):
        for C in (320, 640, 1280):
            d, N = C // H, side * side
            G, R, Nk = B, F * N, N
            q, kv = rnd((G * R, C), 1), rnd((G * Nk, 2 * C), 2)
            o = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
            sp = ops.AttnSpec(q=q, kv=kv, out=o, G=G, heads=H, R=R, Nk=Nk, d=d, dpad=((d + 63) // 64) * 64, ldq=C,
                              ldkv=2 * C, ldo=C, kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1 / _m.sqrt(d))
            us = timeit(lambda: be.attention(sp))
            fl, by = 4.0 * G * R * Nk * C, 2.0 * (2 * G * R * C + 2 * G * Nk * C)
            lines.append(f"| spatial (first-frame keys) | {side}x{side} | {C} | {d} | {us:.1f} | {fl / us / 1e6:.0f} | {by / us / 1e3:.0f} |")
            print(lines[-1], flush=True)
            del q, kv, o
    for Ft in (8, 12, 16, 24):
        for C, N in ((320, 1024), (640, 256), (1280, 64)):
            d = C // H
            qkv = rnd((B * Ft * N, 3 * C), 3)
            o = torch.zeros(B * Ft * N, C, dtype=torch.bfloat16, device=DEV)
            us = timeit(lambda: be.temporal_attention(qkv, o, B, Ft, N, H, d, 1 / _m.sqrt(d)))
            fl, by = 4.0 * B * N * Ft * Ft * C, 2.0 * 4 * B * Ft * N * C
            lines.append(f"| temporal | {Ft} frames x {N} px | {C} | {d} | {us:.1f} | {fl / us / 1e6:.1f} | {by / us / 1e3:.0f} |")
            print(lines[-1], flush=True)
    for Nk in (1, 25, 77, 229, 256):
        for C, N in ((320, 1024), (640, 256), (1280, 64)):
            d = C // H
            G, R = B * F, N
            q, kv = rnd((G * R, C), 4), rnd((G * Nk, 2 * C), 5)
            o = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
            sp = ops.AttnSpec(q=q, kv=kv, out=o, G=G, heads=H, R=R, Nk=Nk, d=d, dpad=((d + 63) // 64) * 64, ldq=C,
                              ldkv=2 * C, ldo=C, kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1 / _m.sqrt(d))
            us = timeit(lambda: be.attention(sp))
            fl, by = 4.0 * G * R * Nk * C, 2.0 * (2 * G * R * C + 2 * G * Nk * C)
            lines.append(f"| cross | {Nk} keys x {N} queries x {G} frames | {C} | {d} | {us:.1f} | {fl / us / 1e6:.1f} | {by / us / 1e3:.0f} |")
            print(lines[-1], flush=True)
    if out_path:
        with open(out_path, "w") as f:
            f.write("\n".join(lines) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true", help="the BASELINE.json config-5 attention microbenchmark sweep")
    ap.add_argument("--once", action="store_true", help="with --sweep: two eager launches per shape (for ncu)")
    ap.add_argument("--shapes", default=",".join(k for k in SHAPES if not k.endswith("_hr")))
    ap.add_argument("--single", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--form", type=int, default=0, help="asva_attn_desc.form: 0 auto, 1 tcgen05, 2 warp-MMA (Nk <= 128)")
    args = ap.parse_args()
    if args.sweep:
        sweep(args.out, args.once)
        return
    be = ops.backend()
    if args.single:
        s = make(args.single)
        for _ in range(3):
            be.attention(s)
        torch.cuda.synchronize()
        return
    lines = ["| shape | G | R | Nk | d | us | core TFLOP/s | Q+O+KV GB/s | rel-L2 vs sim |", "|---|---|---|---|---|---|---|---|---|"]
    for name in args.shapes.split(","):
        s = make(name)
        G, R, Nk, d = SHAPES[name]
        s.form = args.form if (args.form != 2 or Nk <= 128) else 0
        err = float("nan")
        if G * R * Nk * 8 <= 2 * 12288 * 1024 * 8:
            ref = torch.zeros_like(s.out)
            SimBackend().attention(dataclasses.replace(s, out=ref))
            be.attention(s)
            torch.cuda.synchronize()
            err = float((s.out.float() - ref.float()).norm() / ref.float().norm())
        be.attention(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                be.attention(s)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        fl = 4.0 * G * R * Nk * 8 * d
        by = 2.0 * (2 * G * R * 8 * d + 2 * G * Nk * 8 * d)
        lines.append(f"| {name} | {G} | {R} | {Nk} | {d} | {us:.1f} | {fl / us / 1e6:.0f} | {by / us / 1e3:.0f} | {err:.2e} |")
        print(lines[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
