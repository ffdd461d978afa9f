"""Plan sweep for asva_gemm on representative shapes of the headline workload: every (cta_group, block_n, split_k)
candidate is checked against the torch spec interpreter (tests/sim_backend.py) and timed from a CUDA graph.

    python tools/gemm_probe.py [--shapes conv0,lin0,...] [--out gpurun_out/gemm_probe.md]"""
import argparse
import dataclasses
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from asva_b200 import _lib, ops  # noqa: E402
from sim_backend import SimBackend  # noqa: E402

DEV = "cuda"


def rnd(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def conv(n_img, h, w, ci, co):
    x = rnd((n_img * h * w, ci), 1)
    wt = rnd((co, 9 * ci), 2, 1 / math.sqrt(9 * ci))
    b = rnd((co,), 3, dtype=torch.float32)
    return ops.spec_conv3x3(x, wt, torch.empty(n_img * h * w, co, dtype=torch.bfloat16, device=DEV), n_img=n_img, h=h,
                            wd=w, bias=b)


def lin(M, K, N, res=True, geglu=False):
    x = rnd((M, K), 4)
    w = rnd((N, K), 5, 1 / math.sqrt(K))
    b = rnd((N,), 6, dtype=torch.float32)
    out = torch.empty(M, N // 2 if geglu else N, dtype=torch.bfloat16, device=DEV)
    return ops.spec_linear(x, w, out, bias=b, res0=rnd((M, N), 7) if (res and not geglu) else None, geglu=geglu)


def tconv(B, F, N, C):
    y = rnd((B * F * N, C), 8)
    w4 = rnd((C, 4 * C), 9, 0.02)
    return ops.spec_tconv(y, w4, torch.empty(B * F * N, C, dtype=torch.bfloat16, device=DEV), B=B, F=F, N=N,
                          bias=rnd((C,), 10, dtype=torch.float32), tproj=rnd((B, C), 11, dtype=torch.float32),
                          tproj_ld=C, res1=rnd((B * F * N, C), 12))


SHAPES = {
    "conv0": lambda: conv(24, 32, 32, 320, 320),
    "conv0b": lambda: conv(24, 32, 32, 640, 320),
    "conv1": lambda: conv(24, 16, 16, 640, 640),
    "conv2": lambda: conv(24, 8, 8, 1280, 1280),
    "conv3": lambda: conv(24, 4, 4, 1280, 1280),
    "conv3b": lambda: conv(24, 4, 4, 2560, 1280),
    "lin0": lambda: lin(24576, 320, 320),
    "lin0p": lambda: lin(24576, 320, 320, res=False),
    "lin1": lambda: lin(6144, 640, 640),
    "lin2": lambda: lin(1536, 1280, 1280),
    "lin3": lambda: lin(384, 1280, 1280),
    "ff2_0": lambda: lin(24576, 1280, 320),
    "ff2_2": lambda: lin(1536, 5120, 1280),
    "qkv0": lambda: lin(24576, 320, 960, res=False),
    "geglu0": lambda: lin(24576, 320, 2560, geglu=True),
    "geglu1": lambda: lin(6144, 640, 5120, geglu=True),
    "geglu2": lambda: lin(1536, 1280, 10240, geglu=True),
    "tconv0": lambda: tconv(2, 12, 1024, 320),
    "tconv1": lambda: tconv(2, 12, 256, 640),
    "tconv2": lambda: tconv(2, 12, 64, 1280),
    "tconv3": lambda: tconv(2, 12, 16, 1280),
    "tg3": lambda: lin(384, 3840, 1280),          # the gathered conv_temp GEMM of level 3
    "ff2_3": lambda: lin(384, 5120, 1280),
}


def time_spec(be, spec, reps=10):
    be.gemm(spec)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            be.gemm(spec)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default=",".join(SHAPES))
    ap.add_argument("--out", default="")
    ap.add_argument("--splits", default="1,2,4,8")
    ap.add_argument("--timeplan", default="", help="cg,bn,split[,epilogue]: time just this plan")
    ap.add_argument("--single", default="", help="cg,bn,split[,epilogue]: launch just this plan 3 times eagerly (for ncu)")
    args = ap.parse_args()
    be = ops.backend()
    if args.timeplan:
        cg, bn, sp, epi = ([int(x) for x in args.timeplan.split(",")] + [0])[:4]
        for name in args.shapes.split(","):
            spec = SHAPES[name]()
            s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn, split_k=sp, cta_group=cg,
                                    epilogue=epi)
            print(f"{name} cg={cg} bn={bn} split={sp}: {time_spec(be, s):.1f} us")
        return
    if args.single:
        cg, bn, sp, epi = ([int(x) for x in args.single.split(",")] + [0])[:4]
        for name in args.shapes.split(","):
            spec = SHAPES[name]()
            s = dataclasses.replace(spec, out=torch.zeros_like(spec.out), block_n=bn, split_k=sp, cta_group=cg,
                                    epilogue=epi)
            for _ in range(3):
                be.gemm(s)
            torch.cuda.synchronize()
        return
    lines = []
    for name in args.shapes.split(","):
        spec = SHAPES[name]()
        ref = torch.zeros_like(spec.out)
        SimBackend().gemm(dataclasses.replace(spec, out=ref))
        fl = 2.0 * spec.M * spec.N * spec.K
        rows = []
        for cg in (1, 2):
            for bn in ((128,) if spec.geglu else (64, 128, 160, 256)):
                for sp in ((1,) if spec.geglu else [int(x) for x in args.splits.split(",")]):
                    if bn > 64 and bn >= 2 * spec.N:
                        continue
                    epis = (1, 3) if (spec.geglu or spec.out_fp32 or sp > 1) else (1, 2, 3)
                    if cg == 1 and sp in (2, 4, 8) and not spec.geglu and (bn // sp) % 32 == 0:
                        epis = epis + (4,)  # cluster split-K
                    for epi in epis:
                        out = torch.zeros_like(spec.out)
                        s = dataclasses.replace(spec, out=out, block_n=bn, split_k=sp, cta_group=cg, epilogue=epi)
                        try:
                            us = time_spec(be, s)
                        except _lib.AsvaError as e:
                            rows.append((cg, bn, sp, epi, None, str(e)[:60]))
                            continue
                        err = float((out.float() - ref.float()).norm() / ref.float().norm())
                        rows.append((cg, bn, sp, epi, us, err))
        auto = time_spec(be, dataclasses.replace(spec, out=torch.zeros_like(spec.out)))
        ok = [r for r in rows if r[4] is not None]
        best = min(ok, key=lambda r: r[4])
        b1 = min((r for r in ok if r[3] == 1), key=lambda r: r[4])
        b2 = min((r for r in ok if r[3] == 2), key=lambda r: r[4], default=None)
        b3 = min((r for r in ok if r[3] == 3), key=lambda r: r[4], default=None)
        b4 = min((r for r in ok if r[3] == 4), key=lambda r: r[4], default=None)
        lines.append(f"## {name}: M={spec.M} N={spec.N} K={spec.K} segs={len(spec.segs)} box={spec.box}  "
                     f"auto {auto:.1f} us; best cg={best[0]} bn={best[1]} split={best[2]} epi={best[3]} {best[4]:.1f} us "
                     f"({fl / best[4] / 1e6:.0f} TFLOP/s); best panel-epilogue {b1[4]:.1f} us, best per-warp "
                     f"{'-' if b2 is None else format(b2[4], '.1f')} us, best warp-TMA "
                     f"{'-' if b3 is None else format(b3[4], '.1f')} us, best cluster split-K "
                     f"{'-' if b4 is None else f'{b4[4]:.1f} us (bn {b4[1]} x {b4[2]})'}")
        lines.append("| cg | bn | split | epi | us | TFLOP/s | rel-L2 vs sim |")
        lines.append("|---|---|---|---|---|---|---|")
        for cg, bn, sp, epi, us, err in rows:
            if us is None:
                lines.append(f"| {cg} | {bn} | {sp} | {epi} | - | - | {err} |")
            else:
                flag = "" if err < 5e-3 else "  **BAD**"
                lines.append(f"| {cg} | {bn} | {sp} | {epi} | {us:.1f} | {fl / us / 1e6:.0f} | {err:.2e}{flag} |")
        lines.append("")
        print("\n".join(lines[-(len(rows) + 4):]), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
