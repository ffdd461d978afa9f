"""Times the fused GroupNorm and the LayerNorm on the shapes of the headline workload (CUDA-graph replay of 10 calls
over rotating buffers, CUDA events) and prints the achieved bandwidth on the algorithmic bytes (one read, one write).

    python tools/norm_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asva_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=10):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    be = ops.backend()
    if len(sys.argv) > 1 and sys.argv[1] == "--sweep":  # cluster-plan sweep (ASVA_GN_* are read at every call)
        shapes = [(2, 12288, 320, 0), (24, 1024, 320, 0), (2, 3072, 640, 0), (24, 256, 640, 0), (2, 768, 1280, 0),
                  (24, 64, 1280, 0), (2, 3072, 1280, 640), (2, 12288, 320, 320)]
        for n_inst, rows, C0, C1 in shapes:
            C = C0 + C1
            xs0 = [torch.randn(n_inst * rows, C0, device=DEV).bfloat16() for _ in range(4)]
            xs1 = [torch.randn(n_inst * rows, C1, device=DEV).bfloat16() if C1 else None for _ in range(4)]
            outs = [torch.empty(n_inst * rows, C, device=DEV, dtype=torch.bfloat16) for _ in range(4)]
            g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
            res = []
            for cs in (1, 2, 4, 8):
                for t in (128, 256, 512):
                    for k in (0, 16):
                        os.environ.update(ASVA_GN_CS=str(cs), ASVA_GN_T=str(t), ASVA_GN_KMAX=str(k))
                        try:
                            us = timeit(lambda i: be.groupnorm(xs0[i % 4], C0, xs1[i % 4], C1, n_inst, rows, 32, 1e-5,
                                                               g, b, True, outs[i % 4]))
                        except Exception as ex:  # noqa: BLE001
                            continue
                        res.append((us, cs, t, k))
            res.sort()
            print(f"{(n_inst, rows, C0, C1)}: " + "  ".join(f"cs{c}/T{t}/k{k}={u:.1f}" for u, c, t, k in res[:6]),
                  flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--forms":  # debug library: every GroupNorm form on every workload shape
        print("| n_inst | rows | C0 | C1 | default form | cluster us | barrier us | two-launch us |")
        print("|---|---|---|---|---|---|---|---|")
        shapes = [(2, 12288, 320, 0), (24, 1024, 320, 0), (2, 12288, 640, 320), (2, 12288, 320, 320), (2, 3072, 640, 0),
                  (24, 256, 640, 0), (2, 3072, 1280, 640), (2, 3072, 640, 640), (2, 3072, 640, 320), (2, 768, 1280, 0),
                  (24, 64, 1280, 0), (2, 768, 1280, 1280), (2, 768, 1280, 640), (2, 192, 1280, 0), (24, 16, 1280, 0),
                  (2, 192, 1280, 1280)]
        for n_inst, rows, C0, C1 in shapes:
            C = C0 + C1
            xs0 = [torch.randn(n_inst * rows, C0, device=DEV).bfloat16() for _ in range(4)]
            xs1 = [torch.randn(n_inst * rows, C1, device=DEV).bfloat16() if C1 else None for _ in range(4)]
            outs = [torch.empty(n_inst * rows, C, device=DEV, dtype=torch.bfloat16) for _ in range(4)]
            g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
            os.environ.pop("ASVA_GN_FORM", None)
            dflt = be.lib.asva_groupnorm_form(n_inst, rows, C, 32)
            cells = []
            for f in (0, 1, 2):
                os.environ["ASVA_GN_FORM"] = str(f)
                if be.lib.asva_groupnorm_form(n_inst, rows, C, 32) != f:
                    cells.append("-")
                    continue
                us = timeit(lambda i: be.groupnorm(xs0[i % 4], C0, xs1[i % 4], C1, n_inst, rows, 32, 1e-5, g, b, True,
                                                   outs[i % 4]))
                cells.append(f"{us:.1f}")
            os.environ.pop("ASVA_GN_FORM", None)
            print(f"| {n_inst} | {rows} | {C0} | {C1} | {dflt} | " + " | ".join(cells) + " |", flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--single":  # eager launches of a few GroupNorm shapes (for ncu)
        for n_inst, rows, C0, C1 in [(2, 12288, 320, 0), (24, 1024, 320, 0), (2, 3072, 640, 0), (2, 12288, 640, 320)]:
            x0 = torch.randn(n_inst * rows, C0, device=DEV).bfloat16()
            x1 = torch.randn(n_inst * rows, C1, device=DEV).bfloat16() if C1 else None
            o = torch.empty(n_inst * rows, C0 + C1, device=DEV, dtype=torch.bfloat16)
            g, b = torch.randn(C0 + C1, device=DEV), torch.randn(C0 + C1, device=DEV)
            for _ in range(2):
                be.groupnorm(x0, C0, x1, C1, n_inst, rows, 32, 1e-5, g, b, True, o)
            torch.cuda.synchronize()
        return
    print("| op | n_inst | rows | C0 | C1 | us | GB/s (1R+1W) |")
    print("|---|---|---|---|---|---|---|")
    gn = [(2, 12288, 320, 0), (24, 1024, 320, 0), (2, 12288, 640, 320), (2, 12288, 320, 320), (2, 3072, 640, 0),
          (24, 256, 640, 0), (2, 3072, 1280, 640), (2, 768, 1280, 0), (24, 64, 1280, 0), (2, 768, 1280, 1280),
          (2, 192, 1280, 0), (24, 16, 1280, 0), (2, 192, 1280, 1280)]
    NB = 4  # rotating buffers (> L2 for the big shapes only; the UNet itself hands GroupNorm L2-warm data)
    for n_inst, rows, C0, C1 in gn:
        C = C0 + C1
        xs0 = [torch.randn(n_inst * rows, C0, device=DEV).bfloat16() for _ in range(NB)]
        xs1 = [torch.randn(n_inst * rows, C1, device=DEV).bfloat16() if C1 else None for _ in range(NB)]
        outs = [torch.empty(n_inst * rows, C, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
        g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
        us = timeit(lambda i: be.groupnorm(xs0[i % NB], C0, xs1[i % NB], C1, n_inst, rows, 32, 1e-5, g, b, True,
                                           outs[i % NB]))
        by = n_inst * rows * C * 4
        print(f"| groupnorm | {n_inst} | {rows} | {C0} | {C1} | {us:.1f} | {by / us / 1e3:.0f} |", flush=True)
    for M, C in [(24576, 320), (6144, 640), (1536, 1280), (384, 1280), (12288 * 2, 320)]:
        xs = [torch.randn(M, C, device=DEV).bfloat16() for _ in range(NB)]
        outs = [torch.empty(M, C, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
        g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
        us = timeit(lambda i: be.layernorm(xs[i % NB], g, b, None, outs[i % NB], M, C, 1e-5, 1, 1))
        print(f"| layernorm | - | {M} | {C} | 0 | {us:.1f} | {M * C * 4 / us / 1e3:.0f} |", flush=True)


if __name__ == "__main__":
    main()
