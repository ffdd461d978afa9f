"""Per-shape census of the asva_gemm launches of one denoising step of the headline workload.

Records every GemmSpec of one eager step, groups them by signature (M, N, K, taps, epilogue flags) and times each
group on the device (CUDA events, cycling through the group's instances so weights are not L2-resident by accident).
Prints a markdown table: launches/step, us/launch, TFLOP/s, share of the summed GEMM time.

    python tools/gemm_census.py [--workload cfg2] [--reps 10] [--out gpurun_out/gemm_census.md]"""
import argparse
import collections
import os
import sys

os.environ["ASVA_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from asva_b200 import schedulers, synth  # noqa: E402
from avgen.models.unets import AudioUNet3DConditionModel  # noqa: E402
from avgen.pipelines.pipeline_audio_cond_animation import AudioCondAnimationPipeline  # noqa: E402


def signature(s):
    taps = len(s.segs)
    epi = []
    if s.bias is not None:
        epi.append("b")
    if s.add is not None:
        epi.append("a")
    epi += ["r"] * sum(r is not None for r in s.res)
    if s.geglu:
        epi.append("G")
    if s.out_fp32:
        epi.append("f32")
    return (s.M, s.N, s.K, taps, tuple(s.trav), tuple(s.box), "".join(epi))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    F, h, w, _ = bench.WORKLOADS[args.workload]
    chans = bench.CHANS[args.workload]
    sd = bench._build_weights(chans)
    with torch.device("meta"):
        model = AudioUNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                          block_out_channels=chans)
    model.load_state_dict(sd, assign=True)
    model.to("cuda")
    pipe = AudioCondAnimationPipeline(None, None, model, schedulers.DDIMScheduler(), None, None)
    pipe.set_progress_bar_config(disable=True)
    lat, text, audio, mask = synth.synth_inputs(F=F, h=h, w=w, k=2)
    sess = pipe.open_session(text.cuda(), audio.cuda(), mask.cuda(), F, h, w, 50, audio_guidance_scale=4.0)
    be = model.engine().be
    sess.load_latents(lat.cuda())
    sess.step(0)
    recorded = []
    orig = be.gemm

    def rec(spec):
        recorded.append(spec)
        orig(spec)

    be.gemm = rec
    sess.step(1)
    be.gemm = orig
    torch.cuda.synchronize()

    groups = collections.OrderedDict()
    for s in recorded:
        groups.setdefault(signature(s), []).append(s)
    rows = []
    for sig, specs in groups.items():
        for s in specs[:2]:
            be.gemm(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()  # replay from a graph: the host (ctypes + tensor-map encode) is not in the timing
        n = 0
        with torch.cuda.graph(g):
            for _ in range(args.reps):
                for s in specs:
                    be.gemm(s)
                    n += 1
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        M, N, K = sig[0], sig[1], sig[2]
        fl = 2.0 * M * N * K
        # library yardstick: cuBLAS bf16 GEMM of the same M x N x K (plain operands, no epilogue)
        a_ = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        ws_ = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) for _ in range(min(len(specs), 4))]
        o_ = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for w_ in ws_:
            torch.matmul(a_, w_.t(), out=o_)
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        n2 = 0
        with torch.cuda.graph(g2):
            for _ in range(args.reps):
                for w_ in ws_:
                    torch.matmul(a_, w_.t(), out=o_)
                    n2 += 1
        g2.replay()
        torch.cuda.synchronize()
        e0.record()
        g2.replay()
        e1.record()
        torch.cuda.synchronize()
        lib_us = e0.elapsed_time(e1) * 1e3 / n2
        del a_, ws_, o_, g2
        rows.append((sig, len(specs), us, fl / us / 1e6, len(specs) * us, lib_us, be.gemm_plan(specs[0])))
    tot = sum(r[4] for r in rows)
    totfl = sum(2.0 * r[0][0] * r[0][1] * r[0][2] * r[1] for r in rows)
    lines = [f"# asva_gemm census, workload {args.workload}: {len(recorded)} launches/step, summed {tot / 1e3:.3f} ms, "
             f"{totfl / 1e9:.1f} GFLOP executed -> {totfl / tot / 1e6:.1f} TFLOP/s average (back-to-back launches replayed "
             f"from a CUDA graph, CUDA events)", "",
             "| M | N | K | taps | trav | box | epilogue | plan bn/split/cg/stages/epi | launches | us/launch | TFLOP/s | step us | share | cuBLAS us |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for sig, cnt, us, tf, tt, lib, plan in sorted(rows, key=lambda r: -r[4]):
        lines.append(f"| {sig[0]} | {sig[1]} | {sig[2]} | {sig[3]} | {sig[4]} | {sig[5]} | {sig[6]} | {'/'.join(str(x) for x in plan)} | {cnt} | {us:.1f} | "
                     f"{tf:.0f} | {tt:.0f} | {100 * tt / tot:.1f}% | {lib:.1f} |")
    lines.append(f"\ncuBLAS yardstick total for the same shapes: {sum(r[5] * r[1] for r in rows) / 1e3:.3f} ms")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
