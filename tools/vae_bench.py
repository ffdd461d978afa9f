"""VAE decode of one clip's frames on the B200 engine (asva_b200/vae.py): ms per clip, TFLOP/s against the measured
bf16 peak, and the restated torch decoder on this box's host cores on a bounded sample (1 frame).
    python tools/vae_bench.py [--frames 12] [--latent 32] [--reps 10]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asva_b200 import synth, vae  # noqa: E402
from oracle import vae_ref  # noqa: E402


def decode_flops(n, h, w, cfg=None):
    c = dict(vae_ref.DEFAULT_CONFIG)
    c.update(cfg or {})
    ch = list(reversed(c["block_out_channels"]))
    L = c["layers_per_block"] + 1
    fl = 0.0
    conv = lambda ci, co, px, k=3: 2.0 * k * k * ci * co * px * n  # noqa: E731
    px = h * w
    fl += conv(4, ch[0], px)
    fl += 2 * (2 * conv(ch[0], ch[0], px))                       # two mid resnets
    fl += 4 * 2.0 * ch[0] * ch[0] * px * n + 4.0 * px * px * ch[0] * n  # q,k,v,out projections + QK^T + PV
    prev = ch[0]
    for i, co in enumerate(ch):
        for j in range(L):
            ci = prev if j == 0 else co
            fl += conv(ci, co, px) + conv(co, co, px) + (conv(ci, co, px, 1) if ci != co else 0.0)
        if i < len(ch) - 1:
            px *= 4
            fl += conv(co, co, px)
        prev = co
    return fl + conv(ch[-1], 3, px)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu-frames", type=int, default=1)
    a = ap.parse_args()
    sd = synth.synth_state_dict(vae_ref.state_dict_shapes(), seed=3)
    z = torch.randn(a.frames, 4, a.latent, a.latent, generator=torch.Generator().manual_seed(1))
    eng = vae.VAEDecoderEngine(sd, device="cuda")
    zd = z.cuda()
    for _ in range(3):
        y = eng.decode(zd)
    torch.cuda.synchronize()
    n0 = eng.be.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        y = eng.decode(zd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    fl = decode_flops(a.frames, a.latent, a.latent)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:  # noqa: BLE001
        peak = 1400.0
    torch.set_num_threads(os.cpu_count())
    zc = z[: a.cpu_frames]
    with torch.no_grad():
        vae_ref.decode(sd, zc[:1, :, :8, :8])
        t0 = time.perf_counter()
        ref = vae_ref.decode(sd, zc)
        cpu_s = (time.perf_counter() - t0) / a.cpu_frames
    rel = float((y[: a.cpu_frames].cpu() - ref).norm() / ref.norm())
    print(json.dumps({"metric": "VAE decode ms per clip", "frames": a.frames, "latent": a.latent, "ms_per_clip": ms,
                      "gflop_per_clip": fl / 1e9, "tflops": fl / ms / 1e9, "frac_of_measured_bf16_sustained": fl / ms / 1e9 / peak,
                      "launches_per_clip": (eng.be.launches - n0) // a.reps,
                      "cpu_restatement_s_per_frame": cpu_s, "cpu_cores": os.cpu_count(),
                      "speedup_vs_cpu_per_frame": cpu_s * a.frames / (ms * 1e-3), "rel_l2_vs_cpu": rel}))


if __name__ == "__main__":
    main()
