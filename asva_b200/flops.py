"""Algorithmic FLOPs of one CFG denoising step (the numerator of bench.py's roofline), from the analytic model of
SURVEY.md section 8(d): reference work minus provably redundant work (attn1 K/V for frame 0 only; cross-attention
K/V and pos embeddings hoisted out of the loop; audio attention over valid keys only; conv_temp frame-0 term once
per clip).  MAC = 2 FLOP.  Config 2 (B=2, F=12, 32x32): 5 167.8 GFLOP; config 4 (B=2, F=24, 64x64): 45 632.6."""
import math
from typing import Dict


def step_flops(B: int, F: int, h: int, w: int, chans=(320, 640, 1280, 1280), layers: int = 2, cin: int = 4,
               cout: int = 4, n_text: int = 77, n_audio_valid: int = None, attn_levels=(True, True, True, False)
               ) -> Dict[str, float]:
    if n_audio_valid is None:
        n_audio_valid = 1 + 12 * int(math.ceil(19 / F))
    out = dict(conv3x3=0.0, conv1x1=0.0, conv_temp=0.0, proj_inout=0.0, attn1_proj=0.0, attn1_core=0.0,
               cross_proj=0.0, cross_core=0.0, temp_proj=0.0, temp_core=0.0, ff=0.0)
    BF = B * F

    def conv(ci, co, n, k=3):
        out["conv3x3" if k == 3 else "conv1x1"] += 2.0 * k * k * ci * co * n * BF
        out["conv_temp"] += 2.0 * co * co * n * (2 * BF + B)

    def res(ci, co, n):
        conv(ci, co, n)
        conv(co, co, n)
        if ci != co:
            conv(ci, co, n, k=1)

    def tr(c, n):
        out["proj_inout"] += 4.0 * c * c * n * BF
        out["attn1_proj"] += 2.0 * c * c * n * (2 * BF + 2 * B)
        out["attn1_core"] += 4.0 * n * n * c * BF
        out["cross_proj"] += 2 * 4.0 * c * c * n * BF
        out["cross_core"] += 4.0 * n * (n_text + n_audio_valid) * c * BF
        out["temp_proj"] += 8.0 * c * c * F * B * n
        out["temp_core"] += 4.0 * F * F * c * B * n
        out["ff"] += 24.0 * c * c * n * BF

    nlev = len(chans)
    n = h * w
    conv(cin, chans[0], n)
    skips = [chans[0]]
    c_prev = chans[0]
    for i, c in enumerate(chans):
        for j in range(layers):
            res(c_prev if j == 0 else c, c, n)
            if attn_levels[i]:
                tr(c, n)
            skips.append(c)
        c_prev = c
        if i < nlev - 1:
            n //= 4
            conv(c, c, n)
            skips.append(c)
    res(c_prev, c_prev, n)
    tr(c_prev, n)
    res(c_prev, c_prev, n)
    for i, c in enumerate(reversed(chans)):
        lvl = nlev - 1 - i
        for j in range(layers + 1):
            res(c_prev + skips.pop(), c, n)
            c_prev = c
            if attn_levels[lvl]:
                tr(c, n)
        if i < nlev - 1:
            n *= 4
            conv(c, c, n)
    conv(chans[0], cout, n)
    out["total"] = sum(out.values())
    out["gemm"] = out["total"] - out["attn1_core"] - out["cross_core"] - out["temp_core"]
    return out
