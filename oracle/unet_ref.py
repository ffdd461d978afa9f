"""oracle/unet_ref.py — TEST INFRASTRUCTURE.  Clean-room CPU restatement (plain torch, fp32 by default) of ASVA's
audio-conditioned video UNet forward, driven by a reference-format state dict.  It travels to the GPU box (the
reference tree does not) and is the checker for the CUDA path; it is itself pinned against the reference's own
files executed through oracle/ref_loader.py (tests/test_host_cpu.py::test_oracle_vs_reference_live, tests/golden/).

Each function cites the reference lines it follows (paths relative to /root/reference/avgen/models/unets/).
The restatement is functional and applies the result-preserving hoists the product uses:
attn1 keys/values from frame 0 only, conv_temp split into head/prev/cur terms.  Contexts are taken per frame
(B,F,n,768) exactly as the reference does, so frame-varying contexts are also covered."""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

DEFAULT_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=8, cross_attention_dim=768, audio_cross_attention_dim=768, norm_num_groups=32,
    norm_eps=1e-5, flip_sin_to_cos=True, freq_shift=0, sample_size=64,
    down_block_types=("FFSpatioAudioTempCrossAttnDownBlock3D",) * 3 + ("FFSpatioTempResDownBlock3D",),
    mid_block_type="FFSpatioAudioTempCrossAttnUNetMidBlock3D",
    up_block_types=("FFSpatioTempResUpBlock3D",) + ("FFSpatioAudioTempCrossAttnUpBlock3D",) * 3,
)


def sinusoid(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, shift: float = 0.0) -> torch.Tensor:
    """diffusers get_timestep_embedding (embeddings.py), called at audio_cond_unet_3d_condition.py:673 and
    transformers/ff_spatio_audio_temp_transformer_3d.py:348.  fp32; [cos | sin] when flipped."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - shift))
    arg = t.float().view(-1, 1) * freq.view(1, -1)
    s, c = torch.sin(arg), torch.cos(arg)
    return torch.cat([c, s], 1) if flip_sin_to_cos else torch.cat([s, c], 1)


def conv2d(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], **kw) -> torch.Tensor:
    """F.conv2d; low-precision inputs are multiplied in fp32 and the result rounded back (what a bf16 conv with fp32
    accumulation computes) - torch's CPU bf16 convolution returns NaN on few-pixel images (4x2 -> 2x1, stride 2)."""
    if x.dtype == torch.float32:
        return F.conv2d(x, w, b, **kw)
    return F.conv2d(x.float(), w.float(), None if b is None else b.float(), **kw).to(x.dtype)


def lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def time_mlp(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers TimestepEmbedding: linear_1 -> SiLU -> linear_2.  The sinusoid is fp32 and is cast to the model
    dtype before the first linear (audio_cond_unet_3d_condition.py:679)."""
    x = x.to(sd[p + ".linear_1.weight"].dtype)
    return lin(sd, p + ".linear_2", F.silu(lin(sd, p + ".linear_1", x)))


def ff_conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """FFInflatedConv3d.forward, utils.py:34-57.  x: (B,C,F,h,w).  2-D conv per frame, then per pixel
    out_f = y_f + W_head y_0 + W_prev y_{max(f-1,0)} + W_cur y_f + b with W = conv_temp.weight (C, 3C)."""
    B, C, Fr, h, w = x.shape
    wt = sd[p + ".weight"]
    pad = (wt.shape[-1] - 1) // 2
    y = conv2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w), wt, sd[p + ".bias"], stride=stride, padding=pad)
    Co, ho, wo = y.shape[1:]
    y = y.view(B, Fr, Co, ho, wo).permute(0, 1, 3, 4, 2)  # (B,F,h,w,C)
    W = sd[p + ".conv_temp.weight"]
    Wh, Wp, Wc = W[:, :Co], W[:, Co:2 * Co], W[:, 2 * Co:]
    prev = torch.cat([y[:, :1], y[:, :-1]], dim=1)
    out = y + (y[:, :1] @ Wh.t()) + prev @ Wp.t() + y @ Wc.t() + sd[p + ".conv_temp.bias"]
    return out.permute(0, 4, 1, 2, 3)


def resblock(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """FFSpatioTempResnetBlock3D.forward, resnets/ff_spatio_temp_resnet_3d.py:161-191.  GroupNorm sees the 5-D
    tensor, i.e. statistics over (C/groups, F, h, w) per sample (:164,:175)."""
    h = F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = ff_conv(sd, p + ".conv1", F.silu(h))
    h = h + lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None, None]  # temb identical over f (:170-173)
    h = F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = ff_conv(sd, p + ".conv2", F.silu(h))
    if (p + ".conv_shortcut.weight") in sd:
        x = ff_conv(sd, p + ".conv_shortcut", x)
    return x + h  # output_scale_factor == 1 (:189)


def mha(q, k, v, heads: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v per head; q (G,Lq,C), k/v (G,Lk,C); mask bool (G,1|Lq,Lk), True = attend
    (F.scaled_dot_product_attention semantics, utils.py:151-153 / diffusers AttnProcessor2_0)."""
    G, Lq, C = q.shape
    d = C // heads
    qh = q.view(G, Lq, heads, d).transpose(1, 2)
    kh = k.view(G, -1, heads, d).transpose(1, 2)
    vh = v.view(G, -1, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if mask is not None:
        s = s.masked_fill(~mask.view(G, 1, mask.shape[-2], mask.shape[-1]), float("-inf"))
    return (torch.softmax(s, dim=-1) @ vh).transpose(1, 2).reshape(G, Lq, C)


def transformer(sd: SD, p: str, x: torch.Tensor, text: torch.Tensor, audio: torch.Tensor,
                audio_mask: Optional[torch.Tensor], heads: int, groups: int) -> torch.Tensor:
    """FFSpatioAudioTempTransformer3DModel.forward + BasicTransformerBlock.forward,
    transformers/ff_spatio_audio_temp_transformer_3d.py:94-158, 278-373.  x (B,C,F,h,w); text (B,F,77,768);
    audio (B,F,229,768); audio_mask bool (B,F,229)."""
    B, C, Fr, h, w = x.shape
    N = h * w
    xf = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w)
    t = F.group_norm(xf, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)  # per frame, eps 1e-6 (:61)
    t = conv2d(t, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    t = t.permute(0, 2, 3, 1).reshape(B, Fr, N, C)
    b = p + ".transformer_blocks.0"

    def ln(name, v):
        return F.layer_norm(v, (C,), sd[f"{b}.{name}.weight"], sd[f"{b}.{name}.bias"], 1e-5)

    # 1. first-frame spatial attention (utils.py:111-162): queries from every frame, keys/values of frame 0
    n1 = ln("norm1", t)
    q = lin(sd, b + ".attn1.to_q", n1).reshape(B, Fr * N, C)
    k0 = lin(sd, b + ".attn1.to_k", n1[:, 0])
    v0 = lin(sd, b + ".attn1.to_v", n1[:, 0])
    t = t + lin(sd, b + ".attn1.to_out.0", mha(q, k0, v0, heads)).view(B, Fr, N, C)
    # 2. audio cross-attention with the per-frame boolean key mask (:315-325)
    na = ln("norm_audio", t).reshape(B * Fr, N, C)
    au = audio.reshape(B * Fr, -1, audio.shape[-1])
    m = audio_mask.reshape(B * Fr, 1, -1) if audio_mask is not None else None
    o = mha(lin(sd, b + ".attn_audio.to_q", na), lin(sd, b + ".attn_audio.to_k", au),
            lin(sd, b + ".attn_audio.to_v", au), heads, m)
    t = t + lin(sd, b + ".attn_audio.to_out.0", o).view(B, Fr, N, C)
    # 3. text cross-attention (:328-341)
    n2 = ln("norm2", t).reshape(B * Fr, N, C)
    tx = text.reshape(B * Fr, -1, text.shape[-1])
    o = mha(lin(sd, b + ".attn2.to_q", n2), lin(sd, b + ".attn2.to_k", tx), lin(sd, b + ".attn2.to_v", tx), heads)
    t = t + lin(sd, b + ".attn2.to_out.0", o).view(B, Fr, N, C)
    # 4. temporal attention over the frame axis per pixel; pos enters the LayerNorm only (:346-358)
    pos = time_mlp(sd, b + ".pos_embedding_temp", sinusoid(torch.arange(Fr), C))  # (F,C)
    tt = t.permute(0, 2, 1, 3).reshape(B * N, Fr, C)
    nt = ln("norm_temp", tt + pos[None])
    o = mha(lin(sd, b + ".attn_temp.to_q", nt), lin(sd, b + ".attn_temp.to_k", nt),
            lin(sd, b + ".attn_temp.to_v", nt), heads)
    tt = tt + lin(sd, b + ".attn_temp.to_out.0", o)
    t = tt.view(B, N, Fr, C).permute(0, 2, 1, 3)
    # 5. GEGLU feed-forward (:361-371; diffusers FeedForward/GEGLU, erf gelu)
    u = lin(sd, b + ".ff.net.0.proj", ln("norm3", t))
    hval, gate = u.chunk(2, dim=-1)
    t = t + lin(sd, b + ".ff.net.2", hval * F.gelu(gate))
    # proj_out + residual (:142-153)
    o = t.reshape(B * Fr, h, w, C).permute(0, 3, 1, 2)
    o = conv2d(o, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + xf
    return o.view(B, Fr, C, h, w).permute(0, 2, 1, 3, 4)


def unet_forward(sd: SD, cfg: dict, sample: torch.Tensor, timestep, text: torch.Tensor, audio: torch.Tensor,
                 audio_mask: Optional[torch.Tensor], dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """AudioUNet3DConditionModel.forward, audio_cond_unet_3d_condition.py:598-798 with the block sequencing of
    unet_3d_blocks.py:285-302 (res down), :904-938 (attn down), :789-819 (mid), :354-369 (res up), :1019-1064
    (attn up).  sample (B,4,F,h,w) -> (B,4,F,h,w).
    dtype = torch.bfloat16 runs the same restatement with weights and activations in bfloat16 - "torch's own bf16
    eager" of this model, the yardstick the GPU tolerances are anchored to."""
    if dtype != torch.float32:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        sample, text, audio = sample.to(dtype), text.to(dtype), audio.to(dtype)
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    chans, groups, eps, heads = tuple(c["block_out_channels"]), c["norm_num_groups"], c["norm_eps"], c["attention_head_dim"]
    nlev = len(chans)
    B = sample.shape[0]
    t = torch.as_tensor(timestep, dtype=torch.float32).reshape(-1).expand(B)
    temb = time_mlp(sd, "time_embedding", sinusoid(t, chans[0], c["flip_sin_to_cos"], c["freq_shift"]))  # (B, 4*C0)
    x = ff_conv(sd, "conv_in", sample)
    skips = [x]
    for i in range(nlev):
        has_attn = "Attn" in c["down_block_types"][i]
        for j in range(c["layers_per_block"]):
            x = resblock(sd, f"down_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if has_attn:
                x = transformer(sd, f"down_blocks.{i}.attentions.{j}", x, text, audio, audio_mask, heads, groups)
            skips.append(x)
        if i < nlev - 1:
            x = ff_conv(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2)
            skips.append(x)
    x = resblock(sd, "mid_block.resnets.0", x, temb, groups, eps)
    x = transformer(sd, "mid_block.attentions.0", x, text, audio, audio_mask, heads, groups)
    x = resblock(sd, "mid_block.resnets.1", x, temb, groups, eps)
    for i in range(nlev):
        has_attn = "Attn" in c["up_block_types"][i]
        for j in range(c["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resblock(sd, f"up_blocks.{i}.resnets.{j}", x, temb, groups, eps)
            if has_attn:
                x = transformer(sd, f"up_blocks.{i}.attentions.{j}", x, text, audio, audio_mask, heads, groups)
        if i < nlev - 1:
            # F.interpolate(scale=[1,2,2], nearest) then 3x3 FFInflatedConv3d (ff_spatio_temp_resnet_3d.py:47,56)
            x = x.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4)
            x = ff_conv(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    x = F.group_norm(x, groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps)
    return ff_conv(sd, "conv_out", F.silu(x))
