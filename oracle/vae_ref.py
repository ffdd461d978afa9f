"""oracle/vae_ref.py - TEST INFRASTRUCTURE.  CPU restatement (plain torch, fp32) of the SD-1.5 AutoencoderKL DECODER
that the reference pipeline calls after the denoising loop
(/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:205-213 `decode_latents`, :368-370), driven by a
diffusers-format state dict (`post_quant_conv.*`, `decoder.*`).

The class lives in the un-vendored dependency diffusers==0.29.2 (requirements.txt:2; models/autoencoders/
autoencoder_kl.py `AutoencoderKL.decode`, vae.py `Decoder`, unets/unet_2d_blocks.py `UNetMidBlock2D` / `UpDecoderBlock2D`,
resnet.py `ResnetBlock2D`, upsampling.py `Upsample2D`, attention_processor.py `Attention` with
`_from_deprecated_attn_block`).  diffusers cannot be installed here, so this restates its published algorithm for the
SD-1.5 VAE config (block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4, norm_num_groups 32,
act_fn silu, one single-head 512-wide attention in the mid block, eps 1e-6): PARITY UNPINNED against the library itself;
what is pinned is the CUDA decoder / encoder against this restatement.  `encode_moments` restates the ENCODER half
(`encode_latents`, pipeline :199-203,309-310)."""
import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
DEFAULT_CONFIG = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, out_channels=3,
                      norm_num_groups=32, scaling_factor=0.18215)


def state_dict_shapes(cfg: dict = None) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of the decoder half of a diffusers AutoencoderKL checkpoint (vae/diffusion_pytorch_model.*)."""
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    ch = list(reversed(c["block_out_channels"]))  # 512, 512, 256, 128
    L = c["layers_per_block"] + 1
    zc, oc = c["latent_channels"], c["out_channels"]
    out = [("post_quant_conv.weight", (zc, zc, 1, 1)), ("post_quant_conv.bias", (zc,)),
           ("decoder.conv_in.weight", (ch[0], zc, 3, 3)), ("decoder.conv_in.bias", (ch[0],))]

    def res(p, ci, co):
        out.extend([(p + ".norm1.weight", (ci,)), (p + ".norm1.bias", (ci,)),
                    (p + ".conv1.weight", (co, ci, 3, 3)), (p + ".conv1.bias", (co,)),
                    (p + ".norm2.weight", (co,)), (p + ".norm2.bias", (co,)),
                    (p + ".conv2.weight", (co, co, 3, 3)), (p + ".conv2.bias", (co,))])
        if ci != co:
            out.extend([(p + ".conv_shortcut.weight", (co, ci, 1, 1)), (p + ".conv_shortcut.bias", (co,))])

    res("decoder.mid_block.resnets.0", ch[0], ch[0])
    a = "decoder.mid_block.attentions.0"
    out.extend([(a + ".group_norm.weight", (ch[0],)), (a + ".group_norm.bias", (ch[0],))])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out.extend([(f"{a}.{n}.weight", (ch[0], ch[0])), (f"{a}.{n}.bias", (ch[0],))])
    res("decoder.mid_block.resnets.1", ch[0], ch[0])
    prev = ch[0]
    for i, co in enumerate(ch):
        for j in range(L):
            res(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if i < len(ch) - 1:
            out.extend([(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (co, co, 3, 3)),
                        (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (co,))])
        prev = co
    out.extend([("decoder.conv_norm_out.weight", (ch[-1],)), ("decoder.conv_norm_out.bias", (ch[-1],)),
                ("decoder.conv_out.weight", (oc, ch[-1], 3, 3)), ("decoder.conv_out.bias", (oc,))])
    return out


def encoder_state_dict_shapes(cfg: dict = None) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of the encoder half (`encoder.*`, `quant_conv.*`) of a diffusers AutoencoderKL checkpoint."""
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    ch = list(c["block_out_channels"])  # 128, 256, 512, 512
    L = c["layers_per_block"]
    zc = c["latent_channels"]
    out = [("encoder.conv_in.weight", (ch[0], 3, 3, 3)), ("encoder.conv_in.bias", (ch[0],))]

    def res(p, ci, co):
        out.extend([(p + ".norm1.weight", (ci,)), (p + ".norm1.bias", (ci,)),
                    (p + ".conv1.weight", (co, ci, 3, 3)), (p + ".conv1.bias", (co,)),
                    (p + ".norm2.weight", (co,)), (p + ".norm2.bias", (co,)),
                    (p + ".conv2.weight", (co, co, 3, 3)), (p + ".conv2.bias", (co,))])
        if ci != co:
            out.extend([(p + ".conv_shortcut.weight", (co, ci, 1, 1)), (p + ".conv_shortcut.bias", (co,))])

    prev = ch[0]
    for i, co in enumerate(ch):
        for j in range(L):
            res(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if i < len(ch) - 1:
            out.extend([(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (co, co, 3, 3)),
                        (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (co,))])
        prev = co
    res("encoder.mid_block.resnets.0", ch[-1], ch[-1])
    a = "encoder.mid_block.attentions.0"
    out.extend([(a + ".group_norm.weight", (ch[-1],)), (a + ".group_norm.bias", (ch[-1],))])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out.extend([(f"{a}.{n}.weight", (ch[-1], ch[-1])), (f"{a}.{n}.bias", (ch[-1],))])
    res("encoder.mid_block.resnets.1", ch[-1], ch[-1])
    out.extend([("encoder.conv_norm_out.weight", (ch[-1],)), ("encoder.conv_norm_out.bias", (ch[-1],)),
                ("encoder.conv_out.weight", (2 * zc, ch[-1], 3, 3)), ("encoder.conv_out.bias", (2 * zc,)),
                ("quant_conv.weight", (2 * zc, 2 * zc, 1, 1)), ("quant_conv.bias", (2 * zc,))])
    return out


def _conv(x, w, b, **kw):
    """F.conv2d; low-precision inputs are multiplied in fp32 and the result rounded back (a bf16 conv with fp32
    accumulation; torch's CPU bf16 convolution is unreliable on few-pixel images, see oracle/unet_ref.py)."""
    if x.dtype == torch.float32:
        return F.conv2d(x, w, b, **kw)
    return F.conv2d(x.float(), w.float(), b.float(), **kw).to(x.dtype)


def resnet(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """ResnetBlock2D.forward with temb=None, eps 1e-6, output_scale_factor 1 (resnet.py)."""
    h = F.silu(F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6))
    h = _conv(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6))
    h = _conv(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def mid_attention(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """Attention(heads=1, dim_head=C, residual_connection=True, norm_num_groups=groups, bias=True) on a 4-D input
    (attention_processor.py AttnProcessor2_0): GroupNorm over (b, c, hw) -> q/k/v linears -> softmax(q k^T / sqrt(C)) v
    -> to_out[0] -> + residual."""
    b, c, h, w = x.shape
    t = F.group_norm(x.view(b, c, h * w), groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6)
    t = t.transpose(1, 2)  # (b, hw, c)
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1) @ v
    o = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, h, w)


def decode(sd: SD, z: torch.Tensor, cfg: dict = None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """AutoencoderKL.decode(z).sample: z (n, 4, h, w) already divided by scaling_factor -> (n, 3, 8h, 8w).
    dtype = torch.bfloat16: the same restatement with weights and activations in bfloat16 (the tolerance anchor)."""
    if dtype != torch.float32:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        z = z.to(dtype)
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    g = c["norm_num_groups"]
    n_up = len(c["block_out_channels"])
    x = _conv(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = _conv(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = resnet(sd, "decoder.mid_block.resnets.0", x, g)
    x = mid_attention(sd, "decoder.mid_block.attentions.0", x, g)
    x = resnet(sd, "decoder.mid_block.resnets.1", x, g)
    for i in range(n_up):
        for j in range(c["layers_per_block"] + 1):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, g)
        if i < n_up - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(x, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(F.group_norm(x, g, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return _conv(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def encode_moments(sd: SD, x: torch.Tensor, cfg: dict = None, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """AutoencoderKL.encode(x).latent_dist.parameters: x (n, 3, H, W) in [-1, 1] -> moments (n, 2*latent, H/8, W/8) =
    [mean | logvar] (vae.py `Encoder.forward` + `quant_conv`).  Downsample2D of the VAE pads (0, 1, 0, 1) and convolves
    with stride 2 and NO padding (downsampling.py, `padding=0` for DownEncoderBlock2D).  The latent the pipeline uses is
    mean + exp(0.5 * clamp(logvar, -30, 20)) * randn  (DiagonalGaussianDistribution.sample)."""
    if dtype != torch.float32:
        sd = {k: v.to(dtype) for k, v in sd.items()}
        x = x.to(dtype)
    c = dict(DEFAULT_CONFIG)
    c.update(cfg or {})
    g = c["norm_num_groups"]
    n_down = len(c["block_out_channels"])
    x = _conv(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(n_down):
        for j in range(c["layers_per_block"]):
            x = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", x, g)
        if i < n_down - 1:
            x = F.pad(x, (0, 1, 0, 1))
            x = _conv(x, sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                      sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    x = resnet(sd, "encoder.mid_block.resnets.0", x, g)
    x = mid_attention(sd, "encoder.mid_block.attentions.0", x, g)
    x = resnet(sd, "encoder.mid_block.resnets.1", x, g)
    x = F.silu(F.group_norm(x, g, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], 1e-6))
    x = _conv(x, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return _conv(x, sd["quant_conv.weight"], sd["quant_conv.bias"])
