"""SD-1.5 AutoencoderKL on the C-ABI kernels (include/asva_b200.h): DECODE, the step right after the denoising loop
(`decode_latents` of the reference pipeline, /root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:205-213,
:368-370 - all F frames of a clip go through `vae.decode` in one call) and ENCODE of the conditioning image before it
(`encode_latents`, :199-203,309-310).  SURVEY.md section 8(f), rank 1.

The class being replaced lives in diffusers==0.29.2 (AutoencoderKL.decode -> Decoder: conv_in, UNetMidBlock2D with one
single-head 512-wide attention, four UpDecoderBlock2D, GroupNorm + SiLU + conv_out); this engine consumes its state
dict (`post_quant_conv.*`, `decoder.*`) unchanged and reuses the UNet path's kernels:
  * every 3x3 conv is the implicit-GEMM tcgen05 conv (asva_gemm, channels-last bf16, residual add in the epilogue);
  * GroupNorm(+SiLU) is the one-launch fused kernel (per image: n_inst = frames, rows = h*w); the nearest 2x upsample
    is the apply kernel's replicate mode feeding the conv directly;
  * the mid-block attention has ONE head of 512 channels - wider than asva_attention tiles - so it runs as GEMMs:
    S = Q K^T (asva_gemm, fp32 out), asva_softmax_rows, O = P V with V^T produced directly by a GEMM whose
    "activation" operand is W_v (out[c][token] = W_v x^T).  The v bias is folded into the output projection's bias
    (softmax rows sum to one: P (V + 1 b_v^T) = P V + b_v^T).
post_quant_conv / quant_conv are 4x4 / 8x8 matrices on a few channels (16 / 64 MACs per pixel) and run as one torch
einsum on the device.  The encoder's downsamplers pad right / bottom only (diffusers Downsample2D with padding 0): the
implicit-GEMM conv takes its nine taps at offsets 0..2 instead of -1..1 (ops.spec_conv3x3 pad_lo = 0)."""
import math
from typing import Dict, Optional, Tuple

import torch

from . import ops

SD = Dict[str, torch.Tensor]
DEFAULT_CONFIG = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4, out_channels=3,
                      norm_num_groups=32)


def is_autoencoder_kl_state_dict(sd) -> bool:
    return all(k in sd for k in ("post_quant_conv.weight", "decoder.conv_in.weight", "decoder.conv_out.weight",
                                 "decoder.mid_block.attentions.0.to_q.weight"))


class _Conv2d:
    """3x3 (or 1x1) conv weights as the K-major GEMM operand [Cout, k*k*Cin], K index = (ky*3 + kx)*Cin + c."""

    def __init__(self, sd: SD, p: str, dev, dt):
        w = sd[p + ".weight"].float()
        co, ci, kh, kw = w.shape
        self.cin, self.cout, self.k = ci, co, kh
        self.w = w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(dev, dt).contiguous()
        self.b = sd[p + ".bias"].float().to(dev).contiguous()


class _VAEBlocks:
    """What the decoder and the encoder share: weight packing of ResnetBlock2D / the single-head attention, buffers,
    and the launch sequences of GroupNorm(+SiLU), conv, resnet block and attention-as-GEMMs."""

    def __init__(self, cfg: Optional[dict], device, backend, act_dtype):
        c = dict(DEFAULT_CONFIG)
        c.update(cfg or {})
        self.cfg, self.dev, self.dt = c, torch.device(device), act_dtype
        self.be = backend if backend is not None else ops.backend()
        self.groups = c["norm_num_groups"]
        for v in c["block_out_channels"]:
            if v % 64 != 0:
                raise ValueError(f"block_out_channels must be multiples of 64 (got {v})")
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._tuned = set()

    def _f32(self, t):
        return t.float().to(self.dev).contiguous()

    def _bf(self, t):
        return t.float().to(self.dev, self.dt).contiguous()

    def _pack_res(self, sd: SD, p: str) -> dict:
        r = dict(conv1=_Conv2d(sd, p + ".conv1", self.dev, self.dt), conv2=_Conv2d(sd, p + ".conv2", self.dev, self.dt),
                 g1=self._f32(sd[p + ".norm1.weight"]), b1=self._f32(sd[p + ".norm1.bias"]),
                 g2=self._f32(sd[p + ".norm2.weight"]), b2=self._f32(sd[p + ".norm2.bias"]), short=None)
        if (p + ".conv_shortcut.weight") in sd:
            r["short"] = _Conv2d(sd, p + ".conv_shortcut", self.dev, self.dt)
        return r

    def _pack_attn(self, sd: SD, a: str, C: int) -> dict:
        wq, wk, wv, wo = (sd[f"{a}.{n}.weight"].float() for n in ("to_q", "to_k", "to_v", "to_out.0"))
        bq, bk, bv, bo = (sd[f"{a}.{n}.bias"].float() for n in ("to_q", "to_k", "to_v", "to_out.0"))
        return dict(g=self._f32(sd[a + ".group_norm.weight"]), b=self._f32(sd[a + ".group_norm.bias"]),
                    qk_w=self._bf(torch.cat([wq, wk], 0)), qk_b=self._f32(torch.cat([bq, bk], 0)), v_w=self._bf(wv),
                    o_w=self._bf(wo), o_b=self._f32(bo + wo @ bv), C=C)

    def _begin(self, key):
        be = self.be
        tune = getattr(be, "name", "") == "cuda" and key not in self._tuned and not torch.cuda.is_current_stream_capturing()
        if tune:  # first call of a geometry: every new GEMM shape gets its tile plan measured (cached per shape)
            self._tuned.add(key)
            be.tuning = True
        return tune

    # ------------------------------------------------------------------------------------------ helpers
    def buf(self, tag: str, shape: Tuple[int, ...], dtype=None) -> torch.Tensor:
        dtype = self.dt if dtype is None else dtype
        key = (tag, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._bufs[key] = t
        return t

    def _gn(self, x, C, n, rows, gamma, beta, silu, tag="gn"):
        out = self.buf(tag, (n * rows, C))
        self.be.groupnorm(x, C, None, 0, n, rows, self.groups, 1e-6, gamma, beta, silu, out)
        return out

    def _conv3(self, cv: _Conv2d, a, n, h, w, out, res=None, stride=1, pad_lo=1):
        spec = ops.spec_conv3x3(a, cv.w, out, n_img=n, h=h, wd=w, bias=cv.b, stride=stride, pad_lo=pad_lo)
        if res is not None:
            spec.res, spec.res_ld = [res, None], [res.stride(0), 0]
        self.be.gemm(spec)
        return out

    def _resblock(self, r, x, n, h, w, out_tag):
        M, cin, cout = n * h * w, r["conv1"].cin, r["conv1"].cout
        a = self._gn(x, cin, n, h * w, r["g1"], r["b1"], True)
        y = self._conv3(r["conv1"], a, n, h, w, self.buf("conv_y", (M, cout)))
        a2 = self._gn(y, cout, n, h * w, r["g2"], r["b2"], True)
        if r["short"] is not None:
            sc = self.buf("short", (M, cout))
            self.be.gemm(ops.spec_linear(x, r["short"].w, sc, bias=r["short"].b))
            x = sc
        return self._conv3(r["conv2"], a2, n, h, w, self.buf(out_tag, (M, cout)), res=x)

    def _attention(self, x, n, hw):
        be, at = self.be, self.attn
        C, M = at["C"], n * hw
        if hw % 64 != 0:
            raise ValueError(f"latent h*w = {hw} must be a multiple of 64 (the P V product runs as a tcgen05 GEMM)")
        g = self._gn(x, C, n, hw, at["g"], at["b"], False)
        qk = self.buf("qk", (M, 2 * C))
        be.gemm(ops.spec_linear(g, at["qk_w"], qk, bias=at["qk_b"]))
        vt = self.buf("vT", (C, M))  # V^T[c][token] = sum_k W_v[c][k] x[token][k]  (the v bias is folded into o_b)
        be.gemm(ops.spec_linear(at["v_w"], g, vt))
        s = self.buf("scores", (hw, hw), torch.float32)
        p = self.buf("probs", (hw, hw))
        o = self.buf("attn_o", (M, C))
        for i in range(n):
            rows = slice(i * hw, (i + 1) * hw)
            be.gemm(ops.spec_linear(qk[rows, :C], qk[rows, C:], s, out_fp32=True))
            be.softmax_rows(s, p, hw, hw, 1.0 / math.sqrt(C))
            be.gemm(ops.spec_linear(p, vt[:, rows], o[rows]))
        out = self.buf("attn_out", (M, C))
        be.gemm(ops.spec_linear(o, at["o_w"], out, bias=at["o_b"], res0=x))
        return out


class VAEDecoderEngine(_VAEBlocks):
    def __init__(self, sd: SD, cfg: Optional[dict] = None, device="cuda", backend=None, act_dtype=torch.bfloat16):
        super().__init__(cfg, device, backend, act_dtype)
        c = self.cfg
        ch = list(reversed(c["block_out_channels"]))
        zc, oc = c["latent_channels"], c["out_channels"]
        if 9 * zc > 64 or oc > 8:
            raise ValueError("latent channels: 9*latent_channels must be <= 64 and out_channels <= 8")
        self.ch, self.zc, self.oc = ch, zc, oc
        f32, bf = self._f32, self._bf
        self.pq_w = f32(sd["post_quant_conv.weight"].reshape(zc, zc))
        self.pq_b = f32(sd["post_quant_conv.bias"])
        w_in = sd["decoder.conv_in.weight"].float().permute(0, 2, 3, 1).reshape(ch[0], 9 * zc)
        self.in_w = bf(torch.nn.functional.pad(w_in, (0, 64 - 9 * zc)))
        self.in_b = f32(sd["decoder.conv_in.bias"])

        res = lambda p: self._pack_res(sd, p)  # noqa: E731
        self.mid = [res("decoder.mid_block.resnets.0"), res("decoder.mid_block.resnets.1")]
        self.attn = self._pack_attn(sd, "decoder.mid_block.attentions.0", ch[0])
        self.up = []
        L = c["layers_per_block"] + 1
        for i, co in enumerate(ch):
            blk = dict(res=[res(f"decoder.up_blocks.{i}.resnets.{j}") for j in range(L)], up=None)
            if i < len(ch) - 1:
                blk["up"] = _Conv2d(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", self.dev, self.dt)
            self.up.append(blk)
        self.out_g, self.out_b = f32(sd["decoder.conv_norm_out.weight"]), f32(sd["decoder.conv_norm_out.bias"])
        self.out_w = bf(torch.nn.functional.pad(
            sd["decoder.conv_out.weight"].float().permute(0, 2, 3, 1).reshape(oc, 9 * ch[-1]), (0, 0, 0, 8 - oc)))
        self.out_bias = f32(torch.nn.functional.pad(sd["decoder.conv_out.bias"].float(), (0, 8 - oc)))

    # ------------------------------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z (n, 4, h, w), already divided by the scaling factor -> image (n, 3, 8h, 8w) fp32 (AutoencoderKL.decode)."""
        n, zc, h, w = z.shape
        assert zc == self.zc
        be, ch = self.be, self.ch
        tune = self._begin((n, h, w))
        try:
            zq = torch.einsum("oc,nchw->nohw", self.pq_w, z.to(self.dev, torch.float32)) + self.pq_b.view(1, -1, 1, 1)
            col = self.buf("in_col", (n * h * w, 64))
            be.conv_in_im2col(zq.contiguous(), col, n, n, zc, 1, h, w)
            x = self.buf("x_in", (n * h * w, ch[0]))
            be.gemm(ops.spec_linear(col, self.in_w, x, bias=self.in_b))
            x = self._resblock(self.mid[0], x, n, h, w, "mid_a")
            x = self._attention(x, n, h * w)
            x = self._resblock(self.mid[1], x, n, h, w, "mid_b")
            flip = 0
            for blk in self.up:
                for r in blk["res"]:
                    flip ^= 1
                    x = self._resblock(r, x, n, h, w, "up_a" if flip else "up_b")
                if blk["up"] is not None:
                    C = x.shape[1]
                    u = self.buf("up_in", (n * 4 * h * w, C))
                    be.groupnorm_apply(x, C, None, 0, None, n, n, h, w, False, True, u)
                    h, w = 2 * h, 2 * w
                    flip ^= 1
                    x = self._conv3(blk["up"], u, n, h, w, self.buf("up_a" if flip else "up_b", (n * h * w, C)))
            a = self._gn(x, ch[-1], n, h * w, self.out_g, self.out_b, True)
            yo = self.buf("out_y", (n * h * w, 8), torch.float32)
            be.gemm(ops.spec_conv3x3(a, self.out_w, yo, n_img=n, h=h, wd=w, bias=self.out_bias, out_fp32=True))
        finally:
            if tune:
                be.tuning = False
        return yo.view(n, h, w, 8)[..., : self.oc].permute(0, 3, 1, 2).contiguous()


class VAEEncoderEngine(_VAEBlocks):
    """AutoencoderKL.encode up to the distribution parameters (`encoder.*`, `quant_conv.*`): conv_in, four
    DownEncoderBlock2D (two resnets each; the downsampler is a stride-2 conv padded on the right / bottom only), mid
    block with the single-head attention, GroupNorm + SiLU + conv_out, quant_conv.  One conditioning image per clip
    (`encode_latents`, pipeline :199-203,309-310) - small, but it completes the VAE on the kernels of this library."""

    def __init__(self, sd: SD, cfg: Optional[dict] = None, device="cuda", backend=None, act_dtype=torch.bfloat16):
        super().__init__(cfg, device, backend, act_dtype)
        c = self.cfg
        ch = list(c["block_out_channels"])
        self.ch, self.zc = ch, c["latent_channels"]
        if 2 * self.zc > 8:
            raise ValueError("2 * latent_channels must be <= 8")
        w_in = sd["encoder.conv_in.weight"].float()
        self.cin = w_in.shape[1]
        if 9 * self.cin > 64:
            raise ValueError("image channels: 9 * in_channels must be <= 64")
        self.in_w = self._bf(torch.nn.functional.pad(w_in.permute(0, 2, 3, 1).reshape(ch[0], 9 * self.cin),
                                                     (0, 64 - 9 * self.cin)))
        self.in_b = self._f32(sd["encoder.conv_in.bias"])
        self.down = []
        for i in range(len(ch)):
            blk = dict(res=[self._pack_res(sd, f"encoder.down_blocks.{i}.resnets.{j}") for j in range(c["layers_per_block"])],
                       down=None)
            if i < len(ch) - 1:
                blk["down"] = _Conv2d(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", self.dev, self.dt)
            self.down.append(blk)
        self.mid = [self._pack_res(sd, "encoder.mid_block.resnets.0"), self._pack_res(sd, "encoder.mid_block.resnets.1")]
        self.attn = self._pack_attn(sd, "encoder.mid_block.attentions.0", ch[-1])
        self.out_g, self.out_b = self._f32(sd["encoder.conv_norm_out.weight"]), self._f32(sd["encoder.conv_norm_out.bias"])
        co = 2 * self.zc
        self.out_w = self._bf(torch.nn.functional.pad(
            sd["encoder.conv_out.weight"].float().permute(0, 2, 3, 1).reshape(co, 9 * ch[-1]), (0, 0, 0, 8 - co)))
        self.out_bias = self._f32(torch.nn.functional.pad(sd["encoder.conv_out.bias"].float(), (0, 8 - co)))
        self.q_w = self._f32(sd["quant_conv.weight"].reshape(co, co))
        self.q_b = self._f32(sd["quant_conv.bias"])

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """x (n, 3, H, W) in [-1, 1], H and W multiples of 8 -> moments (n, 2*latent, H/8, W/8) fp32 = [mean | logvar]."""
        n, ci, h, w = x.shape
        assert ci == self.cin
        down = 1 << (len(self.ch) - 1)
        if h % down or w % down:
            raise ValueError(f"image {h}x{w} must be divisible by {down}")
        be = self.be
        tune = self._begin((n, h, w))
        try:
            col = self.buf("in_col", (n * h * w, 64))
            be.conv_in_im2col(x.to(self.dev, torch.float32).contiguous(), col, n, n, ci, 1, h, w)
            t = self.buf("x_in", (n * h * w, self.ch[0]))
            be.gemm(ops.spec_linear(col, self.in_w, t, bias=self.in_b))
            flip = 0
            for blk in self.down:
                for r in blk["res"]:
                    flip ^= 1
                    t = self._resblock(r, t, n, h, w, "dn_a" if flip else "dn_b")
                if blk["down"] is not None:
                    C = t.shape[1]
                    flip ^= 1
                    t = self._conv3(blk["down"], t, n, h, w, self.buf("dn_a" if flip else "dn_b", (n * (h // 2) * (w // 2), C)),
                                    stride=2, pad_lo=0)
                    h, w = h // 2, w // 2
            t = self._resblock(self.mid[0], t, n, h, w, "mid_a")
            t = self._attention(t, n, h * w)
            t = self._resblock(self.mid[1], t, n, h, w, "mid_b")
            a = self._gn(t, self.ch[-1], n, h * w, self.out_g, self.out_b, True)
            yo = self.buf("out_y", (n * h * w, 8), torch.float32)
            be.gemm(ops.spec_conv3x3(a, self.out_w, yo, n_img=n, h=h, wd=w, bias=self.out_bias, out_fp32=True))
        finally:
            if tune:
                be.tuning = False
        co = 2 * self.zc
        m = yo.view(n, h, w, 8)[..., :co]
        return (torch.einsum("oc,nhwc->nohw", self.q_w, m) + self.q_b.view(1, -1, 1, 1)).contiguous()


class DiagonalGaussian:
    """diffusers DiagonalGaussianDistribution over the encoder's moments: .sample(generator) / .mode() / .mean / .std."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _EncoderOutput:
    def __init__(self, dist):
        self.latent_dist = dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


class _DecoderOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class FastDecodeVAE(torch.nn.Module):
    """A stock AutoencoderKL with `decode` (and `encode`, when the module carries the encoder keys) routed through the
    engines above.  `config`, `dtype`, parameters and `.to` are the wrapped module's own, so the pipeline code does
    not change."""

    def __init__(self, vae: torch.nn.Module):
        super().__init__()
        self.inner = vae
        self._eng = None
        self._enc, self._enc_checked = None, False

    @property
    def config(self):
        return self.inner.config

    @property
    def dtype(self):
        return next(self.inner.parameters()).dtype

    def encode(self, x, return_dict: bool = True, **kw):
        """The encoder half runs on the engine too when the module carries diffusers' encoder keys; else its own."""
        if self._enc is None and not self._enc_checked:
            self._enc_checked = True
            sd = self.inner.state_dict()
            dev = next(self.inner.parameters()).device
            if dev.type == "cuda" and "encoder.conv_in.weight" in sd and "quant_conv.weight" in sd:
                cfg = {k: getattr(self.inner.config, k) for k in DEFAULT_CONFIG if hasattr(self.inner.config, k)}
                with torch.cuda.device(dev):
                    self._enc = VAEEncoderEngine(sd, cfg, device=dev)
        if self._enc is None:
            return self.inner.encode(x, **kw)
        with torch.cuda.device(self._enc.dev):
            out = _EncoderOutput(DiagonalGaussian(self._enc.encode_moments(x).to(x.dtype)))
        return out if return_dict else (out.latent_dist,)

    def _apply(self, fn, *a, **k):
        self._eng = None
        self._enc, self._enc_checked = None, False
        return super()._apply(fn, *a, **k)

    def engine(self) -> VAEDecoderEngine:
        if self._eng is None:
            dev = next(self.inner.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("FastDecodeVAE decodes on the CUDA engine only: move the VAE to a CUDA device")
            cfg = {k: getattr(self.inner.config, k) for k in DEFAULT_CONFIG if hasattr(self.inner.config, k)}
            with torch.cuda.device(dev):
                self._eng = VAEDecoderEngine(self.inner.state_dict(), cfg, device=dev)
        return self._eng

    @torch.no_grad()
    def decode(self, z, return_dict: bool = True, **kw):
        eng = self.engine()
        with torch.cuda.device(eng.dev):
            img = eng.decode(z).to(z.dtype)
        return _DecoderOutput(img) if return_dict else (img,)


def wrap_vae(vae):
    """AutoencoderKL-shaped module (diffusers state-dict keys) -> FastDecodeVAE; anything else is returned unchanged."""
    if isinstance(vae, torch.nn.Module) and not isinstance(vae, FastDecodeVAE) and is_autoencoder_kl_state_dict(vae.state_dict()):
        return FastDecodeVAE(vae)
    return vae
