"""SD-1.5 AutoencoderKL DECODER on the C-ABI kernels (include/asva_b200.h): the step right after the denoising loop,
`decode_latents` of the reference pipeline (/root/reference/avgen/pipelines/pipeline_audio_cond_animation.py:205-213,
:368-370 - all F frames of a clip go through `vae.decode` in one call).  SURVEY.md section 8(f), rank 1.

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
post_quant_conv is a 4x4 matrix on a 4-channel latent (16 MACs per pixel) and runs as one torch einsum on the device.
VAE.encode (one image per clip) stays the stock module."""
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


class VAEDecoderEngine:
    def __init__(self, sd: SD, cfg: Optional[dict] = None, device="cuda", backend=None, act_dtype=torch.bfloat16):
        c = dict(DEFAULT_CONFIG)
        c.update(cfg or {})
        self.cfg, self.dev, self.dt = c, torch.device(device), act_dtype
        self.be = backend if backend is not None else ops.backend()
        self.groups = c["norm_num_groups"]
        ch = list(reversed(c["block_out_channels"]))
        for v in ch:
            if v % 64 != 0:
                raise ValueError(f"block_out_channels must be multiples of 64 (got {v})")
        zc, oc = c["latent_channels"], c["out_channels"]
        if 9 * zc > 64 or oc > 8:
            raise ValueError("latent channels: 9*latent_channels must be <= 64 and out_channels <= 8")
        self.ch, self.zc, self.oc = ch, zc, oc
        f32 = lambda t: t.float().to(self.dev).contiguous()  # noqa: E731
        bf = lambda t: t.float().to(self.dev, self.dt).contiguous()  # noqa: E731
        self.pq_w = f32(sd["post_quant_conv.weight"].reshape(zc, zc))
        self.pq_b = f32(sd["post_quant_conv.bias"])
        w_in = sd["decoder.conv_in.weight"].float().permute(0, 2, 3, 1).reshape(ch[0], 9 * zc)
        self.in_w = bf(torch.nn.functional.pad(w_in, (0, 64 - 9 * zc)))
        self.in_b = f32(sd["decoder.conv_in.bias"])

        def res(p):
            r = dict(conv1=_Conv2d(sd, p + ".conv1", self.dev, self.dt), conv2=_Conv2d(sd, p + ".conv2", self.dev, self.dt),
                     g1=f32(sd[p + ".norm1.weight"]), b1=f32(sd[p + ".norm1.bias"]),
                     g2=f32(sd[p + ".norm2.weight"]), b2=f32(sd[p + ".norm2.bias"]), short=None)
            if (p + ".conv_shortcut.weight") in sd:
                r["short"] = _Conv2d(sd, p + ".conv_shortcut", self.dev, self.dt)
            return r

        self.mid = [res("decoder.mid_block.resnets.0"), res("decoder.mid_block.resnets.1")]
        a = "decoder.mid_block.attentions.0"
        C = ch[0]
        wq, wk, wv, wo = (sd[f"{a}.{n}.weight"].float() for n in ("to_q", "to_k", "to_v", "to_out.0"))
        bq, bk, bv, bo = (sd[f"{a}.{n}.bias"].float() for n in ("to_q", "to_k", "to_v", "to_out.0"))
        self.attn = dict(g=f32(sd[a + ".group_norm.weight"]), b=f32(sd[a + ".group_norm.bias"]),
                         qk_w=bf(torch.cat([wq, wk], 0)), qk_b=f32(torch.cat([bq, bk], 0)), v_w=bf(wv),
                         o_w=bf(wo), o_b=f32(bo + wo @ bv), C=C)
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
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._tuned = set()

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

    def _conv3(self, cv: _Conv2d, a, n, h, w, out, res=None):
        spec = ops.spec_conv3x3(a, cv.w, out, n_img=n, h=h, wd=w, bias=cv.b)
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

    # ------------------------------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z (n, 4, h, w), already divided by the scaling factor -> image (n, 3, 8h, 8w) fp32 (AutoencoderKL.decode)."""
        n, zc, h, w = z.shape
        assert zc == self.zc
        if (h * w) % 64 != 0:
            raise ValueError(f"latent h*w = {h * w} must be a multiple of 64 (the P V product runs as a tcgen05 GEMM)")
        be, ch = self.be, self.ch
        key = (n, h, w)
        tune = getattr(be, "name", "") == "cuda" and key not in self._tuned and not torch.cuda.is_current_stream_capturing()
        if tune:  # first decode of a geometry: every new GEMM shape gets its tile plan measured (cached per shape)
            self._tuned.add(key)
            be.tuning = True
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


class _DecoderOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class FastDecodeVAE(torch.nn.Module):
    """A stock AutoencoderKL with `decode` routed through VAEDecoderEngine.  Everything else (`encode`, `config`,
    `dtype`, parameters, `.to`) is the wrapped module's own, so the pipeline code does not change."""

    def __init__(self, vae: torch.nn.Module):
        super().__init__()
        self.inner = vae
        self._eng = None

    @property
    def config(self):
        return self.inner.config

    @property
    def dtype(self):
        return next(self.inner.parameters()).dtype

    def encode(self, *a, **k):
        return self.inner.encode(*a, **k)

    def _apply(self, fn, *a, **k):
        self._eng = None
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
