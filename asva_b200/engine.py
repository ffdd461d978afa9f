"""UNet denoising engine: one forward of ASVA's audio-conditioned video UNet as a fixed sequence of C-ABI kernel
launches (include/asva_b200.h) over channels-last bf16 activations  x[b][f][y][x][c].

Replaces AudioUNet3DConditionModel.forward (/root/reference/avgen/models/unets/audio_cond_unet_3d_condition.py:598-798)
and everything below it.  Dataflow per block follows SURVEY.md Appendix B:
  * FFInflatedConv3d (utils.py:22-57)      = implicit-GEMM conv  ->  frame-0 "head" GEMM  ->  temporal 2-tap GEMM
                                             whose epilogue adds head term, time-embedding projection and residual
  * FFSpatioTempResnetBlock3D (:161-191)   = GN stats / GN+SiLU apply / conv / temporal conv  (x2) + shortcut
  * BasicTransformerBlock (:278-373)       = LN -> Q GEMM (head-split epilogue) -> tcgen05 flash attention -> out GEMM
                                             (+residual) for attn1 (keys/values of frame 0 only), attn_audio (masked),
                                             attn2; LN(+pos) -> QKV GEMM -> per-pixel F x F core -> out GEMM;
                                             LN -> GEGLU GEMM -> out GEMM
Cross-attention keys/values depend only on the clip's text/audio context, so `set_context` projects them once per
clip; `pos_embedding_temp` depends only on F and is built in `prepare`.  All buffers are allocated in `prepare`, so
`forward` performs no allocation and no host<->device traffic and can be captured in a CUDA graph."""
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

SD = Dict[str, torch.Tensor]

DEFAULT_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=8, cross_attention_dim=768, audio_cross_attention_dim=768, norm_num_groups=32,
    norm_eps=1e-5, flip_sin_to_cos=True, freq_shift=0, sample_size=64,
    down_block_types=("FFSpatioAudioTempCrossAttnDownBlock3D",) * 3 + ("FFSpatioTempResDownBlock3D",),
    mid_block_type="FFSpatioAudioTempCrossAttnUNetMidBlock3D",
    up_block_types=("FFSpatioTempResUpBlock3D",) + ("FFSpatioAudioTempCrossAttnUpBlock3D",) * 3,
)


def check_supported(cfg: dict) -> None:
    """The kernels cover the architecture family the reference ships (configs/audio-cond_animation/*.yaml:30-32);
    anything else fails loudly instead of silently computing something different."""
    for c in cfg["block_out_channels"]:
        if c % 64 != 0:
            raise ValueError(f"block_out_channels must be multiples of 64 (got {c})")
        if (c // cfg["attention_head_dim"]) % 8 != 0:
            raise ValueError(f"head dim {c // cfg['attention_head_dim']} must be a multiple of 8")
    if cfg["cross_attention_dim"] % 64 or cfg["audio_cross_attention_dim"] % 64:
        raise ValueError("context dims must be multiples of 64")
    if cfg["in_channels"] * 9 > 64 or cfg["out_channels"] > 8:
        raise ValueError("latent channels: in*9 must be <= 64 and out <= 8")
    if cfg["freq_shift"] != 0 or not cfg["flip_sin_to_cos"]:
        raise ValueError("only flip_sin_to_cos=True, freq_shift=0 timestep features are implemented")
    for t in tuple(cfg["down_block_types"]) + tuple(cfg["up_block_types"]) + (cfg["mid_block_type"],):
        if t not in ("FFSpatioAudioTempCrossAttnDownBlock3D", "FFSpatioTempResDownBlock3D",
                     "FFSpatioAudioTempCrossAttnUNetMidBlock3D", "FFSpatioTempResUpBlock3D",
                     "FFSpatioAudioTempCrossAttnUpBlock3D"):
            raise ValueError(f"unsupported block type {t}")


class _Conv:
    """Packed FFInflatedConv3d: spatial weight [Cout, k*k*Cin] (K index = (ky*3+kx)*Cin + c) and the temporal
    weights W4 = [Wc | Wp | Wh | Wh + Wp] (current / previous / first frame; the last block serves frame 0, whose
    "previous frame" is itself - see ops.spec_tconv)."""

    def __init__(self, sd: SD, p: str, dev, dt=torch.bfloat16):
        w = sd[p + ".weight"].float()
        co, ci, kh, kw = w.shape
        self.cin, self.cout, self.k = ci, co, kh
        self.w = w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(dev, dt).contiguous()
        self.b = sd[p + ".bias"].float().to(dev).contiguous()
        wt = sd[p + ".conv_temp.weight"].float()
        wh, wp, wc = wt[:, :co], wt[:, co:2 * co], wt[:, 2 * co:]
        self.wt_full = wt.to(dev).contiguous()
        self.bt = sd[p + ".conv_temp.bias"].float().to(dev).contiguous()
        if co % 64 == 0:
            self.w4 = torch.cat([wc, wp, wh, wh + wp], dim=1).to(dev, dt).contiguous()


class UNetEngine:
    def __init__(self, sd: SD, cfg: Optional[dict] = None, device="cuda", backend=None, act_dtype=torch.bfloat16):
        """act_dtype is bfloat16 in production (the only type the CUDA kernels take); tests run the same launch
        sequence through the torch interpreter in float32 to check the host logic exactly against the oracle."""
        self.dt = act_dtype
        c = dict(DEFAULT_CONFIG)
        c.update(cfg or {})
        check_supported(c)
        self.cfg = c
        self.dev = torch.device(device)
        self.be = backend if backend is not None else ops.backend()
        self.chans = tuple(c["block_out_channels"])
        self.heads = c["attention_head_dim"]
        self.groups = c["norm_num_groups"]
        self.eps = float(c["norm_eps"])
        self.shape = None
        self.ctx, self.ctx_sig = None, None
        # Buffer generation: bumped whenever prepare() / set_context() (re)allocates device buffers.  Anything that
        # captured device addresses (CUDA graphs of the launch sequence) keys itself on it and re-captures on change.
        self.gen = 0
        # LayerNorm fold (ops.LnFold): off by default - measured slower than the LayerNorm passes it removes (DESIGN.md
        # section 5); ASVA_LN_FOLD=1 turns it on for A/B runs
        self.fold_ln = os.environ.get("ASVA_LN_FOLD", "0") == "1"
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._ctx_bufs: Dict[tuple, torch.Tensor] = {}
        self._pack(sd)

    # ------------------------------------------------------------------------------------------ weights
    def _f32(self, t):
        return t.float().to(self.dev).contiguous()

    def _bf(self, t):
        return t.float().to(self.dev, self.dt).contiguous()

    def _fold_ln(self, w: torch.Tensor, bias: Optional[torch.Tensor], norm) -> tuple:
        """LayerNorm(x; gamma, beta) W^T + bias = rstd * (x (W gamma)^T - mean * wsum) + (W beta + bias):
        -> (W gamma in the activation dtype, wsum of exactly those rounded values, folded bias), see ops.LnFold."""
        gamma, beta = norm
        w = w.float().to(self.dev)
        wg = (w * gamma.view(1, -1)).to(self.dt).contiguous()
        fb = w @ beta
        if bias is not None:
            fb = fb + bias.float().to(self.dev)
        return wg, wg.float().sum(1).contiguous(), fb.contiguous()

    def _pack_res(self, sd: SD, p: str) -> dict:
        r = dict(name=p, conv1=_Conv(sd, p + ".conv1", self.dev, self.dt), conv2=_Conv(sd, p + ".conv2", self.dev, self.dt),
                 g1=self._f32(sd[p + ".norm1.weight"]), b1=self._f32(sd[p + ".norm1.bias"]),
                 g2=self._f32(sd[p + ".norm2.weight"]), b2=self._f32(sd[p + ".norm2.bias"]), short=None)
        if (p + ".conv_shortcut.weight") in sd:
            r["short"] = _Conv(sd, p + ".conv_shortcut", self.dev, self.dt)
        r["tproj_off"] = self._tproj_total
        self._tproj_w.append(sd[p + ".time_emb_proj.weight"].float())
        self._tproj_b.append(sd[p + ".time_emb_proj.bias"].float())
        self._tproj_total += r["conv1"].cout
        return r

    def _pack_attn(self, sd: SD, p: str) -> dict:
        b = p + ".transformer_blocks.0"
        C = sd[p + ".proj_in.weight"].shape[0]
        a = dict(name=p, C=C, d=C // self.heads)
        a["dpad"] = ((a["d"] + 63) // 64) * 64
        a["gn_g"], a["gn_b"] = self._f32(sd[p + ".norm.weight"]), self._f32(sd[p + ".norm.bias"])
        a["pi_w"], a["pi_b"] = self._bf(sd[p + ".proj_in.weight"].reshape(C, C)), self._f32(sd[p + ".proj_in.bias"])
        a["po_w"], a["po_b"] = self._bf(sd[p + ".proj_out.weight"].reshape(C, C)), self._f32(sd[p + ".proj_out.bias"])
        for n in ("norm1", "norm_audio", "norm2", "norm_temp", "norm3"):
            a[n] = (self._f32(sd[f"{b}.{n}.weight"]), self._f32(sd[f"{b}.{n}.bias"]))
        for n in ("attn1", "attn_audio", "attn2", "attn_temp"):
            a[n + ".q"] = self._bf(sd[f"{b}.{n}.to_q.weight"])
            a[n + ".kv"] = self._bf(torch.cat([sd[f"{b}.{n}.to_k.weight"].float(), sd[f"{b}.{n}.to_v.weight"].float()], 0))
            a[n + ".o_w"] = self._bf(sd[f"{b}.{n}.to_out.0.weight"])
            a[n + ".o_b"] = self._f32(sd[f"{b}.{n}.to_out.0.bias"])
        a["attn_temp.qkv"] = torch.cat([a["attn_temp.q"], a["attn_temp.kv"]], 0).contiguous()
        # LayerNorm folded into the projection behind it (ops.LnFold): W * gamma, its row sums, W beta (+ bias)
        kv1 = torch.cat([sd[f"{b}.attn1.to_k.weight"].float(), sd[f"{b}.attn1.to_v.weight"].float()], 0)
        a["attn1.qf"] = self._fold_ln(sd[f"{b}.attn1.to_q.weight"], None, a["norm1"])
        a["attn1.kvf"] = self._fold_ln(kv1, None, a["norm1"])
        a["attn_audio.qf"] = self._fold_ln(sd[f"{b}.attn_audio.to_q.weight"], None, a["norm_audio"])
        a["attn2.qf"] = self._fold_ln(sd[f"{b}.attn2.to_q.weight"], None, a["norm2"])
        # GEGLU: every 128-column tile of the first FF GEMM holds [64 value | 64 gate] columns
        w1, b1 = sd[f"{b}.ff.net.0.proj.weight"].float(), sd[f"{b}.ff.net.0.proj.bias"].float()
        inner = w1.shape[0] // 2
        wv, wg = w1[:inner].view(inner // 64, 64, C), w1[inner:].view(inner // 64, 64, C)
        a["ff1_w"] = self._bf(torch.stack([wv, wg], dim=1).reshape(2 * inner, C))
        a["ff1_b"] = self._f32(torch.stack([b1[:inner].view(-1, 64), b1[inner:].view(-1, 64)], dim=1).reshape(-1))
        a["ff1f"] = self._fold_ln(torch.stack([wv, wg], dim=1).reshape(2 * inner, C), a["ff1_b"], a["norm3"])
        a["ff2_w"], a["ff2_b"] = self._bf(sd[f"{b}.ff.net.2.weight"]), self._f32(sd[f"{b}.ff.net.2.bias"])
        a["pos1_w"], a["pos1_b"] = self._bf(sd[f"{b}.pos_embedding_temp.linear_1.weight"]), self._f32(sd[f"{b}.pos_embedding_temp.linear_1.bias"])
        a["pos2_w"], a["pos2_b"] = self._bf(sd[f"{b}.pos_embedding_temp.linear_2.weight"]), self._f32(sd[f"{b}.pos_embedding_temp.linear_2.bias"])
        self.attns.append(a)
        return a

    def _pack(self, sd: SD) -> None:
        c, ch = self.cfg, self.chans
        nlev, L = len(ch), c["layers_per_block"]
        self._tproj_w, self._tproj_b, self._tproj_total = [], [], 0
        self.attns: List[dict] = []
        ci = c["in_channels"]
        w_in = sd["conv_in.weight"].float().permute(0, 2, 3, 1).reshape(ch[0], 9 * ci)
        self.conv_in = _Conv(sd, "conv_in", self.dev, self.dt)
        self.conv_in.w = self._bf(torch.nn.functional.pad(w_in, (0, 64 - 9 * ci)))
        self.te = [self._bf(sd["time_embedding.linear_1.weight"]), self._f32(sd["time_embedding.linear_1.bias"]),
                   self._bf(sd["time_embedding.linear_2.weight"]), self._f32(sd["time_embedding.linear_2.bias"])]
        self.temb_dim = sd["time_embedding.linear_2.weight"].shape[0]
        self.down, self.up = [], []
        for i in range(nlev):
            has_attn = "Attn" in c["down_block_types"][i]
            blk = dict(res=[], attn=[], down=None)
            for j in range(L):
                blk["res"].append(self._pack_res(sd, f"down_blocks.{i}.resnets.{j}"))
                blk["attn"].append(self._pack_attn(sd, f"down_blocks.{i}.attentions.{j}") if has_attn else None)
            if i < nlev - 1:
                blk["down"] = _Conv(sd, f"down_blocks.{i}.downsamplers.0.conv", self.dev, self.dt)
            self.down.append(blk)
        self.mid = dict(res=[self._pack_res(sd, "mid_block.resnets.0"), self._pack_res(sd, "mid_block.resnets.1")],
                        attn=self._pack_attn(sd, "mid_block.attentions.0"))
        for i in range(nlev):
            has_attn = "Attn" in c["up_block_types"][i]
            blk = dict(res=[], attn=[], up=None)
            for j in range(L + 1):
                blk["res"].append(self._pack_res(sd, f"up_blocks.{i}.resnets.{j}"))
                blk["attn"].append(self._pack_attn(sd, f"up_blocks.{i}.attentions.{j}") if has_attn else None)
            if i < nlev - 1:
                blk["up"] = _Conv(sd, f"up_blocks.{i}.upsamplers.0.conv", self.dev, self.dt)
            self.up.append(blk)
        self.out_g, self.out_b = self._f32(sd["conv_norm_out.weight"]), self._f32(sd["conv_norm_out.bias"])
        co = c["out_channels"]
        self.conv_out = _Conv(sd, "conv_out", self.dev, self.dt)
        self.conv_out.w = self._bf(torch.nn.functional.pad(
            sd["conv_out.weight"].float().permute(0, 2, 3, 1).reshape(co, 9 * ch[0]), (0, 0, 0, 8 - co)))
        self.conv_out.b = self._f32(torch.nn.functional.pad(sd["conv_out.bias"].float(), (0, 8 - co)))
        self.tproj_w = self._bf(torch.cat(self._tproj_w, 0))
        self.tproj_b = self._f32(torch.cat(self._tproj_b, 0))
        del self._tproj_w, self._tproj_b

    # ------------------------------------------------------------------------------------------ buffers
    def buf(self, tag: str, shape: Tuple[int, ...], dtype=None, zero: bool = False) -> torch.Tensor:
        dtype = self.dt if dtype is None else dtype
        key = (tag, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            if self._frozen:
                raise RuntimeError(f"engine buffer {key} requested after prepare(); shapes must be static")
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)
            self._bufs[key] = t
        return t

    _frozen = False
    _tuned = False

    def prepare(self, B: int, F: int, h: int, w: int) -> None:
        """Fixes the problem shape, builds the per-block temporal position embeddings and (by a dry run of the
        launch sequence) allocates every buffer the forward needs."""
        nlev = len(self.chans)
        if h % (1 << (nlev - 1)) or w % (1 << (nlev - 1)):
            raise ValueError(f"latent {h}x{w} must be divisible by {1 << (nlev - 1)}")
        if not (1 <= F <= 32):
            raise ValueError("1 <= F <= 32 frames supported")
        self._frozen = False
        self._bufs.clear()
        self._ctx_bufs.clear()
        self.gen += 1
        self.shape = (B, F, h, w)
        be = self.be
        ar = torch.arange(F, dtype=torch.float32, device=self.dev)
        for a in self.attns:  # pos = Linear2(SiLU(Linear1(sinusoid(arange F)))) (ff_..._transformer_3d.py:348-349)
            C = a["C"]
            feat = torch.empty(F, C, dtype=torch.float32, device=self.dev)
            hid = torch.empty(F, C, dtype=torch.float32, device=self.dev)
            a["pos"] = torch.empty(F, C, dtype=torch.float32, device=self.dev)
            be.timestep_features(ar, feat, F, C, True)
            be.small_linear(feat, a["pos1_w"], a["pos1_b"], hid, F, C, C, 0, 1)
            be.small_linear(hid, a["pos2_w"], a["pos2_b"], a["pos"], F, C, C, 0, 0)
        self.ctx, self.ctx_sig = None, None
        self._tuned = False

    def _ctx_buf(self, tag, shape, dtype):
        key = (tag, tuple(shape), dtype)
        t = self._ctx_bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._ctx_bufs[key] = t
            self.gen += 1
        return t

    def set_context(self, text: torch.Tensor, audio: torch.Tensor, audio_mask: Optional[torch.Tensor]) -> None:
        """Projects the clip's conditioning to per-block keys/values once (they are constant over the sampler loop).
        text (B,F,n_t,768) / audio (B,F,n_a,768) / audio_mask bool (B,F,n_a) in the reference's wire format
        (pipeline_audio_cond_animation.py:299,304-306).  Frame-invariant contexts (what the pipeline produces with
        `repeat`, :177) are projected once per batch entry, frame-varying ones once per (b, f).  With an audio mask
        the projected keys/values are then COMPACTED per (b, f) to the keys that frame may attend (25 of 229 at
        F = 12, segmask_imagebind.py:104-114), so the attention kernel sees short dense key lists and no mask.
        Results land in persistent buffers (same addresses for every clip of the same geometry) so a captured CUDA
        graph stays valid; `ctx_sig` changes when the geometry does and owners of graphs must re-capture."""
        B, F, h, w = self.shape
        assert text.shape[:2] == (B, F) and audio.shape[:2] == (B, F), (text.shape, audio.shape, self.shape)

        def frame_invariant(t):
            return t.stride(1) == 0 or F == 1 or bool((t[:, :1] == t).all())

        ctx, sig = {}, []
        tune = getattr(self.be, "name", "") == "cuda" and not torch.cuda.is_current_stream_capturing()
        if tune:
            self.be.tuning = True  # new K/V projection shapes get their tile plan measured once (cached per shape)
        for name, t in (("attn2", text), ("attn_audio", audio)):
            inv = frame_invariant(t)
            src = t[:, 0] if inv else t.reshape(B * F, t.shape[2], t.shape[3])
            G, nk = src.shape[0], src.shape[1]
            x = src.reshape(G * nk, src.shape[-1]).to(self.dev, self.dt).contiguous()
            gidx, cmask, nv = None, None, nk
            if name == "attn_audio" and audio_mask is not None:
                assert audio_mask.shape == (B, F, nk)
                m = audio_mask.reshape(B * F, nk).to(self.dev).bool()
                nv = max(1, int(m.sum(dim=1).max()))
                # stable order: valid keys first (ascending), padded with key 0 where a row has fewer than nv
                order = torch.argsort((~m).to(torch.int8), dim=1, stable=True)[:, :nv]
                valid = torch.gather(m, 1, order)
                base = (torch.arange(B * F, device=self.dev) // (F if inv else 1)) * nk
                gidx = (base.view(-1, 1) + torch.where(valid, order, torch.zeros_like(order))).reshape(-1)
                cmask = None if bool(valid.all()) else valid.to(torch.uint8).contiguous()
            kvs = []
            for i, a in enumerate(self.attns):
                full = self._ctx_buf(f"{name}.full{i}", (G * nk, 2 * a["C"]), self.dt)
                self.be.gemm(ops.spec_linear(x, a[name + ".kv"], full))
                if gidx is not None:
                    kv = self._ctx_buf(f"{name}.kv{i}", (B * F * nv, 2 * a["C"]), self.dt)
                    torch.index_select(full, 0, gidx, out=kv)
                    kvs.append(kv)
                else:
                    kvs.append(full)
            if gidx is not None:
                ctx[name] = dict(inv=False, G=B * F, nk=nv, kv=kvs)
                if cmask is not None:
                    mb = self._ctx_buf("audio.mask", tuple(cmask.shape), torch.uint8)
                    mb.copy_(cmask)
                    cmask = mb
            else:
                ctx[name] = dict(inv=inv, G=G, nk=nk, kv=kvs)
            sig.append((ctx[name]["inv"], ctx[name]["G"], ctx[name]["nk"], cmask is not None))
            if name == "attn_audio":
                ctx["mask"] = cmask
        if tune:
            self.be.tuning = False
        self.ctx = ctx
        self.ctx_sig = tuple(sig)

    # ------------------------------------------------------------------------------------------ building blocks
    def _ffconv_tail(self, cv: _Conv, y, out, B, F, N, tproj=None, res1=None):
        if N < 64:
            # a frame is smaller than half a tile: gather the three taps so the GEMM gets full 128-row tiles
            # (one-frame tiles would leave 3/4 of every MMA empty and re-read the weights 4x as often)
            C = cv.cout
            g = self.buf("tc_gather", (B * F * N, 3 * C))
            self.be.tconv_gather(y, g, B, F, N, C)
            sp = ops.spec_linear(g, cv.w4[:, :3 * C], out, bias=cv.bt, res0=y, res1=res1)
            if tproj is not None:
                sp.add = ops.RowAdd(tproj, self._tproj_total, div=F * N)
            self.be.gemm(sp)
            return
        self.be.gemm(ops.spec_tconv(y, cv.w4, out, B=B, F=F, N=N, bias=cv.bt, tproj=tproj,
                                    tproj_ld=self._tproj_total, res1=res1))

    def _conv3(self, cv: _Conv, a, B, F, h, w, stride=1):
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        y = self.buf("conv_y", (B * F * ho * wo, cv.cout))
        self.be.gemm(ops.spec_conv3x3(a, cv.w, y, n_img=B * F, h=h, wd=w, stride=stride, bias=cv.b))
        return y, ho, wo

    def _resblock(self, r: dict, x0, x1, B, F, h, w, out_tag: str):
        be, N = self.be, h * w
        M = B * F * N
        C0, C1 = x0.shape[1], (x1.shape[1] if x1 is not None else 0)
        cin, cout = C0 + C1, r["conv1"].cout
        a = self.buf("gn_out", (M, cin))
        be.groupnorm(x0, C0, x1, C1, B, F * N, self.groups, self.eps, r["g1"], r["b1"], True, a)
        y, _, _ = self._conv3(r["conv1"], a, B, F, h, w)
        h1 = self.buf("res_h1", (M, cout))
        tp = self.tproj[:, r["tproj_off"]: r["tproj_off"] + cout]
        self._ffconv_tail(r["conv1"], y, h1, B, F, N, tproj=tp)
        a2 = self.buf("gn_out", (M, cout))
        be.groupnorm(h1, cout, None, 0, B, F * N, self.groups, self.eps, r["g2"], r["b2"], True, a2)
        y2, _, _ = self._conv3(r["conv2"], a2, B, F, h, w)
        if r["short"] is not None:
            sc = r["short"]
            ys = self.buf("short_y", (M, cout))
            be.gemm(ops.spec_linear(x0, sc.w, ys, x2=x1, bias=sc.b))
            s = self.buf("short_s", (M, cout))
            self._ffconv_tail(sc, ys, s, B, F, N)
            res = s
        else:
            assert x1 is None
            res = x0
        out = self.buf(out_tag, (M, cout))
        self._ffconv_tail(r["conv2"], y2, out, B, F, N, res1=res)
        return out

    def _ln_fold(self, a: dict, key: str, st, **kw) -> "ops.LnFold":
        _, wsum, _ = a[key]
        return ops.LnFold(stats=st, wsum=wsum, cols=a["C"], eps=1e-5, **kw)

    def _attention(self, a: dict, name: str, t, n, B, F, N, kv, G, R, nk, mask=None, mask_rows=1, st=None,
                   emit=False):
        """t <- t + Wo softmax(Q K^T / sqrt(d)) V + bo   with Q = Wq n (head-split epilogue).
        st: the row statistics of t - then n is unused and Q = Wq LayerNorm(t) by the fold (ops.LnFold); emit: the
        output projection leaves the statistics of the new t in st for the next sub-block."""
        be, C, d, dpad, H = self.be, a["C"], a["d"], a["dpad"], self.heads
        q = self.buf("attn_q", (B * F * N, C))
        if st is not None:
            wq, _, bq = a[name + ".qf"]
            sp = ops.spec_linear(t, wq, q, bias=bq)
            sp.ln = self._ln_fold(a, name + ".qf", st)
            be.gemm(sp)
        else:
            be.gemm(ops.spec_linear(n, a[name + ".q"], q))
        o = self.buf("attn_o", (B * F * N, C))
        be.attention(ops.AttnSpec(q=q, kv=kv, out=o, G=G, heads=H, R=R, Nk=nk, d=d, dpad=dpad, ldq=C, ldkv=2 * C, ldo=C,
                                  kv_rows_per_group=nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d), mask=mask,
                                  mask_ld=(mask.shape[1] if mask is not None else 0), mask_rows=mask_rows))
        sp = ops.spec_linear(o, a[name + ".o_w"], t, bias=a[name + ".o_b"], res0=t)
        if emit:
            sp.stats_out = st
        be.gemm(sp)

    def _transformer(self, a: dict, x, B, F, h, w, idx: int, out_tag: str):
        be, N, C = self.be, h * w, a["C"]
        M = B * F * N
        g = self.buf("gn_out", (M, C))
        be.groupnorm(x, C, None, 0, B * F, N, self.groups, 1e-6, a["gn_g"], a["gn_b"], False, g)
        t = self.buf("tok", (M, C))
        n = self.buf("ln_out", (M, C))
        # LayerNorm fold (ops.LnFold): four of the block's five LayerNorms never run as a pass - the GEMM that writes
        # the token rows t also leaves their row sums in st, and the projection behind the LayerNorm reads t itself.
        # (norm_temp keeps its kernel: its input is t + pos, whose row sums the producer does not know.)
        fold = self.fold_ln and C % 32 == 0
        st = self.buf("ln_stats", (C // 32, M, 2), torch.float32) if fold else None
        sp = ops.spec_linear(g, a["pi_w"], t, bias=a["pi_b"])
        sp.stats_out = st
        be.gemm(sp)
        # 1. first-frame spatial attention: K/V from the frame-0 rows only (utils.py:137-143)
        kv0 = self.buf("kv0", (B * N, 2 * C))
        if fold:
            av = ops.AView(t, (C, N, B, 1), (C, F * N * C, B * F * N * C))
            wkv, _, bkv = a["attn1.kvf"]
            sp = ops.spec_rows3(av, (N, B, 1), wkv, kv0, bias=bkv)
            sp.ln = self._ln_fold(a, "attn1.kvf", st, grp_rows=N, grp_stride=F * N)
            be.gemm(sp)
        else:
            be.layernorm(t, a["norm1"][0], a["norm1"][1], None, n, M, C, 1e-5, N, F)
            av = ops.AView(n, (C, N, B, 1), (C, F * N * C, B * F * N * C))
            be.gemm(ops.spec_rows3(av, (N, B, 1), a["attn1.kv"], kv0))
        self._attention(a, "attn1", t, n, B, F, N, kv0, B, F * N, N, st=st, emit=fold)
        # 2./3. audio (masked) and text cross-attention onto the per-clip projected contexts
        for name, norm in (("attn_audio", "norm_audio"), ("attn2", "norm2")):
            cx = self.ctx[name]
            if not fold:
                be.layernorm(t, a[norm][0], a[norm][1], None, n, M, C, 1e-5, N, F)
            mask = self.ctx["mask"] if name == "attn_audio" else None
            emit = fold and name == "attn_audio"  # attn2's output feeds norm_temp, which keeps its kernel
            if cx["inv"]:
                self._attention(a, name, t, n, B, F, N, cx["kv"][idx], B, F * N, cx["nk"], mask, N, st=st, emit=emit)
            else:
                self._attention(a, name, t, n, B, F, N, cx["kv"][idx], B * F, N, cx["nk"], mask, N, st=st, emit=emit)
        # 4. temporal attention per pixel; pos goes into the LayerNorm input only (:352-358)
        be.layernorm(t, a["norm_temp"][0], a["norm_temp"][1], a["pos"], n, M, C, 1e-5, N, F)
        qkv = self.buf("qkv_t", (M, 3 * C))
        be.gemm(ops.spec_linear(n, a["attn_temp.qkv"], qkv))
        o = self.buf("attn_o", (M, C))
        be.temporal_attention(qkv, o, B, F, N, self.heads, a["d"], 1.0 / math.sqrt(a["d"]))
        sp = ops.spec_linear(o, a["attn_temp.o_w"], t, bias=a["attn_temp.o_b"], res0=t)
        sp.stats_out = st
        be.gemm(sp)
        # 5. GEGLU feed-forward
        ffh = self.buf("ff_h", (M, 4 * C))
        if fold:
            w1, _, b1 = a["ff1f"]
            sp = ops.spec_linear(t, w1, ffh, bias=b1, geglu=True)
            sp.ln = self._ln_fold(a, "ff1f", st)
            be.gemm(sp)
        else:
            be.layernorm(t, a["norm3"][0], a["norm3"][1], None, n, M, C, 1e-5, N, F)
            be.gemm(ops.spec_linear(n, a["ff1_w"], ffh, bias=a["ff1_b"], geglu=True))
        be.gemm(ops.spec_linear(ffh, a["ff2_w"], t, bias=a["ff2_b"], res0=t))
        out = self.buf(out_tag, (M, C))
        be.gemm(ops.spec_linear(t, a["po_w"], out, bias=a["po_b"], res0=x))
        return out

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, latents: torch.Tensor, timesteps: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """latents fp32 (Bs,Cl,F,h,w) with B % Bs == 0 (the CFG copies share one latent); timesteps fp32 (B,) on the
        device; out fp32 (B,Co,F,h,w)."""
        if self.ctx is None:
            raise RuntimeError("set_context() must be called before forward()")
        if not self._tuned and getattr(self.be, "name", "") == "cuda" and not torch.cuda.is_current_stream_capturing():
            # one throw-away pass of the launch sequence in which every new GEMM shape has its tile plan measured on
            # the device (CudaBackend.gemm / asva_gemm_tune); it only scribbles on the engine's own buffers and `out`
            self._tuned = True
            self.be.tuning = True
            try:
                self.forward(latents, timesteps, out)
            finally:
                self.be.tuning = False
            self.be.save_plans()  # no-op unless ASVA_PLAN_CACHE names a file
        B, F, h, w = self.shape
        be, ch, L = self.be, self.chans, self.cfg["layers_per_block"]
        nlev = len(ch)
        Bs, Cl = latents.shape[0], latents.shape[1]
        assert latents.shape[2:] == (F, h, w) and B % Bs == 0 and latents.dtype == torch.float32
        assert timesteps.shape == (B,) and timesteps.dtype == torch.float32
        # time embedding and all 22 time_emb_proj in three skinny launches
        tf = self.buf("t_feat", (B, ch[0]), torch.float32)
        e1 = self.buf("t_e1", (B, self.temb_dim), torch.float32)
        temb = self.buf("t_emb", (B, self.temb_dim), torch.float32)
        self.tproj = self.buf("t_proj", (B, self._tproj_total), torch.float32)
        be.timestep_features(timesteps, tf, B, ch[0], True)
        be.small_linear(tf, self.te[0], self.te[1], e1, B, self.temb_dim, ch[0], 0, 1)
        be.small_linear(e1, self.te[2], self.te[3], temb, B, self.temb_dim, self.temb_dim, 0, 0)
        be.small_linear(temb, self.tproj_w, self.tproj_b, self.tproj, B, self._tproj_total, self.temb_dim, 1, 0)
        # conv_in
        N = h * w
        col = self.buf("in_col", (B * F * N, 64))
        be.conv_in_im2col(latents, col, B, Bs, Cl, F, h, w)
        y = self.buf("conv_y", (B * F * N, ch[0]))
        be.gemm(ops.spec_linear(col, self.conv_in.w, y, bias=self.conv_in.b))
        x = self.buf("skip0", (B * F * N, ch[0]))
        self._ffconv_tail(self.conv_in, y, x, B, F, N)
        skips = [x]
        ai = 0
        hh, ww = h, w
        for i, blk in enumerate(self.down):
            for j in range(L):
                tag = f"d{i}r{j}"
                x = self._resblock(blk["res"][j], x, None, B, F, hh, ww, tag if blk["attn"][j] is None else "res_out")
                if blk["attn"][j] is not None:
                    x = self._transformer(blk["attn"][j], x, B, F, hh, ww, ai, tag)
                    ai += 1
                skips.append(x)
            if blk["down"] is not None:
                y, ho, wo = self._conv3(blk["down"], x, B, F, hh, ww, stride=2)
                x = self.buf(f"d{i}ds", (B * F * ho * wo, blk["down"].cout))
                self._ffconv_tail(blk["down"], y, x, B, F, ho * wo)
                hh, ww = ho, wo
                skips.append(x)
        x = self._resblock(self.mid["res"][0], x, None, B, F, hh, ww, "res_out")
        x = self._transformer(self.mid["attn"], x, B, F, hh, ww, ai, "mid_t")
        ai += 1
        flip = [0]

        def up_tag():  # ping-pong so a block never writes the buffer it is still reading
            flip[0] ^= 1
            return "up_a" if flip[0] else "up_b"

        x = self._resblock(self.mid["res"][1], x, None, B, F, hh, ww, up_tag())
        for i, blk in enumerate(self.up):
            for j in range(L + 1):
                has = blk["attn"][j] is not None
                x = self._resblock(blk["res"][j], x, skips.pop(), B, F, hh, ww, "res_out" if has else up_tag())
                if has:
                    x = self._transformer(blk["attn"][j], x, B, F, hh, ww, ai, up_tag())
                    ai += 1
            if blk["up"] is not None:
                C = x.shape[1]
                u = self.buf("gn_out", (B * F * 4 * hh * ww, C))
                be.groupnorm_apply(x, C, None, 0, None, B, B * F, hh, ww, False, True, u)
                hh, ww = 2 * hh, 2 * ww
                y, _, _ = self._conv3(blk["up"], u, B, F, hh, ww)
                x = self.buf(up_tag(), (B * F * hh * ww, blk["up"].cout))
                self._ffconv_tail(blk["up"], y, x, B, F, hh * ww)
        # conv_norm_out -> SiLU -> conv_out (3x3 to 4 channels, fp32) -> its temporal 3-tap in fp32
        C = ch[0]
        a = self.buf("gn_out", (B * F * N, C))
        be.groupnorm(x, C, None, 0, B, F * N, self.groups, self.eps, self.out_g, self.out_b, True, a)
        yo = self.buf("out_y", (B * F * N, 8), torch.float32)
        be.gemm(ops.spec_conv3x3(a, self.conv_out.w, yo, n_img=B * F, h=h, wd=w, bias=self.conv_out.b, out_fp32=True))
        be.conv_out_finish(yo, 8, self.conv_out.wt_full, self.conv_out.bt, out, B, self.cfg["out_channels"], F, h, w)
        self._frozen = True
        return out


class GraphRunner:
    """Runs a fixed launch sequence: first call eagerly (one-time kernel attribute setup, buffer allocation), second
    call under CUDA-graph capture, later calls as graph replays - ~800 launches per step become one submission.
    ASVA_NO_GRAPH=1 keeps everything eager (debugging, ncu)."""

    def __init__(self, fn, backend, enabled: Optional[bool] = None):
        import os
        self.fn, self.be = fn, backend
        if enabled is None:
            enabled = os.environ.get("ASVA_NO_GRAPH", "0") != "1"
        self.enabled = enabled and getattr(backend, "name", "") == "cuda"
        self.graph = None
        self.calls = 0
        self.launches_per_call = 0
        self.total_launches = 0

    def __call__(self) -> None:
        if not self.enabled or (self.graph is None and self.calls == 0):
            n0 = self.be.launches
            self.fn()
            self.launches_per_call = self.be.launches - n0
        elif self.graph is None:
            # Dead Python cycles that hold CUDA memory (an engine dropped by its model, a finished session) must not be
            # collected INSIDE the capture window: releasing their blocks there invalidates the capture
            # ("operation failed due to a previous error during capture"; torch >= 2.9 no longer runs gc.collect() at
            # capture entry).  Collect them now and keep the cyclic collector off until the capture has ended.
            import gc
            gc.collect()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = self.be.launches
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    self.fn()
            finally:
                if gc_was_on:
                    gc.enable()
            self.launches_per_call = self.be.launches - n0
            self.graph = g
            g.replay()
        else:
            self.graph.replay()
        self.calls += 1
        self.total_launches += self.launches_per_call

    def reset(self) -> None:
        self.graph = None
        self.calls = 0
