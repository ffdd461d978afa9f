"""Torch interpreter of the operator specs in asva_b200/ops.py  —  TEST INFRASTRUCTURE ONLY.

It restates, in plain torch (any device), what each C-ABI entry point of include/asva_b200.h computes from the same
descriptor.  Uses: (1) on CPU, run the whole engine through it and compare with the oracle, which checks the
host-side logic (weight repacking, descriptor construction, op sequencing) without a GPU; (2) on the GPU box, serve
as the per-kernel reference that the CUDA kernels are compared against on identical inputs.  The product package
never imports this file."""
import math
from typing import Optional

import torch

from asva_b200.ops import AttnSpec, GemmSpec


def _flat(t: torch.Tensor) -> torch.Tensor:
    """1-D view of the tensor's storage starting at its first element (so offsets mirror raw pointers)."""
    n = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    return torch.as_strided(t, (n,), (1,))


class SimBackend:
    name = "sim"

    def __init__(self, round_bf16_accum: bool = False) -> None:
        self.launches = 0

    # ------------------------------------------------------------------ gemm
    def _accumulate(self, s: GemmSpec) -> torch.Tensor:
        """fp32 [M, N] = sum over K segments of A_seg @ W[:, wk : wk + len]^T  (rows of frame 0 use wk_first)."""
        D1, D2, D3 = s.out_dims
        dev = s.out.device
        o1 = torch.arange(D1, device=dev).view(1, 1, D1).expand(D3, D2, D1).reshape(-1)
        o2 = torch.arange(D2, device=dev).view(1, D2, 1).expand(D3, D2, D1).reshape(-1)
        o3 = torch.arange(D3, device=dev).view(D3, 1, 1).expand(D3, D2, D1).reshape(-1)
        Wf = torch.as_strided(s.w, (s.N, s.wcols), (s.ldw, 1)).float()
        acc = torch.zeros(o1.numel(), s.N, dtype=torch.float32, device=dev)
        for sg in s.segs:
            av = s.a[sg.src]
            flat = _flat(av.t)
            c_ext, e1, e2, e3 = av.dims
            i1 = o1 * s.trav[0] + sg.off[0]
            i2 = o2 * s.trav[1] + sg.off[1] if sg.fix2 < 0 else torch.full_like(o2, sg.fix2)
            i3 = o3 * s.trav[2] + sg.off[2]
            ok = (i1 >= 0) & (i1 < e1) & (i2 >= 0) & (i2 < e2) & (i3 >= 0) & (i3 < e3)
            base = i1.clamp(0, e1 - 1) * av.strides[0] + i2.clamp(0, e2 - 1) * av.strides[1] + \
                i3.clamp(0, e3 - 1) * av.strides[2]
            n = sg.num_kb * 64
            kk = torch.arange(n, device=dev) + sg.c0
            okc = kk < c_ext
            idx = base.view(-1, 1) + kk.clamp(max=c_ext - 1).view(1, -1)
            vals = flat[idx].float() * (ok.view(-1, 1) & okc.view(1, -1)).float()
            part = vals @ Wf[:, sg.wk: sg.wk + n].t()
            if sg.wk_first >= 0:
                alt = vals @ Wf[:, sg.wk_first: sg.wk_first + n].t()
                part = torch.where((o2 == 0).view(-1, 1), alt, part)
            acc = acc + part
        return acc

    def gemm(self, s: GemmSpec) -> None:
        self.launches += 1
        acc = self._accumulate(s)
        M = acc.shape[0]
        dev = acc.device
        rows = torch.arange(M, device=dev)
        if s.ln is not None:  # LayerNorm fold: rstd * (acc - mean * wsum); the column terms follow as usual
            f = s.ln
            sr = rows if f.grp_rows == 0 else (rows // f.grp_rows) * f.grp_stride + rows % f.grp_rows
            st = f.stats.float()[:, sr, :]
            sm, sq = st[..., 0], st[..., 1]
            tot, tot2 = torch.zeros_like(sm[0]), torch.zeros_like(sq[0])
            for i in range(st.shape[0]):  # slot order, like the kernel
                tot, tot2 = tot + sm[i], tot2 + sq[i]
            mean = tot / f.cols
            rstd = torch.rsqrt((tot2 / f.cols - mean * mean).clamp_min(0.0) + f.eps)
            acc = rstd.view(-1, 1) * (acc - mean.view(-1, 1) * f.wsum[: s.N].float().view(1, -1))
        if s.geglu:
            assert s.N % 128 == 0
            if s.bias is not None:
                acc = acc + s.bias[: s.N].float().view(1, -1)
            t = acc.view(M, s.N // 128, 2, 64)
            hv, gv = t[:, :, 0, :], t[:, :, 1, :]
            val = (hv * torch.nn.functional.gelu(gv)).reshape(M, s.N // 2)
            n_out = s.N // 2
        else:
            val = acc
            if s.bias is not None:
                val = val + s.bias[: s.N].float().view(1, -1)
            colsN = torch.arange(s.N, device=dev)
            if s.add is not None:
                idx = ((rows // s.add.div) * s.add.ld).view(-1, 1) + colsN.view(1, -1)
                val = val + _flat(s.add.t)[idx].float()
            for r, ld in zip(s.res, s.res_ld):
                if r is None:
                    continue
                idx = (rows * ld).view(-1, 1) + colsN.view(1, -1)
                val = val + _flat(r)[idx].float()
            n_out = s.N
            if s.stats_out is not None:  # (sum, sum of squares) of the fp32 values per row and 32-column slot
                assert s.N % 32 == 0
                v32 = val.float().view(M, s.N // 32, 32)
                st = torch.stack([v32.sum(-1), (v32 * v32).sum(-1)], dim=-1).permute(1, 0, 2)
                s.stats_out.view(-1)[: st.numel()].copy_(st.reshape(-1))
        cols = torch.arange(n_out, device=dev)
        off = (rows * s.ldo).view(-1, 1) + cols.view(1, -1)
        outf = _flat(s.out)
        outf[off.reshape(-1)] = val.reshape(-1).to(s.out.dtype)

    # ------------------------------------------------------------------ attention
    def attention(self, s: AttnSpec) -> None:
        self.launches += 1
        G, H, R, d = s.G, s.heads, s.R, s.d
        q = torch.as_strided(_flat(s.q), (G, R, H, d), (R * s.ldq, s.ldq, d, 1)).permute(0, 2, 1, 3).float()
        kvf = _flat(s.kv)
        kv = torch.as_strided(kvf, (G, s.Nk, s.ldkv), (s.kv_rows_per_group * s.ldkv, s.ldkv, 1))
        k = kv[:, :, s.k_col0: s.k_col0 + H * d].reshape(G, s.Nk, H, d).permute(0, 2, 1, 3).float()
        v = kv[:, :, s.v_col0: s.v_col0 + H * d].reshape(G, s.Nk, H, d).permute(0, 2, 1, 3).float()
        sc = torch.einsum("ghrd,ghkd->ghrk", q, k) * s.scale
        if s.mask is not None:
            m = torch.as_strided(_flat(s.mask), (G * R // s.mask_rows, s.Nk), (s.mask_ld, 1)).bool()
            m = m.repeat_interleave(s.mask_rows, dim=0).view(G, 1, R, s.Nk)
            sc = sc.masked_fill(~m, float("-inf"))
        p = torch.softmax(sc, dim=-1)
        o = torch.einsum("ghrk,ghkd->ghrd", p, v)  # [G,H,R,d]
        o = o.permute(0, 2, 1, 3).reshape(G * R, H * d)
        outv = torch.as_strided(_flat(s.out), (G * R, H * d), (s.ldo, 1))
        outv.copy_(o.to(s.out.dtype))

    def temporal_attention(self, qkv, out, B, F, N, heads, d, scale, form=0) -> None:
        self.launches += 1
        C = heads * d
        t = qkv.view(B, F, N, 3, heads, d).float()
        q, k, v = t[:, :, :, 0], t[:, :, :, 1], t[:, :, :, 2]  # [B,F,N,H,d]
        sc = torch.einsum("bfnhd,bgnhd->bnhfg", q, k) * scale
        p = torch.softmax(sc, dim=-1)
        o = torch.einsum("bnhfg,bgnhd->bfnhd", p, v)
        out.view(B, F, N, C).copy_(o.reshape(B, F, N, C).to(out.dtype))

    # ------------------------------------------------------------------ norms
    def layernorm(self, x, gamma, beta, pos, out, M, C, eps, N, F) -> None:
        self.launches += 1
        v = x.view(M, C).float()
        if pos is not None:
            f = (torch.arange(M, device=x.device) // N) % F
            v = v + pos.view(F, C)[f]
        o = torch.nn.functional.layer_norm(v, (C,), gamma.float(), beta.float(), eps)
        out.view(M, C).copy_(o.to(out.dtype))

    def groupnorm_ws_floats(self, n_inst, rows, C) -> int:
        return 16

    def groupnorm_stats(self, x0, C0, x1, C1, n_inst, rows, groups, eps, gamma, beta, stats, ws) -> None:
        self.launches += 2
        v = x0.view(n_inst, rows, C0).float()
        if x1 is not None:
            v = torch.cat([v, x1.view(n_inst, rows, C1).float()], dim=2)
        Ct = v.shape[2]
        g = v.view(n_inst, rows, groups, Ct // groups).permute(0, 2, 1, 3).reshape(n_inst, groups, -1).double()
        mean = g.mean(dim=2)
        rstd = 1.0 / torch.sqrt(g.var(dim=2, unbiased=False) + eps)
        mean_c = mean.repeat_interleave(Ct // groups, dim=1)
        scale = rstd.repeat_interleave(Ct // groups, dim=1) * gamma.double().view(1, Ct)
        shift = beta.double().view(1, Ct) - mean_c * scale
        stats.view(n_inst, Ct, 2).copy_(torch.stack([scale, shift], dim=2).float())

    def groupnorm(self, x0, C0, x1, C1, n_inst, rows, groups, eps, gamma, beta, silu, out) -> None:
        Ct = C0 + (C1 if x1 is not None else 0)
        st = torch.empty(n_inst, Ct, 2, dtype=torch.float32, device=x0.device)
        self.groupnorm_stats(x0, C0, x1, C1, n_inst, rows, groups, eps, gamma, beta, st, None)
        self.groupnorm_apply(x0, C0, x1, C1, st, n_inst, n_inst, 1, rows, silu, False, out)
        self.launches -= 2  # one launch on the device

    def groupnorm_apply(self, x0, C0, x1, C1, stats, n_inst, n_img, h, w, silu, upsample, out) -> None:
        self.launches += 1
        v = x0.view(n_img, h, w, C0).float()
        if x1 is not None:
            v = torch.cat([v, x1.view(n_img, h, w, C1).float()], dim=3)
        Ct = v.shape[3]
        if stats is not None:
            st = stats.view(n_inst, Ct, 2).repeat_interleave(n_img // n_inst, dim=0)
            v = v * st[:, :, 0].view(n_img, 1, 1, Ct) + st[:, :, 1].view(n_img, 1, 1, Ct)
        if silu:
            v = torch.nn.functional.silu(v)
        if upsample:
            v = v.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        out.view(v.shape).copy_(v.to(out.dtype))

    # ------------------------------------------------------------------ small kernels
    def conv_in_im2col(self, lat, out, B, Bs, Cl, F, h, w) -> None:
        self.launches += 1
        x = lat.view(Bs, Cl, F, h, w)[torch.arange(B, device=lat.device) % Bs]  # [B,Cl,F,h,w]
        x = x.permute(0, 2, 1, 3, 4).reshape(B * F, Cl, h, w)
        xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
        cols = []
        for ky in range(3):
            for kx in range(3):
                cols.append(xp[:, :, ky: ky + h, kx: kx + w])  # [BF, Cl, h, w]
        t = torch.stack(cols, dim=1)  # [BF, 9, Cl, h, w]
        t = t.permute(0, 3, 4, 1, 2).reshape(B * F * h * w, 9 * Cl)
        o = torch.zeros(B * F * h * w, 64, device=lat.device, dtype=torch.float32)
        o[:, : 9 * Cl] = t
        out.view(B * F * h * w, 64).copy_(o.to(out.dtype))

    def tconv_gather(self, y, out, B, F, N, C) -> None:
        self.launches += 1
        v = y.view(B, F, N, C)
        prev = torch.cat([v[:, :1], v[:, :-1]], dim=1)
        out.view(B, F, N, 3 * C).copy_(torch.cat([v, prev, v[:, :1].expand_as(v)], dim=3))

    def conv_out_finish(self, y, ldy, wt, bt, out, B, Co, F, h, w) -> None:
        self.launches += 1
        hw = h * w
        v = torch.as_strided(_flat(y), (B, F, hw, Co), (F * hw * ldy, hw * ldy, ldy, 1)).float()
        prev = torch.cat([v[:, :1], v[:, :-1]], dim=1)
        head = v[:, :1].expand_as(v)
        cat = torch.cat([head, prev, v], dim=3)  # [B,F,hw,3Co]
        o = v + cat @ wt.view(Co, 3 * Co).float().t() + bt.float()
        out.view(B, Co, F, hw).copy_(o.permute(0, 3, 1, 2))

    def small_linear(self, x, w, bias, out, M, N, K, act_in, act_out) -> None:
        self.launches += 1
        v = x.view(M, K).float()
        if act_in == 1:
            v = torch.nn.functional.silu(v)
        o = v @ w.view(N, K).float().t()
        if bias is not None:
            o = o + bias.float()
        if act_out == 1:
            o = torch.nn.functional.silu(o)
        out.view(M, N).copy_(o)

    def timestep_features(self, t, out, B, dim, flip) -> None:
        self.launches += 1
        half = dim // 2
        freq = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
        arg = t.view(B, 1).float() * freq.view(1, half)
        s, c = torch.sin(arg), torch.cos(arg)
        out.view(B, dim).copy_(torch.cat([c, s], dim=1) if flip else torch.cat([s, c], dim=1))

    def _cfg(self, eps, k, coef):
        e = 0
        for j in range(k):
            e = e + coef[j] * eps[j]
        return e

    def softmax_rows(self, scores, probs, rows, cols, scale) -> None:
        self.launches += 1
        probs[:rows, :cols] = torch.softmax(scores[:rows, :cols].float() * scale, dim=-1).to(probs.dtype)

    def cfg_ddim_step(self, eps, k, lat, coef, C, F, hw, clips: int = 1) -> None:
        self.launches += 1
        e = self._cfg(eps.view(k, clips, C, F, hw), k, coef)
        lv = lat.view(clips, C, F, hw)
        lv[:, :, 1:] = coef[3] * lv[:, :, 1:] + coef[4] * e[:, :, 1:]

    def cfg_plms_step(self, eps, k, lat, hist, coef, slots, C, F, hw, clips: int = 1) -> None:
        self.launches += 1
        e = self._cfg(eps.view(k, clips, C, F, hw), k, coef)
        hv = hist.view(4, clips, C, F, hw)
        ehat = coef[5] * e
        for j in range(1, 4):
            if float(coef[5 + j]) != 0.0:
                ehat = ehat + coef[5 + j] * hv[int(slots[j])]
        if int(slots[0]) >= 0:
            hv[int(slots[0]), :, :, 1:] = e[:, :, 1:]
        lv = lat.view(clips, C, F, hw)
        lv[:, :, 1:] = coef[3] * lv[:, :, 1:] + coef[4] * ehat[:, :, 1:]
