"""Tensor-level operator layer over the C ABI (include/asva_b200.h).

`GemmSpec` / `AttnSpec` mirror `asva_gemm_desc` / `asva_attn_desc` field for field but hold torch tensors
instead of raw pointers.  The builder functions (`spec_linear`, `spec_conv3x3`, `spec_tconv`, ...) encode how each
reference operator maps onto the generic kernels; the backend object executes a spec.  The only product backend
is `CudaBackend` (ctypes -> libasva_b200.so).  tests/sim_backend.py interprets the same specs with torch on the
CPU so the descriptor logic can be checked against the oracle without a GPU - it is test infrastructure and is
never importable from this package."""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

BIG = 1 << 30


@dataclass
class Seg:
    src: int
    c0: int
    off: Tuple[int, int, int]
    num_kb: int
    wk: int = -1        # first column of W this segment multiplies; -1 = consecutive (filled in by GemmSpec)
    wk_first: int = -1  # >= 0: column of W used instead by tiles whose d2 origin is 0 (box[1] must be 1)
    fix2: int = -1      # >= 0: absolute d2 coordinate of the rows read (box[1] must be 1)


@dataclass
class RowAdd:
    t: torch.Tensor  # fp32; element [(row // div) * ld + col]
    ld: int
    div: int


@dataclass
class AView:
    """4-D (c, d1, d2, d3) channels-innermost view of a bf16 tensor (element strides for d1..d3)."""
    t: torch.Tensor
    dims: Tuple[int, int, int, int]
    strides: Tuple[int, int, int]


@dataclass
class LnFold:
    """LayerNorm folded into the projection that follows it (asva_gemm_desc.ln_*): the GEMM runs on the un-normalised
    rows with W * gamma and its epilogue applies rstd * (acc - mean * wsum) + (W beta + bias); mean / rstd come from
    the per-slot row sums `stats` that the GEMM which produced the rows left (GemmSpec.stats_out)."""
    stats: torch.Tensor   # fp32 [cols/32, rows, 2]: (sum, sum of squares) per 32-column slot
    wsum: torch.Tensor    # fp32 [N]: sum_k (W gamma)[n, k] of the bf16 weight the GEMM reads
    cols: int             # C, the normalised width (= K)
    eps: float
    grp_rows: int = 0     # output row r reads statistics row (r // grp_rows) * grp_stride + r % grp_rows (0: row r)
    grp_stride: int = 0


@dataclass
class GemmSpec:
    a: List[Optional[AView]]
    box: Tuple[int, int, int]
    trav: Tuple[int, int, int]
    out_dims: Tuple[int, int, int]
    segs: List[Seg]
    w: torch.Tensor  # bf16 [>=N, ldw]
    ldw: int
    N: int
    K: int
    out: torch.Tensor  # [M, ldo] row-major (bf16, or fp32 with out_fp32)
    ldo: int = 0
    wcols: int = 0
    bias: Optional[torch.Tensor] = None
    add: Optional[RowAdd] = None
    res: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None, None])
    res_ld: List[int] = field(default_factory=lambda: [0, 0])
    geglu: bool = False
    out_fp32: bool = False
    block_n: int = 0    # 0 = let the library choose (cost model); the engine's tuner sets measured choices
    split_k: int = 0
    cta_group: int = 0
    epilogue: int = 0   # 0 = auto, 1 = panel (TMA) epilogue, 2 = per-warp (direct), 3 = warp-private TMA, 4 = cluster split-K
    ln: Optional[LnFold] = None
    stats_out: Optional[torch.Tensor] = None  # fp32 [N/32, M, 2]: row sums of what this GEMM stores (for a later LnFold)

    def __post_init__(self):
        k = 0
        for sg in self.segs:
            if sg.wk < 0:
                sg.wk = k
            k += 64 * sg.num_kb
        if self.wcols == 0:
            self.wcols = max([self.K] + [max(sg.wk, sg.wk_first) + 64 * sg.num_kb for sg in self.segs])
        if self.ldo == 0:
            self.ldo = self.out.stride(0)

    @property
    def M(self) -> int:
        return self.out_dims[0] * self.out_dims[1] * self.out_dims[2]


@dataclass
class AttnSpec:
    q: torch.Tensor  # bf16 [G*R, ldq] token-major, head h at columns h*d..
    kv: torch.Tensor  # bf16 rows of ldkv
    out: torch.Tensor  # bf16 [G*R, ldo]
    G: int
    heads: int
    R: int
    Nk: int
    d: int
    dpad: int
    ldq: int
    ldkv: int
    ldo: int
    kv_rows_per_group: int
    k_col0: int
    v_col0: int
    scale: float
    mask: Optional[torch.Tensor] = None  # uint8 [G*R/mask_rows, mask_ld]
    mask_ld: int = 0
    mask_rows: int = 1
    form: int = 0  # 0 = auto, 1 = tcgen05 kernel, 2 = warp-MMA kernel for <= 128 keys (asva_attn_desc.form)


# ---------------------------------------------------------------------------------------------------
# spec builders
# ---------------------------------------------------------------------------------------------------
def pick_box(dims: Sequence[int], limit: int = 128) -> Tuple[int, int, int]:
    """Largest (b1, b2, b3) box with b1*b2*b3 <= limit, filling the innermost dimension first."""
    b1 = min(dims[0], limit)
    b2 = min(dims[1], max(1, limit // b1))
    b3 = min(dims[2], max(1, limit // (b1 * b2)))
    return (b1, b2, b3)


def _rows_view(x: torch.Tensor) -> AView:
    assert x.dim() == 2 and x.stride(1) == 1, "expected a [rows, channels] view with contiguous channels"
    M, K = x.shape
    ld = x.stride(0)
    return AView(x, (K, M, 1, 1), (ld, ld * max(M, 1), ld * max(M, 1)))


def spec_linear(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, x2: Optional[torch.Tensor] = None,
                bias: Optional[torch.Tensor] = None, res0: Optional[torch.Tensor] = None,
                res1: Optional[torch.Tensor] = None, geglu: bool = False, out_fp32: bool = False) -> GemmSpec:
    """out[M, N] = [x | x2] @ w[N, K]^T (+bias, +residuals, GEGLU).  x, x2, out: 2-D row views."""
    M, K0 = x.shape
    N, K = w.shape
    segs = [Seg(0, 0, (0, 0, 0), K0 // 64)]
    a = [_rows_view(x), None]
    if x2 is not None:
        assert x2.shape[0] == M
        a[1] = _rows_view(x2)
        segs.append(Seg(1, 0, (0, 0, 0), x2.shape[1] // 64))
    assert sum(s.num_kb for s in segs) * 64 == K, (K0, K)
    spec = GemmSpec(a=a, box=(min(M, 128), 1, 1), trav=(1, 1, 1), out_dims=(M, 1, 1), segs=segs, w=w,
                    ldw=w.stride(0), N=N, K=K, out=out, bias=bias, geglu=geglu, out_fp32=out_fp32)
    spec.res = [res0, res1]
    spec.res_ld = [r.stride(0) if r is not None else 0 for r in spec.res]
    return spec


def spec_rows3(x: AView, box_dims: Tuple[int, int, int], w: torch.Tensor, out: torch.Tensor, *,
               bias: Optional[torch.Tensor] = None, out_fp32: bool = False) -> GemmSpec:
    """Plain GEMM over a strided 3-level row set (e.g. the frame-0 rows of every clip)."""
    N, K = w.shape
    assert K == x.dims[0]
    return GemmSpec(a=[x, None], box=pick_box(box_dims), trav=(1, 1, 1), out_dims=tuple(box_dims),
                    segs=[Seg(0, 0, (0, 0, 0), K // 64)], w=w, ldw=w.stride(0), N=N, K=K, out=out, bias=bias,
                    out_fp32=out_fp32)


def spec_conv3x3(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, n_img: int, h: int, wd: int,
                 stride: int = 1, bias: Optional[torch.Tensor] = None, out_fp32: bool = False,
                 pad_lo: int = 1) -> GemmSpec:
    """Implicit-GEMM 3x3 conv.  x: [n_img*h*wd, Cin] channels-last; w: [Cout, 9*Cin] with K index
    (ky*3 + kx)*Cin + c; out: [n_img*ho*wo, >=Cout].  Padding 1 on every side, or - pad_lo = 0 - only on the
    right / bottom (diffusers' VAE Downsample2D: F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding); the
    out-of-range taps are TMA zero fill either way."""
    Cin = x.shape[1]
    N, K = w.shape
    assert K == 9 * Cin and Cin % 64 == 0 and x.stride(1) == 1 and x.stride(0) == Cin and pad_lo in (0, 1)
    ho, wo = (h + pad_lo + 1 - 3) // stride + 1, (wd + pad_lo + 1 - 3) // stride + 1
    segs = [Seg(0, 0, (kx - pad_lo, ky - pad_lo, 0), Cin // 64) for ky in range(3) for kx in range(3)]
    av = AView(x, (Cin, wd, h, n_img), (Cin, wd * Cin, h * wd * Cin))
    return GemmSpec(a=[av, None], box=pick_box((wo, ho, n_img)), trav=(stride, stride, 1),
                    out_dims=(wo, ho, n_img), segs=segs, w=w, ldw=w.stride(0), N=N, K=K, out=out, bias=bias,
                    out_fp32=out_fp32)


def spec_tconv(y: torch.Tensor, w4: torch.Tensor, out: torch.Tensor, *, B: int, F: int, N: int,
               bias: Optional[torch.Tensor] = None, tproj: Optional[torch.Tensor] = None, tproj_ld: int = 0,
               res1: Optional[torch.Tensor] = None) -> GemmSpec:
    """conv_temp as ONE GEMM (utils.py:43-53 restated):
        out_f = y_f + Wc y_f + Wp y_{max(f-1,0)} + Wh y_0 + bt (+ tproj[b]) (+ res1)
    y, out: [B*F*N, C]; w4: [C, 4C] = [Wc | Wp | Wh | Wh + Wp].  Every tile holds rows of ONE frame (box (n, 1, b)),
    so the three K segments are: the tile's own rows, the same rows one frame earlier (coordinate offset -1; frame 0
    reads TMA zero fill) and the same rows of frame 0 (fix2 = 0), whose weight block is Wh - or Wh + Wp for the
    tiles of frame 0 itself (wk_first), which puts back the previous-frame term the zero fill dropped."""
    Cc = y.shape[1]
    assert w4.shape == (Cc, 4 * Cc) and y.stride(0) == Cc
    av = AView(y, (Cc, N, F, B), (Cc, N * Cc, F * N * Cc))
    kb = Cc // 64
    segs = [Seg(0, 0, (0, 0, 0), kb, wk=0), Seg(0, 0, (0, -1, 0), kb, wk=Cc),
            Seg(0, 0, (0, 0, 0), kb, wk=2 * Cc, wk_first=3 * Cc, fix2=0)]
    b1 = min(N, 128)
    box = (b1, 1, min(B, max(1, 128 // b1)))
    spec = GemmSpec(a=[av, None], box=box, trav=(1, 1, 1), out_dims=(N, F, B), segs=segs, w=w4,
                    ldw=w4.stride(0), N=Cc, K=3 * Cc, out=out, bias=bias)
    if tproj is not None:
        spec.add = RowAdd(tproj, tproj_ld, div=F * N)
    spec.res = [y, res1]
    spec.res_ld = [Cc, res1.stride(0) if res1 is not None else 0]
    return spec


# ---------------------------------------------------------------------------------------------------
# CUDA backend
# ---------------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class CudaBackend:
    """Executes specs through libasva_b200.so on the current CUDA stream. No fallback of any kind."""

    name = "cuda"

    SPLITK_WS_BYTES = 96 << 20

    def __init__(self) -> None:
        self.lib = _lib.load()
        self.launches = 0
        self._gn_sync = {}  # grid-barrier workspace of the fused GroupNorm, one per device
        self._prof = None
        self._nvtx = os.environ.get("ASVA_NVTX", "0") == "1"
        self._ws = {}
        self.tuning = False
        self.plan_cache = {}
        self._plan_file = os.environ.get("ASVA_PLAN_CACHE", "")
        if self._plan_file and os.path.exists(self._plan_file):
            self.load_plans(self._plan_file)

    def load_plans(self, path: str) -> None:
        """Measured tile plans saved by save_plans (a profiler run must not re-measure them: timings taken under
        ncu are meaningless)."""
        import ast
        with open(path) as f:
            for line in f:
                if line.strip():
                    k, v = line.rstrip("\n").split(" => ")
                    v = tuple(ast.literal_eval(v))
                    self.plan_cache[ast.literal_eval(k)] = v + (0,) * (4 - len(v))  # older files: no epilogue field

    def save_plans(self, path: str = "") -> None:
        path = path or self._plan_file
        if path:
            with open(path, "w") as f:
                for k, v in self.plan_cache.items():
                    f.write(f"{k!r} => {list(v)!r}\n")

    def splitk_ws(self) -> torch.Tensor:
        """fp32 scratch for split-K partial tiles, one per device, allocated on first use (before any graph
        capture: the engine's first eager step touches it)."""
        dev = torch.cuda.current_device()
        t = self._ws.get(dev)
        if t is None:
            t = torch.empty(self.SPLITK_WS_BYTES // 4, dtype=torch.float32, device=f"cuda:{dev}")
            self._ws[dev] = t
        return t

    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    # -- live per-family timing (bench.py): CUDA events on the launching stream around every C-ABI call
    def profile_begin(self) -> None:
        self._prof = []

    def profile_end(self):
        """-> {family: (calls, total_ms)}; families: gemm, attention, temporal_attention, layernorm, groupnorm, ..."""
        torch.cuda.synchronize()
        out = {}
        for fam, e0, e1 in self._prof or []:
            c, t = out.get(fam, (0, 0.0))
            out[fam] = (c + 1, t + e0.elapsed_time(e1))
        self._prof = None
        return out

    def _timed(self, fam: str):
        be = self

        class _T:
            def __enter__(self):
                if be._nvtx:  # ASVA_NVTX=1: one NVTX range per C-ABI call, named by kernel family (nsys / ncu --nvtx)
                    torch.cuda.nvtx.range_push("asva:" + fam)
                if be._prof is not None:
                    self.e0 = torch.cuda.Event(enable_timing=True)
                    self.e0.record()

            def __exit__(self, *a):
                if be._nvtx:
                    torch.cuda.nvtx.range_pop()
                if be._prof is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    be._prof.append((fam, self.e0, e1))

        return _T()

    @staticmethod
    def _chk_dev(*ts: Optional[torch.Tensor]) -> None:
        for t in ts:
            if t is not None and not t.is_cuda:
                raise _lib.AsvaError("asva_b200 operators need CUDA tensors (no CPU fallback)")

    # -- measured tile plans: while `tuning` is on, the first launch of every new problem shape sweeps the feasible
    #    (block_n, split_k, cta_group) plans on the device (asva_gemm_tune) and the winner is cached per shape
    @staticmethod
    def gemm_signature(s: GemmSpec) -> tuple:
        return (s.box, s.trav, s.out_dims, tuple((g.src, g.num_kb, g.off, g.wk_first >= 0, g.fix2) for g in s.segs),
                s.N, s.K, s.bias is not None, s.add is not None, sum(r is not None for r in s.res), s.geglu,
                s.out_fp32, tuple(None if a is None else a.dims for a in s.a), s.ln is not None,
                s.stats_out is not None)

    def gemm(self, s: GemmSpec) -> None:
        if s.block_n == 0 and s.split_k == 0 and s.cta_group == 0 and s.epilogue == 0:
            sig = self.gemm_signature(s)
            plan = self.plan_cache.get(sig)
            if plan is None and self.tuning:
                d = self._gemm_desc(s)
                bn, sp, cg, ep, us = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_float(0.0)
                _lib.check(self.lib.asva_gemm_tune(d, self._stream(), 8, C.byref(bn), C.byref(sp), C.byref(cg),
                                                   C.byref(ep), C.byref(us)), "asva_gemm_tune")
                plan = (bn.value, sp.value, cg.value, ep.value)
                self.plan_cache[sig] = plan
            if plan is not None:
                s.block_n, s.split_k, s.cta_group, s.epilogue = (tuple(plan) + (0,) * 4)[:4]
        d = self._gemm_desc(s)
        with self._timed('gemm'):
            _lib.check(self.lib.asva_gemm(d, self._stream()), "asva_gemm")
        pl = self.gemm_plan(s, d)
        self.launches += 2 if (pl[1] > 1 and pl[4] != 4) else 1  # workspace split-K adds the reduce kernel

    def gemm_plan(self, s: GemmSpec, d=None) -> tuple:
        """(block_n, split_k, cta_group, stages, epilogue) the library will use for this spec."""
        d = self._gemm_desc(s) if d is None else d
        v = [C.c_int32(0) for _ in range(5)]
        _lib.check(self.lib.asva_gemm_plan(d, *[C.byref(x) for x in v]), "asva_gemm_plan")
        return tuple(x.value for x in v)

    def _gemm_desc(self, s: GemmSpec) -> "_lib.GemmDesc":
        d = _lib.GemmDesc()
        for i in range(2):
            av = s.a[i]
            if av is None:
                d.a[i] = None
                continue
            self._chk_dev(av.t)
            assert av.t.dtype == torch.bfloat16
            d.a[i] = av.t.data_ptr()
            for j in range(4):
                d.a_dims[i][j] = av.dims[j]
            for j in range(3):
                d.a_strides[i][j] = av.strides[j]
        for j in range(3):
            d.box[j], d.trav[j], d.out_dims[j] = s.box[j], s.trav[j], s.out_dims[j]
        d.nseg = len(s.segs)
        for i, sg in enumerate(s.segs):
            d.seg[i].src, d.seg[i].c0, d.seg[i].num_kb = sg.src, sg.c0, sg.num_kb
            d.seg[i].wk, d.seg[i].wk_first, d.seg[i].fix2 = sg.wk, sg.wk_first, sg.fix2
            for j in range(3):
                d.seg[i].off[j] = sg.off[j]
        self._chk_dev(s.w, s.out, s.bias)
        assert s.w.dtype == torch.bfloat16
        d.w, d.ldw, d.N, d.K, d.wcols = s.w.data_ptr(), s.ldw, s.N, s.K, s.wcols
        if s.bias is not None:
            assert s.bias.dtype == torch.float32 and s.bias.numel() >= s.N
        d.bias = _ptr(s.bias)
        if s.add is not None:
            self._chk_dev(s.add.t)
            assert s.add.t.dtype == torch.float32
            d.add.ptr, d.add.ld, d.add.div = s.add.t.data_ptr(), s.add.ld, s.add.div
        for i in range(2):
            r = s.res[i]
            if r is not None:
                self._chk_dev(r)
                assert r.dtype == torch.bfloat16
            d.res[i] = _ptr(r)
            d.res_ld[i] = s.res_ld[i]
        d.geglu, d.out_fp32 = int(s.geglu), int(s.out_fp32)
        assert s.out.dtype == (torch.float32 if s.out_fp32 else torch.bfloat16)
        d.out, d.ldo = s.out.data_ptr(), s.ldo
        d.block_n, d.split_k, d.cta_group, d.epilogue = s.block_n, s.split_k, s.cta_group, s.epilogue
        ws = self.splitk_ws()
        d.ws, d.ws_bytes = ws.data_ptr(), ws.numel() * 4
        if s.stats_out is not None:
            self._chk_dev(s.stats_out)
            assert s.stats_out.dtype == torch.float32 and s.stats_out.is_contiguous()
            assert s.N % 32 == 0 and s.stats_out.numel() >= (s.N // 32) * s.M * 2
            d.stats_out = s.stats_out.data_ptr()
        if s.ln is not None:
            f = s.ln
            self._chk_dev(f.stats, f.wsum)
            assert f.stats.dtype == torch.float32 and f.wsum.dtype == torch.float32 and f.wsum.numel() >= s.N
            assert f.cols % 32 == 0 and f.stats.dim() == 3 and f.stats.shape[0] == f.cols // 32 and f.stats.shape[2] == 2
            assert f.stats.is_contiguous()
            d.ln_cols, d.ln_stats, d.ln_wsum = f.cols, f.stats.data_ptr(), f.wsum.data_ptr()
            d.ln_stat_rows, d.ln_grp_rows, d.ln_grp_stride, d.ln_eps = f.stats.shape[1], f.grp_rows, f.grp_stride, f.eps
        return d

    def attention(self, s: AttnSpec) -> None:
        self._chk_dev(s.q, s.kv, s.out, s.mask)
        d = _lib.AttnDesc()
        d.q, d.kv, d.mask, d.out = s.q.data_ptr(), s.kv.data_ptr(), _ptr(s.mask), s.out.data_ptr()
        d.ldq, d.ldkv, d.ldo, d.mask_ld = s.ldq, s.ldkv, s.ldo, s.mask_ld
        d.G, d.heads, d.R, d.Nk, d.d, d.dpad = s.G, s.heads, s.R, s.Nk, s.d, s.dpad
        d.kv_rows_per_group, d.k_col0, d.v_col0, d.mask_rows = s.kv_rows_per_group, s.k_col0, s.v_col0, s.mask_rows
        d.scale, d.form = s.scale, s.form
        with self._timed('attention'):
            _lib.check(self.lib.asva_attention(d, self._stream()), "asva_attention")
        self.launches += 1

    def temporal_attention(self, qkv, out, B, F, N, heads, d, scale, form: int = 0) -> None:
        """form: 0 = auto, 1 = tcgen05, 2 = thread per query, 3 = warp-MMA (asva_temporal_attention_form)."""
        self._chk_dev(qkv, out)
        with self._timed('temporal_attention'):
            _lib.check(self.lib.asva_temporal_attention_form(qkv.data_ptr(), out.data_ptr(), B, F, N, heads, d, scale,
                                                         form, self._stream()), "asva_temporal_attention")
        self.launches += 1

    def layernorm(self, x, gamma, beta, pos, out, M, C, eps, N, F) -> None:
        self._chk_dev(x, gamma, beta, pos, out)
        with self._timed('layernorm'):
            _lib.check(self.lib.asva_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(pos),
                                           out.data_ptr(), M, C, eps, N, F, self._stream()), "asva_layernorm")
        self.launches += 1

    def groupnorm_ws_floats(self, n_inst, rows, C) -> int:
        return int(self.lib.asva_groupnorm_ws_floats(n_inst, rows, C))

    def groupnorm_stats(self, x0, C0, x1, C1, n_inst, rows, groups, eps, gamma, beta, stats, ws) -> None:
        """stats: fp32 [n_inst, C0+C1, 2] <- per-channel (scale, shift)."""
        self._chk_dev(x0, x1, gamma, beta, stats, ws)
        Ct = C0 + (C1 if x1 is not None else 0)
        assert ws.numel() >= self.groupnorm_ws_floats(n_inst, rows, Ct) and stats.numel() >= n_inst * Ct * 2
        with self._timed('groupnorm'):
            _lib.check(self.lib.asva_groupnorm_stats(x0.data_ptr(), C0, _ptr(x1), C1, n_inst, rows, groups, eps,
                                                     gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                                     ws.data_ptr(), self._stream()), "asva_groupnorm_stats")
        self.launches += 2

    def groupnorm(self, x0, C0, x1, C1, n_inst, rows, groups, eps, gamma, beta, silu, out) -> None:
        """Fused statistics + apply (+SiLU) of one GroupNorm: out bf16 [n_inst*rows, C0+C1] (one launch)."""
        self._chk_dev(x0, x1, gamma, beta, out)
        sync = self._gn_sync.get(x0.device)
        if sync is None:  # zeroed once per device; the kernel leaves it zeroed
            sync = torch.zeros(int(self.lib.asva_groupnorm_sync_bytes()), dtype=torch.uint8, device=x0.device)
            self._gn_sync[x0.device] = sync
        with self._timed('groupnorm'):
            _lib.check(self.lib.asva_groupnorm(x0.data_ptr(), C0, _ptr(x1), C1, n_inst, rows, groups, eps,
                                               gamma.data_ptr(), beta.data_ptr(), int(silu), out.data_ptr(),
                                               sync.data_ptr(), self._stream()), "asva_groupnorm")
        Ct = C0 + (C1 if x1 is not None else 0)
        self.launches += 2 if self.lib.asva_groupnorm_form(n_inst, rows, Ct, groups) == 2 else 1

    def groupnorm_apply(self, x0, C0, x1, C1, stats, n_inst, n_img, h, w, silu, upsample, out) -> None:
        self._chk_dev(x0, x1, stats, out)
        with self._timed('groupnorm'):
            _lib.check(self.lib.asva_groupnorm_apply(x0.data_ptr(), C0, _ptr(x1), C1, _ptr(stats), n_inst, n_img, h,
                                                     w, int(silu), int(upsample), out.data_ptr(), self._stream()),
                       "asva_groupnorm_apply")
        self.launches += 1

    def conv_in_im2col(self, lat, out, B, Bs, Cl, F, h, w) -> None:
        self._chk_dev(lat, out)
        assert lat.dtype == torch.float32 and lat.is_contiguous()
        with self._timed('misc'):
            _lib.check(self.lib.asva_conv_in_im2col(lat.data_ptr(), out.data_ptr(), B, Bs, Cl, F, h, w, self._stream()),
                   "asva_conv_in_im2col")
        self.launches += 1

    def tconv_gather(self, y, out, B, F, N, C) -> None:
        self._chk_dev(y, out)
        with self._timed('misc'):
            _lib.check(self.lib.asva_tconv_gather(y.data_ptr(), out.data_ptr(), B, F, N, C, self._stream()),
                       "asva_tconv_gather")
        self.launches += 1

    def conv_out_finish(self, y, ldy, wt, bt, out, B, Co, F, h, w) -> None:
        self._chk_dev(y, wt, bt, out)
        with self._timed('misc'):
            _lib.check(self.lib.asva_conv_out_finish(y.data_ptr(), ldy, wt.data_ptr(), bt.data_ptr(), out.data_ptr(), B,
                                                 Co, F, h, w, self._stream()), "asva_conv_out_finish")
        self.launches += 1

    def small_linear(self, x, w, bias, out, M, N, K, act_in, act_out) -> None:
        self._chk_dev(x, w, bias, out)
        with self._timed('small_linear'):
            _lib.check(self.lib.asva_small_linear(x.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), M, N, K,
                                              act_in, act_out, self._stream()), "asva_small_linear")
        self.launches += 1

    def timestep_features(self, t, out, B, dim, flip) -> None:
        self._chk_dev(t, out)
        with self._timed('misc'):
            _lib.check(self.lib.asva_timestep_features(t.data_ptr(), out.data_ptr(), B, dim, int(flip), self._stream()),
                   "asva_timestep_features")
        self.launches += 1

    def softmax_rows(self, scores, probs, rows, cols, scale) -> None:
        """probs bf16 [rows, cols] = softmax(scale * scores fp32 [rows, cols]) (row views with contiguous columns)."""
        self._chk_dev(scores, probs)
        assert scores.dtype == torch.float32 and probs.dtype == torch.bfloat16
        with self._timed('misc'):
            _lib.check(self.lib.asva_softmax_rows(scores.data_ptr(), scores.stride(0), probs.data_ptr(), probs.stride(0),
                                                  rows, cols, scale, self._stream()), "asva_softmax_rows")
        self.launches += 1

    def cfg_ddim_step(self, eps, k, lat, coef, C, F, hw, clips: int = 1) -> None:
        """eps fp32 (k*clips, C, F, hw) branch-major; lat fp32 (clips, C, F, hw)."""
        self._chk_dev(eps, lat, coef)
        assert eps.numel() == k * clips * C * F * hw and lat.numel() == clips * C * F * hw
        with self._timed('cfg_step'):
            _lib.check(self.lib.asva_cfg_ddim_step(eps.data_ptr(), k, clips, lat.data_ptr(), coef.data_ptr(), C, F, hw,
                                               self._stream()), "asva_cfg_ddim_step")
        self.launches += 1

    def cfg_plms_step(self, eps, k, lat, hist, coef, slots, C, F, hw, clips: int = 1) -> None:
        self._chk_dev(eps, lat, hist, coef, slots)
        assert eps.numel() == k * clips * C * F * hw and hist.numel() == 4 * clips * C * F * hw
        with self._timed('cfg_step'):
            _lib.check(self.lib.asva_cfg_plms_step(eps.data_ptr(), k, clips, lat.data_ptr(), hist.data_ptr(),
                                               coef.data_ptr(), slots.data_ptr(), C, F, hw, self._stream()),
                       "asva_cfg_plms_step")
        self.launches += 1


_backend = None


def backend():
    """The process-wide operator backend (CUDA). Raises if libasva_b200.so is unavailable."""
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def _set_backend_for_tests(b) -> None:
    """Test hook: tests/ installs the torch simulator here to check descriptor logic on CPU."""
    global _backend
    _backend = b
