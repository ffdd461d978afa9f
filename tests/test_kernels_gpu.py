"""Per-kernel parity on the B200: every C-ABI entry point of include/asva_b200.h against the torch spec
interpreter (tests/sim_backend.py) on identical seeded inputs.  Tolerances: bf16 outputs rel-L2 <= 4e-3 for GEMMs
(fp32 accumulation on both sides, one bf16 rounding), <= 1e-2 for attention (P is rounded to bf16 before the
PV product in the kernel); fp32 outputs <= 1e-5 (elementwise) / 2e-3 (bf16-weight skinny GEMM)."""
import dataclasses
import math

import pytest
import torch

from asva_b200 import ops
from sim_backend import SimBackend

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def _report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    assert torch.isfinite(got).all(), f"{name}: non-finite output"
    err = (got - ref).norm() / (ref.norm() + 1e-20)
    mx = (got - ref).abs().max()
    if not (err <= tol):
        bad = ((got - ref).abs() > 4 * tol * ref.abs().max()).nonzero()
        rows = bad[:, 0].unique()[:16].tolist() if bad.numel() and bad.dim() == 2 else []
        cols = bad[:, 1].unique()[:16].tolist() if bad.numel() and bad.dim() == 2 and bad.shape[1] > 1 else []
        pytest.fail(f"{name}: rel-L2 {err:.3e} > {tol:.1e}, max|d| {mx:.3e}, |ref|max {ref.abs().max():.3e}, "
                    f"{bad.shape[0]} bad elems; first bad rows {rows} cols {cols}; shape {tuple(ref.shape)}")
    return float(err)


def _run_gemm_pair(be, spec, out_shape, out_dtype, fill=0.0):
    o_ref = torch.full(out_shape, fill, dtype=out_dtype, device=DEV)
    o_cu = torch.full(out_shape, fill, dtype=out_dtype, device=DEV)
    SimBackend().gemm(dataclasses.replace(spec, out=o_ref))
    be.gemm(dataclasses.replace(spec, out=o_cu))
    torch.cuda.synchronize()
    return o_cu, o_ref


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,K,N,bn", [
    (256, 64, 64, 0), (128, 320, 320, 0), (300, 320, 320, 0), (1000, 640, 640, 0), (384, 1280, 1280, 0),
    (256, 768, 640, 64), (256, 768, 640, 128), (200, 1280, 2560, 0), (24576, 320, 320, 0), (77, 768, 2560, 0),
    (16, 1280, 1280, 0),
])
@pytest.mark.parametrize("epi", [1, 2, 3])
def test_gemm_linear(cuda_backend, M, K, N, bn, epi):
    x = _rand((M, K), 1)
    w = _rand((N, K), 2, 1.0 / math.sqrt(K))
    bias = _rand((N,), 3, dtype=torch.float32)
    res = _rand((M, N), 4)
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=torch.bfloat16, device=DEV), bias=bias, res0=res)
    spec.block_n, spec.epilogue = bn, epi
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), torch.bfloat16, fill=7.0)
    _report(f"linear {M}x{K}x{N} epi{epi}", got, ref, 4e-3)


@pytest.mark.parametrize("M,K,N,bn,nres", [(1000, 640, 640, 128, 1), (1100, 320, 320, 160, 2), (2304, 1280, 1288, 256, 0),
                                           (130, 640, 320, 64, 2)])
def test_gemm_cta_pair_direct_epilogue(cuda_backend, M, K, N, bn, nres):
    # CTA pairs (odd tile counts: the last pair's second tile does not exist) with the per-warp epilogue; N = 1288
    # leaves a panel with a single 8-column chunk
    x = _rand((M, K), 71)
    w = _rand((N, K), 72, 1.0 / math.sqrt(K))
    bias = _rand((N,), 73, dtype=torch.float32)
    r0, r1 = (_rand((M, N), 74) if nres > 0 else None), (_rand((M, N), 75) if nres > 1 else None)
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=torch.bfloat16, device=DEV), bias=bias, res0=r0, res1=r1)
    spec.block_n, spec.split_k, spec.cta_group, spec.epilogue = bn, 1, 2, 2
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), torch.bfloat16, fill=7.0)
    pl = cuda_backend.gemm_plan(spec)
    assert pl[2] == 2 and pl[4] == 2, pl
    _report(f"pair+direct {M}x{K}x{N} bn{bn}", got, ref, 4e-3)


def test_gemm_linear_fp32_out_no_epilogue(cuda_backend):
    M, K, N = 512, 640, 320
    x, w = _rand((M, K), 5), _rand((N, K), 6, 1.0 / math.sqrt(K))
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=torch.float32, device=DEV), out_fp32=True)
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), torch.float32)
    _report("linear fp32", got, ref, 1e-5)


def test_gemm_two_sources(cuda_backend):
    M, K0, K1, N = 512, 640, 320, 640
    x0, x1 = _rand((M, K0), 7), _rand((M, K1), 8)
    w = _rand((N, K0 + K1), 9, 1.0 / math.sqrt(K0 + K1))
    spec = ops.spec_linear(x0, w, torch.empty(M, N, dtype=torch.bfloat16, device=DEV), x2=x1)
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), torch.bfloat16)
    _report("linear 2-source", got, ref, 4e-3)


# ---- cluster split-K (epilogue form 4)
@pytest.mark.parametrize("M,K,N,bn,split,nres,f32", [
    (384, 3840, 1280, 256, 8, 2, False), (384, 3840, 1280, 128, 4, 1, False), (384, 1280, 1280, 64, 2, 0, False),
    (384, 11520, 1280, 256, 8, 0, False), (130, 640, 320, 64, 2, 2, False), (1536, 1280, 1288, 128, 4, 1, False),
    (32, 1280, 2560, 256, 4, 0, True), (300, 704, 200, 128, 2, 1, False), (128, 5120, 1280, 160, 5, 0, False),
])
def test_gemm_cluster_split_k(cuda_backend, M, K, N, bn, split, nres, f32):
    """K split over the CTAs of a thread-block cluster (partials in tensor memory, slices exchanged through distributed
    shared memory): every cluster size, ragged M / N tails, bias + residuals, fp32 outputs; 160 / 5 is not a feasible
    slice split and must fall back to another epilogue form."""
    x = _rand((M, K), 321)
    w = _rand((N, K), 322, 1.0 / math.sqrt(K))
    bias = _rand((N,), 323, dtype=torch.float32)
    r0, r1 = (_rand((M, N), 324) if nres > 0 else None), (_rand((M, N), 325) if nres > 1 else None)
    dt = torch.float32 if f32 else torch.bfloat16
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=dt, device=DEV), bias=bias, res0=r0, res1=r1, out_fp32=f32)
    spec.block_n, spec.split_k, spec.cta_group, spec.epilogue = bn, split, 1, 4
    pl = cuda_backend.gemm_plan(spec)
    if split == 5:
        assert pl[4] != 4, pl
        return
    assert pl[0] == bn and pl[1] == split and pl[2] == 1 and pl[4] == 4, pl
    n0 = cuda_backend.launches
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), dt, fill=7.0)
    assert cuda_backend.launches - n0 == 1  # no reduce kernel
    _report(f"cluster split-K {M}x{K}x{N} bn{bn} S{split}", got, ref, 2e-5 if f32 else 4e-3)


def test_gemm_cluster_split_k_conv(cuda_backend):
    # the level-3 3x3 conv (4 x 4 images, K = 9 * 1280): implicit-GEMM boxes through the same path
    n_img, h, wd, ci, co = 24, 4, 4, 1280, 1280
    x = _rand((n_img * h * wd, ci), 331)
    wt = _rand((co, 9 * ci), 332, 1.0 / math.sqrt(9 * ci))
    b = _rand((co,), 333, dtype=torch.float32)
    spec = ops.spec_conv3x3(x, wt, torch.empty(n_img * h * wd, co, dtype=torch.bfloat16, device=DEV), n_img=n_img, h=h,
                            wd=wd, bias=b)
    spec.block_n, spec.split_k, spec.cta_group, spec.epilogue = 256, 8, 1, 4
    assert cuda_backend.gemm_plan(spec)[4] == 4
    got, ref = _run_gemm_pair(cuda_backend, spec, (n_img * h * wd, co), torch.bfloat16, fill=7.0)
    _report("cluster split-K conv3x3 L3", got, ref, 4e-3)


# ---- row statistics + LayerNorm fold (asva_gemm_desc.stats_out / ln_*, ops.LnFold)
def _stats_of(x32, C):
    v = x32.view(x32.shape[0], C // 32, 32)
    return torch.stack([v.sum(-1), (v * v).sum(-1)], dim=-1).permute(1, 0, 2).contiguous()


@pytest.mark.parametrize("M,K,N,bn,cg,split,epi", [
    (256, 320, 320, 0, 1, 1, 1), (256, 320, 320, 0, 1, 1, 3), (1000, 640, 640, 128, 2, 1, 3), (1000, 640, 640, 128, 2, 1, 1),
    (300, 320, 320, 160, 1, 1, 3), (24576, 320, 320, 160, 1, 1, 3), (384, 5120, 1280, 64, 1, 2, 3),
    (384, 5120, 1280, 128, 1, 4, 1), (1536, 1280, 1280, 256, 2, 1, 3), (16, 1280, 1280, 64, 1, 1, 3),
    (1536, 1280, 1280, 128, 1, 1, 2),
])
def test_gemm_row_stats_out(cuda_backend, M, K, N, bn, cg, split, epi):
    """The producer side: (sum, sum of squares) per row and 32-column slot of the fp32 values the GEMM stores, in
    every epilogue form that carries it (a request for form 2 is served by 3), the split-K reduce kernel included."""
    x = _rand((M, K), 301)
    w = _rand((N, K), 302, 1.0 / math.sqrt(K))
    bias = _rand((N,), 303, dtype=torch.float32)
    res = _rand((M, N), 304)
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=torch.bfloat16, device=DEV), bias=bias, res0=res)
    spec.block_n, spec.cta_group, spec.split_k, spec.epilogue = bn, cg, split, epi
    st_ref = torch.zeros(N // 32, M, 2, dtype=torch.float32, device=DEV)
    st_cu = torch.full((N // 32, M, 2), 123.0, dtype=torch.float32, device=DEV)
    o_ref = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    o_cu = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    SimBackend().gemm(dataclasses.replace(spec, out=o_ref, stats_out=st_ref))
    sp = dataclasses.replace(spec, out=o_cu, stats_out=st_cu)
    pl = cuda_backend.gemm_plan(sp)
    assert pl[4] != 2 and pl[1] == split, pl
    cuda_backend.gemm(sp)
    torch.cuda.synchronize()
    _report(f"stats-out gemm {M}x{K}x{N}", o_cu, o_ref, 4e-3)
    _report(f"row sums {M}x{K}x{N} plan {pl}", st_cu[..., 0], st_ref[..., 0], 1e-4)
    _report(f"row sums of squares {M}x{K}x{N} plan {pl}", st_cu[..., 1], st_ref[..., 1], 1e-5)


@pytest.mark.parametrize("M,C,N,bn,cg,epi,bias", [
    (256, 320, 320, 0, 1, 1, False), (256, 320, 320, 0, 1, 3, False), (24576, 320, 320, 160, 1, 3, False),
    (24576, 320, 320, 160, 1, 1, True), (1000, 640, 640, 128, 2, 3, True), (1000, 640, 1280, 256, 2, 1, False),
    (1536, 1280, 1280, 128, 1, 3, False), (300, 320, 960, 64, 1, 3, True), (16, 1280, 1280, 64, 1, 0, False),
])
def test_gemm_ln_fold(cuda_backend, M, C, N, bn, cg, epi, bias):
    """The consumer side against LayerNorm -> projection computed directly: rows with a large common offset (mean /
    std ~ 3) so the mean term matters."""
    g = torch.Generator(device="cpu").manual_seed(311)
    x32 = (torch.randn(M, C, generator=g) * (0.5 + torch.rand(M, 1, generator=g)) + 3.0 * torch.randn(M, 1, generator=g))
    x = x32.to(torch.bfloat16).to(DEV)
    gamma = (1.0 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    w32 = (torch.randn(N, C, generator=g) / math.sqrt(C)).to(DEV)
    b32 = torch.randn(N, generator=g).to(DEV) if bias else None
    wg = (w32 * gamma.view(1, -1)).to(torch.bfloat16).contiguous()
    wsum = wg.float().sum(1).contiguous()
    fb = w32 @ beta + (b32 if bias else 0.0)
    st = _stats_of(x.float(), C)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    spec = ops.spec_linear(x, wg, out, bias=fb.contiguous())
    spec.ln = ops.LnFold(stats=st, wsum=wsum, cols=C, eps=1e-5)
    spec.block_n, spec.cta_group, spec.epilogue = bn, cg, epi
    pl = cuda_backend.gemm_plan(spec)
    assert pl[1] == 1 and pl[4] in (1, 3) and (epi == 0 or pl[4] == epi), pl
    cuda_backend.gemm(spec)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ w32.t() + (b32 if bias else 0.0)
    sim = torch.zeros_like(out)
    SimBackend().gemm(dataclasses.replace(spec, out=sim))
    _report(f"ln-fold vs sim {M}x{C}x{N} plan {pl}", out, sim, 4e-3)
    # against the unfolded computation: bf16 rounding of W*gamma and of the output
    _report(f"ln-fold vs LayerNorm->linear {M}x{C}x{N}", out, ref, 6e-3)


@pytest.mark.parametrize("epi", [1, 3])
def test_gemm_ln_fold_geglu(cuda_backend, epi):
    M, C = 384, 640
    g = torch.Generator(device="cpu").manual_seed(312)
    x = (torch.randn(M, C, generator=g) + 2.0 * torch.randn(M, 1, generator=g)).to(torch.bfloat16).to(DEV)
    gamma, beta = (1.0 + 0.2 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    w32 = (torch.randn(8 * C, C, generator=g) / math.sqrt(C)).to(DEV)
    b32 = torch.randn(8 * C, generator=g).to(DEV)
    wg = (w32 * gamma.view(1, -1)).to(torch.bfloat16).contiguous()
    out = torch.zeros(M, 4 * C, dtype=torch.bfloat16, device=DEV)
    spec = ops.spec_linear(x, wg, out, bias=(w32 @ beta + b32).contiguous(), geglu=True)
    spec.ln = ops.LnFold(stats=_stats_of(x.float(), C), wsum=wg.float().sum(1).contiguous(), cols=C, eps=1e-5)
    spec.epilogue = epi
    assert cuda_backend.gemm_plan(spec)[4] == epi
    cuda_backend.gemm(spec)
    torch.cuda.synchronize()
    y = (torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ w32.t() + b32).view(M, -1, 2, 64)
    ref = (y[:, :, 0] * torch.nn.functional.gelu(y[:, :, 1])).reshape(M, 4 * C)
    sim = torch.zeros_like(out)
    SimBackend().gemm(dataclasses.replace(spec, out=sim))
    _report(f"ln-fold geglu vs sim epi{epi}", out, sim, 4e-3)
    _report(f"ln-fold geglu vs LayerNorm->GEGLU epi{epi}", out, ref, 8e-3)


def test_gemm_ln_fold_grouped_rows_chain(cuda_backend):
    """Producer -> consumer on the device, with the consumer reading the statistics of a strided row subset (the
    frame-0 rows feeding attn1's K/V projection): t = x W0^T + b (+ statistics), kv = LayerNorm(t)[frame 0] Wkv^T."""
    B, F, N, C = 2, 3, 64, 320
    M = B * F * N
    g = torch.Generator(device="cpu").manual_seed(313)
    x = _rand((M, C), 314)
    w0 = _rand((C, C), 315, 1.0 / math.sqrt(C))
    b0 = (2.0 * torch.randn(C, generator=g)).to(DEV)
    gamma, beta = (1.0 + 0.2 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    wkv32 = (torch.randn(2 * C, C, generator=g) / math.sqrt(C)).to(DEV)
    t = torch.zeros(M, C, dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(C // 32, M, 2, dtype=torch.float32, device=DEV)
    sp = ops.spec_linear(x, w0, t, bias=b0)
    sp.stats_out = st
    cuda_backend.gemm(sp)
    wg = (wkv32 * gamma.view(1, -1)).to(torch.bfloat16).contiguous()
    kv = torch.zeros(B * N, 2 * C, dtype=torch.bfloat16, device=DEV)
    av = ops.AView(t, (C, N, B, 1), (C, F * N * C, B * F * N * C))
    sk = ops.spec_rows3(av, (N, B, 1), wg, kv, bias=(wkv32 @ beta).contiguous())
    sk.ln = ops.LnFold(stats=st, wsum=wg.float().sum(1).contiguous(), cols=C, eps=1e-5, grp_rows=N, grp_stride=F * N)
    cuda_backend.gemm(sk)
    torch.cuda.synchronize()
    t0 = t.view(B, F, N, C)[:, 0].reshape(B * N, C).float()
    ref = torch.nn.functional.layer_norm(t0, (C,), gamma, beta, 1e-5) @ wkv32.t()
    _report("chained ln-fold on frame-0 rows", kv, ref, 6e-3)


@pytest.mark.parametrize("M,C", [(256, 320), (384, 640), (200, 1280)])
@pytest.mark.parametrize("epi", [1, 3])
def test_gemm_geglu(cuda_backend, M, C, epi):
    x = _rand((M, C), 10)
    w = _rand((8 * C, C), 11, 1.0 / math.sqrt(C))
    bias = _rand((8 * C,), 12, dtype=torch.float32)
    spec = ops.spec_linear(x, w, torch.empty(M, 4 * C, dtype=torch.bfloat16, device=DEV), bias=bias, geglu=True)
    spec.epilogue = epi
    assert cuda_backend.gemm_plan(spec)[4] == epi
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, 4 * C), torch.bfloat16)
    _report(f"geglu {M}x{C}", got, ref, 4e-3)


@pytest.mark.parametrize("n_img,h,w,Cin,Cout,stride", [
    (2, 16, 16, 64, 64, 1), (3, 32, 32, 320, 320, 1), (5, 8, 8, 640, 1280, 1), (24, 4, 4, 1280, 1280, 1),
    (2, 16, 32, 320, 640, 1), (3, 32, 32, 320, 320, 2), (4, 8, 8, 1280, 1280, 2), (2, 6, 10, 128, 64, 1),
    (2, 16, 16, 960, 320, 1),
])
@pytest.mark.parametrize("epi", [1, 2, 3])
def test_gemm_conv3x3(cuda_backend, n_img, h, w, Cin, Cout, stride, epi):
    x = _rand((n_img * h * w, Cin), 13)
    wt = _rand((Cout, 9 * Cin), 14, 1.0 / math.sqrt(9 * Cin))
    bias = _rand((Cout,), 15, dtype=torch.float32)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    spec = ops.spec_conv3x3(x, wt, torch.empty(n_img * ho * wo, Cout, dtype=torch.bfloat16, device=DEV),
                            n_img=n_img, h=h, wd=w, stride=stride, bias=bias)
    spec.epilogue = epi
    got, ref = _run_gemm_pair(cuda_backend, spec, (n_img * ho * wo, Cout), torch.bfloat16)
    _report(f"conv3x3 {n_img}x{h}x{w} {Cin}->{Cout} s{stride}", got, ref, 4e-3)
    # and against torch's own conv2d (checks the K ordering convention of the spec, not only sim == cuda)
    xi = x.float().view(n_img, h, w, Cin).permute(0, 3, 1, 2)
    wi = wt.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(xi, wi, bias, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    _report("conv3x3 vs F.conv2d", got, y, 6e-3)


@pytest.mark.parametrize("B,F,N,C", [(2, 12, 64, 320), (1, 8, 16, 1280), (2, 5, 100, 640), (2, 3, 256, 320),
                                     (3, 4, 16, 640)])
@pytest.mark.parametrize("epi", [1, 2, 3])
def test_gemm_tconv(cuda_backend, B, F, N, C, epi):
    y = _rand((B * F * N, C), 16)
    w3 = _rand((C, 3 * C), 17, 0.02)
    bt = _rand((C,), 18, 0.1, dtype=torch.float32)
    tproj = _rand((B, C), 19, dtype=torch.float32)
    res1 = _rand((B * F * N, C), 20)
    wh, wp, wc = w3[:, :C], w3[:, C:2 * C], w3[:, 2 * C:]
    w4 = torch.cat([wc, wp, wh, (wh.float() + wp.float()).to(torch.bfloat16)], dim=1).contiguous()
    spec = ops.spec_tconv(y, w4, torch.empty(B * F * N, C, dtype=torch.bfloat16, device=DEV), B=B, F=F, N=N,
                          bias=bt, tproj=tproj, tproj_ld=C, res1=res1)
    spec.epilogue = epi
    got, ref = _run_gemm_pair(cuda_backend, spec, (B * F * N, C), torch.bfloat16)
    _report(f"tconv {B}x{F}x{N}x{C}", got, ref, 4e-3)
    # direct restatement of FFInflatedConv3d's temporal part (utils.py:43-53) + tproj + extra residual
    yf = y.float().view(B, F, N, C)
    prev = torch.cat([yf[:, :1], yf[:, :-1]], dim=1)
    cat = torch.cat([yf[:, :1].expand_as(yf), prev, yf], dim=3)
    want = yf + cat @ w3.float().t() + bt + tproj.view(B, 1, 1, C) + res1.float().view(B, F, N, C)
    _report("tconv vs restatement", got, want.reshape(B * F * N, C), 8e-3)


@pytest.mark.parametrize("M,K,N,split,bn", [(384, 11520, 1280, 5, 128), (1536, 5760, 1280, 2, 128), (384, 2560, 1280, 4, 64),
                                            (200, 1280, 640, 3, 0), (130, 640, 320, 2, 160), (384, 2304, 1280, 0, 0)])
@pytest.mark.parametrize("epi", [1, 3])
def test_gemm_split_k(cuda_backend, M, K, N, split, bn, epi):
    x = _rand((M, K), 51)
    w = _rand((N, K), 52, 1.0 / math.sqrt(K))
    bias = _rand((N,), 53, dtype=torch.float32)
    res = _rand((M, N), 54)
    spec = ops.spec_linear(x, w, torch.empty(M, N, dtype=torch.bfloat16, device=DEV), bias=bias, res0=res)
    spec.block_n, spec.split_k, spec.epilogue = bn, split, epi
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, N), torch.bfloat16)
    _report(f"split-k {M}x{K}x{N} /{split} epi{epi}", got, ref, 4e-3)


@pytest.mark.parametrize("bn", [64, 128, 160, 256])
@pytest.mark.parametrize("epi", [1, 2, 3])
def test_gemm_block_n_and_inplace_residual(cuda_backend, bn, epi):
    # every tile width, N not a multiple of it, two residuals one of which aliases the output (t = t + ...)
    M, K, N = 1000, 640, 960
    x = _rand((M, K), 55)
    w = _rand((N, K), 56, 1.0 / math.sqrt(K))
    bias = _rand((N,), 57, dtype=torch.float32)
    t0 = _rand((M, N), 58)
    r1 = _rand((M, N), 59)
    t_ref, t_cu = t0.clone(), t0.clone()
    s_ref = ops.spec_linear(x, w, t_ref, bias=bias, res0=t_ref, res1=r1)
    s_cu = ops.spec_linear(x, w, t_cu, bias=bias, res0=t_cu, res1=r1)
    s_cu.block_n, s_cu.split_k, s_cu.epilogue = bn, 1, epi
    SimBackend().gemm(s_ref)
    cuda_backend.gemm(s_cu)
    torch.cuda.synchronize()
    _report(f"in-place residual bn={bn}", t_cu, t_ref, 4e-3)


def test_gemm_strided_out_and_narrow(cuda_backend):
    # output written into a column slice of a wider buffer; N = 8 (conv_out-like) in fp32
    M, K = 300, 320
    x = _rand((M, K), 60)
    w = _rand((320, K), 61, 1.0 / math.sqrt(K))
    wide_ref = torch.zeros(M, 1024, dtype=torch.bfloat16, device=DEV)
    wide_cu = torch.zeros_like(wide_ref)
    SimBackend().gemm(ops.spec_linear(x, w, wide_ref[:, 320:640]))
    cuda_backend.gemm(ops.spec_linear(x, w, wide_cu[:, 320:640]))
    torch.cuda.synchronize()
    _report("strided out", wide_cu, wide_ref, 4e-3)
    assert (wide_cu[:, :320] == 0).all() and (wide_cu[:, 640:] == 0).all()
    w8 = _rand((8, K), 62, 1.0 / math.sqrt(K))
    b8 = _rand((8,), 63, dtype=torch.float32)
    spec = ops.spec_linear(x, w8, torch.empty(M, 8, dtype=torch.float32, device=DEV), bias=b8, out_fp32=True)
    got, ref = _run_gemm_pair(cuda_backend, spec, (M, 8), torch.float32)
    _report("narrow fp32", got, ref, 1e-5)


# ------------------------------------------------------------------------------------------------ attention
def _attn_case(be, G, heads, R, Nk, d, masked, seed, form=0):
    dpad = ((d + 63) // 64) * 64
    C = heads * d
    q = _rand((G * R, C), seed)
    kv = _rand((G * Nk, 2 * C), seed + 1)
    mask = None
    mask_rows = 1
    if masked:
        # per (group, frame) key windows, like the audio segment masks (segmask_imagebind.py:104-114)
        nf = 4 if R % 4 == 0 else 1
        mask_rows = R // nf
        mask = torch.zeros(G * nf, Nk, dtype=torch.uint8, device=DEV)
        for i in range(G * nf):
            a = (i * 7) % max(1, Nk - 5)
            mask[i, 0] = 1
            mask[i, a:a + 5] = 1
    out_ref = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
    out_cu = torch.zeros_like(out_ref)
    spec = ops.AttnSpec(q=q, kv=kv, out=out_ref, G=G, heads=heads, R=R, Nk=Nk, d=d, dpad=dpad, ldq=C, ldkv=2 * C, ldo=C,
                        kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d), mask=mask,
                        mask_ld=Nk, mask_rows=mask_rows)
    SimBackend().attention(spec)
    be.attention(dataclasses.replace(spec, out=out_cu, form=form))
    torch.cuda.synchronize()
    _report(f"attention G{G} R{R} Nk{Nk} d{d} mask{masked} form{form}", out_cu, out_ref, 1e-2)


@pytest.mark.parametrize("d", [40, 80, 160])
@pytest.mark.parametrize("R,Nk,masked", [(128, 64, False), (256, 256, False), (1024, 1024, False), (200, 77, False),
                                          (512, 229, True), (64, 16, False), (16, 16, False), (384, 1, False)])
def test_attention(cuda_backend, d, R, Nk, masked):
    _attn_case(cuda_backend, 2, 8, R, Nk, d, masked, 30 + d)


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("G,heads,R,Nk,d,masked", [
    (2, 8, 300, 77, 40, False), (3, 8, 128, 25, 80, True), (2, 8, 768, 64, 160, False), (2, 8, 50, 1, 40, False),
    (1, 8, 1000, 128, 40, False), (2, 8, 256, 100, 16, True), (24, 8, 64, 25, 160, True), (2, 8, 12288, 77, 40, False),
    (4, 8, 256, 7, 80, True), (2, 3, 77, 33, 8, False), (2, 10, 260, 81, 64, False), (1, 8, 16, 16, 160, False),
])
def test_attention_small_key_sets_both_forms(cuda_backend, G, heads, R, Nk, d, masked, form):
    """The cross-attention / low-resolution shapes (Nk <= 128) through the tcgen05 kernel (form 1) and the warp-MMA
    kernel (form 2): ragged row and key counts, head chunks (d = 80 / 160 split the heads over CTAs), per-frame key
    masks, one key."""
    _attn_case(cuda_backend, G, heads, R, Nk, d, masked, 140 + Nk + d, form=form)


def test_attention_mma_all_keys_masked_rows_give_zero(cuda_backend):
    G, heads, R, Nk, d = 2, 8, 64, 20, 40
    C = heads * d
    q, kv = _rand((G * R, C), 170), _rand((G * Nk, 2 * C), 171)
    mask = torch.ones(G * 2, Nk, dtype=torch.uint8, device=DEV)
    mask[1] = 0  # the second half of group 0 may attend nothing
    out = torch.full((G * R, C), 3.0, dtype=torch.bfloat16, device=DEV)
    spec = ops.AttnSpec(q=q, kv=kv, out=out, G=G, heads=heads, R=R, Nk=Nk, d=d, dpad=64, ldq=C, ldkv=2 * C, ldo=C,
                        kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d), mask=mask, mask_ld=Nk,
                        mask_rows=R // 2, form=2)
    cuda_backend.attention(spec)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert float(out[R // 2:R].float().abs().max()) == 0.0
    assert float(out[:R // 2].float().abs().max()) > 0.0


def test_attention_config4_first_frame_shape(cuda_backend):
    """BASELINE config 4's level-0 first-frame attention: per CFG branch 24 frames x 4 096 queries against the 4 096
    keys of frame 0 (G 2, R 98 304, Nk 4 096, d 40) - 64 key tiles per query tile, 12 288 query tiles.  The reference
    is evaluated in fp32 in row chunks (the full score tensor would be 26 GB)."""
    G, heads, R, Nk, d = 2, 8, 98304, 4096, 40
    C = heads * d
    q = _rand((G * R, C), 45)
    kv = _rand((G * Nk, 2 * C), 46)
    out = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
    spec = ops.AttnSpec(q=q, kv=kv, out=out, G=G, heads=heads, R=R, Nk=Nk, d=d, dpad=64, ldq=C, ldkv=2 * C, ldo=C,
                        kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d))
    cuda_backend.attention(spec)
    torch.cuda.synchronize()
    ref = torch.empty_like(out)
    k = kv.view(G, Nk, 2 * C)[:, :, :C].reshape(G, Nk, heads, d).permute(0, 2, 1, 3).float()
    v = kv.view(G, Nk, 2 * C)[:, :, C:].reshape(G, Nk, heads, d).permute(0, 2, 1, 3).float()
    for r0 in range(0, R, 8192):
        qc = q.view(G, R, heads, d)[:, r0:r0 + 8192].permute(0, 2, 1, 3).float()
        oc = torch.nn.functional.scaled_dot_product_attention(qc, k, v)
        ref.view(G, R, heads, d)[:, r0:r0 + 8192] = oc.permute(0, 2, 1, 3).to(torch.bfloat16)
    _report("attention config-4 level 0", out, ref, 1e-2)


def test_attention_large_scores(cuda_backend):
    # online-softmax rescale path: later key tiles carry the maxima
    G, heads, R, Nk, d = 1, 2, 128, 512, 40
    dpad, C = 64, heads * d
    q = _rand((G * R, C), 40, 2.0)
    kv = _rand((G * Nk, 2 * C), 41)
    kv[:, :C] *= torch.linspace(0.2, 3.0, Nk, device=DEV).view(-1, 1).to(torch.bfloat16)
    out_ref = torch.zeros(G * R, C, dtype=torch.bfloat16, device=DEV)
    out_cu = torch.zeros_like(out_ref)
    spec = ops.AttnSpec(q=q, kv=kv, out=out_ref, G=G, heads=heads, R=R, Nk=Nk, d=d, dpad=dpad, ldq=C, ldkv=2 * C, ldo=C,
                        kv_rows_per_group=Nk, k_col0=0, v_col0=C, scale=1.0 / math.sqrt(d))
    SimBackend().attention(spec)
    cuda_backend.attention(dataclasses.replace(spec, out=out_cu))
    torch.cuda.synchronize()
    _report("attention rescale", out_cu, out_ref, 1e-2)


@pytest.mark.parametrize("B,F,N,heads,d", [(2, 12, 64, 8, 40), (1, 8, 16, 8, 160), (2, 24, 33, 8, 80), (1, 1, 8, 8, 40),
                                           (2, 12, 1024, 8, 40), (2, 12, 256, 8, 80), (2, 12, 16, 8, 160),
                                           (1, 16, 35, 8, 40), (1, 24, 5, 8, 160), (1, 32, 9, 4, 40), (1, 40, 6, 8, 40),
                                           (3, 7, 1, 2, 8)])
@pytest.mark.parametrize("form", [0, 1, 2, 3])
def test_temporal_attention(cuda_backend, B, F, N, heads, d, form):
    """asva_temporal_attention (auto) and every form by name: tcgen05 (1), thread per query (2), warp-MMA (3) - the
    last two stream pixel groups of 1-4 (ragged last group) and serve F <= 32."""
    if form in (2, 3) and F > 32:
        pytest.skip("the shared-memory forms serve F <= 32")
    C = heads * d
    qkv = _rand((B, F, N, 3 * C), 50)
    o_ref = torch.zeros(B, F, N, C, dtype=torch.bfloat16, device=DEV)
    o_cu = torch.full_like(o_ref, 7.0)
    SimBackend().temporal_attention(qkv, o_ref, B, F, N, heads, d, 1.0 / math.sqrt(d))
    cuda_backend.temporal_attention(qkv, o_cu, B, F, N, heads, d, 1.0 / math.sqrt(d), form=form)
    torch.cuda.synchronize()
    _report("temporal attention", o_cu.view(-1, C), o_ref.view(-1, C), 4e-3)


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("M,C,with_pos", [(1000, 320, False), (513, 640, True), (96, 1280, True), (24576, 320, True),
                                          (6144, 640, False), (1537, 1280, False), (7, 320, False), (301, 64, True),
                                          (130, 128, False), (77, 256, True), (50, 768, False), (33, 1024, True),
                                          (21, 2048, False), (40, 72, False)])
def test_layernorm(cuda_backend, M, C, with_pos):
    # sub-warp kernel: 8 / 16 / 32 lanes per row x 1..5 chunks (C = 64 .. 1280, ragged row counts); C = 2048 and 72
    # take the one-warp-per-row kernel
    N, F = 3, 4
    x = _rand((M, C), 60, 3.0)
    g, b = _rand((C,), 61, dtype=torch.float32), _rand((C,), 62, dtype=torch.float32)
    pos = _rand((F, C), 63, dtype=torch.float32) if with_pos else None
    o_ref = torch.zeros(M, C, dtype=torch.bfloat16, device=DEV)
    o_cu = torch.zeros_like(o_ref)
    SimBackend().layernorm(x, g, b, pos, o_ref, M, C, 1e-5, N, F)
    cuda_backend.layernorm(x, g, b, pos, o_cu, M, C, 1e-5, N, F)
    torch.cuda.synchronize()
    _report("layernorm", o_cu, o_ref, 4e-3)


@pytest.mark.parametrize("n_inst,rows,C0,C1", [(2, 12 * 1024, 320, 0), (24, 256, 640, 0), (2, 768, 1280, 1280),
                                                (2, 192, 640, 320), (3, 50, 64, 0)])
def test_groupnorm(cuda_backend, n_inst, rows, C0, C1):
    x0 = _rand((n_inst * rows, C0), 70, 2.0) + 0.5
    x1 = _rand((n_inst * rows, C1), 71) if C1 else None
    C = C0 + C1
    g, b = _rand((C,), 72, dtype=torch.float32), _rand((C,), 73, dtype=torch.float32)
    st_ref = torch.zeros(n_inst, C, 2, dtype=torch.float32, device=DEV)
    st_cu = torch.zeros_like(st_ref)
    ws = torch.zeros(max(16, cuda_backend.groupnorm_ws_floats(n_inst, rows, C)), dtype=torch.float32, device=DEV)
    SimBackend().groupnorm_stats(x0, C0, x1, C1, n_inst, rows, 32, 1e-5, g, b, st_ref, ws)
    cuda_backend.groupnorm_stats(x0, C0, x1, C1, n_inst, rows, 32, 1e-5, g, b, st_cu, ws)
    torch.cuda.synchronize()
    _report("gn scale/shift", st_cu.view(-1, 2), st_ref.view(-1, 2), 1e-4)
    # against torch's own group_norm on the concatenated NCHW tensor
    xc = torch.cat([x0.float()] + ([x1.float()] if C1 else []), dim=1).view(n_inst, rows, C).permute(0, 2, 1)
    want = torch.nn.functional.group_norm(xc, 32, g, b, 1e-5).permute(0, 2, 1)
    have = xc.permute(0, 2, 1) * st_cu[:, :, 0].view(n_inst, 1, C) + st_cu[:, :, 1].view(n_inst, 1, C)
    _report("gn vs F.group_norm", have.reshape(-1, C), want.reshape(-1, C), 1e-4)
    # apply: treat every instance as `rows` = h*w pixels of one image; with and without 2x upsample
    h = 1
    for cand in (32, 16, 8, 5, 2, 1):
        if rows % cand == 0:
            h = cand
            break
    w = rows // h
    for up in (0, 1):
        shp = (n_inst * rows * (4 if up else 1), C)
        o_ref = torch.zeros(shp, dtype=torch.bfloat16, device=DEV)
        o_cu = torch.zeros_like(o_ref)
        SimBackend().groupnorm_apply(x0, C0, x1, C1, st_ref, n_inst, n_inst, h, w, 1, up, o_ref)
        cuda_backend.groupnorm_apply(x0, C0, x1, C1, st_ref, n_inst, n_inst, h, w, 1, up, o_cu)
        torch.cuda.synchronize()
        _report(f"gn apply up{up}", o_cu, o_ref, 4e-3)
    o_ref.zero_(), o_cu.zero_()
    SimBackend().groupnorm_apply(x0, C0, x1, C1, None, n_inst, n_inst, h, w, 0, 1, o_ref)
    cuda_backend.groupnorm_apply(x0, C0, x1, C1, None, n_inst, n_inst, h, w, 0, 1, o_cu)
    torch.cuda.synchronize()
    assert torch.equal(o_ref, o_cu)  # pure nearest-upsample copy


@pytest.mark.parametrize("n_inst,rows,C0,C1,silu", [(2, 12 * 1024, 320, 0, 1), (24, 1024, 320, 0, 0),
                                                     (2, 12 * 1024, 640, 320, 1), (24, 256, 640, 0, 0),
                                                     (2, 768, 1280, 1280, 1), (2, 192, 640, 320, 1),
                                                     (2, 192, 1280, 0, 1), (24, 16, 1280, 0, 0), (3, 50, 64, 0, 1),
                                                     (1, 2, 32, 0, 0)])
def test_groupnorm_fused(cuda_backend, n_inst, rows, C0, C1, silu):
    """One-launch GroupNorm (statistics + apply) against torch's group_norm; run three times back to back to prove the
    kernel leaves its barrier workspace clean."""
    x0 = _rand((n_inst * rows, C0), 74, 2.0) + 0.5
    x1 = _rand((n_inst * rows, C1), 75) if C1 else None
    C = C0 + C1
    g, b = _rand((C,), 76, dtype=torch.float32), _rand((C,), 77, dtype=torch.float32)
    xc = torch.cat([x0.float()] + ([x1.float()] if C1 else []), dim=1).view(n_inst, rows, C).permute(0, 2, 1)
    want = torch.nn.functional.group_norm(xc, 32, g, b, 1e-5)
    if silu:
        want = torch.nn.functional.silu(want)
    want = want.permute(0, 2, 1).reshape(-1, C)
    o_sim = torch.zeros(n_inst * rows, C, dtype=torch.bfloat16, device=DEV)
    SimBackend().groupnorm(x0, C0, x1, C1, n_inst, rows, 32, 1e-5, g, b, silu, o_sim)
    outs = []
    for _ in range(3):
        o = torch.zeros_like(o_sim)
        cuda_backend.groupnorm(x0, C0, x1, C1, n_inst, rows, 32, 1e-5, g, b, silu, o)
        outs.append(o)
    torch.cuda.synchronize()
    _report("gn fused vs F.group_norm", outs[0], want, 4e-3)
    _report("gn fused vs sim", outs[0], o_sim, 4e-3)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    for sync in cuda_backend._gn_sync.values():  # counters and accumulators back to zero (one workspace per device)
        assert int(sync[:65544].view(torch.int32).abs().sum()) == 0  # (the two-launch form's partials live past them)
    form = cuda_backend.lib.asva_groupnorm_form(n_inst, rows, C, 32)
    assert form == (2 if (n_inst == 2 and rows == 12 * 1024) else form) and form in (0, 1, 2)


# ------------------------------------------------------------------------------------------------ small kernels
def test_conv_in_and_out(cuda_backend):
    B, Bs, Cl, F, h, w = 2, 1, 4, 5, 8, 12
    lat = _rand((Bs, Cl, F, h, w), 80, dtype=torch.float32)
    o_ref = torch.zeros(B * F * h * w, 64, dtype=torch.bfloat16, device=DEV)
    o_cu = torch.ones_like(o_ref)
    SimBackend().conv_in_im2col(lat, o_ref, B, Bs, Cl, F, h, w)
    cuda_backend.conv_in_im2col(lat, o_cu, B, Bs, Cl, F, h, w)
    torch.cuda.synchronize()
    assert torch.equal(o_ref, o_cu)
    y = _rand((B * F * h * w, 8), 81, dtype=torch.float32)
    wt, bt = _rand((4, 12), 82, dtype=torch.float32), _rand((4,), 83, dtype=torch.float32)
    r_ref = torch.zeros(B, 4, F, h, w, dtype=torch.float32, device=DEV)
    r_cu = torch.zeros_like(r_ref)
    SimBackend().conv_out_finish(y, 8, wt, bt, r_ref, B, 4, F, h, w)
    cuda_backend.conv_out_finish(y, 8, wt, bt, r_cu, B, 4, F, h, w)
    torch.cuda.synchronize()
    _report("conv_out_finish", r_cu.view(-1, w), r_ref.view(-1, w), 1e-5)


@pytest.mark.parametrize("B,F,N,C", [(2, 12, 16, 1280), (1, 5, 7, 64), (2, 1, 4, 320)])
def test_tconv_gather_conv_in_group(cuda_backend, B, F, N, C):
    y = _rand((B * F * N, C), 70)
    got = torch.zeros(B * F * N, 3 * C, dtype=torch.bfloat16, device=DEV)
    ref = torch.zeros_like(got)
    cuda_backend.tconv_gather(y, got, B, F, N, C)
    SimBackend().tconv_gather(y, ref, B, F, N, C)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)  # pure data movement: bit exact


@pytest.mark.parametrize("M,N,K,ai,ao", [(2, 1280, 320, 0, 1), (2, 1280, 1280, 0, 0), (2, 18560, 1280, 1, 0),
                                         (12, 320, 320, 0, 1), (5, 77, 64, 1, 1)])
def test_small_linear(cuda_backend, M, N, K, ai, ao):
    x = _rand((M, K), 90, dtype=torch.float32)
    w = _rand((N, K), 91, 1.0 / math.sqrt(K))
    b = _rand((N,), 92, dtype=torch.float32)
    o_ref = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    o_cu = torch.zeros_like(o_ref)
    SimBackend().small_linear(x, w, b, o_ref, M, N, K, ai, ao)
    cuda_backend.small_linear(x, w, b, o_cu, M, N, K, ai, ao)
    torch.cuda.synchronize()
    _report("small_linear", o_cu, o_ref, 1e-5)


def test_timestep_features(cuda_backend):
    t = torch.tensor([981.0, 1.0, 500.0], device=DEV)
    o_ref = torch.zeros(3, 320, device=DEV)
    o_cu = torch.zeros_like(o_ref)
    SimBackend().timestep_features(t, o_ref, 3, 320, True)
    cuda_backend.timestep_features(t, o_cu, 3, 320, True)
    torch.cuda.synchronize()
    assert (o_cu - o_ref).abs().max() < 2e-4  # sin/cos of arguments up to ~1e3 in fp32


@pytest.mark.parametrize("k,clips", [(1, 1), (2, 1), (3, 1), (2, 3), (3, 2)])
def test_cfg_steps(cuda_backend, k, clips):
    """eps (k, clips, C, F, hw) branch-major; latents (clips, C, F, hw); the clips of a launch share the step."""
    C, F, hw = 4, 6, 80
    eps = _rand((k * clips, C, F, hw), 100, dtype=torch.float32)
    lat0 = _rand((clips, C, F, hw), 101, dtype=torch.float32)
    coef = torch.tensor([-3.0, 4.0, 0.5, 1.01, -0.07, 23 / 12, -16 / 12, 5 / 12, 0.0], device=DEV)
    if k < 3:
        coef[2] = 0.0
    a, b = lat0.clone(), lat0.clone()
    SimBackend().cfg_ddim_step(eps, k, a, coef, C, F, hw, clips)
    cuda_backend.cfg_ddim_step(eps, k, b, coef, C, F, hw, clips)
    torch.cuda.synchronize()
    assert torch.equal(a[:, :, 0], lat0[:, :, 0]) and torch.equal(b[:, :, 0], lat0[:, :, 0])
    _report("cfg ddim", b.view(-1, hw), a.view(-1, hw), 1e-6)
    if clips > 1:  # clip j of a batched launch == the same clip stepped alone
        j = clips - 1
        alone = lat0[j].clone()
        cuda_backend.cfg_ddim_step(eps.view(k, clips, C, F, hw)[:, j].contiguous(), k, alone, coef, C, F, hw, 1)
        torch.cuda.synchronize()
        assert torch.equal(alone, b[j])
    hist0 = _rand((4, clips, C, F, hw), 102, dtype=torch.float32)
    slots = torch.tensor([2, 0, 1, 3], dtype=torch.int32, device=DEV)
    a, b, ha, hb = lat0.clone(), lat0.clone(), hist0.clone(), hist0.clone()
    SimBackend().cfg_plms_step(eps, k, a, ha, coef, slots, C, F, hw, clips)
    cuda_backend.cfg_plms_step(eps, k, b, hb, coef, slots, C, F, hw, clips)
    torch.cuda.synchronize()
    _report("cfg plms", b.view(-1, hw), a.view(-1, hw), 1e-6)
    _report("cfg plms hist", hb[:, :, :, 1:].reshape(-1, hw), ha[:, :, :, 1:].reshape(-1, hw), 1e-6)
