// LayerNorm and GroupNorm kernels (HBM/L2-bound elementwise + reductions; CUDA cores, 16-byte vector access).
#include "common.cuh"
#include "host_common.h"

namespace asva {

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per token row; row kept in registers (C <= 2048), two-pass mean / variance.
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxChunks = 8;  // 8 chunks x 8 elements x 32 lanes = 2048 channels

template <int kChunks>  // ceil(C / 256): keeps the register footprint (and so the occupancy) proportional to C
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ pos,
                                                        __nv_bfloat16* __restrict__ out, int64_t M, int C, float eps,
                                                        int N, int F) {
  pdl_trigger();
  pdl_wait();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int nchunk = C >> 3;
  const __nv_bfloat16* xr = x + row * C;
  const float* pr = (pos != nullptr) ? pos + static_cast<int64_t>((row / N) % F) * C : nullptr;
  float v[kChunks][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nchunk) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + ch * 8);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
      if (pr != nullptr) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pr + ch * 8));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pr + ch * 8) + 1);
        v[i][0] += p0.x; v[i][1] += p0.y; v[i][2] += p0.z; v[i][3] += p0.w;
        v[i][4] += p1.x; v[i][5] += p1.y; v[i][6] += p1.z; v[i][7] += p1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v[i][j] - mean;
        sq += dlt * dlt;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
  __nv_bfloat16* orow = out + row * C;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nchunk) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]);
      u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(orow + ch * 8) = u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics, two deterministic stages.
//   stage 1: grid (splits, n_inst, cblocks); a CTA reduces a row range x a block of 8-channel chunks into
//            per-channel (sum, sumsq) partials  ws[inst][split][channel][2]
//   stage 2: one thread per (inst, group) folds splits x channels-in-group in double precision
// ------------------------------------------------------------------------------------------------
struct GnStatsPlan {
  int splits, cblocks, cw, rows_per_pass;
};

static GnStatsPlan gn_plan(int n_inst, int64_t rows, int Ctot) {
  GnStatsPlan pl;
  const int nchunk = Ctot / 8;
  pl.cblocks = (nchunk + 127) / 128;
  pl.cw = (nchunk + pl.cblocks - 1) / pl.cblocks;
  pl.rows_per_pass = 256 / pl.cw;
  if (pl.rows_per_pass < 1) pl.rows_per_pass = 1;
  int64_t want = (2 * 148 + (int64_t)n_inst * pl.cblocks - 1) / ((int64_t)n_inst * pl.cblocks);
  int64_t max_splits = rows / (4 * (int64_t)pl.rows_per_pass);
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  pl.splits = (int)want;
  return pl;
}

__global__ void __launch_bounds__(256) gn_stats_stage1(const __nv_bfloat16* __restrict__ x0, int C0,
                                                       const __nv_bfloat16* __restrict__ x1, int C1, int64_t rows,
                                                       int splits, int cw, int rows_per_pass,
                                                       float* __restrict__ ws) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];  // [rows_per_pass][cw*8][2]
  const int Ctot = C0 + C1;
  const int nchunk = Ctot >> 3;
  const int split = blockIdx.x, inst = blockIdx.y, cb = blockIdx.z;
  const int rl = threadIdx.x / cw;
  const int cl = threadIdx.x % cw;
  const int ch = cb * cw + cl;
  const bool active = (rl < rows_per_pass) && (ch < nchunk);
  const int64_t rbeg = rows * split / splits, rend = rows * (split + 1) / splits;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    const int c = ch * 8;
    const __nv_bfloat16* src;
    int ld, cc;
    if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
    src += (static_cast<int64_t>(inst) * rows) * ld + cc;
    for (int64_t r = rbeg + rl; r < rend; r += rows_per_pass) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + r * ld);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cz = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      const float f[8] = {a.x, a.y, b.x, b.y, cz.x, cz.y, d.x, d.y};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] += f[j] * f[j];
      }
    }
  }
  if (rl < rows_per_pass && cl < cw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[((rl * cw + cl) * 8 + j) * 2 + 0] = s[j];
      red[((rl * cw + cl) * 8 + j) * 2 + 1] = q[j];
    }
  }
  __syncthreads();
  // fold the row lanes: thread t < cw*8 owns one channel of this block
  for (int t = threadIdx.x; t < cw * 8; t += blockDim.x) {
    const int chn = cb * cw * 8 + t;
    if (chn >= Ctot) continue;
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) {
      ss += red[((r * cw) * 8 + t) * 2 + 0];
      qq += red[((r * cw) * 8 + t) * 2 + 1];
    }
    float* w = ws + ((static_cast<int64_t>(inst) * splits + split) * Ctot + chn) * 2;
    w[0] = ss;
    w[1] = qq;
  }
}

__global__ void __launch_bounds__(128) gn_stats_stage2(const float* __restrict__ ws, int n_inst, int splits, int Ctot,
                                                         int groups, int64_t rows, float eps,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  // one CTA per (instance, group): threads stride over splits x channels-in-group, fp64 reduction, then the
  // per-channel affine  y = x * scale + shift  (scale = rstd * gamma, shift = beta - mean * scale) is written
  __shared__ double red[2][4];
  __shared__ float mr[2];
  const int idx = blockIdx.x;
  const int inst = idx / groups, g = idx % groups;
  const int cpg = Ctot / groups;
  const int total = splits * cpg;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < total; i += 128) {
    const int sp = i / cpg, c = i - sp * cpg;
    const float2 v = *reinterpret_cast<const float2*>(
        ws + ((static_cast<int64_t>(inst) * splits + sp) * Ctot + g * cpg + c) * 2);
    s += static_cast<double>(v.x);
    q += static_cast<double>(v.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    q = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const double cnt = static_cast<double>(rows) * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    mr[0] = static_cast<float>(mean);
    mr[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const float mean = mr[0], rstd = mr[1];
  for (int c = threadIdx.x; c < cpg; c += 128) {
    const int ch = g * cpg + c;
    const float sc = rstd * gamma[ch];
    float* o = stats + (static_cast<int64_t>(inst) * Ctot + ch) * 2;
    o[0] = sc;
    o[1] = beta[ch] - mean * sc;
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU, + nearest 2x upsample, + concat of two sources), elementwise on 8-channel chunks.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                       const __nv_bfloat16* __restrict__ x1, int C1,
                                                       const float* __restrict__ stats, int img_per_inst, int h,
                                                       int w, int silu, int up, __nv_bfloat16* __restrict__ out,
                                                       int64_t total_chunks) {
  pdl_trigger();
  pdl_wait();
  const int Ctot = C0 + C1;
  const int nchunk = Ctot >> 3;
  const int ho = up ? 2 * h : h, wo = up ? 2 * w : w;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_chunks;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % nchunk);
    const int64_t pix = i / nchunk;
    int64_t srow = pix, img = pix / (static_cast<int64_t>(wo) * ho);
    if (up) {
      const int xo = static_cast<int>(pix % wo);
      const int yo = static_cast<int>((pix / wo) % ho);
      srow = (img * h + (yo >> 1)) * w + (xo >> 1);
    }
    const int c = ch * 8;
    const __nv_bfloat16* src = (c < C0) ? x0 + srow * C0 + c : x1 + srow * C1 + (c - C0);
    const uint4 u = *reinterpret_cast<const uint4*>(src);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cz = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    float f[8] = {a.x, a.y, b.x, b.y, cz.x, cz.y, d.x, d.y};
    if (stats != nullptr) {
      const float4* st = reinterpret_cast<const float4*>(stats + (static_cast<int64_t>(img / img_per_inst) * Ctot + c) * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(st + j);  // (scale, shift) of channels c+2j, c+2j+1
        f[2 * j] = fmaf(f[2 * j], t.x, t.y);
        f[2 * j + 1] = fmaf(f[2 * j + 1], t.z, t.w);
      }
    }
    if (silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + pix * Ctot + c) = o;
  }
}

}  // namespace asva

extern "C" int asva_layernorm(const void* x, const float* gamma, const float* beta, const float* pos, void* out,
                              int64_t M, int32_t C, float eps, int32_t N, int32_t F, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(x && gamma && beta && out, "asva_layernorm: null operand");
  ASVA_REQUIRE(C % 8 == 0 && C >= 8 && C <= 8 * 32 * kLnMaxChunks, "asva_layernorm: C=%d unsupported", C);
  ASVA_REQUIRE(M >= 1, "asva_layernorm: M must be positive");
  ASVA_REQUIRE(pos == nullptr || (N >= 1 && F >= 1), "asva_layernorm: pos needs N, F");
  const unsigned blocks = static_cast<unsigned>((M + 7) / 8);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  const int n = N > 0 ? N : 1, f = F > 0 ? F : 1;
#define ASVA_LN_CASE(K) \
  case K: ASVA_CUDA_OK(launch_k(layernorm_kernel<K>, dim3(blocks), dim3(256), 0, stream, 1, xp, gamma, beta, pos, op, M, C, eps, n, f)); break;
  switch ((C / 8 + 31) / 32) {
    ASVA_LN_CASE(1) ASVA_LN_CASE(2) ASVA_LN_CASE(3) ASVA_LN_CASE(4) ASVA_LN_CASE(5) ASVA_LN_CASE(6) ASVA_LN_CASE(7)
    default: ASVA_CUDA_OK(launch_k(layernorm_kernel<8>, dim3(blocks), dim3(256), 0, stream, 1, xp, gamma, beta, pos, op, M, C, eps, n, f)); break;
  }
#undef ASVA_LN_CASE
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int64_t asva_groupnorm_ws_floats(int32_t n_inst, int64_t rows, int32_t C) {
  if (n_inst < 1 || rows < 1 || C < 8) return 0;
  asva::GnStatsPlan pl = asva::gn_plan(n_inst, rows, C);
  return static_cast<int64_t>(n_inst) * pl.splits * C * 2;
}

extern "C" int asva_groupnorm_stats(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t n_inst,
                                    int64_t rows, int32_t groups, float eps, const float* gamma, const float* beta,
                                    float* stats, float* partial_ws, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (x1 == nullptr) C1 = 0;
  const int Ctot = C0 + C1;
  ASVA_REQUIRE(x0 && stats && partial_ws && gamma && beta, "asva_groupnorm_stats: null operand");
  ASVA_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && Ctot >= 8, "asva_groupnorm_stats: channels must be multiples of 8");
  ASVA_REQUIRE(groups >= 1 && Ctot % groups == 0, "asva_groupnorm_stats: C=%d not divisible by groups=%d", Ctot, groups);
  ASVA_REQUIRE(n_inst >= 1 && rows >= 1, "asva_groupnorm_stats: empty problem");
  GnStatsPlan pl = gn_plan(n_inst, rows, Ctot);
  dim3 grid(pl.splits, n_inst, pl.cblocks);
  const size_t smem = static_cast<size_t>(pl.rows_per_pass) * pl.cw * 8 * 2 * sizeof(float);
  ASVA_CUDA_OK(launch_k(gn_stats_stage1, dim3(grid), dim3(256), smem, stream, 1, reinterpret_cast<const __nv_bfloat16*>(x0), C0,
                                               reinterpret_cast<const __nv_bfloat16*>(x1), C1, rows, pl.splits,
                                               pl.cw, pl.rows_per_pass, partial_ws));
  ASVA_CUDA_OK(cudaGetLastError());
  ASVA_CUDA_OK(launch_k(gn_stats_stage2, dim3(n_inst * groups), dim3(128), 0, stream, 1, partial_ws, n_inst, pl.splits, Ctot, groups, rows, eps, gamma,
                                                        beta, stats));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_groupnorm_apply(const void* x0, int32_t C0, const void* x1, int32_t C1, const float* stats,
                                    int32_t n_inst, int32_t n_img, int32_t h, int32_t w, int32_t silu,
                                    int32_t upsample, void* out, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (x1 == nullptr) C1 = 0;
  const int Ctot = C0 + C1;
  ASVA_REQUIRE(x0 && out, "asva_groupnorm_apply: null operand");
  ASVA_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && Ctot >= 8, "asva_groupnorm_apply: channels must be multiples of 8");
  ASVA_REQUIRE(stats == nullptr || (n_inst >= 1 && n_img % n_inst == 0),
               "asva_groupnorm_apply: inconsistent normalisation arguments");
  const int ho = upsample ? 2 * h : h, wo = upsample ? 2 * w : w;
  const int64_t total = static_cast<int64_t>(n_img) * ho * wo * (Ctot / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ASVA_CUDA_OK(launch_k(gn_apply_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, 1, reinterpret_cast<const __nv_bfloat16*>(x0), C0, reinterpret_cast<const __nv_bfloat16*>(x1), C1, stats,
      n_inst > 0 ? n_img / n_inst : 1, h, w, silu, upsample, reinterpret_cast<__nv_bfloat16*>(out), total));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}
