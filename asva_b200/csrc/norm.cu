// LayerNorm and GroupNorm kernels (HBM/L2-bound elementwise + reductions; CUDA cores, 16-byte vector access).
#include <stdlib.h>

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

// LayerNorm, sub-warp form: a row is held by 8, 16 or 32 lanes with NCH 16-byte chunks each (C = 8 * lanes * NCH - the
// SD-1.5 widths 320 / 640 / 1280 are 5 chunks on 8 / 16 / 32 lanes), so a warp normalises 4 / 2 / 1 rows at a time
// with every lane busy (the one-warp-per-row form above leaves 3/4 of the lanes idle on the second chunk of a
// 320-wide row and gives each warp one row: ~12 us for the 24576 x 320 rows of a level-0 block, 2.6 TB/s).  gamma /
// beta sit in shared memory (loaded before the dependent-launch wait), the warps loop over row groups with the next
// group's loads issued before the current group's arithmetic, and the grid is sized for equal trip counts.
template <int NCH>
__global__ void __launch_bounds__(256, 2) layernorm_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                const float* __restrict__ pos,
                                                                __nv_bfloat16* __restrict__ out, int64_t M, int C,
                                                                float eps, int N, int F, int lanes_log2) {
  extern __shared__ float gb[];  // gamma[C] | beta[C]
  pdl_trigger();
  for (int i = threadIdx.x; i < (C >> 2); i += blockDim.x) {
    reinterpret_cast<float4*>(gb)[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    reinterpret_cast<float4*>(gb + C)[i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  __syncthreads();
  pdl_wait();
  const int lanes = 1 << lanes_log2, rows_per_warp = 32 >> lanes_log2;
  const int lane = threadIdx.x & 31;
  const int sub = lane >> lanes_log2, li = lane & (lanes - 1);
  const int wpb = blockDim.x >> 5;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * wpb;
  const float inv_c = 1.0f / static_cast<float>(C);
  int64_t grp = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  uint4 nxt[NCH];
  auto fetch = [&](int64_t g) {
    const int64_t row = g * rows_per_warp + sub;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      nxt[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row < M) nxt[i] = *reinterpret_cast<const uint4*>(x + row * C + (li + lanes * i) * 8);
    }
  };
  if (grp * rows_per_warp < M) fetch(grp);
  for (; grp * rows_per_warp < M; grp += n_warps) {  // warp-uniform trip count
    const int64_t row = grp * rows_per_warp + sub;
    float v[NCH][8];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float2 a = unpack_bf16x2(nxt[i].x), b = unpack_bf16x2(nxt[i].y), c = unpack_bf16x2(nxt[i].z),
                   d = unpack_bf16x2(nxt[i].w);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
      v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
    }
    if ((grp + n_warps) * rows_per_warp < M) fetch(grp + n_warps);
    if (pos != nullptr && row < M) {
      const float* pr = pos + static_cast<int64_t>((row / N) % F) * C;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pr + (li + lanes * i) * 8));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pr + (li + lanes * i) * 8) + 1);
        v[i][0] += p0.x; v[i][1] += p0.y; v[i][2] += p0.z; v[i][3] += p0.w;
        v[i][4] += p1.x; v[i][5] += p1.y; v[i][6] += p1.z; v[i][7] += p1.w;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    for (int o = lanes >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v[i][j] - mean;
        sq = fmaf(dlt, dlt, sq);
      }
    for (int o = lanes >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    if (row < M) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c0 = (li + lanes * i) * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(gb + c0), g1 = *reinterpret_cast<const float4*>(gb + c0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(gb + C + c0), b1 = *reinterpret_cast<const float4*>(gb + C + c0 + 4);
        uint4 u;
        u.x = pack_bf16x2((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y);
        u.y = pack_bf16x2((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w);
        u.z = pack_bf16x2((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y);
        u.w = pack_bf16x2((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w);
        *reinterpret_cast<uint4*>(out + row * C + c0) = u;
      }
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


// ------------------------------------------------------------------------------------------------
// Fused GroupNorm: statistics + apply (+SiLU) in ONE launch.  The grid is sized to be co-resident (<= kGnCtasPerSm
// CTAs per SM), CTA (inst, split, cb) reduces its rows x channel block, folds the sums per group into fp64
// accumulators in global memory (atomicAdd), the whole grid meets at a counter barrier, and every CTA then normalises
// exactly the rows it reduced (second read served by L1 / L2).  Three launches and a partials round trip become one
// launch whose HBM traffic is one read and one write of the tensor.  The accumulators and both counters are left
// zeroed by the last CTA that passes the exit counter, so the workspace only has to be zeroed once, at allocation.
// ------------------------------------------------------------------------------------------------
constexpr int kGnCtasPerSm = 4;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void gn_accum8(const uint4& u, float (&s)[8], float (&q)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  const float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] += f[j];
    q[j] = fmaf(f[j], f[j], q[j]);
  }
}

__device__ __forceinline__ uint4 gn_affine8(const uint4& u, const float4 (&t)[4], int silu) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = fmaf(f[2 * j], t[j].x, t[j].y);
    f[2 * j + 1] = fmaf(f[2 * j + 1], t[j].z, t[j].w);
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
  return o;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm in two launches for the few-instances x many-rows norms (a whole clip per instance at the top resolution:
// 2 x 12288 rows x 320..960 channels).  The one-launch forms pay for their grid-wide meeting: the barrier kernel below
// runs those shapes at 1.1-1.8 TB/s (29 us for 31 MB of traffic).  Here the kernel boundary is the barrier:
//   A  gn2_stats_kernel: CTA (split, inst) reduces full rows of its row range (whole 128-byte lines, no strips) to
//      per-GROUP partials ws[inst][split][group][2] - a few KB per instance instead of per-channel tables;
//   B  gn2_apply_kernel: every CTA folds its instance's splits x groups partials (fp64, fixed order) into per-channel
//      (scale, shift) in shared memory, then streams its rows: one read (L2-warm from A), one write.
// With programmatic dependent launch B's prologue overlaps A's tail.  Deterministic: no atomics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) gn2_stats_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                        const __nv_bfloat16* __restrict__ x1, int C1, int64_t rows,
                                                        int splits, int cw, int rows_per_pass, int groups,
                                                        int cs, float* __restrict__ ws) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];  // [rows_per_pass][cw*8][2], then [cw*8][2] per-channel sums
  const int Ctot = C0 + C1;
  const int split = blockIdx.x, inst = blockIdx.y;
  const int rl = threadIdx.x / cw;
  const int cl = threadIdx.x % cw;
  const bool active = rl < rows_per_pass;
  const int64_t rbeg = rows * split / splits, rend = rows * (split + 1) / splits;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    const int c = cl * 8;
    const __nv_bfloat16* src;
    int ld, cc;
    if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
    src += (static_cast<int64_t>(inst) * rows) * ld + cc;
    int64_t r = rbeg + rl;
    const int64_t st = static_cast<int64_t>(rows_per_pass) * ld;
    for (; r + 3 * rows_per_pass < rend; r += 4 * rows_per_pass) {  // four rows in flight per thread
      const __nv_bfloat16* a = src + r * ld;
      const uint4 u0 = *reinterpret_cast<const uint4*>(a);
      const uint4 u1 = *reinterpret_cast<const uint4*>(a + st);
      const uint4 u2 = *reinterpret_cast<const uint4*>(a + 2 * st);
      const uint4 u3 = *reinterpret_cast<const uint4*>(a + 3 * st);
      gn_accum8(u0, s, q);
      gn_accum8(u1, s, q);
      gn_accum8(u2, s, q);
      gn_accum8(u3, s, q);
    }
    for (; r < rend; r += rows_per_pass) gn_accum8(*reinterpret_cast<const uint4*>(src + r * ld), s, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[((rl * cw + cl) * 8 + j) * 2 + 0] = s[j];
      red[((rl * cw + cl) * 8 + j) * 2 + 1] = q[j];
    }
  }
  __syncthreads();
  float* chan = red + rows_per_pass * cw * 16;
  for (int t = threadIdx.x; t < Ctot; t += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) {
      ss += red[((r * cw) * 8 + t) * 2 + 0];
      qq += red[((r * cw) * 8 + t) * 2 + 1];
    }
    chan[2 * t] = ss;
    chan[2 * t + 1] = qq;
  }
  __syncthreads();
  // per-group sums of this CTA, then of its cluster of 8 (distributed shared memory): the apply kernel then folds
  // splits / 8 partials per group instead of splits (its prologue is paid by every one of its CTAs)
  const int cpg = Ctot / groups;
  float* gsum = chan + 2 * Ctot;  // [groups][2]
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int c = 0; c < cpg; ++c) {
      ss += chan[2 * (g * cpg + c)];
      qq += chan[2 * (g * cpg + c) + 1];
    }
    gsum[2 * g] = ss;
    gsum[2 * g + 1] = qq;
  }
  if (cs == 1) {  // no cluster: this CTA's sums are a partial of their own
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x)
      *reinterpret_cast<float2*>(ws + ((static_cast<int64_t>(inst) * splits + split) * groups + g) * 2) =
          make_float2(gsum[2 * g], gsum[2 * g + 1]);
    return;
  }
  cluster_sync_all();
  if (cluster_ctarank() == 0) {
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      float ss = 0.f, qq = 0.f;
      const uint32_t la = smem_u32(gsum + 2 * g);
      for (uint32_t rk = 0; rk < static_cast<uint32_t>(cs); ++rk) {
        uint32_t ra;
        float a, b;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rk));
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(ra) : "memory");
        ss += a;
        qq += b;
      }
      *reinterpret_cast<float2*>(ws + ((static_cast<int64_t>(inst) * (splits / cs) + (split / cs)) * groups + g) * 2) =
          make_float2(ss, qq);
    }
  }
  cluster_sync_all();  // the peers' shared memory stays alive until rank 0 has read it
}

__global__ void __launch_bounds__(256, 4) gn2_apply_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                        const __nv_bfloat16* __restrict__ x1, int C1, int64_t rows,
                                                        int splits, int groups, float eps,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int silu,
                                                        const float* __restrict__ ws, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float4 tab[];  // [Ctot / 2]: (scale, shift) of two adjacent channels; then float mr[groups][2]
  pdl_trigger();
  const int Ctot = C0 + C1;
  const int nchunk = Ctot >> 3;
  const int cpg = Ctot / groups;
  const int inst = blockIdx.y;
  float* mr = reinterpret_cast<float*>(tab + Ctot / 2);
  pdl_wait();
  {
    // eight lanes per group fold its partials (independent loads: one round trip for all groups), fp64, fixed order
    const int part = threadIdx.x & 7;
    const double cnt = static_cast<double>(rows) * cpg;
    for (int g0 = 0; g0 < groups; g0 += 32) {
      const int g = g0 + (threadIdx.x >> 3);
      double sm = 0.0, sq = 0.0;
      if (g < groups) {
        for (int sp = part; sp < splits; sp += 8) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(ws + ((static_cast<int64_t>(inst) * splits + sp) * groups + g) * 2));
          sm += static_cast<double>(v.x);
          sq += static_cast<double>(v.y);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (part == 0 && g < groups) {
        const double mean = sm / cnt;
        double var = sq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        mr[2 * g] = static_cast<float>(mean);
        mr[2 * g + 1] = rsqrtf(static_cast<float>(var) + eps);
      }
    }
  }
  __syncthreads();
  for (int c2 = threadIdx.x; c2 < Ctot / 2; c2 += blockDim.x) {
    const int c = 2 * c2;
    const int ga = c / cpg, gb = (c + 1) / cpg;
    const float sa = mr[2 * ga + 1] * __ldg(gamma + c), sb = mr[2 * gb + 1] * __ldg(gamma + c + 1);
    tab[c2] = make_float4(sa, __ldg(beta + c) - mr[2 * ga] * sa, sb, __ldg(beta + c + 1) - mr[2 * gb] * sb);
  }
  __syncthreads();
  // CTA x of the instance owns a row range; thread (rl, cl) walks it rpp rows at a time on ONE 8-channel chunk, whose
  // four (scale, shift) pairs stay in registers - no index arithmetic and no table reads inside the loop
  const int cw = nchunk, rpp = blockDim.x / cw;
  const int rl = threadIdx.x / cw, cl = threadIdx.x - rl * cw;
  if (rl >= rpp) return;
  const float4 t[4] = {tab[cl * 4], tab[cl * 4 + 1], tab[cl * 4 + 2], tab[cl * 4 + 3]};
  const int64_t rbeg = rows * blockIdx.x / gridDim.x, rend = rows * (blockIdx.x + 1) / gridDim.x;
  const int c = cl * 8;
  const __nv_bfloat16* src;
  int ld;
  if (c < C0) { src = x0 + c; ld = C0; } else { src = x1 + (c - C0); ld = C1; }
  src += static_cast<int64_t>(inst) * rows * ld;
  __nv_bfloat16* dst = out + static_cast<int64_t>(inst) * rows * Ctot + c;
  int64_t r = rbeg + rl;
  const int64_t si = static_cast<int64_t>(rpp) * ld, so = static_cast<int64_t>(rpp) * Ctot;
  for (; r + 3 * rpp < rend; r += 4 * rpp) {  // four rows in flight per thread
    const __nv_bfloat16* a = src + r * ld;
    __nv_bfloat16* o = dst + r * Ctot;
    const uint4 u0 = *reinterpret_cast<const uint4*>(a);
    const uint4 u1 = *reinterpret_cast<const uint4*>(a + si);
    const uint4 u2 = *reinterpret_cast<const uint4*>(a + 2 * si);
    const uint4 u3 = *reinterpret_cast<const uint4*>(a + 3 * si);
    *reinterpret_cast<uint4*>(o) = gn_affine8(u0, t, silu);
    *reinterpret_cast<uint4*>(o + so) = gn_affine8(u1, t, silu);
    *reinterpret_cast<uint4*>(o + 2 * so) = gn_affine8(u2, t, silu);
    *reinterpret_cast<uint4*>(o + 3 * so) = gn_affine8(u3, t, silu);
  }
  for (; r < rend; r += rpp)
    *reinterpret_cast<uint4*>(dst + r * Ctot) = gn_affine8(*reinterpret_cast<const uint4*>(src + r * ld), t, silu);
}


__global__ void __launch_bounds__(256) gn_fused_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                       const __nv_bfloat16* __restrict__ x1, int C1, int64_t rows,
                                                       int splits, int cblocks, int cw, int rows_per_pass, int groups,
                                                       float eps, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int silu,
                                                       __nv_bfloat16* __restrict__ out, double* __restrict__ acc,
                                                       unsigned* __restrict__ ctr, int n_acc) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256 * 16];  // [rows_per_pass][cw*8][2]
  __shared__ __align__(16) float chs[1024 * 2];  // per-channel (sum, sumsq) of this CTA, later the (scale, shift) table
  __shared__ int s_last;
  const int Ctot = C0 + C1;
  const int nchunk = Ctot >> 3;
  const int cpg = Ctot / groups;
  const int cb = blockIdx.x % cblocks;
  const int split = (blockIdx.x / cblocks) % splits;
  const int inst = blockIdx.x / (cblocks * splits);
  const int rl = threadIdx.x / cw;
  const int cl = threadIdx.x % cw;
  const int ch = cb * cw + cl;
  const bool active = (rl < rows_per_pass) && (ch < nchunk);
  const int64_t rbeg = rows * split / splits, rend = rows * (split + 1) / splits;
  const int c = ch * 8;
  const __nv_bfloat16* src = x0;
  int64_t ld = C0;
  if (active) {
    if (c < C0) { src = x0 + c; } else { src = x1 + (c - C0); ld = C1; }
    src += static_cast<int64_t>(inst) * rows * ld;
  }
  const int64_t rstep = rows_per_pass;
  // ---------------- phase 1: sums over this CTA's rows ----------------
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    if (active) {
      int64_t r = rbeg + rl;
      for (; r + 3 * rstep < rend; r += 4 * rstep) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(src + r * ld);
        const uint4 u1 = *reinterpret_cast<const uint4*>(src + (r + rstep) * ld);
        const uint4 u2 = *reinterpret_cast<const uint4*>(src + (r + 2 * rstep) * ld);
        const uint4 u3 = *reinterpret_cast<const uint4*>(src + (r + 3 * rstep) * ld);
        gn_accum8(u0, s, q);
        gn_accum8(u1, s, q);
        gn_accum8(u2, s, q);
        gn_accum8(u3, s, q);
      }
      for (; r < rend; r += rstep) gn_accum8(*reinterpret_cast<const uint4*>(src + r * ld), s, q);
    }
    if (rl < rows_per_pass) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red[((rl * cw + cl) * 8 + j) * 2 + 0] = s[j];
        red[((rl * cw + cl) * 8 + j) * 2 + 1] = q[j];
      }
    }
  }
  __syncthreads();
  const int c_lo = cb * cw * 8;
  const int c_hi = min(c_lo + cw * 8, Ctot);
  for (int t = threadIdx.x; t < cw * 8; t += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) {
      ss += red[((r * cw) * 8 + t) * 2 + 0];
      qq += red[((r * cw) * 8 + t) * 2 + 1];
    }
    chs[2 * t] = ss;
    chs[2 * t + 1] = qq;
  }
  __syncthreads();
  {
    const int g_lo = c_lo / cpg, g_hi = (c_hi - 1) / cpg;
    for (int g = g_lo + threadIdx.x; g <= g_hi; g += blockDim.x) {
      const int a = max(g * cpg, c_lo), b = min((g + 1) * cpg, c_hi);
      double ss = 0.0, qq = 0.0;
      for (int k = a; k < b; ++k) {
        ss += static_cast<double>(chs[2 * (k - c_lo)]);
        qq += static_cast<double>(chs[2 * (k - c_lo) + 1]);
      }
      double* dst = acc + (static_cast<int64_t>(inst) * groups + g) * 2;
      atomicAdd(dst, ss);
      atomicAdd(dst + 1, qq);
    }
  }
  // ---------------- grid barrier (all CTAs are resident by construction) ----------------
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < gridDim.x) {
      if (clock64() - t0 > (1ll << 32)) __trap();  // ~2 s: never spin forever on a misconfigured launch
    }
    __threadfence();
  }
  __syncthreads();
  // ---------------- phase 2: (scale, shift) of this CTA's channels ----------------
  for (int t = threadIdx.x; t < cw * 8; t += blockDim.x) {
    const int chn = c_lo + t;
    if (chn >= Ctot) continue;
    const int g = chn / cpg;
    const double* a = acc + (static_cast<int64_t>(inst) * groups + g) * 2;
    const double cnt = static_cast<double>(rows) * cpg;
    const double mean = __ldcg(a) / cnt;
    double var = __ldcg(a + 1) / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = rstd * gamma[chn];
    chs[2 * t] = sc;
    chs[2 * t + 1] = beta[chn] - static_cast<float>(mean) * sc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(ctr + 1, 1u);
    s_last = (old == gridDim.x - 1) ? 1 : 0;
  }
  // ---------------- phase 3: normalise the same rows ----------------
  if (active) {
    float4 t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float4*>(&chs[(cl * 8 + 2 * j) * 2]);
    __nv_bfloat16* dst = out + static_cast<int64_t>(inst) * rows * Ctot + c;
    int64_t r = rbeg + rl;
    for (; r + 3 * rstep < rend; r += 4 * rstep) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(src + r * ld);
      const uint4 u1 = *reinterpret_cast<const uint4*>(src + (r + rstep) * ld);
      const uint4 u2 = *reinterpret_cast<const uint4*>(src + (r + 2 * rstep) * ld);
      const uint4 u3 = *reinterpret_cast<const uint4*>(src + (r + 3 * rstep) * ld);
      *reinterpret_cast<uint4*>(dst + r * Ctot) = gn_affine8(u0, t, silu);
      *reinterpret_cast<uint4*>(dst + (r + rstep) * Ctot) = gn_affine8(u1, t, silu);
      *reinterpret_cast<uint4*>(dst + (r + 2 * rstep) * Ctot) = gn_affine8(u2, t, silu);
      *reinterpret_cast<uint4*>(dst + (r + 3 * rstep) * Ctot) = gn_affine8(u3, t, silu);
    }
    for (; r < rend; r += rstep)
      *reinterpret_cast<uint4*>(dst + r * Ctot) = gn_affine8(*reinterpret_cast<const uint4*>(src + r * ld), t, silu);
  }
  __syncthreads();
  if (s_last) {  // every CTA has read the accumulators: leave the workspace zeroed for the next launch
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x == 0) {
      ctr[0] = 0u;
      ctr[1] = 0u;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Fused GroupNorm, cluster form (the fast path): one thread-block cluster owns one (instance, bundle of gb groups)
// unit - a column strip of gb * C/groups channels (a multiple of 8, so 16-byte chunks never leave the strip) - and its
// CTAs split the instance's rows.  Each CTA reduces its rows (fp32 per thread, warp shuffles, fp64 across warps),
// the cluster exchanges the per-group sums through distributed shared memory (one barrier.cluster, no global
// atomics, no grid barrier), and every CTA normalises its own rows: from REGISTERS when its slice is at most KMAX
// chunks per thread (the tensor is read from HBM once), else with a second read that L2 serves.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_dsmem_f64(const double* p, uint32_t rank) {
  uint32_t a = smem_u32(p), ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

constexpr int kGnMaxGb = 8;  // groups per bundle (8 / gcd(C/groups, 8) <= 8)

template <int KMAX>
__global__ void __launch_bounds__(512) gn_cluster_kernel(const __nv_bfloat16* __restrict__ x0, int C0,
                                                         const __nv_bfloat16* __restrict__ x1, int C1, int64_t rows,
                                                         int groups, int gb, int cs, float eps,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int silu,
                                                         __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  __shared__ double wpart[16][kGnMaxGb][2];  // per-warp group sums
  __shared__ double part[kGnMaxGb][2];       // this CTA's group sums (read by the cluster)
  __shared__ float stat[kGnMaxGb][2];        // (mean, rstd) of the bundle's groups
  const int Ctot = C0 + C1;
  const int cpg = Ctot / groups;
  const int bch = gb * cpg;  // channels of a bundle
  const int bw = bch >> 3;   // ... in 16-byte chunks
  const int nb = groups / gb;
  const int rank = (cs > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit = blockIdx.x / cs;
  const int inst = unit / nb, bundle = unit - inst * nb;
  const int rpp = blockDim.x / bw;
  const int rl = threadIdx.x / bw, col = threadIdx.x - rl * bw;
  const bool active = rl < rpp;
  const int64_t rbeg = rows * rank / cs, rend = rows * (rank + 1) / cs;
  const int c = bundle * bch + col * 8;  // first channel of this thread's chunk
  const __nv_bfloat16* src = x0;
  int64_t ld = C0;
  if (c < C0) { src = x0 + c; } else { src = x1 + (c - C0); ld = C1; }
  src += static_cast<int64_t>(inst) * rows * ld;
  pdl_wait();
  // ---------------- phase 1 ----------------
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  uint4 v[KMAX > 0 ? KMAX : 1];
  // rows of this thread: rbeg + rl + k * rpp, k < nk
  const int nk = (active && rbeg + rl < rend) ? static_cast<int>((rend - rbeg - rl + rpp - 1) / rpp) : 0;
  const int64_t pstep = static_cast<int64_t>(rpp) * ld;
  if (active) {
    if constexpr (KMAX > 0) {
      const __nv_bfloat16* pp = src + (rbeg + rl) * ld;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < nk) v[k] = *reinterpret_cast<const uint4*>(pp);
        pp += pstep;
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < nk) gn_accum8(v[k], s, q);
    } else {
      int64_t r = rbeg + rl;
      for (; r + 7 * rpp < rend; r += 8 * rpp) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = *reinterpret_cast<const uint4*>(src + (r + static_cast<int64_t>(k) * rpp) * ld);
#pragma unroll
        for (int k = 0; k < 8; ++k) gn_accum8(u[k], s, q);
      }
      for (; r < rend; r += rpp) gn_accum8(*reinterpret_cast<const uint4*>(src + r * ld), s, q);
    }
  }
  // group of each of this thread's channels inside the bundle
  int gj[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gj[j] = (col * 8 + j) / cpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = 0; g < gb; ++g) {
    float ss = 0.f, qq = 0.f;
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ss += (gj[j] == g) ? s[j] : 0.f;
        qq += (gj[j] == g) ? q[j] : 0.f;
      }
    }
    ss = warp_sum(ss);
    qq = warp_sum(qq);
    if (lane == 0) {
      wpart[warp][g][0] = static_cast<double>(ss);
      wpart[warp][g][1] = static_cast<double>(qq);
    }
  }
  // the affine parameters are fetched here, behind the barriers below (and not earlier: the held rows need the registers)
  float ga[8], be[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c) + 1);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
  __syncthreads();
  if (threadIdx.x < gb * 2) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    double a = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) a += wpart[w][g][k];
    part[g][k] = a;
  }
  if (cs > 1) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < gb) {
    const int g = threadIdx.x;
    double a = 0.0, b = 0.0;
    if (cs > 1) {
      for (int rk = 0; rk < cs; ++rk) {  // same order in every CTA: identical statistics across the cluster
        a += ld_dsmem_f64(&part[g][0], rk);
        b += ld_dsmem_f64(&part[g][1], rk);
      }
    } else {
      a = part[g][0];
      b = part[g][1];
    }
    const double cnt = static_cast<double>(rows) * cpg;
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[g][0] = static_cast<float>(mean);
    stat[g][1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  // ---------------- phase 2: normalise this CTA's rows ----------------
  if (active) {
    float4 t[4];
    {
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = stat[gj[j]][1] * ga[j];
        sh[j] = be[j] - stat[gj[j]][0] * sc[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = make_float4(sc[2 * j], sh[2 * j], sc[2 * j + 1], sh[2 * j + 1]);
    }
    __nv_bfloat16* dst = out + static_cast<int64_t>(inst) * rows * Ctot + c;
    if constexpr (KMAX > 0) {
      __nv_bfloat16* dp = dst + (rbeg + rl) * Ctot;
      const int64_t dstep = static_cast<int64_t>(rpp) * Ctot;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < nk) *reinterpret_cast<uint4*>(dp) = gn_affine8(v[k], t, silu);
        dp += dstep;
      }
    } else {
      int64_t r = rbeg + rl;
      for (; r + 3 * rpp < rend; r += 4 * rpp) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(src + (r + static_cast<int64_t>(k) * rpp) * ld);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(dst + (r + static_cast<int64_t>(k) * rpp) * Ctot) = gn_affine8(u[k], t, silu);
      }
      for (; r < rend; r += rpp)
        *reinterpret_cast<uint4*>(dst + r * Ctot) = gn_affine8(*reinterpret_cast<const uint4*>(src + r * ld), t, silu);
    }
  }
  if (cs > 1) cluster_sync_all();  // no CTA leaves while a peer may still read its sums
}

struct GnClusterPlan {
  int ok, gb, cs, threads, kmax, grid;
};

static GnClusterPlan gn_cluster_plan(int n_inst, int64_t rows, int Ctot, int groups) {
  GnClusterPlan pl;
  memset(&pl, 0, sizeof(pl));
  const int cpg = Ctot / groups;
  int g8 = 8;  // gcd(cpg, 8)
  while (cpg % g8 != 0) g8 >>= 1;
  pl.gb = 8 / g8;
  if (groups % pl.gb != 0) return pl;
  const int bw = pl.gb * cpg / 8;
  if (bw > 128) return pl;
  const int64_t units = static_cast<int64_t>(n_inst) * (groups / pl.gb);
  if (units * 8 > (1ll << 30)) return pl;
  // Measured on B200 (tools/norm_probe.py --sweep): the launch should fill the SMs in ONE wave.  Many units -> one
  // small streaming CTA each; few units -> clusters that split the rows, 512 threads, slices held in registers.
  // Re-measured at the end of round 2 (after SiLU left the IEEE division): a unit per CTA - no cluster - whenever the
  // units alone fill the SMs (clusters cost more at launch than they save: 17.4 vs 18.2 us on 24 x 1024 x 320), its
  // slice held in registers by a 128-thread CTA when it is small (24 x 64 x 1280: 5.7 vs 7.6 us); and clusters of 8 for
  // the few-unit strips with >= 4096 chunks per CTA even past one CTA per SM (2 x 3072 x 1920: 20.7 vs 28.3 us).
  bool hold = true, small_unit = false, big8 = false;
  if (units >= 148) {
    pl.cs = 1;
    small_unit = rows * bw < 4096;
    hold = small_unit;
  } else {
    pl.cs = 8;
    while (pl.cs > 1 && units * pl.cs > 160) pl.cs >>= 1;
    if (pl.cs < 8 && units * 8 <= 296 && (rows / 8) * bw >= 4096) {
      pl.cs = 8;
      big8 = true;  // streamed by 512 threads (the measured best; holding 12+ chunks per thread was 35 vs 21 us)
      hold = false;
    }
  }
  while (pl.cs > 1 && rows / pl.cs < 64) pl.cs >>= 1;
#ifdef ASVA_DEBUG_SWITCHES  // plan overrides for tools/norm_probe.py --sweep (experiment builds only)
  const char* e_cs = getenv("ASVA_GN_CS");
  const char* e_t = getenv("ASVA_GN_T");
  const char* e_k = getenv("ASVA_GN_KMAX");
#else
  const char *e_cs = nullptr, *e_t = nullptr, *e_k = nullptr;
#endif
  if (e_cs != nullptr) pl.cs = atoi(e_cs);
  const int64_t rows_cta = (rows + pl.cs - 1) / pl.cs;
  const int64_t chunks = rows_cta * bw;
  if (small_unit)
    pl.threads = 128;
  else if (hold)
    pl.threads = chunks >= 1024 ? 512 : (chunks >= 256 ? 256 : 128);
  else
    pl.threads = (pl.cs == 1 || big8) ? 512 : (chunks >= 512 ? 256 : 128);
  if (e_t != nullptr) pl.threads = atoi(e_t);
  if (pl.threads < bw) pl.threads = ((bw + 31) / 32) * 32;
  const int rpp = pl.threads / bw;
  const int64_t k = (rows_cta + rpp - 1) / rpp;
  pl.kmax = !hold ? 0 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 0)));
  if (e_k != nullptr) pl.kmax = (k <= atoi(e_k)) ? (k <= 4 ? 4 : (k <= 8 ? 8 : 16)) : 0;
  pl.grid = static_cast<int>(units * pl.cs);
  pl.ok = 1;
  return pl;
}

static int g_gn_cap_d[kMaxDevices] = {0};  // co-resident CTAs of gn_fused_kernel, per device

static GnStatsPlan gn_fused_plan(int n_inst, int64_t rows, int Ctot, int cap) {
  GnStatsPlan pl;
  const int nchunk = Ctot / 8;
  pl.cblocks = (nchunk + 127) / 128;
  pl.cw = (nchunk + pl.cblocks - 1) / pl.cblocks;
  pl.rows_per_pass = 256 / pl.cw;
  if (pl.rows_per_pass < 1) pl.rows_per_pass = 1;
  int64_t want = cap / ((int64_t)n_inst * pl.cblocks);
  int64_t max_splits = rows / (2 * (int64_t)pl.rows_per_pass);
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  pl.splits = (int)want;
  return pl;
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
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  const int n = N > 0 ? N : 1, f = F > 0 ? F : 1;
  // sub-warp form: the fewest lanes per row (8, 16, 32) that hold the row in at most 5 chunks each
  for (int ll = 3; ll <= 5; ++ll) {
    const int lanes = 1 << ll;
    if (C % (8 * lanes) != 0 || C / (8 * lanes) > 5) continue;
    const int nch = C / (8 * lanes);
    const int64_t groups = (M + (32 / lanes) - 1) / (32 / lanes);
    int threads = 256, per_sm = 2;  // CTA size and CTAs per SM the grid is sized for (registers allow 512 threads / SM)
#ifdef ASVA_DEBUG_SWITCHES
    if (const char* e = getenv("ASVA_LNR_T")) threads = atoi(e) == 128 ? 128 : 256;
    if (const char* e = getenv("ASVA_LNR_PER_SM")) per_sm = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : per_sm;
#endif
    const int wpb = threads / 32;
    const int64_t cap = static_cast<int64_t>(device_sms()) * per_sm * wpb;  // resident warps
    const int64_t iters = (groups + cap - 1) / cap;
    const unsigned blocks = static_cast<unsigned>((groups + wpb * iters - 1) / (wpb * iters));
    const size_t smem = static_cast<size_t>(C) * 8;
#define ASVA_LNR_CASE(K) \
  case K: ASVA_CUDA_OK(launch_k(layernorm_rows_kernel<K>, dim3(blocks), dim3(threads), smem, stream, 1, xp, gamma, beta, pos, op, M, C, eps, n, f, ll)); break;
    switch (nch) {
      ASVA_LNR_CASE(1) ASVA_LNR_CASE(2) ASVA_LNR_CASE(3) ASVA_LNR_CASE(4)
      default: ASVA_CUDA_OK(launch_k(layernorm_rows_kernel<5>, dim3(blocks), dim3(threads), smem, stream, 1, xp, gamma, beta, pos, op, M, C, eps, n, f, ll)); break;
    }
#undef ASVA_LNR_CASE
    ASVA_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const unsigned blocks = static_cast<unsigned>((M + 7) / 8);
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

// sync_ws layout: [0, 65544) counters + fp64 accumulators of the grid-barrier kernel (kept zeroed by it);
// [kGn2WsOffset, + kGn2WsBytes) per-group partials of the two-launch form (no initial state)
constexpr int64_t kGn2WsOffset = 66048, kGn2WsBytes = 256 * 1024;
extern "C" int64_t asva_groupnorm_sync_bytes(void) { return kGn2WsOffset + kGn2WsBytes; }

namespace asva {
// Which kernel(s) asva_groupnorm runs: 0 = cluster kernel, 1 = grid-barrier kernel, 2 = two launches (stats, apply).
static int gn_pick_form(int n_inst, int64_t rows, int Ctot, int groups, GnClusterPlan* cpo) {
  const GnClusterPlan cp = gn_cluster_plan(n_inst, rows, Ctot, groups);
  if (cpo != nullptr) *cpo = cp;
  // few units x many rows (a whole clip per instance at the top resolution): 8 CTAs per unit cannot fill the GPU
  // and narrow column strips waste DRAM bursts - those shapes read full rows instead
  const int64_t bytes = static_cast<int64_t>(n_inst) * rows * Ctot * 2;
  const bool narrow = cp.ok && (int64_t)n_inst * (groups / cp.gb) <= 16 && bytes > (12ll << 20);
  int form = (cp.ok && !narrow) ? 0 : 1;
  const bool two_ok = Ctot / 8 <= 128 && rows >= 8 * 4 * (512 / (Ctot / 8));  // 8 statistics CTAs of >= 4 passes
  if (form == 1 && two_ok && bytes > (4ll << 20)) form = 2;
  // whole-clip norms with few, wide-strided units (>= 10 MB, <= 16 units): the two-launch form also beats the cluster
  // kernel's column strips (2 x 12288 x 960: 42.4 vs 54.3 us; 2 x 3072 x 960: 20.0 vs 23.6); with 32 units the cluster
  // kernel streams with 8 CTAs per unit and wins (2 x 12288 x 640: 26.7 vs 32.3 us)
  if (form == 0 && two_ok && n_inst <= 4 && bytes > (10ll << 20) && (int64_t)n_inst * (groups / cp.gb) <= 16) form = 2;
#ifdef ASVA_DEBUG_SWITCHES
  if (const char* e = getenv("ASVA_GN_NO_CLUSTER"))
    if (e[0] == '1' && form == 0) form = 1;
  if (const char* e = getenv("ASVA_GN_FORM")) {  // 0 / 1 / 2 where feasible (tools/norm_probe.py)
    const int f = atoi(e);
    if (f == 0 && cp.ok) form = 0;
    if (f == 1) form = 1;
    if (f == 2 && two_ok) form = 2;
  }
#endif
  return form;
}
}  // namespace asva

extern "C" int asva_groupnorm_form(int32_t n_inst, int64_t rows, int32_t C, int32_t groups) {
  if (n_inst < 1 || rows < 1 || C < 8 || groups < 1 || C % groups != 0) return -1;
  return asva::gn_pick_form(n_inst, rows, C, groups, nullptr);
}

extern "C" int asva_groupnorm(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t n_inst, int64_t rows,
                              int32_t groups, float eps, const float* gamma, const float* beta, int32_t silu,
                              void* out, void* sync_ws, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (x1 == nullptr) C1 = 0;
  const int Ctot = C0 + C1;
  ASVA_REQUIRE(x0 && out && sync_ws && gamma && beta, "asva_groupnorm: null operand");
  ASVA_REQUIRE(C0 % 8 == 0 && C1 % 8 == 0 && Ctot >= 8, "asva_groupnorm: channels must be multiples of 8");
  ASVA_REQUIRE(groups >= 1 && Ctot % groups == 0, "asva_groupnorm: C=%d not divisible by groups=%d", Ctot, groups);
  ASVA_REQUIRE(n_inst >= 1 && rows >= 1, "asva_groupnorm: empty problem");
  ASVA_REQUIRE((int64_t)n_inst * groups <= 4096, "asva_groupnorm: n_inst * groups = %lld exceeds the workspace",
               (long long)n_inst * groups);
  GnClusterPlan cp;
  const int form = gn_pick_form(n_inst, rows, Ctot, groups, &cp);
  if (form == 0) {
    const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(x0);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(x1);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define ASVA_GN_LAUNCH(K)                                                                                          \
  ASVA_CUDA_OK(launch_k(gn_cluster_kernel<K>, dim3(cp.grid), dim3(cp.threads), 0, stream, cp.cs, a0, C0, a1, C1, rows, \
                        groups, cp.gb, cp.cs, eps, gamma, beta, silu, o))
    switch (cp.kmax) {
      case 4: ASVA_GN_LAUNCH(4); break;
      case 8: ASVA_GN_LAUNCH(8); break;
      case 16: ASVA_GN_LAUNCH(16); break;
      default: ASVA_GN_LAUNCH(0); break;
    }
#undef ASVA_GN_LAUNCH
    ASVA_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (form == 2) {
    const int sms = device_sms();
    const int cw = Ctot / 8;
    const int rpp = 256 / cw, rpp_a = 512 / cw;  // rows per pass of the apply (256 threads) / the statistics CTA (512)
    // statistics CTAs per cluster (their per-group sums meet in distributed shared memory).  Measured on 2 x 12288 x 320
    // (tools/gn2_cs_probe.sh): 24.2 / 21.8 / 25.4 / 26.4 us for clusters of 1 / 2 / 4 / 8 - pairs halve the partials
    // the apply kernel folds, larger clusters cost more at launch than they save
    int gcs = 2;
#ifdef ASVA_DEBUG_SWITCHES
    if (const char* e = getenv("ASVA_GN2_CS")) gcs = atoi(e);
    if (gcs != 1 && gcs != 2 && gcs != 4 && gcs != 8) gcs = 2;
#endif
    int st_per_sm = 2, ap_per_sm = 4;  // CTAs per SM the two grids are sized for
#ifdef ASVA_DEBUG_SWITCHES
    if (const char* e = getenv("ASVA_GN2_ST_PER_SM")) st_per_sm = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : st_per_sm;
    if (const char* e = getenv("ASVA_GN2_AP_PER_SM")) ap_per_sm = atoi(e) >= 1 && atoi(e) <= 8 ? atoi(e) : ap_per_sm;
#endif
    int64_t splits = (static_cast<int64_t>(st_per_sm) * sms) / n_inst;  // statistics CTAs per instance, in clusters of gcs
    const int64_t max_splits = rows / (4 * static_cast<int64_t>(rpp_a));
    if (splits > max_splits) splits = max_splits;
    const int64_t ws_cap = kGn2WsBytes / (static_cast<int64_t>(n_inst) * groups * 8) * gcs;
    if (splits > ws_cap) splits = ws_cap;
    splits = splits / gcs * gcs;
    if (splits < gcs) splits = gcs;
    ASVA_REQUIRE(static_cast<int64_t>(n_inst) * (splits / gcs) * groups * 8 <= kGn2WsBytes, "asva_groupnorm: workspace too small");
    float* ws2 = reinterpret_cast<float*>(reinterpret_cast<char*>(sync_ws) + kGn2WsOffset);
    const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(x0);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(x1);
    const size_t smem_a = (static_cast<size_t>(rpp_a) * cw * 16 + static_cast<size_t>(Ctot) * 2 + static_cast<size_t>(groups) * 2) * sizeof(float);
    ASVA_REQUIRE(smem_a <= 48 * 1024, "asva_groupnorm: statistics CTA needs %zu bytes of shared memory", smem_a);
    ASVA_CUDA_OK(launch_k(gn2_stats_kernel, dim3(static_cast<unsigned>(splits), n_inst), dim3(512), smem_a, stream, gcs, a0,
                          C0, a1, C1, rows, static_cast<int>(splits), cw, rpp_a, groups, gcs, ws2));
    ASVA_CUDA_OK(cudaGetLastError());
    int64_t bpi = (static_cast<int64_t>(sms) * ap_per_sm + n_inst - 1) / n_inst;
    const int64_t need = rows / (4 * static_cast<int64_t>(rpp));  // at least four passes of rows per CTA
    if (bpi > need) bpi = need;
    if (bpi < 1) bpi = 1;
    const size_t smem_b = static_cast<size_t>(Ctot) * 8 + static_cast<size_t>(groups) * 8;
    ASVA_CUDA_OK(launch_k(gn2_apply_kernel, dim3(static_cast<unsigned>(bpi), n_inst), dim3(256), smem_b, stream, 1, a0, C0,
                          a1, C1, rows, static_cast<int>(splits / gcs), groups, eps, gamma, beta, silu,
                          static_cast<const float*>(ws2), reinterpret_cast<__nv_bfloat16*>(out)));
    ASVA_CUDA_OK(cudaGetLastError());
    return 0;
  }
  // generic form (channel groups that do not bundle into 16-byte strips): grid-barrier kernel
  int& g_gn_cap = g_gn_cap_d[current_device()];
  if (g_gn_cap == 0) {
    int occ = 0;
    const int sms = device_sms();
    ASVA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_fused_kernel, 256, 0));
    ASVA_REQUIRE(occ >= 1 && sms >= 1, "asva_groupnorm: kernel does not fit an SM");
    int per_sm = occ < kGnCtasPerSm ? occ : kGnCtasPerSm;
#ifdef ASVA_DEBUG_SWITCHES
    const char* e = getenv("ASVA_GN_CTAS_PER_SM");
    if (e != nullptr && atoi(e) >= 1 && atoi(e) <= occ) per_sm = atoi(e);
#endif
    g_gn_cap = per_sm * sms;
  }
  const int nchunk = Ctot / 8;
  const int cblocks = (nchunk + 127) / 128;
  ASVA_REQUIRE((int64_t)n_inst * cblocks <= g_gn_cap, "asva_groupnorm: %d instances do not fit a co-resident grid",
               n_inst);
  GnStatsPlan pl = gn_fused_plan(n_inst, rows, Ctot, g_gn_cap);
  const unsigned grid = static_cast<unsigned>(n_inst) * pl.splits * pl.cblocks;
  unsigned* ctr = reinterpret_cast<unsigned*>(sync_ws);
  double* acc = reinterpret_cast<double*>(reinterpret_cast<char*>(sync_ws) + 8);
  ASVA_CUDA_OK(launch_k(gn_fused_kernel, dim3(grid), dim3(256), 0, stream, 1,
                        reinterpret_cast<const __nv_bfloat16*>(x0), C0, reinterpret_cast<const __nv_bfloat16*>(x1), C1,
                        rows, pl.splits, pl.cblocks, pl.cw, pl.rows_per_pass, groups, eps, gamma, beta, silu,
                        reinterpret_cast<__nv_bfloat16*>(out), acc, ctr, n_inst * groups * 2));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}
