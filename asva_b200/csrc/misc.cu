// Small HBM/latency-bound kernels around the tensor-core path: conv_in im2col, conv_out temporal finish,
// skinny linears (timestep embedding MLP, the 22 time_emb_proj), sinusoidal features, the temporal attention
// core over the frame axis, and the fused CFG + sampler update.
#include "common.cuh"
#include "host_common.h"

namespace asva {

// ------------------------------------------------------------------------------------------------
// conv_in im2col: fp32 latents [Bs][Cl][F][h][w] -> bf16 rows [B*F*h*w][64], column = tap*Cl + c
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in_im2col_kernel(const float* __restrict__ lat,
                                                             __nv_bfloat16* __restrict__ out, int B, int Bs, int Cl,
                                                             int F, int h, int w) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(B) * F * h * w * 8;  // 8 chunks of 8 columns per row
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int chunk = static_cast<int>(i & 7);
  const int64_t row = i >> 3;
  const int x = static_cast<int>(row % w);
  const int y = static_cast<int>((row / w) % h);
  const int f = static_cast<int>((row / (static_cast<int64_t>(w) * h)) % F);
  const int b = static_cast<int>(row / (static_cast<int64_t>(w) * h * F));
  const int bs = b % Bs;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = chunk * 8 + j;
    float val = 0.f;
    if (k < 9 * Cl) {
      const int tap = k / Cl, c = k % Cl;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w)
        val = __ldg(lat + (((static_cast<int64_t>(bs) * Cl + c) * F + f) * h + yy) * w + xx);
    }
    v[j] = val;
  }
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + row * 64 + chunk * 8) = u;
}

// ------------------------------------------------------------------------------------------------
// conv_out finish: y fp32 [B*F*hw][ldy] -> out[b][c][f][n] = y_f + Wt [y_0 ; y_{max(f-1,0)} ; y_f] + bt
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_out_finish_kernel(const float* __restrict__ y, int ldy,
                                                              const float* __restrict__ wt,
                                                              const float* __restrict__ bt, float* __restrict__ out,
                                                              int B, int Co, int F, int hw) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = static_cast<int64_t>(B) * F * hw;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = static_cast<int>(i % hw);
  const int f = static_cast<int>((i / hw) % F);
  const int b = static_cast<int>(i / (static_cast<int64_t>(hw) * F));
  const int fp = f > 0 ? f - 1 : 0;
  const float* y0 = y + ((static_cast<int64_t>(b) * F + 0) * hw + n) * ldy;
  const float* yp = y + ((static_cast<int64_t>(b) * F + fp) * hw + n) * ldy;
  const float* yc = y + ((static_cast<int64_t>(b) * F + f) * hw + n) * ldy;
  float cat[24];
  for (int c = 0; c < Co; ++c) {
    cat[c] = y0[c];
    cat[Co + c] = yp[c];
    cat[2 * Co + c] = yc[c];
  }
  for (int c = 0; c < Co; ++c) {
    float acc = bt[c];
    for (int j = 0; j < 3 * Co; ++j) acc += wt[c * 3 * Co + j] * cat[j];
    out[((static_cast<int64_t>(b) * Co + c) * F + f) * hw + n] = yc[c] + acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Skinny linear (M <= 32 rows): one warp per output column, 16-byte weight loads, rows in groups of 4.
// ------------------------------------------------------------------------------------------------
// The weights never depend on the previous kernel, the activations do: a warp pulls its whole weight row into registers
// (K <= 2048) BEFORE the programmatic-dependent-launch wait, so the 51.6 MB time_emb_proj matrix streams from HBM
// while the tiny kernels in front of it are still running, and only then stages x.
constexpr int kSlChunks = 8;  // 8 x 256 elements per warp row held in registers
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int M, int N, int K, int act_in, int act_out) {
  pdl_trigger();
  extern __shared__ float xs[];  // [4][K]
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool held = K <= 256 * kSlChunks;
  uint4 wreg[kSlChunks];
  if (held && n < N) {
    const __nv_bfloat16* wr = w + static_cast<int64_t>(n) * K;
#pragma unroll
    for (int i = 0; i < kSlChunks; ++i) {
      const int k = lane * 8 + i * 256;
      wreg[i] = (k < K) ? __ldg(reinterpret_cast<const uint4*>(wr + k)) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  pdl_wait();
  const int m0 = blockIdx.y * 4;
  const int mrows = min(4, M - m0);
  for (int r = 0; r < 4; ++r) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float v = (r < mrows) ? x[static_cast<int64_t>(m0 + r) * K + k] : 0.f;
      if (act_in == 1) v = silu_f(v);
      xs[r * K + k] = v;
    }
  }
  __syncthreads();
  if (n >= N) return;
  const __nv_bfloat16* wr = w + static_cast<int64_t>(n) * K;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  auto mac = [&](const uint4& u, int k) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    const float wv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 x0 = *reinterpret_cast<const float4*>(xs + r * K + k);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + r * K + k + 4);
      acc[r] += wv[0] * x0.x + wv[1] * x0.y + wv[2] * x0.z + wv[3] * x0.w + wv[4] * x1.x + wv[5] * x1.y +
                wv[6] * x1.z + wv[7] * x1.w;
    }
  };
  if (held) {
#pragma unroll
    for (int i = 0; i < kSlChunks; ++i) {
      const int k = lane * 8 + i * 256;
      if (k < K) mac(wreg[i], k);
    }
  } else {
    for (int k = lane * 8; k < K; k += 256) mac(*reinterpret_cast<const uint4*>(wr + k), k);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bv = bias != nullptr ? bias[n] : 0.f;
    for (int r = 0; r < mrows; ++r) {
      float v = acc[r] + bv;
      if (act_out == 1) v = silu_f(v);
      out[static_cast<int64_t>(m0 + r) * N + n] = v;
    }
  }
}

__global__ void timestep_features_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim,
                                         int flip) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  // diffusers get_timestep_embedding: exponent = -ln(10000) * k / (half - shift), shift = 0
  const float freq = expf(-logf(10000.f) * static_cast<float>(k) / static_cast<float>(half));
  const float arg = t[b] * freq;
  const float s = sinf(arg), c = cosf(arg);
  float* o = out + static_cast<int64_t>(b) * dim;
  if (flip) {
    o[k] = c;
    o[half + k] = s;
  } else {
    o[k] = s;
    o[half + k] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// CFG combine + sampler update (fp32, frames 1..F-1 only).
// ------------------------------------------------------------------------------------------------
// eps [k][clips][C][F][hw] (CFG branch-major, the pipeline's torch.cat order), lat [clips][C][F][hw],
// hist [4][clips][C][F][hw]; all clips of a launch are at the same sampler step (shared coef / slots).
__global__ void __launch_bounds__(256) cfg_step_kernel(const float* __restrict__ eps, int k, int clips,
                                                       float* __restrict__ lat, float* __restrict__ hist,
                                                       const float* __restrict__ coef,
                                                       const int32_t* __restrict__ slots, int C, int F, int hw,
                                                       int plms) {
  pdl_trigger();
  pdl_wait();
  const int64_t per_c = static_cast<int64_t>(F - 1) * hw;
  const int64_t per_clip = per_c * C;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_clip * clips) return;
  const int64_t clip = i / per_clip;
  const int64_t ic = i - clip * per_clip;
  const int c = static_cast<int>(ic / per_c);
  const int64_t rem = ic % per_c;  // (f-1)*hw + n
  const int64_t stride_b = static_cast<int64_t>(C) * F * hw;
  const int64_t idx = clip * stride_b + (static_cast<int64_t>(c) * F + 1) * hw + rem;
  const int64_t stride_k = stride_b * clips;
  float e = 0.f;
  for (int j = 0; j < k; ++j) e += coef[j] * eps[j * stride_k + idx];
  const float cs = coef[3], ce = coef[4];
  float ehat = e;
  if (plms) {
    // coef[5..8] = weights on (e, hist[slots[1]], hist[slots[2]], hist[slots[3]]); slots[0] = store slot or -1
    ehat = coef[5] * e;
    for (int j = 1; j < 4; ++j) {
      const float a = coef[5 + j];
      if (a != 0.f) ehat += a * hist[static_cast<int64_t>(slots[j]) * stride_k + idx];
    }
    if (slots[0] >= 0) hist[static_cast<int64_t>(slots[0]) * stride_k + idx] = e;
  }
  lat[idx] = cs * lat[idx] + ce * ehat;
}

// ------------------------------------------------------------------------------------------------
// conv_temp operand gather for levels whose frames have fewer than 128 pixels: rows [y_f | y_max(f-1,0) | y_0]
// so that the temporal conv is one plain GEMM with full 128-row tiles (instead of tiles of one frame's few pixels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tconv_gather_kernel(const __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ out, int F, int N, int C,
                                                           int64_t total_chunks) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_chunks) return;
  const int nchunk = 3 * C / 8;
  const int ch = static_cast<int>(i % nchunk);
  const int64_t row = i / nchunk;
  const int part = ch / (C / 8), c = (ch % (C / 8)) * 8;
  const int64_t bf = row / N;
  const int n = static_cast<int>(row % N), f = static_cast<int>(bf % F);
  const int64_t b = bf / F;
  const int fs = part == 0 ? f : (part == 1 ? (f > 0 ? f - 1 : 0) : 0);
  const uint4 u = *reinterpret_cast<const uint4*>(y + ((b * F + fs) * N + n) * C + c);
  *reinterpret_cast<uint4*>(out + row * 3 * C + part * C + c) = u;
}

// ------------------------------------------------------------------------------------------------
// Row softmax for attention whose head dimension exceeds what attn_tc.cu tiles (the VAE's single 512-wide head):
// scores come out of asva_gemm in fp32, probabilities go back into asva_gemm as bf16.  One CTA per row; the row is
// read three times (max, sum, write) - it is a few KB and L2 resident.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int64_t lds,
                                                           __nv_bfloat16* __restrict__ p, int64_t ldp, int cols,
                                                           float scale_log2) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8];
  const float* row = s + static_cast<int64_t>(blockIdx.x) * lds;
  __nv_bfloat16* out = p + static_cast<int64_t>(blockIdx.x) * ldp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float off = m * scale_log2;
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    sum += exp2f(fmaf(v.x, scale_log2, -off)) + exp2f(fmaf(v.y, scale_log2, -off)) +
           exp2f(fmaf(v.z, scale_log2, -off)) + exp2f(fmaf(v.w, scale_log2, -off));
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    uint2 u;
    u.x = pack_bf16x2(exp2f(fmaf(v.x, scale_log2, -off)) * inv, exp2f(fmaf(v.y, scale_log2, -off)) * inv);
    u.y = pack_bf16x2(exp2f(fmaf(v.z, scale_log2, -off)) * inv, exp2f(fmaf(v.w, scale_log2, -off)) * inv);
    *reinterpret_cast<uint2*>(out + c) = u;
  }
}

// ------------------------------------------------------------------------------------------------
// Temporal self-attention over the frame axis, memory-bound form (ff_spatio_audio_temp_transformer_3d.py:343-360: every
// pixel attends over its own F frames; F^2 * d MACs per pixel and head against 4 * C * F bytes of q / k / v / out).
// The tcgen05 form (attn_tc.cu) spends a 128-row MMA tile on 10 pixels x 12 frames and runs at ~1.1 TB/s; here a CTA
// owns P consecutive pixels of one clip: the F row slabs [P][3C] (contiguous in qkv) arrive by bulk-async copies on one
// mbarrier, one thread per (pixel, head, query frame) computes its F scores and its d outputs in fp32 straight from
// shared memory (lanes of the same pixel and head read identical K / V addresses: broadcasts), overwrites its own q
// slot with the bf16 output, and the F * P output rows leave by bulk-async stores.  Several CTAs per SM keep loads,
// arithmetic and stores of different pixel groups overlapped.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int FT>  // frames rounded up (loops over frames unroll, scores stay in registers)
__global__ void __launch_bounds__(512) temporal_rows_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int F, int N, int H, int d,
                                                            int P, float scale_log2) {
  extern __shared__ __align__(128) uint8_t tsm[];
  const int C = H * d;
  const uint32_t row_bytes = static_cast<uint32_t>(3 * C) * 2u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tsm + static_cast<size_t>(F) * P * row_bytes);
  const int gpb = (N + P - 1) / P;
  const int b = blockIdx.x / gpb;
  const int n0 = (blockIdx.x - b * gpb) * P;
  const int pe = min(P, N - n0);
  const uint32_t base = smem_u32(tsm);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(F * pe) * row_bytes);
    for (int f = 0; f < F; ++f)
      bulk_load_1d(base + static_cast<uint32_t>(f * P) * row_bytes,
                   qkv + (static_cast<int64_t>(b * F + f) * N + n0) * 3 * C, static_cast<uint32_t>(pe) * row_bytes, bar);
  }
  mbar_wait(bar, 0);
  const int item = threadIdx.x;
  if (item < pe * H * F) {
    const int fq = item % F, ph = item / F;
    const int h = ph % H, p = ph / H;
    const uint32_t qa = base + static_cast<uint32_t>(fq * P + p) * row_bytes + static_cast<uint32_t>(h * d) * 2u;
    const uint32_t ka = base + static_cast<uint32_t>(p) * row_bytes + static_cast<uint32_t>(C + h * d) * 2u;
    const uint32_t va = ka + static_cast<uint32_t>(C) * 2u;
    const uint32_t fstride = static_cast<uint32_t>(P) * row_bytes;
    float s[FT];
#pragma unroll
    for (int j = 0; j < FT; ++j) s[j] = 0.f;
    for (int c = 0; c < d; c += 8) {
      float q[8];
      bf16x8_to_f32(lds128(qa + c * 2), q);
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        if (j < F) {
          float k[8];
          bf16x8_to_f32(lds128(ka + j * fstride + c * 2), k);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[j] = fmaf(q[e], k[e], s[j]);
        }
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < FT; ++j)
      if (j < F) m = fmaxf(m, s[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < FT; ++j) {
      s[j] = (j < F) ? exp2f((s[j] - m) * scale_log2) : 0.f;
      l += s[j];
    }
    const float inv = 1.0f / l;
    for (int c = 0; c < d; c += 8) {
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < FT; ++j) {
        if (j < F) {
          float v[8];
          bf16x8_to_f32(lds128(va + j * fstride + c * 2), v);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], v[e], o[e]);
        }
      }
      // the output takes the place of this thread's own q chunk (nobody else reads it)
      st_shared_v4(qa + c * 2, pack_bf16x2(o[0] * inv, o[1] * inv), pack_bf16x2(o[2] * inv, o[3] * inv),
                   pack_bf16x2(o[4] * inv, o[5] * inv), pack_bf16x2(o[6] * inv, o[7] * inv));
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  for (int i = threadIdx.x; i < F * pe; i += blockDim.x) {
    const int f = i / pe, p = i - f * pe;
    bulk_store_1d(out + (static_cast<int64_t>(b * F + f) * N + n0 + p) * C,
                  base + static_cast<uint32_t>(f * P + p) * row_bytes, static_cast<uint32_t>(C) * 2u);
  }
  bulk_commit();
  bulk_wait_read<0>();
}

template <int FT>
static int launch_temporal_rows(const void* qkv, void* out, int B, int F, int N, int H, int d, int P, float scale,
                                int threads, size_t smem, cudaStream_t stream) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(temporal_rows_kernel<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev] = true;
  }
  const unsigned grid = static_cast<unsigned>(B) * static_cast<unsigned>((N + P - 1) / P);
  ASVA_CUDA_OK(launch_k(temporal_rows_kernel<FT>, dim3(grid), dim3(threads), smem, stream, 1,
                        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), F, N, H, d, P,
                        scale * 1.4426950408889634f));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

// -> 0 launched, < 0 error, 1 = shape not served by this form (the caller uses the tcgen05 kernel)
int temporal_attention_rows(const void* qkv, void* out, int B, int F, int N, int H, int d, float scale,
                            cudaStream_t stream, bool force) {
  const int C = H * d;
  const int64_t slab = static_cast<int64_t>(F) * 3 * C * 2;  // one pixel: F rows of q | k | v
  if (F > 32 || H * F > 512 || slab + 16 > 200 * 1024 || static_cast<int64_t>(B) * N > (1ll << 30)) return 1;
  // Measured (tools/temporal_probe.py, profiles/r2_temporal_probe.md): one thread per (pixel, head, query) wins for
  // short heads and clips (d = 40: 32.9 vs 55.9 us at 12 frames x 1024 px); with d >= 80 a thread's serial dot
  // products, with F > 16 its F^2 growth, make the tcgen05 form the faster one.
  if (!force && (d > 40 || F > 16)) return 1;
  int P = static_cast<int>((56 * 1024) / slab);  // ~4 CTAs per SM
  if (P < 1) P = 1;
  if (P > 4) P = 4;
  while (P > 1 && P * H * F > 512) --P;
  if (P > N) P = N;
  const int threads = ((P * H * F + 31) / 32) * 32;
  const size_t smem = static_cast<size_t>(P) * slab + 16;
  if (F <= 8) return launch_temporal_rows<8>(qkv, out, B, F, N, H, d, P, scale, threads, smem, stream);
  if (F <= 12) return launch_temporal_rows<12>(qkv, out, B, F, N, H, d, P, scale, threads, smem, stream);
  if (F <= 16) return launch_temporal_rows<16>(qkv, out, B, F, N, H, d, P, scale, threads, smem, stream);
  if (F <= 24) return launch_temporal_rows<24>(qkv, out, B, F, N, H, d, P, scale, threads, smem, stream);
  return launch_temporal_rows<32>(qkv, out, B, F, N, H, d, P, scale, threads, smem, stream);
}

// ------------------------------------------------------------------------------------------------
// Temporal attention, warp-MMA form.  Same data path as temporal_rows_kernel (bulk-async slabs in, bulk-async rows
// out, outputs written over the q slots) but the F x F problem of a (pixel, head) is ONE warp's job on mma.sync
// m16n8k16 (bf16 in, fp32 accumulate): S = Q K^T from ldmatrix fragments, softmax on the accumulator fragments (a row
// lives in the four lanes of a quad), P re-used in registers as the A operand of P V (V through ldmatrix.trans).
// ~120 warp instructions per problem instead of ~850: the kernel is then bound by its 8 C F bytes per pixel.  tcgen05
// does not fit this shape - its smallest tile is 64 x 8 x 16 with the accumulator in TMEM, here a problem is 12 x 12 x 40.
// Shared memory rows are padded by 16 bytes so the 8 rows of an ldmatrix tile fall into distinct bank groups.
// ------------------------------------------------------------------------------------------------
template <int MT>  // 16-frame tiles: F <= 16 * MT
__global__ void __launch_bounds__(256) temporal_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ out, int F, int N, int H, int d,
                                                           int P, float scale_log2) {
  extern __shared__ __align__(128) uint8_t tsm[];
  constexpr int NT = 2 * MT;  // 8-key tiles
  const int C = H * d;
  const uint32_t row_bytes = static_cast<uint32_t>(3 * C) * 2u;
  const uint32_t pitch = row_bytes + 16u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tsm + static_cast<size_t>(F) * P * pitch);
  const int gpb = (N + P - 1) / P;
  const int b = blockIdx.x / gpb;
  const int n0 = (blockIdx.x - b * gpb) * P;
  const int pe = min(P, N - n0);
  const uint32_t base = smem_u32(tsm);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(F * pe) * row_bytes);
  __syncthreads();
  for (int i = threadIdx.x; i < F * pe; i += blockDim.x) {  // row (p, f) of the slab <- qkv row (b, f, n0 + p)
    const int p = i / F, f = i - p * F;
    bulk_load_1d(base + static_cast<uint32_t>(i) * pitch, qkv + (static_cast<int64_t>(b * F + f) * N + n0 + p) * 3 * C,
                 row_bytes, bar);
  }
  mbar_wait(bar, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int lr = lane >> 2, lc = (lane & 3) * 2;  // accumulator fragment: rows lr, lr + 8; columns lc, lc + 1
  const int ksteps = (d + 15) >> 4;
  for (int ph = warp; ph < pe * H; ph += nwarps) {
    const int p = ph / H, h = ph - p * H;
    const uint32_t q0 = base + static_cast<uint32_t>(p * F) * pitch + static_cast<uint32_t>(h * d) * 2u;  // row f: + f * pitch
    const uint32_t k0 = q0 + static_cast<uint32_t>(C) * 2u, v0 = k0 + static_cast<uint32_t>(C) * 2u;
    float s[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
    for (int ks = 0; ks < ksteps; ++ks) {
      const bool half = (ks * 16 + 8 >= d);  // the step's upper 8 channels lie past the head: zero them in A
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = min(mt * 16 + (lane & 15), F - 1);
        ldsm_x4(q0 + static_cast<uint32_t>(row) * pitch + static_cast<uint32_t>(ks * 16 + (lane >> 4) * 8) * 2u, a[mt]);
        if (half) a[mt][2] = a[mt][3] = 0u;
      }
#pragma unroll
      for (int np = 0; np < MT; ++np) {  // 16 keys per ldmatrix.x4: two 8-key tiles
        const int key = min(np * 16 + (lane & 7) + ((lane >> 4) << 3), F - 1);
        uint32_t bk[4];
        ldsm_x4(k0 + static_cast<uint32_t>(key) * pitch + static_cast<uint32_t>(ks * 16 + ((lane >> 3) & 1) * 8) * 2u, bk);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(s[mt][2 * np], a[mt], bk[0], bk[1]);
          mma_bf16_16816(s[mt][2 * np + 1], a[mt], bk[2], bk[3]);
        }
      }
    }
    // softmax over the keys of each row (keys >= F masked); probabilities packed as the A operand of P V
    uint32_t pa[MT][MT][4];  // [m tile][16-key step]
    float inv[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = nt * 8 + lc + (e & 1);
          if (key >= F) s[mt][nt][e] = -INFINITY;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[mt][nt][e]);
        }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f((s[mt][nt][e] - mx[e >> 1]) * scale_log2);
          s[mt][nt][e] = pv;
          sum[e >> 1] += pv;
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
        inv[mt][r] = 1.0f / sum[r];
      }
#pragma unroll
      for (int kt = 0; kt < MT; ++kt) {
        pa[mt][kt][0] = pack_bf16x2(s[mt][2 * kt][0], s[mt][2 * kt][1]);
        pa[mt][kt][1] = pack_bf16x2(s[mt][2 * kt][2], s[mt][2 * kt][3]);
        pa[mt][kt][2] = pack_bf16x2(s[mt][2 * kt + 1][0], s[mt][2 * kt + 1][1]);
        pa[mt][kt][3] = pack_bf16x2(s[mt][2 * kt + 1][2], s[mt][2 * kt + 1][3]);
      }
    }
    __syncwarp();  // every lane is done with q before the outputs take its place
    // O = P V, 16 channels (two 8-channel tiles) per step; the last step of a head with d % 16 == 8 keeps its first tile
    for (int c0 = 0; c0 < d; c0 += 16) {
      float o[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int e = 0; e < 4; ++e) o[mt][t][e] = 0.f;
      const bool two = c0 + 8 < d;
#pragma unroll
      for (int kt = 0; kt < MT; ++kt) {
        const int key = min(kt * 16 + (lane & 15), F - 1);
        const int ch = c0 + ((lane >> 4) << 3);
        uint32_t bv[4];
        ldsm_x4_t(v0 + static_cast<uint32_t>(key) * pitch + static_cast<uint32_t>(two ? ch : c0) * 2u, bv);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16_16816(o[mt][0], pa[mt][kt], bv[0], bv[1]);
          if (two) mma_bf16_16816(o[mt][1], pa[mt][kt], bv[2], bv[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t == 1 && !two) continue;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int row = mt * 16 + lr + r * 8;
            if (row < F)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(q0 + static_cast<uint32_t>(row) * pitch +
                                                          static_cast<uint32_t>(c0 + t * 8 + lc) * 2u),
                           "r"(pack_bf16x2(o[mt][t][2 * r] * inv[mt][r], o[mt][t][2 * r + 1] * inv[mt][r]))
                           : "memory");
          }
        }
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  for (int i = threadIdx.x; i < F * pe; i += blockDim.x) {
    const int p = i / F, f = i - p * F;
    bulk_store_1d(out + (static_cast<int64_t>(b * F + f) * N + n0 + p) * C, base + static_cast<uint32_t>(i) * pitch,
                  static_cast<uint32_t>(C) * 2u);
  }
  bulk_commit();
  bulk_wait_read<0>();
}

template <int MT>
static int launch_temporal_mma(const void* qkv, void* out, int B, int F, int N, int H, int d, int P, float scale,
                               size_t smem, cudaStream_t stream) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(temporal_mma_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev] = true;
  }
  const unsigned grid = static_cast<unsigned>(B) * static_cast<unsigned>((N + P - 1) / P);
  // one warp per (pixel, head) for the longer problems; two heads per warp for d = 40 x <= 16 frames, where more CTAs in
  // flight beat more warps per CTA (18.6 vs 19.8 us at level 0; 6.6 vs 7.6 us the other way at d = 160)
  int threads = (P * H >= 8 && (d >= 80 || F > 16)) ? 256 : 128;
#ifdef ASVA_DEBUG_SWITCHES
  if (const char* e = getenv("ASVA_TMMA_T")) threads = atoi(e) == 128 ? 128 : 256;
#endif
  ASVA_CUDA_OK(launch_k(temporal_mma_kernel<MT>, dim3(grid), dim3(threads), smem, stream, 1,
                        reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), F, N, H, d, P,
                        scale * 1.4426950408889634f));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

// -> 0 launched, 1 = shape not served (F > 32 or a pixel's slab beyond shared memory), other = error code
int temporal_attention_mma(const void* qkv, void* out, int B, int F, int N, int H, int d, float scale,
                           cudaStream_t stream) {
  const int C = H * d;
  const int64_t slab = static_cast<int64_t>(F) * (3 * C * 2 + 16);  // one pixel: F padded rows of q | k | v
  if (F > 32 || slab + 16 > 220 * 1024 || static_cast<int64_t>(B) * N > (1ll << 30)) return 1;
  // one pixel per CTA: measured best on every level (tools/temporal_probe.py with ASVA_TMMA_P = 1 / 2 / 3 / 4: 18.3 /
  // 22.4 / 26.0 / 28.3 us at 12 frames x 1024 px x 320 channels) - small CTAs, many in flight, overlap loads and stores
  int P = 1;
#ifdef ASVA_DEBUG_SWITCHES
  if (const char* e = getenv("ASVA_TMMA_P")) {  // pixels per CTA (tools/temporal_probe.py sweeps)
    const int v = atoi(e);
    if (v >= 1 && v <= 8 && static_cast<int64_t>(v) * slab + 16 <= 220 * 1024) P = v;
  }
#endif
  if (P > N) P = N;
  const size_t smem = static_cast<size_t>(P) * slab + 16;
  if (F <= 16) return launch_temporal_mma<1>(qkv, out, B, F, N, H, d, P, scale, smem, stream);
  return launch_temporal_mma<2>(qkv, out, B, F, N, H, d, P, scale, smem, stream);
}

}  // namespace asva

extern "C" int asva_softmax_rows(const float* scores, int64_t lds, void* probs, int64_t ldp, int64_t rows,
                                 int32_t cols, float scale, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(scores && probs, "asva_softmax_rows: null operand");
  ASVA_REQUIRE(rows >= 1 && rows < (1ll << 31) && cols >= 4 && cols % 4 == 0, "asva_softmax_rows: bad shape");
  ASVA_REQUIRE(lds % 4 == 0 && ldp % 4 == 0 && lds >= cols && ldp >= cols, "asva_softmax_rows: bad leading dimension");
  ASVA_REQUIRE(scale > 0.f, "asva_softmax_rows: scale must be positive");
  ASVA_CUDA_OK(launch_k(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0, stream, 1, scores, lds,
                        reinterpret_cast<__nv_bfloat16*>(probs), ldp, cols, scale * 1.4426950408889634f));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_tconv_gather(const void* y, void* out, int32_t B, int32_t F, int32_t N, int32_t C,
                                 asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(y && out, "asva_tconv_gather: null operand");
  ASVA_REQUIRE(B >= 1 && F >= 1 && N >= 1 && C >= 8 && C % 8 == 0, "asva_tconv_gather: bad shape");
  const int64_t total = static_cast<int64_t>(B) * F * N * (3 * C / 8);
  ASVA_CUDA_OK(launch_k(tconv_gather_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 1,
                        reinterpret_cast<const __nv_bfloat16*>(y), reinterpret_cast<__nv_bfloat16*>(out), F, N, C,
                        total));
  return 0;
}

extern "C" int asva_conv_in_im2col(const float* latents, void* out, int32_t B, int32_t Bs, int32_t Cl, int32_t F,
                                   int32_t h, int32_t w, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(latents && out, "asva_conv_in_im2col: null operand");
  ASVA_REQUIRE(Cl >= 1 && 9 * Cl <= 64, "asva_conv_in_im2col: Cl=%d unsupported (9*Cl must be <= 64)", Cl);
  ASVA_REQUIRE(B >= 1 && Bs >= 1 && F >= 1 && h >= 1 && w >= 1, "asva_conv_in_im2col: empty problem");
  const int64_t total = static_cast<int64_t>(B) * F * h * w * 8;
  ASVA_CUDA_OK(launch_k(conv_in_im2col_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 1, latents, reinterpret_cast<__nv_bfloat16*>(out), B, Bs, Cl, F, h, w));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_conv_out_finish(const float* y, int32_t ldy, const float* wt, const float* bt, float* out,
                                    int32_t B, int32_t Co, int32_t F, int32_t h, int32_t w, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(y && wt && bt && out, "asva_conv_out_finish: null operand");
  ASVA_REQUIRE(Co >= 1 && Co <= 8 && ldy >= Co, "asva_conv_out_finish: Co=%d unsupported", Co);
  const int64_t total = static_cast<int64_t>(B) * F * h * w;
  ASVA_CUDA_OK(launch_k(conv_out_finish_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 1, y, ldy, wt, bt, out, B, Co,
                                                                                         F, h * w));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_small_linear(const float* x, const void* w, const float* bias, float* out, int32_t M, int32_t N,
                                 int32_t K, int32_t act_in, int32_t act_out, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(x && w && out, "asva_small_linear: null operand");
  ASVA_REQUIRE(M >= 1 && M <= 32 && N >= 1 && K >= 8 && K % 8 == 0 && K <= 8192, "asva_small_linear: bad shape");
  static bool configured[kMaxDevices] = {false};  // per-device function attribute
  const int dev = current_device();
  if (!configured[dev]) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(small_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 * 4));
    configured[dev] = true;
  }
  dim3 grid((N + 7) / 8, (M + 3) / 4);
  ASVA_CUDA_OK(launch_k(small_linear_kernel, dim3(grid), dim3(256), static_cast<size_t>(4) * K * sizeof(float), stream, 1, x, reinterpret_cast<const __nv_bfloat16*>(w), bias, out, M, N, K, act_in, act_out));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_timestep_features(const float* t, float* out, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                                      asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(t && out && B >= 1 && dim >= 2 && dim % 2 == 0, "asva_timestep_features: bad arguments");
  const int total = B * (dim / 2);
  ASVA_CUDA_OK(launch_k(timestep_features_kernel, dim3((total + 127) / 128), dim3(128), 0, stream, 1, t, out, B, dim, flip_sin_to_cos));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_cfg_ddim_step(const float* eps, int32_t k, int32_t clips, float* latents, const float* coef,
                                  int32_t C, int32_t F, int32_t hw, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(eps && latents && coef && k >= 1 && k <= 3 && clips >= 1 && F >= 2,
               "asva_cfg_ddim_step: bad arguments");
  const int64_t total = static_cast<int64_t>(clips) * C * (F - 1) * hw;
  ASVA_CUDA_OK(launch_k(cfg_step_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 1,
                        eps, k, clips, latents, nullptr, coef, nullptr, C, F, hw, 0));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int asva_cfg_plms_step(const float* eps, int32_t k, int32_t clips, float* latents, float* hist,
                                  const float* coef, const int32_t* slots, int32_t C, int32_t F, int32_t hw,
                                  asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(eps && latents && hist && coef && slots && k >= 1 && k <= 3 && clips >= 1 && F >= 2,
               "asva_cfg_plms_step: bad arguments");
  const int64_t total = static_cast<int64_t>(clips) * C * (F - 1) * hw;
  ASVA_CUDA_OK(launch_k(cfg_step_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 1,
                        eps, k, clips, latents, hist, coef, slots, C, F, hw, 1));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}
