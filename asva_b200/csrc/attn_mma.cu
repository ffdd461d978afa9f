// Attention against a SMALL key set (Nk <= 128: the text and audio cross-attentions, the level-2/3 first-frame
// attention) on warp-level MMA.  avgen/models/unets/utils.py:151-153 and diffusers AttnProcessor2_0
// (ff_spatio_audio_temp_transformer_3d.py:315-341) restated for the case where the whole K / V of a group fits in
// shared memory.
//
// Why not the tcgen05 kernel (attn_tc.cu): with one or two key tiles an item there is a serial chain - Q tile by TMA,
// S MMA, two softmax passes over TMEM, P through shared memory, P V MMA, O out of TMEM - of ~2.5 us, of which the
// tensor core works for ~0.1 us; 10 items per CTA run back to back at ~1.1-1.3 TB/s of Q + O traffic (26-30 us for the
// 31 MB of a level-0 cross-attention, 24 us even with ONE key).  The op is memory-bound: 4 Nk d flops per query and
// head against 4 d bytes.  Here:
//   * a CTA stages the K and V rows of (group, chunk of heads spanning <= 320 channels) once - cp.async into rows
//     padded by 16 bytes (conflict-free ldmatrix) - and serves 128 query rows with them;
//   * each of its 8 warps owns 16 query rows: Q fragments come straight from global memory (a quad reads the 32-byte
//     sector of a row's 16 channels), one head ahead of the arithmetic; S = Q K^T, the masked softmax and P V run on
//     mma.sync m16n8k16 fragments exactly as in the temporal kernel (misc.cu); outputs go straight to global memory;
//   * two CTAs per SM (<= 80 keys), a single wave: loads, arithmetic and stores of different warps overlap.
// MEASURED (profiles/r2_attn_mma.md): slower than the tcgen05 kernel on every shape of the workload - 38.5 vs 25.6 us
// (77 keys, level 0), 22.2 vs 21.6 us (25 keys) - and its time follows the mma.sync count (55 per 16 rows and head at
// 77 keys), not the bytes: the legacy warp-MMA path of sm_100 issues far below the tcgen05 rate.  It is therefore NOT
// the default (asva_attn_desc.form = 2 selects it); the temporal attention (misc.cu), with 11 MMAs per problem, is
// where the same fragment code wins.
#include "common.cuh"
#include "host_common.h"

namespace asva {

struct MmaAttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* kv;
  const uint8_t* mask;
  __nv_bfloat16* out;
  int64_t ldq, ldkv, ldo, mask_ld;
  int32_t R, Nk, d, heads, hc_heads, n_chunks, n_row_blocks;
  int32_t kv_rows_per_group, k_col0, v_col0, mask_rows;
  float scale_log2;
};

constexpr int kMmaRowsPerCta = 128;

// NKT: 8-key tiles the scores of a row span (Nk <= 8 * NKT, NKT even); KQ: 16-channel steps of a head (d <= 16 * KQ)
template <int NKT, int KQ>
__global__ void __launch_bounds__(256, (NKT <= 10 && KQ <= 5) ? 2 : 1) attn_mma_kernel(const MmaAttnParams p) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int rb = blockIdx.x % p.n_row_blocks;
  const int hc = (blockIdx.x / p.n_row_blocks) % p.n_chunks;
  const int g = blockIdx.x / (p.n_row_blocks * p.n_chunks);
  const int h0 = hc * p.hc_heads;
  const int hcount = min(p.hc_heads, p.heads - h0);
  const int d = p.d;
  const uint32_t chunk_bytes = static_cast<uint32_t>(hcount * d) * 2u;
  const uint32_t pitch = static_cast<uint32_t>(p.hc_heads * d) * 2u + 16u;
  const int nkp = (p.Nk + 15) & ~15;  // key rows in shared memory (the rows past Nk and the pads are zeroed)
  const uint32_t ks_base = smem_u32(sm), vs_base = ks_base + static_cast<uint32_t>(nkp) * pitch;
  // zero what the copies do not write: rows [Nk, nkp) and every row's tail past the chunk (incl. the 16-byte pad) -
  // MMA operands read there (multiplied by zero probabilities / zeroed A columns) and must be finite
  for (int i = threadIdx.x; i < 2 * nkp; i += blockDim.x) {
    const int row = i % nkp;
    const uint32_t a = (i < nkp ? ks_base : vs_base) + static_cast<uint32_t>(row) * pitch;
    const uint32_t from = (row < p.Nk) ? chunk_bytes : 0u;
    for (uint32_t o = from; o < pitch; o += 16u) st_shared_v4(a + o, 0u, 0u, 0u, 0u);
  }
  pdl_trigger();
  pdl_wait();
  // K and V rows of the chunk: 16-byte cp.async pieces, all threads (a bulk copy per 640-byte row - 154 of them for
  // 77 keys - cost ~0.12 us each in the TMA unit: measured 34.6 us per level-0 text launch, profiles/r2_attn_mma.md)
  {
    const int ppr = static_cast<int>(chunk_bytes >> 4);
    const int total = 2 * p.Nk * ppr;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int rowi = i / ppr, off = i - rowi * ppr;
      const bool isv = rowi >= p.Nk;
      const int row = isv ? rowi - p.Nk : rowi;
      const __nv_bfloat16* src = p.kv + (static_cast<int64_t>(g) * p.kv_rows_per_group + row) * p.ldkv +
                                 (isv ? p.v_col0 : p.k_col0) + h0 * d + off * 8;
      const uint32_t dst = (isv ? vs_base : ks_base) + static_cast<uint32_t>(row) * pitch + static_cast<uint32_t>(off) * 16u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lc = (lane & 3) * 2;
  const int ksteps = (d + 15) >> 4;
  const int row0 = rb * kMmaRowsPerCta + warp * 16;
  const bool warp_on = row0 < p.R;
  // this lane's two query rows (clamped for loads; stores are predicated)
  const int ra = min(row0 + lr, p.R - 1), rbw = min(row0 + lr + 8, p.R - 1);
  const bool oka = row0 + lr < p.R, okb = row0 + lr + 8 < p.R;
  const __nv_bfloat16* qa = p.q + (static_cast<int64_t>(g) * p.R + ra) * p.ldq + h0 * d + lc;
  const __nv_bfloat16* qb = p.q + (static_cast<int64_t>(g) * p.R + rbw) * p.ldq + h0 * d + lc;
  // key-validity bits of the lane's score columns: bit (2 * nt + j) <-> key 8 * nt + lc + j
  uint32_t va = 0u, vb = 0u;
  {
    const uint8_t* ma = nullptr;
    const uint8_t* mb = nullptr;
    if (p.mask != nullptr) {
      ma = p.mask + ((static_cast<int64_t>(g) * p.R + ra) / p.mask_rows) * p.mask_ld;
      mb = p.mask + ((static_cast<int64_t>(g) * p.R + rbw) / p.mask_rows) * p.mask_ld;
    }
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = nt * 8 + lc + j;
        if (key < p.Nk) {
          if (ma == nullptr || __ldg(ma + key) != 0) va |= 1u << (2 * nt + j);
          if (mb == nullptr || __ldg(mb + key) != 0) vb |= 1u << (2 * nt + j);
        }
      }
  }
  constexpr int kMaxQ = 4 * KQ;  // Q fragment registers of one head: 4 per 16-channel step
  uint32_t qn[kMaxQ];
  auto load_q = [&](int h, uint32_t (&dst)[kMaxQ]) {
#pragma unroll
    for (int ks = 0; ks < KQ; ++ks) {
      if (ks < ksteps) {
        const int c = h * d + ks * 16;
        const bool half = (ks * 16 + 8 >= d);
        dst[4 * ks + 0] = __ldg(reinterpret_cast<const uint32_t*>(qa + c));
        dst[4 * ks + 1] = __ldg(reinterpret_cast<const uint32_t*>(qb + c));
        dst[4 * ks + 2] = half ? 0u : __ldg(reinterpret_cast<const uint32_t*>(qa + c + 8));
        dst[4 * ks + 3] = half ? 0u : __ldg(reinterpret_cast<const uint32_t*>(qb + c + 8));
      }
    }
  };
  if (warp_on) load_q(0, qn);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!warp_on) return;

  for (int h = 0; h < hcount; ++h) {
    const uint32_t kh = ks_base + static_cast<uint32_t>(h * d) * 2u, vh = vs_base + static_cast<uint32_t>(h * d) * 2u;
    float s[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
    {
      uint32_t qc[kMaxQ];
#pragma unroll
      for (int i = 0; i < kMaxQ; ++i) qc[i] = qn[i];
      if (h + 1 < hcount) load_q(h + 1, qn);  // the next head's fragments fly under this head's arithmetic
#pragma unroll
      for (int ks = 0; ks < KQ; ++ks) {
        if (ks < ksteps) {
          const uint32_t a[4] = {qc[4 * ks], qc[4 * ks + 1], qc[4 * ks + 2], qc[4 * ks + 3]};
#pragma unroll
          for (int np = 0; np < NKT / 2; ++np) {
            const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
            uint32_t bk[4];
            ldsm_x4(kh + static_cast<uint32_t>(key) * pitch + static_cast<uint32_t>(ks * 16 + ((lane >> 3) & 1) * 8) * 2u, bk);
            mma_bf16_16816(s[2 * np], a, bk[0], bk[1]);
            mma_bf16_16816(s[2 * np + 1], a, bk[2], bk[3]);
          }
        }
      }
    }
    // masked softmax of the two rows (a row with every key masked yields zeros, like the tcgen05 kernel)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool on = (((e < 2) ? va : vb) >> (2 * nt + (e & 1))) & 1u;
        if (!on) s[nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float sum[2] = {0.f, 0.f}, inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      if (mx[r] == -INFINITY) mx[r] = 0.f;
    }
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f((s[nt][e] - mx[e >> 1]) * p.scale_log2);
        s[nt][e] = pv;
        sum[e >> 1] += pv;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      inv[r] = sum[r] > 0.f ? 1.0f / sum[r] : 0.f;
    }
    uint32_t pa[NKT / 2][4];
#pragma unroll
    for (int kt = 0; kt < NKT / 2; ++kt) {
      pa[kt][0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
      pa[kt][1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
      pa[kt][2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      pa[kt][3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
    }
    __nv_bfloat16* oa = p.out + (static_cast<int64_t>(g) * p.R + ra) * p.ldo + (h0 + h) * d + lc;
    __nv_bfloat16* ob = p.out + (static_cast<int64_t>(g) * p.R + rbw) * p.ldo + (h0 + h) * d + lc;
    for (int c0 = 0; c0 < d; c0 += 16) {
      float o[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[t][e] = 0.f;
      const bool two = c0 + 8 < d;
#pragma unroll
      for (int kt = 0; kt < NKT / 2; ++kt) {
        if (kt * 16 < nkp) {
          const int key = kt * 16 + (lane & 15);
          uint32_t bv[4];
          ldsm_x4_t(vh + static_cast<uint32_t>(key) * pitch + static_cast<uint32_t>(two ? c0 + ((lane >> 4) << 3) : c0) * 2u, bv);
          mma_bf16_16816(o[0], pa[kt], bv[0], bv[1]);
          if (two) mma_bf16_16816(o[1], pa[kt], bv[2], bv[3]);
        }
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 1 && !two) continue;
        if (oka) *reinterpret_cast<uint32_t*>(oa + c0 + t * 8) = pack_bf16x2(o[t][0] * inv[0], o[t][1] * inv[0]);
        if (okb) *reinterpret_cast<uint32_t*>(ob + c0 + t * 8) = pack_bf16x2(o[t][2] * inv[1], o[t][3] * inv[1]);
      }
    }
  }
}

template <int NKT, int KQ>
static int launch_attn_mma(const MmaAttnParams& p, int grid, size_t smem, cudaStream_t stream) {
  static bool configured[kMaxDevices] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(attn_mma_kernel<NKT, KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[dev] = true;
  }
  ASVA_CUDA_OK(launch_k(attn_mma_kernel<NKT, KQ>, dim3(static_cast<unsigned>(grid)), dim3(256), smem, stream, 1, p));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

// -> 0 launched, 1 = shape not served (more than 128 keys, or K / V of a head chunk beyond shared memory), other = error
int attention_small_keys(const asva_attn_desc* d, cudaStream_t stream) {
  if (d->Nk > 128 || d->d % 8 != 0 || d->d > 160) return 1;
  MmaAttnParams p;
  memset(&p, 0, sizeof(p));
  int hc = 320 / d->d;  // heads per chunk: <= 320 channels of K and of V per CTA
  if (hc < 1) hc = 1;
  if (hc > d->heads) hc = d->heads;
  const int nkp = (d->Nk + 15) & ~15;
  const size_t pitch = static_cast<size_t>(hc) * d->d * 2 + 16;
  const size_t smem = 2 * static_cast<size_t>(nkp) * pitch;
  if (smem > 220 * 1024) return 1;
  p.q = reinterpret_cast<const __nv_bfloat16*>(d->q);
  p.kv = reinterpret_cast<const __nv_bfloat16*>(d->kv);
  p.mask = d->mask;
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.ldq = d->ldq;
  p.ldkv = d->ldkv;
  p.ldo = d->ldo;
  p.mask_ld = d->mask_ld;
  p.R = d->R;
  p.Nk = d->Nk;
  p.d = d->d;
  p.heads = d->heads;
  p.hc_heads = hc;
  p.n_chunks = (d->heads + hc - 1) / hc;
  p.n_row_blocks = (d->R + kMmaRowsPerCta - 1) / kMmaRowsPerCta;
  p.kv_rows_per_group = d->kv_rows_per_group;
  p.k_col0 = d->k_col0;
  p.v_col0 = d->v_col0;
  p.mask_rows = d->mask_rows > 0 ? d->mask_rows : 1;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  const int64_t grid = static_cast<int64_t>(d->G) * p.n_chunks * p.n_row_blocks;
  if (grid >= (1ll << 31)) return 1;
  const int g32 = static_cast<int>(grid);
  if (d->d <= 80) {
    if (d->Nk <= 32) return launch_attn_mma<4, 5>(p, g32, smem, stream);
    if (d->Nk <= 80) return launch_attn_mma<10, 5>(p, g32, smem, stream);
    return launch_attn_mma<16, 5>(p, g32, smem, stream);
  }
  if (d->Nk <= 32) return launch_attn_mma<4, 10>(p, g32, smem, stream);
  if (d->Nk <= 80) return launch_attn_mma<10, 10>(p, g32, smem, stream);
  return launch_attn_mma<16, 10>(p, g32, smem, stream);
}

}  // namespace asva
