// Fused attention  O = softmax(Q K^T * scale [+ mask]) V  on tcgen05 for sm_100a.
//
// Persistent kernel, one CTA per SM. A work item is one tile of 128 query rows of one (group, head); the CTA keeps
// NWG (1 or 2) items in flight, one per softmax warpgroup, so the tensor core works for one warpgroup while the other
// is in its softmax ("ping-pong"), and loads of the next item overlap the tail of the current one.
//   warp 0          TMA producer: the Q tile of each warpgroup's next item, then the key tiles of both in-flight items
//                   interleaved through one shared ring: K box(es) (K-major) and V box(es) (MN-major, read straight
//                   from the token-major [rows][channels] projection output - no transposed copy). Q comes from the
//                   token-major projection too: the tensor map is (d, heads, rows) and the 64-wide box zero-fills
//                   the padding past d.
//   warp 1, last    TMEM allocator (warp 1) + one single-thread MMA issuer per warpgroup:  S_w = Q_w K^T  (TMEM cols
//                   [w*KV, ..)),  O_w += P_w V
//   warps 2..5/6..9 softmax warpgroups (thread = query row = TMEM lane). Pass 1 reads S for the row maximum, pass 2
//                   re-reads S, exponentiates (exp2, log2e folded into the scale), writes P as bf16 into a
//                   128B-swizzled smem tile (the A operand of the second MMA) and rescales O in TMEM when the running
//                   maximum moved; after the last key tile it normalises O and stores the rows.
// Reference semantics: F.scaled_dot_product_attention with an optional boolean keep-mask
// (avgen/models/unets/utils.py:151-153 and diffusers AttnProcessor2_0).
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>
#include <string.h>

namespace asva {

struct AttnKParams {
  CUtensorMap tmQ, tmKV;   // temporal mode: tmQ = the q columns, tmKV = the k columns ...
  CUtensorMap tmV;         // ... and tmV = the v columns, all as 5-D (d, heads, N, F, B) views of the qkv buffer
  const uint8_t* mask;
  __nv_bfloat16* out;
  int64_t ldo, mask_ld;
  int32_t R, Nk, d, dN, heads, k_col0, v_col0, mask_rows;
  int32_t ksteps_qk;  // ceil(d/16)
  int32_t nvb;        // number of 64-wide V column blocks = ceil(dN/64)
  int32_t n_qt;       // query tiles per (group, head)
  int32_t total_items;
  long long* trace;         // ASVA_ATTN_TRACE=1: per-phase clock64 sums of block 0, warpgroup 0, thread 0 (debug)
  int32_t tP, tF, tN, tPB;  // temporal mode: pixels per tile, frames, pixels per frame, pixel blocks per clip
  float scale_log2;   // scale * log2(e)
};

constexpr int kPolyOf8 = 4;  // of every 8 exponentials of an unmasked tile, this many avoid the MUFU pipe (packed FFMA2 form)

// NB = S buffers in TMEM and P buffers in shared memory per warpgroup. NB = 2 lets the tensor core compute S(j+1) while
// the softmax works on S(j), and lets the softmax write P(j) while P V(j-1) still reads P(j-1): the per-tile critical
// path shrinks from (MMA issue + softmax) to the softmax alone. It needs 64-key tiles for head dims <= 64 and does
// not fit shared memory for larger heads.
template <int DKA, int KV, int NWG, int STAGES, int NB = 1>
struct AttnCfg {
  static constexpr int kQBytes = DKA * 128 * 128;
  static constexpr int kKBytes = DKA * KV * 128;
  static constexpr int kVBytes = DKA * KV * 128;  // nvb <= DKA
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kPBytes = (KV / 64) * 128 * 128;
  static constexpr int kONCols = DKA * 64 + 16;  // TMEM columns per O accumulator + the 16 row-sum columns after it
  static constexpr int kTmemNeed = NWG * (NB * KV + kONCols);
  static constexpr int kTmemCols = kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512);
  static constexpr int kOnesBytes = 2048;  // 16 x 64 bf16 ones: B operand of the row-sum MMA
  static constexpr int kSmemBytes = NWG * (kQBytes + NB * kPBytes) + STAGES * kStageBytes + kOnesBytes + 512;
  static constexpr int kThreads = 64 + 128 * NWG + (NWG > 1 ? 32 : 0);  // + the second MMA-issuing warp
};

__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2, flush-to-zero; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / integer pipes (no MUFU): round-to-nearest split x = n + f through the 1.5 * 2^23 magic number,
// degree-3 minimax 2^f on [-0.5, 0.5] (relative error 7.7e-5, far below the bf16 rounding of P), n added into the
// exponent field. The softmax of the spatial attention is bound by the MUFU pipe (measured: ~12 cycles per warp-wide
// ex2 with two warps per scheduler), so a share of the exponentials is computed this way instead, as in
// FlashAttention-4.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float q = fmaf(0.05508868f, f, 0.24260405f);
  q = fmaf(q, f, 0.69327623f);
  q = fmaf(q, f, 0.99992895f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2: two fp32 lanes of a 64-bit register pair per instruction). The
// exponential pass issues at the rate of the FMA / ALU pipes (two softmax warps per scheduler, ~310 instructions per
// 64-key tile and warp): the scale FFMA and the polynomial's FMAs / ADDs run on pairs, halving their issue slots.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// poly_exp2 on a pair
__device__ __forceinline__ void poly_exp2_x2(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x2 = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t2 = add_f32x2(x2, pack_f32x2(12582912.f, 12582912.f));
  const uint64_t r2 = add_f32x2(t2, pack_f32x2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma_f32x2(r2, pack_f32x2(-1.f, -1.f), x2);  // x - round(x)
  uint64_t q2 = fma_f32x2(pack_f32x2(0.05508868f, 0.05508868f), f2, pack_f32x2(0.24260405f, 0.24260405f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.69327623f, 0.69327623f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.99992895f, 0.99992895f));
  float q0, q1, t0, t1;
  unpack_f32x2(q2, q0, q1);
  unpack_f32x2(t2, t0, t1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// idesc for P(bf16, K-major, from smem) x V(bf16, MN-major): b_major bit 16 set
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(uint32_t M, uint32_t N) {
  return make_idesc_bf16(M, N) | (1u << 16);
}
// smem descriptor for an MN-major, 128B-swizzled B operand stored as [64-col block][key][64 cols]:
//   LBO = byte pitch between 64-column blocks, SBO = 1024 (8-key groups)
__device__ __forceinline__ uint64_t make_sdesc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// TEMPORAL: attention over the frame axis per pixel (ff_spatio_audio_temp_transformer_3d.py:352-358). An item is a
// block of tP pixels of one (clip, head): its tP*tF (<= KV) rows, ordered (frame, pixel), are both the queries and the
// single key tile, fetched with 5-D boxes straight from the [b][f][n][3C] projection output; row r attends key c iff
// they belong to the same pixel (c % tP == r % tP) - a block-diagonal mask evaluated arithmetically.
template <int DKA, int KV, int NWG, int STAGES, int NB, bool TEMPORAL, int POLY = kPolyOf8, bool PIPE = true>
__global__ void __launch_bounds__(64 + 128 * NWG + (NWG > 1 ? 32 : 0), 1) attn_tc_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<DKA, KV, NWG, STAGES, NB>;
  static_assert(STAGES >= (NB + 1) * NWG, "key ring too shallow for the tiles in flight");
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared memory of a kernel without static shared memory starts at the (1024-byte aligned) base of the
  // CTA's window; the 128B-swizzled tiles need that alignment and there is no room for slack - checked, not assumed
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("[asva] attention: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;                                // [NWG][kQBytes]
  uint8_t* sKV = sQ + NWG * Cfg::kQBytes;            // [STAGES][K | V]
  uint8_t* sP = sKV + STAGES * Cfg::kStageBytes;     // [NWG][NB][kPBytes]
  uint8_t* sOnes = sP + NWG * NB * Cfg::kPBytes;     // [16][64] bf16 1.0
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::kOnesBytes);
  uint64_t* q_full = bars;              // [2]      Q tile of the warpgroup's item landed
  uint64_t* q_free = bars + 2;          // [2]      every S = Q K^T of the item has been issued and completed
  uint64_t* o_free = bars + 4;          // [2]      the warpgroup has read the item's O out of TMEM
  uint64_t* s_full = bars + 6;          // [2][NB]  per warpgroup and S buffer
  uint64_t* p_full = bars + 6 + 2 * NB;   // [2][NB]  P tile written (and S consumed)
  uint64_t* pv_done = bars + 6 + 4 * NB;  // [2][NB]
  uint64_t* kv_full = bars + 6 + 6 * NB;  // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + STAGES);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.Nk + KV - 1) / KV;
  const int item_stride = gridDim.x * NWG;

  if (threadIdx.x == 0) {
    for (int w = 0; w < 2; ++w) {
      mbar_init(&q_full[w], 1);
      mbar_init(&q_free[w], 1);
      mbar_init(&o_free[w], 128);
      for (int b = 0; b < NB; ++b) {
        mbar_init(&s_full[w * NB + b], 1);
        mbar_init(&p_full[w * NB + b], 128);
        mbar_init(&pv_done[w * NB + b], 1);
      }
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;  // bf16 1.0 pairs
  fence_proxy_async_smem();
  if constexpr (TEMPORAL) {
    // the boxes write tP*tF rows; the rows above them are multiplied too (as masked keys / unused queries) and must
    // not hold NaN bit patterns: clear Q and the key ring once
    for (int i = threadIdx.x; i < (NWG * Cfg::kQBytes + STAGES * Cfg::kStageBytes) / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      const uint32_t kv_tx = static_cast<uint32_t>(DKA + p.nvb) * KV * 128u;
      uint32_t pos = 0;
      for (int round = 0;; ++round) {
        const int item0 = round * item_stride + static_cast<int>(blockIdx.x) * NWG;
        if (item0 >= p.total_items) break;
        int gh[NWG];
#pragma unroll
        for (int w = 0; w < NWG; ++w) {
          const int item = item0 + w;
          gh[w] = -1;
          if (item >= p.total_items) continue;
          const int qt = item % p.n_qt;
          gh[w] = item / p.n_qt;  // g * heads + head
          const int g = gh[w] / p.heads, head = gh[w] % p.heads;
          mbar_wait(&q_free[w], (round & 1) ^ 1);
          if constexpr (TEMPORAL) {
            mbar_arrive_expect_tx(&q_full[w], static_cast<uint32_t>(DKA) * p.tP * p.tF * 128u);
#pragma unroll
            for (int a = 0; a < DKA; ++a)
              tma_load_5d(sQ + w * Cfg::kQBytes + a * 128 * 128, &p.tmQ, &q_full[w], a * 64, head, qt * p.tP, 0, g);
          } else {
            mbar_arrive_expect_tx(&q_full[w], Cfg::kQBytes);
#pragma unroll
            for (int a = 0; a < DKA; ++a)
              tma_load_3d(sQ + w * Cfg::kQBytes + a * 128 * 128, &p.tmQ, &q_full[w], a * 64, head, g * p.R + qt * 128);
          }
        }
        for (int j = 0; j < n_tiles; ++j) {
#pragma unroll
          for (int w = 0; w < NWG; ++w) {
            if (gh[w] < 0) continue;
            const int g = gh[w] / p.heads, head = gh[w] % p.heads;
            const uint32_t s = pos % STAGES, ph = (pos / STAGES) & 1u;
            mbar_wait(&kv_empty[s], ph ^ 1u);
            uint8_t* sk = sKV + s * Cfg::kStageBytes;
            uint8_t* sv = sk + Cfg::kKBytes;
            if constexpr (TEMPORAL) {
              const int qt = (item0 + w) % p.n_qt;
              mbar_arrive_expect_tx(&kv_full[s], static_cast<uint32_t>(DKA + p.nvb) * p.tP * p.tF * 128u);
#pragma unroll
              for (int a = 0; a < DKA; ++a)
                tma_load_5d(sk + a * KV * 128, &p.tmKV, &kv_full[s], a * 64, head, qt * p.tP, 0, g);
              for (int a = 0; a < p.nvb; ++a)
                tma_load_5d(sv + a * KV * 128, &p.tmV, &kv_full[s], a * 64, head, qt * p.tP, 0, g);
            } else {
              mbar_arrive_expect_tx(&kv_full[s], kv_tx);
#pragma unroll
              for (int a = 0; a < DKA; ++a)
                tma_load_3d(sk + a * KV * 128, &p.tmKV, &kv_full[s], p.k_col0 + head * p.d + a * 64, j * KV, g);
              for (int a = 0; a < p.nvb; ++a)
                tma_load_3d(sv + a * KV * 128, &p.tmKV, &kv_full[s], p.v_col0 + head * p.d + a * 64, j * KV, g);
            }
            ++pos;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2 + 4 * NWG) {
    // ---------------- MMA issuers: one thread per warpgroup (warp 1 -> warpgroup 0, the last warp -> warpgroup 1), so a
    // warpgroup's S = Q K^T never queues behind the other warpgroup's barrier waits ----------------
    const int w = (warp == 1) ? 0 : 1;
    if (w < NWG) {
      // warp-uniform loop: every lane runs it with identical values, the tcgen05 instructions are predicated on one
      // elected lane - the issue cost per MMA (descriptor adds in uniform registers) is what paces this warp
      const uint32_t el = elect_one();
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KV);
      constexpr uint32_t idesc_l = make_idesc_bf16(128, 16);
      const uint32_t idesc_o = make_idesc_bf16_bmn(128, static_cast<uint32_t>(p.dN));
      constexpr uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);          // K-major, SW128
      constexpr uint64_t desc_hi_mn = desc_hi | (static_cast<uint64_t>((KV * 128u) >> 4) << 16);  // MN-major: LBO
      auto lo = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
      const uint32_t q_lo = lo(smem_u32(sQ + w * Cfg::kQBytes)) | (1u << 16);
      const uint32_t p_lo0 = lo(smem_u32(sP + w * NB * Cfg::kPBytes)) | (1u << 16);
      const uint32_t kv_lo0 = lo(smem_u32(sKV));
      const uint64_t ones_desc = desc_hi | (lo(smem_u32(sOnes)) | (1u << 16));
      const uint32_t tmem_s0 = tmem_base + w * (NB * KV + Cfg::kONCols);
      const uint32_t tmem_o = tmem_s0 + NB * KV;
      const uint32_t tmem_l = tmem_o + p.dN;
      const uint32_t kvf0 = smem_u32(kv_full), kve0 = smem_u32(kv_empty);
      const uint32_t sfull0 = smem_u32(&s_full[w * NB]), pfull0 = smem_u32(&p_full[w * NB]);
      const uint32_t pvdone0 = smem_u32(&pv_done[w * NB]);
      const uint32_t qfree = smem_u32(&q_free[w]), qfull = smem_u32(&q_full[w]), ofree = smem_u32(&o_free[w]);
      // this warpgroup's items are item(it) = (it * gridDim.x + blockIdx.x) * NWG + w; its key tiles form one sequence
      // c = it * n_tiles + j, tile c lives in S / P buffer c % NB
      int n_items = 0;
      for (int it = 0; it * item_stride + static_cast<int>(blockIdx.x) * NWG + w < p.total_items; ++it) ++n_items;
      const int total_c = n_items * n_tiles;
      auto ring_pos = [&](int c) -> uint32_t {  // the producer's order: rounds, then key tiles, then active warpgroups
        const int it = c / n_tiles, j = c - it * n_tiles;
        const int item0 = it * item_stride + static_cast<int>(blockIdx.x) * NWG;
        const uint32_t nact = (item0 + NWG - 1 < p.total_items) ? NWG : 1;
        return static_cast<uint32_t>(it) * n_tiles * NWG + static_cast<uint32_t>(j) * nact + w;
      };
      auto issue_s = [&](int c) {
        const int it = c / n_tiles, j = c - it * n_tiles;
        if (j == 0) mbar_wait_a(qfull, it & 1);  // this item's Q tile has landed
        const uint32_t pos = ring_pos(c);
        const uint32_t s = pos % STAGES, ph = (pos / STAGES) & 1u;
        mbar_wait_a(kvf0 + 8u * s, ph);
        tc_fence_after();
        const uint32_t k_lo = (kv_lo0 + s * (Cfg::kStageBytes >> 4)) | (1u << 16);
        const uint32_t tmem_s = tmem_s0 + static_cast<uint32_t>(c % NB) * KV;
        for (int ks = 0; ks < p.ksteps_qk; ++ks) {
          const uint32_t off = static_cast<uint32_t>(ks >> 2) * ((128u * 128u) >> 4) + static_cast<uint32_t>(ks & 3) * 2u;
          const uint32_t koff = static_cast<uint32_t>(ks >> 2) * ((KV * 128u) >> 4) + static_cast<uint32_t>(ks & 3) * 2u;
          umma_bf16_ss_p(el, tmem_s, desc_hi | (q_lo + off), desc_hi | (k_lo + koff), idesc_s, ks != 0 ? 1u : 0u);
        }
        tc_commit_p(el, sfull0 + 8u * static_cast<uint32_t>(c % NB));
        if (j == n_tiles - 1) tc_commit_p(el, qfree);  // the item's Q tile is no longer needed once this completes
      };
      for (int c = 0; c < NB && c < total_c; ++c) issue_s(c);
      for (int c = 0; c < total_c; ++c) {
        const int it = c / n_tiles, j = c - it * n_tiles;
        const uint32_t b = static_cast<uint32_t>(c % NB), bph = static_cast<uint32_t>(c / NB) & 1u;
        mbar_wait_a(pfull0 + 8u * b, bph);
        // P(c) is ready and S buffer b is free: P V of tile c first, then S(c + NB). The tensor core executes one
        // thread's MMAs in issue order, so the completion of S(c + NB) - which the softmax waits for anyway before it
        // touches tile c + NB - also tells it that P V(c) has finished reading P buffer b: the softmax needs no
        // separate wait for pv_done before it overwrites the buffer (one barrier round trip less per key tile). S(c + NB)
        // starts ~150 cycles later than it could; it is not needed for another ~2 000.
        if (j == 0) mbar_wait_a(ofree, (it & 1) ^ 1);  // the previous item's O has been read out
        tc_fence_after();
        const uint32_t s = ring_pos(c) % STAGES;
        const uint32_t v_lo = kv_lo0 + s * (Cfg::kStageBytes >> 4) + (Cfg::kKBytes >> 4);
        const uint32_t p_lo = p_lo0 + b * (Cfg::kPBytes >> 4);
        const int kvalid = p.Nk - j * KV;  // keys of this tile that exist; P is zero beyond them
        const int ksteps = kvalid >= KV ? KV / 16 : (kvalid + 15) >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t adesc = desc_hi | (p_lo + static_cast<uint32_t>(ks >> 2) * ((128u * 128u) >> 4) +
                                            static_cast<uint32_t>(ks & 3) * 2u);
          const uint64_t bdesc = desc_hi_mn | (v_lo + static_cast<uint32_t>(ks) * ((16u * 128u) >> 4));
          const uint32_t accf = (j | ks) != 0 ? 1u : 0u;
          umma_bf16_ss_p(el, tmem_o, adesc, bdesc, idesc_o, accf);
          // row sums l += P 1: the same P against a tile of ones, 16 columns right after O - the softmax warps
          // never add the probabilities up themselves, and l is exactly the sum of the bf16 P that P V uses
          umma_bf16_ss_p(el, tmem_l, adesc, ones_desc, idesc_l, accf);
        }
        tc_commit_p(el, kve0 + 8u * s);
        tc_commit_p(el, pvdone0 + 8u * b);
        if (c + NB < total_c) issue_s(c + NB);
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax + output (warpgroup w) ----------------
    const int w = (warp - 2) >> 2;  // warps 2..5 -> warpgroup 0, 6..9 -> warpgroup 1
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tmem_S0 = tmem_base + w * (NB * KV + Cfg::kONCols);
    const uint32_t tmem_O = tmem_S0 + NB * KV;
    const uint32_t prow_s0 = smem_u32(sP + w * NB * Cfg::kPBytes + r * 128);  // this row of P buffer 0 (shared address)
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t cnt = 0;
    for (int round = 0;; ++round) {
      const int item = round * item_stride + static_cast<int>(blockIdx.x) * NWG + w;
      if (item >= p.total_items) break;
      const int qt = item % p.n_qt;
      const int ghh = item / p.n_qt;
      const int g = ghh / p.heads, head = ghh % p.heads;
      int row = qt * 128 + r;
      bool valid = row < p.R;
      const uint8_t* mrow = nullptr;
      int t_mod = 0;  // temporal: this row's pixel within the block
      if constexpr (TEMPORAL) {
        t_mod = r % p.tP;
        const int f = r / p.tP, n = qt * p.tP + t_mod;
        valid = (f < p.tF) && (n < p.tN);
        row = f * p.tN + n;  // row inside clip g (p.R = tF * tN)
      } else if (p.mask != nullptr && valid) {
        mrow = p.mask + ((static_cast<int64_t>(g) * p.R + row) / p.mask_rows) * p.mask_ld;
      }
      // Online softmax with a LAZY exponent offset: tile 0 is anchored on its exact row maximum (pass 1); later tiles
      // are exponentiated in ONE pass against the offset already in use while their own maximum is tracked, and the
      // offset moves (O and the row sum are rescaled, the tile redone) only when a row's maximum outgrows it by more
      // than 2^8 - probabilities then stay <= 256, exact in bf16 range and harmless in the fp32 accumulators, and the
      // final O / l normalisation is unchanged. This removes the second read of S and nearly every O rescale.
      float m_used = 0.f;
      bool anchored = false;  // m_used comes from a real maximum of this row (false while every key was masked)
      const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && warp == 2 && lane == 0;
      long long t_a = 0, t_b = 0;
      for (int j = 0; j < n_tiles; ++j, ++cnt) {
        if (tr) t_a = clock64();
        const uint32_t b = cnt % NB, bph = (cnt / NB) & 1u;  // S / P buffer of this tile and its barrier phase
        const uint32_t tmem_S = tmem_S0 + b * KV;
        const uint32_t prow_s = prow_s0 + b * Cfg::kPBytes;
        mbar_wait(&s_full[w * NB + b], bph);
        tc_fence_after();
        if (tr) { t_b = clock64(); p.trace[0] += t_b - t_a; t_a = t_b; }
        const int key0 = j * KV;
        // interior tiles without a mask skip every per-key predicate
        const bool plain = !TEMPORAL && (mrow == nullptr) && (key0 + KV <= p.Nk);
        // columns of this tile that can hold keys, rounded up to the 16-key granularity of the P V product and to
        // the 32-column chunks processed here (everything past Nk is written as zero probability)
        const int cols = (p.Nk - key0 >= KV) ? KV : (((p.Nk - key0 + 15) >> 4) << 4);
        auto chunk_mask = [&](int c) -> uint32_t {  // bit i: key key0 + c + i may be attended (non-plain tiles)
          uint32_t mbits = 0xffffffffu;
          if constexpr (TEMPORAL) {
            mbits = 0;  // same pixel <=> same residue modulo tP
            int rem = c % p.tP;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (rem == t_mod) mbits |= (1u << i);
              if (++rem == p.tP) rem = 0;
            }
          } else if (mrow != nullptr) {
            mbits = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (key0 + c + i < p.Nk && mrow[key0 + c + i] != 0) mbits |= (1u << i);
          }
          const int left = p.Nk - key0 - c;  // keys past Nk do not exist
          if (left < 32) mbits &= (left <= 0) ? 0u : ((1u << left) - 1u);
          return mbits;
        };
        auto row_max = [&]() -> float {  // pass 1 (tile 0 only)
          float m = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < cols; c += 32) {
            uint32_t sv[32];
            tmem_ld_x32(tmem_S + lane_base + c, sv);
            tmem_ld_wait();
            if (plain) {
#pragma unroll
              for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(sv[i]));
            } else {
              const uint32_t mbits = chunk_mask(c);
#pragma unroll
              for (int i = 0; i < 32; ++i) m = fmaxf(m, ((mbits >> i) & 1u) ? __uint_as_float(sv[i]) : -INFINITY);
            }
          }
          return m;
        };
        // exponentiate against offset m_off, write the bf16 P tile (K-major, 128B swizzle), return the raw row max
        // (the row sums come out of the tensor core: see the ones MMA next to P V)
        // one 32-column chunk: exponentiate, pack to bf16, store into the swizzled P tile
        auto exp_chunk = [&](const uint32_t (&sv)[32], int c, float m_off, float& mt) {
          float pv[32];
          if (plain && (POLY % 2 == 0)) {
            const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), off2 = pack_f32x2(-m_off, -m_off);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
              mt = fmaxf(mt, fmaxf(s0, s1));
              const uint64_t x2 = fma_f32x2(pack_f32x2(s0, s1), sc2, off2);
              if ((i & 7) < POLY) {  // balance the MUFU and FMA pipes
                poly_exp2_x2(x2, pv[i], pv[i + 1]);
              } else {
                float x0, x1;
                unpack_f32x2(x2, x0, x1);
                pv[i] = fast_exp2(x0);
                pv[i + 1] = fast_exp2(x1);
              }
            }
          } else if (plain) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              mt = fmaxf(mt, __uint_as_float(sv[i]));
              const float x = fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_off);
              pv[i] = ((i & 7) < POLY) ? poly_exp2(x) : fast_exp2(x);  // balance the MUFU and FMA pipes
            }
          } else {
            const uint32_t mbits = chunk_mask(c);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const bool keep = (mbits >> i) & 1u;
              mt = fmaxf(mt, keep ? __uint_as_float(sv[i]) : -INFINITY);
              pv[i] = keep ? fast_exp2(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_off)) : 0.f;
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(pv[2 * i], pv[2 * i + 1]);
          const uint32_t patom = prow_s + static_cast<uint32_t>(c >> 6) * (128u * 128u);
          const uint32_t chunk0 = static_cast<uint32_t>((c & 63) >> 3);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint32_t phys = (chunk0 + v) ^ sw;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(patom + phys * 16u), "r"(pk[4 * v]),
                         "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
          }
        };
        // exponentiate against offset m_off, write the bf16 P tile (K-major, 128B swizzle), return the raw row max
        // (the row sums come out of the tensor core: see the ones MMA next to P V)
        auto exp_pass = [&](float m_off, float& mt) {
          mt = -INFINITY;
          if constexpr (KV == 64 && !TEMPORAL && PIPE) {
            // two chunks: the second TMEM load is in flight while the first chunk is exponentiated
            uint32_t s0[32], s1[32];
            tmem_ld_x32(tmem_S + lane_base, s0);
            tmem_ld_wait();
            const bool two = cols > 32;
            if (two) tmem_ld_x32(tmem_S + lane_base + 32, s1);
            exp_chunk(s0, 0, m_off, mt);
            if (two) {
              tmem_ld_wait();
              exp_chunk(s1, 32, m_off, mt);
            }
          } else {
#pragma unroll 1
            for (int c = 0; c < cols; c += 32) {
              uint32_t sv[32];
              tmem_ld_x32(tmem_S + lane_base + c, sv);
              tmem_ld_wait();
              exp_chunk(sv, c, m_off, mt);
            }
          }
        };
        // (P buffer b was last read by P V of tile cnt - NB, which the issuer orders BEFORE S(cnt): s_full above
        //  already implies it has finished)
        if (j == 0) {
          const float mt0 = row_max();
          anchored = mt0 > -INFINITY;
          m_used = anchored ? mt0 * p.scale_log2 : 0.f;  // scale > 0: max commutes with the scaling
          if (tr) { t_b = clock64(); p.trace[1] += t_b - t_a; t_a = t_b; }
        }
        if (tr) { t_b = clock64(); p.trace[2] += t_b - t_a; t_a = t_b; }
        float mt;
        exp_pass(m_used, mt);
        if (j > 0) {
          const float mts = mt * p.scale_log2;
          const bool re = (mt > -INFINITY) && (!anchored || mts > m_used + 8.f);
          if (__any_sync(0xffffffffu, re)) {  // rare: some row of this warp outgrew its offset - re-anchor and redo
            if (NB > 1) {  // O must hold every product up to tile cnt - 1 before it is rescaled
              mbar_wait(&pv_done[w * NB + (cnt - 1) % NB], ((cnt - 1) / NB) & 1u);
              tc_fence_after();
            }
            const float m_new = re ? mts : m_used;
            const float alpha = (re && anchored) ? fast_exp2(m_used - m_new) : 1.f;  // unanchored rows hold O = 0
#pragma unroll 1
            for (int c = 0; c < p.dN + 16; c += 16) {  // O and the row-sum columns behind it
              uint32_t ov[16];
              tmem_ld_x16(tmem_O + lane_base + c, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st_x16(tmem_O + lane_base + c, ov);
            }
            tmem_st_wait();
            m_used = m_new;
            anchored = anchored || re;
            exp_pass(m_used, mt);
          }
        }
        if (tr) { t_b = clock64(); p.trace[3] += t_b - t_a; t_a = t_b; }
        tc_fence_before();
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core
        mbar_arrive(&p_full[w * NB + b]);  // also: S consumed (a later Q K^T may overwrite this buffer)
        if (tr) { t_b = clock64(); p.trace[4] += t_b - t_a; p.trace[5] += 1; }
      }
      // ---------------- output ----------------
      mbar_wait(&pv_done[w * NB + (cnt - 1) % NB], ((cnt - 1) / NB) & 1u);
      tc_fence_after();
      float l_run;
      {
        uint32_t lv[16];
        tmem_ld_x16(tmem_O + lane_base + p.dN, lv);
        tmem_ld_wait();
        l_run = __uint_as_float(lv[0]);
      }
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(g) * p.R + row) * p.ldo + head * p.d;
#pragma unroll 1
      for (int c = 0; c < p.dN; c += 16) {
        uint32_t ov[16];
        tmem_ld_x16(tmem_O + lane_base + c, ov);
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h * 8 < p.d) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[h * 8 + 0]) * inv, __uint_as_float(ov[h * 8 + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[h * 8 + 2]) * inv, __uint_as_float(ov[h * 8 + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[h * 8 + 4]) * inv, __uint_as_float(ov[h * 8 + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[h * 8 + 6]) * inv, __uint_as_float(ov[h * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + h * 8) = u;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[w]);  // the next item's first P V may overwrite O
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int DKA, int KV, int NWG, int STAGES, int NB, bool TEMPORAL, int POLY = kPolyOf8, bool PIPE = true>
static int launch_attn(const AttnKParams& kp, cudaStream_t stream) {
  using Cfg = AttnCfg<DKA, KV, NWG, STAGES, NB>;
  static_assert(Cfg::kSmemBytes <= 232448, "attention configuration exceeds the shared memory of an SM");
  static bool configured[kMaxDevices] = {false};  // the opt-in shared-memory size is a per-device attribute
  const int dev = current_device();
  if (!configured[dev]) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<DKA, KV, NWG, STAGES, NB, TEMPORAL, POLY, PIPE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured[dev] = true;
  }
  const int sms = device_sms();
  ASVA_REQUIRE(sms > 0, "asva_attention: cannot query the device");
  int grid = (kp.total_items + NWG - 1) / NWG;
  if (grid > sms) grid = sms;
  ASVA_CUDA_OK(launch_k(attn_tc_kernel<DKA, KV, NWG, STAGES, NB, TEMPORAL, POLY, PIPE>, dim3(grid), dim3(Cfg::kThreads),
                        Cfg::kSmemBytes, stream, 1, kp));
  return 0;
}

}  // namespace asva

extern "C" int asva_attention(const asva_attn_desc* d, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(d != nullptr && d->q && d->kv && d->out, "asva_attention: null operand");
  ASVA_REQUIRE(d->d >= 8 && d->d % 8 == 0 && d->d <= 192, "asva_attention: head dim %d unsupported", d->d);
  ASVA_REQUIRE(d->dpad % 64 == 0 && d->dpad >= d->d && d->dpad <= 192, "asva_attention: dpad %d invalid", d->dpad);
  ASVA_REQUIRE(d->G >= 1 && d->heads >= 1 && d->R >= 1 && d->Nk >= 1, "asva_attention: empty problem");
  ASVA_REQUIRE(d->scale > 0.f, "asva_attention: scale must be positive");
  ASVA_REQUIRE(d->ldkv % 8 == 0 && d->ldo % 8 == 0 && d->k_col0 % 8 == 0 && d->v_col0 % 8 == 0,
               "asva_attention: ldkv/ldo/k_col0/v_col0 must be multiples of 8");
  ASVA_REQUIRE(d->ldq % 8 == 0 && d->ldq >= (int64_t)d->heads * d->d, "asva_attention: ldq=%lld invalid",
               (long long)d->ldq);
  ASVA_REQUIRE(d->mask == nullptr || d->mask_rows >= 1, "asva_attention: mask_rows must be >= 1");
  ASVA_REQUIRE(d->kv_rows_per_group >= d->Nk, "asva_attention: kv_rows_per_group < Nk");
  ASVA_REQUIRE(d->form >= 0 && d->form <= 2, "asva_attention: form %d", d->form);
  if (d->form == 2) {  // the warp-MMA kernel for small key sets (attn_mma.cu), by request only: measured slower than the
                       // tcgen05 kernel on every workload shape (profiles/r2_attn_mma.md)
    const int rc = attention_small_keys(d, stream);
    ASVA_REQUIRE(rc != 1, "asva_attention: the warp-MMA form does not serve Nk=%d, d=%d", d->Nk, d->d);
    return rc;
  }

  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.mask = d->mask;
  kp.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  kp.ldo = d->ldo;
  kp.mask_ld = d->mask_ld;
  kp.R = d->R;
  kp.Nk = d->Nk;
  kp.d = d->d;
  kp.dN = ((d->d + 15) / 16) * 16;
  kp.heads = d->heads;
  kp.k_col0 = d->k_col0;
  kp.v_col0 = d->v_col0;
  kp.mask_rows = d->mask_rows > 0 ? d->mask_rows : 1;
  kp.ksteps_qk = (d->d + 15) / 16;
  kp.nvb = (kp.dN + 63) / 64;
  kp.scale_log2 = d->scale * 1.4426950408889634f;
  kp.n_qt = (d->R + 127) / 128;
  const int64_t items = (int64_t)kp.n_qt * d->heads * d->G;
  ASVA_REQUIRE(items < (1ll << 30), "asva_attention: too many query tiles");
  kp.total_items = (int)items;
  const int dka = d->dpad / 64;
  const int kv = 64;  // key tile of every non-temporal configuration below
  {
    uint64_t dims[3] = {(uint64_t)d->d, (uint64_t)d->heads, (uint64_t)d->G * (uint64_t)d->R};
    uint64_t strides[2] = {(uint64_t)d->d * 2u, (uint64_t)d->ldq * 2u};
    uint32_t box[3] = {64u, 1u, 128u};
    uint32_t el[3] = {1u, 1u, 1u};
    int rc = make_tmap_bf16(&kp.tmQ, d->q, 3, dims, strides, box, el);
    if (rc != 0) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->ldkv, (uint64_t)d->Nk, (uint64_t)d->G};
    uint64_t strides[2] = {(uint64_t)d->ldkv * 2u, (uint64_t)d->kv_rows_per_group * (uint64_t)d->ldkv * 2u};
    uint32_t box[3] = {64u, (uint32_t)kv, 1u};
    uint32_t el[3] = {1u, 1u, 1u};
    int rc = make_tmap_bf16(&kp.tmKV, d->kv, 3, dims, strides, box, el);
    if (rc != 0) return rc;
  }
#ifdef ASVA_DEBUG_SWITCHES  // per-phase clock64 trace and A/B variants: experiment builds only
  static long long* trace_buf = nullptr;
  static int trace_on = -1;
  if (trace_on < 0) {
    const char* e = getenv("ASVA_ATTN_TRACE");
    trace_on = (e != nullptr && e[0] == '1') ? 1 : 0;
    if (trace_on) cudaMalloc(&trace_buf, 8 * sizeof(long long));
  }
  if (trace_on) {
    cudaMemsetAsync(trace_buf, 0, 8 * sizeof(long long), stream);
    kp.trace = trace_buf;
  }
#endif
  int rc;
  switch (dka) {
    case 1: {
#ifdef ASVA_DEBUG_SWITCHES
      static int poly = -1;  // ASVA_ATTN_POLY=0..4: exponentials per 8 computed on the FMA pipes (experiments)
      static bool pipe = true;  // ASVA_ATTN_PIPE=0: serial TMEM loads (the previous form, for A/B timing)
      if (poly < 0) {
        const char* e = getenv("ASVA_ATTN_POLY");
        poly = (e != nullptr && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : kPolyOf8;
        e = getenv("ASVA_ATTN_PIPE");
        pipe = !(e != nullptr && e[0] == '0');
      }
#else
      constexpr bool pipe = true;
#endif
      // overlapping the second chunk's TMEM load pays once an item has several key tiles (self-attention: -3 %); the
      // one- and two-tile items of the cross-attentions run ~4 % faster with the leaner serial form (measured)
      if (!pipe || d->Nk < 256) {
        rc = launch_attn<1, 64, 2, 8, 2, false, kPolyOf8, false>(kp, stream);
        break;
      }
#ifdef ASVA_DEBUG_SWITCHES
      switch (poly) {
        case 0: rc = launch_attn<1, 64, 2, 8, 2, false, 0>(kp, stream); break;
        case 1: rc = launch_attn<1, 64, 2, 8, 2, false, 1>(kp, stream); break;
        case 4: rc = launch_attn<1, 64, 2, 8, 2, false, 4>(kp, stream); break;
        case 3: rc = launch_attn<1, 64, 2, 8, 2, false, 3>(kp, stream); break;
        default: rc = launch_attn<1, 64, 2, 8, 2, false, kPolyOf8>(kp, stream); break;
      }
#else
      rc = launch_attn<1, 64, 2, 8, 2, false, kPolyOf8>(kp, stream);
#endif
      break;
    }
    case 2: rc = launch_attn<2, 64, 2, 4, 1, false>(kp, stream); break;
    default: rc = launch_attn<3, 64, 1, 3, 1, false>(kp, stream); break;
  }
#ifdef ASVA_DEBUG_SWITCHES
  if (trace_on && rc == 0) {
    long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    if (h[5] > 0)
      printf("[asva attn trace] tiles %lld: wait S %.0f | pass1 %.0f | wait PV + rescale %.0f | pass2 %.0f | arrive %.0f cycles per tile\n",
             h[5], (double)h[0] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5], (double)h[4] / h[5]);
  }
#endif
  return rc;
}

static int asva_temporal_attention_tc(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                                      int32_t d, float scale, asva_stream_t stream_);

static int temporal_checks(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads, int32_t d,
                           float scale) {
  using namespace asva;
  ASVA_REQUIRE(qkv && out, "asva_temporal_attention: null operand");
  ASVA_REQUIRE(d % 8 == 0 && d >= 8 && d <= 192, "asva_temporal_attention: head dim %d unsupported", d);
  ASVA_REQUIRE(B >= 1 && N >= 1 && heads >= 1 && F >= 1 && F <= 64, "asva_temporal_attention: bad shape (F=%d)", F);
  ASVA_REQUIRE(scale > 0.f, "asva_temporal_attention: scale must be positive");
  return 0;
}

extern "C" int asva_temporal_attention(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                                       int32_t d, float scale, asva_stream_t stream_) {
  return asva_temporal_attention_form(qkv, out, B, F, N, heads, d, scale, 0, stream_);
}

extern "C" int asva_temporal_attention_form(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                                            int32_t d, float scale, int32_t form, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = temporal_checks(qkv, out, B, F, N, heads, d, scale)) return rc;
  ASVA_REQUIRE(form >= 0 && form <= 3, "asva_temporal_attention_form: form %d", form);
  if (form == 0 || form == 3) {  // warp-MMA form (misc.cu) for every shape it serves
    const int rc = temporal_attention_mma(qkv, out, B, F, N, heads, d, scale, stream);
    if (rc != 1) return rc;
    ASVA_REQUIRE(form == 0, "asva_temporal_attention_form: warp-MMA form does not serve F=%d, C=%d", F, heads * d);
  }
  if (form == 2) {
    const int rc = temporal_attention_rows(qkv, out, B, F, N, heads, d, scale, stream, true);
    ASVA_REQUIRE(rc != 1, "asva_temporal_attention_form: thread-per-query form does not serve F=%d, C=%d", F, heads * d);
    return rc;
  }
  return asva_temporal_attention_tc(qkv, out, B, F, N, heads, d, scale, stream_);
}

static int asva_temporal_attention_tc(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                                          int32_t d, float scale, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(qkv && out, "asva_temporal_attention: null operand");
  ASVA_REQUIRE(d % 8 == 0 && d >= 8 && d <= 192, "asva_temporal_attention: head dim %d unsupported", d);
  ASVA_REQUIRE(B >= 1 && N >= 1 && heads >= 1 && F >= 1 && F <= 64, "asva_temporal_attention: bad shape (F=%d)", F);
  ASVA_REQUIRE(scale > 0.f, "asva_temporal_attention: scale must be positive");
  const int C = heads * d;
  const int dka = (d + 63) / 64;
  const int kv = (dka == 1) ? 128 : 64;
  int P = kv / F;  // pixels per tile: tP * tF rows are queries and keys at once
  if (P > N) P = N;
  ASVA_REQUIRE(P >= 1, "asva_temporal_attention: F=%d exceeds the key tile of %d", F, kv);

  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.out = reinterpret_cast<__nv_bfloat16*>(out);
  kp.ldo = C;
  kp.R = F * N;
  kp.Nk = P * F;
  kp.d = d;
  kp.dN = ((d + 15) / 16) * 16;
  kp.heads = heads;
  kp.mask_rows = 1;
  kp.ksteps_qk = (d + 15) / 16;
  kp.nvb = (kp.dN + 63) / 64;
  kp.scale_log2 = scale * 1.4426950408889634f;
  kp.tP = P;
  kp.tF = F;
  kp.tN = N;
  kp.tPB = (N + P - 1) / P;
  kp.n_qt = kp.tPB;
  const int64_t items = (int64_t)kp.tPB * heads * B;
  ASVA_REQUIRE(items < (1ll << 30), "asva_temporal_attention: too many tiles");
  kp.total_items = (int)items;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv);
  for (int part = 0; part < 3; ++part) {
    uint64_t dims[5] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)F, (uint64_t)B};
    uint64_t strides[4] = {(uint64_t)d * 2u, (uint64_t)3 * C * 2u, (uint64_t)N * 3 * C * 2u,
                           (uint64_t)F * N * 3 * C * 2u};
    uint32_t box[5] = {64u, 1u, (uint32_t)P, (uint32_t)F, 1u};
    uint32_t el[5] = {1u, 1u, 1u, 1u, 1u};
    CUtensorMap* tm = part == 0 ? &kp.tmQ : (part == 1 ? &kp.tmKV : &kp.tmV);
    int rc = make_tmap_bf16(tm, base + (int64_t)part * C, 5, dims, strides, box, el);
    if (rc != 0) return rc;
  }
  switch (dka) {
    case 1: return launch_attn<1, 128, 2, 4, 1, true>(kp, stream);
    case 2: return launch_attn<2, 64, 2, 4, 1, true>(kp, stream);
    default: return launch_attn<3, 64, 1, 3, 1, true>(kp, stream);
  }
}
