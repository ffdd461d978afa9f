// tcgen05 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] * W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Persistent kernel, one CTA per SM walking 128 x BN output tiles round-robin; every global access is a TMA
// transfer, so no warp ever waits on a global load. Warp roles (384 threads):
//   warps 0, 3   TMA producers (one thread each). The smem ring holds stages of TWO 64-wide K blocks; producer j fills
//                block j of every stage: one 4-D box load of A (table driven: implicit-GEMM conv taps, temporal
//                taps, concat sources) and one 2-D box load of W, 128B-swizzled. Two threads and two K blocks per
//                barrier round trip because a single thread's wait -> arm -> issue chain costs ~520 cycles per K
//                block (measured, tools/ubench) - more than the MMA time of a 128 x 160 x 64 block (320 cycles)
//   warp 1       TMEM allocator + single-thread tcgen05.mma issuer (8 x K=16 per stage, one barrier wait and one
//                commit per stage, the next stage's barrier probed before the MMAs are issued); the accumulator is
//                double buffered in TMEM (2 x BN columns), so tile t+1's main loop runs under tile t's epilogue
//   warps 4..11  epilogue, two groups of 4 warps (warp % 4 = TMEM lane quadrant, thread = output row); the groups
//                take alternate 32-column panels: tcgen05.ld -> + bias / per-row-block addend / residual panels
//                (from smem) [or GEGLU] -> swizzled staging slot -> one elected thread issues the TMA store
//                (tails are clipped by the tensor map). Staging slots are double buffered per group. Residual
//                operands arrive by TMA in a private ring per group (64B-swizzled [128 rows][32 bf16] slots) that one
//                lane of the group keeps filled D panels ahead of the group's own position, across tile boundaries.
// Split-K (small-M, long-K problems): tiles are (split, m, n) triples writing fp32 partial tiles to a scratch
// buffer through the same TMA-store path; splitk_finalize_kernel then reduces them and applies the epilogue.
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>
#include <string.h>

namespace asva {

constexpr int kMaxStages = 8;
constexpr int kMaxResSlots = 4;       // residual-panel slots per epilogue group (runtime: n_res_slots)
constexpr int kResSlotBytes = 8192;   // 128 rows x 32 bf16
constexpr int kGemmThreads = 384;
constexpr int kSmemLimit = 232448;    // 227 KB opt-in maximum per CTA
constexpr int kBarBytes = 512;
constexpr int kDefaultEpi = 1;          // epilogue form when neither the descriptor nor ASVA_GEMM_EPI chooses

// Experiment switches (ASVA_GEMM_DBG: skip TMA loads / MMAs for timing; ASVA_GEMM_BN / SPLIT / CG / EPI / STAGES / PF:
// plan overrides) exist only in builds made with -DASVA_DEBUG_SWITCHES (ASVA_NVCC_EXTRA, tools/gemm_dbg_sweep.py).
// The default library never reads the environment on the launch path and has no work-skipping branch.
#ifdef ASVA_DEBUG_SWITCHES
#define ASVA_DBG(p) ((p).dbg)
// ASVA_GEMM_TRACE=1: CTA 0 stamps %globaltimer at the marked points into a [warp][event] table (see asva_gemm)
#define ASVA_TR(p, w, e)                                                                      \
  do {                                                                                        \
    if ((p).trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (e) < 64) {     \
      unsigned long long _t;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                  \
      (p).trace[(w) * 64 + (e)] = static_cast<long long>(_t);                                 \
    }                                                                                         \
  } while (0)
#else
#define ASVA_DBG(p) 0
#define ASVA_TR(p, w, e) do { } while (0)
#endif

struct SegK {
  int32_t src, c0, off1, off2, off3, num_kb, wk, wk_first, fix2;
};

struct GemmKParams {
  CUtensorMap tmA0, tmA1, tmW, tmR0, tmR1, tmO;
  CUtensorMap tmO2;  // epi == 3, bf16: the 32-column tail panel of a tile whose width is not a multiple of 64
  SegK seg[ASVA_GEMM_MAX_SEG];
  int32_t box[3], trav[3], out_dims[3], tiles[3];
  int32_t rows_per_tile, N, n_out, num_kb, n_tiles_n, mn_tiles, total_tiles, split_k, kb_per_split;
  int32_t n_stages, n_res, n_res_slots, out_fp32, dbg;
  long long* trace;  // debug builds only (ASVA_GEMM_TRACE)
  int32_t sub_rows;  // epi == 3: rows of a warp's slice of the tile (32, or the whole tile when it has fewer)
  int32_t pf_blocks, w_kblocks;  // W prefetch: blocks per CTA (0 = off), 64-column blocks of W that exist  // dbg (ASVA_GEMM_DBG): 1 = skip TMA loads, 2 = skip MMAs (timing only)
  const float* bias;
  const float* add_ptr;
  int64_t add_ld;
  int32_t add_div;
  int32_t epi;  // 1 = panel epilogue (TMA residual ring, TMA stores), 2 = per-warp epilogue (direct loads / stores),
                // 3 = warp-private TMA epilogue (every warp moves its own 32-row slice of a panel with its own TMA ops)
                // 4 = cluster split-K: the csplit CTAs of a cluster hold the K partials of ONE tile in TMEM, exchange
                //     column slices through distributed shared memory and each finishes one slice (no workspace, no
                //     reduce launch)
  int32_t csplit;          // epi == 4: cluster size = K splits of a tile (2 / 4 / 8)
  int32_t csplit_f32;      // epi == 4: the output is fp32 (out_fp32 stays 0 there: no staging area is laid out)
  __nv_bfloat16* out_ptr;  // epi == 2 and 4 (direct stores; fp32 outputs through the same pointer)
  int64_t ldo;
  const __nv_bfloat16* res_ptr[2];
  int64_t res_ld[2];
  // row statistics / LayerNorm fold (asva_gemm_desc.stats_out, .ln_*)
  float* stats_out;       // [N/32][stats_rows][2]: (sum, sum of squares) of every stored row per 32-column slot
  int64_t stats_rows;
  const float* ln_stats;  // [ln_slots][ln_stat_rows][2]
  const float* ln_wsum;   // [N]
  int64_t ln_stat_rows;
  int32_t ln_slots, ln_grp_rows, ln_grp_stride;
  float ln_inv_c, ln_eps;
};

struct TileCoord {
  int n0, o1, o2, o3, split, kb0, kb1;
};

// CG = 1: `tile` indexes (split, m tile, n tile). CG = 2 (CTA pair): it indexes (split, PAIR of adjacent m tiles,
// n tile) and the CTA of rank r takes m tile 2*pair + r (an m tile past the end reads zero fill / stores nothing).
template <int BN, int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile, int rank, int srank = -1) {
  TileCoord t;
  // srank >= 0 (cluster split-K): `tile` indexes (m tile, n tile) and the CTA's rank in its cluster is the K split
  t.split = srank >= 0 ? srank : tile / p.mn_tiles;
  const int rem = srank >= 0 ? tile : tile - t.split * p.mn_tiles;
  const int nt = rem % p.n_tiles_n;
  const int mt = (rem / p.n_tiles_n) * CG + rank;
  t.n0 = nt * BN;
  t.o1 = (mt % p.tiles[0]) * p.box[0];
  t.o2 = ((mt / p.tiles[0]) % p.tiles[1]) * p.box[1];
  t.o3 = (mt / (p.tiles[0] * p.tiles[1])) * p.box[2];
  t.kb0 = t.split * p.kb_per_split;
  t.kb1 = min(p.num_kb, t.kb0 + p.kb_per_split);
  return t;
}

template <int BN, bool GEGLU>
__device__ __forceinline__ int tile_panels(const GemmKParams& p, int n0) {
  constexpr int kOutCols = GEGLU ? 64 : BN;
  const int n0_out = GEGLU ? (n0 >> 1) : n0;
  const int left = p.n_out - n0_out;
  return ((left < kOutCols ? left : kOutCols) + 31) >> 5;
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Global output row of tile row (r1, r2, r3) (-1: the row lies outside the tile or the tensor).
__device__ __forceinline__ int64_t tile_out_row(const GemmKParams& p, const TileCoord& tc, int r, int r1, int r2, int r3) {
  const int a1 = tc.o1 + r1, a2 = tc.o2 + r2, a3 = tc.o3 + r3;
  const bool valid = (r < p.rows_per_tile) && (a1 < p.out_dims[0]) && (a2 < p.out_dims[1]) && (a3 < p.out_dims[2]);
  return valid ? (static_cast<int64_t>(a3) * p.out_dims[1] + a2) * p.out_dims[0] + a1 : -1;
}

// LayerNorm fold: out = rs * acc + (bias + rm * wsum[col]) with rs = rstd, rm = -mean * rstd of the row the
// accumulator row was computed from; the sums were left per 32-column slot by the producing GEMM (or are 1 / 0).
__device__ __forceinline__ void row_fold(const GemmKParams& p, int64_t row, float& rs, float& rm) {
  rs = 1.f;
  rm = 0.f;
  if (p.ln_stats == nullptr || row < 0) return;
  int64_t sr = row;
  if (p.ln_grp_rows > 0) sr = (row / p.ln_grp_rows) * p.ln_grp_stride + row % p.ln_grp_rows;
  const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + sr;
  float sum = 0.f, sq = 0.f;
  for (int i = 0; i < p.ln_slots; ++i) {
    const float2 v = __ldg(st + i * p.ln_stat_rows);
    sum += v.x;
    sq += v.y;
  }
  const float mean = sum * p.ln_inv_c;
  const float var = fmaxf(sq * p.ln_inv_c - mean * mean, 0.f);
  rs = rsqrtf(var + p.ln_eps);
  rm = -mean * rs;
}

// b += rm * wsum[col0 .. col0+31] (the folded LayerNorm's mean term joins the bias registers)
__device__ __forceinline__ void fold_bias(const GemmKParams& p, float4 (&b)[8], int col0, float rm) {
  if (p.ln_wsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (col0 + j * 4 < p.N) {
      const float4 w = ldg4(p.ln_wsum + col0 + j * 4);
      b[j].x = fmaf(rm, w.x, b[j].x);
      b[j].y = fmaf(rm, w.y, b[j].y);
      b[j].z = fmaf(rm, w.z, b[j].z);
      b[j].w = fmaf(rm, w.w, b[j].w);
    }
  }
}

// The additive column terms of 4 adjacent outputs of one row: bias + row addend + (LayerNorm fold) rm * wsum.
__device__ __forceinline__ float4 col_terms(const GemmKParams& p, const float* addp, int col, float rm) {
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < p.N) {
    if (p.bias != nullptr) b = ldg4(p.bias + col);
    if (addp != nullptr) {
      const float4 a = ldg4(addp + col);
      b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
    }
    if (p.ln_wsum != nullptr) {
      const float4 w = ldg4(p.ln_wsum + col);
      b.x = fmaf(rm, w.x, b.x); b.y = fmaf(rm, w.y, b.y); b.z = fmaf(rm, w.z, b.z); b.w = fmaf(rm, w.w, b.w);
    }
  }
  return b;
}

// (sum, sum of squares) of the 32 fp32 values a thread is about to store for its row -> stats[slot][row]
__device__ __forceinline__ void emit_row_stats(const GemmKParams& p, const uint32_t (&v)[32], int col0, int64_t row) {
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = __uint_as_float(v[j]);
    s[j & 3] += x;
    q[j & 3] = fmaf(x, x, q[j & 3]);
  }
  if (row >= 0 && col0 < p.N)
    *reinterpret_cast<float2*>(p.stats_out + (static_cast<int64_t>(col0 >> 5) * p.stats_rows + row) * 2) =
        make_float2((s[0] + s[1]) + (s[2] + s[3]), (q[0] + q[1]) + (q[2] + q[3]));
}


// Per-warp epilogue (epi == 2; bf16 output, no GEGLU, no split-K). Every epilogue warp owns the 32 x 32 sub-panels of
// its TMEM lane quadrant in alternate panels and runs on its own - no group barriers, no TMA residual ring, no
// staging/TMA-store round trip: tcgen05.ld (thread = row) -> fp32 transpose through a private 4 KB shared slab
// (16-byte chunks XOR-swizzled by the row, conflict free both ways) -> 4 lanes per row add bias / row addend /
// residuals to their 8 columns and store 16 bytes each (a warp instruction covers 8 rows x 64 contiguous bytes, whole
// sectors). Residuals are plain global loads in the same layout, fetched one panel ahead - the first panel of a tile
// while the warp still waits for the tile's accumulator - so L2 latency is off the per-panel chain.
template <int BN, int CG>
__device__ __forceinline__ void epilogue_direct(const GemmKParams& p, int warp, int lane, int rank, int tile0,
                                                int tile_step, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                uint64_t* tmem_empty_bar, uint8_t* slabs) {
  const uint32_t g = static_cast<uint32_t>(warp - 4) >> 2;
  const int qd = warp & 3;
  const uint32_t slab = smem_u32(slabs) + static_cast<uint32_t>(warp - 4) * 4096u;
  const int cp = lane & 3;     // which 8-column chunk of a panel this lane finishes
  const int rsub = lane >> 2;  // its row inside each 8-row slab
  int q1[4], q2[4], q3[4];
  bool in_tile[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int tr = qd * 32 + it * 8 + rsub;
    q1[it] = tr % p.box[0];
    q2[it] = (tr / p.box[0]) % p.box[1];
    q3[it] = tr / (p.box[0] * p.box[1]);
    in_tile[it] = tr < p.rows_per_tile;
  }
  const uint32_t wr_base = slab + static_cast<uint32_t>(lane) * 128u;
  const uint32_t wr_x = static_cast<uint32_t>(lane & 7);
  const uint32_t rd0 = static_cast<uint32_t>(((2 * cp) ^ rsub) << 4), rd1 = static_cast<uint32_t>(((2 * cp + 1) ^ rsub) << 4);
  uint32_t pc = 0, t = 0;
  for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++t) {
    const TileCoord tc = decode_tile<BN, CG>(p, tile, rank);
    const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
    const int n_panels = tile_panels<BN, false>(p, tc.n0);
    int grow[4], arow[4];  // global output row of each of the lane's 4 rows (-1: does not exist), its addend row
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int a1 = tc.o1 + q1[it], a2 = tc.o2 + q2[it], a3 = tc.o3 + q3[it];
      const bool ok = in_tile[it] && (a1 < p.out_dims[0]) && (a2 < p.out_dims[1]) && (a3 < p.out_dims[2]);
      grow[it] = ok ? (a3 * p.out_dims[1] + a2) * p.out_dims[0] + a1 : -1;
      arow[it] = ok ? grow[it] / p.add_div : 0;
    }
    const int q_first = ((pc & 1u) == g) ? 0 : 1;
    int q_last = n_panels - 1;
    if (((pc + q_last) & 1u) != g) --q_last;
    uint4 cur[2][4];
    auto load_res = [&](int q, uint4 (&dst)[2][4]) {
      const int col = tc.n0 + q * 32 + cp * 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          dst[i][it] = make_uint4(0u, 0u, 0u, 0u);
          if (i < p.n_res && grow[it] >= 0 && col < p.N)
            dst[i][it] = *reinterpret_cast<const uint4*>(p.res_ptr[i] + static_cast<int64_t>(grow[it]) * p.res_ld[i] + col);
        }
      }
    };
    if (q_first <= q_last) load_res(q_first, cur);
    mbar_wait(&tmem_full_bar[acc], acc_ph);
    tc_fence_after();
    if (q_last < q_first) {  // no panel of this tile is ours: hand the accumulator back right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_pair_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(qd * 32) << 16);
#pragma unroll 1
    for (int q = q_first; q <= q_last; q += 2) {
      const int col = tc.n0 + q * 32 + cp * 8;
      const bool cok = col < p.N;
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (p.bias != nullptr && cok) {
        b0 = ldg4(p.bias + col);
        b1 = ldg4(p.bias + col + 4);
      }
      {
        uint32_t v[32];
        tmem_ld_x32(taddr + q * 32, v);
        tmem_ld_wait();
        if (q == q_last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_pair_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(wr_base + ((static_cast<uint32_t>(j) ^ wr_x) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      __syncwarp();
      uint4 nxt[2][4];
      const bool more = q + 2 <= q_last;
      if (more) load_res(q + 2, nxt);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const uint32_t ra = slab + static_cast<uint32_t>(it * 8 + rsub) * 128u;
        float4 x0 = ld_shared_f4(ra + rd0), x1 = ld_shared_f4(ra + rd1);
        x0.x += b0.x; x0.y += b0.y; x0.z += b0.z; x0.w += b0.w;
        x1.x += b1.x; x1.y += b1.y; x1.z += b1.z; x1.w += b1.w;
        const bool ok = grow[it] >= 0 && cok;
        if (p.add_ptr != nullptr && ok) {
          const float* ap = p.add_ptr + static_cast<int64_t>(arow[it]) * p.add_ld + col;
          const float4 a0 = ldg4(ap), a1 = ldg4(ap + 4);
          x0.x += a0.x; x0.y += a0.y; x0.z += a0.z; x0.w += a0.w;
          x1.x += a1.x; x1.y += a1.y; x1.z += a1.z; x1.w += a1.w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i < p.n_res) {
            const uint4 w = cur[i][it];
            const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z), f3 = unpack_bf16x2(w.w);
            x0.x += f0.x; x0.y += f0.y; x0.z += f1.x; x0.w += f1.y;
            x1.x += f2.x; x1.y += f2.y; x1.z += f3.x; x1.w += f3.y;
          }
        }
        if (ok) {
          uint4 w;
          w.x = pack_bf16x2(x0.x, x0.y);
          w.y = pack_bf16x2(x0.z, x0.w);
          w.z = pack_bf16x2(x1.x, x1.y);
          w.w = pack_bf16x2(x1.z, x1.w);
          *reinterpret_cast<uint4*>(p.out_ptr + static_cast<int64_t>(grow[it]) * p.ldo + col) = w;
        }
      }
      __syncwarp();  // every lane is done reading the slab before the next panel overwrites it
      if (more) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int it = 0; it < 4; ++it) cur[i][it] = nxt[i][it];
      }
    }
    pc += n_panels;
  }
}

// Warp-private TMA epilogue (epi == 3). Measured with a %globaltimer trace (profiles/r2_gemm_trace.md): a panel of the
// panel epilogue costs a warp ~1.5 us of which ~350 instructions are issued by only two warps per scheduler - bias loads
// that miss the (almost absent) L1, bf16 unpack + fp32 adds for the residual, one lane decoding tiles and issuing the
// residual TMA loads, 64-byte rows on both the residual and the output side. Every short-K launch runs at that pace.
// This form removes work from the chain instead of overlapping it:
//   * each of the eight epilogue warps owns the 32 rows of its TMEM lane quadrant: private staging, private TMA
//     stores through a tensor map whose row box is the quadrant's slice of the tile (host: sub_box) - no block barrier;
//   * bf16 outputs move 64-column panels (128-byte rows: whole lines in, whole lines out, half the per-panel fixed cost);
//   * residual panels arrive through a ring shared by both groups and are issued by warp 2, which has nothing else to
//     do, so no epilogue lane decodes tiles or issues loads; they are added in fp32 before the single bf16 rounding
//     (a packed bf16x2 add was 5 % faster but cost 8 % of the whole-UNet error budget: measured, not shipped);
//   * bias loads are issued before the wait for the TMEM load, so their L2 latency hides under it.
// GEGLU and fp32 outputs keep 32-column panels (epilogue_warp_tma_narrow).
template <int BN, int CG>
__device__ __forceinline__ int wide_chunks(const GemmKParams& p, int n0) {  // 32-column chunks of the tile that exist
  const int left = p.n_out - n0;
  return ((left < BN ? left : BN) + 31) >> 5;
}

// warp 2, one lane: residual panels in the order the epilogue consumes them (tile, panel, residual).
// The D slots are split into one private ring per epilogue group (group 0: ceil(D/2) slots, group 1: the rest): a slot is
// then always consumed by the same four warps, in order, and the parity of its barriers cannot alias. (With one ring
// shared by both groups a group could reach "slot s, phase p+1" before phase p had even landed - mbarrier parity waits
// tell adjacent phases apart, not phases two apart - read stale data, release the slot early and desynchronise the ring:
// seen as a hang when the loads came from DRAM in a different order, profiles/r2_gemm_trace.md.)
template <int BN, int CG>
__device__ __forceinline__ void residual_issuer_wide(const GemmKParams& p, int rank, int tile0, int tile_step,
                                                     uint64_t* res_full, uint64_t* res_empty, uint8_t* res_ring) {
  const uint32_t D = static_cast<uint32_t>(p.n_res_slots);
  const uint32_t dg[2] = {(D + 1u) >> 1, D >> 1}, base[2] = {0u, (D + 1u) >> 1};
  const uint32_t tx = static_cast<uint32_t>(p.rows_per_tile) * 128u;
  uint32_t pos[2] = {0u, 0u};  // entries issued into each group's ring
  uint32_t pc = 0;
  for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
    const TileCoord tc = decode_tile<BN, CG>(p, tile, rank);
    const int np = (wide_chunks<BN, CG>(p, tc.n0) + 1) >> 1;
    for (int q = 0; q < np; ++q) {
      const uint32_t g = (pc + static_cast<uint32_t>(q)) & 1u;
      for (int i = 0; i < p.n_res; ++i) {
        const uint32_t slot = base[g] + pos[g] % dg[g];
        mbar_wait(&res_empty[slot], ((pos[g] / dg[g]) & 1u) ^ 1u);  // fresh barriers count as released
        mbar_arrive_expect_tx(&res_full[slot], tx);
        tma_load_4d(res_ring + slot * 16384u, i ? &p.tmR1 : &p.tmR0, &res_full[slot], tc.n0 + q * 64, tc.o1, tc.o2,
                    tc.o3);
        ++pos[g];
      }
    }
    pc += static_cast<uint32_t>(np);
  }
}

template <int BN, int CG, bool RS>
__device__ __forceinline__ void epilogue_warp_tma_wide(const GemmKParams& p, int warp, int lane, int rank, int tile0,
                                                       int tile_step, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                       uint64_t* tmem_empty_bar, uint64_t* res_full,
                                                       uint64_t* res_empty, uint8_t* res_ring, uint8_t* out_ring) {
  const int ew = warp - 4;
  const uint32_t g = static_cast<uint32_t>(ew) >> 2;
  const int qd = warp & 3;
  const int r0 = qd * 32;
  const bool has_rows = r0 < p.rows_per_tile;
  const int q1 = r0 % p.box[0], q2 = (r0 / p.box[0]) % p.box[1], q3 = r0 / (p.box[0] * p.box[1]);
  const int r = r0 + lane;
  const int r1 = r % p.box[0], r2 = (r / p.box[0]) % p.box[1], r3 = r / (p.box[0] * p.box[1]);
  const uint32_t stage = smem_u32(out_ring) + static_cast<uint32_t>(ew) * 4096u;  // 32 rows x 128 B, one slot
  const uint32_t my_row = stage + static_cast<uint32_t>(lane) * 128u;
  const uint32_t res_row = static_cast<uint32_t>(r) * 128u;                          // this row inside a residual slot
  const uint32_t swz = static_cast<uint32_t>(lane & 7);                              // r & 7 == lane & 7
  const uint32_t D = static_cast<uint32_t>(p.n_res_slots);
  const uint32_t nres = static_cast<uint32_t>(p.n_res);
  const uint32_t ring_d = g ? (D >> 1) : ((D + 1u) >> 1), ring_base = g ? ((D + 1u) >> 1) : 0u;  // this group's ring
  uint32_t rpos = 0;  // residual entries this group has consumed
  auto release_acc = [](uint64_t* bar) {
    if constexpr (CG == 2) mbar_arrive_pair_leader(bar); else mbar_arrive(bar);
  };
  uint32_t pc = 0, t = 0;  // pc = panels of all earlier tiles (panel parity -> group)
  for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++t) {
    const TileCoord tc = decode_tile<BN, CG>(p, tile, rank);
    const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
    const int chunks = wide_chunks<BN, CG>(p, tc.n0);
    const int n_panels = (chunks + 1) >> 1;
    int64_t grow = -1;
    if (RS || p.add_ptr != nullptr) grow = tile_out_row(p, tc, r, r1, r2, r3);
    const float* addp = nullptr;
    if (p.add_ptr != nullptr) addp = p.add_ptr + ((grow < 0 ? 0 : grow) / p.add_div) * p.add_ld;
    float rs = 1.f, rm = 0.f;
    if constexpr (RS) row_fold(p, grow, rs, rm);
    int q_last = n_panels - 1;
    if (((pc + q_last) & 1u) != g) --q_last;
    mbar_wait(&tmem_full_bar[acc], acc_ph);
    tc_fence_after();
    ASVA_TR(p, warp, 3 + 20 * t);
    int trp = 0;
    if (q_last < 0 || !has_rows) {  // nothing of this tile is ours: hand the accumulator back right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty_bar[acc]);
      pc += n_panels;
      continue;
    }
    const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(r0) << 16);
#pragma unroll 1
    for (int q = 0; q < n_panels; ++q) {
      if (((pc + q) & 1u) != g) continue;
      const bool two = (2 * q + 1) < chunks;  // the panel has its second 32-column half
      const int acol = tc.n0 + q * 64;
      uint32_t v0[32], v1[32];
      tmem_ld_x32(taddr + q * 64, v0);
      if (two) tmem_ld_x32(taddr + q * 64 + 32, v1);
      // bias (+ per-row addend) of one 32-column half at a time: the first half's loads fly under the TMEM load, the
      // second half's under the first half's arithmetic - both halves at once would not fit the register file
      float4 b[8];
      auto load_bias = [&](int col0, bool on) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on && p.bias != nullptr && col0 + j * 4 < p.N) b[j] = ldg4(p.bias + col0 + j * 4);
        }
        if (on && addp != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col0 + j * 4 < p.N) {
              const float4 a = ldg4(addp + col0 + j * 4);
              b[j].x += a.x; b[j].y += a.y; b[j].z += a.z; b[j].w += a.w;
            }
          }
        }
        if constexpr (RS) {
          if (on) fold_bias(p, b, col0, rm);
        }
      };
      load_bias(acol, true);
      tmem_ld_wait();
      if (q == q_last) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&tmem_empty_bar[acc]);
      }
      ASVA_TR(p, warp, 4 + 20 * t + 6 * trp);
      // accumulator (* the row's rstd under the LayerNorm fold, else * 1: exact) + bias (+ row addend) in fp32, in place
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v0[4 * j + 0] = __float_as_uint(RS ? fmaf(__uint_as_float(v0[4 * j + 0]), rs, b[j].x) : __uint_as_float(v0[4 * j + 0]) + b[j].x);
        v0[4 * j + 1] = __float_as_uint(RS ? fmaf(__uint_as_float(v0[4 * j + 1]), rs, b[j].y) : __uint_as_float(v0[4 * j + 1]) + b[j].y);
        v0[4 * j + 2] = __float_as_uint(RS ? fmaf(__uint_as_float(v0[4 * j + 2]), rs, b[j].z) : __uint_as_float(v0[4 * j + 2]) + b[j].z);
        v0[4 * j + 3] = __float_as_uint(RS ? fmaf(__uint_as_float(v0[4 * j + 3]), rs, b[j].w) : __uint_as_float(v0[4 * j + 3]) + b[j].w);
      }
      if (two) {
        load_bias(acol + 32, true);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v1[4 * j + 0] = __float_as_uint(RS ? fmaf(__uint_as_float(v1[4 * j + 0]), rs, b[j].x) : __uint_as_float(v1[4 * j + 0]) + b[j].x);
          v1[4 * j + 1] = __float_as_uint(RS ? fmaf(__uint_as_float(v1[4 * j + 1]), rs, b[j].y) : __uint_as_float(v1[4 * j + 1]) + b[j].y);
          v1[4 * j + 2] = __float_as_uint(RS ? fmaf(__uint_as_float(v1[4 * j + 2]), rs, b[j].z) : __uint_as_float(v1[4 * j + 2]) + b[j].z);
          v1[4 * j + 3] = __float_as_uint(RS ? fmaf(__uint_as_float(v1[4 * j + 3]), rs, b[j].w) : __uint_as_float(v1[4 * j + 3]) + b[j].w);
        }
      }
      ASVA_TR(p, warp, 5 + 20 * t + 6 * trp);
      // residual panels: added in fp32 (the stored value is rounded to bf16 exactly once)
      for (uint32_t i = 0; i < nres; ++i) {
        const uint32_t slot = ring_base + rpos % ring_d, ph = (rpos / ring_d) & 1u;
        ++rpos;
        mbar_wait(&res_full[slot], ph);
        const uint32_t rp = smem_u32(res_ring) + slot * 16384u + res_row;
        auto add_half = [&](uint32_t (&v)[32], int c0) {  // 32 columns = 4 chunks of 8 bf16
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(rp + ((static_cast<uint32_t>(c0 + c) ^ swz) << 4))
                         : "memory");
            v[8 * c + 0] = __float_as_uint(__uint_as_float(v[8 * c + 0]) + __uint_as_float(w0 << 16));
            v[8 * c + 1] = __float_as_uint(__uint_as_float(v[8 * c + 1]) + __uint_as_float(w0 & 0xffff0000u));
            v[8 * c + 2] = __float_as_uint(__uint_as_float(v[8 * c + 2]) + __uint_as_float(w1 << 16));
            v[8 * c + 3] = __float_as_uint(__uint_as_float(v[8 * c + 3]) + __uint_as_float(w1 & 0xffff0000u));
            v[8 * c + 4] = __float_as_uint(__uint_as_float(v[8 * c + 4]) + __uint_as_float(w2 << 16));
            v[8 * c + 5] = __float_as_uint(__uint_as_float(v[8 * c + 5]) + __uint_as_float(w2 & 0xffff0000u));
            v[8 * c + 6] = __float_as_uint(__uint_as_float(v[8 * c + 6]) + __uint_as_float(w3 << 16));
            v[8 * c + 7] = __float_as_uint(__uint_as_float(v[8 * c + 7]) + __uint_as_float(w3 & 0xffff0000u));
          }
        };
        add_half(v0, 0);
        if (two) add_half(v1, 4);
        __syncwarp();  // every lane of this warp has read its rows of the slot
        if (lane == 0) mbar_arrive(&res_empty[slot]);
      }
      if constexpr (RS) {
        if (p.stats_out != nullptr) {
          emit_row_stats(p, v0, acol, grow);
          if (two) emit_row_stats(p, v1, acol + 32, grow);
        }
      }
      uint32_t pk[32];  // the row's 64 bf16 outputs, packed
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
      if (two) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
      }
      ASVA_TR(p, warp, 6 + 20 * t + 6 * trp);
      if (lane == 0) bulk_wait_read<0>();  // this warp's previous store has finished reading the staging slot
      __syncwarp();
      ASVA_TR(p, warp, 7 + 20 * t + 6 * trp);
      if (two) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(my_row + ((static_cast<uint32_t>(c) ^ swz) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      } else {
        // 32-column tail panel: rows of 64 B under the 64B swizzle of its own tensor map (tmO2)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t lin = static_cast<uint32_t>(lane) * 64u + static_cast<uint32_t>(c) * 16u;
          st_shared_v4(stage + (lin ^ (((lin >> 7) & 3u) << 4)), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
      ASVA_TR(p, warp, 8 + 20 * t + 6 * trp);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(two ? &p.tmO : &p.tmO2)), "r"(stage), "r"(acol), "r"(tc.o1 + q1),
                     "r"(tc.o2 + q2), "r"(tc.o3 + q3), "r"(tc.split)
                     : "memory");
        bulk_commit();
      }
      ASVA_TR(p, warp, 9 + 20 * t + 6 * trp);
      ++trp;
    }
    pc += n_panels;
  }
  if (lane == 0) bulk_wait_read<0>();
  ASVA_TR(p, warp, 63);
}

// Narrow form of the warp-private TMA epilogue (GEGLU and fp32 outputs: 32-column panels).
// (original text) The panel epilogue above moves a 128-row x 32-column panel per GROUP of four
// warps: two named barriers and one elected thread's TMA store per panel, one lane feeding the group's residual ring -
// two serial latency chains per CTA, which is what bounds every short-K launch. Here each of the
// eight epilogue warps owns the 32 rows of its TMEM lane quadrant outright: its own staging slots, its own residual
// ring and mbarriers, its own TMA loads and stores through tensor maps whose row box is the quadrant's 32-row slice of
// the tile (host: sub-box feasibility). No block-level barrier is left in the epilogue - eight independent chains per
// CTA keep the TMA / TMEM / L2 latencies of one panel under the work of the others.
template <int BN, bool GEGLU, int CG, bool RS>
__device__ __forceinline__ void epilogue_warp_tma_narrow(const GemmKParams& p, int warp, int lane, int rank, int tile0,
                                                  int tile_step, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                  uint64_t* tmem_empty_bar, uint64_t* res_full_bar, uint8_t* res_ring,
                                                  uint8_t* out_ring) {
  const int ew = warp - 4;
  const uint32_t g = static_cast<uint32_t>(ew) >> 2;
  const int qd = warp & 3;
  const int r0 = qd * 32;                       // first tile row of this warp's slice
  const bool has_rows = r0 < p.rows_per_tile;   // short tiles: the upper quadrants hold no rows
  const int q1 = r0 % p.box[0], q2 = (r0 / p.box[0]) % p.box[1], q3 = r0 / (p.box[0] * p.box[1]);
  const int r = r0 + lane;                      // this thread's row in the tile (== its TMEM lane)
  const int r1 = r % p.box[0], r2 = (r / p.box[0]) % p.box[1], r3 = r / (p.box[0] * p.box[1]);
  const uint32_t slot_bytes = p.out_fp32 ? 4096u : 2048u;  // 32 rows x 32 columns
  uint8_t* my_out = out_ring + static_cast<uint32_t>(ew) * 2u * slot_bytes;
  const uint32_t D = static_cast<uint32_t>(p.n_res_slots);
  uint8_t* my_res = res_ring + static_cast<uint32_t>(ew) * D * 2048u;
  uint64_t* my_bar = res_full_bar + ew * kMaxResSlots;
  auto release_acc = [](uint64_t* bar) {
    if constexpr (CG == 2) mbar_arrive_pair_leader(bar); else mbar_arrive(bar);
  };
  // residual prefetch, D panels ahead of the warp's own position (across tile boundaries); lane 0 issues
  const bool pf_lane = has_rows && (lane == 0) && (p.n_res > 0);
  uint32_t rslot = 0, rph = 0;
  int pf_tile = tile0, pf_q = 0, pf_i = 0;
  uint32_t pf_pc = 0, pf_slot = 0;
  TileCoord pf_tc = decode_tile<BN, CG>(p, tile0 < p.total_tiles ? tile0 : 0, rank);
  int pf_np = tile_panels<BN, GEGLU>(p, pf_tc.n0);
  auto pf_issue = [&]() {
    while (pf_tile < p.total_tiles) {
      while (pf_q < pf_np && (((pf_pc + pf_q) & 1u) != g)) ++pf_q;
      if (pf_q < pf_np) break;
      pf_pc += pf_np;
      pf_tile += tile_step;
      pf_q = 0;
      if (pf_tile < p.total_tiles) {
        pf_tc = decode_tile<BN, CG>(p, pf_tile, rank);
        pf_np = tile_panels<BN, GEGLU>(p, pf_tc.n0);
      }
    }
    if (pf_tile >= p.total_tiles) return;
    mbar_arrive_expect_tx(&my_bar[pf_slot], static_cast<uint32_t>(p.sub_rows) * 64u);
    tma_load_4d(my_res + pf_slot * 2048u, pf_i ? &p.tmR1 : &p.tmR0, &my_bar[pf_slot], pf_tc.n0 + pf_q * 32,
                pf_tc.o1 + q1, pf_tc.o2 + q2, pf_tc.o3 + q3);
    if (++pf_slot == D) pf_slot = 0;
    if (++pf_i == p.n_res) {
      pf_i = 0;
      ++pf_q;
    }
  };
  if (pf_lane)
    for (uint32_t i = 0; i < D; ++i) pf_issue();
  uint32_t pc = 0, ocnt = 0, t = 0;
  for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++t) {
    const TileCoord tc = decode_tile<BN, CG>(p, tile, rank);
    const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
    const int n_panels = tile_panels<BN, GEGLU>(p, tc.n0);
    const int n0_out = GEGLU ? (tc.n0 >> 1) : tc.n0;
    int64_t grow = -1;
    if (RS || p.add_ptr != nullptr) grow = tile_out_row(p, tc, r, r1, r2, r3);
    const float* addp = nullptr;
    if (p.add_ptr != nullptr) addp = p.add_ptr + ((grow < 0 ? 0 : grow) / p.add_div) * p.add_ld;
    float rs = 1.f, rm = 0.f;
    if constexpr (RS) row_fold(p, grow, rs, rm);
    int q_last = n_panels - 1;
    if (((pc + q_last) & 1u) != g) --q_last;
    mbar_wait(&tmem_full_bar[acc], acc_ph);
    tc_fence_after();
    ASVA_TR(p, warp, 3 + 20 * t);
    int trp = 0;
    if (q_last < 0 || !has_rows) {  // nothing of this tile is ours: hand the accumulator back right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(&tmem_empty_bar[acc]);
      pc += n_panels;
      continue;
    }
    const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(r0) << 16);
#pragma unroll 1
    for (int q = 0; q < n_panels; ++q) {
      if (((pc + q) & 1u) != g) continue;
      uint32_t v[32];
      if constexpr (!GEGLU) {
        tmem_ld_x32(taddr + q * 32, v);
        tmem_ld_wait();
        ASVA_TR(p, warp, 4 + 20 * t + 6 * trp);
        if (q == q_last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(&tmem_empty_bar[acc]);
        }
        const int acol = tc.n0 + q * 32;
        if constexpr (RS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // acc * rstd (1 without the LayerNorm fold: exact) + column terms
            const float4 b = col_terms(p, addp, acol + j * 4, rm);
            v[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 0]), rs, b.x));
            v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), rs, b.y));
            v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), rs, b.z));
            v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), rs, b.w));
          }
        } else {
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (acol + j * 4 < p.N) {
                const float4 b = ldg4(p.bias + acol + j * 4);
                v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
                v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
              }
            }
          }
          if (addp != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (acol + j * 4 < p.N) {
                const float4 b = ldg4(addp + acol + j * 4);
                v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
                v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
              }
            }
          }
        }
        ASVA_TR(p, warp, 5 + 20 * t + 6 * trp);
        for (int i = 0; i < p.n_res; ++i) {
          mbar_wait(&my_bar[rslot], rph);
          const uint8_t* rp = my_res + rslot * 2048u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t lin = static_cast<uint32_t>(lane) * 64u + j * 16u;
            const uint4 w = *reinterpret_cast<const uint4*>(rp + (lin ^ (((lin >> 7) & 3u) << 4)));
            const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z),
                         f3 = unpack_bf16x2(w.w);
            v[8 * j + 0] = __float_as_uint(__uint_as_float(v[8 * j + 0]) + f0.x);
            v[8 * j + 1] = __float_as_uint(__uint_as_float(v[8 * j + 1]) + f0.y);
            v[8 * j + 2] = __float_as_uint(__uint_as_float(v[8 * j + 2]) + f1.x);
            v[8 * j + 3] = __float_as_uint(__uint_as_float(v[8 * j + 3]) + f1.y);
            v[8 * j + 4] = __float_as_uint(__uint_as_float(v[8 * j + 4]) + f2.x);
            v[8 * j + 5] = __float_as_uint(__uint_as_float(v[8 * j + 5]) + f2.y);
            v[8 * j + 6] = __float_as_uint(__uint_as_float(v[8 * j + 6]) + f3.x);
            v[8 * j + 7] = __float_as_uint(__uint_as_float(v[8 * j + 7]) + f3.y);
          }
          if (++rslot == D) {
            rslot = 0;
            rph ^= 1u;
          }
        }
        if constexpr (RS) {
          if (p.stats_out != nullptr) emit_row_stats(p, v, acol, grow);
        }
      } else {
        // tile columns [0,64) = value h, [64,128) = gate g; output = (h + bh) * gelu(g + bg)
        uint32_t gv[32];
        tmem_ld_x32(taddr + q * 32, v);
        tmem_ld_x32(taddr + 64 + q * 32, gv);
        tmem_ld_wait();
        if (q == q_last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(&tmem_empty_bar[acc]);
        }
        const int acol = tc.n0 + q * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (RS) {
            const float4 bh = col_terms(p, nullptr, acol + j * 4, rm), bg = col_terms(p, nullptr, acol + 64 + j * 4, rm);
            v[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 0]), rs, bh.x) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 0]), rs, bg.x)));
            v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), rs, bh.y) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 1]), rs, bg.y)));
            v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), rs, bh.z) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 2]), rs, bg.z)));
            v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), rs, bh.w) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 3]), rs, bg.w)));
          } else {
            float4 bh = make_float4(0.f, 0.f, 0.f, 0.f), bg = bh;
            if (p.bias != nullptr) {
              bh = ldg4(p.bias + acol + j * 4);
              bg = ldg4(p.bias + acol + 64 + j * 4);
            }
            v[4 * j + 0] = __float_as_uint((__uint_as_float(v[4 * j + 0]) + bh.x) * gelu_erf_f(__uint_as_float(gv[4 * j + 0]) + bg.x));
            v[4 * j + 1] = __float_as_uint((__uint_as_float(v[4 * j + 1]) + bh.y) * gelu_erf_f(__uint_as_float(gv[4 * j + 1]) + bg.y));
            v[4 * j + 2] = __float_as_uint((__uint_as_float(v[4 * j + 2]) + bh.z) * gelu_erf_f(__uint_as_float(gv[4 * j + 2]) + bg.z));
            v[4 * j + 3] = __float_as_uint((__uint_as_float(v[4 * j + 3]) + bh.w) * gelu_erf_f(__uint_as_float(gv[4 * j + 3]) + bg.w));
          }
        }
      }
      ASVA_TR(p, warp, 6 + 20 * t + 6 * trp);
      // every lane is past its reads of this panel's residual slots (and of the staging slot's previous content)
      uint8_t* op = my_out + (ocnt & 1u) * slot_bytes;
      if (lane == 0) bulk_wait_read<1>();  // the store that last used this staging slot has finished reading it
      __syncwarp();
      ASVA_TR(p, warp, 7 + 20 * t + 6 * trp);
      if (pf_lane)
        for (int i = 0; i < p.n_res; ++i) pf_issue();
      if (!p.out_fp32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
          w.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
          w.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
          w.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
          const uint32_t lin = static_cast<uint32_t>(lane) * 64u + j * 16u;
          *reinterpret_cast<uint4*>(op + (lin ^ (((lin >> 7) & 3u) << 4))) = w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t lin = static_cast<uint32_t>(lane) * 128u + j * 16u;
          *reinterpret_cast<uint4*>(op + (lin ^ (((lin >> 7) & 7u) << 4))) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      ASVA_TR(p, warp, 8 + 20 * t + 6 * trp);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_5d(&p.tmO, op, n0_out + q * 32, tc.o1 + q1, tc.o2 + q2, tc.o3 + q3, tc.split);
        bulk_commit();
      }
      ASVA_TR(p, warp, 9 + 20 * t + 6 * trp);
      ++trp;
      ++ocnt;
    }
    pc += n_panels;
  }
  if (lane == 0) bulk_wait_read<0>();
  ASVA_TR(p, warp, 63);
}

// Cluster split-K (epi == 4).  For problems with few rows and a long K (M = 384 at the deepest level: three m tiles)
// every tile plan is either operand-feed-bound (narrow tiles re-read A per n tile) or leaves most SMs idle; splitting K
// fixes both but the workspace form pays a second launch and a round trip of fp32 partials through L2.  Here the S
// splits of a tile are the S CTAs of one thread-block cluster: each runs the unchanged main loop on its K range into
// TMEM, then (phase B) every CTA sends each 32-column chunk of its fp32 partial to the CTA that owns that column slice -
// st.shared::cluster into a receive area that reuses the ring's shared memory - and (phase C) every CTA sums the S
// partials of its own slice in rank order (deterministic), applies bias / row addend / residuals and stores the rows.
// Two barrier.cluster round trips replace the reduce kernel.  One tile per CTA (grid = tiles x S).
template <int BN>
__device__ __forceinline__ void csplit_scatter(const GemmKParams& p, int warp, int lane, uint32_t srank,
                                               uint32_t tmem_base, uint32_t recv_base) {
  const int qd = warp & 3;
  const uint32_t g = static_cast<uint32_t>(warp - 4) >> 2;
  const int row = qd * 32 + lane;
  const int wsl = BN / p.csplit;       // slice width (a multiple of 32)
  const int nlc = wsl >> 5;            // 32-column chunks per slice
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
  const uint32_t sw = static_cast<uint32_t>(row & 7);
#pragma unroll 1
  for (int c = static_cast<int>(g); c < BN / 32; c += 2) {
    uint32_t v[32];
    tmem_ld_x32(taddr + c * 32, v);
    tmem_ld_wait();
    const uint32_t owner = static_cast<uint32_t>((c * 32) / wsl);
    const int lc = c - static_cast<int>(owner) * nlc;
    // receive area of the owner: [source rank][slice chunk][row][32 fp32], 16-byte pieces XOR-swizzled by the row
    const uint32_t local = recv_base + ((srank * nlc + lc) * 128u + static_cast<uint32_t>(row)) * 128u;
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(owner));
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra + ((static_cast<uint32_t>(j) ^ sw) << 4)),
                   "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                   : "memory");
  }
}

template <int BN, int CG>
__device__ __forceinline__ void csplit_finish(const GemmKParams& p, int warp, int lane, uint32_t srank, int tile,
                                              uint32_t recv_base) {
  const int qd = warp & 3;
  const uint32_t g = static_cast<uint32_t>(warp - 4) >> 2;
  const int r = qd * 32 + lane;
  const int r1 = r % p.box[0], r2 = (r / p.box[0]) % p.box[1], r3 = r / (p.box[0] * p.box[1]);
  const TileCoord tc = decode_tile<BN, CG>(p, tile, 0, static_cast<int>(srank));
  const int64_t grow = tile_out_row(p, tc, r, r1, r2, r3);
  const int wsl = BN / p.csplit, nlc = wsl >> 5;
  const uint32_t sw = static_cast<uint32_t>(r & 7);
  const float* addp = (p.add_ptr != nullptr && grow >= 0) ? p.add_ptr + (grow / p.add_div) * p.add_ld : nullptr;
#pragma unroll 1
  for (int lc = static_cast<int>(g); lc < nlc; lc += 2) {
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int src = 0; src < p.csplit; ++src) {  // rank order: the sum does not depend on arrival order
      const uint32_t a = recv_base + ((static_cast<uint32_t>(src) * nlc + lc) * 128u + static_cast<uint32_t>(r)) * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = ld_shared_f4(a + ((static_cast<uint32_t>(j) ^ sw) << 4));
        acc[4 * j] += x.x; acc[4 * j + 1] += x.y; acc[4 * j + 2] += x.z; acc[4 * j + 3] += x.w;
      }
    }
    if (grow < 0) continue;
    const int col0 = tc.n0 + static_cast<int>(srank) * wsl + lc * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 8 columns at a time
      const int col = col0 + 8 * j;
      if (col >= p.N) continue;
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = acc[8 * j + e];
      if (p.bias != nullptr) {
        const float4 b0 = ldg4(p.bias + col), b1 = ldg4(p.bias + col + 4);
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
      if (addp != nullptr) {
        const float4 b0 = ldg4(addp + col), b1 = ldg4(addp + col + 4);
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i < p.n_res) {
          const uint4 w = *reinterpret_cast<const uint4*>(p.res_ptr[i] + grow * p.res_ld[i] + col);
          const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z), f3 = unpack_bf16x2(w.w);
          x[0] += f0.x; x[1] += f0.y; x[2] += f1.x; x[3] += f1.y; x[4] += f2.x; x[5] += f2.y; x[6] += f3.x; x[7] += f3.y;
        }
      }
      if (p.csplit_f32) {
        float* o = reinterpret_cast<float*>(p.out_ptr) + grow * p.ldo + col;
        *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
      } else {
        uint4 w;
        w.x = pack_bf16x2(x[0], x[1]);
        w.y = pack_bf16x2(x[2], x[3]);
        w.z = pack_bf16x2(x[4], x[5]);
        w.w = pack_bf16x2(x[6], x[7]);
        *reinterpret_cast<uint4*>(p.out_ptr + grow * p.ldo + col) = w;
      }
    }
  }
}

// RS: the instantiations that carry the row-statistics / LayerNorm-fold code (asva_gemm_desc.stats_out, .ln_*); the
// others are the plain epilogues, untouched by it (the fold's extra loads in the shared path cost the plain launches
// 10 - 60 %, measured: tools/lnfold_probe.py).
template <int BN, bool GEGLU, int CG, bool RS>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  constexpr int kABytes = 128 * 128;
  constexpr int kBBytes = (BN / CG) * 128;  // a CTA of a pair holds half of the W tile's rows
  constexpr int kStageBytes = kABytes + kBBytes;   // one 64-wide K block
  constexpr int kSuperBytes = 2 * kStageBytes;     // a ring stage = two K blocks
  constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_stages = p.n_stages;
  uint8_t* res_ring = smem + n_stages * kSuperBytes;
  uint8_t* out_ring = res_ring + 2 * p.n_res_slots * kResSlotBytes;
  const int out_slot_bytes = p.out_fp32 ? 16384 : 8192;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_ring + 4 * out_slot_bytes);
  uint64_t* full_bar = bars;                     // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;       // [kMaxStages]
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;       // [2 groups][kMaxResSlots] (epi 1) / [8 warps][kMaxResSlots] (epi 3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full_bar + 8 * kMaxResSlots);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  ASVA_TR(p, warp, 0);
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool csplit = (CG == 1) && p.csplit > 1;  // cluster split-K: the cluster rank is the K split of ONE tile
  const int srank = csplit ? static_cast<int>(cluster_ctarank()) : -1;
  const int cdiv = csplit ? p.csplit : CG;
  const int tile0 = blockIdx.x / cdiv, tile_step = gridDim.x / cdiv;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrival per producer thread (of the pair's even CTA)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8 * CG);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    // epi 3 (wide): [0..3] = residual slot full, [8..11] = slot released by every quadrant warp that holds rows
    const int nq = (p.rows_per_tile + 31) / 32;
    for (int s = 0; s < 8 * kMaxResSlots; ++s)
      mbar_init(&res_full_bar[s], (p.epi == 3 && s >= 8 && s < 12) ? static_cast<uint32_t>(nq < 4 ? nq : 4) : 1u);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmO);
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Weights do not depend on the previous kernel: while that kernel is still finishing, the CTAs of this one pull W
  // from HBM into L2 (the 2.3 GB of weights per step never survive in L2 from one step to the next) - the whole matrix,
  // dealt out block by block (64 columns x the TMA box rows) over the grid, earliest K blocks first, up to a cap per CTA.
  if (warp == 2 && lane == 0 && p.pf_blocks > 0) {
    const int row_boxes = (p.N + BN / CG - 1) / (BN / CG);
    const int total = p.w_kblocks * row_boxes;
    int n = 0;
    for (int b = blockIdx.x; b < total && n < p.pf_blocks; b += gridDim.x, ++n) {
      const int kbx = b / row_boxes, rbx = b - kbx * row_boxes;
      tma_prefetch_l2_2d(&p.tmW, kbx * 64, rbx * (BN / CG));
    }
  }
  ASVA_TR(p, warp, 1);
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are read from here on
  ASVA_TR(p, warp, 2);

  // The producer and MMA loops run in one thread each and are paced by their own instruction and mbarrier latency,
  // so they are written for a minimal dependent-instruction count: ring position kept as (stage, phase) counters,
  // shared addresses and descriptors advanced by adds, segment parameters reloaded only when a segment ends, and the
  // NEXT stage's barrier probed (non-blocking test_wait) before the current stage's work is issued.
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  if (warp == 0 || warp == 3) {
    // ---------------- TMA producers (A, W): warp j loads K block j of every stage ----------------
    // (all 32 lanes run the loop with identical values; `el` predicates the TMA / arrive instructions on one lane)
    {
      const uint32_t el = elect_one();
      const int j = (warp == 3) ? 1 : 0;
      const uint32_t tx_bytes = (static_cast<uint32_t>(p.rows_per_tile) * 128u + kBBytes) * CG;
      const uint32_t full0_l = (CG == 2) ? (full0 & kPeerBitMask) : full0;  // pair: the even CTA's barriers
      const uint32_t el_arm = (rank == 0) ? el : 0u;  // only the even CTA's producers arrive on its barriers
      const uint32_t n_st = static_cast<uint32_t>(n_stages);
      uint32_t s = 0, ph = 1;  // ph = parity to wait for on the empty barrier
      bool ready = true;       // fresh barriers: the "previous phase" of every empty barrier counts as complete
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
        const TileCoord tc = decode_tile<BN, CG>(p, tile, rank, srank);
        const int i1 = tc.o1 * p.trav[0], i2 = tc.o2 * p.trav[1], i3 = tc.o3 * p.trav[2];
        const int wn0 = tc.n0 + rank * (BN / CG);
        const bool first2 = (tc.o2 == 0);
        int kb = tc.kb0 + j, seg = 0, kin = kb;
        SegK sg = p.seg[0];
        const CUtensorMap* tmA = &p.tmA0;
        int c1 = 0, c2 = 0, c3 = 0, wbase = 0;
        auto derive = [&]() {
          tmA = sg.src ? &p.tmA1 : &p.tmA0;
          c1 = i1 + sg.off1;
          c3 = i3 + sg.off3;
          c2 = sg.fix2 >= 0 ? sg.fix2 : i2 + sg.off2;
          wbase = (sg.wk_first >= 0 && first2) ? sg.wk_first : sg.wk;
        };
        if (kb < tc.kb1) {
          while (kin >= sg.num_kb) {
            kin -= sg.num_kb;
            sg = p.seg[++seg];
          }
          derive();
        }
        for (int i = (tc.kb1 - tc.kb0 + 1) >> 1; i > 0; --i) {
          if (!ready) mbar_wait_a(empty0 + 8u * s, ph);
          uint32_t s2 = s + 1, ph2 = ph;
          if (s2 == n_st) {
            s2 = 0;
            ph2 ^= 1u;
          }
          ready = mbar_test_wait_a(empty0 + 8u * s2, ph2);
          const uint32_t sa = smem_a0 + s * kSuperBytes + j * kStageBytes;
          const uint32_t fb = full0_l + 8u * s;
          if constexpr (CG == 2) {
            // CTA pairs: the stage's instructions are issued by ONE thread from inside a branch on elect.sync - ptxas
            // then keeps the operands in uniform registers and emits the instructions back to back, where predicating
            // each instruction on an elected lane costs a vote / elect / retry sequence (~15 instructions) per TMA or
            // MMA instruction. With neither loads nor MMAs a stage then costs ~300 instead of ~590 cycles
            // (profiles/r1_gemm_dbg_sweep_v12.md); single-CTA plans keep the predicated form, whose per-MMA overhead
            // interleaves with the tensor core (back-to-back issue measured 3-7 % slower there).
            if (kb < tc.kb1 && !(ASVA_DBG(p) & 1)) {
              const int ccol = sg.c0 + kin * 64, wcol = wbase + kin * 64;
              if (elect_one()) {
                // both CTAs load their half; all bytes are credited to the even CTA's barrier, armed by its producers
                if (rank == 0) mbar_arrive_expect_tx_a(fb, tx_bytes);
                tma_load_4d_pair_a(sa, tmA, fb, ccol, c1, c2, c3);
                tma_load_2d_pair_a(sa + kABytes, &p.tmW, fb, wcol, wn0);
              }
            } else {
              if (elect_one() && rank == 0) mbar_arrive_a(fb);  // odd K-block count: no second block in the last stage
            }
            __syncwarp();
          } else {
            if (kb < tc.kb1 && !(ASVA_DBG(p) & 1)) {
              const int ccol = sg.c0 + kin * 64, wcol = wbase + kin * 64;
              mbar_arrive_expect_tx_p(el_arm, fb, tx_bytes);
              tma_load_4d_p(el, sa, tmA, fb, ccol, c1, c2, c3);
              tma_load_2d_p(el, sa + kABytes, &p.tmW, fb, wcol, wn0);
            } else {
              mbar_arrive_p(el_arm, fb);  // odd K-block count: the last stage of the tile has no second block
            }
          }
          kb += 2;
          kin += 2;
          if (kb < tc.kb1 && kin >= sg.num_kb) {
            do {
              kin -= sg.num_kb;
              sg = p.seg[++seg];
            } while (kin >= sg.num_kb);
            derive();
          }
          s = s2;
          ph = ph2;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform loop, instructions predicated on one elected lane) ----------------
    if (rank == 0) {
      const uint32_t el = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN);
      // smem descriptor (K-major, 128B swizzle): low word = (addr >> 4) | LBO(1) << 16, high word constant
      constexpr uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t a_lo0 = ((smem_a0 & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t n_st = static_cast<uint32_t>(n_stages);
      const uint32_t tfull0 = smem_u32(tmem_full_bar);
      // a value the compiler can prove warp-uniform (shared-memory loads are not), so the MMAs take it straight from
      // a uniform register instead of an elect-and-broadcast loop per instruction
      const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);
      uint32_t s = 0, ph = 0, t = 0;
      bool ready = false;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++t) {
        const TileCoord tc = decode_tile<BN, CG>(p, tile, 0, srank);
        const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator
        ASVA_TR(p, warp, 4 + 3 * t);
        const uint32_t tmem_d = tmem_u + acc * BN;
        uint32_t accumulate = 0;
        for (int n = tc.kb1 - tc.kb0; n > 0; n -= 2) {
          if (!ready) mbar_wait_a(full0 + 8u * s, ph);
          if (n == tc.kb1 - tc.kb0) ASVA_TR(p, warp, 5 + 3 * t);
          uint32_t s2 = s + 1, ph2 = ph;
          if (s2 == n_st) {
            s2 = 0;
            ph2 ^= 1u;
          }
          ready = mbar_test_wait_a(full0 + 8u * s2, ph2);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (kSuperBytes >> 4);
          if constexpr (CG == 2) {
            if (elect_one()) {  // one thread issues the stage's MMAs and their commit (see the producer loop)
              if (!(ASVA_DBG(p) & 2)) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  if (h == 1 && n < 2) break;  // odd K-block count: the tile's last stage holds one block
                  const uint64_t adesc = desc_hi | (a_lo + h * (kStageBytes >> 4));
                  const uint64_t bdesc = desc_hi | (a_lo + h * (kStageBytes >> 4) + (kABytes >> 4));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_ss_pair(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (h == 0 && k == 0) ? accumulate : 1u);
                }
              }
              tc_commit_pair_a(empty0 + 8u * s, 3);
            }
            __syncwarp();
          } else {
            const uint32_t el_mma = (ASVA_DBG(p) & 2) ? 0u : el;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t el_h = (h == 1 && n < 2) ? 0u : el_mma;
              const uint64_t adesc = desc_hi | (a_lo + h * (kStageBytes >> 4));
              const uint64_t bdesc = desc_hi | (a_lo + h * (kStageBytes >> 4) + (kABytes >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss_p(el_h, tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (h == 0 && k == 0) ? accumulate : 1u);
            }
            tc_commit_p(el, empty0 + 8u * s);
          }
          accumulate = 1u;
          s = s2;
          ph = ph2;
        }
        if constexpr (CG == 2) {
          if (elect_one()) tc_commit_pair_a(tfull0 + 8u * acc, 3);
          __syncwarp();
        } else {
          tc_commit_p(el, tfull0 + 8u * acc);
        }
        ASVA_TR(p, warp, 6 + 3 * t);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && p.epi == 4) {
    // cluster split-K: the tile's partial is complete in TMEM buffer 0; the exchange happens below, where every
    // thread of the cluster meets
    if (tile0 < p.total_tiles) {
      mbar_wait(&tmem_full_bar[0], 0);
      tc_fence_after();
    }
  } else if (warp >= 4 && !GEGLU && p.epi == 2) {
    epilogue_direct<BN, CG>(p, warp, lane, rank, tile0, tile_step, tmem_base, tmem_full_bar, tmem_empty_bar, out_ring);
  } else if (warp == 2) {
    if constexpr (!GEGLU) {
      if (p.epi == 3 && !p.out_fp32 && p.n_res > 0 && lane == 0)
        residual_issuer_wide<BN, CG>(p, rank, tile0, tile_step, res_full_bar, res_full_bar + 8, res_ring);
    }
    __syncwarp();
  } else if (warp >= 4 && p.epi == 3) {
    if constexpr (!GEGLU) {
      if (!p.out_fp32)
        epilogue_warp_tma_wide<BN, CG, RS>(p, warp, lane, rank, tile0, tile_step, tmem_base, tmem_full_bar, tmem_empty_bar,
                                       res_full_bar, res_full_bar + 8, res_ring, out_ring);
      else
        epilogue_warp_tma_narrow<BN, GEGLU, CG, RS>(p, warp, lane, rank, tile0, tile_step, tmem_base, tmem_full_bar,
                                                tmem_empty_bar, res_full_bar, res_ring, out_ring);
    } else {
      epilogue_warp_tma_narrow<BN, GEGLU, CG, RS>(p, warp, lane, rank, tile0, tile_step, tmem_base, tmem_full_bar,
                                              tmem_empty_bar, res_full_bar, res_ring, out_ring);
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const uint32_t g = static_cast<uint32_t>(warp - 4) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // TMEM lane == row inside the tile
    const int r1 = r % p.box[0];
    const int r2 = (r / p.box[0]) % p.box[1];
    const int r3 = r / (p.box[0] * p.box[1]);
    const bool leader = (qd == 0) && (lane == 0);
    auto release_acc = [](uint64_t* bar) {  // the accumulator-free barrier lives in the pair's even CTA
      if constexpr (CG == 2) mbar_arrive_pair_leader(bar); else mbar_arrive(bar);
    };
    uint8_t* my_out = out_ring + g * 2 * out_slot_bytes;
    uint32_t pc = 0, ocnt = 0, t = 0;
    // Residual operands: each group streams the residual panels of ITS OWN upcoming output panels through a private
    // ring of D slots, D panels ahead of where it is working (across tile boundaries). One lane of the group issues
    // the TMA loads; a slot is refilled right after the group barrier that follows its last read, so no "empty"
    // barrier is needed, and no single thread has to issue the residual loads of both groups.
    const uint32_t D = static_cast<uint32_t>(p.n_res_slots);
    const bool pf_lane = (qd == 1) && (lane == 0) && (p.n_res > 0);
    const uint32_t res_tx = static_cast<uint32_t>(p.rows_per_tile) * 64u;
    uint32_t rslot = 0, rph = 0;              // consumer position in the ring
    int pf_tile = tile0, pf_q = 0, pf_i = 0;  // producer position: next (tile, panel, residual) to fetch
    uint32_t pf_pc = 0, pf_slot = 0;
    TileCoord pf_tc = decode_tile<BN, CG>(p, tile0 < p.total_tiles ? tile0 : 0, rank);
    int pf_np = tile_panels<BN, GEGLU>(p, pf_tc.n0);
    auto pf_issue = [&]() {
      while (pf_tile < p.total_tiles) {  // advance to the next panel this group owns
        while (pf_q < pf_np && (((pf_pc + pf_q) & 1u) != g)) ++pf_q;
        if (pf_q < pf_np) break;
        pf_pc += pf_np;
        pf_tile += tile_step;
        pf_q = 0;
        if (pf_tile < p.total_tiles) {
          pf_tc = decode_tile<BN, CG>(p, pf_tile, rank);
          pf_np = tile_panels<BN, GEGLU>(p, pf_tc.n0);
        }
      }
      if (pf_tile >= p.total_tiles) return;
      const uint32_t slot = g * kMaxResSlots + pf_slot;
      mbar_arrive_expect_tx(&res_full_bar[slot], res_tx);
      tma_load_4d(res_ring + (g * D + pf_slot) * kResSlotBytes, pf_i ? &p.tmR1 : &p.tmR0, &res_full_bar[slot],
                  pf_tc.n0 + pf_q * 32, pf_tc.o1, pf_tc.o2, pf_tc.o3);
      if (++pf_slot == D) pf_slot = 0;
      if (++pf_i == p.n_res) {
        pf_i = 0;
        ++pf_q;
      }
    };
    if (pf_lane)
      for (uint32_t i = 0; i < D; ++i) pf_issue();
    for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++t) {
      const TileCoord tc = decode_tile<BN, CG>(p, tile, rank);
      const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const int n_panels = tile_panels<BN, GEGLU>(p, tc.n0);
      const int n0_out = GEGLU ? (tc.n0 >> 1) : tc.n0;
      int64_t grow = -1;
      if (RS || p.add_ptr != nullptr) grow = tile_out_row(p, tc, r, r1, r2, r3);
      const float* addp = nullptr;
      if (p.add_ptr != nullptr) addp = p.add_ptr + ((grow < 0 ? 0 : grow) / p.add_div) * p.add_ld;
      float rs = 1.f, rm = 0.f;
      if constexpr (RS) row_fold(p, grow, rs, rm);
      int q_last = n_panels - 1;
      if (((pc + q_last) & 1u) != g) --q_last;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      if (q_last < 0) {  // no panel of this tile is ours: hand the accumulator back right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&tmem_empty_bar[acc]);
      }
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(qd * 32) << 16);
#pragma unroll 1
      for (int q = 0; q < n_panels; ++q) {
        if (((pc + q) & 1u) != g) continue;
        uint32_t v[32];
        if constexpr (!GEGLU) {
          tmem_ld_x32(taddr + q * 32, v);
          tmem_ld_wait();
          if (q == q_last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(&tmem_empty_bar[acc]);
          }
          const int acol = tc.n0 + q * 32;
          if constexpr (RS) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // acc * rstd (1 without the LayerNorm fold: exact) + column terms
              const float4 b = col_terms(p, addp, acol + j * 4, rm);
              v[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 0]), rs, b.x));
              v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), rs, b.y));
              v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), rs, b.z));
              v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), rs, b.w));
            }
          } else {
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (acol + j * 4 < p.N) {
                  const float4 b = ldg4(p.bias + acol + j * 4);
                  v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
                  v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                  v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                  v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
                }
              }
            }
            if (addp != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (acol + j * 4 < p.N) {
                  const float4 b = ldg4(addp + acol + j * 4);
                  v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
                  v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
                  v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
                  v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
                }
              }
            }
          }
          for (int i = 0; i < p.n_res; ++i) {
            mbar_wait(&res_full_bar[g * kMaxResSlots + rslot], rph);
            const uint8_t* rp = res_ring + (g * D + rslot) * kResSlotBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t lin = static_cast<uint32_t>(r) * 64u + j * 16u;
              const uint4 w = *reinterpret_cast<const uint4*>(rp + (lin ^ (((lin >> 7) & 3u) << 4)));
              const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z),
                           f3 = unpack_bf16x2(w.w);
              v[8 * j + 0] = __float_as_uint(__uint_as_float(v[8 * j + 0]) + f0.x);
              v[8 * j + 1] = __float_as_uint(__uint_as_float(v[8 * j + 1]) + f0.y);
              v[8 * j + 2] = __float_as_uint(__uint_as_float(v[8 * j + 2]) + f1.x);
              v[8 * j + 3] = __float_as_uint(__uint_as_float(v[8 * j + 3]) + f1.y);
              v[8 * j + 4] = __float_as_uint(__uint_as_float(v[8 * j + 4]) + f2.x);
              v[8 * j + 5] = __float_as_uint(__uint_as_float(v[8 * j + 5]) + f2.y);
              v[8 * j + 6] = __float_as_uint(__uint_as_float(v[8 * j + 6]) + f3.x);
              v[8 * j + 7] = __float_as_uint(__uint_as_float(v[8 * j + 7]) + f3.y);
            }
            if (++rslot == D) {
              rslot = 0;
              rph ^= 1u;
            }
          }
          if constexpr (RS) {
            if (p.stats_out != nullptr) emit_row_stats(p, v, acol, grow);
          }
        } else {
          // tile columns [0,64) = value h, [64,128) = gate g; output = (h + bh) * gelu(g + bg)
          uint32_t gv[32];
          tmem_ld_x32(taddr + q * 32, v);
          tmem_ld_x32(taddr + 64 + q * 32, gv);
          tmem_ld_wait();
          if (q == q_last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(&tmem_empty_bar[acc]);
          }
          const int acol = tc.n0 + q * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if constexpr (RS) {
              const float4 bh = col_terms(p, nullptr, acol + j * 4, rm), bg = col_terms(p, nullptr, acol + 64 + j * 4, rm);
              v[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 0]), rs, bh.x) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 0]), rs, bg.x)));
              v[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), rs, bh.y) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 1]), rs, bg.y)));
              v[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), rs, bh.z) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 2]), rs, bg.z)));
              v[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), rs, bh.w) * gelu_erf_f(fmaf(__uint_as_float(gv[4 * j + 3]), rs, bg.w)));
            } else {
              float4 bh = make_float4(0.f, 0.f, 0.f, 0.f), bg = bh;
              if (p.bias != nullptr) {
                bh = ldg4(p.bias + acol + j * 4);
                bg = ldg4(p.bias + acol + 64 + j * 4);
              }
              v[4 * j + 0] = __float_as_uint((__uint_as_float(v[4 * j + 0]) + bh.x) * gelu_erf_f(__uint_as_float(gv[4 * j + 0]) + bg.x));
              v[4 * j + 1] = __float_as_uint((__uint_as_float(v[4 * j + 1]) + bh.y) * gelu_erf_f(__uint_as_float(gv[4 * j + 1]) + bg.y));
              v[4 * j + 2] = __float_as_uint((__uint_as_float(v[4 * j + 2]) + bh.z) * gelu_erf_f(__uint_as_float(gv[4 * j + 2]) + bg.z));
              v[4 * j + 3] = __float_as_uint((__uint_as_float(v[4 * j + 3]) + bh.w) * gelu_erf_f(__uint_as_float(gv[4 * j + 3]) + bg.w));
            }
          }
        }
        // ---- stage the panel (swizzled exactly as the output tensor map expects) and store it with TMA
        uint8_t* op = my_out + (ocnt & 1u) * out_slot_bytes;
        if (leader) bulk_wait_read<1>();  // the store that last used this slot has finished reading it
        named_bar_sync(1 + g, 128);
        if (pf_lane)  // every warp of the group is past its reads of this panel's residual slots: refill them
          for (int i = 0; i < p.n_res; ++i) pf_issue();
        if (!p.out_fp32) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            w.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            w.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            w.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            const uint32_t lin = static_cast<uint32_t>(r) * 64u + j * 16u;
            *reinterpret_cast<uint4*>(op + (lin ^ (((lin >> 7) & 3u) << 4))) = w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t lin = static_cast<uint32_t>(r) * 128u + j * 16u;
            *reinterpret_cast<uint4*>(op + (lin ^ (((lin >> 7) & 7u) << 4))) =
                make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + g, 128);
        if (leader) {
          tma_store_5d(&p.tmO, op, n0_out + q * 32, tc.o1, tc.o2, tc.o3, tc.split);
          bulk_commit();
        }
        ++ocnt;
      }
      pc += n_panels;
    }
    if (leader) bulk_wait_read<0>();  // shared memory may be released; completion of the writes is the grid's completion
  }
  if constexpr (CG == 1 && !GEGLU) {
    if (csplit) {
      // every CTA's main loop is over (its epilogue warps saw the accumulator complete, its producers / issuer left
      // their loops): the ring's shared memory is free to receive partial slices
      tc_fence_before();
      cluster_sync_all();
      const uint32_t recv = smem_u32(smem);
      if (warp >= 4 && tile0 < p.total_tiles) {
        tc_fence_after();
        csplit_scatter<BN>(p, warp, lane, static_cast<uint32_t>(srank), tmem_base, recv);
      }
      cluster_sync_all();  // all slices have landed (release / acquire at cluster scope)
      if (warp >= 4 && tile0 < p.total_tiles) csplit_finish<BN, CG>(p, warp, lane, static_cast<uint32_t>(srank), tile0, recv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA of a pair may leave while the other still uses it
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Split-K second pass: out = sum_s ws[s] + bias + addend + residuals, 8 columns per thread.
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int S, int64_t M, int N,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ add_ptr, int64_t add_ld,
                                                              int add_div, const __nv_bfloat16* res0, int64_t ld0,
                                                              const __nv_bfloat16* res1, int64_t ld1, void* out,
                                                              int64_t ldo, int out_fp32, float* stats_out) {
  pdl_trigger();
  pdl_wait();
  const int nchunk = N >> 3;
  const int64_t total = M * nchunk;
  const int64_t plane = M * static_cast<int64_t>(N);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool on = i < total;
    if (__ballot_sync(0xffffffffu, on) == 0u) break;  // the whole warp leaves together (row statistics shuffle)
    const int64_t row = on ? i / nchunk : 0;
    const int col = on ? static_cast<int>(i % nchunk) * 8 : 0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (on) {
    const float* src = ws + row * N + col;
    for (int s = 0; s < S; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(src + s * plane);
      const float4 b = *reinterpret_cast<const float4*>(src + s * plane + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (bias != nullptr) {
      const float4 a = ldg4(bias + col), b = ldg4(bias + col + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (add_ptr != nullptr) {
      const float* ap = add_ptr + (row / add_div) * add_ld + col;
      const float4 a = ldg4(ap), b = ldg4(ap + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const __nv_bfloat16* rp = k ? res1 : res0;
      if (rp == nullptr) continue;
      const uint4 w = *reinterpret_cast<const uint4*>(rp + row * (k ? ld1 : ld0) + col);
      const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z), f3 = unpack_bf16x2(w.w);
      v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y; v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
    }
    if (out_fp32) {
      float* o = reinterpret_cast<float*>(out) + row * ldo + col;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 w;
      w.x = pack_bf16x2(v[0], v[1]);
      w.y = pack_bf16x2(v[2], v[3]);
      w.z = pack_bf16x2(v[4], v[5]);
      w.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + row * ldo + col) = w;
    }
    }
    if (stats_out != nullptr) {
      // N % 32 == 0: the four lanes of an aligned quad hold one row's 32-column slot (asva_gemm_desc.stats_out)
      float sm = 0.f, sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sm += v[j];
        sq = fmaf(v[j], v[j], sq);
      }
      sm += __shfl_xor_sync(0xffffffffu, sm, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sm += __shfl_xor_sync(0xffffffffu, sm, 2);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      if (on && (threadIdx.x & 3) == 0)
        *reinterpret_cast<float2*>(stats_out + (static_cast<int64_t>(col >> 5) * M + row) * 2) = make_float2(sm, sq);
    }
  }
}

struct GemmPlan {
  int bn, split, stages, cg, epi;
};

static int fixed_smem(int res_slots, int out_fp32) {
  return 1024 /*align*/ + kBarBytes + 4 * (out_fp32 ? 16384 : 8192) + 2 * res_slots * kResSlotBytes;
}
// ring stages of two 64-wide K blocks each, given the residual slots per epilogue group
static int stages_with(int bn, int cg, int res_slots, int out_fp32) {
  const int s = (kSmemLimit - fixed_smem(res_slots, out_fp32)) / (2 * (16384 + (bn / cg) * 128));
  return s > kMaxStages ? kMaxStages : s;
}
// Shared memory split between the main-loop ring and the residual ring: as many residual slots (<= 4 per group) as
// leave the main loop the stages it can use (3, or fewer for a short K loop), never fewer than 2 of either.
static int res_slots_for(int bn, int cg, int n_res, int out_fp32, int num_kb) {
  if (n_res == 0) return 0;
  int want = (num_kb + 1) / 2;
  if (want > 3) want = 3;
  if (want < 2) want = 2;
  for (int d = kMaxResSlots; d > 2; --d)
    if (stages_with(bn, cg, d, out_fp32) >= want) return d;
  return 2;
}
static int stages_for(int bn, int cg, int n_res, int out_fp32, int num_kb) {
  return stages_with(bn, cg, res_slots_for(bn, cg, n_res, out_fp32, num_kb), out_fp32);
}
static int smem_for(int bn, int cg, int stages, int res_slots, int out_fp32) {
  return fixed_smem(res_slots, out_fp32) + stages * 2 * (16384 + (bn / cg) * 128);
}

// Cost model (cycles) behind the automatic tile-width / CTA-pair / split-K choice. Per 64-wide K block a CTA needs
// max(MMA issue time = 2*BN cycles, operand bytes / its share of the L2->SM bandwidth); a launch takes
// ceil(tiles / SMs) waves of (K blocks * that + fixed per-tile cost); split-K adds the reduce pass. Callers that
// can measure (the engine's tuner) pass block_n / cta_group / split_k explicitly instead.
static double plan_cost(int bn, int cg, int split, int N, int64_t m_tiles, int num_kb, int64_t M, int sms) {
  const int n_tiles = (N + bn - 1) / bn;
  const int kbps = (num_kb + split - 1) / split;
  const int64_t tiles = ((m_tiles + cg - 1) / cg) * cg * n_tiles * split;  // in CTAs
  const int64_t ctas = tiles < sms ? tiles : sms;
  const int64_t waves = (tiles + sms - 1) / sms;
  double bw = 6500.0 / static_cast<double>(ctas);  // bytes / cycle / SM
  if (bw > 80.0) bw = 80.0;
  const double feed = (16384.0 + (bn / cg) * 128.0) / bw;
  const double mma = 2.0 * bn;
  const double per_kb = feed > mma ? feed : mma;
  const double epi = 300.0 * ((bn + 31) / 32) / 2.0 + 400.0;  // per tile, two groups in parallel
  double tile = kbps * per_kb;
  if (tile < epi) tile = epi;
  double cost = 2500.0 + waves * (tile + 700.0) + (cg == 2 ? 600.0 : 0.0);
  if (split > 1) cost += 5000.0 + static_cast<double>(M) * N * 4.0 * (split + 1) / 3000.0;
  return cost;
}

// Warp-private TMA epilogue: the 32 tile rows of TMEM lane quadrant q (linear tile rows 32q .. 32q+31, first box
// dimension fastest) must themselves be a (d1, d2, d3) box. Returns false when the tile's row box does not split so.
static bool sub_box(const asva_gemm_desc* d, int sub[3]) {
  const int b1 = d->box[0], b2 = d->box[1], b3 = d->box[2];
  const int rows = b1 * b2 * b3;
  if (rows <= 32) {
    sub[0] = b1; sub[1] = b2; sub[2] = b3;
    return true;
  }
  if (b1 >= 32) {
    // a single line of rows may be ragged: the rows past it lie past the tensor and the TMA clips them
    const bool one_line = (b2 == 1 && b3 == 1 && b1 >= d->out_dims[0]);
    if (b1 % 32 != 0 && !one_line) return false;
    sub[0] = 32; sub[1] = 1; sub[2] = 1;
    return true;
  }
  if (32 % b1 != 0 || rows % 32 != 0) return false;
  if (b1 * b2 >= 32) {
    if ((b1 * b2) % 32 != 0) return false;
    sub[0] = b1; sub[1] = 32 / b1; sub[2] = 1;
    return true;
  }
  if (32 % (b1 * b2) != 0) return false;
  sub[0] = b1; sub[1] = b2; sub[2] = 32 / (b1 * b2);
  return true;
}

static int env_int(const char* name) {
#ifdef ASVA_DEBUG_SWITCHES
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

static GemmPlan plan_gemm(const asva_gemm_desc* d, int64_t m_tiles, int64_t M, int num_kb, int n_res, int sms) {
  static int env_bn = -1, env_split = -1, env_cg = -1, env_epi = -1;
  if (env_bn < 0) {
    env_bn = env_int("ASVA_GEMM_BN");
    env_split = env_int("ASVA_GEMM_SPLIT");
    env_cg = env_int("ASVA_GEMM_CG");
    env_epi = env_int("ASVA_GEMM_EPI");
  }
  GemmPlan best{128, 1, 2, 1, 1};
  // epilogue form: the per-warp one exists for bf16, non-GEGLU, non-split outputs; explicit request > env > default
  // default (measured, profiles/r1_gemm_probe_v9.md): the per-warp form wins whenever there is a residual to add
  // 3 (warp-private TMA) serves every output type but needs a tile whose quadrants are boxes (sub_box)
  int sub[3];
  const bool sub_ok = sub_box(d, sub);
  int want_epi = d->epilogue ? d->epilogue : (env_epi ? env_epi : (n_res > 0 ? 2 : kDefaultEpi));
  const bool row_stats = d->stats_out != nullptr || d->ln_cols > 0;  // forms 1 and 3 carry the statistics / fold code
  if (want_epi == 2 && row_stats) want_epi = sub_ok ? 3 : 1;
  if (want_epi == 2 && (d->geglu || d->out_fp32)) want_epi = 1;
  if (want_epi == 3 && (!sub_ok || (n_res > 0 && d->out_fp32))) want_epi = 1;
  // 4 = cluster split-K: the splits of a tile are the CTAs of one cluster (2 / 4 / 8), single-CTA tiles, slices of
  // whole 32-column chunks; no GEGLU, no row statistics / fold
  const bool want_cs = want_epi == 4 && !d->geglu && !row_stats;
  if (want_epi == 4 && !want_cs) want_epi = sub_ok ? 3 : 1;
  if (want_epi != 2 && want_epi != 3 && want_epi != 4) want_epi = 1;
  const int bns[4] = {64, 128, 160, 256};
  const int splits[10] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
  const int want_bn = d->geglu ? 128 : (d->block_n ? d->block_n : env_bn);
  const int want_split = (d->geglu || d->ln_cols > 0) ? 1 : (d->split_k ? d->split_k : env_split);
  const int want_cg = d->cta_group ? d->cta_group : env_cg;
  const int64_t max_split = (d->ws != nullptr) ? d->ws_bytes / (M * static_cast<int64_t>(d->N) * 4) : 1;
  double best_cost = 1e300;
  bool pair_ok = true;  // a pair shares ONE W tile: both m tiles must agree on wk vs wk_first (same d2 origin)
  for (int s = 0; s < d->nseg; ++s)
    if (d->seg[s].wk_first >= 0 && ((d->out_dims[0] + d->box[0] - 1) / d->box[0]) % 2 != 0) pair_ok = false;
  for (int cg = 1; cg <= 2; ++cg) {
    if (want_cg && cg != want_cg && !(cg == 1 && !pair_ok)) continue;
    if (cg == 2 && !pair_ok) continue;
    if (want_cs && cg != 1) continue;
    for (int bi = 0; bi < 4; ++bi) {
      const int bn = bns[bi];
      if (want_bn && bn != want_bn) continue;
      if (bn > 64 && bn >= 2 * ((d->N + 31) / 32) * 32) continue;  // more than half the tile would be padding
      for (int si = 0; si < 10; ++si) {
        const int sp = splits[si];
        if (want_cs) {
          if (sp != 2 && sp != 4 && sp != 8) continue;
          if (want_split && sp != want_split) continue;
          if ((bn / sp) % 32 != 0 || sp > num_kb) continue;
        } else {
          if (want_split && sp != want_split && !(want_split > max_split && sp == 1)) continue;
          if (sp > max_split || sp > num_kb) continue;
        }
        const int kbps = (num_kb + sp - 1) / sp;
        if ((num_kb + kbps - 1) / kbps != sp) continue;  // would leave an empty split
        const int nr = (sp > 1 || want_epi == 2 || want_cs) ? 0 : n_res;
        if (stages_for(bn, cg, nr, want_cs ? 0 : (sp > 1 ? 1 : d->out_fp32), (num_kb + sp - 1) / sp) < 2) continue;
        const double c = plan_cost(bn, cg, sp, d->N, m_tiles, num_kb, M, sms);
        if (c < best_cost) {
          best_cost = c;
          best.bn = bn;
          best.split = sp;
          best.cg = cg;
        }
      }
    }
  }
  if (want_cs && best_cost >= 1e300) {  // no cluster split-K plan is feasible: the caller sees a different epilogue back
    asva_gemm_desc c = *d;
    c.epilogue = sub_ok ? 3 : 1;
    return plan_gemm(&c, m_tiles, M, num_kb, n_res, sms);
  }
  best.epi = want_cs ? 4 : (best.split > 1 ? (want_epi == 3 ? 3 : 1) : want_epi);
  const int nr = (best.split > 1 || best.epi == 2) ? 0 : n_res;
  best.stages = stages_for(best.bn, best.cg, nr, best.epi == 4 ? 0 : (best.split > 1 ? 1 : d->out_fp32),
                           (num_kb + best.split - 1) / best.split);
  const int cap = env_int("ASVA_GEMM_STAGES");
  if (cap >= 2 && best.stages > cap) best.stages = cap;
  return best;
}

template <int BN, bool GEGLU, int CG, bool RS>
static int launch_gemm(const GemmKParams& kp, int smem_bytes, cudaStream_t stream) {
  static bool configured_d[kMaxDevices] = {false};  // function attributes and occupancy are per device
  static int max_ctas_d[kMaxDevices] = {0};
  const int dev = current_device();
  const int g_num_sms = device_sms();
  bool& configured = configured_d[dev];
  int& max_ctas = max_ctas_d[dev];
  if (!configured) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, GEGLU, CG, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kSmemLimit));
    max_ctas = g_num_sms;
    if (CG == 2) {  // how many CTA pairs the GPU can hold at once (GPCs with an odd SM count lose one)
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(g_num_sms / 2 * 2, 1, 1);
      q.blockDim = dim3(kGemmThreads, 1, 1);
      q.dynamicSmemBytes = kSmemLimit;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = 2;
      at.val.clusterDim.y = 1;
      at.val.clusterDim.z = 1;
      q.attrs = &at;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, GEGLU, CG, RS>, &q) == cudaSuccess && n > 0)
        max_ctas = 2 * n;
      else
        max_ctas = g_num_sms / 2 * 2;
      (void)cudaGetLastError();
    }
    configured = true;
  }
  if (CG == 1 && kp.csplit > 1) {  // cluster split-K: one tile per cluster, not persistent
    ASVA_CUDA_OK(launch_k(gemm_tc_kernel<BN, GEGLU, CG, RS>, dim3(kp.total_tiles * kp.csplit), dim3(kGemmThreads),
                          smem_bytes, stream, kp.csplit, kp));
    ASVA_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int want = kp.total_tiles * CG;
  const int grid = want < max_ctas ? want : max_ctas;
  ASVA_CUDA_OK(launch_k(gemm_tc_kernel<BN, GEGLU, CG, RS>, dim3(grid), dim3(kGemmThreads), smem_bytes, stream, CG, kp));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int CG, bool RS>
static int dispatch_gemm(const GemmKParams& kp, int bn, bool geglu, int smem, cudaStream_t stream) {
  if (geglu) return launch_gemm<128, true, CG, RS>(kp, smem, stream);
  switch (bn) {
    case 64: return launch_gemm<64, false, CG, RS>(kp, smem, stream);
    case 128: return launch_gemm<128, false, CG, RS>(kp, smem, stream);
    case 160: return launch_gemm<160, false, CG, RS>(kp, smem, stream);
    default: return launch_gemm<256, false, CG, RS>(kp, smem, stream);
  }
}

}  // namespace asva

// The plan asva_gemm would run for this descriptor (explicit block_n / split_k / cta_group requests that are not
// feasible fall back to the cost model's choice, so callers compare the result with what they asked for).
static int compute_plan(const asva_gemm_desc* d, asva::GemmPlan* out) {
  using namespace asva;
  if (d == nullptr || d->nseg < 1 || d->nseg > ASVA_GEMM_MAX_SEG) return fail(ASVA_ERR_INVALID, "bad descriptor");
  const int g_num_sms = device_sms();
  ASVA_REQUIRE(g_num_sms > 0, "asva_gemm: cannot query the device");
  int64_t m_tiles = 1, M = 1;
  for (int i = 0; i < 3; ++i) {
    if (d->box[i] < 1 || d->out_dims[i] < 1) return fail(ASVA_ERR_INVALID, "bad box/out_dims");
    m_tiles *= (d->out_dims[i] + d->box[i] - 1) / d->box[i];
    M *= d->out_dims[i];
  }
  int kb_total = 0, n_res = 0;
  for (int s = 0; s < d->nseg; ++s) kb_total += d->seg[s].num_kb;
  for (int i = 0; i < 2; ++i) n_res += d->res[i] != nullptr;
  *out = plan_gemm(d, m_tiles, M, kb_total, n_res, g_num_sms);
  return 0;
}

extern "C" int asva_gemm_plan(const asva_gemm_desc* d, int32_t* block_n, int32_t* split_k, int32_t* cta_group,
                              int32_t* stages, int32_t* epilogue) {
  asva::GemmPlan pl;
  const int rc = compute_plan(d, &pl);
  if (rc != 0) return rc;
  if (block_n) *block_n = pl.bn;
  if (split_k) *split_k = pl.split;
  if (cta_group) *cta_group = pl.cg;
  if (stages) *stages = pl.stages;
  if (epilogue) *epilogue = pl.epi;
  return 0;
}

extern "C" int asva_gemm_tune(const asva_gemm_desc* d, asva_stream_t stream_, int32_t reps, int32_t* block_n,
                              int32_t* split_k, int32_t* cta_group, int32_t* epilogue, float* best_us) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(d != nullptr && block_n && split_k && cta_group, "asva_gemm_tune: null argument");
  if (reps < 1) reps = 3;
  cudaEvent_t e0, e1;
  ASVA_CUDA_OK(cudaEventCreate(&e0));
  ASVA_CUDA_OK(cudaEventCreate(&e1));
  const int bns[4] = {64, 128, 160, 256};
  const int splits[6] = {1, 2, 3, 4, 6, 8};
  float best = 1e30f;
  int rc = 0, bb = 0, bs = 0, bc = 0, be = 0;
  for (int ce = 0; ce < 2 * 4 && rc == 0; ++ce) {
    const int cg = 1 + ce / 4, epi = 1 + ce % 4;
    if (epi == 4 && cg != 1) continue;
    for (int bi = 0; bi < 4 && rc == 0; ++bi) {
      for (int si = 0; si < 6 && rc == 0; ++si) {
        asva_gemm_desc c = *d;
        c.block_n = d->geglu ? 128 : bns[bi];
        c.split_k = splits[si];
        c.cta_group = cg;
        c.epilogue = epi;
        if (d->geglu && (bi != 1 || si != 0)) continue;
        if (d->ln_cols > 0 && si != 0) continue;  // folded launches do not split K
        GemmPlan pl;
        if (compute_plan(&c, &pl) != 0) continue;
        if (pl.bn != c.block_n || pl.split != c.split_k || pl.cg != c.cta_group || pl.epi != epi) continue;  // not feasible
#ifdef ASVA_DEBUG_SWITCHES
        if (getenv("ASVA_TUNE_LOG")) {
          printf("[tune] cg=%d bn=%d split=%d epi=%d stages=%d\n", cg, c.block_n, c.split_k, epi, pl.stages);
          fflush(stdout);
        }
#endif
        if ((rc = asva_gemm(&c, stream_)) != 0) break;  // warm-up (also first-use kernel attribute setup)
        cudaEventRecord(e0, stream);
        for (int r = 0; r < reps && rc == 0; ++r) rc = asva_gemm(&c, stream_);
        cudaEventRecord(e1, stream);
        if (rc != 0) break;
        if (cudaEventSynchronize(e1) != cudaSuccess) {
          rc = fail(ASVA_ERR_CUDA, "asva_gemm_tune: %s", cudaGetErrorString(cudaGetLastError()));
          break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const float us = ms * 1e3f / reps;
        if (us < best) {
          best = us;
          bb = c.block_n;
          bs = c.split_k;
          bc = cg;
          be = epi;
        }
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != 0) return rc;
  ASVA_REQUIRE(bb != 0, "asva_gemm_tune: no feasible plan");
  *block_n = bb;
  *split_k = bs;
  *cta_group = bc;
  if (epilogue) *epilogue = be;
  if (best_us) *best_us = best;
  return 0;
}

extern "C" int asva_gemm(const asva_gemm_desc* d, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(d != nullptr, "asva_gemm: null descriptor");
  ASVA_REQUIRE(d->a[0] != nullptr && d->w != nullptr && d->out != nullptr, "asva_gemm: null operand");
  ASVA_REQUIRE(d->nseg >= 1 && d->nseg <= ASVA_GEMM_MAX_SEG, "asva_gemm: nseg=%d out of range", d->nseg);
  ASVA_REQUIRE(d->N >= 8 && d->N % 8 == 0, "asva_gemm: N=%d must be a positive multiple of 8", d->N);
  ASVA_REQUIRE(d->K > 0 && d->K % 64 == 0, "asva_gemm: K=%d must be a positive multiple of 64", d->K);
  ASVA_REQUIRE(d->wcols >= 64 && d->ldw >= d->wcols && d->ldw % 8 == 0, "asva_gemm: ldw=%lld / wcols=%d invalid",
               (long long)d->ldw, d->wcols);
  ASVA_REQUIRE(d->ldo % (d->out_fp32 ? 4 : 8) == 0 && d->ldo >= (d->geglu ? d->N / 2 : d->N),
               "asva_gemm: ldo=%lld invalid", (long long)d->ldo);
  const int g_num_sms = device_sms();
  ASVA_REQUIRE(g_num_sms > 0, "asva_gemm: cannot query the device");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  int rows = 1;
  int64_t m_tiles = 1, M = 1;
  for (int i = 0; i < 3; ++i) {
    ASVA_REQUIRE(d->box[i] >= 1 && d->out_dims[i] >= 1 && (d->trav[i] == 1 || d->trav[i] == 2),
                 "asva_gemm: bad box/out_dims/trav at dim %d", i);
    kp.box[i] = d->box[i];
    kp.trav[i] = d->trav[i];
    kp.out_dims[i] = d->out_dims[i];
    kp.tiles[i] = (d->out_dims[i] + d->box[i] - 1) / d->box[i];
    rows *= d->box[i];
    m_tiles *= kp.tiles[i];
    M *= d->out_dims[i];
  }
  ASVA_REQUIRE(rows <= 128, "asva_gemm: tile of %d rows exceeds 128", rows);
  ASVA_REQUIRE(m_tiles <= (1 << 24), "asva_gemm: %lld M tiles", (long long)m_tiles);
  kp.rows_per_tile = rows;
  kp.N = d->N;
  kp.n_out = d->geglu ? d->N / 2 : d->N;

  int kb_total = 0;
  bool uses_src1 = false;
  for (int s = 0; s < d->nseg; ++s) {
    const asva_gemm_seg& g = d->seg[s];
    ASVA_REQUIRE(g.num_kb >= 1 && (g.src == 0 || g.src == 1), "asva_gemm: bad segment %d", s);
    ASVA_REQUIRE(g.c0 >= 0 && g.c0 + 64 * (int64_t)g.num_kb <= d->a_dims[g.src][0],
                 "asva_gemm: segment %d channels [%d, %lld) exceed source extent %lld", s, g.c0,
                 (long long)(g.c0 + 64 * (int64_t)g.num_kb), (long long)d->a_dims[g.src][0]);
    ASVA_REQUIRE(g.wk >= 0 && g.wk % 64 == 0 && g.wk + 64 * g.num_kb <= d->wcols,
                 "asva_gemm: segment %d weight columns [%d, %d) exceed wcols=%d", s, g.wk, g.wk + 64 * g.num_kb,
                 d->wcols);
    ASVA_REQUIRE(g.wk_first < 0 || (g.wk_first % 64 == 0 && g.wk_first + 64 * g.num_kb <= d->wcols),
                 "asva_gemm: segment %d wk_first=%d invalid", s, g.wk_first);
    ASVA_REQUIRE((g.wk_first < 0 && g.fix2 < 0) || d->box[1] == 1,
                 "asva_gemm: segment %d uses wk_first/fix2, which need box[1] == 1", s);
    kp.seg[s] = SegK{g.src, g.c0, g.off[0], g.off[1], g.off[2], g.num_kb, g.wk, g.wk_first, g.fix2};
    kb_total += g.num_kb;
    uses_src1 |= (g.src == 1);
  }
  ASVA_REQUIRE(kb_total * 64 == d->K, "asva_gemm: segments cover K=%d but desc says K=%d", kb_total * 64, d->K);
  ASVA_REQUIRE(!uses_src1 || d->a[1] != nullptr, "asva_gemm: segment references missing source 1");
  kp.num_kb = kb_total;

  int n_res = 0;
  const void* res[2] = {nullptr, nullptr};
  int64_t res_ld[2] = {0, 0};
  for (int i = 0; i < 2; ++i) {
    if (d->res[i] == nullptr) continue;
    ASVA_REQUIRE(d->res_ld[i] % 8 == 0 && d->res_ld[i] >= d->N, "asva_gemm: residual ld must be a multiple of 8");
    res[n_res] = d->res[i];
    res_ld[n_res] = d->res_ld[i];
    ++n_res;
  }
  ASVA_REQUIRE(!d->geglu || (n_res == 0 && d->add.ptr == nullptr && !d->out_fp32 && d->N % 128 == 0),
               "asva_gemm: GEGLU takes bias only, writes bf16 and needs N %% 128 == 0 (N=%d)", d->N);
  ASVA_REQUIRE(d->add.ptr == nullptr || (d->add.ld % 4 == 0 && d->add.div >= 1),
               "asva_gemm: rowadd ld must be a multiple of 4 and div >= 1");

  ASVA_REQUIRE(d->stats_out == nullptr || (!d->geglu && !d->out_fp32 && d->N % 32 == 0),
               "asva_gemm: stats_out needs a bf16, non-GEGLU output with N %% 32 == 0 (N=%d)", d->N);
  ASVA_REQUIRE(d->ln_cols == 0 || (d->ln_cols > 0 && d->ln_cols % 32 == 0 && d->ln_stats != nullptr &&
                                   d->ln_wsum != nullptr && d->ln_stat_rows > 0 && d->ln_grp_rows >= 0),
               "asva_gemm: LayerNorm fold needs ln_stats, ln_wsum, ln_stat_rows and ln_cols %% 32 == 0");

  const GemmPlan plan = plan_gemm(d, m_tiles, M, kb_total, n_res, g_num_sms);
  const int bn = plan.bn;
  const bool csplit = plan.epi == 4;             // K split over the CTAs of a cluster: no workspace, no second launch
  const bool split = plan.split > 1 && !csplit;  // K split through the fp32 workspace + reduce kernel
  kp.csplit = csplit ? plan.split : 0;
  kp.split_k = plan.split;
  kp.kb_per_split = (kb_total + plan.split - 1) / plan.split;
  kp.n_res = split ? 0 : n_res;
  kp.out_fp32 = split ? 1 : (csplit ? 0 : d->out_fp32);
  kp.csplit_f32 = (csplit && d->out_fp32) ? 1 : 0;
  kp.n_stages = plan.stages;
  kp.epi = plan.epi;
  const bool direct = plan.epi == 2 || csplit;  // plain pointers for residuals and output
  ASVA_REQUIRE(!csplit || (plan.cg == 1 && (plan.split == 2 || plan.split == 4 || plan.split == 8) &&
                           (bn / plan.split) % 32 == 0),
               "asva_gemm: cluster split-K needs cta_group 1, 2 / 4 / 8 splits and slices of whole 32-column chunks");
  kp.n_res_slots = direct ? 0 : res_slots_for(bn, plan.cg, kp.n_res, kp.out_fp32, kp.kb_per_split);
  if (direct) {
    kp.out_ptr = reinterpret_cast<__nv_bfloat16*>(d->out);
    kp.ldo = d->ldo;
    for (int i = 0; i < n_res; ++i) {
      kp.res_ptr[i] = reinterpret_cast<const __nv_bfloat16*>(res[i]);
      kp.res_ld[i] = res_ld[i];
    }
  }
  kp.dbg = env_int("ASVA_GEMM_DBG");
  {
    int pf = 64;  // W blocks each CTA prefetches into L2 before the dependent-launch wait
#ifdef ASVA_DEBUG_SWITCHES
    if (const char* e = getenv("ASVA_GEMM_PF")) pf = atoi(e);
#endif
    kp.pf_blocks = pf;
    kp.w_kblocks = d->wcols / 64;
  }
  ASVA_REQUIRE(plan.stages >= 2, "asva_gemm: no shared memory left for a pipeline (block_n=%d)", bn);
  ASVA_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256, "asva_gemm: unsupported block_n=%d", bn);
  ASVA_REQUIRE(plan.cg == 1 || plan.cg == 2, "asva_gemm: cta_group must be 0 (auto), 1 or 2");
  if (!split) {
    kp.bias = d->bias;
    kp.add_ptr = d->add.ptr;
    kp.add_ld = d->add.ld;
    kp.add_div = d->add.div > 0 ? d->add.div : 1;
    kp.stats_out = d->stats_out;  // (a split launch leaves them to its reduce kernel)
    kp.stats_rows = M;
  }
  if (d->ln_cols > 0) {
    ASVA_REQUIRE(!split && plan.epi != 2, "asva_gemm: LayerNorm fold with split-K / epilogue 2");
    kp.ln_stats = d->ln_stats;
    kp.ln_wsum = d->ln_wsum;
    kp.ln_stat_rows = d->ln_stat_rows;
    kp.ln_slots = d->ln_cols / 32;
    kp.ln_grp_rows = d->ln_grp_rows;
    kp.ln_grp_stride = d->ln_grp_stride;
    kp.ln_inv_c = 1.0f / static_cast<float>(d->ln_cols);
    kp.ln_eps = d->ln_eps;
  }
  ASVA_REQUIRE(d->stats_out == nullptr || plan.epi != 2, "asva_gemm: stats_out with epilogue 2");

  for (int src = 0; src < 2; ++src) {
    if (d->a[src] == nullptr) continue;
    uint64_t dims[4], strides[3];
    uint32_t box[4], el[4];
    dims[0] = (uint64_t)d->a_dims[src][0];
    box[0] = 64;
    el[0] = 1;
    for (int i = 0; i < 3; ++i) {
      dims[i + 1] = (uint64_t)d->a_dims[src][i + 1];
      strides[i] = (uint64_t)d->a_strides[src][i] * 2u;
      box[i + 1] = (uint32_t)(d->box[i] * d->trav[i]);
      el[i + 1] = (uint32_t)d->trav[i];
    }
    int rc = make_tmap_bf16(src == 0 ? &kp.tmA0 : &kp.tmA1, d->a[src], 4, dims, strides, box, el);
    if (rc != 0) return rc;
  }
  if (d->a[1] == nullptr) kp.tmA1 = kp.tmA0;
  {
    uint64_t dims[2] = {(uint64_t)d->wcols, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)d->ldw * 2u};
    uint32_t box[2] = {64u, (uint32_t)(bn / plan.cg)};
    uint32_t el[2] = {1u, 1u};
    int rc = make_tmap_bf16(&kp.tmW, d->w, 2, dims, strides, box, el);
    if (rc != 0) return rc;
  }
  // output (or split-K scratch) and residuals: (cols, d1, d2, d3[, split]) boxes of 32 columns x the row box
  uint32_t rbox[5] = {32u, (uint32_t)d->box[0], (uint32_t)d->box[1], (uint32_t)d->box[2], 1u};
  uint32_t obox[5] = {32u, (uint32_t)d->box[0], (uint32_t)d->box[1], (uint32_t)d->box[2], 1u};
  kp.sub_rows = rows < 32 ? rows : 32;
  const bool wide = plan.epi == 3 && !kp.out_fp32 && !d->geglu;  // bf16 panels of 64 columns
  if (plan.epi == 3) {  // every epilogue warp stores its own quadrant of a panel
    int sub[3];
    ASVA_REQUIRE(sub_box(d, sub), "asva_gemm: epilogue 3 needs a tile whose 32-row quadrants are boxes");
    for (int i = 0; i < 3; ++i) obox[i + 1] = (uint32_t)sub[i];
    if (wide) obox[0] = rbox[0] = 64u;  // residual panels keep the whole-tile row box (one load per panel)
  }
  const uint32_t rel[5] = {1u, 1u, 1u, 1u, 1u};
  {
    const bool f32 = kp.out_fp32 != 0;
    const uint64_t es = f32 ? 4u : 2u;
    const uint64_t ld = split ? (uint64_t)d->N : (uint64_t)d->ldo;
    void* base = split ? d->ws : d->out;
    uint64_t dims[5] = {(uint64_t)kp.n_out, (uint64_t)d->out_dims[0], (uint64_t)d->out_dims[1],
                        (uint64_t)d->out_dims[2], (uint64_t)(csplit ? 1 : plan.split)};
    uint64_t strides[4] = {ld * es, ld * es * dims[1], ld * es * dims[1] * dims[2],
                           ld * es * dims[1] * dims[2] * dims[3]};
    int rc = make_tmap(&kp.tmO, base, f32 ? TMAP_F32 : TMAP_BF16, (f32 || wide) ? TMAP_SW128 : TMAP_SW64, 5, dims,
                       strides, obox, rel);
    if (rc != 0) return rc;
    kp.tmO2 = kp.tmO;
    if (wide) {
      obox[0] = 32u;
      rc = make_tmap(&kp.tmO2, base, TMAP_BF16, TMAP_SW64, 5, dims, strides, obox, rel);
      if (rc != 0) return rc;
    }
  }
  for (int i = 0; i < (direct ? 0 : kp.n_res); ++i) {
    uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->out_dims[0], (uint64_t)d->out_dims[1], (uint64_t)d->out_dims[2]};
    const uint64_t ld = (uint64_t)res_ld[i] * 2u;
    uint64_t strides[3] = {ld, ld * dims[1], ld * dims[1] * dims[2]};
    int rc = make_tmap(i == 0 ? &kp.tmR0 : &kp.tmR1, res[i], TMAP_BF16, wide ? TMAP_SW128 : TMAP_SW64, 4, dims, strides,
                       rbox, rel);
    if (rc != 0) return rc;
  }
  if (kp.n_res < 2) kp.tmR1 = kp.tmR0;
  if (kp.n_res < 1 || direct) kp.tmR0 = kp.tmR1 = kp.tmO;

#ifdef ASVA_DEBUG_SWITCHES
  static long long* trace_buf = nullptr;
  static int trace_on = -1;
  if (trace_on < 0) {
    const char* e = getenv("ASVA_GEMM_TRACE");
    trace_on = (e != nullptr && e[0] == '1') ? 1 : 0;
    if (trace_on) cudaMalloc(&trace_buf, 12 * 64 * sizeof(long long));
  }
  if (trace_on) {
    cudaMemsetAsync(trace_buf, 0, 12 * 64 * sizeof(long long), stream);
    kp.trace = trace_buf;
  }
#endif
  kp.n_tiles_n = (d->N + bn - 1) / bn;
  ASVA_REQUIRE(m_tiles * kp.n_tiles_n * plan.split < (1ll << 31), "asva_gemm: too many tiles");
  kp.mn_tiles = (int)(((m_tiles + plan.cg - 1) / plan.cg) * kp.n_tiles_n);  // pairs of m tiles when cg == 2
  kp.total_tiles = csplit ? kp.mn_tiles : kp.mn_tiles * plan.split;
  const int smem = smem_for(bn, plan.cg, plan.stages, kp.n_res_slots, kp.out_fp32);
  const bool rowstat = kp.stats_out != nullptr || kp.ln_stats != nullptr;
  const int rc = plan.cg == 2 ? (rowstat ? dispatch_gemm<2, true>(kp, bn, d->geglu != 0, smem, stream)
                                         : dispatch_gemm<2, false>(kp, bn, d->geglu != 0, smem, stream))
                              : (rowstat ? dispatch_gemm<1, true>(kp, bn, d->geglu != 0, smem, stream)
                                         : dispatch_gemm<1, false>(kp, bn, d->geglu != 0, smem, stream));
#ifdef ASVA_DEBUG_SWITCHES
  if (trace_on && rc == 0) {
    static long long h[12 * 64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (int w = 0; w < 12; ++w)
      if (h[w * 64] != 0 && (t0 == 0 || h[w * 64] < t0)) t0 = h[w * 64];
    printf("[asva gemm trace] M tiles %lld N=%d K=%d bn=%d cg=%d epi=%d stages=%d (ns since CTA 0 entry)\n",
           (long long)m_tiles, d->N, d->K, bn, plan.cg, plan.epi, plan.stages);
    for (int w = 0; w < 12; ++w) {
      printf("  warp %2d:", w);
      for (int e = 0; e < 64; ++e)
        if (h[w * 64 + e] != 0) printf(" e%d=%lld", e, h[w * 64 + e] - t0);
      printf("\n");
    }
  }
#endif
  if (rc != 0 || !split) return rc;
  const int64_t chunks = M * (d->N / 8);
  int64_t blocks = (chunks + 255) / 256;
  if (blocks > g_num_sms * 8) blocks = g_num_sms * 8;
  ASVA_CUDA_OK(launch_k(splitk_finalize_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 1,
                        reinterpret_cast<const float*>(d->ws), plan.split, M, d->N, d->bias, d->add.ptr, d->add.ld,
                        d->add.div > 0 ? d->add.div : 1, reinterpret_cast<const __nv_bfloat16*>(res[0]), res_ld[0],
                        reinterpret_cast<const __nv_bfloat16*>(res[1]), res_ld[1], d->out, d->ldo, d->out_fp32,
                        d->stats_out));
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}
