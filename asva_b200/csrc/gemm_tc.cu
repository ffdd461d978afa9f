// tcgen05 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] * W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Persistent kernel: one CTA per SM walks output tiles (128 x BN) round-robin. Warp roles (192 threads):
//   warp 0      TMA producer: per 64-wide K block one 4-D box load of A (table-driven: implicit-GEMM conv taps,
//               temporal taps, concat sources) and one 2-D box load of W into a 128B-swizzled smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per K block); the accumulator is double
//               buffered in TMEM (2 x BN columns), so tile t+1's main loop runs under tile t's epilogue
//   warps 2..5  epilogue, two phases per 64-column panel:
//                 1. tcgen05.ld (lane = output row) -> fp32 staging tile in smem (odd 16-byte pitch: conflict free)
//                 2. threads re-map to (row, 8-column chunk) so that consecutive lanes touch consecutive 16 bytes of
//                    one output row: bias / broadcast adds / residual loads and the bf16|fp32 stores are coalesced
#include "common.cuh"
#include "host_common.h"
#include <string.h>

namespace asva {

struct SegK {
  int32_t src, c0, off1, off2, off3, num_kb;
};
struct RowAddK {
  const float* ptr;
  int64_t ld;
  int32_t div_outer, mul_outer, mod_inner, sel_lt, sel_off;
};

struct GemmKParams {
  CUtensorMap tmA0, tmA1, tmW;
  SegK seg[ASVA_GEMM_MAX_SEG];
  int32_t box[3], trav[3], out_dims[3], tiles[3];
  int32_t rows_per_tile, N, num_kb, n_tiles_n, total_tiles;
  const float* bias;
  RowAddK add[2];
  const __nv_bfloat16* res[2];
  int64_t res_ld[2];
  void* out;
  int64_t row_s1, row_s0, col_s1;
  int32_t row_div, col_div, out_fp32;
};

struct RowInfo {  // per output row of the current tile, written by the thread that owns the TMEM lane
  int64_t row;      // global output row, -1 = outside the problem
  int64_t out_off;  // element offset of the row in `out`
  int64_t add_off[2];
};

template <int BN, bool GEGLU>
struct GemmCfg {
  static constexpr int kStages = (BN <= 64) ? 6 : 4;
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutCols = GEGLU ? 64 : BN;        // output columns per tile
  static constexpr int kPanel = 64;                        // columns staged at a time
  static constexpr int kPitch = kPanel * 4 + 16;           // bytes per staged row (odd multiple of 16)
  static constexpr int kStagingBytes = 128 * kPitch;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 128 * (int)sizeof(RowInfo) +
                                    1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int BN, bool GEGLU>
__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN, GEGLU>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;
  RowInfo* rowinfo = reinterpret_cast<RowInfo*>(staging + Cfg::kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(rowinfo + 128);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmW);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = static_cast<uint32_t>(p.rows_per_tile) * 128u + Cfg::kBBytes;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles_n) * BN;
        const int mt = tile / p.n_tiles_n;
        const int t1 = mt % p.tiles[0];
        const int t2 = (mt / p.tiles[0]) % p.tiles[1];
        const int t3 = mt / (p.tiles[0] * p.tiles[1]);
        const int i1 = t1 * p.box[0] * p.trav[0], i2 = t2 * p.box[1] * p.trav[1], i3 = t3 * p.box[2] * p.trav[2];
        int seg = 0, kin = 0;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          const SegK sg = p.seg[seg];
          tma_load_4d(sa, sg.src ? &p.tmA1 : &p.tmA0, &full_bar[s], sg.c0 + kin * 64, i1 + sg.off1, i2 + sg.off2,
                      i3 + sg.off3);
          tma_load_2d(sa + Cfg::kABytes, &p.tmW, &full_bar[s], kb * 64, n0);
          if (++kin == sg.num_kb) {
            kin = 0;
            ++seg;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = make_sdesc_sw128(sa);
          const uint64_t bdesc = make_sdesc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tmem_full_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;  // TMEM lane == row inside the tile; also the linear thread id of phase 2
    const int r1 = r % p.box[0];
    const int r2 = (r / p.box[0]) % p.box[1];
    const int r3 = r / (p.box[0] * p.box[1]);
    uint8_t* my_stage = staging + r * Cfg::kPitch;
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1u, acc_ph = (t >> 1) & 1u;
      const int n0 = (tile % p.n_tiles_n) * BN;
      const int mt = tile / p.n_tiles_n;
      const int o1 = (mt % p.tiles[0]) * p.box[0];
      const int o2 = ((mt / p.tiles[0]) % p.tiles[1]) * p.box[1];
      const int o3 = (mt / (p.tiles[0] * p.tiles[1])) * p.box[2];
      {
        const bool valid = (r < p.rows_per_tile) && (o1 + r1 < p.out_dims[0]) && (o2 + r2 < p.out_dims[1]) &&
                           (o3 + r3 < p.out_dims[2]);
        RowInfo ri;
        ri.row = -1;
        ri.out_off = 0;
        ri.add_off[0] = ri.add_off[1] = 0;
        if (valid) {
          const int64_t row =
              (static_cast<int64_t>(o3 + r3) * p.out_dims[1] + (o2 + r2)) * p.out_dims[0] + (o1 + r1);
          ri.row = row;
          ri.out_off = (row / p.row_div) * p.row_s1 + (row % p.row_div) * p.row_s0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (p.add[i].ptr != nullptr) {
              const RowAddK& a = p.add[i];
              const int64_t arow = (row / a.div_outer) * a.mul_outer + (row % a.mod_inner);
              ri.add_off[i] = arow * a.ld + (((row % a.div_outer) < a.sel_lt) ? a.sel_off : 0);
            }
          }
        }
        rowinfo[r] = ri;  // published by the first epi_bar_sync below
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);

#pragma unroll 1
      for (int pc = 0; pc < Cfg::kOutCols; pc += Cfg::kPanel) {
        const int pw = (Cfg::kOutCols - pc) < Cfg::kPanel ? (Cfg::kOutCols - pc) : Cfg::kPanel;  // 64 or 32
        // ---- phase 1: accumulator -> fp32 staging (thread = row)
#pragma unroll 1
        for (int c = 0; c < pw; c += 32) {
          uint32_t v[32];
          if constexpr (!GEGLU) {
            tmem_ld_x32(taddr + pc + c, v);
            tmem_ld_wait();
          } else {
            // tile columns [0,64) = value h, [64,128) = gate g; staged value = (h + bh) * gelu(g + bg)
            uint32_t gv[32];
            tmem_ld_x32(taddr + pc + c, v);
            tmem_ld_x32(taddr + 64 + pc + c, gv);
            tmem_ld_wait();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float bh[8], bg[8];
              const int tc = pc + c + g8 * 8;
              if (p.bias != nullptr && n0 + 64 + tc + 8 <= p.N) {
                load8_f32(p.bias + n0 + tc, bh);
                load8_f32(p.bias + n0 + 64 + tc, bg);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) bh[j] = bg[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float h = __uint_as_float(v[g8 * 8 + j]) + bh[j];
                const float gg = __uint_as_float(gv[g8 * 8 + j]) + bg[j];
                v[g8 * 8 + j] = __float_as_uint(h * gelu_erf_f(gg));
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(my_stage + (c + j * 4) * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (pc + Cfg::kPanel >= Cfg::kOutCols) {  // whole accumulator read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        epi_bar_sync();
        // ---- phase 2: (row, 8-column chunk) per thread, coalesced global access
        const int cpp_shift = (pw == 64) ? 3 : 2;  // chunks per row in this panel: 8 or 4
        const int nchunks = 128 << cpp_shift;
        constexpr int U = 4;
#pragma unroll 1
        for (int id0 = r; id0 < nchunks; id0 += 128 * U) {
          float v[U][8];
          bool live[U];
          int64_t ooff[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int id = id0 + u * 128;
            const int row_l = id >> cpp_shift;
            const int cc = id & ((1 << cpp_shift) - 1);
            const RowInfo ri = rowinfo[row_l];
            const int ocol = (GEGLU ? (n0 >> 1) : n0) + pc + cc * 8;  // output column
            const int ncol = GEGLU ? (n0 + 64 + pc + cc * 8) : ocol;    // last accumulator column this chunk needs
            live[u] = (id < nchunks) && (ri.row >= 0) && (ncol < p.N);
            if (!live[u]) continue;
            const float4* sp = reinterpret_cast<const float4*>(staging + row_l * Cfg::kPitch + cc * 32);
            const float4 x0 = sp[0], x1 = sp[1];
            v[u][0] = x0.x; v[u][1] = x0.y; v[u][2] = x0.z; v[u][3] = x0.w;
            v[u][4] = x1.x; v[u][5] = x1.y; v[u][6] = x1.z; v[u][7] = x1.w;
            ooff[u] = ri.out_off + static_cast<int64_t>(ocol / p.col_div) * p.col_s1 + (ocol % p.col_div);
            if constexpr (!GEGLU) {
              if (p.bias != nullptr) {
                float b[8];
                load8_f32(p.bias + ocol, b);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] += b[j];
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (p.add[i].ptr != nullptr) {
                  float b[8];
                  load8_f32(p.add[i].ptr + ri.add_off[i] + ocol, b);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[u][j] += b[j];
                }
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (p.res[i] != nullptr) {
                  const uint4 w = *reinterpret_cast<const uint4*>(p.res[i] + ri.row * p.res_ld[i] + ocol);
                  const float2 f0 = unpack_bf16x2(w.x), f1 = unpack_bf16x2(w.y), f2 = unpack_bf16x2(w.z),
                               f3 = unpack_bf16x2(w.w);
                  v[u][0] += f0.x; v[u][1] += f0.y; v[u][2] += f1.x; v[u][3] += f1.y;
                  v[u][4] += f2.x; v[u][5] += f2.y; v[u][6] += f3.x; v[u][7] += f3.y;
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            if (p.out_fp32) {
              float* o = reinterpret_cast<float*>(p.out) + ooff[u];
              *reinterpret_cast<float4*>(o) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
            } else {
              uint4 w;
              w.x = pack_bf16x2(v[u][0], v[u][1]);
              w.y = pack_bf16x2(v[u][2], v[u][3]);
              w.z = pack_bf16x2(v[u][4], v[u][5]);
              w.w = pack_bf16x2(v[u][6], v[u][7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ooff[u]) = w;
            }
          }
        }
        epi_bar_sync();  // staging (and, after the last panel, rowinfo) may be overwritten
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

static int g_num_sms = 0;

template <int BN, bool GEGLU>
static int launch_gemm(const GemmKParams& kp, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, GEGLU>;
  static bool configured = false;
  if (!configured) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    ASVA_CUDA_OK(cudaGetDevice(&dev));
    ASVA_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = kp.total_tiles < g_num_sms ? kp.total_tiles : g_num_sms;
  gemm_tc_kernel<BN, GEGLU><<<grid, 192, Cfg::kSmemBytes, stream>>>(kp);
  ASVA_CUDA_OK(cudaGetLastError());
  return 0;
}

static int pick_block_n(int N, int64_t m_tiles) {
  // prefer exact tilings; small grids take narrower tiles to put more CTAs in flight
  if (N % 128 == 0) {
    if (m_tiles * (N / 128) < 148 && N % 64 == 0) return 64;
    return 128;
  }
  if (N % 160 == 0) return 160;
  if (N % 64 == 0) return 64;
  if (N <= 64) return 64;
  return 128;
}

}  // namespace asva

extern "C" int asva_gemm(const asva_gemm_desc* d, asva_stream_t stream_) {
  using namespace asva;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ASVA_REQUIRE(d != nullptr, "asva_gemm: null descriptor");
  ASVA_REQUIRE(d->a[0] != nullptr && d->w != nullptr && d->out != nullptr, "asva_gemm: null operand");
  ASVA_REQUIRE(d->nseg >= 1 && d->nseg <= ASVA_GEMM_MAX_SEG, "asva_gemm: nseg=%d out of range", d->nseg);
  ASVA_REQUIRE(d->N >= 8 && d->N % 8 == 0, "asva_gemm: N=%d must be a positive multiple of 8", d->N);
  ASVA_REQUIRE(d->K > 0 && d->K % 64 == 0, "asva_gemm: K=%d must be a positive multiple of 64", d->K);
  ASVA_REQUIRE(d->ldw >= d->K && d->ldw % 8 == 0, "asva_gemm: ldw=%lld invalid", (long long)d->ldw);
  ASVA_REQUIRE(d->row_div > 0 && d->col_div > 0 && d->col_div % 8 == 0, "asva_gemm: bad output addressing");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  int rows = 1;
  int64_t m_tiles = 1;
  for (int i = 0; i < 3; ++i) {
    ASVA_REQUIRE(d->box[i] >= 1 && d->out_dims[i] >= 1 && (d->trav[i] == 1 || d->trav[i] == 2),
                 "asva_gemm: bad box/out_dims/trav at dim %d", i);
    kp.box[i] = d->box[i];
    kp.trav[i] = d->trav[i];
    kp.out_dims[i] = d->out_dims[i];
    kp.tiles[i] = (d->out_dims[i] + d->box[i] - 1) / d->box[i];
    rows *= d->box[i];
    m_tiles *= kp.tiles[i];
  }
  ASVA_REQUIRE(rows <= 128, "asva_gemm: tile of %d rows exceeds 128", rows);
  ASVA_REQUIRE(m_tiles <= (1 << 24), "asva_gemm: %lld M tiles", (long long)m_tiles);
  kp.rows_per_tile = rows;
  kp.N = d->N;

  int kb_total = 0;
  bool uses_src1 = false;
  for (int s = 0; s < d->nseg; ++s) {
    const asva_gemm_seg& g = d->seg[s];
    ASVA_REQUIRE(g.num_kb >= 1 && (g.src == 0 || g.src == 1), "asva_gemm: bad segment %d", s);
    ASVA_REQUIRE(g.c0 >= 0 && g.c0 + 64 * (int64_t)g.num_kb <= d->a_dims[g.src][0],
                 "asva_gemm: segment %d channels [%d, %lld) exceed source extent %lld", s, g.c0,
                 (long long)(g.c0 + 64 * (int64_t)g.num_kb), (long long)d->a_dims[g.src][0]);
    kp.seg[s] = SegK{g.src, g.c0, g.off[0], g.off[1], g.off[2], g.num_kb};
    kb_total += g.num_kb;
    uses_src1 |= (g.src == 1);
  }
  ASVA_REQUIRE(kb_total * 64 == d->K, "asva_gemm: segments cover K=%d but desc says K=%d", kb_total * 64, d->K);
  ASVA_REQUIRE(!uses_src1 || d->a[1] != nullptr, "asva_gemm: segment references missing source 1");
  kp.num_kb = kb_total;

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

  int bn = d->block_n;
  if (d->geglu) {
    ASVA_REQUIRE(d->N % 128 == 0, "asva_gemm: GEGLU needs N %% 128 == 0 (N=%d)", d->N);
    ASVA_REQUIRE(!d->out_fp32, "asva_gemm: GEGLU writes bf16");
    bn = 128;
  } else if (bn == 0) {
    bn = pick_block_n(d->N, m_tiles);
  }
  ASVA_REQUIRE(bn == 64 || bn == 128 || bn == 160, "asva_gemm: unsupported block_n=%d", bn);
  {
    uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)d->ldw * 2u};
    uint32_t box[2] = {64u, (uint32_t)bn};
    uint32_t el[2] = {1u, 1u};
    int rc = make_tmap_bf16(&kp.tmW, d->w, 2, dims, strides, box, el);
    if (rc != 0) return rc;
  }

  kp.bias = d->bias;
  for (int i = 0; i < 2; ++i) {
    kp.add[i].ptr = d->add[i].ptr;
    kp.add[i].ld = d->add[i].ld;
    kp.add[i].div_outer = d->add[i].div_outer > 0 ? d->add[i].div_outer : 1;
    kp.add[i].mul_outer = d->add[i].mul_outer;
    kp.add[i].mod_inner = d->add[i].mod_inner > 0 ? d->add[i].mod_inner : 1;
    kp.add[i].sel_lt = d->add[i].sel_lt;
    kp.add[i].sel_off = d->add[i].sel_off;
    kp.res[i] = reinterpret_cast<const __nv_bfloat16*>(d->res[i]);
    kp.res_ld[i] = d->res_ld[i];
    ASVA_REQUIRE(d->res[i] == nullptr || d->res_ld[i] % 8 == 0, "asva_gemm: residual ld must be a multiple of 8");
    ASVA_REQUIRE(d->add[i].ptr == nullptr || (d->add[i].ld % 4 == 0 && d->add[i].sel_off % 4 == 0),
                 "asva_gemm: rowadd ld/sel_off must be multiples of 4");
  }
  kp.out = d->out;
  kp.row_s1 = d->row_s1;
  kp.row_s0 = d->row_s0;
  kp.col_s1 = d->col_s1;
  kp.row_div = d->row_div;
  kp.col_div = d->col_div;
  kp.out_fp32 = d->out_fp32;

  kp.n_tiles_n = (d->N + bn - 1) / bn;
  ASVA_REQUIRE(m_tiles * kp.n_tiles_n < (1ll << 31), "asva_gemm: too many tiles");
  kp.total_tiles = (int)(m_tiles * kp.n_tiles_n);
  if (d->geglu) return launch_gemm<128, true>(kp, stream);
  switch (bn) {
    case 64: return launch_gemm<64, false>(kp, stream);
    case 128: return launch_gemm<128, false>(kp, stream);
    default: return launch_gemm<160, false>(kp, stream);
  }
}
