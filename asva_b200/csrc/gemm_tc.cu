// tcgen05 GEMM for sm_100a:  out[M,N] = epilogue( A[M,K] * W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA computes one 128 x BN output tile. Warp roles (192 threads):
//   warp 0      TMA producer: per 64-wide K block one 4-D box load of A (table-driven: implicit-GEMM conv taps,
//               temporal taps, concat sources) and one 2-D box load of W into a 128B-swizzled smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per K block), commits free the ring
//   warps 2..5  epilogue: tcgen05.ld the accumulator (lane = output row), fused bias / broadcast adds /
//               residuals / GEGLU, bf16 or fp32 stores with two-level row/column addressing
// Two CTAs are resident per SM (<= 110 KB smem, <= 256 TMEM columns each) so one tile's epilogue overlaps the
// other's main loop.
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
  int32_t rows_per_tile, N, num_kb;
  const float* bias;
  RowAddK add[2];
  const __nv_bfloat16* res[2];
  int64_t res_ld[2];
  void* out;
  int64_t row_s1, row_s0, col_s1;
  int32_t row_div, col_div, out_fp32;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN <= 64) ? 4 : 3;
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 128 /*barriers*/;
};

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <int BN, bool GEGLU>
__global__ void __launch_bounds__(192, 2) gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;
  const int t1 = mt % p.tiles[0];
  const int t2 = (mt / p.tiles[0]) % p.tiles[1];
  const int t3 = mt / (p.tiles[0] * p.tiles[1]);
  const int o1 = t1 * p.box[0], o2 = t2 * p.box[1], o3 = t3 * p.box[2];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
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
      const int i1 = o1 * p.trav[0], i2 = o2 * p.trav[1], i3 = o3 * p.trav[2];
      const uint32_t tx_bytes = static_cast<uint32_t>(p.rows_per_tile) * 128u + Cfg::kBBytes;
      int seg = 0, kin = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
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
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint64_t adesc = make_sdesc_sw128(sa);
        const uint64_t bdesc = make_sdesc_sw128(sa + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int r1 = r % p.box[0];
    const int r2 = (r / p.box[0]) % p.box[1];
    const int r3 = r / (p.box[0] * p.box[1]);
    const bool valid = (r < p.rows_per_tile) && (o1 + r1 < p.out_dims[0]) && (o2 + r2 < p.out_dims[1]) &&
                       (o3 + r3 < p.out_dims[2]);
    const int64_t row = (static_cast<int64_t>(o3 + r3) * p.out_dims[1] + (o2 + r2)) * p.out_dims[0] + (o1 + r1);
    const int64_t out_row_off = valid ? (row / p.row_div) * p.row_s1 + (row % p.row_div) * p.row_s0 : 0;
    const float* addp[2] = {nullptr, nullptr};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (valid && p.add[i].ptr != nullptr) {
        const RowAddK& a = p.add[i];
        const int64_t arow = (row / a.div_outer) * a.mul_outer + (row % a.mod_inner);
        const int sel = ((row % a.div_outer) < a.sel_lt) ? a.sel_off : 0;
        addp[i] = a.ptr + arow * a.ld + sel;
      }
    }
    const __nv_bfloat16* resp[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) resp[i] = (valid && p.res[i] != nullptr) ? p.res[i] + row * p.res_ld[i] : nullptr;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if constexpr (!GEGLU) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t acc[32];
        tmem_ld_x32(taddr + c, acc);
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = n0 + c + g * 8;
          if (col >= p.N) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
          if (p.bias != nullptr) {
            float b[8];
            load8_f32(p.bias + col, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += b[j];
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (addp[i] != nullptr) {
              float b[8];
              load8_f32(addp[i] + col, b);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += b[j];
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (resp[i] != nullptr) {
              const uint4 u = *reinterpret_cast<const uint4*>(resp[i] + col);
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z),
                           f3 = unpack_bf16x2(u.w);
              v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
              v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
            }
          }
          const int64_t off = out_row_off + static_cast<int64_t>(col / p.col_div) * p.col_s1 + (col % p.col_div);
          if (p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + off;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]);
            u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]);
            u.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = u;
          }
        }
      }
    } else {
      // GEGLU: tile columns [0,64) = value h, [64,128) = gate g; out[:, n0/2 + j] = h_j * gelu(g_j)
      static_assert(!GEGLU || BN == 128, "GEGLU epilogue needs BN == 128");
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t hv[32], gv[32];
        tmem_ld_x32(taddr + c, hv);
        tmem_ld_x32(taddr + 64 + c, gv);
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int tc = c + g * 8;  // column inside the value half
          if (n0 + tc >= p.N) break;
          float bh[8], bg[8];
          if (p.bias != nullptr) {
            load8_f32(p.bias + n0 + tc, bh);
            load8_f32(p.bias + n0 + 64 + tc, bg);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bh[j] = bg[j] = 0.f;
          }
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float h = __uint_as_float(hv[g * 8 + j]) + bh[j];
            const float gg = __uint_as_float(gv[g * 8 + j]) + bg[j];
            v[j] = h * gelu_erf_f(gg);
          }
          const int col = (n0 >> 1) + tc;
          const int64_t off = out_row_off + static_cast<int64_t>(col / p.col_div) * p.col_s1 + (col % p.col_div);
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]);
          u.y = pack_bf16x2(v[2], v[3]);
          u.z = pack_bf16x2(v[4], v[5]);
          u.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BN, bool GEGLU>
static int launch_gemm(const GemmKParams& kp, int n_tiles, int m_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    ASVA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
    configured = true;
  }
  dim3 grid(n_tiles, m_tiles, 1);
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
  ASVA_REQUIRE(m_tiles <= 65535, "asva_gemm: %lld M tiles exceed grid.y", (long long)m_tiles);
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

  const int n_tiles = (d->N + bn - 1) / bn;
  if (d->geglu) return launch_gemm<128, true>(kp, n_tiles, (int)m_tiles, stream);
  switch (bn) {
    case 64: return launch_gemm<64, false>(kp, n_tiles, (int)m_tiles, stream);
    case 128: return launch_gemm<128, false>(kp, n_tiles, (int)m_tiles, stream);
    default: return launch_gemm<160, false>(kp, n_tiles, (int)m_tiles, stream);
  }
}
