// Hardware characterisation for the GEMM main loop (B200): how fast can ONE thread issue tcgen05.mma / commit, what
// does an mbarrier ping-pong between a producer and a consumer thread cost, and what does the TMA deliver per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I asva_b200/csrc -o ubench tools/ubench/ubench.cu -lcuda
// Prints cycles per K block (64-wide) for each experiment; numbers go to profiles/.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace asva;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int kIters = 2000;

// ---- 1. MMA issue rate: one thread, 4 x (128 x N x 16) per K block over resident smem operands, commit per block
template <int N, bool COMMIT>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, done;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t adesc = make_sdesc_sw128(smem_u32(smem));
    const uint64_t bdesc = make_sdesc_sw128(smem_u32(smem) + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
      if (COMMIT) tc_commit(&bar);
    }
    tc_commit(&done);  // completes after every MMA issued above
    const long long t1 = clock64();
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// ---- 2. producer/consumer ping-pong over a ring of S stages, no TMA, no MMA.
//      mode bit0: consumer releases with tcgen05.commit (else plain mbarrier.arrive); bit1: tcgen05.fence after wait
template <int S>
__global__ void __launch_bounds__(128, 1) pingpong_kernel(long long* out, int mode) {
  __shared__ uint64_t full[S], empty[S];
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
  if (warp == 1 && lane == 0) {
    uint32_t s = 0, ph = 1;
    for (int it = 0; it < kIters; ++it) {
      mbar_wait_a(empty0 + 8 * s, ph);
      mbar_arrive_expect_tx_a(full0 + 8 * s, 0);
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && lane == 0) {
    uint32_t s = 0, ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
      mbar_wait_a(full0 + 8 * s, ph);
      if (mode & 2) tc_fence_after();
      if (mode & 1) tc_commit_a(empty0 + 8 * s);
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
      if (++s == S) { s = 0; ph ^= 1; }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_slot, 32);
}

// ---- 3. full main loop skeleton: TMA A (2-D box 64 x 128 rows) + W (64 x N rows) per K block into a ring of S
//      stages, consumer = MMA thread (mode bit0: issue the MMAs; else just commit)
template <int N, int S>
__global__ void __launch_bounds__(128, 1) mainloop_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmW, long long* out,
                                                          int mode, int kblocks, int rows_a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStage = 16384 + N * 128;
  __shared__ uint64_t full[S], empty[S], done;
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), sm0 = smem_u32(smem);
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    uint32_t s = 0, ph = 1;
    const int row0 = (blockIdx.x * 128) % rows_a;
    for (int it = 0; it < kIters; ++it) {
      mbar_wait_a(empty0 + 8 * s, ph);
      const uint32_t sa = sm0 + s * kStage, fb = full0 + 8 * s;
      mbar_arrive_expect_tx_a(fb, kStage);
      const int kc = (it % kblocks) * 64;
      tma_load_2d_a(sa, &tmA, fb, kc, row0);
      tma_load_2d_a(sa + 16384, &tmW, fb, kc, 0);
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    constexpr uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t a_lo0 = ((sm0 & 0x3FFFFu) >> 4) | (1u << 16);
    uint32_t s = 0, ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
      mbar_wait_a(full0 + 8 * s, ph);
      tc_fence_after();
      if (mode & 1) {
        const uint32_t a_lo = a_lo0 + s * (kStage >> 4);
        const uint64_t adesc = desc_hi | a_lo, bdesc = desc_hi | (a_lo + 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + 2u * k, bdesc + 2u * k, idesc, 1u);
      }
      tc_commit_a(empty0 + 8 * s);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    tc_commit(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make2d(EncodeTiledFn fn, void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t el[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("tensor map encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <int N, bool COMMIT>
static void run_mma_rate(long long* d_out, int grid) {
  const int smem = 16384 + N * 128 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel<N, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<N, COMMIT><<<grid, 128, smem>>>(d_out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("mma_rate  N=%3d commit/kb=%d grid=%3d : issue %.1f cyc/kb, complete %.1f cyc/kb (math floor %d)\n", N, (int)COMMIT,
         grid, (double)h[0] / kIters, (double)h[1] / kIters, 2 * N);
}

template <int S>
static void run_pingpong(long long* d_out, int grid, int mode) {
  pingpong_kernel<S><<<grid, 128>>>(d_out, mode);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("pingpong  S=%d release=%s fence=%d grid=%3d : %.1f cyc/iter\n", S, (mode & 1) ? "tcgen05.commit" : "mbarrier.arrive",
         (mode >> 1) & 1, grid, (double)h / kIters);
}

template <int N, int S>
static void run_mainloop(EncodeTiledFn fn, void* a, void* w, long long* d_out, int grid, int mode, int rows_a, int K) {
  CUtensorMap tmA = make2d(fn, a, K, rows_a, 128), tmW = make2d(fn, w, K, N, N);
  const int smem = S * (16384 + N * 128) + 1024;
  CK(cudaFuncSetAttribute(mainloop_kernel<N, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mainloop_kernel<N, S><<<grid, 128, smem>>>(tmA, tmW, d_out, mode, K / 64, rows_a);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("mainloop  N=%3d S=%d mma=%d grid=%3d A=%dx%d : %.1f cyc/kb (math floor %d, bytes/kb %d)\n", N, S, mode & 1, grid,
         rows_a, K, (double)h / kIters, 2 * N, 16384 + N * 128);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  const int rows_a = 24576, K = 2880;
  void *a, *w;
  CK(cudaMalloc(&a, (size_t)rows_a * K * 2));
  CK(cudaMalloc(&w, (size_t)256 * K * 2));
  CK(cudaMemset(a, 0, (size_t)rows_a * K * 2));
  CK(cudaMemset(w, 0, (size_t)256 * K * 2));
  printf("SMs %d, %d iterations per experiment\n", sms, kIters);
  for (int grid : {1, sms}) {
    run_mma_rate<64, true>(d_out, grid);
    run_mma_rate<128, true>(d_out, grid);
    run_mma_rate<160, true>(d_out, grid);
    run_mma_rate<256, true>(d_out, grid);
    run_mma_rate<160, false>(d_out, grid);
  }
  for (int mode = 0; mode < 4; ++mode) {
    run_pingpong<2>(d_out, sms, mode);
    run_pingpong<4>(d_out, sms, mode);
    run_pingpong<8>(d_out, sms, mode);
  }
  for (int grid : {1, sms}) {
    for (int mode = 0; mode < 2; ++mode) {
      run_mainloop<64, 8>(fn, a, w, d_out, grid, mode, rows_a, K);
      run_mainloop<160, 5>(fn, a, w, d_out, grid, mode, rows_a, K);
      run_mainloop<256, 4>(fn, a, w, d_out, grid, mode, rows_a, K);
    }
  }
  return 0;
}
