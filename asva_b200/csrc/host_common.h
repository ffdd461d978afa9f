// Host-side helpers shared by the C-ABI translation units: error reporting, driver entry point for
// cuTensorMapEncodeTiled (looked up at run time so the library does not link libcuda), launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/asva_b200.h"

namespace asva {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

// One-time launch state (opt-in shared-memory attributes, occupancy-derived grid caps, SM counts) is per DEVICE:
// translation units keep it in arrays of kMaxDevices indexed by current_device().
constexpr int kMaxDevices = 64;
int current_device();   // cudaGetDevice, clamped to [0, kMaxDevices)
int device_sms();       // multiprocessor count of the current device (cached per device; 0 on failure)

// Builds a bf16 tiled tensor map with 128-byte swizzle. dims/box/elem_strides have `rank` entries (innermost
// first); strides_bytes has rank-1 entries (dims 1..rank-1). Returns 0 or a negative asva_status.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// General form: dtype TMAP_BF16 / TMAP_F32, swizzle TMAP_SW_NONE / TMAP_SW64 / TMAP_SW128 (the box's innermost
// extent in bytes must not exceed the swizzle span).
enum { TMAP_BF16 = 0, TMAP_F32 = 1 };
enum { TMAP_SW_NONE = 0, TMAP_SW64 = 1, TMAP_SW128 = 2 };
int make_tmap(CUtensorMap* out, const void* base, int dtype, int swizzle, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// Kernel launch with the programmatic-dependent-launch attribute (and optionally a cluster of `cluster_x` CTAs).
// ASVA_NO_PDL=1 in the environment launches plainly (A/B measurements, debugging).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define ASVA_CUDA_OK(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::asva::fail(ASVA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                          __FILE__, __LINE__);                                                     \
  } while (0)

#define ASVA_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return ::asva::fail(ASVA_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

// misc.cu: the memory-bound temporal attention; 0 = launched, 1 = shape not served by it (force: not served at all;
// else: also the shapes where the tcgen05 form measured faster), other = error code
int temporal_attention_rows(const void* qkv, void* out, int B, int F, int N, int H, int d, float scale,
                            cudaStream_t stream, bool force);
// attn_mma.cu: attention against <= 128 keys on warp MMA; 0 = launched, 1 = shape not served, other = error code
int attention_small_keys(const asva_attn_desc* d, cudaStream_t stream);
// misc.cu: the warp-MMA temporal attention; 0 = launched, 1 = shape not served, other = error code
int temporal_attention_mma(const void* qkv, void* out, int B, int F, int N, int H, int d, float scale,
                           cudaStream_t stream);

}  // namespace asva
