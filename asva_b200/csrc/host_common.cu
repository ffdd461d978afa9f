#include "host_common.h"

#include <stdlib.h>
#include <string.h>

namespace asva {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int device_sms() {
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) sms[dev] = n;
  }
  return sms[dev];
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ASVA_NO_PDL");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_tmap(out, base, TMAP_BF16, TMAP_SW128, rank, dims, strides_bytes, box, elem_strides);
}

int make_tmap(CUtensorMap* out, const void* base, int dtype, int swizzle, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(ASVA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0)
    return fail(ASVA_ERR_INVALID, "tensor base %p is not 16-byte aligned", base);
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5], gel[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    gel[i] = elem_strides[i];
    if (box[i] == 0 || box[i] > 256) return fail(ASVA_ERR_INVALID, "TMA box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if ((strides_bytes[i] & 15u) != 0 || strides_bytes[i] == 0)
      return fail(ASVA_ERR_INVALID, "TMA stride[%d]=%llu bytes must be a non-zero multiple of 16", i,
                  (unsigned long long)strides_bytes[i]);
  }
  const CUtensorMapDataType dt = (dtype == TMAP_F32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = (swizzle == TMAP_SW128)  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle == TMAP_SW64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                         : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, gel,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(ASVA_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r,
                rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

}  // namespace asva

extern "C" const char* asva_last_error(void) { return asva::g_err; }
extern "C" int asva_version(void) { return 100; }
extern "C" int asva_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return asva::fail(ASVA_ERR_CUDA, "no CUDA device");
  if (prop.major != 10) return asva::fail(ASVA_ERR_DEVICE, "device is sm_%d%d, need sm_100", prop.major, prop.minor);
  return 0;
}
