// Error plumbing, TMA tensor-map construction and device queries for the C ABI.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace crag {

namespace {
thread_local char g_err[512] = {0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
}  // namespace

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

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(CRAG_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  if (rows == 0 || cols == 0) return fail(CRAG_ERR_INVALID, "tensor map over an empty tensor");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d; rows=%llu cols=%llu stride=%llu box=%ux%u)",
                int(r), (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes,
                box_rows, box_cols);
  return CRAG_OK;
}

int sm_count() {
  // cached per device; a handful of devices at most
  static int cached[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    fail(CRAG_ERR_CUDA, "no usable CUDA device (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    return -1;
  }
  if (cached[dev] == 0) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
      fail(CRAG_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace crag

extern "C" int crag_version(void) { return 1000; }
extern "C" const char* crag_last_error(void) { return crag::g_err; }
extern "C" int crag_sm_count(void) { return crag::sm_count(); }
