// Host-side helpers shared by the translation units of libcomorag_b200:
// error reporting for the C ABI, TMA tensor-map construction (driver entry
// point fetched through the runtime, so the library links only cudart), and
// device queries.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/comorag_b200.h"

namespace crag {

// Thread-local last-error text (the C ABI never throws; see crag_last_error()).
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define CRAG_CUDA_OK(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::crag::fail(CRAG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                                 \
  } while (0)

// 2-D bf16 row-major tensor [rows, cols] with `row_stride_bytes` between rows;
// box = box_rows x 64 columns (one 128-byte swizzle span), SWIZZLE_128B,
// out-of-bounds elements read as zero.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols = 64);

int sm_count();

}  // namespace crag
