// Host stand-in for <cuda_bf16.h> (tests/warp_emu only): bf16 storage type and the conversions the encoder's SIMT kernels
// use, round-to-nearest-even as __float2bfloat16_rn.
#pragma once
#include <stdint.h>
#include <string.h>

struct __nv_bfloat16 { uint16_t bits; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };

static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return __nv_bfloat16{uint16_t(0x7FFF)};   // NaN
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __nv_bfloat16{uint16_t(u >> 16)};
}
static inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = uint32_t(h.bits) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
