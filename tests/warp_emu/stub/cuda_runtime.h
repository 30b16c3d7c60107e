// Host stand-in for <cuda_runtime.h>, seen ONLY by tests/warp_emu (g++ -I tests/warp_emu/stub): it lets the pure-SIMT
// headers of the search kernel (csrc/topk.cuh, csrc/pool_floor.cuh) compile for the CPU with the warp intrinsics they
// use mapped onto 32 cooperatively scheduled lanes (warp_emu.h).  Test infrastructure, not product code.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../warp_emu.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))

static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }

template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::barrier(); }

template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return warp_emu::exchange(v, [&](int lane) { return lane ^ lane_mask; });
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  return warp_emu::exchange(v, [&](int) { return src & 31; });
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  return warp_emu::reduce<unsigned>(pred ? (1u << warp_emu::lane()) : 0u, [](unsigned a, unsigned b) { return a | b; });
}
static inline int __reduce_add_sync(unsigned, int v) { return warp_emu::reduce<int>(v, [](int a, int b) { return a + b; }); }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return warp_emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a + b; }); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return warp_emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return warp_emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return warp_emu::reduce<unsigned>(v, [](unsigned a, unsigned b) { return a | b; }); }
