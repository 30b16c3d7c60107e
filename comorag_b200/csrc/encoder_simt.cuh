// The encoder's non-GEMM, non-attention kernels: embedding gather + LayerNorm, LayerNorm, masked mean pool + L2
// normalise (K3; BGEEmbedding.py:15-28, :127) and the sequence-classification head.  SIMT code with bf16 loads / stores
// and fp32 arithmetic -- no tensor-core, TMA or cp.async instruction -- kept in a header so tests/warp_emu can run
// exactly these kernels on emulated thread blocks against double-precision models.
#pragma once
#include <math.h>
#include <stdint.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#ifndef CRAG_EMULATED_PTX   // tests/warp_emu supplies host versions of these two
// dynamic shared memory of the kernel, as an array named `name`
#define CRAG_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
namespace crag {
// named barrier 1 over the first 128 threads of the block (pool_normalize_kernel's token group 0)
__device__ __forceinline__ void bar_sync_group0_128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
}  // namespace crag
#endif

namespace crag {

// --------------------------------------------------------------- LayerNorm
// One warp per row; each lane owns VPL 16-byte vectors (8 bf16) of the row.
template <int VPL>
__device__ __forceinline__ void warp_layernorm_store(float (&x)[VPL][8], int H, int lane, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps,
                                                     __nv_bfloat16* __restrict__ orow) {
  const int nvec = H / 8;
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
    if (lane + v * 32 < nvec)
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += x[v][j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / float(H);
  float var = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
    if (lane + v * 32 < nvec)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[v][j] - mean;
        var += d * d;
      }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / float(H) + eps);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = lane + v * 32;
    if (vec < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vec * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vec * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vec * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vec * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      __nv_bfloat162 o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = __floats2bfloat162_rn((x[v][2 * j] - mean) * rstd * g[2 * j] + b[2 * j],
                                     (x[v][2 * j + 1] - mean) * rstd * g[2 * j + 1] + b[2 * j + 1]);
      *reinterpret_cast<uint4*>(orow + vec * 8) = *reinterpret_cast<uint4*>(o);
    }
  }
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __bfloat162float(p[j].x);
    f[2 * j + 1] = __bfloat162float(p[j].y);
  }
}

template <int VPL>
__global__ void __launch_bounds__(128) embed_layernorm_kernel(const int32_t* __restrict__ token_ids,
                                                              const int32_t* __restrict__ cu_seqlens, int n_seqs,
                                                              int T, int H, int vocab, int max_pos, int pos_offset,
                                                              const __nv_bfloat16* __restrict__ word_emb,
                                                              const __nv_bfloat16* __restrict__ pos_emb,
                                                              const __nv_bfloat16* __restrict__ type_emb,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              __nv_bfloat16* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  // position within the owning sequence: binary search cu_seqlens (n_seqs is small)
  int lo = 0, hi = n_seqs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cu_seqlens + mid) <= t) lo = mid; else hi = mid;
  }
  int pos = t - __ldg(cu_seqlens + lo) + pos_offset;
  pos = pos < max_pos ? pos : max_pos - 1;
  int id = __ldg(token_ids + t);
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const int nvec = H / 8;
  float x[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = lane + v * 32;
    if (vec < nvec) {
      float a[8], b[8], c[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(word_emb + int64_t(id) * H + vec * 8)), a);
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(pos_emb + int64_t(pos) * H + vec * 8)), b);
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(type_emb + vec * 8)), c);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[v][j] = a[j] + b[j] + c[j];
    }
  }
  warp_layernorm_store<VPL>(x, H, lane, gamma, beta, eps, out + int64_t(t) * H);
}

template <int VPL>
__global__ void __launch_bounds__(128) layernorm_kernel(const __nv_bfloat16* __restrict__ in, int T, int H,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        __nv_bfloat16* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= T) return;
  const int nvec = H / 8;
  float x[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int vec = lane + v * 32;
    if (vec < nvec) bf16x8_to_float(*reinterpret_cast<const uint4*>(in + int64_t(t) * H + vec * 8), x[v]);
  }
  warp_layernorm_store<VPL>(x, H, lane, gamma, beta, eps, out + int64_t(t) * H);
}

// ----------------------------------------------------------- pool + normalise
// one block per sequence: kPoolGroups token groups x 128 column threads (8 bf16 columns each, up to 2 vectors:
// H <= 2048); groups stride over the tokens, partial sums meet in shared memory.
constexpr int kPoolGroups = 8;
__global__ void __launch_bounds__(128 * kPoolGroups) pool_normalize_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                                           const int32_t* __restrict__ cu_seqlens, int H,
                                                                           int normalize, float* __restrict__ out_f32,
                                                                           __nv_bfloat16* __restrict__ out_bf16,
                                                                           int64_t out_bf16_stride) {
  CRAG_DYNAMIC_SHARED(float, s_pool);          // [kPoolGroups][H] partial sums, then [4] norm partials
  const int seq = blockIdx.x;
  const int tid = threadIdx.x & 127, grp = threadIdx.x >> 7;
  const int start = __ldg(cu_seqlens + seq);
  const int L = __ldg(cu_seqlens + seq + 1) - start;
  const int nvec = H / 8;
  float acc[2][8];
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
  for (int t = grp; t < L; t += kPoolGroups) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int vec = tid + v * 128;
      if (vec < nvec) {
        float f[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(hidden + int64_t(start + t) * H + vec * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[v][j] += f[j];
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int vec = tid + v * 128;
    if (vec < nvec)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_pool[grp * H + vec * 8 + j] = acc[v][j];
  }
  __syncthreads();
  if (grp != 0) return;
  const float inv_len = 1.f / float(L);  // L == 0 -> inf/nan row, as the reference's 0/0 would give
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int vec = tid + v * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sum = 0.f;
      if (vec < nvec)
        for (int g2 = 0; g2 < kPoolGroups; ++g2) sum += s_pool[g2 * H + vec * 8 + j];  // fixed order: deterministic
      acc[v][j] = sum * inv_len;
      ss += acc[v][j] * acc[v][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float* s_part = s_pool + kPoolGroups * H;
  if ((tid & 31) == 0) s_part[tid >> 5] = ss;
  bar_sync_group0_128();                           // group 0 only
  const float norm = sqrtf(s_part[0] + s_part[1] + s_part[2] + s_part[3]);
  const float inv = normalize ? 1.f / fmaxf(norm, 1e-12f) : 1.f;  // F.normalize eps
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int vec = tid + v * 128;
    if (vec < nvec) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = acc[v][j] * inv;
      if (out_f32) {
        float4* dst = reinterpret_cast<float4*>(out_f32 + int64_t(seq) * H + vec * 8);
        dst[0] = make_float4(r[0], r[1], r[2], r[3]);
        dst[1] = make_float4(r[4], r[5], r[6], r[7]);
      }
      if (out_bf16) {
        __nv_bfloat162 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
        *reinterpret_cast<uint4*>(out_bf16 + int64_t(seq) * out_bf16_stride + vec * 8) = *reinterpret_cast<uint4*>(o);
      }
    }
  }
}

// ------------------------------------------------- classification head
// Sequence-classification head on the FIRST token of every sequence (the cross-encoder rerank score):
//   logits = W_out tanh(W_dense h[<s>] + b_dense) + b_out
// (HF XLMRobertaClassificationHead; dropout is the identity in eval).  One block per sequence, 8 warps: a warp owns
// output features o, o+8, ... and its lanes stride the H inputs in 16-byte vectors, so every weight row is read as
// contiguous 512-byte segments (L2-resident after the first block).  H <= 1024, multiple of 8.
constexpr int kClsWarps = 8;
__device__ __forceinline__ float warp_dot_bf16(const __nv_bfloat16* __restrict__ wr, const float* s_in, int nvec, int lane) {
  float acc = 0.f;
  for (int vec = lane; vec < nvec; vec += 32) {
    float f[8];
    bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(wr + vec * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(f[j], s_in[vec * 8 + j], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

__global__ void __launch_bounds__(32 * kClsWarps) cls_head_kernel(const __nv_bfloat16* __restrict__ hidden,
                                                                  const int32_t* __restrict__ cu_seqlens, int H,
                                                                  const __nv_bfloat16* __restrict__ w_dense,
                                                                  const float* __restrict__ b_dense,
                                                                  const __nv_bfloat16* __restrict__ w_out,
                                                                  const float* __restrict__ b_out, int n_labels,
                                                                  float* __restrict__ logits) {
  __shared__ float s_x[1024];
  __shared__ float s_y[1024];
  const int seq = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = H / 8;
  const __nv_bfloat16* row = hidden + int64_t(__ldg(cu_seqlens + seq)) * H;
  for (int vec = threadIdx.x; vec < nvec; vec += 32 * kClsWarps) {
    float f[8];
    bf16x8_to_float(*reinterpret_cast<const uint4*>(row + vec * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s_x[vec * 8 + j] = f[j];
  }
  __syncthreads();
  for (int o = warp; o < H; o += kClsWarps) {
    const float acc = warp_dot_bf16(w_dense + int64_t(o) * H, s_x, nvec, lane);
    if (lane == 0) s_y[o] = tanhf(acc + __ldg(b_dense + o));
  }
  __syncthreads();
  for (int o = warp; o < n_labels; o += kClsWarps) {
    const float acc = warp_dot_bf16(w_out + int64_t(o) * H, s_y, nvec, lane);
    if (lane == 0) logits[int64_t(seq) * n_labels + o] = acc + __ldg(b_out + o);
  }
}

}  // namespace crag
