// Full descending ranking of one query's score array on the device: the
//   sorted_doc_ids = np.argsort(query_doc_scores)[::-1]; sorted_doc_scores = query_doc_scores[sorted_doc_ids]
// step of the reference's dense_passage_retrieval (ComoRAG.py:965-966), whose whole permutation feeds the PPR reset
// weights (ComoRAG.py:1034-1042).  Equal scores rank by ascending row (the engine's deterministic tie rule; numpy's
// argsort[::-1] leaves ties unspecified).
//
// Stable LSD radix sort, 4 passes of 8 bits over key = ~orderable(score) (ascending key == descending score) with
// the row index as payload.  HBM-bound integer work: per pass every element is read twice (histogram, scatter) and
// written once; nothing here is GEMM-shaped.
//   pass p:  hist_kernel    each warp owns a contiguous run of elements and counts its 256 digit bins
//            scan_kernel    block d turns hist[d][*] into exclusive global offsets (digit base + running sum)
//            scatter_kernel each warp replays its run in order, 32 elements a round; equal digits inside a round
//                           are ranked with match.any so the scatter stays stable
// Pass 0 builds the keys from the fp32 scores on the fly, the last pass writes int64 ids + fp32 scores.
#include "common.cuh"
#include "topk.cuh"

namespace crag {
namespace {

constexpr int kSortThreads = 256;                       // 8 warps per CTA
constexpr int kWarpsPerCta = kSortThreads / 32;

struct SortPlan {
  int run;         // elements per warp (multiple of 32)
  int n_warps;     // total warps
  int grid;
  size_t key_bytes, val_bytes, hist_bytes, total;
};

SortPlan plan_sort(int64_t n) {
  SortPlan p;
  // runs of 2048 elements; for big arrays grow the run so the histogram table stays small (<= 8192 warps)
  int64_t run = 2048;
  while ((n + run - 1) / run > 8192) run *= 2;
  p.run = int(run);
  p.n_warps = int((n + run - 1) / run);
  if (p.n_warps < 1) p.n_warps = 1;
  p.grid = (p.n_warps + kWarpsPerCta - 1) / kWarpsPerCta;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  p.key_bytes = up(size_t(n) * 4);
  p.val_bytes = up(size_t(n) * 4);
  p.hist_bytes = up(size_t(256) * p.n_warps * 4 + 256 * 4);
  p.total = 2 * p.key_bytes + 2 * p.val_bytes + p.hist_bytes;
  return p;
}

__device__ __forceinline__ uint32_t sort_key(float s) { return ~orderable_f32(s); }

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) hist_kernel(const float* __restrict__ scores,
                                                            const uint32_t* __restrict__ keys_in, int64_t n, int run,
                                                            int n_warps, int shift, uint32_t* __restrict__ hist,
                                                            uint32_t* __restrict__ totals) {
  __shared__ uint32_t s_cnt[kWarpsPerCta][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerCta + w;
  for (int i = lane; i < 256; i += 32) s_cnt[w][i] = 0;
  __syncwarp();
  if (gw < n_warps) {
    const int64_t lo = int64_t(gw) * run;
    const int64_t hi = lo + run < n ? lo + run : n;
    for (int64_t i = lo + lane; i < hi; i += 32) {
      const uint32_t key = FIRST ? sort_key(scores[i]) : keys_in[i];
      atomicAdd(&s_cnt[w][(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    for (int d = lane; d < 256; d += 32) {
      const uint32_t c = s_cnt[w][d];
      hist[size_t(d) * n_warps + gw] = c;
      if (c) atomicAdd(&totals[d], c);
    }
  }
}

// block d: exclusive offsets of digit d's per-warp counts, starting at the number of elements with a smaller digit
__global__ void __launch_bounds__(256) scan_kernel(uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals,
                                                   int n_warps) {
  __shared__ uint32_t s_part[256];
  __shared__ uint32_t s_base;
  const int d = blockIdx.x, t = threadIdx.x;
  s_part[t] = t < d ? totals[t] : 0u;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_part[t] += s_part[t + o];
    __syncthreads();
  }
  if (t == 0) s_base = s_part[0];
  __syncthreads();
  uint32_t carry = s_base;
  uint32_t* row = hist + size_t(d) * n_warps;
  for (int base = 0; base < n_warps; base += 256) {
    const int i = base + t;
    const uint32_t v = i < n_warps ? row[i] : 0u;
    // inclusive block scan (Hillis-Steele over 256 entries)
    __syncthreads();
    s_part[t] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const uint32_t add = t >= o ? s_part[t - o] : 0u;
      __syncthreads();
      s_part[t] += add;
      __syncthreads();
    }
    if (i < n_warps) row[i] = carry + s_part[t] - v;
    carry += s_part[255];
  }
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kSortThreads) scatter_kernel(const float* __restrict__ scores,
                                                               const uint32_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in, int64_t n, int run,
                                                               int n_warps, int shift, const uint32_t* __restrict__ hist,
                                                               uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out,
                                                               int64_t* __restrict__ out_ids,
                                                               float* __restrict__ out_scores) {
  __shared__ uint32_t s_off[kWarpsPerCta][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerCta + w;
  if (gw >= n_warps) return;
  for (int d = lane; d < 256; d += 32) s_off[w][d] = hist[size_t(d) * n_warps + gw];
  __syncwarp();
  const int64_t lo = int64_t(gw) * run;
  const int64_t hi = lo + run < n ? lo + run : n;
  const uint32_t lt = (1u << lane) - 1u;
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t i = base + lane;
    const bool live = i < hi;
    uint32_t key = 0, val = 0;
    if (live) {
      key = FIRST ? sort_key(scores[i]) : keys_in[i];
      val = FIRST ? uint32_t(i) : vals_in[i];
    }
    // dead lanes get digits no live lane can have, so they never share a match group
    const uint32_t digit = live ? ((key >> shift) & 255u) : (256u + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & lt);
    uint32_t pos = 0;
    if (live) pos = s_off[w][digit] + rank;
    __syncwarp();
    if (live && rank == 0) s_off[w][digit] += __popc(peers);
    __syncwarp();
    if (live) {
      if (LAST) {
        out_ids[pos] = int64_t(val);
        out_scores[pos] = unorderable_f32(~key);
      } else {
        keys_out[pos] = key;
        vals_out[pos] = val;
      }
    }
  }
}

}  // namespace
}  // namespace crag

using namespace crag;

extern "C" size_t crag_rank_workspace_bytes(int64_t n) {
  if (n < 0 || n >= (int64_t(1) << 31)) return 0;
  return plan_sort(n).total;
}

extern "C" int crag_rank_scores(const float* scores, int64_t n, int64_t* out_ids, float* out_scores, void* workspace,
                                size_t workspace_bytes, crag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || n >= (int64_t(1) << 31)) return fail(CRAG_ERR_INVALID, "crag_rank_scores: n out of range (%lld)", (long long)n);
  if (n == 0) return CRAG_OK;
  if (!scores || !out_ids || !out_scores || !workspace) return fail(CRAG_ERR_INVALID, "crag_rank_scores: null pointer");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(CRAG_ERR_INVALID, "crag_rank_scores: workspace must be 256-byte aligned");
  const SortPlan p = plan_sort(n);
  if (workspace_bytes < p.total) return fail(CRAG_ERR_WORKSPACE, "crag_rank_scores: workspace %zu < %zu bytes", workspace_bytes, p.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint32_t* keys[2] = {reinterpret_cast<uint32_t*>(ws), reinterpret_cast<uint32_t*>(ws + p.key_bytes)};
  uint32_t* vals[2] = {reinterpret_cast<uint32_t*>(ws + 2 * p.key_bytes), reinterpret_cast<uint32_t*>(ws + 2 * p.key_bytes + p.val_bytes)};
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + 2 * p.key_bytes + 2 * p.val_bytes);
  uint32_t* totals = hist + size_t(256) * p.n_warps;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    const uint32_t* kin = pass ? keys[(pass - 1) & 1] : nullptr;
    const uint32_t* vin = pass ? vals[(pass - 1) & 1] : nullptr;
    uint32_t* kout = keys[pass & 1];
    uint32_t* vout = vals[pass & 1];
    CRAG_CUDA_OK(cudaMemsetAsync(totals, 0, 256 * 4, stream));
    if (pass == 0) hist_kernel<true><<<p.grid, kSortThreads, 0, stream>>>(scores, nullptr, n, p.run, p.n_warps, shift, hist, totals);
    else hist_kernel<false><<<p.grid, kSortThreads, 0, stream>>>(nullptr, kin, n, p.run, p.n_warps, shift, hist, totals);
    scan_kernel<<<256, 256, 0, stream>>>(hist, totals, p.n_warps);
    if (pass == 0) scatter_kernel<true, false><<<p.grid, kSortThreads, 0, stream>>>(scores, nullptr, nullptr, n, p.run, p.n_warps, shift, hist, kout, vout, nullptr, nullptr);
    else if (pass < 3) scatter_kernel<false, false><<<p.grid, kSortThreads, 0, stream>>>(nullptr, kin, vin, n, p.run, p.n_warps, shift, hist, kout, vout, nullptr, nullptr);
    else scatter_kernel<false, true><<<p.grid, kSortThreads, 0, stream>>>(nullptr, kin, vin, n, p.run, p.n_warps, shift, hist, nullptr, nullptr, out_ids, out_scores);
    CRAG_CUDA_OK(cudaGetLastError());
  }
  return CRAG_OK;
}
