// Constants, shared-memory layout and argument structs of the shard scan kernel (search.cu), in a header of their own
// so that tests/warp_emu can build the select warps' code (select_warps.inc.cuh) for the host.
#pragma once
#include <math.h>
#include <stdint.h>
#include <cuda_runtime.h>

#include "ivf_kernels.cuh"   // kTileRows
#include "pool_floor.cuh"    // kNQ

namespace crag {

constexpr int kBlockK = 64;     // bf16 per 128-byte swizzle row
constexpr int kStageBytes = kTileRows * kBlockK * 2;  // 16 KB
constexpr int kQBlockBytes = kNQ * kBlockK * 2;       // 4 KB
constexpr int kSearchThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kAccStages = 16;       // score-tile buffers in TMEM: the scan may run 16 tiles ahead of the select warps
constexpr uint32_t kTmemCols = kAccStages * kNQ;  // 512 columns = all of TMEM (1 CTA per SM)

template <int KLIST, int CAP, int STAGES>
struct SearchLayout {
  static constexpr int kKeysPerQuery = KLIST + CAP;
  __host__ __device__ static constexpr size_t keys_bytes() { return size_t(kNQ) * kKeysPerQuery * 8; }
  __host__ __device__ static constexpr size_t misc_bytes() {
    return (2 * STAGES + 2 * kAccStages + 1) * 8    // mbarriers
           + kNQ * 8               // thr_key
           + kNQ * 4               // thr_f
           + kNQ * 4               // cnt
           + kNQ * 8 + kNQ * 4     // continuation bound (key, score)
           + kNQ * 8               // pooled admission floor (key)
           + 4 * kNQ * 8           // per-warp partial floors of a refresh
           + 4 * kNQ * 2 * 4       // min/max cross-warp reduction
           + 16;                   // tmem base
  }
  __host__ static size_t smem_bytes(int num_kb) {
    return 1024 + size_t(STAGES) * kStageBytes + size_t(num_kb) * kQBlockBytes + keys_bytes() + misc_bytes();
  }
};

// IVF variant of the scan (BASELINE config 4; semantic: oracle/ivf_oracle.py): the shard holds bf16 RESIDUALS grouped
// by coarse list, every list padded to whole 128-row tiles, and a pass touches only the tiles of probed lists.
//   work[i]   = (first stored row of the tile, valid rows in it, list id, 0), written by ivf_plan_kernel
//   n_work    = number of work items (device scalar: the plan is built on the device, no host round trip)
//   list_mask = per list, bit q set when query q probes it; coarse[list * 32 + q] = q . c_list from the coarse pass
// The flat instantiations (IVF = false) take an empty struct instead and compile to the same SASS as before.
// r[q] for a runtime q without sending the score registers to local memory: a 5-level select tree on the bits of q
__device__ __forceinline__ uint32_t pick32(const uint32_t (&r)[32], int q) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (q & 16) ? r[i + 16] : r[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (q & 8) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (q & 4) ? b[i + 4] : b[i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (q & 2) ? c[i + 2] : c[i];
  return (q & 1) ? d[1] : d[0];
}

struct IvfArgs {
  const int4* work;
  const int* n_work;
  const uint32_t* list_mask;
  const float* coarse;
};
struct NoIvfArgs {};
// Score-all variant (SCORES = true): the full-array contracts of the reference -- get_fact_scores returns the score
// of EVERY fact row (ComoRAG.py:937-948) and dense_passage_retrieval a permutation of ALL rows (:950-967, consumed
// rank by rank by PPR at :1034-1042).  Same TMA -> tcgen05 -> TMEM stream; the select warps write the fp32 scores
// (out[q * ld + row], one coalesced 128-byte store per warp and query) instead of running the selector.
struct ScoreArgs {
  float* out;
  int64_t ld;
  // assignment mode (best_id != nullptr): instead of storing the scores, each row keeps the running argmax over the
  // query blocks it has met -- the IVF build's "which centroid does this row belong to" (rows = corpus, centroids =
  // queries, 32 per pass).  A row is owned by one thread per pass and passes are stream-ordered: plain read-modify-write.
  float* best_score;
  int32_t* best_id;
  int32_t base_id;
};
template <bool IVF, bool SCORES> struct IvfParam { using type = NoIvfArgs; };
template <> struct IvfParam<true, false> { using type = IvfArgs; };
template <> struct IvfParam<false, true> { using type = ScoreArgs; };

// multiplier of the group permutation g -> (g * P) mod n_groups: near n_groups / golden ratio, coprime to n_groups
// (0 = identity / permutation off)
inline uint32_t perm_multiplier(int64_t num_tiles) {
  if (num_tiles < 4) return 0u;
  auto gcd = [](uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; };
  uint64_t p = (uint64_t(double(num_tiles) * 0.6180339887498949) | 1ull);
  while (gcd(p, uint64_t(num_tiles)) != 1) p += 2;
  return uint32_t(p % uint64_t(num_tiles));
}

}  // namespace crag
