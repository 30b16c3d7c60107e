// The encoder's non-GEMM, non-attention kernels (csrc/encoder_simt.cuh) on emulated thread blocks, against
// double-precision models of what the reference computes with HF / torch:
//   embed_layernorm_kernel  word + position (+ offset: XLM-R) + token-type-0 embedding gather -> LayerNorm
//   layernorm_kernel        torch.nn.LayerNorm over the last dimension (fp32 statistics, bf16 in / out)
//   pool_normalize_kernel   mean_pooling (BGEEmbedding.py:15-28) over the UNPADDED token stream + F.normalize (:127),
//                           fp32 rows and the bf16 row written into a strided shard
//   cls_head_kernel         out_proj(tanh(dense(h[first token]))) -- XLMRobertaClassificationHead
// Tolerances: bf16 outputs within one bf16 ulp of the rounded model value (the GPU's and the host's fp32 summation
// orders differ); fp32 outputs within 2e-6 relative.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <vector>

#include <cuda_runtime.h>   // the stubs
#include <cuda_bf16.h>

#include "encoder_simt.cuh"

using namespace crag;

static std::mt19937_64 rng(777);
static float frand(float lo, float hi) { return lo + (hi - lo) * float(rng() % 1000001) / 1e6f; }
static float gauss() { float s = 0; for (int i = 0; i < 6; ++i) s += frand(-1.f, 1.f); return s * 0.7f; }

#define REQUIRE(cond, ...)                              \
  do {                                                  \
    if (!(cond)) {                                      \
      fprintf(stderr, "FAILED %s:%d: %s\n  ", __FILE__, __LINE__, #cond); \
      fprintf(stderr, __VA_ARGS__);                     \
      fprintf(stderr, "\n");                            \
      exit(1);                                          \
    }                                                   \
  } while (0)

static std::vector<__nv_bfloat16> bf16_vec(size_t n, float scale) {
  std::vector<__nv_bfloat16> v(n);
  for (auto& x : v) x = __float2bfloat16_rn(gauss() * scale);
  return v;
}
static float ulp_bf16(float x) { return std::max(std::fabs(x), 1e-3f) * (1.0f / 128.0f); }   // 2^-7: one bf16 step at |x|

// LayerNorm of one row in double precision
static std::vector<double> layernorm_model(const std::vector<double>& x, const std::vector<float>& g, const std::vector<float>& b, double eps) {
  const size_t H = x.size();
  double mean = 0, var = 0;
  for (double v : x) mean += v;
  mean /= double(H);
  for (double v : x) var += (v - mean) * (v - mean);
  var /= double(H);
  std::vector<double> y(H);
  for (size_t i = 0; i < H; ++i) y[i] = (x[i] - mean) / std::sqrt(var + eps) * g[i] + b[i];
  return y;
}

template <int VPL>
static void test_layernorm(int H, int T) {
  std::vector<__nv_bfloat16> in = bf16_vec(size_t(T) * H, 1.5f), out(size_t(T) * H);
  std::vector<float> g(H), b(H);
  for (int i = 0; i < H; ++i) { g[i] = frand(0.5f, 1.5f); b[i] = frand(-0.3f, 0.3f); }
  warp_emu::launch((T + 3) / 4, 128, [&] { layernorm_kernel<VPL>(in.data(), T, H, g.data(), b.data(), 1e-12f, out.data()); });
  for (int t = 0; t < T; ++t) {
    std::vector<double> x(H);
    for (int i = 0; i < H; ++i) x[i] = __bfloat162float(in[size_t(t) * H + i]);
    std::vector<double> y = layernorm_model(x, g, b, 1e-12);
    for (int i = 0; i < H; ++i) {
      const float got = __bfloat162float(out[size_t(t) * H + i]);
      REQUIRE(std::fabs(got - float(y[i])) <= ulp_bf16(float(y[i])), "layernorm<%d> H=%d row %d col %d: %g vs %g", VPL, H, t, i, got, y[i]);
    }
  }
  printf("ok  layernorm_kernel<%d>: H = %d, %d rows\n", VPL, H, T);
}

template <int VPL>
static void test_embed_layernorm(int H, int pos_offset) {
  const int vocab = 500, max_pos = 64 + pos_offset, n_seqs = 5;
  const int lens[n_seqs] = {1, 17, 64, 3, 40};
  std::vector<int32_t> cu(n_seqs + 1, 0), ids;
  for (int s = 0; s < n_seqs; ++s) {
    cu[s + 1] = cu[s] + lens[s];
    for (int j = 0; j < lens[s]; ++j) ids.push_back(int32_t(rng() % vocab));
  }
  const int T = cu[n_seqs];
  std::vector<__nv_bfloat16> we = bf16_vec(size_t(vocab) * H, 1.f), pe = bf16_vec(size_t(max_pos) * H, 0.5f), te = bf16_vec(size_t(2) * H, 0.2f), out(size_t(T) * H);
  std::vector<float> g(H), b(H);
  for (int i = 0; i < H; ++i) { g[i] = frand(0.8f, 1.2f); b[i] = frand(-0.1f, 0.1f); }
  warp_emu::launch((T + 3) / 4, 128, [&] {
    embed_layernorm_kernel<VPL>(ids.data(), cu.data(), n_seqs, T, H, vocab, max_pos, pos_offset, we.data(), pe.data(), te.data(), g.data(), b.data(), 1e-12f, out.data());
  });
  for (int s = 0; s < n_seqs; ++s)
    for (int j = 0; j < lens[s]; ++j) {
      const int t = cu[s] + j, pos = j + pos_offset;        // position restarts in every sequence (+ XLM-R's padding offset)
      std::vector<double> x(H);
      for (int i = 0; i < H; ++i)
        x[i] = double(__bfloat162float(we[size_t(ids[t]) * H + i])) + __bfloat162float(pe[size_t(pos) * H + i]) + __bfloat162float(te[i]);
      std::vector<double> y = layernorm_model(x, g, b, 1e-12);
      for (int i = 0; i < H; ++i) {
        const float got = __bfloat162float(out[size_t(t) * H + i]);
        REQUIRE(std::fabs(got - float(y[i])) <= ulp_bf16(float(y[i])), "embed_layernorm<%d> H=%d token %d col %d: %g vs %g", VPL, H, t, i, got, y[i]);
      }
    }
  printf("ok  embed_layernorm_kernel<%d>: H = %d, position offset %d, %d packed tokens of %d sequences\n", VPL, H, pos_offset, T, n_seqs);
}

static void test_pool_normalize(int H, int normalize) {
  const int n_seqs = 6;
  const int lens[n_seqs] = {1, 7, 8, 9, 130, 33};       // fewer tokens than token groups, exactly 8, more than 8 ...
  std::vector<int32_t> cu(n_seqs + 1, 0);
  for (int s = 0; s < n_seqs; ++s) cu[s + 1] = cu[s] + lens[s];
  const int T = cu[n_seqs];
  std::vector<__nv_bfloat16> hidden = bf16_vec(size_t(T) * H, 1.f);
  std::vector<float> out(size_t(n_seqs) * H, -9.f);
  const int64_t stride = H + 64;                           // the corpus shard's padded row stride
  std::vector<__nv_bfloat16> shard(size_t(n_seqs) * stride, __nv_bfloat16{0x7FC0});
  const size_t smem = (size_t(kPoolGroups) * H + 4) * sizeof(float);
  warp_emu::launch(n_seqs, 128 * kPoolGroups, [&] {
    pool_normalize_kernel(hidden.data(), cu.data(), H, normalize, out.data(), shard.data(), stride);
  }, smem);
  for (int s = 0; s < n_seqs; ++s) {
    std::vector<double> m(H, 0.0);
    for (int t = cu[s]; t < cu[s + 1]; ++t)
      for (int i = 0; i < H; ++i) m[i] += __bfloat162float(hidden[size_t(t) * H + i]);
    double n2 = 0;
    for (int i = 0; i < H; ++i) { m[i] /= double(lens[s]); n2 += m[i] * m[i]; }
    const double inv = normalize ? 1.0 / std::max(std::sqrt(n2), 1e-12) : 1.0;
    for (int i = 0; i < H; ++i) {
      const double want = m[i] * inv;
      const float got = out[size_t(s) * H + i];
      REQUIRE(std::fabs(got - want) <= 2e-6 * std::max(1.0, std::fabs(want)), "pool_normalize H=%d seq %d col %d: %g vs %g", H, s, i, got, want);
      const float gb = __bfloat162float(shard[size_t(s) * stride + i]);
      REQUIRE(gb == __bfloat162float(__float2bfloat16_rn(got)), "pool_normalize bf16 row: seq %d col %d", s, i);
    }
    for (int i = H; i < stride; ++i) REQUIRE(shard[size_t(s) * stride + i].bits == 0x7FC0, "pool_normalize wrote past the row (seq %d col %d)", s, i);
  }
  printf("ok  pool_normalize_kernel: H = %d, normalize = %d, lengths 1..130, fp32 rows + bf16 rows in a strided shard\n", H, normalize);
}

static void test_cls_head(int H, int n_labels) {
  const int n_seqs = 4;
  const int lens[n_seqs] = {5, 1, 12, 3};
  std::vector<int32_t> cu(n_seqs + 1, 0);
  for (int s = 0; s < n_seqs; ++s) cu[s + 1] = cu[s] + lens[s];
  std::vector<__nv_bfloat16> hidden = bf16_vec(size_t(cu[n_seqs]) * H, 1.f);
  std::vector<__nv_bfloat16> wd = bf16_vec(size_t(H) * H, 1.f / std::sqrt(float(H))), wo = bf16_vec(size_t(n_labels) * H, 1.f / std::sqrt(float(H)));
  std::vector<float> bd(H), bo(n_labels), logits(size_t(n_seqs) * n_labels, -9.f);
  for (auto& x : bd) x = frand(-0.2f, 0.2f);
  for (auto& x : bo) x = frand(-0.2f, 0.2f);
  warp_emu::launch(n_seqs, 32 * kClsWarps, [&] {
    cls_head_kernel(hidden.data(), cu.data(), H, wd.data(), bd.data(), wo.data(), bo.data(), n_labels, logits.data());
  });
  for (int s = 0; s < n_seqs; ++s) {
    const __nv_bfloat16* row = hidden.data() + size_t(cu[s]) * H;      // the sequence's FIRST token
    std::vector<double> y(H);
    for (int o = 0; o < H; ++o) {
      double acc = bd[o];
      for (int i = 0; i < H; ++i) acc += double(__bfloat162float(wd[size_t(o) * H + i])) * __bfloat162float(row[i]);
      y[o] = std::tanh(acc);
    }
    for (int o = 0; o < n_labels; ++o) {
      double acc = bo[o];
      for (int i = 0; i < H; ++i) acc += double(__bfloat162float(wo[size_t(o) * H + i])) * y[i];
      REQUIRE(std::fabs(logits[size_t(s) * n_labels + o] - acc) <= 1e-4, "cls_head H=%d seq %d label %d: %g vs %g", H, s, o, logits[size_t(s) * n_labels + o], acc);
    }
  }
  printf("ok  cls_head_kernel: H = %d, %d label(s)\n", H, n_labels);
}

int main() {
  test_layernorm<1>(128, 9);
  test_layernorm<2>(384, 6);
  test_layernorm<4>(768, 5);
  test_layernorm<4>(1024, 5);
  test_embed_layernorm<1>(128, 0);
  test_embed_layernorm<2>(384, 0);
  test_embed_layernorm<4>(1024, 2);
  test_pool_normalize(128, 1);
  test_pool_normalize(384, 1);
  test_pool_normalize(1024, 1);
  test_pool_normalize(768, 0);
  test_cls_head(128, 1);
  test_cls_head(1024, 1);
  test_cls_head(256, 3);
  printf("ALL OK\n");
  return 0;
}
