// CPU check of the search kernel's selector primitives, compiled from the SAME headers the kernel uses
// (comorag_b200/csrc/topk.cuh, pool_floor.cuh) with the warp intrinsics emulated by 32 cooperatively scheduled lanes
// (warp_emu.h, stub/cuda_runtime.h).  Properties checked, each against a plain std::sort model:
//   1. warp_sort_desc<EPL>      sorts 32*EPL keys descending (blocked layout), duplicates and zeros included
//   2. flush_query<KLIST, CAP>  list u first c buffered candidates -> best k, descending, zero padded; thr = k-th key
//   3. insert_few<KLIST, CAP>   the same contract for c <= kInsertMax candidates
//   3b. select_stream<KLIST, CAP> sorted partial lists (per CTA / per rank) streamed through one selector -> global best k
//   4. pooled floors            lane_kth_of_pool / pooled_floor_batch8 / pooled_kth_key / pooled_max_kth: the floor a
//                               refresh derives from the published keys is reached by >= k distinct published keys
//                               (the exactness invariant of the admission floor), and the bisection variants return
//                               EXACTLY the k-th largest published key, ties on the score word included
// Prints one line per group and exits non-zero on the first violated property.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <vector>

#include <cuda_runtime.h>   // the stub

#include "pool_floor.cuh"
#include "topk.cuh"

using namespace crag;

static std::mt19937_64 rng(20260924);

#define REQUIRE(cond, ...)                              \
  do {                                                  \
    if (!(cond)) {                                      \
      fprintf(stderr, "FAILED %s:%d: %s\n  ", __FILE__, __LINE__, #cond); \
      fprintf(stderr, __VA_ARGS__);                     \
      fprintf(stderr, "\n");                            \
      exit(1);                                          \
    }                                                   \
  } while (0)

// a key as the kernel builds it: distinct rows -> distinct keys; `score_levels` small => many ties on the score word
static uint64_t random_key(int score_levels, uint32_t row) {
  const float s = score_levels > 0 ? float(int(rng() % score_levels)) / float(score_levels) - 0.25f
                                   : float(double(rng() % 2000001) / 1e6 - 1.0);
  return make_key(s, row);
}

static std::vector<uint64_t> distinct_keys(int n, int score_levels) {
  std::vector<uint32_t> rows(n);
  for (int i = 0; i < n; ++i) rows[i] = uint32_t(rng() % 4000000000u);
  std::sort(rows.begin(), rows.end());
  rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
  while (int(rows.size()) < n) rows.push_back(rows.back() + 1 + uint32_t(rows.size()));
  std::shuffle(rows.begin(), rows.end(), rng);
  std::vector<uint64_t> k(n);
  for (int i = 0; i < n; ++i) k[i] = random_key(score_levels, rows[i]);
  return k;
}

// ---------------------------------------------------------------------------------------------- 1. warp_sort_desc
template <int EPL>
static void test_sort(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    std::vector<uint64_t> in(32 * EPL);
    const int levels = (r % 3 == 0) ? 4 : 0;
    for (auto& x : in) x = (rng() % 7 == 0) ? 0ull : random_key(levels, uint32_t(rng() % 1000));   // duplicates allowed here
    std::vector<uint64_t> out(in.size());
    warp_emu::run_warp([&](int lane) {
      uint64_t v[EPL];
      for (int j = 0; j < EPL; ++j) v[j] = in[lane * EPL + j];
      warp_sort_desc<EPL>(v, lane);
      for (int j = 0; j < EPL; ++j) out[lane * EPL + j] = v[j];
    });
    std::vector<uint64_t> want = in;
    std::sort(want.begin(), want.end(), std::greater<uint64_t>());
    REQUIRE(out == want, "warp_sort_desc<%d> round %d", EPL, r);
  }
  printf("ok  warp_sort_desc<%d>: %d random inputs\n", EPL, rounds);
}

// ------------------------------------------------------------------------------------------ 2./3. flush / insert
struct ListCase {
  std::vector<uint64_t> mem;    // KLIST + CAP keys as the kernel holds them in shared memory
  std::vector<uint64_t> want;   // expected list area afterwards
  uint64_t want_thr;
};

template <int KLIST, int CAP>
static ListCase make_case(int k, int held, int c, int levels, bool sorted_list) {
  ListCase cs;
  cs.mem.assign(KLIST + CAP, 0ull);
  std::vector<uint64_t> keys = distinct_keys(held + c + 8, levels);
  std::vector<uint64_t> list(keys.begin(), keys.begin() + held);
  if (sorted_list) std::sort(list.begin(), list.end(), std::greater<uint64_t>());
  for (int i = 0; i < held; ++i) cs.mem[i] = list[i];
  for (int i = 0; i < c; ++i) cs.mem[KLIST + i] = keys[held + i];
  for (int i = c; i < CAP; ++i) cs.mem[KLIST + i] = keys[held + c + (i % 8)] | 1ull << 63;   // stale garbage beyond c: must be ignored
  std::vector<uint64_t> all(keys.begin(), keys.begin() + held + c);
  std::sort(all.begin(), all.end(), std::greater<uint64_t>());
  cs.want.assign(KLIST, 0ull);
  for (int i = 0; i < k && i < int(all.size()); ++i) cs.want[i] = all[i];
  cs.want_thr = int(all.size()) >= k ? all[k - 1] : 0ull;
  return cs;
}

template <int KLIST, int CAP>
static void test_flush(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int k = 1 + int(rng() % KLIST);
    const int held = int(rng() % (k + 1));                    // the list never holds more than k keys
    const int c = CAP ? int(rng() % (CAP + 1)) : 0;
    ListCase cs = make_case<KLIST, CAP>(k, held, c, (r % 4 == 0) ? 3 : 0, true);
    uint64_t thr = 0xdeadbeefull;
    warp_emu::run_warp([&](int lane) { flush_query<KLIST, CAP>(cs.mem.data(), c, k, &thr, lane); });
    REQUIRE(std::equal(cs.want.begin(), cs.want.end(), cs.mem.begin()), "flush_query<%d,%d> k=%d held=%d c=%d", KLIST, CAP, k, held, c);
    REQUIRE(thr == cs.want_thr || (cs.want_thr == 0 && thr == 0), "flush_query<%d,%d> thr k=%d held=%d c=%d", KLIST, CAP, k, held, c);
  }
  printf("ok  flush_query<%d, %d>: %d random (k, list, candidates)\n", KLIST, CAP, rounds);
}

// the first tile of a pass: 128 unsorted keys (some empty) sit in slots 0..127 and one flush builds the list
template <int KLIST>
static void test_direct_first(int rounds) {
  constexpr int CAPD = 128 - KLIST;
  for (int r = 0; r < rounds; ++r) {
    const int k = 1 + int(rng() % KLIST);
    const int live = int(rng() % 129);
    std::vector<uint64_t> keys = distinct_keys(128, (r % 3 == 0) ? 2 : 0);
    std::vector<uint64_t> mem(KLIST + (CAPD ? CAPD : 0) + 128, 0ull);
    std::vector<int> pos(128);
    for (int i = 0; i < 128; ++i) pos[i] = i;
    std::shuffle(pos.begin(), pos.end(), rng);
    std::vector<uint64_t> all;
    for (int i = 0; i < live; ++i) { mem[pos[i]] = keys[i]; all.push_back(keys[i]); }
    std::sort(all.begin(), all.end(), std::greater<uint64_t>());
    uint64_t thr = 1;
    warp_emu::run_warp([&](int lane) { flush_query<KLIST, CAPD>(mem.data(), CAPD, k, &thr, lane); });
    for (int i = 0; i < KLIST; ++i) {
      const uint64_t w = (i < k && i < int(all.size())) ? all[i] : 0ull;
      REQUIRE(mem[i] == w, "direct first tile KLIST=%d k=%d live=%d slot %d", KLIST, k, live, i);
    }
    REQUIRE(thr == (int(all.size()) >= k ? all[k - 1] : 0ull), "direct first tile thr KLIST=%d k=%d live=%d", KLIST, k, live);
  }
  printf("ok  first-tile flush_query<%d, %d>: %d random tiles\n", KLIST, 128 - KLIST, rounds);
}

template <int KLIST, int CAP>
static void test_insert_few(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int k = 1 + int(rng() % KLIST);
    const int held = int(rng() % (k + 1));
    const int c = 1 + int(rng() % kInsertMax);
    ListCase cs = make_case<KLIST, CAP>(k, held, c, (r % 4 == 0) ? 3 : 0, true);
    uint64_t thr = 0xdeadbeefull;
    warp_emu::run_warp([&](int lane) { insert_few<KLIST, CAP>(cs.mem.data(), c, k, &thr, lane); });
    REQUIRE(std::equal(cs.want.begin(), cs.want.end(), cs.mem.begin()), "insert_few<%d,%d> k=%d held=%d c=%d", KLIST, CAP, k, held, c);
    REQUIRE(thr == cs.want_thr, "insert_few<%d,%d> thr k=%d held=%d c=%d", KLIST, CAP, k, held, c);
  }
  printf("ok  insert_few<%d, %d>: %d random (k, list, <= %d candidates)\n", KLIST, CAP, rounds, kInsertMax);
}

// ------------------------------------------------------------------------------------------------ select_stream
// the core of merge_topk_kernel / finalize_exchange_kernel: `parts` sorted lists of k keys each (the per-CTA or per-rank
// partial results, zero padded when a part holds fewer) streamed through one selector, pre-filtered by the largest k-th
// key of any part
template <int KLIST, int CAP>
static void test_select_stream(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int k = 1 + int(rng() % KLIST);
    const int parts = 1 + int(rng() % 40);
    const int levels = (r % 3 == 0) ? 3 : 0;
    std::vector<uint64_t> pool = distinct_keys(parts * k, levels);
    std::vector<uint64_t> lists(size_t(parts) * k, 0ull), all;
    uint64_t bound = 0;
    for (int p = 0; p < parts; ++p) {
      const int have = (rng() % 4 == 0) ? int(rng() % (k + 1)) : k;        // some parts saw fewer than k rows
      std::vector<uint64_t> mine(pool.begin() + size_t(p) * k, pool.begin() + size_t(p) * k + have);
      std::sort(mine.begin(), mine.end(), std::greater<uint64_t>());
      for (int j = 0; j < have; ++j) { lists[size_t(p) * k + j] = mine[j]; all.push_back(mine[j]); }
      if (have == k) bound = std::max(bound, mine[k - 1]);
    }
    if (r % 5 == 4) bound = 0;                                             // PAIRS mode streams without a bound
    std::sort(all.begin(), all.end(), std::greater<uint64_t>());
    std::vector<uint64_t> keys(KLIST + CAP, ~0ull);
    uint64_t thr = ~0ull;
    warp_emu::run_warp([&](int lane) {
      select_stream<KLIST, CAP>(keys.data(), &thr, lane, k, parts * k, bound, [&](int idx) -> uint64_t { return lists[idx]; });
    });
    for (int j = 0; j < KLIST; ++j) {
      const uint64_t w = (j < k && j < int(all.size())) ? all[j] : 0ull;
      REQUIRE(keys[j] == w, "select_stream<%d,%d> k=%d parts=%d rank %d", KLIST, CAP, k, parts, j);
    }
  }
  printf("ok  select_stream<%d, %d>: %d random sets of sorted partial lists\n", KLIST, CAP, rounds);
}

// ---------------------------------------------------------------------------------------------- 4. pooled floors
struct Pool {
  int n_ctas;
  std::vector<uint64_t> t;   // [cta][slot][q]
  uint64_t& at(int c, int m, int q) { return t[(size_t(c) * kPoolSlots + m) * kNQ + q]; }
};

// what the CTAs would have published: per CTA and query its best kPoolM keys (descending; possibly fewer or none) and
// its own k-th key in slot kPoolM (0 or a key not above its 4th best); all keys of one query distinct
static Pool make_pool(int n_ctas, int levels, double p_empty) {
  Pool p;
  p.n_ctas = n_ctas;
  p.t.assign(size_t(n_ctas) * kPoolSlots * kNQ, 0ull);
  for (int q = 0; q < kNQ; ++q) {
    std::vector<uint64_t> keys = distinct_keys(n_ctas * (kPoolM + 1), levels);
    for (int c = 0; c < n_ctas; ++c) {
      if (double(rng() % 1000) / 1000.0 < p_empty) continue;
      std::vector<uint64_t> mine(keys.begin() + c * (kPoolM + 1), keys.begin() + (c + 1) * (kPoolM + 1));
      std::sort(mine.begin(), mine.end(), std::greater<uint64_t>());
      const int have = 1 + int(rng() % kPoolM);
      for (int m = 0; m < have; ++m) p.at(c, m, q) = mine[m];
      if (have == kPoolM && rng() % 2) p.at(c, kPoolM, q) = mine[kPoolM];
    }
  }
  return p;
}

// the DISTINCT rows the pool vouches for (a torn read can show one row of a CTA in two of its slots)
static std::vector<uint64_t> published(Pool& p, int q, int slots) {
  std::vector<uint64_t> v;
  for (int c = 0; c < p.n_ctas; ++c)
    for (int m = 0; m < slots; ++m)
      if (p.at(c, m, q)) v.push_back(p.at(c, m, q));
  std::sort(v.begin(), v.end(), std::greater<uint64_t>());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}
// what a reader can see when a CTA publishes between its loads of slot m and slot m + 1: the old key of slot m, then
// the new key of slot m + 1 -- which is that same row, moved down by the better row that arrived
static void tear_some_entries(Pool& p, int q, int how_many) {
  for (int t = 0; t < how_many; ++t) {
    const int c = int(rng() % p.n_ctas), m = int(rng() % (kPoolM - 1));
    if (p.at(c, m, q)) p.at(c, m + 1, q) = p.at(c, m, q);
  }
}
static int count_ge(const std::vector<uint64_t>& v, uint64_t f) {
  return int(std::count_if(v.begin(), v.end(), [&](uint64_t x) { return x >= f; }));
}

static void test_small_k_floor(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int n_ctas = (r % 5 == 0) ? 148 : 1 + int(rng() % kPoolMaxCtas);
    const int k = 1 + int(rng() % kPoolSmallK);
    const int kp = (k + 3) / 4;
    Pool p = make_pool(n_ctas, (r % 3 == 0) ? 3 : 0, (r % 4 == 0) ? 0.5 : 0.02);
    uint64_t part[4][kNQ];
    for (int w = 0; w < 4; ++w)     // the four select warps, each over the CTAs c = w (mod 4); lane = query
      warp_emu::run_warp([&](int lane) { part[w][lane] = lane_kth_of_pool<4>(p.t.data(), n_ctas, w, lane, 1, kp); });
    for (int q = 0; q < kNQ; ++q) {
      uint64_t floor = part[0][q];
      for (int w = 1; w < 4; ++w) floor = std::min(floor, part[w][q]);
      std::vector<uint64_t> best = published(p, q, 1);
      if (floor) REQUIRE(count_ge(best, floor) >= k, "small-k floor admits too few: k=%d ctas=%d q=%d: %d keys >= floor", k, n_ctas, q, count_ge(best, floor));
      for (int w = 0; w < 4; ++w) {                 // each warp's part is exactly the kp-th largest of its quarter
        std::vector<uint64_t> quarter;
        for (int c = w; c < n_ctas; c += 4) if (p.at(c, 0, q)) quarter.push_back(p.at(c, 0, q));
        std::sort(quarter.begin(), quarter.end(), std::greater<uint64_t>());
        const uint64_t want = int(quarter.size()) >= kp ? quarter[kp - 1] : 0ull;
        REQUIRE(part[w][q] == want, "lane_kth_of_pool k=%d kp=%d ctas=%d w=%d q=%d", k, kp, n_ctas, w, q);
      }
    }
  }
  printf("ok  lane_kth_of_pool<4> (k <= %d): %d random pools, floor reached by >= k published keys\n", kPoolSmallK, rounds);
}

static void test_batch8_floor(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int n_ctas = (r % 3 == 0) ? 148 : 20 + int(rng() % (kPoolMaxCtas - 19));
    const int k = kPoolSmallK + 1 + int(rng() % std::max(1, 4 * n_ctas / 5 - kPoolSmallK));
    const int nq = (r % 4 == 0) ? 1 + int(rng() % kNQ) : kNQ;
    const int levels = (r % 2 == 0) ? 2 + int(rng() % 3) : 0;        // every other pool: massive ties on the score word
    Pool p = make_pool(n_ctas, levels, (r % 5 == 0) ? 0.4 : 0.0);
    for (int ew = 0; ew < 4; ++ew) {
      uint64_t out[8];
      uint32_t ties = 0;
      warp_emu::run_warp([&](int lane) {
        uint64_t o[8];
        const uint32_t t = pooled_floor_batch8(p.t.data(), n_ctas, ew, nq, k, lane, o);
        if (lane == 0) { ties = t; for (int j = 0; j < 8; ++j) out[j] = o[j]; }
      });
      for (int j = 0; j < 8; ++j) {
        const int q = ew + 4 * j;
        if (q >= nq) { REQUIRE(out[j] == 0ull, "batch8: query %d >= nq %d got a floor", q, nq); continue; }
        std::vector<uint64_t> best = published(p, q, 1);
        const uint64_t want = int(best.size()) >= k ? best[k - 1] : 0ull;
        REQUIRE(out[j] == want, "pooled_floor_batch8 k=%d ctas=%d q=%d levels=%d: got %llx want %llx", k, n_ctas, q, levels,
                (unsigned long long)out[j], (unsigned long long)want);
        if (want) {
          const int same_score = int(std::count_if(best.begin(), best.end(), [&](uint64_t x) { return (x >> 32) == (want >> 32); }));
          REQUIRE(((ties >> j) & 1u) == (same_score > 1 ? 1u : 0u), "batch8 tie flag k=%d q=%d", k, q);
        }
      }
    }
  }
  printf("ok  pooled_floor_batch8 (16 < k <= 0.8 x CTAs): %d random pools, floor == k-th largest best key, tie flags exact\n", rounds);
}

static void test_all_keys_floor(int rounds) {
  for (int r = 0; r < rounds; ++r) {
    const int n_ctas = (r % 3 == 0) ? 148 : 8 + int(rng() % (kPoolMaxCtas - 7));
    const int k = 1 + int(rng() % 128);
    const int levels = (r % 2 == 0) ? 2 + int(rng() % 3) : 0;
    Pool p = make_pool(n_ctas, levels, (r % 5 == 0) ? 0.3 : 0.0);
    const int q = int(rng() % kNQ);
    if (r % 2) tear_some_entries(p, q, 1 + int(rng() % 24));
    uint64_t got = 1, got_max = 1;
    warp_emu::run_warp([&](int lane) {
      const uint64_t f = pooled_kth_key(p.t.data(), n_ctas, q, k, lane);
      const uint64_t m = pooled_max_kth(p.t.data(), n_ctas, q, lane);
      if (lane == 7) { got = f; got_max = m; }
    });
    std::vector<uint64_t> all = published(p, q, kPoolM);
    REQUIRE(got == (int(all.size()) >= k ? all[k - 1] : 0ull), "pooled_kth_key k=%d ctas=%d q=%d levels=%d", k, n_ctas, q, levels);
    uint64_t want_max = 0;
    for (int c = 0; c < n_ctas; ++c) want_max = std::max(want_max, p.at(c, kPoolM, q));
    REQUIRE(got_max == want_max, "pooled_max_kth ctas=%d q=%d", n_ctas, q);
  }
  printf("ok  pooled_kth_key / pooled_max_kth: %d random pools\n", rounds);
}

int main(int argc, char** argv) {
  const int scale = argc > 1 ? atoi(argv[1]) : 1;
  test_sort<1>(20 * scale);
  test_sort<2>(20 * scale);
  test_sort<4>(20 * scale);
  test_sort<8>(20 * scale);
  test_flush<32, 32>(60 * scale);
  test_flush<64, 64>(60 * scale);
  test_flush<128, 128>(60 * scale);
  test_direct_first<64>(30 * scale);
  test_direct_first<128>(30 * scale);
  test_insert_few<64, 64>(60 * scale);
  test_insert_few<128, 128>(60 * scale);
  test_select_stream<32, 32>(40 * scale);
  test_select_stream<64, 64>(40 * scale);
  test_select_stream<128, 128>(40 * scale);
  test_small_k_floor(12 * scale);
  test_batch8_floor(12 * scale);
  test_all_keys_floor(40 * scale);
  printf("ALL OK\n");
  return 0;
}
