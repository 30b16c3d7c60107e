// 32 emulated lanes of one warp on the CPU: each lane is a ucontext fiber, a warp-collective (shuffle, reduction,
// ballot, __syncwarp) is a barrier at which the fibers hand values over through a slot array.  Lanes run one at a time
// and only switch at collectives, so plain shared arrays behave like shared memory between __syncwarp()s.  Every lane
// must reach every collective (as on the GPU with a full mask); a lane that exits while others wait is reported as a
// deadlock instead of hanging the test.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

namespace warp_emu {

constexpr int kLanes = 32;
constexpr size_t kStackBytes = 512 << 10;

struct Warp {
  ucontext_t main, ctx[kLanes];
  char* stacks[kLanes];
  bool done[kLanes];
  int cur = 0;
  int arrived = 0;
  uint64_t gen = 0;
  uint64_t slot[kLanes];
  std::function<void(int)> body;
};

inline Warp*& current() {
  static Warp* w = nullptr;
  return w;
}
inline int lane() { return current()->cur; }

inline void yield_lane() {
  Warp* w = current();
  swapcontext(&w->ctx[w->cur], &w->main);
}

inline void barrier() {
  Warp* w = current();
  const uint64_t my = w->gen;
  if (++w->arrived == kLanes) {
    w->arrived = 0;
    ++w->gen;
  }
  while (w->gen == my) yield_lane();
}

// every lane deposits v; lane l receives the value of lane src_of(l)
template <class T, class F>
inline T exchange(T v, F src_of) {
  static_assert(sizeof(T) <= 8, "exchange moves at most 64 bits");
  Warp* w = current();
  const int me = w->cur;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w->slot[me] = bits;
  barrier();
  const uint64_t got = w->slot[src_of(me) & (kLanes - 1)];
  barrier();                       // nobody overwrites a slot before every lane has read
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}

template <class T, class F>
inline T reduce(T v, F op) {
  Warp* w = current();
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w->slot[w->cur] = bits;
  barrier();
  T acc;
  memcpy(&acc, &w->slot[0], sizeof(T));
  for (int l = 1; l < kLanes; ++l) {
    T x;
    memcpy(&x, &w->slot[l], sizeof(T));
    acc = op(acc, x);
  }
  barrier();
  return acc;
}

inline void trampoline() {
  Warp* w = current();
  w->body(w->cur);
  w->done[w->cur] = true;
}

// run body(lane) on 32 lanes to completion
inline void run_warp(const std::function<void(int)>& body) {
  Warp* w = new Warp;
  Warp* outer = current();
  current() = w;
  w->body = body;
  for (int l = 0; l < kLanes; ++l) {
    w->done[l] = false;
    w->stacks[l] = static_cast<char*>(malloc(kStackBytes));
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = w->stacks[l];
    w->ctx[l].uc_stack.ss_size = kStackBytes;
    w->ctx[l].uc_link = &w->main;
    makecontext(&w->ctx[l], reinterpret_cast<void (*)()>(trampoline), 0);
  }
  int stuck_sweeps = 0;
  while (true) {
    int alive = 0, finished_before = 0;
    for (int l = 0; l < kLanes; ++l) finished_before += w->done[l];
    const uint64_t gen_before = w->gen;
    const int arrived_before = w->arrived;
    for (int l = 0; l < kLanes; ++l) {
      if (w->done[l]) continue;
      ++alive;
      w->cur = l;
      swapcontext(&w->main, &w->ctx[l]);
    }
    if (alive == 0) break;
    int finished_after = 0;
    for (int l = 0; l < kLanes; ++l) finished_after += w->done[l];
    if (w->gen == gen_before && w->arrived == arrived_before && finished_after == finished_before) {
      if (++stuck_sweeps > 4) {
        fprintf(stderr, "warp_emu: deadlock -- %d lanes wait at a collective that %d finished lanes never reach\n",
                w->arrived, finished_after);
        abort();
      }
    } else {
      stuck_sweeps = 0;
    }
  }
  for (int l = 0; l < kLanes; ++l) free(w->stacks[l]);
  delete w;
  current() = outer;
}

}  // namespace warp_emu
