// tests/sim: a 32-lane SIMT emulator for the HOST, so that the warp kernels of breakmer_b200/csrc/nw.cuh -- the very
// source nvcc compiles for sm_100a -- can be run and checked in the build container, which has no GPU.
//
// TEST TOOL ONLY (compiled with -DBK_SIMT by tests/test_simt_nw.py).  It is not part of the product library, is never
// loaded by breakmer_b200 and is not a fallback: the product fails without a CUDA device.
//
// Model: every lane of the warp is a fiber (ucontext) running the same function.  Lanes run one after the other until
// they reach a warp collective (__shfl_*_sync, __ballot_sync, __syncwarp); a collective is a barrier across the 32
// fibers with an exchange buffer.  Kernels that are correct under this model are race free at warp level as long as
// every cross-lane communication goes through a collective or is separated by __syncwarp() -- which is exactly what is
// worth checking: a missing __syncwarp() shows up as a wrong result here (lanes run to the next collective one at a
// time, in lane order, the most adversarial interleaving for producer/consumer code).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__

struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace simt {

constexpr int LANES = 32;
constexpr size_t STACK_BYTES = 256 * 1024;

struct Warp {
  ucontext_t sched;
  ucontext_t ctx[LANES];
  char* stacks = nullptr;
  bool done[LANES];
  int cur = 0;
  // collective state
  int arrived = 0;
  unsigned generation = 0;
  int op = 0;                      // kind of the collective in progress (all lanes must agree)
  long long buf[LANES];
  std::function<void(int)> body;
};

inline Warp*& current() {
  static thread_local Warp* w = nullptr;
  return w;
}

inline void trampoline() {
  Warp* w = current();
  const int l = w->cur;
  w->body(l);
  w->done[l] = true;
  swapcontext(&w->ctx[l], &w->sched);
}

// run `body(lane)` on 32 emulated lanes until all of them have returned
inline void run_warp(const std::function<void(int)>& body) {
  Warp w;
  w.body = body;
  w.stacks = (char*)malloc(STACK_BYTES * LANES);
  Warp* prev = current();
  current() = &w;
  for (int l = 0; l < LANES; ++l) {
    w.done[l] = false;
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stacks + STACK_BYTES * l;
    w.ctx[l].uc_stack.ss_size = STACK_BYTES;
    w.ctx[l].uc_link = &w.sched;
    makecontext(&w.ctx[l], (void (*)())trampoline, 0);
  }
  int live = LANES;
  while (live > 0) {
    live = 0;
    for (int l = 0; l < LANES; ++l) {
      if (w.done[l]) continue;
      w.cur = l;
      swapcontext(&w.sched, &w.ctx[l]);
      if (!w.done[l]) ++live;
    }
    if (live > 0 && live < LANES && w.arrived > 0 && w.arrived == live) {
      fprintf(stderr, "simt: %d lanes wait at a collective that %d lanes have left the kernel without\n", live, LANES - live);
      abort();
    }
  }
  current() = prev;
  free(w.stacks);
}

inline int lane_id() { return current()->cur; }

inline void yield_lane() {
  Warp* w = current();
  swapcontext(&w->ctx[w->cur], &w->sched);
}

// barrier across the 32 fibers; `op` identifies the call site's kind so that divergent collectives are caught
inline void barrier(int op) {
  Warp* w = current();
  if (w->arrived == 0) w->op = op;
  else if (w->op != op) { fprintf(stderr, "simt: lanes diverged into different collectives (%d vs %d)\n", w->op, op); abort(); }
  const unsigned gen = w->generation;
  if (++w->arrived == LANES) {
    w->arrived = 0;
    ++w->generation;
  } else {
    while (w->generation == gen) yield_lane();
  }
}

template <typename T>
inline T exchange(T v, int src_lane, int op) {
  static_assert(sizeof(T) <= sizeof(long long), "exchange type too wide");
  Warp* w = current();
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->buf[w->cur] = raw;
  barrier(op);
  T out = v;
  if (src_lane >= 0 && src_lane < LANES) memcpy(&out, &w->buf[src_lane], sizeof(T));
  barrier(op + 1);                 // nobody overwrites the buffer before everybody has read it
  return out;
}

}  // namespace simt

// ---- the intrinsics nw.cuh uses --------------------------------------------------------------------------------
struct SimtThreadIdx { int x_get() const { return simt::lane_id(); } };
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src & 31, 10); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int l = simt::lane_id();
  return simt::exchange(v, l - (int)d >= 0 ? l - (int)d : l, 20);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, simt::lane_id() ^ m, 30); }
inline unsigned __ballot_sync(unsigned, bool p) {
  simt::Warp* w = simt::current();
  w->buf[w->cur] = p ? 1 : 0;
  simt::barrier(40);
  unsigned m = 0;
  for (int l = 0; l < simt::LANES; ++l) m |= (unsigned)(w->buf[l] & 1) << l;
  simt::barrier(41);
  return m;
}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier(50); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __vimax3_s32(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
inline int __viaddmax_s32(int a, int b, int c) { const int s = a + b; return s > c ? s : c; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;       // msb replication mode
    r |= b << (8 * i);
  }
  return r;
}
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcg(T* p, T v) { *p = v; }
