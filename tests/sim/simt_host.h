// tests/sim: a 32-lane SIMT emulator for the HOST, so that the warp kernels of breakmer_b200/csrc/nw.cuh -- the very
// source nvcc compiles for sm_100a -- can be run and checked in the build container, which has no GPU.
//
// TEST TOOL ONLY (compiled with -DBK_SIMT by tests/test_simt_nw.py).  It is not part of the product library, is never
// loaded by breakmer_b200 and is not a fallback: the product fails without a CUDA device.
//
// Model: every thread of a block (up to 8 warps of 32 lanes) is a fiber (ucontext) running the same function.  Threads
// run one after the other until they reach a warp collective (__shfl_*_sync, __ballot_sync, __syncwarp) -- a barrier
// across the 32 fibers of that warp with an exchange buffer -- or __syncthreads(), a barrier across the block.  Kernels that are correct under this model are race free at warp level as long as
// every cross-lane communication goes through a collective or is separated by __syncwarp() -- which is exactly what is
// worth checking: a missing __syncwarp() shows up as a wrong result here (lanes run to the next collective one at a
// time, in lane order, the most adversarial interleaving for producer/consumer code).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ucontext.h>

#include <functional>
#include <mutex>

// SIMT_TSAN: the emulator under ThreadSanitizer.  Every emulated thread is a TSan fiber, fiber switches carry NO
// synchronisation, and the only happens-before edges are the ones the kernel asks for: warp collectives, __syncthreads,
// kernel boundaries (atomics are real atomics).  TSan then reports what CUDA calls a race: two threads of a block
// touching the same address, at least one writing, with no barrier in between.
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define SIMT_ASAN 1
#endif
#ifdef SIMT_TSAN
#include <sanitizer/tsan_interface.h>
#define SIMT_NO_TSAN __attribute__((no_sanitize("thread")))
#else
#define SIMT_NO_TSAN
#endif

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
constexpr int MAX_WARPS = 32;
constexpr size_t STACK_BYTES = 512 * 1024;

struct WarpState {               // collective in progress on one warp
  int lanes = LANES;             // threads of this warp (the last warp of a block may be partial)
  int arrived = 0;
  unsigned generation = 0;
  int op = 0;                    // kind of the collective (all lanes must agree)
  long long buf[LANES];
};

// Fiber switch.  glibc's swapcontext saves and restores the signal mask with two system calls per switch, which is most
// of the emulator's run time; on x86-64 a switch is therefore a dozen instructions of our own (callee-saved registers
// and the stack pointer), ucontext elsewhere.
#if defined(__x86_64__)
#define SIMT_ASM_SWITCH 1
extern "C" void simt_switch(void** save_sp, void* load_sp);
__asm__(
    ".text\n"
    ".globl simt_switch\n"
    ".hidden simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch,.-simt_switch\n");
typedef void* Ctx;
#else
typedef ucontext_t Ctx;
#endif

struct Block {
  int n_threads = 0;
  unsigned block_idx = 0, grid_dim = 1;
  Ctx sched;
  Ctx ctx[MAX_WARPS * LANES];
  bool done[MAX_WARPS * LANES];
  char* stacks = nullptr;
  int cur = 0;
  WarpState warp[MAX_WARPS];
  int b_arrived = 0;             // __syncthreads
  unsigned b_generation = 0;
  unsigned long long progress = 0;
#ifdef SIMT_TSAN
  void* tsan_fiber[MAX_WARPS * LANES];
  void* tsan_sched = nullptr;
  // addresses TSan hangs the happens-before edges on.  Two per barrier, alternating with its generation: a thread that
  // is still waking from barrier n must not acquire what faster threads release when they arrive at barrier n + 1
  char sync_launch = 0, sync_done = 0, sync_block[2] = {0, 0};
  char sync_warp[MAX_WARPS][2];
#endif
  void* where[MAX_WARPS * LANES];   // return address of the barrier each thread waits at (deadlock report)
  unsigned n_warp_bar[MAX_WARPS * LANES] = {0}, n_block_bar[MAX_WARPS * LANES] = {0};
  void* hist[MAX_WARPS * LANES][16];
  std::function<void()> body;
};

SIMT_NO_TSAN inline Block*& current() {
  static Block* b = nullptr;       // (guarded by device_mutex: one block at a time)
  return b;
}

inline void report(Block* b) {
  fprintf(stderr, "simt: %d threads at __syncthreads;", b->b_arrived);
  for (int w = 0; w < b->n_threads / LANES; ++w)
    fprintf(stderr, " warp %d: %d lanes at collective %d;", w, b->warp[w].arrived, b->warp[w].op);
  fprintf(stderr, "\n");
  for (int t = 0; t < b->n_threads; ++t)
    if (!b->done[t] && (t == 0 || b->where[t] != b->where[t - 1]))
      fprintf(stderr, "simt:   thread %d waits at %p (its %u-th warp collective, %u-th __syncthreads)\n", t, b->where[t],
              b->n_warp_bar[t], b->n_block_bar[t]);
  if (getenv("SIMT_HISTORY"))
    for (int t = 0; t < b->n_threads; ++t)
      if (!b->done[t] && (t == 0 || b->where[t] != b->where[t - 1])) {
        fprintf(stderr, "simt:   thread %d last warp collectives (oldest first):", t);
        for (unsigned i = b->n_warp_bar[t] > 16 ? b->n_warp_bar[t] - 16 : 0; i < b->n_warp_bar[t]; ++i) fprintf(stderr, " %p", b->hist[t][i & 15]);
        fprintf(stderr, "\n");
      }
}

SIMT_NO_TSAN inline void switch_ctx(Ctx* from, Ctx* to) {
#ifdef SIMT_ASM_SWITCH
  simt_switch(from, *to);
#else
  swapcontext(from, to);
#endif
}

#ifdef SIMT_ASAN
// AddressSanitizer has to be told about every stack switch (stack bounds of the destination)
struct AsanStacks { const void* sched_bottom = nullptr; size_t sched_size = 0; };
inline AsanStacks& asan_stacks() { static AsanStacks a; return a; }
#endif

SIMT_NO_TSAN inline void trampoline() {
  Block* b = current();
  const int t = b->cur;
#ifdef SIMT_ASAN
  __sanitizer_finish_switch_fiber(nullptr, &asan_stacks().sched_bottom, &asan_stacks().sched_size);
#endif
#ifdef SIMT_TSAN
  __tsan_acquire(&b->sync_launch);           // everything before the launch happens before the kernel's threads
#endif
  b->body();
#ifdef SIMT_TSAN
  __tsan_release(&b->sync_done);             // and the threads happen before whatever follows the block
#endif
  b->done[t] = true;
  ++b->progress;
#ifdef SIMT_TSAN
  __tsan_switch_to_fiber(b->tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
#ifdef SIMT_ASAN
  __sanitizer_start_switch_fiber(nullptr, asan_stacks().sched_bottom, asan_stacks().sched_size);   // this fiber is done
#endif
  switch_ctx(&b->ctx[t], &b->sched);
  abort();                         // a finished fiber is never resumed
}

// run `body()` on n_warps x 32 emulated threads of one block until all of them have returned
// stacks of the fibers: one pool, grown on demand and kept (blocks run one at a time)
inline uint8_t* dyn_smem();
inline char* stack_pool(size_t bytes) {
  static char* pool = nullptr;
  static size_t cap = 0;
  if (bytes > cap) {
    free(pool);
    pool = (char*)malloc(bytes);
    cap = bytes;
  }
  return pool;
}

// one block runs at a time, process-wide: the stack pool, the dynamic shared memory and the function-scope __shared__
// variables are single copies (host threads that launch concurrently -- two DevicePipelines on one device -- take turns)
inline std::recursive_mutex& device_mutex() {
  static std::recursive_mutex m;
  return m;
}

inline void run_threads(int n_threads, unsigned block_idx, const std::function<void()>& body, unsigned grid_dim) {
  std::lock_guard<std::recursive_mutex> hold(device_mutex());
  if (n_threads < 1 || n_threads > MAX_WARPS * LANES) { fprintf(stderr, "simt: block of %d threads\n", n_threads); abort(); }
  static Block* reuse = nullptr;
  Block* b = reuse ? reuse : (reuse = new Block);
  b->n_threads = n_threads;
  b->block_idx = block_idx;
  b->grid_dim = grid_dim ? grid_dim : block_idx + 1;
  b->body = body;
  b->b_arrived = 0; b->b_generation = 0; b->progress = 0;
  for (int w = 0; w < MAX_WARPS; ++w) {
    b->warp[w] = WarpState();
    const int left = n_threads - w * LANES;
    b->warp[w].lanes = left >= LANES ? LANES : (left > 0 ? left : 0);
  }
  memset(b->n_warp_bar, 0, sizeof b->n_warp_bar);
  memset(b->n_block_bar, 0, sizeof b->n_block_bar);
  b->stacks = stack_pool(STACK_BYTES * (size_t)b->n_threads);
  if (!getenv("SIMT_KEEP_SMEM")) memset(dyn_smem(), 0xEE, 232 * 1024);  // shared memory of a new block holds garbage
#ifdef SIMT_TSAN
  b->tsan_sched = __tsan_get_current_fiber();
  for (int t = 0; t < b->n_threads; ++t) b->tsan_fiber[t] = __tsan_create_fiber(0);
  __tsan_release(&b->sync_launch);
#endif
  Block* prev = current();
  current() = b;
  for (int t = 0; t < b->n_threads; ++t) {
    b->done[t] = false;
#ifdef SIMT_ASM_SWITCH
    // a fresh stack that simt_switch "returns" into trampoline() from: six callee-saved registers, the entry address,
    // and a null return address above it (rsp is 8 mod 16 at the entry, as after a call)
    uintptr_t top = ((uintptr_t)(b->stacks + STACK_BYTES * (t + 1))) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;
    *--sp = (void*)(void (*)())trampoline;
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    b->ctx[t] = (void*)sp;
#else
    getcontext(&b->ctx[t]);
    b->ctx[t].uc_stack.ss_sp = b->stacks + STACK_BYTES * t;
    b->ctx[t].uc_stack.ss_size = STACK_BYTES;
    b->ctx[t].uc_link = &b->sched;
    makecontext(&b->ctx[t], (void (*)())trampoline, 0);
#endif
  }
  // order in which the runnable threads get their turn in a pass: ascending (default), descending, or a fresh
  // pseudo-random permutation per pass (SIMT_ORDER=reverse|random[:seed]) -- different interleavings of the same kernel
  const char* ord = getenv("SIMT_ORDER");
  const int mode = !ord ? 0 : (ord[0] == 'r' && ord[1] == 'e') ? 1 : 2;
  unsigned long long rng = 0x9E3779B97F4A7C15ull ^ (ord && strchr(ord, ':') ? strtoull(strchr(ord, ':') + 1, nullptr, 10) : 1);
  int order[MAX_WARPS * LANES];
  for (int t = 0; t < b->n_threads; ++t) order[t] = mode == 1 ? b->n_threads - 1 - t : t;
  for (;;) {
    int live = 0;
    const unsigned long long before = b->progress;
    if (mode == 2)
      for (int t = b->n_threads - 1; t > 0; --t) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        const int j = (int)(rng % (unsigned)(t + 1));
        const int tmp = order[t]; order[t] = order[j]; order[j] = tmp;
      }
    for (int oi = 0; oi < b->n_threads; ++oi) {
      const int t = order[oi];
      if (b->done[t]) continue;
      b->cur = t;
#ifdef SIMT_TSAN
      __tsan_switch_to_fiber(b->tsan_fiber[t], __tsan_switch_to_fiber_no_sync);
#endif
#ifdef SIMT_ASAN
      void* fake = nullptr;
      __sanitizer_start_switch_fiber(&fake, b->stacks + STACK_BYTES * (size_t)t, STACK_BYTES);
#endif
      switch_ctx(&b->sched, &b->ctx[t]);
#ifdef SIMT_ASAN
      __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
      if (!b->done[t]) ++live;
    }
    if (live == 0) break;
    if (b->progress == before) {
      fprintf(stderr, "simt: deadlock -- %d threads wait at barriers that the others never reach "
                      "(a collective inside divergent control flow, or threads that left the kernel early)\n", live);
      report(b);
      abort();
    }
  }
#ifdef SIMT_TSAN
  __tsan_acquire(&b->sync_done);
  for (int t = 0; t < b->n_threads; ++t) __tsan_destroy_fiber(b->tsan_fiber[t]);
#endif
  current() = prev;
}

inline void run_block(int n_warps, unsigned block_idx, const std::function<void()>& body, unsigned grid_dim = 0) {
  run_threads(n_warps * LANES, block_idx, body, grid_dim);
}

// kernel<<<grid, block, smem, stream>>>(args): the blocks one after the other (tests/sim/gen_simt_sources.py rewrites
// the launch statements of the product's host code into calls of this)
inline void launch(long long grid, long long block, const std::function<void()>& body) {
  for (long long g = 0; g < grid; ++g) run_threads((int)block, (unsigned)g, body, (unsigned)grid);
}

// one warp: body(lane)
inline int lane_id();
inline void run_warp(const std::function<void(int)>& body) {
  run_block(1, 0, [&]() { body(lane_id()); });
}

SIMT_NO_TSAN inline int thread_id() { return current()->cur; }
SIMT_NO_TSAN inline int lane_id() { return current()->cur & 31; }
SIMT_NO_TSAN inline int warp_id() { return current()->cur >> 5; }

SIMT_NO_TSAN inline void yield_thread() {
  Block* b = current();
#ifdef SIMT_TSAN
  __tsan_switch_to_fiber(b->tsan_sched, __tsan_switch_to_fiber_no_sync);
#endif
#ifdef SIMT_ASAN
  void* fake = nullptr;
  __sanitizer_start_switch_fiber(&fake, asan_stacks().sched_bottom, asan_stacks().sched_size);
#endif
  switch_ctx(&b->ctx[b->cur], &b->sched);
#ifdef SIMT_ASAN
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
}

// barrier across the 32 fibers of the calling warp; `op` identifies the kind of collective so that lanes that
// diverged into different collectives are caught
SIMT_NO_TSAN __attribute__((noinline)) inline void barrier(int op) {
  Block* b = current();
  b->where[b->cur] = __builtin_return_address(0);
  b->hist[b->cur][b->n_warp_bar[b->cur] & 15] = __builtin_return_address(0);
  ++b->n_warp_bar[b->cur];
  WarpState* w = &b->warp[b->cur >> 5];
  if (w->arrived == 0) w->op = op;
  else if (w->op != op) {
    fprintf(stderr, "simt: lanes of a warp diverged into different collectives (%d vs %d; thread %d arrives at %p)\n", w->op, op, b->cur,
            __builtin_return_address(0));
    report(b);
    abort();
  }
  const unsigned gen = w->generation;
#ifdef SIMT_TSAN
  char* sync = &b->sync_warp[b->cur >> 5][gen & 1];
  __tsan_release(sync);
#endif
  if (++w->arrived == w->lanes) {
    w->arrived = 0;
    ++w->generation;
    ++b->progress;
  } else {
    while (w->generation == gen) yield_thread();
  }
#ifdef SIMT_TSAN
  __tsan_acquire(sync);
#endif
}

SIMT_NO_TSAN __attribute__((noinline)) inline void block_barrier() {
  Block* b = current();
  b->where[b->cur] = __builtin_return_address(0);
  ++b->n_block_bar[b->cur];
  const unsigned gen = b->b_generation;
#ifdef SIMT_TSAN
  char* sync = &b->sync_block[gen & 1];
  __tsan_release(sync);
#endif
  if (++b->b_arrived == b->n_threads) {
    b->b_arrived = 0;
    ++b->b_generation;
    ++b->progress;
  } else {
    while (b->b_generation == gen) yield_thread();
  }
#ifdef SIMT_TSAN
  __tsan_acquire(sync);
#endif
}

template <typename T>
SIMT_NO_TSAN inline T exchange(T v, int src_lane, int op) {
  static_assert(sizeof(T) <= sizeof(long long), "exchange type too wide");
  Block* b = current();
  WarpState* w = &b->warp[b->cur >> 5];
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->buf[b->cur & 31] = raw;
  barrier(op);
  T out = v;
  if (src_lane >= 0 && src_lane < LANES) memcpy(&out, &w->buf[src_lane], sizeof(T));
  barrier(op + 1);                 // nobody overwrites the buffer before everybody has read it
  return out;
}

struct Dim { unsigned x; };
SIMT_NO_TSAN inline Dim thread_idx() { return Dim{(unsigned)current()->cur}; }
SIMT_NO_TSAN inline Dim block_idx() { return Dim{current()->block_idx}; }
SIMT_NO_TSAN inline Dim block_dim() { return Dim{(unsigned)current()->n_threads}; }
SIMT_NO_TSAN inline Dim grid_dim() { return Dim{current()->grid_dim}; }

// the dynamic shared memory of the block that is running (blocks run one at a time)
inline uint8_t* dyn_smem() {
  alignas(16) static uint8_t buf[256 * 1024];
  return buf;
}

}  // namespace simt

#define threadIdx (simt::thread_idx())
#define blockIdx (simt::block_idx())
#define blockDim (simt::block_dim())
#define gridDim (simt::grid_dim())
#define __shared__ static          // a function-scope __shared__ variable: one copy for all fibers of the (single) running block
#define __align__(n)
#define __launch_bounds__(...)

// ---- the intrinsics nw.cuh uses --------------------------------------------------------------------------------
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src & 31, 10); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int l = simt::lane_id();
  return simt::exchange(v, l - (int)d >= 0 ? l - (int)d : l, 20);
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, simt::lane_id() ^ m, 30); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const int l = simt::lane_id();
  return simt::exchange(v, l + (int)d < simt::LANES ? l + (int)d : l, 60);
}
SIMT_NO_TSAN inline unsigned __ballot_sync(unsigned, bool p) {
  simt::Block* b = simt::current();
  simt::WarpState* w = &b->warp[b->cur >> 5];
  w->buf[b->cur & 31] = p ? 1 : 0;
  simt::barrier(40);
  unsigned m = 0;
  for (int l = 0; l < w->lanes; ++l) m |= (unsigned)(w->buf[l] & 1) << l;
  simt::barrier(41);
  return m;
}
template <typename T> SIMT_NO_TSAN inline unsigned __match_any_sync(unsigned, T v) {
  static_assert(sizeof(T) <= sizeof(long long), "match type too wide");
  simt::Block* b = simt::current();
  simt::WarpState* w = &b->warp[b->cur >> 5];
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->buf[b->cur & 31] = raw;
  simt::barrier(70);
  unsigned m = 0;
  for (int l = 0; l < w->lanes; ++l) m |= (unsigned)(w->buf[l] == raw) << l;
  simt::barrier(71);
  return m;
}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier(50); }
inline void __syncthreads() { simt::block_barrier(); }
// (threads run one at a time between barriers, so plain read-modify-write is atomic here; under SIMT_TSAN they are real
// atomics so that TSan knows them as such)
#ifdef SIMT_TSAN
template <typename T, typename V> inline T atomicAdd(T* p, V v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename V> inline T atomicOr(T* p, V v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_RELAXED); }
template <typename T, typename V> inline T atomicMax(T* p, V v) {
  T o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while ((T)v > o && !__atomic_compare_exchange_n(p, &o, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
template <typename T, typename V> inline T atomicCAS(T* p, V cmp, V val) {
  T o = (T)cmp;
  __atomic_compare_exchange_n(p, &o, (T)val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return o;
}
#else
template <typename T, typename V> inline T atomicAdd(T* p, V v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename V> inline T atomicOr(T* p, V v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename V> inline T atomicMax(T* p, V v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename V> inline T atomicCAS(T* p, V cmp, V val) { const T o = *p; if (o == (T)cmp) *p = (T)val; return o; }
#endif
inline void __threadfence_block() {}
inline void __threadfence() {}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline float __log2f(float x) { return log2f(x); }
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
