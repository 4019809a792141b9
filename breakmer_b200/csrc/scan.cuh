// Device-wide exclusive prefix sum over uint32 (three launches: tile sums, spine,
// apply).  Used for compaction offsets and for the radix sort's digit tables.
// HBM traffic: reads the input twice, writes the output once (12 B/element).
#pragma once
#include "common.cuh"

namespace bk {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a block of SCAN_THREADS threads;
// returns the exclusive prefix, *total gets the block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[SCAN_THREADS / 32];
  __shared__ uint32_t btotal;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const uint32_t inc = warp_incl_scan(v);
  if (l == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (l < SCAN_THREADS / 32) ? wsum[l] : 0;
    const uint32_t si = warp_incl_scan(s);
    if (l < SCAN_THREADS / 32) wsum[l] = si - s;
    if (l == SCAN_THREADS / 32 - 1) btotal = si;
  }
  __syncthreads();
  const uint32_t r = inc - v + wsum[w];
  *total = btotal;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                       uint32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place; writes the grand total
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(uint32_t* __restrict__ tile_sums, int64_t n_tiles,
                                                                   uint32_t* __restrict__ grand_total) {
  uint32_t carry = 0;
  for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = (i < n_tiles) ? tile_sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                   const uint32_t* __restrict__ tile_offsets,
                                                                   uint32_t* __restrict__ out) {
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}

// host helper.  `tmp` needs ceil(n / SCAN_TILE) uint32.  out may alias in.
inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* tmp, uint32_t* grand_total,
                               cudaStream_t st) {
  if (n <= 0) {
    if (grand_total) cudaMemsetAsync(grand_total, 0, sizeof(uint32_t), st);
    return;
  }
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, tmp);
  scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(tmp, tiles, grand_total);
  scan_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, tmp, out);
}
inline int64_t scan_tmp_elems(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

}  // namespace bk
