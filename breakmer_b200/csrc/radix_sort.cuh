// Stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass.
//
// This is the "sort-based counting" engine of the k-mer stage (replaces the
// jellyfish hash + `dump -c` + `load_kmers`, utils.py:160-168,287-297): k-mer
// occurrences keyed by (region, mer, set tag) are sorted once, after which
// counting is a run-length pass (kmers.cuh).
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the
// digit-major [256][n_tiles] table, (3) stable scatter (warp match-any ranking,
// tile reordered in shared memory so that stores are coalesced per digit run).
// Only the low `key_bits` bits are sorted: the caller knows how many bits the
// composite key occupies.  HBM traffic per pass: keys read twice, values once,
// both written once = 32 B/element; every access is coalesced except the
// scatter, whose writes are digit-clustered.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "host_util.cuh"

namespace bk {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // keys per tile
constexpr int RS_WCHUNK = 32 * RS_ITEMS;             // contiguous keys per warp

// n_dev (optional): the element count lives on the device (n is then only the upper bound the grid was sized for), so
// that sorts of device-sized arrays need no host round trip.  Tiles past the count write empty histograms / do nothing.
__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                               uint32_t* __restrict__ table, int64_t n_tiles,
                                                               const uint32_t* __restrict__ n_dev) {
  __shared__ uint32_t hist[256];
  if (n_dev) n = (int64_t)*n_dev;
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  table[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

// Stable scatter of one tile.  Ranks come from warp match-any; the tile is first put in
// digit order in shared memory, so the global stores of a warp go to consecutive
// addresses of (at most a few) digit runs instead of 32 scattered sectors.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint64_t* __restrict__ keys_out,
                                                                 uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                 const uint32_t* __restrict__ table, int64_t n_tiles,
                                                                 const uint32_t* __restrict__ n_dev) {
  if (n_dev) n = (int64_t)*n_dev;
  if ((int64_t)blockIdx.x * RS_TILE >= n) return;
  __shared__ uint32_t wcount[RS_WARPS][256];
  __shared__ uint64_t skeys[RS_TILE];
  __shared__ uint32_t svals[RS_TILE];
  __shared__ uint32_t dbase[256];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int d = l; d < 256; d += 32) wcount[w][d] = 0;
  __syncwarp();
  const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
  const int64_t base = tile0 + (int64_t)w * RS_WCHUNK;
  const unsigned lt = (1u << l) - 1u;
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k * 32 + l;
    const bool ok = i < n;
    key[k] = ok ? keys_in[i] : ~0ull;
    val[k] = ok ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const unsigned d = (unsigned)(key[k] >> shift) & 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const uint32_t before = wcount[w][d];
    __syncwarp();
    if ((peers & lt) == 0) wcount[w][d] = before + __popc(peers);
    __syncwarp();
    rank[k] = before + __popc(peers & lt);
  }
  __syncthreads();
  {
    // one thread per digit: tile-local start of the digit, then per-warp starts inside it
    const int d = threadIdx.x;
    uint32_t tot = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) tot += wcount[ww][d];
    uint32_t tile_total;
    const uint32_t tstart = block_excl_scan(tot, &tile_total);
    uint32_t run = tstart;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      const uint32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
    dbase[d] = table[(int64_t)d * n_tiles + blockIdx.x] - tstart;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const unsigned d = (unsigned)(key[k] >> shift) & 255u;
    const uint32_t lpos = wcount[w][d] + rank[k];
    skeys[lpos] = key[k];
    svals[lpos] = val[k];
  }
  __syncthreads();
  const int64_t left = n - tile0;
  const int tile_n = left < RS_TILE ? (int)left : RS_TILE;     // padding sorts to the end of digit 255, i.e. past tile_n
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int i = k * RS_THREADS + threadIdx.x;
    if (i < tile_n) {
      const uint64_t kk = skeys[i];
      const uint32_t dst = dbase[(unsigned)(kk >> shift) & 255u] + (uint32_t)i;
      keys_out[dst] = kk;
      vals_out[dst] = svals[i];
    }
  }
}

struct RadixSortScratch {
  uint64_t* keys_alt;
  uint32_t* vals_alt;
  uint32_t* table;      // 256 * n_tiles
  uint32_t* scan_tmp;   // scan_tmp_elems(256 * n_tiles)
};
inline int64_t rs_num_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts the pairs; the result is left in whichever of the two buffer pairs the
// last pass wrote, returned through *keys_sorted / *vals_sorted.  Returns the
// number of passes run.
inline int radix_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int key_bits, const RadixSortScratch& s,
                            cudaStream_t st, uint64_t** keys_sorted, uint32_t** vals_sorted, KernelTimers& kt,
                            bool aux = false, const uint32_t* n_dev = nullptr) {
  uint64_t* ka = keys; uint64_t* kb = s.keys_alt;
  uint32_t* va = vals; uint32_t* vb = s.vals_alt;
  int passes = 0;
  if (n > 1 || (n_dev && n > 0)) {
    const int64_t tiles = rs_num_tiles(n);
    passes = (key_bits + 7) / 8;
    if (passes > 8) passes = 8;
    for (int p = 0; p < passes; ++p) {
      const int shift = 8 * p;
      {
        TimedLaunch t(kt, st, aux ? KF_AUX_SORT : KF_SORT_COUNT);
        rs_count_kernel<<<(unsigned)tiles, RS_THREADS, 0, st>>>(ka, n, shift, s.table, tiles, n_dev);
      }
      {
        TimedLaunch t(kt, st, aux ? KF_AUX_SORT : KF_SORT_SCAN, 3);
        exclusive_scan_u32(s.table, s.table, 256 * tiles, s.scan_tmp, nullptr, st);
      }
      {
        TimedLaunch t(kt, st, aux ? KF_AUX_SORT : KF_SORT_SCATTER);
        rs_scatter_kernel<<<(unsigned)tiles, RS_THREADS, 0, st>>>(ka, va, kb, vb, n, shift, s.table, tiles, n_dev);
      }
      uint64_t* tk = ka; ka = kb; kb = tk;
      uint32_t* tv = va; va = vb; vb = tv;
    }
  }
  *keys_sorted = ka;
  *vals_sorted = va;
  return passes;
}

}  // namespace bk
