// Batched `olc.nw` over independent pairs: one warp per pair, both directions
// per sweep (nw.cuh).  Backs the C-ABI entry bk_nw_batch, i.e. the drop-in for
// the call sites sv_assembly.py:451-452 / olc.py:40.  When alignment strings
// are requested (the first two fields of the tuple olc.nw returns) the pointer
// table of direction A is written to scratch and lane 0 walks it back exactly as
// olc.py:90-105 does.
#pragma once
#include "nw.cuh"

namespace bk {

constexpr int NWB_WARPS = 4;
constexpr int NWB_SEQ_CAP = 4096;   // bytes of shared memory per sequence per warp

struct NwBatchParams {
  const uint8_t* seqs;
  const int64_t* seq_off;     // n_seq + 1
  const int32_t* pair_a;      // seq1 of each pair
  const int32_t* pair_b;      // seq2 of each pair
  int64_t n_pairs;
  int32_t* out;               // n_pairs * 10 : A{prej,j0,prei,i0,score}, B{...}
  int2* edge;                 // per warp 2 * edge_stride int2, or null when no seq1 is longer than 256
  int edge_stride;
  uint2* lastcol;             // per warp NW_MAX_LEN / 2 + 1 uint2 (last DP column of the sweep in flight)
  uint8_t* tab;               // per warp NW_TAB_BYTES (score-table scratch of the score pass + traceback), or null
  int want_aln;
  uint8_t* ptr_scratch;       // sum (n+1)(m+1) bytes
  const int64_t* ptr_off;     // n_pairs
  uint8_t* aln1;              // sum (m+n) bytes
  uint8_t* aln2;
  const int64_t* aln_off;     // n_pairs
  int32_t* aln_len;           // n_pairs
};

template <bool PTR>
__device__ __forceinline__ void nw_dispatch(const uint8_t* cs, int m, const uint8_t* rs, int n, int2* e0, int2* e1,
                                            uint2* lastcol, uint8_t* tab, uint8_t* ptrmat, NwDual& out) {
  if (PTR) {
    if (m <= 128) nw_dual_warp<4, true>(cs, m, rs, n, e0, e1, ptrmat, out);
    else nw_dual_warp<8, true>(cs, m, rs, n, e0, e1, ptrmat, out);
  } else {
    nw_dual_dispatch(cs, m, rs, n, e0, e1, lastcol, tab, out);
  }
}

__global__ void __launch_bounds__(NWB_WARPS * 32) nw_batch_kernel(NwBatchParams p) {
  __shared__ __align__(16) uint8_t smem[NWB_WARPS][2][NWB_SEQ_CAP];
  const int w = threadIdx.x >> 5, L = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * NWB_WARPS + w;
  const int64_t nw_total = (int64_t)gridDim.x * NWB_WARPS;
  uint8_t* s1 = smem[w][0];
  uint8_t* s2 = smem[w][1];
  int2* e0 = p.edge ? p.edge + (size_t)gw * 2 * p.edge_stride : nullptr;
  int2* e1 = p.edge ? e0 + p.edge_stride : nullptr;
  uint2* lastcol = p.lastcol + (size_t)gw * (NW_MAX_LEN / 2 + 1);
  uint8_t* tab = p.tab ? p.tab + (size_t)gw * NW_TAB_BYTES : nullptr;
  for (int64_t pi = gw; pi < p.n_pairs; pi += nw_total) {
    const int ia = p.pair_a[pi], ib = p.pair_b[pi];
    const int64_t oa = p.seq_off[ia], ob = p.seq_off[ib];
    const int m = (int)(p.seq_off[ia + 1] - oa), n = (int)(p.seq_off[ib + 1] - ob);
    if (m > NW_MAX_LEN || n > NW_MAX_LEN) continue;                      // nw_long_kernel's (nw_long.cuh); warp-uniform
    __syncwarp();
    for (int x = L; x < m; x += 32) s1[x] = p.seqs[oa + x];
    for (int x = L; x < n; x += 32) s2[x] = p.seqs[ob + x];
    __syncwarp();
    NwDual r;
    if (p.want_aln) {
      uint8_t* pm = p.ptr_scratch + p.ptr_off[pi];
      for (int i = L; i <= n; i += 32) pm[(size_t)i * (m + 1)] = 1;      // olc.py:56-57
      for (int j = L; j <= m; j += 32) pm[j] = 2;                         // olc.py:58-59
      __syncwarp();
      nw_dispatch<true>(s1, m, s2, n, e0, e1, lastcol, nullptr, pm, r);
      __syncwarp();
      if (L == 0) {                                                       // olc.py:86-105
        int i = r.a.prei, j = m, len = 0;
        uint8_t* a1 = p.aln1 + p.aln_off[pi];
        uint8_t* a2 = p.aln2 + p.aln_off[pi];
        for (;;) {
          const int t = pm[(size_t)i * (m + 1) + j];
          if (t == 3)      { a1[len] = s1[j - 1]; a2[len] = s2[i - 1]; --i; --j; }
          else if (t == 2) { a1[len] = s1[j - 1]; a2[len] = '-'; --j; }
          else             { a1[len] = '-'; a2[len] = s2[i - 1]; --i; }
          ++len;
          if (i == 0 || j == 0) break;
        }
        p.aln_len[pi] = len;       // strings are stored reversed; the host shim flips them
      }
    } else {
      nw_dispatch<false>(s1, m, s2, n, e0, e1, lastcol, tab, nullptr, r);
    }
    if (L == 0) {
      int32_t* o = p.out + pi * 10;
      o[0] = r.a.prej; o[1] = r.a.j0; o[2] = r.a.prei; o[3] = r.a.i0; o[4] = r.a.score;
      o[5] = r.b.prej; o[6] = r.b.j0; o[7] = r.b.prei; o[8] = r.b.i0; o[9] = r.b.score;
    }
  }
}

}  // namespace bk
