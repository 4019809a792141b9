// `olc.nw` for pairs with a sequence above NW_MAX_LEN (4095) bases -- olc.py:40-107 has no length limit, the packed-cell
// warp kernels of nw.cuh have (14-bit score / 12-bit origin fields).  Rare path of bk_nw_batch: one CTA per (pair,
// direction) sweeps the table by anti-diagonals in 32-bit scores.  Cell (i, j) (row i = base i of seq2, column j = base
// j of seq1) lies on diagonal d = i + j and needs (i-1, j-1) of diagonal d-2 and (i-1, j), (i, j-1) of diagonal d-1:
// three rolling diagonals indexed by row, O(m + n) memory, one __syncthreads per diagonal.
//
// What the reference's traceback (olc.py:90-105) would return is carried FORWARD with the score: every cell holds the
// (i, j) at which a walk started from it stops (the first position with i == 0 or j == 0), taken over from the
// predecessor the pointer priority diag(3) > up(2) > left(1) of olc.py:62-74 picks.  So the five integers of the tuple
// need no pointer table; only when the alignment STRINGS are asked for (direction A of bk_nw_batch with want_aln) is
// the (n+1)(m+1)-byte pointer table written and walked back by one thread as olc.py does.
#pragma once
#include "common.cuh"

namespace bk {

constexpr int NWL_THREADS = 256;

struct NwLongParams {
  const uint8_t* seqs;
  const int64_t* seq_off;
  const int32_t* pair_a;
  const int32_t* pair_b;
  const int64_t* long_idx;    // n_long pair indices (the pairs nw_batch_kernel skips)
  int32_t* out;               // n_pairs * 10, as NwBatchParams::out
  int32_t* diag;              // per CTA 9 * diag_stride: 3 rolling diagonals x {score, stop row, stop column}
  int64_t diag_stride;        // >= longest sequence of the long pairs + 1
  int want_aln;
  uint8_t* ptr_scratch;       // want_aln: (n+1)(m+1) bytes per long pair (direction A only)
  const int64_t* ptr_off;     // n_long
  uint8_t* aln1;
  uint8_t* aln2;
  const int64_t* aln_off;     // n_pairs (the caller's)
  int32_t* aln_len;           // n_pairs
};

// grid = 2 * n_long: CTA 2q = nw(seq1, seq2) of long pair q (fields 0-4 of its output row), CTA 2q+1 = nw(seq2, seq1)
__global__ void __launch_bounds__(NWL_THREADS) nw_long_kernel(NwLongParams p) {
  __shared__ int red_v[NWL_THREADS];
  __shared__ int red_i[NWL_THREADS];
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x >> 1;
  const int dir = blockIdx.x & 1;
  const int64_t pi = p.long_idx[q];
  const int ia = dir ? p.pair_b[pi] : p.pair_a[pi];
  const int ib = dir ? p.pair_a[pi] : p.pair_b[pi];
  const uint8_t* s1 = p.seqs + p.seq_off[ia];                       // columns (seq1 of this direction)
  const uint8_t* s2 = p.seqs + p.seq_off[ib];                       // rows
  const int m = (int)(p.seq_off[ia + 1] - p.seq_off[ia]);
  const int n = (int)(p.seq_off[ib + 1] - p.seq_off[ib]);
  int32_t* base = p.diag + (size_t)blockIdx.x * 9 * p.diag_stride;
  const size_t stride = (size_t)p.diag_stride;                      // buffer b, field f (score, stop row, stop column) at (3b + f) * stride
  const bool want_ptr = p.want_aln && dir == 0;
  uint8_t* pm = want_ptr ? p.ptr_scratch + p.ptr_off[q] : nullptr;
  const size_t width = (size_t)m + 1;

  // end cell of olc.py:79-83 over the last column (row 0 scores 0); per thread in ascending rows with >=
  int best_v = tid == 0 ? 0 : -2147483647 - 1, best_i = 0, best_oi = 0, best_oj = 0;
  int cur = 0;                                                       // buffer of diagonal d; (cur+2)%3 = d-1, (cur+1)%3 = d-2
  for (int d = 2; d <= m + n; ++d) {
    const int32_t* s_1 = base + (size_t)(3 * ((cur + 2) % 3)) * stride;     // diagonal d-1
    const int32_t* i_1 = s_1 + stride;
    const int32_t* j_1 = i_1 + stride;
    const int32_t* s_2 = base + (size_t)(3 * ((cur + 1) % 3)) * stride;     // diagonal d-2
    const int32_t* i_2 = s_2 + stride;
    const int32_t* j_2 = i_2 + stride;
    int32_t* s_0 = base + (size_t)(3 * cur) * stride;                       // diagonal d
    int32_t* i_0 = s_0 + stride;
    int32_t* j_0 = i_0 + stride;
    const int ilo = d - m > 1 ? d - m : 1;
    const int ihi = d - 1 < n ? d - 1 : n;
    for (int i = ilo + tid; i <= ihi; i += NWL_THREADS) {
      const int j = d - i;
      const bool top = i == 1, lft = j == 1;                         // the neighbours in row 0 / column 0 score 0
      const int dg = ((top || lft) ? 0 : s_2[i - 1]) + (s1[j - 1] == s2[i - 1] ? 1 : -2);   // olc.py:18-20, 63
      const int up = (lft ? 0 : s_1[i]) - 2;                         // score[i][j-1]: consumes seq1 (olc.py:64)
      const int lf = (top ? 0 : s_1[i - 1]) - 2;                     // score[i-1][j]: consumes seq2 (olc.py:65)
      int best = lf > up ? lf : up;
      best = dg > best ? dg : best;
      int si, sj;
      uint8_t t;
      if (best == dg)      { t = 3; if (top || lft) { si = i - 1; sj = j - 1; } else { si = i_2[i - 1]; sj = j_2[i - 1]; } }
      else if (best == up) { t = 2; if (lft)        { si = i;     sj = 0;     } else { si = i_1[i];     sj = j_1[i];     } }
      else                 { t = 1; if (top)        { si = 0;     sj = j;     } else { si = i_1[i - 1]; sj = j_1[i - 1]; } }
      s_0[i] = best;
      i_0[i] = si;
      j_0[i] = sj;
      if (want_ptr) pm[(size_t)i * width + j] = t;
      if (j == m && best >= best_v) { best_v = best; best_i = i; best_oi = si; best_oj = sj; }
    }
    __syncthreads();
    cur = (cur + 1) % 3;
  }
  // the largest row among the maxima wins (>= scan in ascending rows)
  red_v[tid] = best_v;
  red_i[tid] = best_i;
  __syncthreads();
  for (int s = NWL_THREADS / 2; s > 0; s >>= 1) {
    if (tid < s) {
      const int v = red_v[tid + s], i2 = red_i[tid + s];
      if (v > red_v[tid] || (v == red_v[tid] && i2 > red_i[tid])) { red_v[tid] = v; red_i[tid] = i2; }
    }
    __syncthreads();
  }
  const int end_v = red_v[0], end_i = red_i[0];
  if (best_v == end_v && best_i == end_i && (end_i > 0 || tid == 0)) {         // exactly one thread owns the end cell
    int32_t* o = p.out + pi * 10 + dir * 5;
    if (end_i == 0) { best_oi = 0; best_oj = m - 1; }               // pointer[0][m] = 2 (olc.py:58-59): one step along row 0
    o[0] = m; o[1] = best_oj; o[2] = end_i; o[3] = best_oi; o[4] = end_v;
  }
  if (want_ptr && tid == 0) {                                        // olc.py:86-105; stored end-first, the host flips
    int i = end_i, j = m, len = 0;
    uint8_t* a1 = p.aln1 + p.aln_off[pi];
    uint8_t* a2 = p.aln2 + p.aln_off[pi];
    for (;;) {
      const int t = i == 0 ? 2 : pm[(size_t)i * width + j];
      if (t == 3)      { a1[len] = s1[j - 1]; a2[len] = s2[i - 1]; --i; --j; }
      else if (t == 2) { a1[len] = s1[j - 1]; a2[len] = '-'; --j; }
      else             { a1[len] = '-'; a2[len] = s2[i - 1]; --i; }
      ++len;
      if (i == 0 || j == 0) break;
    }
    p.aln_len[pi] = len;
  }
}

}  // namespace bk
