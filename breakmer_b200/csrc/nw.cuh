// Overlap aligner: one warp computes BOTH directions of the reference's
// `olc.nw` for a (contig, read) pair in a single sweep of the DP table.
//
// Replaces olc.py:40-107 (`nw`) as called from sv_assembly.py:451-452
// (`v1 = nw(contig, read)`, `v2 = nw(read, contig)`).
//
// Observations the kernel is built on (each is checked by tests against the
// oracle, which is pinned to the reference's own output):
//   1. The score table of nw(b, a) is the transpose of the score table of
//      nw(a, b): the recurrence max(diag, up-2, left-2) is symmetric.  Only the
//      pointer priority (diag > up > left, olc.py:69-74) and the end-cell scan
//      (last column, largest row on ties, olc.py:79-83) are direction specific.
//   2. The caller never needs the alignment strings, only where the traceback
//      ends.  That origin can be carried forward with the score: a cell
//      inherits the origin of the predecessor its pointer selects, and the
//      traceback stops at the first cell with i == 0 or j == 0 (olc.py:105).
//      An origin is therefore a boundary cell, encoded as (j - i).
//   3. score, pointer tag and origin pack into one 32-bit word
//          [ score : 14 | tag : 2 | origin+32768 : 16 ]
//      so that a signed max over the three candidates selects the score, breaks
//      ties by pointer priority and carries the origin -- 4 integer ops per
//      direction per cell with the DPX add-max instruction (VIADDMNMX).
//   Two words are kept per cell, one per direction (they differ in the tag
//   order of the two gap moves).
//
// Layout: the "column" sequence (length m) is spread over the 32 lanes, C
// consecutive columns per lane in registers; rows are swept systolically (lane
// L works on row t-L at step t, its left neighbour's column arrives by
// shuffle).  Column sequences longer than 32*C are processed in column blocks,
// the block boundary column travelling through a scratch buffer.
//
// In the terms of olc.nw, for dev-frame A = nw(colseq, rowseq): colseq is seq1
// (index j), rowseq is seq2 (index i); B = nw(rowseq, colseq).
#pragma once
#include "common.cuh"

namespace bk {

struct NwOut {   // fields [2:7] of the tuple olc.nw returns (olc.py:107)
  int prej, j0, prei, i0, score;
};
struct NwDual {
  NwOut a;   // nw(colseq, rowseq)
  NwOut b;   // nw(rowseq, colseq)
};

constexpr int NW_TAB_BYTES = 160 * 1024;   // per-warp score-table scratch of the score pass + traceback (below)
constexpr int NW_SHIFT = 18;
constexpr int NW_BIAS = 32768;
constexpr int NW_TAG_CLEAR = ~(3 << 16);
constexpr int NW_ONE = 1 << NW_SHIFT;
// candidate increments: score delta in the top field, pointer tag in bits 16-17
constexpr int NW_D_MATCH = 1 * NW_ONE + (3 << 16);
constexpr int NW_D_MISM = -2 * NW_ONE + (3 << 16);
constexpr int NW_GAP_HI = -2 * NW_ONE + (2 << 16);   // the gap move with priority 2
constexpr int NW_GAP_LO = -2 * NW_ONE + (1 << 16);   // the gap move with priority 1

BK_HD void nw_decode_a(int best_q, int best_i, int m, NwOut& o) {
  o.prej = m;
  o.prei = best_i;
  o.score = best_q >> NW_SHIFT;
  if (best_i == 0) {           // pointer[0][m] == 2: one step left (olc.py:58-59,96-99)
    o.j0 = m - 1;
    o.i0 = 0;
  } else {
    int e = (best_q & 0xffff) - NW_BIAS;
    o.j0 = e >= 0 ? e : 0;
    o.i0 = e >= 0 ? 0 : -e;
  }
}
// B = nw(rowseq, colseq): its seq1 is rowseq (length n), its rows run over colseq
BK_HD void nw_decode_b(int best_q, int best_j, int n, NwOut& o) {
  o.prej = n;
  o.prei = best_j;
  o.score = best_q >> NW_SHIFT;
  if (best_j == 0) {
    o.j0 = n - 1;
    o.i0 = 0;
  } else {
    int e = (best_q & 0xffff) - NW_BIAS;   // dev-frame origin (i0d, j0d)
    int j0d = e >= 0 ? e : 0, i0d = e >= 0 ? 0 : -e;
    o.j0 = i0d;   // start in rowseq
    o.i0 = j0d;   // start in colseq
  }
}

#ifdef BK_SIM
// ---------------------------------------------------------------------------
// tests/sim only: scalar stand-in with the same contract (host debugging of the
// assembler control logic).  Not compiled into the product library.
// ---------------------------------------------------------------------------
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
template <int C, bool PTR>
inline void nw_dual_warp(const uint8_t* cs, int m, const uint8_t* rs, int n, int2*, int2*, uint8_t* ptrmat, NwDual& out) {
  (void)ptrmat;
  static thread_local int *qa = nullptr, *qb = nullptr;
  static thread_local size_t cap = 0;
  size_t need = (size_t)(m + 1) * 2;
  if (need > cap) { delete[] qa; delete[] qb; qa = new int[need]; qb = new int[need]; cap = need; }
  int* pa = qa; int* ca = qa + (m + 1);
  int* pb = qb; int* cb = qb + (m + 1);
  for (int j = 0; j <= m; ++j) pa[j] = pb[j] = NW_BIAS + j;
  int best_a = pa[m], best_ai = 0;
  int best_b = NW_BIAS, best_bj = 0;
  if (n == 0) { /* unreachable: callers reject empty sequences */ }
  for (int i = 1; i <= n; ++i) {
    ca[0] = cb[0] = NW_BIAS - i;
    for (int j = 1; j <= m; ++j) {
      int s = (cs[j - 1] == rs[i - 1]) ? NW_D_MATCH : NW_D_MISM;
      int a = pa[j - 1] + s, h = ca[j - 1] + NW_GAP_HI, v = pa[j] + NW_GAP_LO;
      int r = a > h ? a : h; r = r > v ? r : v;
      ca[j] = r & NW_TAG_CLEAR;
      a = pb[j - 1] + s; h = cb[j - 1] + NW_GAP_LO; v = pb[j] + NW_GAP_HI;
      r = a > h ? a : h; r = r > v ? r : v;
      cb[j] = r & NW_TAG_CLEAR;
    }
    if ((ca[m] >> NW_SHIFT) >= (best_a >> NW_SHIFT)) { best_a = ca[m]; best_ai = i; }
    int* t = pa; pa = ca; ca = t; t = pb; pb = cb; cb = t;
  }
  for (int j = 0; j <= m; ++j)
    if ((pb[j] >> NW_SHIFT) >= (best_b >> NW_SHIFT)) { best_b = pb[j]; best_bj = j; }
  nw_decode_a(best_a, best_ai, m, out.a);
  nw_decode_b(best_b, best_bj, n, out.b);
}
template <int C>
inline void nw_dual_warp_fast(const uint8_t* cs, int m, const uint8_t* rs, int n, int2* e0, int2* e1, NwDual& out) {
  nw_dual_warp<C, false>(cs, m, rs, n, e0, e1, nullptr, out);
}
template <bool LAZY = false>
inline void nw_dual_dispatch(const uint8_t* cs, int m, const uint8_t* rs, int n, int2* e0, int2* e1, uint2* /*lastcol*/,
                             uint8_t* /*tab*/, NwDual& out) {
  nw_dual_warp<4, false>(cs, m, rs, n, e0, e1, nullptr, out);
}
#else
// ---------------------------------------------------------------------------
// The product kernel.
// ---------------------------------------------------------------------------
template <int C, bool PTR>
__device__ __forceinline__ void nw_dual_warp(const uint8_t* __restrict__ cs, int m,
                                             const uint8_t* __restrict__ rs, int n,
                                             int2* edge0, int2* edge1, uint8_t* ptrmat, NwDual& out) {
  const unsigned FULL = 0xffffffffu;
  const int L = lane();
  constexpr int W = 32 * C;
  const int nblk = (m + W - 1) / W;
  int best_a = NW_BIAS + m, best_ai = 0;        // score[0][m] = 0 (olc.py:79-83 starts at row 0)
  int best_b = NW_BIAS, best_bj = 0;            // score[n][0] = 0
  for (int b = 0; b < nblk; ++b) {
    const int jb = b * W;
    const int jfirst = jb + L * C + 1;          // 1-based column held in slot 0
    int colA[C], colB[C], ch[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = jfirst + c;
      colA[c] = colB[c] = NW_BIAS + j;          // row 0: score 0, origin (0, j)
      ch[c] = (j <= m) ? (int)cs[j - 1] : 0x100;
    }
    int diagA = NW_BIAS + (jfirst - 1), diagB = diagA;
    int lastA = colA[C - 1], lastB = colB[C - 1];
    const int2* ein = (b & 1) ? edge1 : edge0;
    int2* eout = (b & 1) ? edge0 : edge1;
    const bool more = (b + 1 < nblk);
    const bool has_last = (m > jb) && (m <= jb + W);
    const int own_lane = (m - 1 - jb) / C, own_c = (m - 1 - jb) % C;
    const int steps = n + 31;
    for (int t = 0; t < steps; ++t) {
      int hA = __shfl_up_sync(FULL, lastA, 1);
      int hB = __shfl_up_sync(FULL, lastB, 1);
      const int i = t - L + 1;
      const bool active = (i >= 1) && (i <= n);
      if (L == 0) {
        if (b == 0) {
          hA = hB = NW_BIAS - i;                // column 0: score 0, origin (i, 0)
        } else if (active) {
          const int2 e = ein[i];
          hA = e.x; hB = e.y;
        }
      }
      if (active) {
        const int rc = (int)rs[i - 1];
        int dA = diagA, dB = diagB;
        diagA = hA; diagB = hB;
        unsigned ptrs = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int vA = colA[c], vB = colB[c];
          const int s = (ch[c] == rc) ? NW_D_MATCH : NW_D_MISM;
          // direction A (olc.py:64-74): diag(3) > score[i][j-1] (2) > score[i-1][j] (1)
          int tA = dA + s;
          tA = __viaddmax_s32(hA, NW_GAP_HI, tA);
          tA = __viaddmax_s32(vA, NW_GAP_LO, tA);
          // direction B is the transpose, so the two gap moves swap priority
          int tB = dB + s;
          tB = __viaddmax_s32(vB, NW_GAP_HI, tB);
          tB = __viaddmax_s32(hB, NW_GAP_LO, tB);
          if (PTR) ptrs |= ((unsigned)(tA >> 16) & 3u) << (2 * c);
          dA = vA; dB = vB;
          hA = tA & NW_TAG_CLEAR; hB = tB & NW_TAG_CLEAR;
          colA[c] = hA; colB[c] = hB;
        }
        lastA = colA[C - 1]; lastB = colB[C - 1];
        if (PTR) {
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (jfirst + c <= m) ptrmat[(size_t)i * (m + 1) + jfirst + c] = (uint8_t)((ptrs >> (2 * c)) & 3u);
        }
        if (more && L == 31) eout[i] = make_int2(lastA, lastB);
        if (has_last && L == own_lane) {
          int cand = colA[0];
#pragma unroll
          for (int c = 1; c < C; ++c) cand = (c == own_c) ? colA[c] : cand;
          if ((cand >> NW_SHIFT) >= (best_a >> NW_SHIFT)) { best_a = cand; best_ai = i; }
        }
        if (i == n) {
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (jfirst + c <= m && (colB[c] >> NW_SHIFT) >= (best_b >> NW_SHIFT)) { best_b = colB[c]; best_bj = jfirst + c; }
        }
      }
    }
    __syncwarp();
  }
  // A lives on the lane that owns column m of the last block
  {
    const int jb = (nblk - 1) * W;
    const int own_lane = (m - 1 - jb) / C;
    best_a = __shfl_sync(FULL, best_a, own_lane);
    best_ai = __shfl_sync(FULL, best_ai, own_lane);
  }
  // B: max score over the last row, largest column on ties (olc.py:81 uses >=)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const int oq = __shfl_xor_sync(FULL, best_b, off);
    const int oj = __shfl_xor_sync(FULL, best_bj, off);
    const int s0 = best_b >> NW_SHIFT, s1 = oq >> NW_SHIFT;
    if (s1 > s0 || (s1 == s0 && oj > best_bj)) { best_b = oq; best_bj = oj; }
  }
  nw_decode_a(best_a, best_ai, m, out.a);
  nw_decode_b(best_b, best_bj, n, out.b);
}

// Fast path (no pointer table): TWO rows per systolic step.  Lane L works on rows
// (2(t-L)+1, 2(t-L)+2) at step t; within the lane's C-column strip the 2 x C tile is
// evaluated as straight-line code, so the four independent dependency chains (two
// rows x two directions) overlap in the pipeline, and the step count is halved.
// Row characters are fetched one step ahead.  `rs` must be 2-byte aligned and
// readable up to index n (one byte past the sequence).
// OWN_C >= 0: column m is known at compile time to sit in slot OWN_C of its lane (single
// column block), which turns the last-column tracker into two compares; OWN_C = -1 is
// the general case.
// The end cell of direction A (max over the LAST COLUMN, largest row on ties, olc.py:79-83) is not tracked inside the
// sweep: the lane that owns column m stores its two cells of every step to `lastcol` (one 8-byte store, off the ALU
// pipe) and the warp scans those n values once after the sweep -- ~8 ALU instructions per step less in the hot loop.
// lastcol: (n + 1) / 2 + 1 uint2 of per-warp scratch.
template <int C, int OWN_C = -1>
__device__ __forceinline__ void nw_dual_warp_fast(const uint8_t* __restrict__ cs, int m,
                                                  const uint8_t* __restrict__ rs, int n,
                                                  int2* edge0, int2* edge1, uint2* __restrict__ lastcol, NwDual& out) {
  const unsigned FULL = 0xffffffffu;
  constexpr bool SINGLE = OWN_C >= 0;           // one column block, last column in a known slot
  constexpr int LOW = (1 << NW_SHIFT) - 1;      // everything below the score field
  const int L = lane();
  constexpr int W = 32 * C;
  const int nblk = SINGLE ? 1 : (m + W - 1) / W;
  const int hn = (n + 1) >> 1;                  // row pairs
  int best_a = NW_BIAS + m, best_ai = 0;        // score[0][m] = 0 (olc.py:79-83 starts at row 0)
  int best_b = NW_BIAS, best_bj = 0;            // score[n][0] = 0
  for (int b = 0; b < nblk; ++b) {
    const int jb = b * W;
    const int jfirst = jb + L * C + 1;          // 1-based column held in slot 0
    // the sweep ends when the last lane that holds a real column has done the last row pair
    const int cols_here = (m - jb) < W ? (m - jb) : W;
    const int steps = hn + (cols_here + C - 1) / C - 1;
    int colA[C], colB[C], ch[C], r0A[C], r0B[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = jfirst + c;
      colA[c] = colB[c] = NW_BIAS + j;          // row 0: score 0, origin (0, j)
      r0A[c] = r0B[c] = 0;
      ch[c] = (j <= m) ? (int)cs[j - 1] : 0x100;
    }
    int diagA = NW_BIAS + (jfirst - 1), diagB = diagA;
    int lastA0 = 0, lastB0 = 0, lastA1 = colA[C - 1], lastB1 = colB[C - 1];
    const int2* ein = (b & 1) ? edge1 : edge0;
    int2* eout = (b & 1) ? edge0 : edge1;
    const bool more = !SINGLE && (b + 1 < nblk);
    const bool own = (m > jb) && (m <= jb + W) && (L == (m - 1 - jb) / C);
    const int own_c = (m - 1 - jb) % C;
    const bool lane0_first = (L == 0) && (b == 0);
    // row characters of the first step this lane is active in (t == L): rows 1 and 2
    int rc0_next = rs[0], rc1_next = rs[1];
    for (int t = 0; t < steps; ++t) {
      int hA0 = __shfl_up_sync(FULL, lastA0, 1);
      int hB0 = __shfl_up_sync(FULL, lastB0, 1);
      int hA1 = __shfl_up_sync(FULL, lastA1, 1);
      int hB1 = __shfl_up_sync(FULL, lastB1, 1);
      const int q = t - L;                      // row pair this lane works on
      const bool active = (unsigned)q < (unsigned)hn;
      const int i0 = 2 * q + 1, i1 = i0 + 1;
      if (lane0_first) {
        hA0 = hB0 = NW_BIAS - i0;               // column 0: score 0, origin (i, 0)
        hA1 = hB1 = NW_BIAS - i1;
      }
      if (!SINGLE) {
        if (L == 0 && b > 0 && active) {
          const int2 e0 = ein[i0];
          hA0 = e0.x; hB0 = e0.y;
          if (i1 <= n) { const int2 e1 = ein[i1]; hA1 = e1.x; hB1 = e1.y; }
        }
      }
      if (active) {
        const int rc0 = rc0_next, rc1 = rc1_next;
        {   // prefetch the two row characters of the next step (clamped, always in bounds)
          int nx = i0 + 1;                      // index of row i0+2 in rs
          nx = nx > NW_MAX_LEN - 2 ? NW_MAX_LEN - 2 : nx;
          rc0_next = rs[nx]; rc1_next = rs[nx + 1];
        }
        // ---- row i0 ----
        {
          int dA = diagA, dB = diagB, hA = hA0, hB = hB0;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int vA = colA[c], vB = colB[c];
            const int s = (ch[c] == rc0) ? NW_D_MATCH : NW_D_MISM;
            // the horizontal input (hA / hB) is the only one that depends on the previous column of this row:
            // it goes last, so the dependent chain along a row is one max + the tag clear per cell
            int tA = dA + s;
            tA = __viaddmax_s32(vA, NW_GAP_LO, tA);
            tA = __viaddmax_s32(hA, NW_GAP_HI, tA);
            int tB = dB + s;
            tB = __viaddmax_s32(vB, NW_GAP_HI, tB);
            tB = __viaddmax_s32(hB, NW_GAP_LO, tB);
            dA = vA; dB = vB;
            hA = tA & NW_TAG_CLEAR; hB = tB & NW_TAG_CLEAR;
            r0A[c] = hA; r0B[c] = hB;
          }
        }
        // ---- row i1 (garbage when i1 == n + 1; never observed) ----
        {
          int dA = hA0, dB = hB0, hA = hA1, hB = hB1;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int vA = r0A[c], vB = r0B[c];
            const int s = (ch[c] == rc1) ? NW_D_MATCH : NW_D_MISM;
            // the horizontal input (hA / hB) is the only one that depends on the previous column of this row:
            // it goes last, so the dependent chain along a row is one max + the tag clear per cell
            int tA = dA + s;
            tA = __viaddmax_s32(vA, NW_GAP_LO, tA);
            tA = __viaddmax_s32(hA, NW_GAP_HI, tA);
            int tB = dB + s;
            tB = __viaddmax_s32(vB, NW_GAP_HI, tB);
            tB = __viaddmax_s32(hB, NW_GAP_LO, tB);
            dA = vA; dB = vB;
            hA = tA & NW_TAG_CLEAR; hB = tB & NW_TAG_CLEAR;
            colA[c] = hA; colB[c] = hB;
          }
        }
        diagA = hA1; diagB = hB1;               // Q(i1, jfirst-1): diagonal input of the next step's first row
        lastA0 = r0A[C - 1]; lastB0 = r0B[C - 1];
        lastA1 = colA[C - 1]; lastB1 = colB[C - 1];
        if (!SINGLE) {
          if (more && L == 31) {
            eout[i0] = make_int2(lastA0, lastB0);
            if (i1 <= n) eout[i1] = make_int2(lastA1, lastB1);
          }
        }
        {   // last column (direction A): handed to the post-sweep scan
          int c0, c1;
          if (SINGLE) {
            c0 = r0A[SINGLE ? OWN_C : 0]; c1 = colA[SINGLE ? OWN_C : 0];
          } else {
            c0 = r0A[0]; c1 = colA[0];
#pragma unroll
            for (int c = 1; c < C; ++c) { c0 = (c == own_c) ? r0A[c] : c0; c1 = (c == own_c) ? colA[c] : c1; }
          }
          if (own) lastcol[q] = make_uint2((unsigned)c0, (unsigned)c1);
        }
      }
    }
    // last row (direction B): every lane's registers still hold the last row pair it worked on -- row n is its
    // first row if n is odd, its second if n is even.  Columns in increasing order, >= keeps the largest column.
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int v = (n & 1) ? r0B[c] : colB[c];
      if (jfirst + c <= m && (v | LOW) >= best_b) { best_b = v; best_bj = jfirst + c; }
    }
    __syncwarp();
  }
  {
    // scan of the stored last column: rows in increasing order, >= keeps the largest row; row 0 (score 0) starts it.
    // Lane L takes a contiguous chunk of row pairs, then the lanes combine (higher score, then higher row).
    const int per = (hn + 31) >> 5;
    const int q0 = L * per, q1 = (q0 + per) < hn ? (q0 + per) : hn;
    int ba = (L == 0) ? best_a : (int)0x80000000, bi = (L == 0) ? 0 : -1;
    for (int q = q0; q < q1; ++q) {
      const uint2 v = lastcol[q];
      const int i0 = 2 * q + 1;
      if (((int)v.x | LOW) >= ba) { ba = (int)v.x; bi = i0; }
      if (i0 + 1 <= n && ((int)v.y | LOW) >= ba) { ba = (int)v.y; bi = i0 + 1; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int oq = __shfl_xor_sync(FULL, ba, off);
      const int oi = __shfl_xor_sync(FULL, bi, off);
      const int s0 = ba >> NW_SHIFT, s1 = oq >> NW_SHIFT;
      if (s1 > s0 || (s1 == s0 && oi > bi)) { ba = oq; bi = oi; }
    }
    best_a = ba; best_ai = bi;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const int oq = __shfl_xor_sync(FULL, best_b, off);
    const int oj = __shfl_xor_sync(FULL, best_bj, off);
    const int s0 = best_b >> NW_SHIFT, s1 = oq >> NW_SHIFT;
    if (s1 > s0 || (s1 == s0 && oj > best_bj)) { best_b = oq; best_bj = oj; }
  }
  nw_decode_a(best_a, best_ai, m, out.a);
  nw_decode_b(best_b, best_bj, n, out.b);
}
// ---------------------------------------------------------------------------------------------------------------
// Score pass + warp-parallel traceback (the path almost every alignment of the assembler takes).
//
// The packed cell above spends 8 ALU-pipe instructions per cell because both directions carry their traceback origin
// through EVERY cell.  But the two directions share one score table (observation 1), and the origin is only needed
// for two cells: the two end cells.  So:
//
//   pass 1  computes the score table ONCE, tag-free, in the shifted form  T[i][j] = score[i][j] + 2 (i + j):
//             diagonal move   T[i-1][j-1] + (match ? 5 : 2)        (+1 / -2, plus the shift of 4)
//             either gap move T[i][j-1], T[i-1][j] unchanged       (-2, plus the shift of 2)
//           i.e. one add and one three-way max (VIMNMX3) per cell; T is non-negative, non-decreasing along rows and
//           columns and differs by at most 5 between neighbours.  The low byte of every T goes to a per-warp scratch
//           table (L2 resident), laid out by (step, lane) so that each lane's two rows x C columns of a step are one
//           8- or 16-byte store and a warp's stores are contiguous.
//   ends    direction A: the last column (kept in full by the lane that owns column m, as before), largest row on
//           ties; direction B: the last row, largest column on ties (olc.py:79-83 with its `>=`).
//   pass 2  walks the two tracebacks (olc.py:90-105) from the end cells.  The pointer of a cell is recomputed from
//           the table: 3 iff T == T_diag + c, else -- direction A -- 2 iff T == T[i][j-1], else 1; direction B is the
//           transpose, so its second choice is T[i-1][j].  Neighbour differences are < 256, so the low bytes decide
//           these equalities exactly.  Alignments are long diagonal runs: the 32 lanes test the next 32 cells of the
//           diagonal at once and the walk jumps to the first cell that is not a diagonal step, so a traceback costs
//           a handful of warp iterations instead of one step per base.
//
// Results are identical to nw_dual_warp_fast (tests/test_gpu_nw.py runs both paths against the reference's outputs).
// Limits: one column block (m <= 32 C) and (rows / 2 + 32) * 64 C bytes of table <= NW_TAB_BYTES; anything else takes
// the packed-cell kernel above.

template <int C>
__device__ __forceinline__ bool nw_trace_fits(int m, int n) {
  return m <= 32 * C && (size_t)(((n + 1) >> 1) + 32) * 64 * C <= (size_t)NW_TAB_BYTES;
}

// byte offset of cell (i, j), i >= 1, j >= 1, in the (step, lane) table layout
template <int C>
__device__ __forceinline__ unsigned nw_tab_off(int i, int j) {
  const unsigned Lc = (unsigned)(j - 1) / C, c = (unsigned)(j - 1) % C;
  const unsigned q = (unsigned)(i - 1) >> 1, r = (unsigned)(i - 1) & 1u;
  return (((q + Lc) * 32u + Lc) * 2u + r) * C + c;
}
// low byte of T at (i, j), boundary included (T[0][j] = 2 j, T[i][0] = 2 i)
// (branch-free: the load is always issued, from a clamped in-table address, so that the loads of one traceback
// iteration overlap instead of waiting for each other behind divergent branches)
template <int C>
__device__ __forceinline__ unsigned nw_tab_get(const uint8_t* __restrict__ tab, int i, int j) {
  const bool inside = i >= 1 && j >= 1;
  const unsigned v = (unsigned)__ldcg(tab + (inside ? nw_tab_off<C>(i, j) : 0u));
  const unsigned edge = (unsigned)(2 * (i + j)) & 255u;          // T[0][j] = 2 j, T[i][0] = 2 i
  return inside ? v : edge;
}

// The traceback(s) of pass 2.  Walks direction A from (ia, ja) and / or direction B from (ib, jb) to the boundary; on
// return the coordinates are the traceback origins.  Both walks share the loop (their table loads overlap).
template <int C>
__device__ __forceinline__ void nw_trace_walk(const uint8_t* __restrict__ cs, const uint8_t* __restrict__ rs,
                                              const uint8_t* __restrict__ tab, bool goA, bool goB, int& ia, int& ja, int& ib, int& jb) {
  const unsigned FULL = 0xffffffffu;
  const int L = lane();
  while (goA || goB) {
    bool ndA = true, hA = false, ndB = true, vB = false;
    {
      // cells (i - L, j - L) of both walks; lanes past the boundary (and a walk that is over) read clamped cells
      const int ai = ia - L, aj = ja - L, bi = ib - L, bj = jb - L;
      const bool okA = goA && ai >= 1 && aj >= 1, okB = goB && bi >= 1 && bj >= 1;
      const int ai1 = okA ? ai : 1, aj1 = okA ? aj : 1, bi1 = okB ? bi : 1, bj1 = okB ? bj : 1;
      const unsigned tcA = nw_tab_get<C>(tab, ai1, aj1), tdA = nw_tab_get<C>(tab, ai1 - 1, aj1 - 1), thA = nw_tab_get<C>(tab, ai1, aj1 - 1);
      const unsigned tcB = nw_tab_get<C>(tab, bi1, bj1), tdB = nw_tab_get<C>(tab, bi1 - 1, bj1 - 1), tvB = nw_tab_get<C>(tab, bi1 - 1, bj1);
      const unsigned cA = (cs[aj1 - 1] == rs[ai1 - 1]) ? 5u : 2u, cB = (cs[bj1 - 1] == rs[bi1 - 1]) ? 5u : 2u;
      if (okA) { ndA = ((tcA - tdA - cA) & 255u) != 0u; hA = ((tcA - thA) & 255u) == 0u; }
      if (okB) { ndB = ((tcB - tdB - cB) & 255u) != 0u; vB = ((tcB - tvB) & 255u) == 0u; }
    }
    if (goA) {
      const unsigned mk = __ballot_sync(FULL, ndA);
      const int run = mk ? __ffs((int)mk) - 1 : 32;
      ia -= run; ja -= run;
      if (ia == 0 || ja == 0) goA = false;
      else if (run < 32) {
        if (__shfl_sync(FULL, (int)hA, run)) ja -= 1; else ia -= 1;      // pointer 2: score[i][j-1]; pointer 1: score[i-1][j]
        if (ia == 0 || ja == 0) goA = false;
      }
    }
    if (goB) {
      const unsigned mk = __ballot_sync(FULL, ndB);
      const int run = mk ? __ffs((int)mk) - 1 : 32;
      ib -= run; jb -= run;
      if (ib == 0 || jb == 0) goB = false;
      else if (run < 32) {
        if (__shfl_sync(FULL, (int)vB, run)) ib -= 1; else jb -= 1;      // transposed: its pointer 2 is the vertical neighbour
        if (ib == 0 || jb == 0) goB = false;
      }
    }
  }
}

// LAZY: the caller is contig.check_align (sv_assembly.py:449-504), which reads a direction's traceback origin only
//   - for the higher-scoring direction, unless its score already fails `score < min_len / 4` (:459-460);
//   - for the other direction only if the first one fails the identity test (:461-464) or the scores tie.
// The direction that is not read is typically the one in which the read overhangs the contig: a low-scoring path of
// dozens of single gap steps, i.e. dozens of dependent table look-ups.  Its origin is left at the end cell.
#define BK_TAB_ST(p, v) __stcg((p), (v))
#ifdef BK_NW_PROF
__device__ long long bk_nw_prof[8];
#define BK_NW_T(i) if (lane() == 0) bk_nw_prof[i] = clock64();
#else
#define BK_NW_T(i)
#endif
template <int C, int OWN_C, bool LAZY>
__device__ __forceinline__ void nw_dual_trace(const uint8_t* __restrict__ cs, int m, const uint8_t* __restrict__ rs, int n,
                                              uint2* __restrict__ lastcol, uint8_t* __restrict__ tab, NwDual& out) {
  const unsigned FULL = 0xffffffffu;
  const int L = lane();
  const int hn = (n + 1) >> 1;                  // row pairs
  const int jfirst = L * C + 1;                 // 1-based column held in slot 0
  const int steps = hn + (m + C - 1) / C - 1;   // the sweep ends when the last lane with a real column has done the last pair
  int col[C], r0[C], ch[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = jfirst + c;
    col[c] = 2 * j;                             // row 0: score 0
    r0[c] = 0;
    ch[c] = (j <= m) ? (int)cs[j - 1] : 0x100;
  }
  int diag = 2 * (jfirst - 1);
  int last0 = 0, last1 = col[C - 1];
  const bool own = L == (m - 1) / C;             // (the slot of column m in that lane is OWN_C, resolved at compile time:
  int rc0_next = rs[0], rc1_next = rs[1];        //  selecting it at run time costs 7 % of the sweep)
  BK_NW_T(0)
  // ---- pass 1: scores ----
  for (int t = 0; t < steps; ++t) {
    int h0 = __shfl_up_sync(FULL, last0, 1);
    int h1 = __shfl_up_sync(FULL, last1, 1);
    const int q = t - L;
    const bool active = (unsigned)q < (unsigned)hn;
    if (L == 0) { h0 = 4 * q + 2; h1 = 4 * q + 4; }          // column 0: T[i][0] = 2 i, rows 2q+1 and 2q+2
    if (active) {
      const int rc0 = rc0_next, rc1 = rc1_next;
      {
        int nx = 2 * q + 2;                                    // row i0 + 2 (0-based index in rs), clamped in bounds
        nx = nx > NW_MAX_LEN - 2 ? NW_MAX_LEN - 2 : nx;
        rc0_next = rs[nx]; rc1_next = rs[nx + 1];
      }
      {
        int d = diag, h = h0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int v = col[c];
          const int tt = __vimax3_s32(d + ((ch[c] == rc0) ? 5 : 2), h, v);
          d = v; h = tt; r0[c] = tt;
        }
      }
      {
        int d = h0, h = h1;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int v = r0[c];
          const int tt = __vimax3_s32(d + ((ch[c] == rc1) ? 5 : 2), h, v);
          d = v; h = tt; col[c] = tt;
        }
      }
      diag = h1;
      last0 = r0[C - 1]; last1 = col[C - 1];
      if (own) BK_TAB_ST(lastcol + q, make_uint2((unsigned)r0[OWN_C], (unsigned)col[OWN_C]));
      // low bytes of the 2 x C cells of this step -> table[(t, L)]  (st.global.cg: the table is read back once, from L2)
      if (C == 4) {
        const unsigned w0 = __byte_perm(__byte_perm(r0[0], r0[1], 0x0040), __byte_perm(r0[2], r0[3], 0x0040), 0x5410);
        const unsigned w1 = __byte_perm(__byte_perm(col[0], col[1], 0x0040), __byte_perm(col[2], col[3], 0x0040), 0x5410);
        BK_TAB_ST(reinterpret_cast<uint2*>(tab) + (t * 32 + L), make_uint2(w0, w1));
      } else {
        unsigned w[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          w[g] = __byte_perm(__byte_perm(r0[4 * g], r0[4 * g + 1], 0x0040), __byte_perm(r0[4 * g + 2], r0[4 * g + 3], 0x0040), 0x5410);
          w[2 + g] = __byte_perm(__byte_perm(col[4 * g], col[4 * g + 1], 0x0040), __byte_perm(col[4 * g + 2], col[4 * g + 3], 0x0040), 0x5410);
        }
        BK_TAB_ST(reinterpret_cast<uint4*>(tab) + (t * 32 + L), make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
  }
  BK_NW_T(1)
  // ---- end cells ----
  // direction B: last row n (the lane's first row if n is odd, its second if n is even); key = T - 2 j = score + 2 n
  int best_b = 2 * n, best_bj = 0;              // score[n][0] = 0
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = jfirst + c;
    const int v = ((n & 1) ? r0[c] : col[c]) - 2 * j;
    if (j <= m && v >= best_b) { best_b = v; best_bj = j; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const int ov = __shfl_xor_sync(FULL, best_b, off);
    const int oj = __shfl_xor_sync(FULL, best_bj, off);
    if (ov > best_b || (ov == best_b && oj > best_bj)) { best_b = ov; best_bj = oj; }
  }
  __syncwarp();                                  // lastcol / table stores of all lanes are visible to the warp
  // direction A: last column m; key = T - 2 i = score + 2 m; rows ascending, >= keeps the largest row
  int best_a, best_ai;
  {
    const int per = (hn + 31) >> 5;
    const int q0 = L * per, q1 = (q0 + per) < hn ? (q0 + per) : hn;
    int ba = (L == 0) ? 2 * m : (int)0x80000000, bi = (L == 0) ? 0 : -1;
    for (int q = q0; q < q1; ++q) {
      const uint2 v = lastcol[q];
      const int i0 = 2 * q + 1;
      const int k0 = (int)v.x - 2 * i0, k1 = (int)v.y - 2 * (i0 + 1);
      if (k0 >= ba) { ba = k0; bi = i0; }
      if (i0 + 1 <= n && k1 >= ba) { ba = k1; bi = i0 + 1; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int ov = __shfl_xor_sync(FULL, ba, off);
      const int oi = __shfl_xor_sync(FULL, bi, off);
      if (ov > ba || (ov == ba && oi > bi)) { ba = ov; bi = oi; }
    }
    best_a = ba; best_ai = bi;
  }
  out.a.prej = m; out.a.prei = best_ai; out.a.score = best_a - 2 * m;
  out.b.prej = n; out.b.prei = best_bj; out.b.score = best_b - 2 * n;
  BK_NW_T(2)
  // ---- pass 2: traceback(s) ----
  int ia = best_ai, ja = m, ib = n, jb = best_bj;
  bool goA = best_ai > 0, goB = best_bj > 0;
  if (!goA) { ja = m - 1; }                      // pointer[0][m] == 2: one step left (olc.py:58-59, 96-99)
  if (!goB) { ib = n - 1; }                      // same in the transposed problem
  if (LAZY) {
    const int sA = out.a.score, sB = out.b.score, mn = m < n ? m : n;
    const bool lowA = 4 * sA < mn, lowB = 4 * sB < mn;          // `score < min_len / 4`: fails whatever the span is
    if (sA != sB) {
      const bool firstA = sA > sB;
      if (firstA ? lowA : lowB) {
        goA = false; goB = false;                                // both directions fail on their scores alone
      } else {
        const bool g1A = goA && firstA, g1B = goB && !firstA;
        nw_trace_walk<C>(cs, rs, tab, g1A, g1B, ia, ja, ib, jb);
        // identity test of the first direction (:461-464, integer form Q27); the other one is read only if it fails
        const int span = firstA ? (m - ja) : (n - ib);
        const int s1 = firstA ? sA : sB;
        const bool bad1 = 200 * s1 < 179 * span;
        if (firstA) goA = false; else goB = false;
        if (!bad1 || (firstA ? lowB : lowA)) { goA = false; goB = false; }
      }
    } else if (lowA) {
      goA = false; goB = false;
    }
  }
  nw_trace_walk<C>(cs, rs, tab, goA, goB, ia, ja, ib, jb);
  out.a.j0 = ja; out.a.i0 = ia;                  // start in colseq (seq1 of A), start in rowseq
  out.b.j0 = ib; out.b.i0 = jb;                  // B: seq1 is rowseq
  __syncwarp();
  BK_NW_T(3)
}

// dispatch on the read length: 4 columns per lane up to 128 bases (with the last-column
// slot resolved at compile time), 8 beyond; the score pass + traceback when its table fits
template <bool LAZY = false>
__device__ __forceinline__ void nw_dual_dispatch(const uint8_t* __restrict__ cs, int m, const uint8_t* __restrict__ rs, int n,
                                                 int2* e0, int2* e1, uint2* lastcol, uint8_t* tab, NwDual& out) {
  if (m <= 128) {
    if (tab && nw_trace_fits<4>(m, n)) {
      switch ((m - 1) & 3) {
        case 0: nw_dual_trace<4, 0, LAZY>(cs, m, rs, n, lastcol, tab, out); break;
        case 1: nw_dual_trace<4, 1, LAZY>(cs, m, rs, n, lastcol, tab, out); break;
        case 2: nw_dual_trace<4, 2, LAZY>(cs, m, rs, n, lastcol, tab, out); break;
        default: nw_dual_trace<4, 3, LAZY>(cs, m, rs, n, lastcol, tab, out); break;
      }
      return;
    }
    nw_dual_warp_fast<4>(cs, m, rs, n, e0, e1, lastcol, out);      // (rare now: contigs beyond the score table's reach)
  } else {
    nw_dual_warp_fast<8>(cs, m, rs, n, e0, e1, lastcol, out);
  }
}
#endif  // BK_SIM

}  // namespace bk
