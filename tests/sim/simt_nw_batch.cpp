// tests/sim: nw_batch_kernel (breakmer_b200/csrc/nw_batch.cuh, the device side of bk_nw_batch: one warp per pair, four
// warps per block, optional pointer table + alignment strings) on the host SIMT emulator.  TEST TOOL ONLY.  The host side
// restates what api.cu does around the launch: scratch sizing and the final flip of the alignment strings.
#define BK_SIMT 1
#include <algorithm>
#include <vector>

#ifdef BK_SIMT
#include "nw_batch.cuh"                  // the copy under tests/sim/_gen (gen_simt_sources.py)
#include "nw_long.cuh"
#else
#include "../../breakmer_b200/csrc/nw_batch.cuh"
#include "../../breakmer_b200/csrc/nw_long.cuh"
#endif

using namespace bk;

// out: n_pairs x 10 ints; aln1 / aln2: caller buffers with aln_off[p] + len(seq1) + len(seq2) bytes per pair
extern "C" int simt_nw_batch(const uint8_t* seqs, const int64_t* seq_off, int n_seq, const int32_t* pair_a, const int32_t* pair_b,
                             int64_t n_pairs, int grid, int use_tab, int32_t* out, int want_aln, uint8_t* aln1, uint8_t* aln2,
                             const int64_t* aln_off, int32_t* aln_len) {
  (void)n_seq;
  NwBatchParams P;
  memset(&P, 0, sizeof P);
  P.seqs = seqs; P.seq_off = seq_off; P.pair_a = pair_a; P.pair_b = pair_b; P.n_pairs = n_pairs; P.out = out;
  if (grid < 1) grid = 1;
  const int64_t warps = (int64_t)grid * NWB_WARPS;
  std::vector<int2> edge((size_t)warps * 2 * (NW_MAX_LEN + 1));
  std::vector<uint2> lastcol((size_t)warps * (NW_MAX_LEN / 2 + 1));
  std::vector<uint8_t> tab(use_tab ? (size_t)warps * NW_TAB_BYTES : 1);
  P.edge = edge.data(); P.edge_stride = NW_MAX_LEN + 1;
  P.lastcol = lastcol.data();
  P.tab = use_tab ? tab.data() : nullptr;
  P.want_aln = want_aln;
  std::vector<int64_t> ptr_off(n_pairs + 1, 0);
  std::vector<uint8_t> ptr;
  if (want_aln) {
    for (int64_t p = 0; p < n_pairs; ++p) {
      const int64_t m = seq_off[pair_a[p] + 1] - seq_off[pair_a[p]], n = seq_off[pair_b[p] + 1] - seq_off[pair_b[p]];
      ptr_off[p + 1] = ptr_off[p] + (n + 1) * (m + 1);
    }
    ptr.assign((size_t)ptr_off[n_pairs] + 1, 0xCC);
    P.ptr_scratch = ptr.data(); P.ptr_off = ptr_off.data();
    P.aln1 = aln1; P.aln2 = aln2; P.aln_off = aln_off; P.aln_len = aln_len;
  }
  for (int b = 0; b < grid; ++b)
    simt::run_block(NWB_WARPS, (unsigned)b, [&]() { nw_batch_kernel(P); }, (unsigned)grid);
  if (want_aln)
    for (int64_t p = 0; p < n_pairs; ++p) {                       // api.cu: the strings are stored reversed
      std::reverse(aln1 + aln_off[p], aln1 + aln_off[p] + aln_len[p]);
      std::reverse(aln2 + aln_off[p], aln2 + aln_off[p] + aln_len[p]);
    }
  return 0;
}

// nw_long_kernel (nw_long.cuh: bk_nw_batch's path for pairs above 4095 bases) on EVERY pair given -- the kernel itself
// has no lower bound, so the reference's golden pairs exercise it.  One block of NWL_THREADS per (pair, direction).
extern "C" int simt_nw_long(const uint8_t* seqs, const int64_t* seq_off, int n_seq, const int32_t* pair_a, const int32_t* pair_b,
                            int64_t n_pairs, int32_t* out, int want_aln, uint8_t* aln1, uint8_t* aln2, const int64_t* aln_off,
                            int32_t* aln_len) {
  NwLongParams Q;
  memset(&Q, 0, sizeof Q);
  Q.seqs = seqs; Q.seq_off = seq_off; Q.pair_a = pair_a; Q.pair_b = pair_b; Q.out = out;
  std::vector<int64_t> idx(n_pairs), ptr_off(n_pairs + 1, 0);
  int64_t longest = 0;
  for (int q = 0; q < n_seq; ++q) longest = std::max(longest, seq_off[q + 1] - seq_off[q]);
  for (int64_t p = 0; p < n_pairs; ++p) {
    idx[p] = p;
    const int64_t m = seq_off[pair_a[p] + 1] - seq_off[pair_a[p]], n = seq_off[pair_b[p] + 1] - seq_off[pair_b[p]];
    ptr_off[p + 1] = ptr_off[p] + (n + 1) * (m + 1);
  }
  Q.long_idx = idx.data();
  Q.diag_stride = longest + 1;                                         // (exact: one element past it is another block's)
  std::vector<int32_t> diag((size_t)(2 * n_pairs * 9 * Q.diag_stride) + 1, (int32_t)0xCCCCCCCC);
  Q.diag = diag.data();
  Q.want_aln = want_aln;
  std::vector<uint8_t> ptr((size_t)ptr_off[n_pairs] + 1, 0xCC);
  if (want_aln) {
    Q.ptr_scratch = ptr.data(); Q.ptr_off = ptr_off.data();
    Q.aln1 = aln1; Q.aln2 = aln2; Q.aln_off = aln_off; Q.aln_len = aln_len;
  }
  const unsigned grid = (unsigned)(2 * n_pairs);
  for (unsigned b = 0; b < grid; ++b)
    simt::run_block(NWL_THREADS / 32, b, [&]() { nw_long_kernel(Q); }, grid);
  if (want_aln)
    for (int64_t p = 0; p < n_pairs; ++p) {
      std::reverse(aln1 + aln_off[p], aln1 + aln_off[p] + aln_len[p]);
      std::reverse(aln2 + aln_off[p], aln2 + aln_off[p] + aln_len[p]);
    }
  return 0;
}
