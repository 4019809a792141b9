// Read-redundancy replay (host side of bk_dedup_reads).
//
// The reference's older assembler variant drops redundant reads from the batch found for one seed k-mer, one read at
// a time (read_batch.check_mer_read, sv_assembly_mm2.py:309-355): each decision needs two overlap alignments against
// the batch's most recent read, and which read that is depends on the earlier decisions.  Here the alignments of ALL
// ordered pairs of the batch are computed in one kernel launch first, and this file only replays the decision chain
// over the score table.
//
// Pair (i < j) of a batch lives at row j*(j-1)/2 + i of `tab`; a row holds fields [2:7] of nw(seq_i, seq_j) in
// [0..4] and of nw(seq_j, seq_i) in [5..9] (prej, j, prei, i, max_i -- olc.py:107).
#pragma once
#include <cstdint>
#include <unordered_map>
#include <vector>

// flags written to `flags`: BK_DEDUP_ADDED / _REDUNDANT / _DELETED of include/breakmer_b200.h
constexpr int64_t DEDUP_MAX_PAIRS = int64_t(1) << 23;

namespace dedup_detail {

struct Sub { bool ok; bool has_score; int score; };   // subseq()'s (True, None) / (True, s) / (False, s)

// subseq(seq1, seq2) (sv_assembly_mm2.py:74-85): aln = nw(seq2, seq1); f = that call's five fields.
inline Sub subseq(const int32_t* f, int64_t len1, int64_t len2, double frac) {
  const int score = f[4];
  if (f[0] == len2 && f[1] == 0 && (double)score >= frac * (double)len2)
    return len2 < len1 ? Sub{true, false, 0} : Sub{true, true, score};
  return Sub{false, true, score};
}
inline bool truthy(const Sub& s) { return s.has_score && s.score != 0; }   // Python's truth of the tuple's 2nd field

}  // namespace dedup_detail

// Reads lo..hi-1 form one batch; read lo opens it (read_batch.__init__, :290-294).
inline void dedup_replay(const int64_t* seq_off, const int32_t* mer_pos, int64_t lo, int64_t hi, const int32_t* tab,
                         double frac, uint8_t* check, uint8_t* flags) {
  using namespace dedup_detail;
  std::unordered_map<int32_t, std::vector<int64_t>> at_pos;   // mer_pos_d, holding read numbers
  auto len = [&](int64_t r) { return seq_off[r + 1] - seq_off[r]; };
  check[lo] = 1;
  flags[lo] = BK_DEDUP_ADDED;
  at_pos[mer_pos[lo]].push_back(lo);
  int64_t last = lo;                                           // batch_reads[-1]
  for (int64_t r = lo + 1; r < hi; ++r) {
    flags[r] = 0;
    check[r] = 0;
    // sim_seqs() is true for every batch read not flagged redundant: its `same_reads(..) or subseq(..)` tests a
    // non-empty tuple (sv_assembly_mm2.py:92), so no score enters this step.
    bool dup = false;
    auto it = at_pos.find(mer_pos[r]);
    if (it != at_pos.end())
      for (int64_t x : it->second) dup |= !(flags[x] & BK_DEDUP_REDUNDANT);
    if (dup) { flags[r] = BK_DEDUP_DELETED; continue; }
    const int64_t j = r - lo, i = last - lo;
    const int32_t* row = tab + (j * (j - 1) / 2 + i) * 10;
    const Sub ss1 = subseq(row + 5, len(last), len(r), frac);  // nw(read, last): the new read inside the last one?
    if (ss1.ok && !truthy(ss1)) { flags[r] = BK_DEDUP_DELETED; continue; }
    const Sub ss2 = subseq(row, len(r), len(last), frac);      // nw(last, read): the last read inside the new one?
    if (ss2.ok && !truthy(ss2)) {
      flags[last] |= BK_DEDUP_DELETED | BK_DEDUP_REDUNDANT;
    } else if ((ss1.ok && truthy(ss1)) || (ss2.ok && truthy(ss2))) {
      if (ss1.ok && ss1.score >= ss2.score) { flags[r] = BK_DEDUP_DELETED; continue; }
      if (ss2.ok && ss2.score >= ss1.score) flags[last] |= BK_DEDUP_DELETED | BK_DEDUP_REDUNDANT;
    }
    at_pos[mer_pos[r]].push_back(r);
    flags[r] = BK_DEDUP_ADDED;
    check[r] = 1;
    last = r;
  }
}
