// Read-redundancy replay (host side of bk_dedup_reads).
//
// The reference's older assembler variant drops redundant reads from the batch found for one seed k-mer, one read at
// a time (read_batch.check_mer_read, sv_assembly_mm2.py:309-355): each decision needs the two overlap alignments of
// the new read against the batch's most recent read, and which read that is depends on the earlier decisions.
//
// Here the alignments are computed ahead of the decisions, for all batches of a call at once:
//   round 0   every read against its DEDUP_BAND predecessors (right whenever fewer than DEDUP_BAND reads in a row were
//             dropped -- the common case), one kernel launch;
//   round n   the decision chains of all batches are replayed as far as the score table reaches; a batch that needs a
//             pair outside the table asks for its current `last` read against the next DEDUP_WINDOW reads, and all such
//             requests go out in one more launch.
// Every launch serves every unfinished batch, so the number of launches follows the runs of more than DEDUP_BAND dropped
// reads in the worst batch of the call, not the number of reads.  (Measured on a B200, 100 batches of 350 reads: a
// window that doubles per request aligned twice as many pairs for the same 29-30 launches -- the stalls are many short
// runs, not a few long ones -- and took 26 ms instead of 12.6 ms per call.)
//
// A score-table row holds fields [2:7] of nw(seq_i, seq_j) in [0..4] and of nw(seq_j, seq_i) in [5..9], i < j
// (prej, j, prei, i, max_i -- olc.py:107).  Flags written: BK_DEDUP_* of include/breakmer_b200.h.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

constexpr int64_t DEDUP_MAX_PAIRS = int64_t(1) << 23;   // per launch
constexpr int DEDUP_BAND = 4;
constexpr int DEDUP_WINDOW = 8;

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

struct Chain {                                                  // one read_batch
  int64_t lo, hi, cur, last;
  std::unordered_map<int32_t, std::vector<int64_t>> at_pos;     // mer_pos_d, holding read numbers
  std::unordered_map<uint64_t, int64_t> extra;                  // (i, j) outside the band -> table row
};

}  // namespace dedup_detail

// align(pair_a, pair_b, n, out) fills n rows of 10 ints and returns 0 or an error code (the product passes
// bk_nw_batch, the host test harness the oracle's nw).
template <typename Align>
int dedup_run(const int64_t* seq_off, const int32_t* mer_pos, const int64_t* batch_off, int64_t n_batches, double frac,
              uint8_t* check, uint8_t* flags, int64_t* n_pairs_out, int* n_rounds_out, Align&& align) {
  using namespace dedup_detail;
  std::vector<Chain> chains((size_t)n_batches);
  const int64_t n_reads = n_batches ? batch_off[n_batches] : 0;
  std::vector<int64_t> band_row((size_t)n_reads + 1, 0);        // first band row of read j: pairs (j-1, j), (j-2, j), ...
  std::vector<int32_t> pa, pb, tab;
  int64_t n_rows = 0;
  for (int64_t b = 0; b < n_batches; ++b) {
    Chain& c = chains[b];
    c.lo = batch_off[b]; c.hi = batch_off[b + 1]; c.cur = c.lo + 1; c.last = c.lo;
    check[c.lo] = 1;                                            // the opener (read_batch.__init__, :290-294)
    flags[c.lo] = BK_DEDUP_ADDED;
    c.at_pos[mer_pos[c.lo]].push_back(c.lo);
    for (int64_t j = c.lo; j < c.hi; ++j) {
      band_row[j] = n_rows;
      const int64_t w = std::min<int64_t>(DEDUP_BAND, j - c.lo);
      for (int64_t d = 1; d <= w; ++d) { pa.push_back((int32_t)(j - d)); pb.push_back((int32_t)j); }
      n_rows += w;
    }
  }
  auto len = [&](int64_t r) { return seq_off[r + 1] - seq_off[r]; };
  int64_t n_pairs = 0;
  int rounds = 0;
  const bool trace = getenv("BK_DEDUP_TRACE") != nullptr;      // per-round host timing on stderr (tools/dedup_profile.py)
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  for (;;) {
    const auto t0 = now();
    const int64_t asked = (int64_t)pa.size();
    if (!pa.empty()) {
      if ((int64_t)pa.size() > DEDUP_MAX_PAIRS) return BK_ERR_CAPACITY;
      const int64_t have = (int64_t)tab.size() / 10;
      tab.resize((size_t)(have + (int64_t)pa.size()) * 10);
      const int rc = align(pa.data(), pb.data(), (int64_t)pa.size(), tab.data() + have * 10);
      if (rc != 0) return rc;
      n_pairs += (int64_t)pa.size();
      ++rounds;
      pa.clear(); pb.clear();
    }
    const auto t1 = now();
    bool all_done = true;
    for (Chain& c : chains) {
      while (c.cur < c.hi) {
        const int64_t r = c.cur;
        // sim_seqs() is true for every batch read not flagged redundant: its `same_reads(..) or subseq(..)` tests a
        // non-empty tuple (sv_assembly_mm2.py:92), so no score enters this step.
        bool dup = false;
        auto it = c.at_pos.find(mer_pos[r]);
        if (it != c.at_pos.end())
          for (int64_t x : it->second) dup |= !(flags[x] & BK_DEDUP_REDUNDANT);
        if (dup) { check[r] = 0; flags[r] = BK_DEDUP_DELETED; ++c.cur; continue; }
        int64_t row;
        if (r - c.last <= DEDUP_BAND) {
          row = band_row[r] + (r - c.last - 1);
        } else {
          auto e = c.extra.find(((uint64_t)c.last << 32) | (uint64_t)r);
          if (e == c.extra.end()) {                             // ask for `last` against the next reads, resume next round
            for (int64_t q = r; q < std::min(c.hi, r + DEDUP_WINDOW); ++q) {
              c.extra[((uint64_t)c.last << 32) | (uint64_t)q] = n_rows++;
              pa.push_back((int32_t)c.last); pb.push_back((int32_t)q);
            }
            break;
          }
          row = e->second;
        }
        const int32_t* f = tab.data() + row * 10;
        const int64_t last = c.last;
        ++c.cur;
        check[r] = 0;
        const Sub ss1 = subseq(f + 5, len(last), len(r), frac);  // nw(read, last): the new read inside the last one?
        if (ss1.ok && !truthy(ss1)) { flags[r] = BK_DEDUP_DELETED; continue; }
        const Sub ss2 = subseq(f, len(r), len(last), frac);      // nw(last, read): the last read inside the new one?
        if (ss2.ok && !truthy(ss2)) {
          flags[last] |= BK_DEDUP_DELETED | BK_DEDUP_REDUNDANT;
        } else if ((ss1.ok && truthy(ss1)) || (ss2.ok && truthy(ss2))) {
          if (ss1.ok && ss1.score >= ss2.score) { flags[r] = BK_DEDUP_DELETED; continue; }
          if (ss2.ok && ss2.score >= ss1.score) flags[last] |= BK_DEDUP_DELETED | BK_DEDUP_REDUNDANT;
        }
        c.at_pos[mer_pos[r]].push_back(r);
        flags[r] = BK_DEDUP_ADDED;
        check[r] = 1;
        c.last = r;
      }
      all_done &= c.cur >= c.hi;
    }
    if (trace) fprintf(stderr, "dedup round %d: %lld pairs aligned in %.3f ms, replay %.3f ms\n", rounds, (long long)asked,
                       ms(t0, t1), ms(t1, now()));
    if (all_done) break;
  }
  if (n_pairs_out) *n_pairs_out = n_pairs;
  if (n_rounds_out) *n_rounds_out = rounds;
  return 0;
}
