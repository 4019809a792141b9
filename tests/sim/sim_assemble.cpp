// tests/sim: host-compiled builds of the device assembler (breakmer_b200/csrc/assemble.cuh).  TEST/DEBUG TOOL ONLY:
// they let the kernel source be checked against the oracle in the build container, which has no GPU.  They are not
// part of the product library, are never loaded by breakmer_b200, and are not a fallback -- the product path runs the
// same source compiled by nvcc for sm_100a.
//   default (BK_SIM):  single-lane build of the control logic (a "warp" is one lane, scalar stand-in for the DP)
//   -DBK_SIMT:         the kernel itself -- assemble_kernel<W>, W warps of 32 lanes, the round protocol between the
//                      controller and the aligner warps, the warp DP kernels -- on the fiber emulator of simt_host.h
#ifndef BK_SIMT
#define BK_SIM 1
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifdef BK_SIMT
#include "assemble.cuh"                  // the copy under tests/sim/_gen (gen_simt_sources.py)
#else
#include "../../breakmer_b200/csrc/assemble.cuh"
#endif

using namespace bk;

#ifdef BK_SIMT
static int use_tab = 1;
// 1 (default): aligner warps get a score table (score pass + traceback); 0: the packed-cell kernel, as with BK_NW_PACKED=1
extern "C" void sim_use_score_table(int on) { use_tab = on; }
#endif

extern "C" int sim_assemble_region(
    const uint8_t* rbases, const int64_t* roff, int n_reads, const uint32_t* mult, const uint8_t* io,
    const uint64_t* mers, const uint32_t* counts, int n_mers, int k, int rc_thresh, int read_len,
    // outputs (caller allocated, capacities given)
    int spec_w, int64_t cap, uint8_t* o_seq, int32_t* o_locs, int32_t* o_io, int32_t* o_ot, int32_t* o_reads,
    uint64_t* o_kmer_mer, int32_t* o_kmer_pos, int32_t* o_kmer_meta, int64_t* o_desc, int64_t* n_contigs,
    uint64_t* stats_out) {
  AsmParams P;
  memset(&P, 0, sizeof P);
  P.n_regions = 1; P.k = k; P.rc_thresh = rc_thresh; P.read_cap = ASM_CAP;
  P.rbases = rbases; P.roff = roff;
  int64_t u_off[2] = {0, n_reads};
  std::vector<int32_t> u_rec(n_reads);
  for (int i = 0; i < n_reads; ++i) u_rec[i] = i;
  std::vector<int32_t> u_len(n_reads + 1);
  for (int i = 0; i < n_reads; ++i) u_len[i] = (int32_t)(roff[i + 1] - roff[i]);
  P.u_off = u_off; P.u_rec = u_rec.data(); P.u_mult = mult; P.u_io = io; P.u_len = u_len.data();
  int32_t rl = read_len;
  P.read_len = &rl;
  int64_t so_off[2] = {0, n_mers};
  P.so_off = so_off; P.so_mer = mers; P.so_cnt = counts;
  // seed order and the homopolymer filter are evaluated inside the assembler (next_seed, bind_region)
  std::vector<uint8_t> alive(n_mers + 1), mused(n_mers + 1, 0);
  std::vector<uint64_t> seed_a(n_mers + 1), seed_b(n_mers + 1);
  P.seed_a = seed_a.data(); P.seed_b = seed_b.data();
  std::vector<int32_t> l_mused(n_mers + 1);
  P.l_mused = l_mused.data();
  // posting lists (kernel: index)
  std::vector<std::vector<std::pair<int, int>>> post(n_mers);
  for (int u = 0; u < n_reads; ++u) {
    const uint8_t* s = rbases + roff[u];
    int len = (int)(roff[u + 1] - roff[u]);
    for (int x = 0; x + k <= len; ++x) {
      uint64_t code;
      if (!window_code(s, x, k, code)) continue;
      const uint64_t* it = std::lower_bound(mers, mers + n_mers, code);
      if (it == mers + n_mers || *it != code) continue;
      int si = (int)(it - mers);
      if (!post[si].empty() && post[si].back().first == u) continue;   // first position only
      post[si].push_back({u, x});
    }
  }
  std::vector<int64_t> post_off(n_mers + 1, 0);
  std::vector<int32_t> post_read, post_pos;
  for (int i = 0; i < n_mers; ++i) {
    post_off[i] = (int64_t)post_read.size();
    for (auto& pr : post[i]) { post_read.push_back(pr.first); post_pos.push_back(pr.second); }
  }
  post_off[n_mers] = (int64_t)post_read.size();
  post_read.push_back(0); post_pos.push_back(0);
  P.post_off = post_off.data(); P.post_read = post_read.data(); P.post_pos = post_pos.data();
  // read -> k-mers lists (kernel: index), ascending local mer index
  std::vector<std::vector<std::pair<int, int>>> rk(n_reads);
  for (int si = 0; si < n_mers; ++si)
    for (auto& pr : post[si]) rk[pr.first].push_back({si, pr.second});
  std::vector<int64_t> rk_off(n_reads + 1, 0);
  std::vector<int32_t> rk_s, rk_pos;
  for (int u = 0; u < n_reads; ++u) {
    rk_off[u] = (int64_t)rk_s.size();
    for (auto& e : rk[u]) { rk_s.push_back(e.first); rk_pos.push_back(e.second); }
  }
  rk_off[n_reads] = (int64_t)rk_s.size();
  rk_s.push_back(0); rk_pos.push_back(0);
  P.rk_off = rk_off.data(); P.rk_s = rk_s.data(); P.rk_pos = rk_pos.data();
  std::vector<uint32_t> m_checked(n_mers + 1, 0), m_taken(n_mers + 1, 0), m_first(n_mers + 1, 0);
  P.m_alive = alive.data(); P.m_used = mused.data(); P.m_checked = m_checked.data(); P.m_taken = m_taken.data(); P.m_first = m_first.data();
  const int U = n_reads + 1;
  std::vector<uint8_t> r_used(U, 0), r_deleted(U, 0), r_queued(U, 0);
  std::vector<uint32_t> r_buf(U, 0), r_inreads(U, 0);
  std::vector<int32_t> q_read(U), q_seed(U), l_alt(U), l_del(U), hit_u(U), hit_pos(U), hit2_u(U), hit2_pos(U);
  P.r_used = r_used.data(); P.r_deleted = r_deleted.data(); P.r_queued = r_queued.data();
  P.r_buf = r_buf.data(); P.r_inreads = r_inreads.data();
  P.q_read = q_read.data(); P.q_seed = q_seed.data(); P.l_alt = l_alt.data(); P.l_del = l_del.data();
  P.hit_u = hit_u.data(); P.hit_pos = hit_pos.data(); P.hit2_u = hit2_u.data(); P.hit2_pos = hit2_pos.data();
  std::vector<uint64_t> sort_a(U), sort_b(U);
  P.sort_a = sort_a.data(); P.sort_b = sort_b.data();
  std::vector<uint8_t> w_cseq(ASM_BUF);
  std::vector<int32_t> w_cnt(4 * ASM_BUF), w_K(4 * ASM_KCAP), w_NK(4 * ASM_KCAP), w_diff(ASM_CAP + 1);
  std::vector<uint64_t> w_wcode(ASM_CAP);
  P.w_cseq = w_cseq.data(); P.w_cnt = w_cnt.data(); P.w_K = w_K.data(); P.w_NK = w_NK.data();
  P.w_wcode = w_wcode.data(); P.w_diff = w_diff.data(); P.w_edge = nullptr;
  unsigned long long cursor[5] = {0, 0, 0, 0, 0};
  P.out_cursor = cursor;
  P.cap_seq = P.cap_cnt = P.cap_reads = P.cap_kmers = P.cap_ctg = (unsigned long long)cap;
  P.o_seq = o_seq; P.o_locs = o_locs; P.o_io = o_io; P.o_ot = o_ot; P.o_reads = o_reads;
  P.o_kmer_mer = o_kmer_mer; P.o_kmer_pos = o_kmer_pos; P.o_kmer_meta = o_kmer_meta; P.o_desc = o_desc;
  int32_t status = 0, ncontigs = 0;
  P.region_status = &status; P.region_ncontigs = &ncontigs;
  unsigned long long stats[16] = {0};
  P.stats = stats;
#ifdef BK_SIMT
  // one CTA with blockIdx.x = SLOT (so that the controller role sits on hardware warp SLOT % W, not on warp 0); the
  // per-slot scratch arrays are re-pointed at arrays with SLOT + 1 slots
  const int W = spec_w;
  const int SLOT = 1;
  const size_t NS = SLOT + 1;
  std::vector<uint8_t> xw_cseq(NS * ASM_BUF);
  std::vector<int32_t> xw_cnt(NS * 4 * ASM_BUF), xw_K(NS * 4 * ASM_KCAP), xw_NK(NS * 4 * ASM_KCAP), xw_diff(NS * (ASM_CAP + 1));
  std::vector<uint64_t> xw_wcode(NS * ASM_CAP);
  std::vector<int2> xw_edge(NS * W * 2 * ASM_CAP);
  std::vector<uint2> xw_lastcol(NS * W * ASM_LASTCOL);
  std::vector<uint8_t> xw_tab(NS * W * (size_t)NW_TAB_BYTES);
  P.w_cseq = xw_cseq.data(); P.w_cnt = xw_cnt.data(); P.w_K = xw_K.data(); P.w_NK = xw_NK.data();
  P.w_wcode = xw_wcode.data(); P.w_diff = xw_diff.data(); P.w_edge = xw_edge.data();
  P.w_lastcol = xw_lastcol.data(); P.w_tab = use_tab ? xw_tab.data() : nullptr;
  int max_len = 0;
  for (int i = 0; i < n_reads; ++i) max_len = std::max(max_len, (int)(roff[i + 1] - roff[i]));
  P.read_cap = std::min((int)ASM_CAP, (max_len + 2 + 15) & ~15);
  int work_counter = 0;
  int32_t work_order[1] = {0};
  P.work_counter = &work_counter; P.work_order = work_order;
  unsigned long long region_cells[1] = {0};
  P.region_cells = region_cells;
  if (assemble_smem_bytes(W, P.read_cap) > 256 * 1024) return -100;
  memset(simt::dyn_smem(), 0, 256 * 1024);
  const AsmParams PP = P;
  auto body = [&]() {
    switch (W) {
      case 1: assemble_kernel<1, 0>(PP); break;
      case 2: assemble_kernel<2, 0>(PP); break;
      case 8: assemble_kernel<8, 0>(PP); break;
      default: assemble_kernel<4, 0>(PP); break;
    }
  };
  simt::run_block(W, SLOT, body);
  const int c_status = status;
#else
  std::vector<uint8_t> s_reads((size_t)ASM_SPEC_W * ASM_CAP), s_contig(ASM_CAP), s_pred((size_t)ASM_SPEC_W * ASM_CAP);
  SpecShared sp;
  memset(&sp, 0, sizeof sp);
  RegionCtx c;
  std::vector<int32_t> s_hash(MER_HASH_SIZE);
  bind_region(c, P, 0, 0, s_reads.data(), s_contig.data(), s_pred.data(), s_hash.data(), &sp, spec_w);
  assemble_region(c);
  stats[0] += c.n_align; stats[1] += c.n_cells;
  const int c_status = c.status;
#endif
  *n_contigs = (int64_t)cursor[4];
  for (int i = 0; i < 4; ++i) stats_out[i] = stats[i];
  return c_status;
}
