// tests/sim: the WHOLE library (generated emulator sources, see gen_simt_sources.py) in one executable, for runs under
// ThreadSanitizer: -fsanitize=thread -DSIMT_TSAN makes every emulated GPU thread a TSan fiber whose only
// happens-before edges are the barriers the kernels execute (simt_host.h), so TSan reports warp- and block-level data
// races of the kernel source.  TEST TOOL ONLY.
//
//   simt_tsan_driver <manifest> <k> <rc_thresh>
// manifest: one line per region, "ref.fa<TAB>reads.fastq<TAB>sc.fa[<TAB>normal.fastq]".  The regions go through
// bk_ingest_files + bk_compare_kmers_batch; per region the sample-only k-mer count and the contig sequences are printed.
#include <stdio.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "all.cpp"

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s manifest k rc_thresh\n", argv[0]); return 2; }
  std::ifstream mf(argv[1]);
  std::vector<std::string> ref, fq, sc, nm;
  std::string line;
  bool any_normal = false;
  while (std::getline(mf, line)) {
    if (line.empty()) continue;
    std::vector<std::string> f;
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, '\t')) f.push_back(tok);
    if (f.size() < 3) { fprintf(stderr, "bad manifest line\n"); return 2; }
    ref.push_back(f[0]); fq.push_back(f[1]); sc.push_back(f[2]);
    nm.push_back(f.size() > 3 ? f[3] : "");
    any_normal = any_normal || f.size() > 3;
  }
  const int R = (int)ref.size();
  std::vector<const char*> pr, pf, ps, pn;
  for (int i = 0; i < R; ++i) { pr.push_back(ref[i].c_str()); pf.push_back(fq[i].c_str()); ps.push_back(sc[i].c_str()); pn.push_back(nm[i].c_str()); }
  bk_ingest_t g = nullptr;
  if (bk_ingest_create(2, 0, &g) != BK_OK) { fprintf(stderr, "ingest create failed\n"); return 1; }
  bk_batch_input in;
  bk_ingest_text text;
  if (bk_ingest_files(g, R, pr.data(), pf.data(), ps.data(), any_normal ? pn.data() : nullptr, &in, &text) != BK_OK) {
    fprintf(stderr, "ingest: %s\n", bk_ingest_last_error(g));
    return 1;
  }
  in.k = atoi(argv[2]); in.rc_thresh = atoi(argv[3]); in.have_mers = 0;
  bk_handle_t h = nullptr;
  if (bk_create(0, &h) != BK_OK) { fprintf(stderr, "bk_create failed\n"); return 1; }
  bk_batch_result res;
  if (bk_compare_kmers_batch(h, &in, &res) != BK_OK) { fprintf(stderr, "batch: %s\n", bk_last_error(h)); return 1; }
  for (int r = 0; r < R; ++r) {
    printf("region %d status %d sample_only %lld contigs %lld\n", r, res.region_status[r],
           (long long)(res.so_off[r + 1] - res.so_off[r]), (long long)(res.ctg_reg_off[r + 1] - res.ctg_reg_off[r]));
    for (int64_t c = res.ctg_reg_off[r]; c < res.ctg_reg_off[r + 1]; ++c)
      printf("contig %.*s\n", (int)res.ctg_seq_off[2 * c + 1], res.ctg_seq + res.ctg_seq_off[2 * c]);
  }
  // ---- the other entry points, on the records of region 0 (their kernels are not on the batched path) -------------
  {
    const int64_t r0 = in.read_reg_off[0], r1 = in.read_reg_off[1];
    const int64_t n_rec = r1 - r0;
    std::vector<int64_t> off(n_rec + 1);
    for (int64_t i = 0; i <= n_rec; ++i) off[i] = in.read_off[r0 + i] - in.read_off[r0];
    const char* bases = in.read_bases + in.read_off[r0];
    const uint64_t *m_case, *m_sc, *m_ref, *m_only;
    const uint32_t *c_case, *c_sc, *c_ref, *c_only;
    int64_t n_case = 0, n_sc = 0, n_ref = 0, n_only = 0;
    int rc = bk_count_kmers(h, bases, off.data(), n_rec, nullptr, in.k, &m_case, &c_case, &n_case);
    std::vector<uint64_t> case_m(m_case, m_case + n_case);
    std::vector<uint32_t> case_c(c_case, c_case + n_case);
    const int64_t s0 = in.sc_reg_off[0], s1 = in.sc_reg_off[1];
    std::vector<int64_t> soff(s1 - s0 + 1);
    for (int64_t i = 0; i <= s1 - s0; ++i) soff[i] = in.sc_off[s0 + i] - in.sc_off[s0];
    rc |= bk_count_kmers(h, in.sc_bases + in.sc_off[s0], soff.data(), s1 - s0, nullptr, in.k, &m_sc, &c_sc, &n_sc);
    std::vector<uint64_t> sc_m(m_sc, m_sc + n_sc);
    int64_t roff[2] = {0, in.ref_off[1] - in.ref_off[0]};
    rc |= bk_count_kmers(h, in.ref_bases, roff, 1, nullptr, in.k, &m_ref, &c_ref, &n_ref);
    std::vector<uint64_t> ref_m(m_ref, m_ref + n_ref);
    rc |= bk_sample_only(h, in.k, case_m.data(), case_c.data(), n_case, sc_m.data(), n_sc, ref_m.data(), n_ref, nullptr, 0,
                         &m_only, &c_only, &n_only);
    printf("count_kmers rc %d case %lld sc %lld ref %lld sample_only(forward reference only) %lld\n", rc, (long long)n_case,
           (long long)n_sc, (long long)n_ref, (long long)n_only);
    // olc.nw over neighbouring reads, with and without alignment strings
    const int64_t n_pairs = n_rec > 41 ? 40 : (n_rec > 1 ? n_rec - 1 : 0);
    if (n_pairs > 0) {
      std::vector<int32_t> pa(n_pairs), pb(n_pairs), out(n_pairs * 10), alen(n_pairs);
      std::vector<int64_t> aoff(n_pairs + 1, 0);
      for (int64_t p = 0; p < n_pairs; ++p) {
        pa[p] = (int32_t)p; pb[p] = (int32_t)p + 1;
        aoff[p + 1] = aoff[p] + (off[p + 1] - off[p]) + (off[p + 2] - off[p + 1]);
      }
      std::vector<char> a1(aoff[n_pairs] + 1), a2(aoff[n_pairs] + 1);
      int rn = bk_nw_batch(h, bases, off.data(), n_rec, pa.data(), pb.data(), n_pairs, out.data(), 1, a1.data(), a2.data(), aoff.data(), alen.data());
      long long sum = 0;
      for (int64_t p = 0; p < n_pairs; ++p) sum += out[p * 10 + 4] + alen[p];
      rn |= bk_nw_batch(h, bases, off.data(), n_rec, pa.data(), pb.data(), n_pairs, out.data(), 0, nullptr, nullptr, nullptr, nullptr);
      printf("nw_batch rc %d pairs %lld checksum %lld\n", rn, (long long)n_pairs, sum);
      // read redundancy: the first reads as two batches
      const int64_t nd = n_rec > 16 ? 16 : n_rec;
      std::vector<int32_t> mer_pos(nd, 0);
      int64_t boff[3] = {0, nd / 2, nd};
      std::vector<uint8_t> check(nd), flags(nd);
      int64_t npairs = 0; int32_t nl = 0;
      int rd = bk_dedup_reads(h, bases, off.data(), nd, mer_pos.data(), boff, 2, 0.90, check.data(), flags.data(), &npairs, &nl);
      printf("dedup rc %d alignments %lld launches %d\n", rd, (long long)npairs, nl);
    }
    // olc.nw above the packed-cell kernels' 4095 bases (nw_long_kernel): the reference sequence of region 0 repeated to
    // 4,300 bases against its mutated tail; checksum only (parity is tests/test_gpu_nw.py's job)
    {
      const int64_t L0 = in.ref_off[1] - in.ref_off[0];
      std::string a, b;
      while ((int64_t)a.size() < 4300) a.append(in.ref_bases, (size_t)std::min<int64_t>(L0, 4300 - (int64_t)a.size()));
      b = a.substr(4000) + a.substr(100, 150);
      for (size_t x = 7; x < b.size(); x += 41) b[x] = b[x] == 'A' ? 'C' : 'A';
      std::string both = a + b;
      int64_t loff[3] = {0, (int64_t)a.size(), (int64_t)both.size()};
      int32_t lpa[1] = {0}, lpb[1] = {1}, lout[10], lalen[1];
      int64_t laoff[2] = {0, (int64_t)both.size()};
      std::vector<char> l1(both.size() + 1), l2(both.size() + 1);
      int rl = bk_nw_batch(h, both.data(), loff, 2, lpa, lpb, 1, lout, 1, l1.data(), l2.data(), laoff, lalen);
      const int score_with_strings = lout[4];
      rl |= bk_nw_batch(h, both.data(), loff, 2, lpa, lpb, 1, lout, 0, nullptr, nullptr, nullptr, nullptr);
      printf("nw_long rc %d score %d (with strings %d) aln_len %d\n", rl, lout[4], score_with_strings, lalen[0]);
    }
    // the reference k-mer cache instead of the reference sequences
    const long long contigs_before = (long long)res.n_contigs;
    int rcache = bk_ref_cache_build(h, in.ref_bases, in.ref_off, R, in.k);
    bk_batch_input in2 = in;
    in2.ref_bases = nullptr; in2.ref_off = nullptr;
    bk_batch_result res2;
    rcache |= bk_compare_kmers_batch(h, &in2, &res2);
    printf("ref_cache rc %d contigs %lld (with sequences: %lld)\n", rcache, (long long)res2.n_contigs, contigs_before);
    bk_ref_cache_clear(h);
  }
  bk_destroy(h);
  bk_ingest_destroy(g);
  return 0;
}
