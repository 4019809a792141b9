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
  bk_destroy(h);
  bk_ingest_destroy(g);
  return 0;
}
