// tests/sim: the batched k-mer stage -- region_kmer_kernel (512 threads per region: shared- or global-memory hash table,
// tiled window enumeration, block-wide bitonic sort) and region_compact_kernel of breakmer_b200/csrc/region_kmers.cuh,
// the product source -- on the host SIMT emulator of simt_host.h.  TEST TOOL ONLY.  Built and driven by
// tests/test_simt_kstage.py.  The host side below restates what pipeline_impl.cuh does around the two launches
// (record tables without empty records, table placement, the prefix sum of the per-region counts).
#define BK_SIMT 1
#include <algorithm>
#include <vector>

#ifdef BK_SIMT
#include "region_kmers.cuh"                  // the copy under tests/sim/_gen (gen_simt_sources.py)
#else
#include "../../breakmer_b200/csrc/region_kmers.cuh"
#endif

using namespace bk;

namespace {
struct HostSet {
  std::vector<int64_t> koff, reg_base, reg_krec;
  RkSet dev{nullptr, nullptr, nullptr, nullptr};
};
// pipeline_impl.cuh: upload_record_set
void make_set(const uint8_t* bases, const int64_t* off, const int64_t* reg_off, int R, HostSet& s) {
  if (!off || !reg_off) return;
  s.reg_base.assign(R + 1, 0);
  s.reg_krec.assign(R + 1, 0);
  for (int r = 0; r < R; ++r) {
    s.reg_base[r] = off[reg_off[r]];
    s.reg_krec[r] = (int64_t)s.koff.size();
    for (int64_t i = reg_off[r]; i < reg_off[r + 1]; ++i)
      if (off[i + 1] > off[i]) s.koff.push_back(off[i]);
  }
  const int64_t n_bases = off[reg_off[R]];
  s.reg_base[R] = n_bases;
  s.reg_krec[R] = (int64_t)s.koff.size();
  s.koff.push_back(n_bases);
  s.dev = RkSet{bases, s.koff.data(), s.reg_base.data(), s.reg_krec.data()};
}
}  // namespace

// sets: 0 soft-clip, 1 reads, 2 reference, 3 normal (bases[i] == NULL: absent).  Outputs: so_off[R + 1], and so_mer / so_cnt
// with room for one entry per soft-clip base.  force_global: every region's table is a slice of the global table.
extern "C" int simt_region_kmers(int R, int k, const uint8_t* const* bases, const int64_t* const* off, const int64_t* const* reg_off,
                                 int force_global, int grid, int64_t* so_off, uint64_t* so_mer, uint32_t* so_cnt) {
  HostSet hs[4];
  for (int i = 0; i < 4; ++i)
    if (bases[i] || off[i]) make_set(bases[i], off[i], reg_off[i], R, hs[i]);
  RegionKmerParams P;
  memset(&P, 0, sizeof P);
  P.n_regions = R; P.k = k;
  P.sc = hs[0].dev; P.reads = hs[1].dev; P.ref = hs[2].dev; P.normal = hs[3].dev;
  if (!P.sc.bases || !P.reads.bases) return -1;
  // pipeline_impl.cuh: pipeline_upload_into (table placement)
  std::vector<uint32_t> cap(R, 1024);
  std::vector<int64_t> goff(R, -1);
  int smem_cap = 1024;
  int64_t gslots = 0;
  for (int r = 0; r < R; ++r) {
    const int64_t n_sc = hs[0].reg_base[r + 1] - hs[0].reg_base[r];
    uint64_t c = 1024;
    while (c <= (uint64_t)n_sc) c <<= 1;
    cap[r] = (uint32_t)c;
    if (!force_global && c <= (uint64_t)RK_SMEM_CAP_MAX) smem_cap = std::max<int>(smem_cap, (int)c);
    else { goff[r] = gslots; gslots += (int64_t)c; }
  }
  std::vector<uint64_t> gkeys(gslots + 1, 0x1234);
  std::vector<uint32_t> gcnt(gslots + 1, 77);
  const int64_t n_sc_bases = hs[0].reg_base[R];
  std::vector<uint64_t> st_mer(n_sc_bases + 1);
  std::vector<uint32_t> st_cnt(n_sc_bases + 1);
  std::vector<uint32_t> seg(R + 1, 0xdeadbeef);
  P.smem_cap = smem_cap; P.tab_cap = cap.data(); P.gtab_off = goff.data();
  P.gkeys = gkeys.data(); P.gcnt = gcnt.data();
  P.st_mer = st_mer.data(); P.st_cnt = st_cnt.data(); P.seg_counts = seg.data();
  if ((size_t)smem_cap * 12 + RK_TILE + 64 > 256 * 1024) return -2;
  memset(simt::dyn_smem(), 0xEE, 256 * 1024);            // shared memory is not zero on entry
  if (grid < 1) grid = 1;
  if (grid > R) grid = R > 0 ? R : 1;
  for (int b = 0; b < grid; ++b)
    simt::run_block(RK_THREADS / 32, (unsigned)b, [&]() { region_kmer_kernel(P); }, (unsigned)grid);
  so_off[0] = 0;
  for (int r = 0; r < R; ++r) so_off[r + 1] = so_off[r] + seg[r];      // (the device pass: exclusive_scan_u32)
  for (int b = 0; b < grid; ++b)
    simt::run_block(256 / 32, (unsigned)b, [&]() {
      region_compact_kernel(st_mer.data(), st_cnt.data(), hs[0].reg_base.data(), so_off, R, so_mer, so_cnt);
    }, (unsigned)grid);
  return 0;
}
