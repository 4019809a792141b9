// tests/sim: the warp DP kernels of breakmer_b200/csrc/nw.cuh (the product source, unmodified) on the 32-lane host
// emulator of simt_host.h.  TEST TOOL ONLY -- see simt_host.h.  Built and driven by tests/test_simt_nw.py.
#define BK_SIMT 1
#include <vector>

#ifdef BK_SIMT
#include "nw.cuh"                  // the copy under tests/sim/_gen (gen_simt_sources.py)
#else
#include "../../breakmer_b200/csrc/nw.cuh"
#endif

using namespace bk;

// mode 0: nw_dual_dispatch<false> (score pass + both tracebacks where its table fits, else the packed-cell kernel)
// mode 1: nw_dual_dispatch<true>  (LAZY: the origin check_align does not read is left at the end cell)
// mode 2: the packed-cell kernel alone (tab = NULL)
// mode 3: nw_dual_warp<4, false>, the one-row-per-step packed kernel with column blocks (what bk_nw_batch uses for strings)
// out: 10 ints = a.{prej, j0, prei, i0, score}, b.{...}; every lane must agree (checked here)
extern "C" int simt_nw_dual(const uint8_t* cs_in, int m, const uint8_t* rs_in, int n, int mode, int* out10) {
  if (m < 1 || n < 1 || m > NW_MAX_LEN || n > NW_MAX_LEN) return -1;
  // the kernels read the row characters one step ahead (clamped to NW_MAX_LEN) and want them 2-byte aligned
  std::vector<uint16_t> csb((NW_MAX_LEN + 8) / 2, 0), rsb((NW_MAX_LEN + 8) / 2, 0);
  uint8_t* cs = (uint8_t*)csb.data();
  uint8_t* rs = (uint8_t*)rsb.data();
  memcpy(cs, cs_in, m);
  memcpy(rs, rs_in, n);
  std::vector<int2> e0(NW_MAX_LEN + 2), e1(NW_MAX_LEN + 2);
  std::vector<uint2> lastcol(NW_MAX_LEN / 2 + 4);
  std::vector<uint8_t> tab(NW_TAB_BYTES + 64, 0xAB);
  NwDual res[32];
  simt::run_warp([&](int l) {
    NwDual o;
    memset(&o, 0, sizeof o);
    switch (mode) {
      case 0: nw_dual_dispatch<false>(cs, m, rs, n, e0.data(), e1.data(), lastcol.data(), tab.data(), o); break;
      case 1: nw_dual_dispatch<true>(cs, m, rs, n, e0.data(), e1.data(), lastcol.data(), tab.data(), o); break;
      case 2: nw_dual_dispatch<false>(cs, m, rs, n, e0.data(), e1.data(), lastcol.data(), nullptr, o); break;
      default: nw_dual_warp<4, false>(cs, m, rs, n, e0.data(), e1.data(), nullptr, o); break;
    }
    res[l] = o;
  });
  for (int l = 1; l < 32; ++l)
    if (memcmp(&res[l], &res[0], sizeof(NwDual)) != 0) return -2 - l;     // lanes disagree
  const NwOut* p[2] = {&res[0].a, &res[0].b};
  for (int d = 0; d < 2; ++d) {
    out10[5 * d + 0] = p[d]->prej; out10[5 * d + 1] = p[d]->j0; out10[5 * d + 2] = p[d]->prei;
    out10[5 * d + 3] = p[d]->i0; out10[5 * d + 4] = p[d]->score;
  }
  return 0;
}

extern "C" int simt_nw_trace_fits(int m, int n) { return m <= 128 && nw_trace_fits<4>(m, n) ? 1 : 0; }
