// Phase timing of one alignment on one warp (experiments only): cycles of the score pass, the end-cell scans and the
// traceback of nw.cuh's nw_dual_trace, and of the packed-cell kernel, for a read overlapping the end of a contig.
#define BK_NW_PROF 1
#include "../breakmer_b200/csrc/nw.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
using namespace bk;
__global__ void k(const uint8_t* rd, int m, const uint8_t* ct, int n, uint2* lastcol, uint8_t* tab, int2* edge, int mode, int* res, long long* cyc) {
  __shared__ __align__(16) uint8_t s1[4096], s2[4096];
  for (int x = threadIdx.x; x < m; x += 32) s1[x] = rd[x];
  for (int x = threadIdx.x; x <= n; x += 32) s2[x] = x < n ? ct[x] : 0;
  __syncwarp();
  NwDual r;
  long long t0 = clock64();
  if (mode == 0) nw_dual_dispatch<false>(s1, m, s2, n, edge, edge + 4096, lastcol, tab, r);
  else if (mode == 1) nw_dual_dispatch<true>(s1, m, s2, n, edge, edge + 4096, lastcol, tab, r);
  else nw_dual_dispatch<false>(s1, m, s2, n, edge, edge + 4096, lastcol, nullptr, r);
  long long t1 = clock64();
  if (threadIdx.x == 0) {
    res[0] = r.a.score; res[1] = r.a.j0; res[2] = r.a.i0; res[3] = r.b.score; res[4] = r.b.j0; res[5] = r.b.i0;
    cyc[0] = t1 - t0;
    for (int i = 0; i < 4; ++i) cyc[1 + i] = bk_nw_prof[i] - t0;
  }
}
int main(int argc, char** argv) {
  int m = argc > 1 ? atoi(argv[1]) : 100, n = argc > 2 ? atoi(argv[2]) : 200, over = argc > 3 ? atoi(argv[3]) : 40;
  char* g = (char*)malloc(n + m + 1);
  srand(7);
  for (int i = 0; i < n + m; ++i) g[i] = "ACGT"[rand() & 3];
  uint8_t *rd, *ct, *tab; uint2* lc; int2* edge; int* res; long long* cyc;
  cudaMalloc(&rd, m); cudaMalloc(&ct, n + 1); cudaMalloc(&tab, NW_TAB_BYTES); cudaMalloc(&lc, 8 * 4096); cudaMalloc(&edge, 16 * 8192);
  cudaMalloc(&res, 64); cudaMalloc(&cyc, 64);
  cudaMemcpy(ct, g, n, cudaMemcpyHostToDevice);
  cudaMemcpy(rd, g + n - (m - over), m, cudaMemcpyHostToDevice);      // read overlaps the contig's last m - over bases
  const char* names[] = {"trace (both)", "trace (lazy)", "packed"};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 3; ++mode) {
      k<<<1, 32>>>(rd, m, ct, n, lc, tab, edge, mode, res, cyc);
      int hres[6]; long long hc[5];
      cudaMemcpy(hres, res, sizeof hres, cudaMemcpyDeviceToHost);
      cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
      if (rep) printf("%-14s total %7lld cycles | start %lld pass1-end %lld ends %lld done %lld | A(s=%d j0=%d i0=%d) B(s=%d j0=%d i0=%d)\n",
                      names[mode], hc[0], hc[1], hc[2], hc[3], hc[4], hres[0], hres[1], hres[2], hres[3], hres[4], hres[5]);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
