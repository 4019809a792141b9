// Measured ceiling of the overlap DP: the rate at which the DP's own cell update issues on a full B200 when nothing
// else is in the way -- same instruction sequence as the score pass of breakmer_b200/csrc/nw.cuh (compare the two
// characters, add 5 or 2 to the diagonal value, three-way max with the two gap moves), C = 4 columns per thread in
// registers, two rows per step, no shuffles, no stores, 16 warps per SM sub-partition group so that dependent-chain
// latency is hidden.  bench.py uses the result (profiles/r2_int_peak.json) as the denominator of roofline_alu.
// Not part of the product library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(512) cell_rate_kernel(int steps, const uint8_t* __restrict__ rows, int* __restrict__ sink) {
  constexpr int C = 4;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int col[C], r0[C], ch[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { col[c] = 2 * (c + 1); r0[c] = 0; ch[c] = (tid * 7 + c * 3) & 3; }
  int diag = 0, h0 = 2, h1 = 4;
  for (int t = 0; t < steps; ++t) {
    const int rc0 = rows[(2 * t) & 1023], rc1 = rows[(2 * t + 1) & 1023];
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
    diag = h1; h0 += 4; h1 += 4;
  }
  int s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += col[c] + r0[c];
  if (s == 0x7fffffff) sink[0] = s;             // keeps the loop alive
}

extern "C" int int_peak_run(int device, int steps, int blocks_per_sm, double* cells_per_s, double* ms_out, int* sm_count_out) {
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  uint8_t* rows; int* sink;
  cudaMalloc(&rows, 1024); cudaMalloc(&sink, 4);
  uint8_t hrows[1024];
  for (int i = 0; i < 1024; ++i) hrows[i] = (uint8_t)((i * 2654435761u >> 7) & 3);
  cudaMemcpy(rows, hrows, 1024, cudaMemcpyHostToDevice);
  const int grid = sms * blocks_per_sm;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cell_rate_kernel<<<grid, 512>>>(steps / 8, rows, sink);      // warm-up
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    cell_rate_kernel<<<grid, 512>>>(steps, rows, sink);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) return -2;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  *ms_out = best;
  *cells_per_s = (double)grid * 512.0 * 8.0 * (double)steps / (best * 1e-3);
  *sm_count_out = sms;
  cudaFree(rows); cudaFree(sink);
  cudaEventDestroy(a); cudaEventDestroy(b);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
