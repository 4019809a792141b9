// assemble_kernel<4, 4>: width 4 with the register allocation bounded for 4 resident CTAs per SM.
#include "assemble.cuh"
#include "assemble_launch.cuh"

namespace bk {
cudaError_t launch_assemble_w4c4(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st) {
  std::lock_guard<std::mutex> hold(launch_attr_mutex());
  cudaError_t e = cudaFuncSetAttribute(assemble_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(assemble_kernel<4, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  if (e != cudaSuccess) return e;
  assemble_kernel<4, 4><<<grid, 128, dyn_smem, st>>>(A);
  return cudaGetLastError();
}
}  // namespace bk
