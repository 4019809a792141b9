// assemble_kernel<4, 5>: width 4 with the register allocation bounded for 5 resident CTAs per SM.
#include "assemble.cuh"
#include "assemble_launch.cuh"

namespace bk {
cudaError_t launch_assemble_w4c5(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st) {
  std::lock_guard<std::mutex> hold(launch_attr_mutex());
  cudaError_t e = cudaFuncSetAttribute(assemble_kernel<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(assemble_kernel<4, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  if (e != cudaSuccess) return e;
  assemble_kernel<4, 5><<<grid, 128, dyn_smem, st>>>(A);
  return cudaGetLastError();
}
}  // namespace bk
