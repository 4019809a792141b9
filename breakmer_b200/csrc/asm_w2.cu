// assemble_kernel<2> (assemble.cuh) and its launcher.
#include "assemble.cuh"
#include "assemble_launch.cuh"

namespace bk {
cudaError_t launch_assemble_w2(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st) {
  std::lock_guard<std::mutex> hold(launch_attr_mutex());
  cudaError_t e = cudaFuncSetAttribute(assemble_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(assemble_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
  if (e != cudaSuccess) return e;
  assemble_kernel<2><<<grid, 32 * 2, dyn_smem, st>>>(A);
  return cudaGetLastError();
}
}  // namespace bk
