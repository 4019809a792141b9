// Launchers of assemble_kernel<W>, one translation unit per width (asm_w1.cu ... asm_w8.cu) so that the four
// instantiations compile in parallel; api.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>

#include <mutex>

namespace bk {
struct AsmParams;
// Function attributes (dynamic shared-memory limit, carve-out) belong to the kernel, not to a launch: two host threads
// that set them to different values for their own batches and then launch would race (the launch of one is validated
// against the limit the other just set: "too many resources requested for launch").  Every "set attributes, launch"
// pair of the library runs under this one mutex; the launches themselves are asynchronous, so it is held for microseconds.
inline std::mutex& launch_attr_mutex() {
  static std::mutex m;
  return m;
}
// sets the dynamic shared-memory limit and carve-out preference of the kernel, then launches it
cudaError_t launch_assemble_w1(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w2(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4c4(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4c5(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w8(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
}  // namespace bk
