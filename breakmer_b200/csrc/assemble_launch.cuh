// Launchers of assemble_kernel<W>, one translation unit per width (asm_w1.cu ... asm_w8.cu) so that the four
// instantiations compile in parallel; api.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>

namespace bk {
struct AsmParams;
// sets the dynamic shared-memory limit and carve-out preference of the kernel, then launches it
cudaError_t launch_assemble_w1(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w2(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4c4(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w4c5(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
cudaError_t launch_assemble_w8(const AsmParams& A, int grid, int dyn_smem, int carveout, cudaStream_t st);
}  // namespace bk
