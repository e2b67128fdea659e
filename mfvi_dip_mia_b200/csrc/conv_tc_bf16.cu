// bf16 instantiation of the per-tap weight-gradient kernel (conv_tc_wgrad.cuh), in a translation unit of its own so that the
// tf32 kernels of conv_tc.cu compile to the same machine code as before (bf16-operand mode, DESIGN.md section 8, stage B).
#include "conv_tc_wgrad.cuh"

namespace mfvi {
namespace tc {

cudaError_t wgrad_tc_bf16_set_smem(int bytes) {
  return cudaFuncSetAttribute(k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t wgrad_tc_bf16_launch(dim3 grid, size_t smem, cudaStream_t st, const CUtensorMap& tmDy, const CUtensorMap& tmX,
                                 const TcWgradArgs& a) {
  return launch_k(k_wgrad_tc<true>, grid, kThreads, smem, st, tmDy, tmX, a);
}

}  // namespace tc
}  // namespace mfvi
