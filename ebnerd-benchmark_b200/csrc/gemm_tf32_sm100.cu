// placeholder until the tcgen05 kernel lands
#include "ebk_common.cuh"
namespace ebk {
int gemm_tf32(const GemmOperandA&, const float*, int, bool, float*, int, int, int, int, float, cudaStream_t) {
  set_error("gemm_tf32: not built yet");
  return EBK_ERR_UNSUPPORTED;
}
}  // namespace ebk
