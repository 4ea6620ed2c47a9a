// Weight gradient of the implicit GEMM on tensor cores (bf16 operands, fp32 accumulation in TMEM) -- see the kernel comment.
#include "tcgen05_common.cuh"

namespace ccvpe {

int wgrad_tcgen05_supported(const ccvpe_wgrad_desc& d) { (void)d; return 0; }
int64_t wgrad_tcgen05_workspace_elems(const ccvpe_wgrad_desc& d) { (void)d; return 0; }
int wgrad_tcgen05(const ccvpe_wgrad_desc& d, cudaStream_t st) {
  (void)d; (void)st;
  return fail(CCVPE_ERR_UNSUPPORTED, "wgrad_tcgen05: not built");
}

}  // namespace ccvpe
