// placeholder until the tcgen05 backend lands
#include "common.cuh"
namespace ccvpe {
bool igemm_tcgen05_supported(const ccvpe_igemm_desc&) { return false; }
int igemm_tcgen05(const ccvpe_igemm_desc&, cudaStream_t) {
  return fail(CCVPE_ERR_UNSUPPORTED, "ccvpe_igemm: tcgen05 backend not built into this library version");
}
}  // namespace ccvpe
