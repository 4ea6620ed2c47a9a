// Library-wide state: error string, launch counter, device properties.
#include "common.cuh"

namespace ccvpe {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> n{0};      // process wide: autograd runs the backward kernels on its own thread
  return n;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace ccvpe

extern "C" {
int ccvpe_abi_version(void) { return CCVPE_ABI_VERSION; }
const char* ccvpe_last_error(void) { return ccvpe::last_error_buffer(); }
int64_t ccvpe_launch_count(void) { return ccvpe::launch_counter().load(); }
void ccvpe_reset_launch_count(void) { ccvpe::launch_counter().store(0); }
}
