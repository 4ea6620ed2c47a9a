// Development probe: how fast can 128-thread "one output row per thread" epilogues push bf16 rows to HBM?
//   mode 0: each thread stores its row with 16-byte st.global (rows rowbytes apart: 32 different lines per warp instr)
//   mode 1: each thread writes its row to shared memory and issues ONE cp.async.bulk (TMA 1-D bulk store) for it
//   mode 2: rows staged in shared memory, then warp-cooperative coalesced 16-byte st.global (lanes walk the row)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/store_probe.bin scripts/store_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint8_t* out, int mode, int rowbytes, int pitch, int tiles_per_cta) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int t = threadIdx.x;
  uint8_t* my = smem + (size_t)t * pitch;
  const uint4 v = make_uint4(t, t + 1, t + 2, t + 3);
  for (int i = 0; i < tiles_per_cta; ++i) {
    const size_t tile = (size_t)blockIdx.x * tiles_per_cta + i;
    uint8_t* g = out + (tile * 128 + t) * (size_t)rowbytes;
    if (mode == 0) {
      for (int c = 0; c < rowbytes; c += 16) *reinterpret_cast<uint4*>(g + c) = v;
    } else if (mode == 1) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // my previous row has left shared memory
      for (int c = 0; c < rowbytes; c += 16) *reinterpret_cast<uint4*>(my + c) = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(my)), "r"(rowbytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else {
      for (int c = 0; c < rowbytes; c += 16) *reinterpret_cast<uint4*>(my + c) = v;
      __syncwarp();
      // the warp's 32 rows are contiguous in global memory: lanes walk them 16 bytes at a time
      const int warp = t >> 5, lane = t & 31;
      uint8_t* gw = out + (tile * 128 + warp * 32) * (size_t)rowbytes;
      const int per_row = rowbytes >> 4, total = 32 * per_row;
      for (int q = lane; q < total; q += 32) {
        const int r = q / per_row, c = q - r * per_row;
        *reinterpret_cast<uint4*>(gw + (size_t)q * 16) = *reinterpret_cast<const uint4*>(smem + (size_t)(warp * 32 + r) * pitch + c * 16);
      }
      __syncwarp();
    }
  }
  if (mode == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const int ctas_per_sm = argc > 1 ? atoi(argv[1]) : 2;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm;
  const size_t total = 2ull << 30;   // 2 GB written per launch
  uint8_t* out;
  cudaMalloc(&out, total + (1 << 20));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rowbytes : {32, 64, 128, 192, 256, 512}) {
    for (int mode = 0; mode < 3; ++mode) {
      const int pitch = rowbytes + (((rowbytes >> 4) & 1) ? 0 : 16);
      const int tiles_per_cta = (int)(total / ((size_t)grid * 128 * rowbytes));
      const size_t smem = (size_t)128 * pitch;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      probe<<<grid, 128, smem>>>(out, mode, rowbytes, pitch, tiles_per_cta);
      cudaEventRecord(e0);
      probe<<<grid, 128, smem>>>(out, mode, rowbytes, pitch, tiles_per_cta);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)grid * tiles_per_cta * 128 * rowbytes;
      printf("rowbytes %4d mode %d (%s): %.3f ms  %.0f GB/s  [%s]\n", rowbytes, mode,
             mode == 0 ? "row-per-thread STG" : (mode == 1 ? "per-row bulk store" : "smem + coalesced STG"), ms, bytes / ms / 1e6,
             cudaGetErrorString(err));
    }
  }
  return 0;
}
