// Development probe: cycles per TMA tensor load for box shapes used by the implicit-GEMM kernel (one CTA per SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe scripts/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma5(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int STAGES = 8;
// mode 0: 5-D map [C][W][H][B][1], box {kw, 128, 1, 1, 1}; mode 1: 2-D map [C][M], box {kw, 128}
__global__ void probe(const __grid_constant__ CUtensorMap tm, int mode, int box_bytes, int iters, int W, int H, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    long long t0 = clock64();
    for (int i = 0; i < iters + STAGES; ++i) {
      int s = i % STAGES;
      if (i >= STAGES) mbar_wait(smem_u32(&bar[s]), ((i / STAGES) - 1) & 1);
      if (i < iters) {
        mbar_expect(smem_u32(&bar[s]), box_bytes);
        int tile = (blockIdx.x * iters + i);
        if (mode == 0) {
          int xs = tile % (W / 128), y = (tile / (W / 128)) % H, b = tile / ((W / 128) * H);
          tma5(base + s * 16384, &tm, smem_u32(&bar[s]), 0, xs * 128, y, b, 0);
        } else if (mode == 2) {
          int xs = tile % (W / 128), yb = tile / (W / 128);
          tma3(base + s * 16384, &tm, smem_u32(&bar[s]), 0, xs * 128, yb);
        } else {
          tma2(base + s * 16384, &tm, smem_u32(&bar[s]), 0, tile * 128);
        }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &sym, 12000, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)sym;
  const int B = 16, H = 512, W = 512;
  int sms = 148, iters = 400;
  long long* out; cudaMalloc(&out, sms * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 16384 + 2048);
  int chans[] = {16, 32, 40, 64, 128};
  for (int ci = 0; ci < 5; ++ci) for (int mode = 0; mode < 3; ++mode) {
    int C = chans[ci];
    int kw = C <= 16 ? 16 : (C < 96 ? 32 : 64);
    void* x; size_t n = (size_t)B * H * W * C * 2; cudaMalloc(&x, n); cudaMemset(x, 0, n);
    CUtensorMap tm;
    CUtensorMapSwizzle sw = kw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r;
    if (mode == 0) {
      cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
      cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)B * H * W * C * 2};
      cuuint32_t box[5] = {(cuuint32_t)kw, 128, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (mode == 2) {
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)B * H};
      cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
      cuuint32_t box[3] = {(cuuint32_t)kw, 128, 1}, es[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)B * H * W};
      cuuint64_t str[1] = {(cuuint64_t)C * 2};
      cuuint32_t box[2] = {(cuuint32_t)kw, 128}, es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    int box_bytes = kw * 2 * 128;
    for (int rep = 0; rep < 2; ++rep) probe<<<sms, 32, STAGES * 16384 + 2048>>>(tm, mode, box_bytes, iters, W, H, out);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    printf("C=%3d kw=%2d mode=%s box=%5d B : %.0f cycles/load  (%.1f B/cycle/SM useful, 148 SMs concurrently)\n", C, kw, mode == 0 ? "5D" : (mode == 1 ? "2D" : "3D"), box_bytes, avg / iters, box_bytes / (avg / iters));
    cudaFree(x);
  }
  return 0;
}
