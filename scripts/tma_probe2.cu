// Development probe 2: what bounds the activation loads of the row-ring convolution?  All 148 SMs stream distinct
// 130-pixel row segments of a channels-last bf16 image (> L2) with 8 loads in flight per SM, one issuing thread per CTA:
//   tensor : cp.async.bulk.tensor.3d box [kw ch x 130 px x 1 row] (what conv_ring_tcgen05_kernel issues), pixel pitch ld
//   bulk   : cp.async.bulk (1-D) of the same segment's 130 * C * 2 contiguous bytes
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tma_probe2.bin scripts/tma_probe2.cu -lcuda
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
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int STAGES = 8;
constexpr int SLOT = 26 * 1024;   // (two 64-wide boxes of 130 rows: 2 x 17408 B would need 34 KB -- those configs are skipped)
// mode 0: tensor boxes (nblk boxes of kw channels per segment); mode 1: one 1-D bulk copy per segment
__global__ void probe(const __grid_constant__ CUtensorMap tm, const uint8_t* x, int mode, int C, int ld, int kw, int nblk, int iters, int W,
                      int rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int strips = W / 128;
    const uint32_t tx = mode == 0 ? (uint32_t)(nblk * kw * 2 * 130) : (uint32_t)(129 * C * 2);
    long long t0 = clock64();
    for (int i = 0; i < iters + STAGES; ++i) {
      int s = i % STAGES;
      if (i >= STAGES) mbar_wait(smem_u32(&bar[s]), ((i / STAGES) - 1) & 1);
      if (i < iters) {
        mbar_expect(smem_u32(&bar[s]), tx);
        // consecutive CTAs take consecutive segments of the image (like the ring kernel's strips), iteration i moves down
        // (32-bit index math only: the first version's 64-bit div / mod cost ~600 cycles per iteration and hid everything)
        const unsigned seg = (unsigned)i * gridDim.x + blockIdx.x;
        const int xs = (int)(seg & (unsigned)(strips - 1));           // strips is a power of two (W = 512)
        const unsigned row = (seg >> 2) & (unsigned)(rows - 1);        // rows is a power of two
        if (mode == 0) {
          for (int b = 0; b < nblk; ++b) tma3(base + s * SLOT + b * ((130 * kw * 2 + 1023) / 1024 * 1024), &tm, smem_u32(&bar[s]), b * kw, xs * 128 - 1, (int)row);
        } else {
          int x0 = xs * 128 - 1; if (x0 < 0) x0 = 0;      // (129 pixels: stays inside the row at both borders)
          bulk1d(base + s * SLOT, x + ((size_t)row * W + x0) * ld * 2, 129 * C * 2, smem_u32(&bar[s]));
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
  int sms = 148, iters = 600;
  long long* out; cudaMalloc(&out, sms * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * SLOT + 2048);
  // {C, ld, kw}
  int cfg[][3] = {{16, 16, 16}, {32, 32, 32}, {40, 40, 32}, {40, 40, 64}, {40, 64, 64}, {64, 64, 64}, {16, 64, 16}};
  const int W = 512;
  for (auto& c : cfg) {
    const int C = c[0], ld = c[1], kw = c[2];
    int rows = (int)((size_t)400 * 1024 * 1024 / ((size_t)W * ld * 2));   // 400 MB image stack: every segment is cold
    { int p2 = 1; while (p2 * 2 <= rows) p2 *= 2; rows = p2; }             // (>= 200 MB: still larger than L2)
    void* x; size_t n = (size_t)rows * W * ld * 2; cudaMalloc(&x, n); cudaMemset(x, 0, n);
    CUtensorMap tm;
    CUtensorMapSwizzle sw = kw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)rows};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)kw, 130, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int nblk = (C + kw - 1) / kw;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 1 && ld != C) continue;
      for (int rep = 0; rep < 2; ++rep) probe<<<sms, 32, STAGES * SLOT + 2048>>>(tm, (const uint8_t*)x, mode, C, ld, kw, nblk, iters, W, rows, out);
      cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
      const double useful = 130.0 * C * 2;
      printf("C=%3d ld=%3d kw=%2d %-6s: %6.0f cycles per 130-px segment (%d box rows)  %.1f useful B/cycle/SM = %.2f TB/s at 1.9 GHz x 148\n", C, ld, kw,
             mode == 0 ? "tensor" : "bulk", avg / iters, mode == 0 ? nblk * 130 : 1, useful / (avg / iters), useful / (avg / iters) * 148 * 1.9e9 / 1e12);
    }
    cudaFree(x);
  }
  return 0;
}
