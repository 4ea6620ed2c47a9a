// Encoder stem on the tensor cores (reference efficientnet_pytorch/model.py:309-311 in eval mode with
// BN folded, preceded -- uint8 variant -- by the ToTensor / Normalize / roll / crop of train_VIGOR.py:55-70, 272-273):
//
//   out[b, ho + lo, wo + lo, :] = SiLU( sum_{ci, ky, kx} W[:, ci, ky, kx] * x[b, ci, 2 ho + ky - in_lo, 2 wo + kx - in_lo] + bias )
//
// The CUDA-core version (stem_conv_silu_kernel, encoder_ops.cu) issues 864 packed FMAs per pixel pair and measured 0.32 ms
// for the 512 x 512 aerial batch of 64 (1.4 TB/s, a quarter of what the bytes need).  Here the 3x3x3 window is an implicit
// GEMM with K = 27 (padded to 32), N = 32 output channels, M = 128 pixels of one output row:
//   * the 128 threads of a CTA each gather one pixel's 27 inputs (planar fp32, or uint8 with the normalisation applied
//     on the fly), round them to bf16 and store the pixel's 64-byte row of the K-major, 64B-swizzled A tile with four
//     16-byte shared stores (chunk index XOR (row / 2) % 4: the pattern TMA would have written);
//   * one elected thread issues two tcgen05.mma (K = 16 each) against the resident [32 x 32] weight tile; accumulators are
//     double buffered in TMEM, so the MMAs of tile i run under the epilogue of tile i - 1;
//   * epilogue: thread = pixel: tcgen05.ld of its 32 accumulators, bias + SiLU, one 64-byte channels-last store (plus
//     the wrap-around columns of the circularly padded panorama buffer).
// No TMA (the planar, stride-2, possibly uint8 input is not a box), many small CTAs per SM instead of warp specialisation.
// Operands are rounded to bf16 like every other layer of the bf16 plan; accumulation is fp32.
#include "tcgen05_common.cuh"

namespace ccvpe {

constexpr int STEM_CO = 32;

template <bool CIRC, bool U8>
__global__ void __launch_bounds__(128) stem_tcgen05_kernel(const __grid_constant__ StemTcParams p) {
  __shared__ __align__(1024) uint8_t s_a[2][TC_BM * 64];
  __shared__ __align__(1024) uint8_t s_b[STEM_CO * 64];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[STEM_CO];

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t swz = (uint32_t)((tid >> 1) & 3);       // 64B swizzle of row `tid`: 16-byte chunk c lives at c ^ swz

  {  // weights: thread = (n = tid / 4, chunk = tid % 4): Bs[n][k0 .. k0 + 8) = W[k][n], k = (ci * 3 + ky) * 3 + kx, zero for k >= 27
    const int n = tid >> 2, c = tid & 3;
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = c * 8 + 2 * j;
      const float w0 = k < 27 ? __ldg(p.w + k * STEM_CO + n) : 0.f;
      const float w1 = k + 1 < 27 ? __ldg(p.w + (k + 1) * STEM_CO + n) : 0.f;
      h[j] = __floats2bfloat162_rn(w0, w1);
    }
    *reinterpret_cast<uint4*>(s_b + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = q;
    if (tid < STEM_CO) s_bias[tid] = 0.5f * __ldg(p.bias + tid);      // SiLU(a) = h + h tanh(h), h = a / 2
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(STEM_CO >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint64_t hi = make_smem_desc(0, 32) & 0xFFFFFFFF00000000ull;
  const uint64_t bdesc = hi | (((smem_u32(s_b) & 0x3FFFFu) >> 4) | (1u << 16));

  int roll_cache_b = -1, roll = 0;
  auto build = [&](int tile, int s) {
    const int xs = tile % p.strips;
    const int q = tile / p.strips;
    const int ho = q % p.Ho, b = q / p.Ho;
    const int wo = xs * TC_BM + tid;
    if (U8 && p.shift && b != roll_cache_b) {
      roll = p.shift[b] % p.Wsrc;
      if (roll < 0) roll += p.Wsrc;
      roll_cache_b = b;
    }
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    const bool valid = wo < p.Wo;
    float v[28];
    v[27] = 0.f;
    auto load1 = [&](int ci, int ih, int iw) -> float {       // one in-range input sample (normalised on the fly for uint8)
      if (U8) {
        int ws = iw - roll;                                   // torch.roll: out[w] = in[(w - shift) mod Wsrc]
        if (ws < 0) ws += p.Wsrc;
        return fmaf((float)__ldg(p.x8 + ((int64_t)(b * 3 + ci) * p.H + ih) * p.Wsrc + ws), p.sc[ci], p.sh[ci]);
      }
      return __ldg(p.x + ((int64_t)(b * 3 + ci) * p.H + ih) * p.W + iw);
    };
    // (A variant with one 8-byte load for columns 2 wo, 2 wo + 1 and the third column taken from the neighbouring lane cut the
    // L1 sector traffic three-fold but measured 3-45 % SLOWER on every shape -- profiles/r03_stem_ab.txt -- and was dropped.)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ih = 2 * ho + ky - p.in_lo;
        const bool row_ok = valid && ih >= 0 && ih < p.H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          int iw = 2 * wo + kx - p.in_lo;
          bool ok = row_ok;
          if (CIRC) iw = iw < 0 ? iw + p.W : (iw >= p.W ? iw - p.W : iw);
          else ok = ok && iw >= 0 && iw < p.W;
          v[(ci * 3 + ky) * 3 + kx] = ok ? load1(ci, ih, iw) : 0.f;
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
    uint8_t* row = s_a[s] + tid * 64;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(row + ((c ^ swz) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  };

  auto epilogue = [&](int tile, int s, uint32_t phase) {
    const int xs = tile % p.strips;
    const int q = tile / p.strips;
    const int ho = q % p.Ho, b = q / p.Ho;
    const int wo = xs * TC_BM + tid;
    mbar_wait(smem_u32(&bar[s]), phase);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * STEM_CO), v);
    tmem_ld_wait();
    tc_fence_before();
    if (wo >= p.Wo) return;
    uint4 o[4];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
    for (int c = 0; c < STEM_CO; c += 2) {
      const float a0 = fmaf(__uint_as_float(v[c]), 0.5f, s_bias[c]), a1 = fmaf(__uint_as_float(v[c + 1]), 0.5f, s_bias[c + 1]);
      h[c >> 1] = __floats2bfloat162_rn(fmaf(a0, tanh_approx(a0), a0), fmaf(a1, tanh_approx(a1), a1));
    }
    __nv_bfloat16* orow = p.out + ((int64_t)b * p.Hp + ho + p.out_lo) * p.Wp * STEM_CO;
    uint4* dst = reinterpret_cast<uint4*>(orow + (int64_t)(wo + p.out_lo) * STEM_CO);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = o[i];
    if (CIRC) {   // wrap columns of the circularly padded image: [0, out_lo) <- last columns, [out_lo + Wo, Wp) <- first ones
      const int hi_pad = p.Wp - p.out_lo - p.Wo;
      if (wo >= p.Wo - p.out_lo) {
        uint4* d2 = reinterpret_cast<uint4*>(orow + (int64_t)(wo - (p.Wo - p.out_lo)) * STEM_CO);
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[i] = o[i];
      }
      if (wo < hi_pad) {
        uint4* d2 = reinterpret_cast<uint4*>(orow + (int64_t)(p.out_lo + p.Wo + wo) * STEM_CO);
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[i] = o[i];
      }
    }
  };

  int it = 0, prev = -1;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    build(tile, s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t adesc = hi | (((smem_u32(s_a[s]) & 0x3FFFFu) >> 4) | (1u << 16));
        const uint32_t tmem_d = tmem_base + (uint32_t)(s * STEM_CO);
        umma_bf16(tmem_d, adesc, bdesc, idesc, 0u);
        umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
        umma_commit(smem_u32(&bar[s]));
      }
      __syncwarp();
    }
    if (prev >= 0) epilogue(prev, s ^ 1, (uint32_t)((it - 1) >> 1) & 1u);
    prev = tile;
  }
  if (prev >= 0) epilogue(prev, (it - 1) & 1, (uint32_t)((it - 1) >> 1) & 1u);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

int stem_tcgen05(StemTcParams p, bool circular, bool u8, cudaStream_t st) {
  p.strips = (p.Wo + TC_BM - 1) / TC_BM;
  const int64_t total = (int64_t)p.B * p.Ho * p.strips;
  if (total >= (1LL << 31)) return fail(CCVPE_ERR_BAD_ARGUMENT, "stem: too many tiles");
  p.total_tiles = (int)total;
  const int max_grid = 8 * sm_count();          // 64 TMEM columns per CTA -> at most eight resident CTAs per SM
  const int grid = p.total_tiles < max_grid ? p.total_tiles : max_grid;
  if (circular) {
    if (u8) stem_tcgen05_kernel<true, true><<<grid, 128, 0, st>>>(p);
    else stem_tcgen05_kernel<true, false><<<grid, 128, 0, st>>>(p);
  } else {
    if (u8) stem_tcgen05_kernel<false, true><<<grid, 128, 0, st>>>(p);
    else stem_tcgen05_kernel<false, false><<<grid, 128, 0, st>>>(p);
  }
  return check_launch("stem_tcgen05_kernel");
}

}  // namespace ccvpe
