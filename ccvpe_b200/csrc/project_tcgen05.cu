// MBConv projection of the EfficientNet-B0 encoders (reference efficientnet_pytorch/model.py:115-131):
//
//   y[b, p, :] = Wg[b] . d[b, p, :]  (+ x[b, p, :])            Wg[b] = W_proj . diag(gate[b])   (ccvpe_se_gate_scale)
//
// i.e. the squeeze-excite broadcast multiply, the 1x1 projection conv with its folded BatchNorm scale and the identity
// skip in ONE pass over the 6x-expanded depthwise output d.  Because the gate differs per image this is a batch of small
// GEMMs with PER-IMAGE weights: pixels on the UMMA M axis (128-pixel tiles that never straddle an image: rows past the
// image's last pixel are TMA zero fill and are dropped by the epilogue), output channels on N, expanded channels on K.
// The second, optional output y + bias is the block output the decoder reads as a skip (the folded BatchNorm shift is
// otherwise carried into the next block's expand bias, see fast_encoder.py).
//
//   warp 0 = TMA producer: per K block one [kw ch x 128 px] box of d[b] and one [kw ch x block_n] box of Wg[b] (3-D maps
//            over [channels, rows, image]) into a ring of stages;
//   warp 1 = MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = block_n, fp32 accumulators double buffered in
//            TMEM so that the epilogue of item i overlaps the K loop of item i + 1);
//   warps 2..5 = epilogue: tcgen05.ld -> + residual -> bf16 -> 16-byte stores (one accumulator row per thread).
// HBM-bound by the read of d (mid = 6 x cin channels per pixel against cout written): the roofline is the HBM peak.
#include "tcgen05_common.cuh"

namespace ccvpe {

constexpr int PJ_MAX_STAGES = 8;
constexpr int PJ_SMEM_BUDGET = 200 * 1024;
constexpr int PJ_EPI_WARPS = 4;
constexpr int PJ_THREADS = 64 + 32 * PJ_EPI_WARPS;

struct ProjParams {
  CUtensorMap tm_a, tm_b;
  int B, HW, mid, cout;
  int kw, nkb, tail16;      // K-block width, number of K blocks, K16 slices of the last block
  int block_n, n_tiles, m_tiles, total_items;
  int stages, a_bytes, stage_bytes, tmem_cols;
  const __nv_bfloat16* res;
  const __nv_bfloat16* bias;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
};

__global__ void __launch_bounds__(PJ_THREADS, 1) project_tcgen05_kernel(const __grid_constant__ ProjParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[PJ_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[PJ_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tmem_full[a]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[a]), PJ_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_a);
    prefetch_tmap(&p.tm_b);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int per_image = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = (uint32_t)((TC_BM + p.block_n) * p.kw * 2);
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int b = item / per_image, r = item - b * per_image;
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t dst = smem_base + (uint32_t)(stage * p.stage_bytes);
          mbar_arrive_expect_tx(full, tx);
          tma_load_3d(dst, &p.tm_a, full, kb * p.kw, mt * TC_BM, b);
          tma_load_3d(dst + (uint32_t)p.a_bytes, &p.tm_b, full, kb * p.kw, nt * p.block_n, b);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                           ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t hi = (uint32_t)(make_smem_desc(0, p.kw) >> 32);
    const uint32_t lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t lo_stage = (uint32_t)p.stage_bytes >> 4, lo_b = (uint32_t)p.a_bytes >> 4;
    const int full16 = p.kw >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&bar_tmem_empty[acc]), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.block_n);
      uint32_t accumulate = 0;
#pragma unroll 1
      for (int kb = 0; kb < p.nkb; ++kb) {
        const int nk16 = (kb == p.nkb - 1) ? p.tail16 : full16;
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        const uint32_t a_lo = lo_base + (uint32_t)stage * lo_stage;
        const uint64_t adesc = ((uint64_t)hi << 32) | a_lo;
        const uint64_t bdesc = ((uint64_t)hi << 32) | (a_lo + lo_b);
        if (elect_one()) {
          umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
          if (nk16 > 1) umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
          if (nk16 > 2) umma_bf16(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
          if (nk16 > 3) umma_bf16(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
          umma_commit(smem_u32(&bar_empty[stage]));
        }
        accumulate = 1;
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit(smem_u32(&bar_tmem_full[acc]));
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warp w owns TMEM lanes 32 * (w % 4) .. + 31 = accumulator rows =====================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const bool has_res = p.res != nullptr, has_out2 = p.out2 != nullptr;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const int b = item / per_image, r = item - b * per_image;
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      const int pix = mt * TC_BM + row;
      const bool valid = pix < p.HW;
      const int n0 = nt * p.block_n;
      const int ncols = min(p.block_n, p.cout - n0);
      const int64_t off = ((int64_t)b * p.HW + pix) * p.cout + n0;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      // the residual of the first chunk is requested before the accumulator is waited for
      uint4 rq[4];
      auto load_res = [&](int c0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rq[g] = make_uint4(0u, 0u, 0u, 0u);
          if (has_res && valid && c0 + g * 8 < ncols) rq[g] = __ldg(reinterpret_cast<const uint4*>(p.res + off + c0 + g * 8));
        }
      };
      load_res(0);
      mbar_wait(smem_u32(&bar_tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.block_n);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= ncols) {    // last chunk of this accumulator stage is in registers: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));
        }
        uint4 cur[4] = {rq[0], rq[1], rq[2], rq[3]};
        if (c0 + 32 < ncols) load_res(c0 + 32);
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int cl = c0 + g * 8;
            if (cl < ncols) {
              const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&cur[g]);
              uint4 pk;
              __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 rr = __bfloat1622float2(rh[j]);
                ph[j] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * j]) + rr.x,
                                              __uint_as_float(v[g * 8 + 2 * j + 1]) + rr.y);
              }
              *reinterpret_cast<uint4*>(p.out + off + cl) = pk;
              if (has_out2) {
                // the decoder's copy: (bf16 y) + (bf16 bias), rounded once more -- what `cur + b_proj` computes in torch
                const uint4 bq = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + cl));
                const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&bq);
                uint4 pk2;
                __nv_bfloat162* qh = reinterpret_cast<__nv_bfloat162*>(&pk2);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 y = __bfloat1622float2(ph[j]), bb = __bfloat1622float2(bh[j]);
                  qh[j] = __floats2bfloat162_rn(y.x + bb.x, y.y + bb.y);
                }
                *reinterpret_cast<uint4*>(p.out2 + off + cl) = pk2;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace ccvpe

extern "C" int ccvpe_mbconv_project_nhwc(const void* d, const void* wg, const void* residual, const void* bias, void* out,
                                         void* out_biased, int B, int HW, int mid, int cout, void* stream) {
  using namespace ccvpe;
  CCVPE_REQUIRE(d && wg && out, "ccvpe_mbconv_project_nhwc: null pointer");
  CCVPE_REQUIRE(B > 0 && HW > 0 && mid > 0 && cout > 0 && mid % 8 == 0 && cout % 8 == 0,
                "ccvpe_mbconv_project_nhwc: bad shape B=%d HW=%d mid=%d cout=%d (channel counts must be multiples of 8)", B, HW,
                mid, cout);
  CCVPE_REQUIRE(!out_biased || bias, "ccvpe_mbconv_project_nhwc: out_biased needs a bias");
  CCVPE_REQUIRE(aligned16(d) && aligned16(wg) && aligned16(out) && aligned16(residual) && aligned16(bias) && aligned16(out_biased),
                "ccvpe_mbconv_project_nhwc: pointers must be 16-byte aligned");
  CCVPE_REQUIRE((int64_t)B * HW * (int64_t)(mid > cout ? mid : cout) < (1LL << 40), "ccvpe_mbconv_project_nhwc: tensor too large");
  static thread_local ProjParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.HW = HW; p.mid = mid; p.cout = cout;
  p.kw = tc_block_width(mid);
  p.nkb = (mid + p.kw - 1) / p.kw;
  p.tail16 = (mid - (p.nkb - 1) * p.kw + 15) / 16;
  p.n_tiles = (cout + TC_MAX_N - 1) / TC_MAX_N;
  p.block_n = ((cout + p.n_tiles - 1) / p.n_tiles + 15) / 16 * 16;
  p.m_tiles = (HW + TC_BM - 1) / TC_BM;
  const int64_t total = (int64_t)B * p.m_tiles * p.n_tiles;
  CCVPE_REQUIRE(total < (1LL << 30), "ccvpe_mbconv_project_nhwc: too many tiles");
  p.total_items = (int)total;
  p.a_bytes = TC_BM * p.kw * 2;
  p.stage_bytes = p.a_bytes + (p.block_n * p.kw * 2 + 1023) / 1024 * 1024;
  p.stages = PJ_SMEM_BUDGET / p.stage_bytes;
  if (p.stages > PJ_MAX_STAGES) p.stages = PJ_MAX_STAGES;
  // two accumulator stages; the epilogue reads whole 32-column chunks, so the second stage's last chunk may reach up to
  // block_n + 32 columns past the first stage's start: keep that inside the allocation (matters for block_n = 16)
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.block_n || p.tmem_cols < p.block_n + ((p.block_n + 31) / 32) * 32) p.tmem_cols <<= 1;
  p.res = (const __nv_bfloat16*)residual;
  p.bias = (const __nv_bfloat16*)bias;
  p.out = (__nv_bfloat16*)out;
  p.out2 = (__nv_bfloat16*)out_biased;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)mid, (uint64_t)HW, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)mid * 2, (uint64_t)HW * mid * 2};
    uint32_t box[3] = {(uint32_t)p.kw, TC_BM, 1};
    if ((rc = encode_map(&p.tm_a, d, 3, dims, str, box, p.kw)) != CCVPE_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)mid, (uint64_t)cout, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)mid * 2, (uint64_t)cout * mid * 2};
    uint32_t box[3] = {(uint32_t)p.kw, (uint32_t)p.block_n, 1};
    if ((rc = encode_map(&p.tm_b, wg, 3, dims, str, box, p.kw)) != CCVPE_OK) return rc;
  }
  const int smem = p.stages * p.stage_bytes + 1024;
  static thread_local uint64_t attr = 0;
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(project_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PJ_SMEM_BUDGET + 1024);
    if (e != cudaSuccess) return fail(CCVPE_ERR_CUDA, "cudaFuncSetAttribute(project): %s", cudaGetErrorString(e));
  }
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  project_tcgen05_kernel<<<grid, PJ_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("project_tcgen05_kernel");
}
