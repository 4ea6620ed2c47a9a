// a4/a5/a6 on tensor cores -- the rolled cosine correlation as a TMA-fed tcgen05 GEMM (bf16 path; full-circle L == C and
// windowed L < C: limited FoV, KITTI, Oxford's centred window).
//
//   dot[b, p, i] = sum_c x[b, p, c] * G[b, i, c],      G = banded circulant of the ground descriptor (match.cu)
//
// Pixels sit on the UMMA M axis (128 per tile), the R <= 32 orientations on N (padded to 32), channels on K.  A tiles
// [128 px x kw ch] arrive by 2-D TMA straight from the channels-last map, G tiles [32 x kw] likewise; accumulators live in
// TMEM (double buffered, 2 x 32 columns).  While the tensor pipe consumes a K block, four "norm" warps read the same
// shared-memory tile (one thread per pixel, chunk order rotated per row so the 128-bit loads are conflict free) and
// accumulate sum_c x^2, which gives both the cosine denominator and the 1/||x|| that F.normalize needs.  The same warps
// run the epilogue: tcgen05.ld -> divide -> planar fp32 score stores (one pixel per lane: coalesced per orientation),
// max over the selected orientations, 1/norm, and (bottleneck level only) the channels-last score copy and x-hat.
// Windowed levels (L < C) need one denominator PER ORIENTATION: sum of x^2 over the channels of window i.  All windows of
// a level start and end on multiples of u = gcd(C, L, window bases), so the norm warps accumulate
//     wsq[i] += m[unit(c)][i] * x_c^2          m = 0/1 window-membership table, one row per UNIT of channels,
// with the table in shared memory (rows read as broadcast LDS.128) and the R accumulators packed in fp32 pairs (FFMA2).
// UNIT = a whole K block when u % kw == 0 (KITTI levels 1-3, FoV levels 1-2: one table row per K block, R/2 FFMA2 per 64
// channels), 8 channels when u % 8 == 0, single channels otherwise (Oxford's levels 3-6, limited-FoV levels 5-6).
// The correlation moves 2*R flops per loaded element (AI 8-20 flop/B in bf16) so the roofline that bounds it is HBM;
// the tensor pipe is what keeps the arithmetic off the critical path.
#include "tcgen05_common.cuh"

namespace ccvpe {

constexpr int MT_THREADS = 256;     // warps 0-3 control, 4-7 norm + epilogue
constexpr int MT_MAX_STAGES = 8;
constexpr int MT_N = 32;

struct MatchTcParams {
  CUtensorMap tm_x, tm_g;
  int kw, nb, C, HW, B, tiles_per_img, total_tiles, stages, a_bytes, stage_bytes;
  int n_rolls, ld_scores_cl;
  int wt_rows, wt_pitch;                 // windowed: membership table [wt_rows][wt_pitch] floats (pitch/4 odd: conflict free)
  const float* wtab;
  uint32_t max_mask;
  const float* gnorm;
  const __nv_bfloat16* x;
  float* scores;
  __nv_bfloat16* scores_cl;
  float* max_out;
  float* inv_norm;
  __nv_bfloat16* xhat;
};

struct ShiftArg {
  int s[MT_N];
};

__global__ void build_rolled_descriptor_bf16_kernel(const float* __restrict__ g, int L, int C, int offset,
                                                    const ShiftArg shifts, int n_rolls,
                                                    __nv_bfloat16* __restrict__ G, float* __restrict__ gnorm,
                                                    float* __restrict__ wtab, int wt_rows, int wt_pitch, int wt_unit) {
  const int b = blockIdx.y, i = blockIdx.x;
  if (i > MT_N) {
    // window-membership table (batch independent: written by the b == 0 blocks): row r covers channels
    // [r * wt_unit, (r + 1) * wt_unit), which lie entirely inside or outside every window by construction of wt_unit
    if (b != 0 || !wtab) return;
    for (int e = (i - MT_N - 1) * blockDim.x + threadIdx.x; e < wt_rows * wt_pitch; e += (gridDim.x - MT_N - 1) * blockDim.x) {
      const int r = e / wt_pitch, col = e - r * wt_pitch;
      const int c = r * wt_unit;
      float m = 0.f;
      if (col < n_rolls && c < C) {
        const int base = ((offset + shifts.s[col]) % C + C) % C;
        int k = c - base;
        if (k < 0) k += C;
        m = k < L ? 1.f : 0.f;
      }
      wtab[e] = m;
    }
    return;
  }
  if (i < MT_N) {
    __nv_bfloat16* row = G + ((int64_t)b * MT_N + i) * C;
    if (i < n_rolls) {
      const int base = ((offset + shifts.s[i]) % C + C) % C;
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int k = c - base;
        if (k < 0) k += C;
        row[c] = __float2bfloat16_rn(k < L ? g[(int64_t)b * L + k] : 0.f);
      }
    } else {
      for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] = __float2bfloat16_rn(0.f);
    }
  } else {
    __shared__ float red[32];
    float s = 0.f;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
      float v = g[(int64_t)b * L + k];
      s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) gnorm[b] = sqrtf(t);
    }
  }
}

// UNIT: 0 = full circle (one norm for all orientations) | 1 / 8 = membership table row per channel / per 8 channels |
//       64 = one row per K block
constexpr int MT_WMAX = 24;              // orientations of a windowed level (C ABI limit)

template <int UNIT>
__device__ __forceinline__ void window_accumulate(uint64_t (&wsq)[MT_WMAX / 2], float s, const float* __restrict__ row,
                                                  int pitch) {
  const uint64_t ss = pack_f32x2(s, s);
#pragma unroll
  for (int i4 = 0; i4 < MT_WMAX / 4; ++i4) {
    if (i4 * 4 < pitch) {                  // uniform
      const float4 m = *reinterpret_cast<const float4*>(row + i4 * 4);
      wsq[2 * i4] = ffma2(ss, pack_f32x2(m.x, m.y), wsq[2 * i4]);
      wsq[2 * i4 + 1] = ffma2(ss, pack_f32x2(m.z, m.w), wsq[2 * i4 + 1]);
    }
  }
}

template <int UNIT>
__global__ void __launch_bounds__(MT_THREADS, UNIT == 0 ? 4 : 2) match_tcgen05_kernel(const __grid_constant__ MatchTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MT_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[MT_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1 + 4);     // MMA commit + one arrival per norm warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tmem_full[a]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[a]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_x);
    prefetch_tmap(&p.tm_g);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(64)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float* s_wtab = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * p.stage_bytes);
  if (UNIT != 0) {
    const int n4 = (p.wt_rows * p.wt_pitch) >> 2;
    for (int e = threadIdx.x; e < n4; e += MT_THREADS)
      reinterpret_cast<float4*>(s_wtab)[e] = __ldg(reinterpret_cast<const float4*>(p.wtab) + e);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, single-lane issue via elect_one) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)((TC_BM + MT_N) * p.kw * 2);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_img;
        const int prow = b * p.HW + (tile - b * p.tiles_per_img) * TC_BM;   // first pixel row of the tile
        for (int kb = 0; kb < p.nb; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = smem_u32(&bar_full[stage]);
            const uint32_t dst = smem_base + (uint32_t)(stage * p.stage_bytes);
            mbar_arrive_expect_tx(full, tx);
            tma_load_2d(dst, &p.tm_x, full, kb * p.kw, prow);
            tma_load_2d(dst + (uint32_t)p.a_bytes, &p.tm_g, full, kb * p.kw, b * MT_N);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, elect_one issue) =====================
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MT_N >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      const uint64_t hi = (uint64_t)(uint32_t)(make_smem_desc(0, p.kw) >> 32) << 32;
      const int nfull = p.kw >> 4, ntail = (p.C - (p.nb - 1) * p.kw + 15) >> 4;
      const uint32_t lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t lo_stage = (uint32_t)p.stage_bytes >> 4, lo_b = (uint32_t)p.a_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = (uint32_t)it & 1u;
        mbar_wait(smem_u32(&bar_tmem_empty[acc]), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)MT_N;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kb = 0; kb < p.nb; ++kb) {
          const int nk16 = (kb == p.nb - 1) ? ntail : nfull;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_lo = lo_base + (uint32_t)stage * lo_stage;
          const uint64_t adesc = hi | a_lo, bdesc = hi | (a_lo + lo_b);
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
    }
  } else if (warp >= 4) {
    // ===================== norm + epilogue warps: thread = pixel row of the tile =====================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const int row_bytes = p.kw * 2;
    const int chunks = row_bytes >> 4;                    // 16-byte chunks per row: 2, 4 or 8
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_img;
      const int pix = (tile - b * p.tiles_per_img) * TC_BM + row;     // pixel within the image
      const bool valid = pix < p.HW;
      float sq = 0.f;
      uint64_t wsq[MT_WMAX / 2];           // windowed: per-orientation sum of squares, packed fp32 pairs
#pragma unroll
      for (int i = 0; i < MT_WMAX / 2; ++i) wsq[i] = 0ull;
      // inverse of the TMA swizzle: 16-byte chunk position within the row ^ swz = logical chunk (128B / 64B / 32B swizzle)
      const int swz = p.kw == 64 ? (row & 7) : (p.kw == 32 ? ((row >> 1) & 3) : ((row >> 2) & 1));
#pragma unroll 1
      for (int kb = 0; kb < p.nb; ++kb) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        const uint8_t* tile_ptr = smem_raw + (smem_base - smem_u32(smem_raw)) + stage * p.stage_bytes + row * row_bytes;
        float s0 = 0.f, s1 = 0.f;
        for (int j = 0; j < chunks; ++j) {
          // UNIT == 0: rotate the chunk order by the row index: lanes of a warp then cover all bank groups (conflict free);
          // the swizzle only permutes chunks inside a row, and a sum of squares does not care about their order.
          // Windowed: every lane reads LOGICAL chunk j (physical j ^ swz -- conflict free by construction of the
          // swizzle), so the membership-table row is the same for the whole warp (broadcast loads).
          const int phys = UNIT == 0 ? ((j + row) & (chunks - 1)) : (j ^ swz);
          const uint4 q = *reinterpret_cast<const uint4*>(tile_ptr + (phys << 4));
          const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
          const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
          const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.z));
          const float2 f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.w));
          if (UNIT == 1) {
            const float* trow = s_wtab + (kb * p.kw + j * 8) * p.wt_pitch;
            const float e[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float x2 = e[t] * e[t];
              s0 += x2;
              window_accumulate<UNIT>(wsq, x2, trow + t * p.wt_pitch, p.wt_pitch);
            }
          } else if (UNIT == 8) {
            float a = f0.x * f0.x, b2 = f0.y * f0.y;
            a = fmaf(f1.x, f1.x, a); b2 = fmaf(f1.y, f1.y, b2);
            a = fmaf(f2.x, f2.x, a); b2 = fmaf(f2.y, f2.y, b2);
            a = fmaf(f3.x, f3.x, a); b2 = fmaf(f3.y, f3.y, b2);
            const float s8 = a + b2;
            s0 += s8;
            window_accumulate<UNIT>(wsq, s8, s_wtab + (kb * chunks + j) * p.wt_pitch, p.wt_pitch);
          } else {
            s0 = fmaf(f0.x, f0.x, s0); s1 = fmaf(f0.y, f0.y, s1);
            s0 = fmaf(f1.x, f1.x, s0); s1 = fmaf(f1.y, f1.y, s1);
            s0 = fmaf(f2.x, f2.x, s0); s1 = fmaf(f2.y, f2.y, s1);
            s0 = fmaf(f3.x, f3.x, s0); s1 = fmaf(f3.y, f3.y, s1);
          }
        }
        sq += s0 + s1;
        if (UNIT == 64) window_accumulate<UNIT>(wsq, s0 + s1, s_wtab + kb * p.wt_pitch, p.wt_pitch);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[stage]));
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // ---- epilogue ----
      const uint32_t acc = (uint32_t)it & 1u;
      mbar_wait(smem_u32(&bar_tmem_full[acc]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * (uint32_t)MT_N, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));   // accumulator is in registers: release it early
      if (valid) {
        const float nrm = sqrtf(sq);
        const float inv = 1.f / fmaxf(nrm, 1e-12f);
        // one reciprocal of the cosine denominator, no epsilon: den == 0 gives inf and 0 * inf = NaN like the
        // reference's 0 / 0 (models.py:196)
        const float gn = __ldg(p.gnorm + b);
        const float rden = 1.f / (nrm * gn);
        const float rgn = 1.f / gn;
        const int64_t gp = (int64_t)b * p.HW + pix;
        float best = -INFINITY;
        bool any_nan = false;
        float sc[MT_N];
        float* sp = p.scores ? p.scores + (int64_t)b * p.n_rolls * p.HW + pix : nullptr;
#pragma unroll
        for (int i = 0; i < MT_N; ++i) {
          if (i < p.n_rolls) {
            if (UNIT != 0 && i < MT_WMAX) {
              // windowed cosine denominator sqrt(sum_window x^2) * ||g||; a zero window gives 0 * inf = NaN like the
              // reference's 0 / 0 (models.py:196)
              float wlo, whi;
              unpack_f32x2(wsq[i >> 1], wlo, whi);
              sc[i] = __uint_as_float(v[i]) * (rsqrtf((i & 1) ? whi : wlo) * rgn);
            } else {
              sc[i] = __uint_as_float(v[i]) * rden;
            }
            if (sp) sp[(int64_t)i * p.HW] = sc[i];
            if ((p.max_mask >> i) & 1u) {
              any_nan |= (sc[i] != sc[i]);
              best = fmaxf(best, sc[i]);
            }
          } else {
            sc[i] = 0.f;
          }
        }
        if (any_nan) best = __int_as_float(0x7fc00000);
        if (p.max_out) p.max_out[gp] = best;
        if (p.inv_norm) p.inv_norm[gp] = inv;
        if (p.scores_cl) {
          __nv_bfloat16* dst = p.scores_cl + gp * p.ld_scores_cl;
#pragma unroll
          for (int i = 0; i < MT_N; i += 8) {
            if (i < p.ld_scores_cl) {
              uint4 pk;
              __nv_bfloat162 q0 = __floats2bfloat162_rn(sc[i + 0], sc[i + 1]), q1 = __floats2bfloat162_rn(sc[i + 2], sc[i + 3]);
              __nv_bfloat162 q2 = __floats2bfloat162_rn(sc[i + 4], sc[i + 5]), q3 = __floats2bfloat162_rn(sc[i + 6], sc[i + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&q0);
              pk.y = *reinterpret_cast<uint32_t*>(&q1);
              pk.z = *reinterpret_cast<uint32_t*>(&q2);
              pk.w = *reinterpret_cast<uint32_t*>(&q3);
              *reinterpret_cast<uint4*>(dst + i) = pk;
            }
          }
          for (int i = MT_N; i < p.ld_scores_cl; ++i) dst[i] = __float2bfloat16_rn(0.f);
        }
        if (p.xhat) {   // bottleneck level only (64 pixels per image): one more read of this pixel's vector
          const __nv_bfloat16* src = p.x + gp * p.C;
          __nv_bfloat16* dst = p.xhat + gp * p.C;
          for (int c = 0; c < p.C; c += 8) {
            uint4 q = *reinterpret_cast<const uint4*>(src + c);
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = __bfloat1622float2(h[j]);
              h[j] = __floats2bfloat162_rn(f.x * inv, f.y * inv);
            }
            *reinterpret_cast<uint4*>(dst + c) = q;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

// Windowed levels: granularity of the window-membership table and its shape.
struct WindowPlan {
  int unit;       // 0 (L == C), 1, 8 or 64 (= one row per K block)
  int rows, pitch;
};

static int gcd_int(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a < 0 ? -a : a;
}

static WindowPlan plan_windows(int C, int L, int offset, const int32_t* shifts_host, int n_rolls) {
  WindowPlan w = {0, 0, 0};
  if (L == C) return w;
  int u = gcd_int(C, L);
  for (int i = 0; i < n_rolls; ++i) u = gcd_int(u, ((offset + shifts_host[i]) % C + C) % C);
  const int kw = tc_block_width(C), nb = (C + kw - 1) / kw;
  if (u % kw == 0) {
    w.unit = 64;
    w.rows = nb;
  } else if (u % 8 == 0) {
    w.unit = 8;
    w.rows = nb * kw / 8;
  } else {
    w.unit = 1;
    w.rows = nb * kw;
  }
  const int r4 = (n_rolls + 3) / 4;
  w.pitch = 4 * (r4 | 1);          // pitch / 4 odd: table rows 16*odd bytes apart never share a bank group
  return w;
}

constexpr int MT_WTAB_MAX_BYTES = 40 * 1024;

bool match_tcgen05_supported(int dtype, int C, int L, int offset, const int32_t* shifts_host, int n_rolls,
                             int ld_scores_cl) {
  if (!(dtype == CCVPE_BF16 && L <= C && C % 8 == 0 && n_rolls <= MT_N && ld_scores_cl % 8 == 0)) return false;
  if (L == C) return true;
  if (n_rolls > MT_WMAX) return false;
  const WindowPlan w = plan_windows(C, L, offset, shifts_host, n_rolls);
  return w.rows * w.pitch * 4 <= MT_WTAB_MAX_BYTES;
}

// scratch layout (bytes from `scratch`): G bf16 [B, 32, C] | gnorm fp32 [B] | window table fp32 [rows, pitch]
int match_tcgen05(const void* x, int B, int HW, int C, const float* g, int L, int offset, const int32_t* shifts_host,
                  int n_rolls, uint32_t max_mask, float* scores, void* scores_cl, int ld_scores_cl, float* max_out,
                  float* inv_norm, void* xhat, float* scratch, cudaStream_t st) {
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  uint8_t* base = reinterpret_cast<uint8_t*>(scratch);
  __nv_bfloat16* G = reinterpret_cast<__nv_bfloat16*>(base);
  float* gnorm = reinterpret_cast<float*>(base + up((int64_t)B * MT_N * C * 2));
  float* wtab = reinterpret_cast<float*>(base + up((int64_t)B * MT_N * C * 2) + up((int64_t)B * 4));
  const WindowPlan wp = plan_windows(C, L, offset, shifts_host, n_rolls);
  ShiftArg sa;   // the roll table travels as a kernel argument (no copy, no sync, graph-capturable)
  for (int i = 0; i < MT_N; ++i) sa.s[i] = i < n_rolls ? shifts_host[i] : 0;
  const int tab_blocks = wp.unit ? (wp.rows * wp.pitch + 255) / 256 : 0;
  build_rolled_descriptor_bf16_kernel<<<dim3(MT_N + 1 + (tab_blocks > 8 ? 8 : tab_blocks), B), 256, 0, st>>>(
      g, L, C, offset, sa, n_rolls, G, gnorm, wp.unit ? wtab : nullptr, wp.rows, wp.pitch,
      wp.unit == 64 ? tc_block_width(C) : wp.unit);
  int rc = check_launch("build_rolled_descriptor_bf16_kernel");
  if (rc != CCVPE_OK) return rc;

  static thread_local MatchTcParams p;
  memset(&p, 0, sizeof(p));
  p.kw = tc_block_width(C);
  p.nb = (C + p.kw - 1) / p.kw;
  p.C = C;
  p.HW = HW;
  p.B = B;
  p.tiles_per_img = (HW + TC_BM - 1) / TC_BM;
  p.total_tiles = B * p.tiles_per_img;
  p.a_bytes = TC_BM * p.kw * 2;
  p.stage_bytes = p.a_bytes + ((MT_N * p.kw * 2 + 1023) / 1024) * 1024;
  // ~52 KB of pipeline per CTA so that four CTAs (16 norm/epilogue warps) share an SM: the epilogue is scalar-ALU and
  // store bound and needs the extra warps to hide its own latency (windowed: two CTAs per SM -- the per-orientation
  // accumulators double the register footprint -- with a deeper ring each)
  int stages = ((wp.unit ? 64 : 52) * 1024) / p.stage_bytes;
  if (stages < 2) stages = 2;
  p.stages = stages > MT_MAX_STAGES ? MT_MAX_STAGES : stages;
  p.n_rolls = n_rolls;
  p.ld_scores_cl = scores_cl ? ld_scores_cl : 0;
  p.wt_rows = wp.rows;
  p.wt_pitch = wp.pitch;
  p.wtab = wtab;
  p.max_mask = max_mask;
  p.gnorm = gnorm;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.scores = scores;
  p.scores_cl = static_cast<__nv_bfloat16*>(scores_cl);
  p.max_out = max_out;
  p.inv_norm = inv_norm;
  p.xhat = static_cast<__nv_bfloat16*>(xhat);
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)B * HW};
    uint64_t str[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {(uint32_t)p.kw, TC_BM};
    if ((rc = encode_map(&p.tm_x, x, 2, dims, str, box, p.kw)) != CCVPE_OK) return rc;
    uint64_t gdims[2] = {(uint64_t)C, (uint64_t)B * MT_N};
    uint32_t gbox[2] = {(uint32_t)p.kw, MT_N};
    if ((rc = encode_map(&p.tm_g, G, 2, gdims, str, gbox, p.kw)) != CCVPE_OK) return rc;
  }
  const int smem = p.stages * p.stage_bytes + 1024 + (wp.unit ? wp.rows * wp.pitch * 4 : 0);
  static thread_local uint64_t attr_set = 0;
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaSuccess;
#define CCVPE_MT_ATTR(U)                                                                                             \
  if (e == cudaSuccess)                                                                                              \
    e = cudaFuncSetAttribute(match_tcgen05_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024)
    CCVPE_MT_ATTR(0); CCVPE_MT_ATTR(1); CCVPE_MT_ATTR(8); CCVPE_MT_ATTR(64);
#undef CCVPE_MT_ATTR
    if (e != cudaSuccess) return fail(CCVPE_ERR_CUDA, "cudaFuncSetAttribute(match): %s", cudaGetErrorString(e));
  }
  const int per_sm = wp.unit ? 2 : 4;
  const int grid = p.total_tiles < per_sm * sm_count() ? p.total_tiles : per_sm * sm_count();
  switch (wp.unit) {
    case 0: match_tcgen05_kernel<0><<<grid, MT_THREADS, smem, st>>>(p); break;
    case 1: match_tcgen05_kernel<1><<<grid, MT_THREADS, smem, st>>>(p); break;
    case 8: match_tcgen05_kernel<8><<<grid, MT_THREADS, smem, st>>>(p); break;
    default: match_tcgen05_kernel<64><<<grid, MT_THREADS, smem, st>>>(p); break;
  }
  return check_launch("match_tcgen05_kernel");
}

}  // namespace ccvpe
