// tcgen05 backend of the generic implicit GEMM: the bf16 throughput path of a3 (cell descriptors), a8 (transposed
// convs) and a9 (3x3 convs).  sm_100a only.
//
//   * A (activations, channels-last bf16) is never im2col'ed: for every (tap, 64-channel block) ONE TMA box load of
//     [128 pixels x 64 channels] at spatially shifted coordinates lands a K-major, 128B-swizzled UMMA operand tile in
//     shared memory; the conv halo (pad 1) is TMA out-of-bounds zero fill, channel tails are zero filled too, and the
//     skip concat is a second tensor map walked by the same K loop.
//   * B (weights, [N][taps][K] bf16, K contiguous) tiles arrive through a 2-D tensor map the same way.
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n<=256, K=16) into TMEM; accumulators
//     are double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (tcgen05.ld -> bias / row-scale /
//     rank-1 / ReLU -> bf16 or fp32 stores incl. the transposed-conv pixel shuffle).  Persistent: grid = min(tiles, SMs).
//   * every mbarrier wait is bounded (clock64 watchdog -> __trap) so a protocol bug fails the launch instead of
//     hanging the GPU.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace ccvpe {

constexpr int TC_BM = 128;          // pixels per tile (UMMA M)
constexpr int TC_BK = 64;           // bf16 channels per K block (= one 128-byte swizzle row)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_MAX_N = 256;
constexpr int TC_THREADS = 256;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_BUDGET = 200 * 1024;

struct TcParams {
  CUtensorMap tm_a0, tm_a1, tm_b0, tm_b1;
  int nb0, nb1, c0, c1, kw0, kw1, kpad0, kpad1, taps, n_tiles_n, total_tiles, block_n, stages;
  int tiles[4], box[4];
  int tap_off[9][4];
  int out_stride[4], extent[4];
  int N, HWo, Wout, Hout;
  const float* bias;
  const float* row_scale;
  const float* row_r1;
  const float* r1_w;
  int relu, out_mode, out_f32, ldo;
  void* out;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a pipeline protocol bug must not hang the device
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled shared-memory matrix descriptor.  A K block of kw bf16 channels is one swizzle row of 2*kw bytes
// (kw = 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B); 8 rows form a group, groups are 16*kw bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int kw) {
  const uint64_t layout = kw == 64 ? 2 : (kw == 32 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), 16 B
  d |= (uint64_t)((16 * kw) >> 4) << 32;     // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

// ---- kernel ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) igemm_tcgen05_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tmem_full[2];
  __shared__ __align__(8) uint64_t bar_tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[2][TC_MAX_N];
  __shared__ __align__(16) float s_r1w[2][TC_MAX_N];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = TC_A_BYTES + p.block_n * TC_BK * 2;   // sized for the widest (64-channel) K block
  const int nkb = p.taps * (p.nb0 + p.nb1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tmem_full[a]), 1);
      mbar_init(smem_u32(&bar_tmem_empty[a]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tm_a0);
    prefetch_tmap(&p.tm_b0);
    if (p.nb1) {
      prefetch_tmap(&p.tm_a1);
      prefetch_tmap(&p.tm_b1);
    }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        int base[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          base[d] = (mt % p.tiles[d]) * p.box[d];
          mt /= p.tiles[d];
        }
        const int n0 = nt * p.block_n;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int c1 = base[0] + p.tap_off[tap][0], c2 = base[1] + p.tap_off[tap][1];
          const int c3 = base[2] + p.tap_off[tap][2], c4 = base[3] + p.tap_off[tap][3];
          for (int src = 0; src < 2; ++src) {
            const int nb = src ? p.nb1 : p.nb0;
            const int kw = src ? p.kw1 : p.kw0;
            const CUtensorMap* tm = src ? &p.tm_a1 : &p.tm_a0;
            const CUtensorMap* tmb = src ? &p.tm_b1 : &p.tm_b0;
            const int kbase = tap * (p.kpad0 + p.kpad1) + (src ? p.kpad0 : 0);   // column of this source in W[n][.]
            const uint32_t tx = (uint32_t)((TC_BM + p.block_n) * kw * 2);
            for (int cb = 0; cb < nb; ++cb) {
              mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&bar_full[stage]);
              const uint32_t a_dst = smem_base + stage * stage_bytes;
              mbar_arrive_expect_tx(full, tx);
              tma_load_5d(a_dst, tm, full, cb * kw, c1, c2, c3, c4);
              tma_load_2d(a_dst + TC_A_BYTES, tmb, full, kbase + cb * kw, n0);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=block_n
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(smem_u32(&bar_tmem_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_MAX_N);
        for (int kb = 0; kb < nkb; ++kb) {
          // number of 16-channel MMAs that carry data in this K block (the rest is TMA zero fill)
          const int within = kb % (p.nb0 + p.nb1);
          const bool s1 = within >= p.nb0;
          const int kw = s1 ? p.kw1 : p.kw0;
          const int cvalid = s1 ? p.c1 - (within - p.nb0) * kw : p.c0 - within * kw;
          const int nk16 = cvalid >= kw ? kw / 16 : (cvalid + 15) >> 4;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          const uint64_t adesc = make_smem_desc(a_addr, kw);
          const uint64_t bdesc = make_smem_desc(a_addr + TC_A_BYTES, kw);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the 16-byte address field
            if (k < nk16) umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));   // frees the smem slot once these MMAs have read it
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&bar_tmem_full[acc]));    // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (4 warps, warp w owns TMEM lanes 32*(w%4) .. +31) =====================
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;                 // 0..127 within the epilogue group
    const bool fixed_n = (p.n_tiles_n == 1);
    const bool has_r1 = (p.row_r1 != nullptr), has_rs = (p.row_scale != nullptr);

    // per-column epilogue vectors live in shared memory (one copy per accumulator stage)
    auto stage_vectors = [&](int buf, int n0) {
      for (int c = et; c < p.block_n; c += 128) {
        const int n = n0 + c;
        s_bias[buf][c] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
        s_r1w[buf][c] = (has_r1 && n < p.N) ? __ldg(p.r1_w + n) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    // row -> output pixel of a tile
    auto locate = [&](int tile, int& m_glob, bool& valid) {
      int mt = tile / p.n_tiles_n;
      int r = row;
      m_glob = 0;
      valid = true;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int coord = (mt % p.tiles[d]) * p.box[d] + (r % p.box[d]);
        mt /= p.tiles[d];
        r /= p.box[d];
        valid = valid && (coord < p.extent[d]);
        m_glob += coord * p.out_stride[d];
      }
    };

    if (fixed_n) stage_vectors(0, 0);
    int m_glob, m_next = 0;
    bool valid, valid_next = false;
    float rs = 1.f, r1 = 0.f, rs_next = 1.f, r1_next = 0.f;
    if ((int)blockIdx.x < p.total_tiles) {
      locate(blockIdx.x, m_next, valid_next);
      if (valid_next) {
        if (has_rs) rs_next = __ldg(p.row_scale + m_next);
        if (has_r1) r1_next = __ldg(p.row_r1 + m_next);
      }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      m_glob = m_next; valid = valid_next; rs = rs_next; r1 = r1_next;
      {  // prefetch the per-row scalars of the next tile: their latency hides behind this tile's epilogue
        const int nxt = tile + gridDim.x;
        rs_next = 1.f; r1_next = 0.f; valid_next = false;
        if (nxt < p.total_tiles) {
          locate(nxt, m_next, valid_next);
          if (valid_next) {
            if (has_rs) rs_next = __ldg(p.row_scale + m_next);
            if (has_r1) r1_next = __ldg(p.row_r1 + m_next);
          }
        }
      }
      const int nt = tile % p.n_tiles_n;
      const int n0 = nt * p.block_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int vb = fixed_n ? 0 : acc;
      if (!fixed_n) stage_vectors(acc, n0);

      // output addressing (hoisted out of the column loop)
      int64_t base_px[4];
      int64_t plane0 = 0;
      if (p.out_mode == 0) {
        base_px[0] = (int64_t)m_glob * p.ldo;
      } else {
        const int b_img = m_glob / p.HWo;
        const int hw = m_glob - b_img * p.HWo;
        if (p.out_mode == 2) {
          plane0 = (int64_t)b_img * p.N * p.HWo + hw;
        } else {
          const int h = hw / p.Wout, w = hw - h * p.Wout;
#pragma unroll
          for (int ij = 0; ij < 4; ++ij)
            base_px[ij] = (((int64_t)b_img * 2 * p.Hout + 2 * h + (ij >> 1)) * (2 * p.Wout) + 2 * w + (ij & 1)) * p.ldo;
        }
      }
      const int cout = p.out_mode == 1 ? (p.N >> 2) : p.N;
      int ij = 0, co = n0;                       // running (quadrant, channel) of column n0 + c
      if (p.out_mode == 1) {
        ij = n0 / cout;
        co = n0 - ij * cout;
      }

      mbar_wait(smem_u32(&bar_tmem_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TC_MAX_N);

      auto process = [&](const uint32_t (&v)[32], int c0) {
        if (!valid) return;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int cl = c0 + g8 * 8;             // column within the tile
          const int n = n0 + cl;
          if (n >= p.N || cl >= p.block_n) break;
          const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[vb][cl]);
          const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[vb][cl + 4]);
          float y[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          if (has_r1) {
            const float4 w0 = *reinterpret_cast<const float4*>(&s_r1w[vb][cl]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_r1w[vb][cl + 4]);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaf(r1, wv[j], y[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            y[j] = fmaf(__uint_as_float(v[g8 * 8 + j]), rs, y[j]);
            if (p.relu) y[j] = fmaxf(y[j], 0.f);
          }
          if (p.out_mode == 2) {
            float* o = static_cast<float*>(p.out) + plane0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (n + j < p.N) o[(int64_t)(n + j) * p.HWo] = y[j];
            continue;
          }
          const int64_t off = (p.out_mode == 0 ? base_px[0] : base_px[ij]) + co;
          const bool full8 = (n + 8 <= p.N);
          if (p.out_f32) {
            float* o = static_cast<float*>(p.out) + off;
            if (full8 && (p.ldo & 3) == 0) {
              *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n + j < p.N) o[j] = y[j];
            }
          } else {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + off;
            if (full8 && (p.ldo & 7) == 0) {
              __nv_bfloat162 q0 = __floats2bfloat162_rn(y[0], y[1]), q1 = __floats2bfloat162_rn(y[2], y[3]);
              __nv_bfloat162 q2 = __floats2bfloat162_rn(y[4], y[5]), q3 = __floats2bfloat162_rn(y[6], y[7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&q0);
              pk.y = *reinterpret_cast<uint32_t*>(&q1);
              pk.z = *reinterpret_cast<uint32_t*>(&q2);
              pk.w = *reinterpret_cast<uint32_t*>(&q3);
              *reinterpret_cast<uint4*>(o) = pk;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n + j < p.N) o[j] = __float2bfloat16_rn(y[j]);
            }
          }
          co += 8;
          if (p.out_mode == 1 && co >= cout) {
            co -= cout;
            ++ij;
          }
        }
      };

      // TMEM -> registers, double buffered: the load of chunk c+1 is in flight while chunk c is processed
      uint32_t va[32], vb2[32];
      tmem_ld32(taddr, va);
      for (int c0 = 0; c0 < p.block_n; c0 += 64) {
        tmem_ld_wait();
        const bool more1 = (c0 + 32 < p.block_n);
        if (more1) tmem_ld32(taddr + (uint32_t)(c0 + 32), vb2);
        process(va, c0);
        if (more1) {
          tmem_ld_wait();
          if (c0 + 64 < p.block_n) tmem_ld32(taddr + (uint32_t)(c0 + 64), va);
          process(vb2, c0 + 32);
        }
      }
      // release the accumulator stage back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &sym, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

static int encode_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int kw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(CCVPE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CCVPE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,..] box=[%u,%u,%u,..]",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
  return CCVPE_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// K-block width (bf16 channels) used for a source with c channels; one block is one swizzle row of 2*width bytes.
// Narrow sources get narrow blocks so TMA neither over-fetches nor zero-fills most of the tile (include/ccvpe_b200.h
// documents the same rule for the w_nk weight layout).
static int tc_block_width(int c) { return c <= 16 ? 16 : (c < 96 ? 32 : 64); }

struct TcGeometry {
  bool cell;        // k2 s2 cell-descriptor gather
  int tw, th, tb;   // conv tile: tw x th pixels x tb images = 128 rows
};

static bool tc_geometry(const ccvpe_igemm_desc& d, TcGeometry* g) {
  if (d.stride == 2 && d.kh == 2 && d.kw == 2 && d.pad == 0) {
    if (d.Hin != 16 || d.Win != 16 || d.Hout != 8 || d.Wout != 8 || d.c1 != 0) return false;
    g->cell = true;
    g->tw = g->th = g->tb = 0;
    return true;
  }
  if (d.stride != 1 || d.kh != d.kw || !(d.kh == 1 || d.kh == 3) || d.pad != (d.kh - 1) / 2) return false;
  if (d.Hin != d.Hout || d.Win != d.Wout) return false;
  if (!is_pow2(d.Wout) || !is_pow2(d.Hout)) return false;
  int tw = d.Wout < TC_BM ? d.Wout : TC_BM;
  int th = TC_BM / tw;
  if (th > d.Hout) th = d.Hout;
  int tb = TC_BM / (tw * th);
  if (tw * th * tb != TC_BM || tb > 256) return false;
  g->cell = false;
  g->tw = tw;
  g->th = th;
  g->tb = tb;
  return true;
}

bool igemm_tcgen05_supported(const ccvpe_igemm_desc& d) {
  TcGeometry g;
  if (d.dtype != CCVPE_BF16 || !d.w_nk) return false;
  if (!tc_geometry(d, &g)) return false;
  if (d.out_mode == 1 && (d.kh != 1)) return false;
  return true;
}

int igemm_tcgen05(const ccvpe_igemm_desc& d, cudaStream_t st) {
  TcGeometry g;
  if (d.dtype != CCVPE_BF16) return fail(CCVPE_ERR_UNSUPPORTED, "ccvpe_igemm(tcgen05): bf16 operands only");
  if (!d.w_nk) return fail(CCVPE_ERR_BAD_ARGUMENT, "ccvpe_igemm(tcgen05): w_nk is null");
  if (!tc_geometry(d, &g)) return fail(CCVPE_ERR_UNSUPPORTED, "ccvpe_igemm(tcgen05): unsupported geometry");
  if (!aligned16(d.w_nk)) return fail(CCVPE_ERR_BAD_ARGUMENT, "ccvpe_igemm(tcgen05): w_nk must be 16-byte aligned");

  static thread_local TcParams p;   // > 1 KB: keep it off the stack; it is copied at launch
  memset(&p, 0, sizeof(p));
  p.kw0 = tc_block_width(d.c0);
  p.kw1 = d.c1 ? tc_block_width(d.c1) : 64;
  p.nb0 = (d.c0 + p.kw0 - 1) / p.kw0;
  p.nb1 = (d.c1 + p.kw1 - 1) / p.kw1;
  p.kpad0 = p.nb0 * p.kw0;
  p.kpad1 = p.nb1 * p.kw1;
  p.c0 = d.c0;
  p.c1 = d.c1;
  p.taps = d.kh * d.kw;
  const int n_tiles_n = (d.N + TC_MAX_N - 1) / TC_MAX_N;
  int block_n = ((d.N + n_tiles_n - 1) / n_tiles_n + 15) / 16 * 16;
  p.n_tiles_n = n_tiles_n;
  p.block_n = block_n;
  const int stage_bytes = TC_A_BYTES + block_n * TC_BK * 2;
  int stages = TC_SMEM_BUDGET / stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  p.stages = stages;
  const int64_t ktot = (int64_t)p.taps * (p.kpad0 + p.kpad1);

  int rc;
  const uint64_t esz = 2;
  if (g.cell) {
    // [c][dw=2][j=8][dh=2][ib=8*B]
    uint64_t dims[5] = {(uint64_t)d.c0, 2, 8, 2, (uint64_t)8 * d.B};
    uint64_t str[4] = {(uint64_t)d.ld0 * esz, 2ull * d.ld0 * esz, (uint64_t)d.Win * d.ld0 * esz,
                       2ull * d.Win * d.ld0 * esz};
    uint32_t box[5] = {(uint32_t)p.kw0, 1, 8, 1, 16};
    if ((rc = encode_map(&p.tm_a0, d.a0, 5, dims, str, box, p.kw0)) != CCVPE_OK) return rc;
    p.tiles[0] = 1; p.tiles[1] = 1; p.tiles[2] = 1; p.tiles[3] = (8 * d.B + 15) / 16;
    p.box[0] = 1; p.box[1] = 8; p.box[2] = 1; p.box[3] = 16;
    for (int t = 0; t < 4; ++t) {
      p.tap_off[t][0] = t & 1;   // dw
      p.tap_off[t][1] = 0;
      p.tap_off[t][2] = t >> 1;  // dh
      p.tap_off[t][3] = 0;
    }
    p.out_stride[0] = 0; p.out_stride[1] = 1; p.out_stride[2] = 0; p.out_stride[3] = 8;
    p.extent[0] = 1 << 30; p.extent[1] = 8; p.extent[2] = 1 << 30; p.extent[3] = 8 * d.B;
  } else {
    for (int s = 0; s < 2; ++s) {
      const void* base = s ? d.a1 : d.a0;
      const int c = s ? d.c1 : d.c0, ld = s ? d.ld1 : d.ld0;
      if (!c) continue;
      uint64_t dims[5] = {(uint64_t)c, (uint64_t)d.Win, (uint64_t)d.Hin, (uint64_t)d.B, 1};
      uint64_t str[4] = {(uint64_t)ld * esz, (uint64_t)d.Win * ld * esz, (uint64_t)d.Hin * d.Win * ld * esz,
                         (uint64_t)d.B * d.Hin * d.Win * ld * esz};
      const int kw = s ? p.kw1 : p.kw0;
      uint32_t box[5] = {(uint32_t)kw, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.tb, 1};
      if ((rc = encode_map(s ? &p.tm_a1 : &p.tm_a0, base, 5, dims, str, box, kw)) != CCVPE_OK) return rc;
    }
    p.tiles[0] = d.Wout / g.tw; p.tiles[1] = d.Hout / g.th; p.tiles[2] = (d.B + g.tb - 1) / g.tb; p.tiles[3] = 1;
    p.box[0] = g.tw; p.box[1] = g.th; p.box[2] = g.tb; p.box[3] = 1;
    for (int t = 0; t < p.taps; ++t) {
      p.tap_off[t][0] = (t % d.kw) - d.pad;
      p.tap_off[t][1] = (t / d.kw) - d.pad;
      p.tap_off[t][2] = 0;
      p.tap_off[t][3] = 0;
    }
    p.out_stride[0] = 1; p.out_stride[1] = d.Wout; p.out_stride[2] = d.Hout * d.Wout; p.out_stride[3] = 0;
    p.extent[0] = d.Wout; p.extent[1] = d.Hout; p.extent[2] = d.B; p.extent[3] = 1;
  }
  for (int s = 0; s < (d.c1 ? 2 : 1); ++s) {
    const int kw = s ? p.kw1 : p.kw0;
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d.N};
    uint64_t str[1] = {(uint64_t)ktot * esz};
    uint32_t box[2] = {(uint32_t)kw, (uint32_t)block_n};
    if ((rc = encode_map(s ? &p.tm_b1 : &p.tm_b0, d.w_nk, 2, dims, str, box, kw)) != CCVPE_OK) return rc;
  }
  const int m_tiles = p.tiles[0] * p.tiles[1] * p.tiles[2] * p.tiles[3];
  p.total_tiles = m_tiles * n_tiles_n;
  p.N = d.N;
  p.HWo = d.Hout * d.Wout;
  p.Wout = d.Wout;
  p.Hout = d.Hout;
  p.bias = d.bias;
  p.row_scale = d.row_scale;
  p.row_r1 = d.row_r1;
  p.r1_w = d.r1_w;
  p.relu = d.relu;
  p.out_mode = d.out_mode;
  p.out_f32 = (d.out_dtype == CCVPE_F32);
  p.ldo = d.ldo;
  p.out = d.out;

  const int smem = stages * stage_bytes + 1024;
  static thread_local int smem_attr_set = 0;
  if (smem_attr_set < smem) {
    cudaError_t e = cudaFuncSetAttribute(igemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8192);
    if (e != cudaSuccess) return fail(CCVPE_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_attr_set = 227 * 1024;
  }
  int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  igemm_tcgen05_kernel<<<grid, TC_THREADS, smem, st>>>(p);
  return check_launch("igemm_tcgen05_kernel");
}

}  // namespace ccvpe
