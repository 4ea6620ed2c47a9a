// a4/a5/a6 -- rolled cosine matching of one decoder level (reference models.py:186-202 and its 23 sibling loops).
//
// The reference evaluates, per orientation i, roll -> slice -> norm -> mul with a tiled ground map -> sum -> div -> cat
// (7 full passes over the aerial map per orientation, x20).  Here the orientation loop is turned inside out: the
// rolled, zero-padded ground descriptor is a banded circulant matrix  G[b, i, c] = g[b, (c - offset - shift_i) mod C]
// (0 where that index >= L), so   dot[b, p, i] = sum_c x[b, p, c] * G[b, i, c]   is ONE pass over x with R accumulators
// per pixel; the window norms use the 0/1 band indicator the same way, and the full-C sum of squares that
// F.normalize (models.py:205) needs falls out of the same pass.
//
// This file is the CUDA-core (fp32 FMA) backend: exact fp32 arithmetic for the parity path, any dtype in.
// HBM roofline: reads x once (C*HW*elem per pair), writes R*HW fp32 scores + 2*HW floats.
#include "common.cuh"

namespace ccvpe {

constexpr int kMatchTile = 128;   // pixels per CTA (one thread per pixel)
constexpr int kMatchChunk = 32;   // channels staged per step
constexpr int kMaxRolls = 32;

struct RollTable {
  int shift[kMaxRolls];
};

// G[b][i][c] and the band indicator M[i][c]; gnorm[b] = ||g_b||_2 (models.py:189: norm of the tiled map == norm of g).
__global__ void build_rolled_descriptor_kernel(const float* __restrict__ g, int B, int L, int C, int offset,
                                               RollTable rolls, int n_rolls, float* __restrict__ G,
                                               float* __restrict__ M, float* __restrict__ gnorm) {
  int b = blockIdx.y;
  int i = blockIdx.x;
  if (i < n_rolls) {
    int base = ((offset + rolls.shift[i]) % C + C) % C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      int k = c - base;
      if (k < 0) k += C;
      bool in = k < L;
      G[((int64_t)b * n_rolls + i) * C + c] = in ? g[(int64_t)b * L + k] : 0.f;
      if (b == 0 && M) M[(int64_t)i * C + c] = in ? 1.f : 0.f;
    }
  } else if (i == n_rolls) {  // one extra block per batch element: the descriptor norm
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

template <typename T, int RMAX, bool WINDOWED>
__global__ void __launch_bounds__(kMatchTile)
match_level_simt_kernel(const T* __restrict__ x, int HW, int C, const float* __restrict__ G,
                        const float* __restrict__ M, const float* __restrict__ gnorm, int n_rolls, uint32_t max_mask,
                        float* __restrict__ scores, T* __restrict__ scores_cl, int ld_scores_cl,
                        float* __restrict__ max_out, float* __restrict__ inv_norm, T* __restrict__ xhat) {
  __shared__ float xs[kMatchTile][kMatchChunk + 1];
  __shared__ __align__(16) float Gs[RMAX][kMatchChunk];
  __shared__ __align__(16) float Ms[WINDOWED ? RMAX : 1][kMatchChunk];
  __shared__ float invs[kMatchTile];

  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kMatchTile;
  const int t = threadIdx.x;
  const int rows = min(kMatchTile, HW - p0);
  const T* xt = x + ((int64_t)b * HW + p0) * C;
  const float* Gb = G + (int64_t)b * n_rolls * C;

  float acc[RMAX], wsq[WINDOWED ? RMAX : 1];
#pragma unroll
  for (int i = 0; i < RMAX; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (WINDOWED ? RMAX : 1); ++i) wsq[i] = 0.f;
  float sq = 0.f;

  for (int c0 = 0; c0 < C; c0 += kMatchChunk) {
    // stage x[p0:p0+128, c0:c0+32] (8 float4 per row)
#pragma unroll
    for (int j = 0; j < (kMatchTile * kMatchChunk / 4) / kMatchTile; ++j) {
      int v = t + j * kMatchTile;
      int row = v >> 3, q = (v & 7) * 4;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows && c0 + q < C) val = load4(xt + (int64_t)row * C + c0 + q);
      xs[row][q + 0] = val.x;
      xs[row][q + 1] = val.y;
      xs[row][q + 2] = val.z;
      xs[row][q + 3] = val.w;
    }
    for (int v = t; v < RMAX * (kMatchChunk / 4); v += kMatchTile) {
      int i = v >> 3, q = (v & 7) * 4;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), mv = gv;
      if (i < n_rolls && c0 + q < C) {
        gv = *reinterpret_cast<const float4*>(Gb + (int64_t)i * C + c0 + q);
        if (WINDOWED) mv = *reinterpret_cast<const float4*>(M + (int64_t)i * C + c0 + q);
      }
      *reinterpret_cast<float4*>(&Gs[i][q]) = gv;
      if (WINDOWED) *reinterpret_cast<float4*>(&Ms[i][q]) = mv;
    }
    __syncthreads();
#pragma unroll 2
    for (int cc = 0; cc < kMatchChunk; cc += 4) {
      float x0 = xs[t][cc], x1 = xs[t][cc + 1], x2 = xs[t][cc + 2], x3 = xs[t][cc + 3];
      float s0 = x0 * x0, s1 = x1 * x1, s2 = x2 * x2, s3 = x3 * x3;
      sq += (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int i = 0; i < RMAX; ++i) {
        float4 gv = *reinterpret_cast<const float4*>(&Gs[i][cc]);
        acc[i] = fmaf(x0, gv.x, acc[i]);
        acc[i] = fmaf(x1, gv.y, acc[i]);
        acc[i] = fmaf(x2, gv.z, acc[i]);
        acc[i] = fmaf(x3, gv.w, acc[i]);
        if (WINDOWED) {
          float4 mv = *reinterpret_cast<const float4*>(&Ms[i][cc]);
          wsq[i] = fmaf(s0, mv.x, wsq[i]);
          wsq[i] = fmaf(s1, mv.y, wsq[i]);
          wsq[i] = fmaf(s2, mv.z, wsq[i]);
          wsq[i] = fmaf(s3, mv.w, wsq[i]);
        }
      }
    }
    __syncthreads();
  }

  const float gn = gnorm[b];
  const float inv = 1.f / fmaxf(sqrtf(sq), 1e-12f);
  invs[t] = inv;
  if (t < rows) {
    const int64_t p = (int64_t)b * HW + p0 + t;
    float best = -INFINITY;
    bool any_nan = false;
#pragma unroll
    for (int i = 0; i < RMAX; ++i) {
      if (i < n_rolls) {
        float wn = sqrtf(WINDOWED ? wsq[i] : sq);
        float s = acc[i] / (wn * gn);  // no epsilon, like models.py:196 (0/0 -> NaN)
        acc[i] = s;
        if (scores) scores[((int64_t)b * n_rolls + i) * HW + p0 + t] = s;
        if ((max_mask >> i) & 1u) {
          any_nan |= (s != s);
          best = fmaxf(best, s);
        }
      }
    }
    if (any_nan) best = __int_as_float(0x7fc00000);  // torch.max propagates NaN
    if (max_out) max_out[p] = best;
    if (inv_norm) inv_norm[p] = inv;
    if (scores_cl) {
      T* dst = scores_cl + p * ld_scores_cl;
#pragma unroll
      for (int i = 0; i < RMAX; ++i)
        if (i < ld_scores_cl) dst[i] = from_float<T>(i < n_rolls ? acc[i] : 0.f);
      for (int i = RMAX; i < ld_scores_cl; ++i) dst[i] = from_float<T>(0.f);
    }
  }
  if (xhat) {  // normalised map (only requested at the bottleneck level; x tile is L2 resident)
    __syncthreads();
    T* xo = xhat + ((int64_t)b * HW + p0) * C;
    int64_t n4 = (int64_t)rows * C / 4;
    for (int64_t v = t; v < n4; v += kMatchTile) {
      int row = (int)((v * 4) / C);
      float4 val = load4(xt + v * 4);
      float s = invs[row];
      val.x *= s; val.y *= s; val.z *= s; val.w *= s;
      store4(xo + v * 4, val);
    }
  }
}

template <typename T>
int launch_match_simt(const T* x, int B, int HW, int C, const float* G, const float* M, const float* gnorm,
                      int n_rolls, uint32_t max_mask, bool windowed, float* scores, T* scores_cl, int ld_scores_cl,
                      float* max_out, float* inv_norm, T* xhat, cudaStream_t st) {
  dim3 grid((HW + kMatchTile - 1) / kMatchTile, B);
#define CCVPE_MATCH_CASE(R, W)                                                                                    \
  match_level_simt_kernel<T, R, W><<<grid, kMatchTile, 0, st>>>(x, HW, C, G, M, gnorm, n_rolls, max_mask, scores, \
                                                                scores_cl, ld_scores_cl, max_out, inv_norm, xhat)
  if (n_rolls <= 16) {
    if (windowed) CCVPE_MATCH_CASE(16, true); else CCVPE_MATCH_CASE(16, false);
  } else {
    if (windowed) CCVPE_MATCH_CASE(24, true); else CCVPE_MATCH_CASE(24, false);
  }
#undef CCVPE_MATCH_CASE
  return check_launch("match_level_simt_kernel");
}

// shared with the backward kernel (train_ops.cu): G [B, R, C], M [R, C] (NULL when the windows cover all channels), gnorm [B]
int build_rolled_descriptor_f32(const float* g, int B, int L, int C, int offset, const int32_t* shifts_host, int n_rolls,
                                float* G, float* M, float* gnorm, cudaStream_t st) {
  RollTable rt;
  for (int i = 0; i < kMaxRolls; ++i) rt.shift[i] = i < n_rolls ? shifts_host[i] : 0;
  build_rolled_descriptor_kernel<<<dim3(n_rolls + 1, B), 256, 0, st>>>(g, B, L, C, offset, rt, n_rolls, G, M, gnorm);
  return check_launch("build_rolled_descriptor_kernel");
}

bool match_tcgen05_supported(int dtype, int C, int L, int offset, const int32_t* shifts_host, int n_rolls,
                             int ld_scores_cl);   // match_tcgen05.cu
int match_tcgen05(const void* x, int B, int HW, int C, const float* g, int L, int offset, const int32_t* shifts_host,
                  int n_rolls, uint32_t max_mask, float* scores, void* scores_cl, int ld_scores_cl, float* max_out,
                  float* inv_norm, void* xhat, float* scratch, cudaStream_t st);

}  // namespace ccvpe

extern "C" int64_t ccvpe_match_scratch_elems(int B, int C, int n_rolls) {
  // SIMT backend: G fp32 [B, R, C] + M [R, C] + gnorm [B], each region rounded up to 64 floats (256 B);
  // tcgen05 backend: G bf16 [B, 32, C] (= B*16*C floats) + gnorm [B] + the window-membership table (<= 40 KB).
  // Sized for the larger of the two.
  auto up = [](int64_t v) { return (v + 63) / 64 * 64; };
  const int r = n_rolls > 17 ? n_rolls : 17;
  return up((int64_t)B * r * C) + up((int64_t)n_rolls * C) + up(B) + 256 + 10240 + 128;
}

extern "C" int ccvpe_match_plan(int dtype, int C, int L, int offset, const int32_t* shifts_host, int n_rolls,
                                int ld_scores_cl, int backend) {
  using namespace ccvpe;
  CCVPE_REQUIRE(shifts_host && C > 0 && L > 0 && L <= C && n_rolls > 0 && n_rolls <= 24, "ccvpe_match_plan: bad argument");
  if (backend == CCVPE_BACKEND_SIMT) return 0;
  return match_tcgen05_supported(dtype, C, L, offset, shifts_host, n_rolls, ld_scores_cl) ? 1 : 0;
}

extern "C" int ccvpe_match_level(const void* x, int dtype, int B, int HW, int C, const float* g, int L, int offset,
                                 const int32_t* shifts_host, int n_rolls, uint32_t max_mask, float* scores,
                                 void* scores_cl, int ld_scores_cl, float* max_out, float* inv_norm, void* xhat,
                                 float* scratch, int backend, void* stream) {
  using namespace ccvpe;
  CCVPE_REQUIRE(x && g && shifts_host && scratch, "ccvpe_match_level: null pointer");
  CCVPE_REQUIRE(B > 0 && HW > 0 && C > 0 && L > 0 && L <= C, "ccvpe_match_level: bad shape B=%d HW=%d C=%d L=%d", B, HW, C,
                L);
  CCVPE_REQUIRE(C % 8 == 0, "ccvpe_match_level: C=%d must be a multiple of 8", C);
  CCVPE_REQUIRE(n_rolls > 0 && n_rolls <= 24, "ccvpe_match_level: n_rolls=%d out of range [1, 24]", n_rolls);
  CCVPE_REQUIRE(dtype == CCVPE_F32 || dtype == CCVPE_BF16, "ccvpe_match_level: bad dtype %d", dtype);
  CCVPE_REQUIRE(aligned16(x) && aligned16(scratch), "ccvpe_match_level: x/scratch must be 16-byte aligned");
  CCVPE_REQUIRE(!scores_cl || ld_scores_cl >= n_rolls, "ccvpe_match_level: ld_scores_cl=%d < n_rolls", ld_scores_cl);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = match_tcgen05_supported(dtype, C, L, offset, shifts_host, n_rolls, scores_cl ? ld_scores_cl : 0);
  if (backend == CCVPE_BACKEND_TCGEN05 && !tc_ok)
    return fail(CCVPE_ERR_UNSUPPORTED, "ccvpe_match_level: tcgen05 backend needs bf16, n_rolls <= 24 and a window table <= 40 KB");
  if ((backend == CCVPE_BACKEND_TCGEN05 || backend == CCVPE_BACKEND_AUTO) && tc_ok)
    return match_tcgen05(x, B, HW, C, g, L, offset, shifts_host, n_rolls, max_mask, scores, scores_cl, ld_scores_cl,
                         max_out, inv_norm, xhat, scratch, st);
  auto up = [](int64_t v) { return (v + 63) / 64 * 64; };
  float* G = scratch;
  float* M = G + up((int64_t)B * n_rolls * C);
  float* gnorm = M + up((int64_t)n_rolls * C);
  const bool windowed = L < C;
  {
    const int rc = build_rolled_descriptor_f32(g, B, L, C, offset, shifts_host, n_rolls, G, windowed ? M : nullptr, gnorm, st);
    if (rc != CCVPE_OK) return rc;
  }
  if (dtype == CCVPE_F32)
    return launch_match_simt<float>((const float*)x, B, HW, C, G, M, gnorm, n_rolls, max_mask, windowed, scores,
                                    (float*)scores_cl, ld_scores_cl, max_out, inv_norm, (float*)xhat, st);
  return launch_match_simt<__nv_bfloat16>((const __nv_bfloat16*)x, B, HW, C, G, M, gnorm, n_rolls, max_mask, windowed,
                                          scores, (__nv_bfloat16*)scores_cl, ld_scores_cl, max_out, inv_norm,
                                          (__nv_bfloat16*)xhat, st);
}
