// a1 -- ground descriptor heads (reference models.py:22-31, 57-97, 152-157).
//
//   g[b, w*c + ch] = sum_h v[h] * (sum_k W[ch,k] F[b,k,h,w] + b1[ch]) + b2
//                  = sum_k W[ch,k] * S[b,w,k] + b1[ch] * sum_h v[h] + b2,     S[b,w,k] = sum_h v[h] F[b,k,h,w]
//
// Two small kernels: the height reduction (reads the 1 MB/pair feature volume once, any strides) and the projection over
// K = 1280 (single head: a warp-per-output dot product; all heads: shared-memory tiled GEMMs).  HBM-bound, ~1 MB/pair.
#include <cstring>

#include "common.cuh"

namespace ccvpe {

template <typename T>
__global__ void grd_height_reduce_kernel(const T* __restrict__ feat, int B, int K, int H, int W, int64_t sb,
                                         int64_t sk, int64_t sh, int64_t sw, const float* __restrict__ v,
                                         float* __restrict__ S, bool k_fastest) {
  int64_t total = (int64_t)B * K * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b, k, w;
    if (k_fastest) {
      k = (int)(i % K);
      w = (int)((i / K) % W);
      b = (int)(i / ((int64_t)K * W));
    } else {
      w = (int)(i % W);
      k = (int)((i / W) % K);
      b = (int)(i / ((int64_t)K * W));
    }
    const T* p = feat + b * sb + k * sk + w * sw;
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc = fmaf(__ldg(v + h), to_float(p[h * sh]), acc);
    S[((int64_t)b * W + w) * K + k] = acc;
  }
}

__global__ void grd_project_kernel(const float* __restrict__ S, const float* __restrict__ w1,
                                   const float* __restrict__ b1, const float* __restrict__ v,
                                   const float* __restrict__ b2, int BW, int K, int H, int c,
                                   float* __restrict__ out) {
  int warps_per_block = blockDim.x >> 5;
  int64_t o = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (o >= (int64_t)BW * c) return;
  int ch = (int)(o % c);
  int64_t bw = o / c;
  const float* s = S + bw * K;
  const float* w = w1 + (int64_t)ch * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(w + k), s[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float vs = 0.f;
    for (int h = 0; h < H; ++h) vs += v[h];
    out[o] = acc + b1[ch] * vs + b2[0];  // out index = (b*W + w)*c + ch  ==  b*(W*c) + w*c + ch
  }
}

// ---- all six heads in two launches -----------------------------------------------------------------------------------
constexpr int kMaxHeads = 6;
struct HeadTable {
  const float* w1[kMaxHeads];
  const float* b1[kMaxHeads];
  const float* w2[kMaxHeads];
  const float* b2[kMaxHeads];
  float* out[kMaxHeads];
  int c[kMaxHeads];
  int c_prefix[kMaxHeads + 1];
  int n;
};

// S[l][b][w][k] = sum_h v_l[h] F[b,k,h,w]: the feature volume is read ONCE for all heads
template <typename T>
__global__ void grd_height_reduce_all_kernel(const T* __restrict__ feat, int B, int K, int H, int W, int64_t sb, int64_t sk,
                                             int64_t sh, int64_t sw, HeadTable t, float* __restrict__ S, bool k_fastest) {
  const int64_t total = (int64_t)B * K * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b, k, w;
    if (k_fastest) {
      k = (int)(i % K);
      w = (int)((i / K) % W);
      b = (int)(i / ((int64_t)K * W));
    } else {
      w = (int)(i % W);
      k = (int)((i / W) % K);
      b = (int)(i / ((int64_t)K * W));
    }
    const T* p = feat + b * sb + k * sk + w * sw;
    float acc[kMaxHeads];
#pragma unroll
    for (int l = 0; l < kMaxHeads; ++l) acc[l] = 0.f;
    for (int h = 0; h < H; ++h) {
      const float f = to_float(p[h * sh]);
#pragma unroll
      for (int l = 0; l < kMaxHeads; ++l)
        if (l < t.n) acc[l] = fmaf(__ldg(t.w2[l] + h), f, acc[l]);
    }
#pragma unroll
    for (int l = 0; l < kMaxHeads; ++l)
      if (l < t.n) S[(((int64_t)l * B + b) * W + w) * K + k] = acc[l];
  }
}

// Shared-memory tiled projection (six small fp32 GEMMs [B*W x K] . [K x c_l]).  The first version, one warp per output
// element, re-read an S row once per output channel (126 x 5 KB per row from L2: 0.8 GB for B = 64, measured 0.14 ms =
// 3.6 % of the HBM roofline for 39 MB of operands).  Here a block owns 8 rows x (up to) 64 channels of one head and walks
// K in steps of 128 (ten barrier rounds for K = 1280; a first tiled attempt with 32-row tiles and 32-wide K steps had too
// few blocks and forty exposed load latencies each and was slower than the warp-per-output kernel).  grid = (row tiles,
// heads); 256 threads; a thread accumulates 2 rows x 1 channel with float4 shared loads along K.
constexpr int GP_ROWS = 8, GP_COLS = 64, GP_KT = 128, GP_PITCH = GP_KT + 4;
__global__ void __launch_bounds__(256) grd_project_tiled_kernel(const float* __restrict__ S, HeadTable t, int B, int W, int K, int H) {
  __shared__ __align__(16) float sS[GP_ROWS][GP_PITCH];
  __shared__ __align__(16) float sW[GP_COLS][GP_PITCH];
  const int l = blockIdx.y;
  const int c = t.c[l];
  const int BW = B * W;
  const int r0 = blockIdx.x * GP_ROWS;
  const int tid = threadIdx.x, col = tid & 63, rg = tid >> 6;
  const float* Sl = S + (int64_t)l * BW * K;
  const float* Wl = t.w1[l];
  float vs = 0.f;
  for (int h = 0; h < H; ++h) vs += t.w2[l][h];
  for (int n0 = 0; n0 < c; n0 += GP_COLS) {
    const int nc = min(GP_COLS, c - n0);
    float acc0 = 0.f, acc1 = 0.f;
    for (int k0 = 0; k0 < K; k0 += GP_KT) {
      for (int i = tid; i < GP_ROWS * (GP_KT / 4); i += 256) {          // S tile: one float4 per thread
        const int r = i / (GP_KT / 4), k4 = (i - r * (GP_KT / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < BW && k0 + k4 < K) v = *reinterpret_cast<const float4*>(Sl + (int64_t)(r0 + r) * K + k0 + k4);
        *reinterpret_cast<float4*>(&sS[r][k4]) = v;
      }
      for (int i = tid; i < nc * (GP_KT / 4); i += 256) {               // W tile: up to eight float4 per thread, all in flight
        const int n = i / (GP_KT / 4), k4 = (i - n * (GP_KT / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + k4 < K) v = __ldg(reinterpret_cast<const float4*>(Wl + (int64_t)(n0 + n) * K + k0 + k4));
        *reinterpret_cast<float4*>(&sW[n][k4]) = v;
      }
      __syncthreads();
      if (col < nc) {
#pragma unroll 8
        for (int k4 = 0; k4 < GP_KT; k4 += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&sW[col][k4]);
          const float4 a0 = *reinterpret_cast<const float4*>(&sS[2 * rg][k4]);
          const float4 a1 = *reinterpret_cast<const float4*>(&sS[2 * rg + 1][k4]);
          acc0 = fmaf(a0.x, b.x, acc0); acc0 = fmaf(a0.y, b.y, acc0); acc0 = fmaf(a0.z, b.z, acc0); acc0 = fmaf(a0.w, b.w, acc0);
          acc1 = fmaf(a1.x, b.x, acc1); acc1 = fmaf(a1.y, b.y, acc1); acc1 = fmaf(a1.z, b.z, acc1); acc1 = fmaf(a1.w, b.w, acc1);
        }
      }
      __syncthreads();
    }
    if (col < nc) {
      const int ch = n0 + col;
      const float add = t.b1[l][ch] * vs + t.b2[l][0];
      const int r = r0 + 2 * rg;
      if (r < BW) t.out[l][(int64_t)r * c + ch] = acc0 + add;
      if (r + 1 < BW) t.out[l][(int64_t)(r + 1) * c + ch] = acc1 + add;
    }
  }
}

}  // namespace ccvpe

extern "C" int ccvpe_grd_descriptors(const void* feat, int dtype, int B, int K, int H, int W, int64_t sb, int64_t sk,
                                     int64_t sh, int64_t sw, int n_heads, const float* const* w1, const float* const* b1,
                                     const float* const* w2, const float* const* b2, const int32_t* c, float* const* out,
                                     float* scratch, void* stream) {
  using namespace ccvpe;
  CCVPE_REQUIRE(feat && w1 && b1 && w2 && b2 && c && out && scratch, "ccvpe_grd_descriptors: null pointer");
  CCVPE_REQUIRE(n_heads >= 1 && n_heads <= kMaxHeads, "ccvpe_grd_descriptors: n_heads=%d out of range", n_heads);
  CCVPE_REQUIRE(B > 0 && K > 0 && K % 4 == 0 && H > 0 && W > 0, "ccvpe_grd_descriptors: bad shape B=%d K=%d H=%d W=%d", B, K, H, W);
  CCVPE_REQUIRE(dtype == CCVPE_F32 || dtype == CCVPE_BF16, "ccvpe_grd_descriptors: bad dtype %d", dtype);
  HeadTable t;
  memset(&t, 0, sizeof(t));
  t.n = n_heads;
  for (int l = 0; l < n_heads; ++l) {
    CCVPE_REQUIRE(w1[l] && b1[l] && w2[l] && b2[l] && out[l] && c[l] > 0, "ccvpe_grd_descriptors: bad head %d", l);
    CCVPE_REQUIRE(aligned16(w1[l]), "ccvpe_grd_descriptors: w1 must be 16-byte aligned");
    t.w1[l] = w1[l]; t.b1[l] = b1[l]; t.w2[l] = w2[l]; t.b2[l] = b2[l]; t.out[l] = out[l]; t.c[l] = c[l];
    t.c_prefix[l + 1] = t.c_prefix[l] + c[l];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * K * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  const bool k_fastest = (sk == 1);
  if (dtype == CCVPE_F32)
    grd_height_reduce_all_kernel<float><<<blocks, 256, 0, st>>>((const float*)feat, B, K, H, W, sb, sk, sh, sw, t, scratch, k_fastest);
  else
    grd_height_reduce_all_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)feat, B, K, H, W, sb, sk, sh, sw, t, scratch, k_fastest);
  CCVPE_LAUNCH_CHECK("grd_height_reduce_all_kernel");
  const dim3 pgrid((unsigned)(((int64_t)B * W + GP_ROWS - 1) / GP_ROWS), (unsigned)n_heads);
  grd_project_tiled_kernel<<<pgrid, 256, 0, st>>>(scratch, t, B, W, K, H);
  CCVPE_LAUNCH_CHECK("grd_project_tiled_kernel");
  return CCVPE_OK;
}

extern "C" int ccvpe_grd_descriptor(const void* feat, int dtype, int B, int K, int H, int W, int64_t sb, int64_t sk,
                                    int64_t sh, int64_t sw, const float* w1, const float* b1, const float* w2,
                                    const float* b2, int c, float* out, float* scratch, void* stream) {
  using namespace ccvpe;
  CCVPE_REQUIRE(feat && w1 && b1 && w2 && b2 && out && scratch, "ccvpe_grd_descriptor: null pointer");
  CCVPE_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0 && c > 0, "ccvpe_grd_descriptor: bad shape B=%d K=%d H=%d W=%d c=%d", B,
                K, H, W, c);
  CCVPE_REQUIRE(dtype == CCVPE_F32 || dtype == CCVPE_BF16, "ccvpe_grd_descriptor: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = (int64_t)B * K * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  bool k_fastest = (sk == 1);
  if (dtype == CCVPE_F32)
    grd_height_reduce_kernel<float><<<blocks, 256, 0, st>>>((const float*)feat, B, K, H, W, sb, sk, sh, sw, w2, scratch,
                                                            k_fastest);
  else
    grd_height_reduce_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)feat, B, K, H, W, sb, sk, sh,
                                                                    sw, w2, scratch, k_fastest);
  CCVPE_LAUNCH_CHECK("grd_height_reduce_kernel");
  int64_t outputs = (int64_t)B * W * c;
  int wpb = 8;
  grd_project_kernel<<<(unsigned)((outputs + wpb - 1) / wpb), wpb * 32, 0, st>>>(scratch, w1, b1, w2, b2, B * W, K, H, c,
                                                                                out);
  CCVPE_LAUNCH_CHECK("grd_project_kernel");
  return CCVPE_OK;
}
