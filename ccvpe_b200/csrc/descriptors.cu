// a1 -- ground descriptor heads (reference models.py:22-31, 57-97, 152-157).
//
//   g[b, w*c + ch] = sum_h v[h] * (sum_k W[ch,k] F[b,k,h,w] + b1[ch]) + b2
//                  = sum_k W[ch,k] * S[b,w,k] + b1[ch] * sum_h v[h] + b2,     S[b,w,k] = sum_h v[h] F[b,k,h,w]
//
// Two small kernels: the height reduction (reads the 1 MB/pair feature volume once, any strides) and a warp-per-output
// dot product over K = 1280 with coalesced reads of both operands.  HBM-bound, ~1 MB/pair; launch latency dominates.
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

}  // namespace ccvpe

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
