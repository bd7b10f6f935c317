// Down-sampled keys / values of the self-attention block (/root/reference/layers.py:96,100,113):
// shared by the tensor-core forward (attn_tc.cu) and backward (attn_tc_bwd.cu) preparation kernels.
#pragma once
#include "common.cuh"

namespace sagan {

template <int C>
__device__ __forceinline__ void pooled_kv(const float* __restrict__ X, const float* sWk, const float* sbk, const float* sWv,
                                          const float* sbv, int b, int H, int W, int ph, int pw, float* kk, float* vv,
                                          uint8_t* ik, uint8_t* iv) {
  constexpr int D = C / 8, DV = C / 2;
  const long long t0 = (long long)b * H * W + (long long)(2 * ph) * W + 2 * pw;
  const long long tk[4] = {t0, t0 + 1, t0 + W, t0 + W + 1};
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    float x[C];
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 v = ld4(X + tk[w] * C + c);
      x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float a = sbk[j];
#pragma unroll
      for (int c = 0; c < C; ++c) a = fmaf(x[c], sWk[c * D + j], a);
      if (w == 0 || a > kk[j]) { kk[j] = a; ik[j] = (uint8_t)w; }
    }
#pragma unroll
    for (int j = 0; j < DV; ++j) {
      float a = sbv[j];
#pragma unroll
      for (int c = 0; c < C; ++c) a = fmaf(x[c], sWv[c * DV + j], a);
      if (w == 0 || a > vv[j]) { vv[j] = a; iv[j] = (uint8_t)w; }
    }
  }
}

}  // namespace sagan
