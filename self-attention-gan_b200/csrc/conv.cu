// Implicit-GEMM convolution kernels, NHWC fp32, Keras kernel layouts (FP32_STRICT math mode:
// fp32 operands, fp32 accumulate on the CUDA cores).
//
//   conv_fwd    y = act(conv(x, w) + b)          Conv2D            discriminator.py:8, generator.py:36,
//                                                 1x1 convs         layers.py:82-85, Dense generator.py:25
//   conv_dgrad  dx = conv^T(dy, w)               backward-data; ALSO the forward of Conv2DTranspose
//                                                 (generator.py:8-9), split into stride^2 parity classes
//                                                 so no multiply-by-zero work is done
//   conv_wgrad  dw = x^T (*) dy, db = sum dy     backward-filter, split over the pixel axis
//
// GEMM views (m = output pixel, n = output channel, k = reduction):
//   fwd    M = B*Ho*Wo, N = Cout, K = kh*kw*Cin; A[m,k] gathered from x, B[k,n] = w (HWIO is [K,N] row-major)
//   dgrad  per class: M = B*Hc*Wc, N = Cin, K = taps*Cout; A[m,k] gathered from dy, B[k,n] = w[kh,kw,n,co]
//   wgrad  M' = kh*kw*Cin, N = Cout, reduction over pixels; atomically accumulated over grid.z splits
#include "common.cuh"

namespace sagan {

// tensor-core variants (conv_tc.cu)
bool conv_tc_fwd_ok(const CG& g, const float* x, const float* w);
bool conv_tc_dgrad_ok(const CG& g, const float* dy, const float* w);
bool conv_tc_wgrad_ok(const CG& g, const float* x, const float* dy);
int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, const CG& g, int act, float slope, cudaStream_t st);
int conv_tc_dgrad(const float* dy, const float* w, float* dx, const CG& g, cudaStream_t st);
int conv_tc_wgrad(const float* x, const float* dy, float* dw, float* dbias, const CG& g, cudaStream_t st);
int conv_tc_get_precision();
void conv_tc_set_precision(int p);
// direct kernels for the 3-channel image side (conv_small.cu)
bool conv_small_ok(const CG& g, const void* a, const void* b);
bool conv_small_fwd_ok(const CG& g, const void* a, const void* b);
bool conv_small_wgrad_ok(const CG& g, const void* a, const void* b);
int conv_small_fwd(const float* x, const float* w, const float* bias, float* y, const CG& g, int act, float slope, cudaStream_t st);
int conv_small_dgrad(const float* dy, const float* w, float* dx, const CG& g, cudaStream_t st);
int conv_small_wgrad(const float* x, const float* dy, float* dw, float* db, const CG& g, cudaStream_t st);

constexpr int CV_THREADS = 256;
constexpr int CV_BK = 16;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == SAGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == SAGAN_ACT_TANH) return tanhf(v);
  return v;
}

// ============================================================================ forward
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CV_THREADS)
conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ y, CG g, int act, float slope, int vecA, int vecB) {
  constexpr int BK = CV_BK;
  constexpr int A_PER = BM * BK / 4 / CV_THREADS;   // float4 loads per thread for the A tile
  static_assert(A_PER >= 1, "tile too small");
  static_assert((BM / TM) * (BN / TN) == CV_THREADS, "thread tiling");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  int a_b[A_PER], a_hi0[A_PER], a_wi0[A_PER];
  bool a_ok[A_PER];
  const int kq = (tid & 3) * 4;
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    const int m = m0 + (tid >> 2) + j * 64;
    a_ok[j] = m < g.M;
    const int mm = a_ok[j] ? m : 0;
    const int b = mm / (g.Ho * g.Wo), rem = mm - b * (g.Ho * g.Wo);
    const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
    a_b[j] = b; a_hi0[j] = ho * g.S - g.PT; a_wi0[j] = wo * g.S - g.PL;
  }
  constexpr int B_F4 = BK * BN / 4;                 // float4 slots in the B tile
  const int b_row = tid / (BN / 4), b_col = (tid % (BN / 4)) * 4;
  const bool b_active = tid < B_F4;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 a_reg[A_PER];
  float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + kq;
      if (a_ok[j]) {
        if (vecA) {
          if (k < g.K) {
            const int tap = k / g.Cin, ci = k - tap * g.Cin;
            const int kh = tap / g.KW, kw = tap - kh * g.KW;
            const int hi = a_hi0[j] + kh, wi = a_wi0[j] + kw;
            if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
              const float4 t = ld4(x + ((size_t)(a_b[j] * g.H + hi) * g.W + wi) * g.Cin + ci);
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kk = k + q;
            if (kk < g.K) {
              const int tap = kk / g.Cin, ci = kk - tap * g.Cin;
              const int kh = tap / g.KW, kw = tap - kh * g.KW;
              const int hi = a_hi0[j] + kh, wi = a_wi0[j] + kw;
              if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
                v[q] = x[((size_t)(a_b[j] * g.H + hi) * g.W + wi) * g.Cin + ci];
            }
          }
        }
      }
      a_reg[j] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (b_active) {
      const int k = k0 + b_row, n = n0 + b_col;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < g.K) {
        if (vecB) {
          if (n < g.Cout) {
            const float4 t = ld4(w + (size_t)k * g.Cout + n);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (n + q < g.Cout) v[q] = w[(size_t)k * g.Cout + n + q];
        }
      }
      b_reg = make_float4(v[0], v[1], v[2], v[3]);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int row = (tid >> 2) + j * 64;
      As[kq + 0][row] = a_reg[j].x; As[kq + 1][row] = a_reg[j].y;
      As[kq + 2][row] = a_reg[j].z; As[kq + 3][row] = a_reg[j].w;
    }
    if (b_active) *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = b_reg;
  };

  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Cout) {
        float v = acc[i][j] + (bias ? bias[n] : 0.f);
        y[(size_t)m * g.Cout + n] = apply_act(v, act, slope);
      }
    }
  }
}

// ============================================================================ dgrad / transposed conv
// One parity class (rh, rw) per blockIdx.z: output pixels hi with (hi + PT) % S == rh use taps kh = rh + S*th.
struct DG {
  int hi_first, wi_first, Hc, Wc;  // first pixel / pixel count of this class along each axis
  int nth, ntw;                    // taps of this class along each axis
};

__device__ __forceinline__ DG dgrad_class(const CG& g, int rh, int rw) {
  DG d;
  d.hi_first = ((rh - g.PT) % g.S + g.S) % g.S;
  d.wi_first = ((rw - g.PL) % g.S + g.S) % g.S;
  d.Hc = d.hi_first < g.H ? (g.H - d.hi_first + g.S - 1) / g.S : 0;
  d.Wc = d.wi_first < g.W ? (g.W - d.wi_first + g.S - 1) / g.S : 0;
  d.nth = rh < g.KH ? (g.KH - rh + g.S - 1) / g.S : 0;
  d.ntw = rw < g.KW ? (g.KW - rw + g.S - 1) / g.S : 0;
  return d;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CV_THREADS)
conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, CG g,
                  int vecA) {
  constexpr int BK = CV_BK;
  constexpr int A_PER = BM * BK / 4 / CV_THREADS;
  static_assert((BM / TM) * (BN / TN) == CV_THREADS, "thread tiling");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int rh = blockIdx.z / g.S, rw = blockIdx.z - rh * g.S;
  const DG c = dgrad_class(g, rh, rw);
  const int Mc = g.B * c.Hc * c.Wc;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= Mc) return;
  const int Kc = c.nth * c.ntw * g.Cout;

  // output pixel of each A row handled by this thread
  int a_b[A_PER], a_hq[A_PER], a_wq[A_PER];
  bool a_ok[A_PER];
  const int kq = (tid & 3) * 4;
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    const int m = m0 + (tid >> 2) + j * 64;
    a_ok[j] = m < Mc;
    const int mm = a_ok[j] ? m : 0;
    const int b = mm / (c.Hc * c.Wc), rem = mm - b * (c.Hc * c.Wc);
    const int ih = rem / c.Wc, iw = rem - ih * c.Wc;
    const int hi = c.hi_first + ih * g.S, wi = c.wi_first + iw * g.S;
    a_b[j] = b;
    a_hq[j] = (hi + g.PT - rh) / g.S;   // ho = hq - th
    a_wq[j] = (wi + g.PL - rw) / g.S;
  }
  // B tile [BK][BN]: element (k, n) = w[kh, kw, n, co]; contiguous along co = k -> float4 along k
  constexpr int B_SLOTS = BN * (BK / 4);
  const int b_n = tid / (BK / 4), b_kq = (tid % (BK / 4)) * 4;
  const bool b_active = tid < B_SLOTS;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 a_reg[A_PER];
  float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + kq;
      if (a_ok[j]) {
        if (vecA) {
          if (k < Kc) {
            const int tap = k / g.Cout, co = k - tap * g.Cout;
            const int th = tap / c.ntw, tw = tap - th * c.ntw;
            const int ho = a_hq[j] - th, wo = a_wq[j] - tw;
            if (ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo) {
              const float4 t = ld4(dy + ((size_t)(a_b[j] * g.Ho + ho) * g.Wo + wo) * g.Cout + co);
              v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kk = k + q;
            if (kk < Kc) {
              const int tap = kk / g.Cout, co = kk - tap * g.Cout;
              const int th = tap / c.ntw, tw = tap - th * c.ntw;
              const int ho = a_hq[j] - th, wo = a_wq[j] - tw;
              if (ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo)
                v[q] = dy[((size_t)(a_b[j] * g.Ho + ho) * g.Wo + wo) * g.Cout + co];
            }
          }
        }
      }
      a_reg[j] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (b_active) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int n = n0 + b_n;
      if (n < g.Cin) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kk = k0 + b_kq + q;
          if (kk < Kc) {
            const int tap = kk / g.Cout, co = kk - tap * g.Cout;
            const int th = tap / c.ntw, tw = tap - th * c.ntw;
            const int kh = rh + th * g.S, kw = rw + tw * g.S;
            v[q] = w[((size_t)(kh * g.KW + kw) * g.Cin + n) * g.Cout + co];
          }
        }
      }
      b_reg = make_float4(v[0], v[1], v[2], v[3]);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int row = (tid >> 2) + j * 64;
      As[kq + 0][row] = a_reg[j].x; As[kq + 1][row] = a_reg[j].y;
      As[kq + 2][row] = a_reg[j].z; As[kq + 3][row] = a_reg[j].w;
    }
    if (b_active) {
      Bs[b_kq + 0][b_n] = b_reg.x; Bs[b_kq + 1][b_n] = b_reg.y;
      Bs[b_kq + 2][b_n] = b_reg.z; Bs[b_kq + 3][b_n] = b_reg.w;
    }
  };

  if (Kc > 0) fetch(0);
  for (int k0 = 0; k0 < Kc; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < Kc) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= Mc) continue;
    const int b = m / (c.Hc * c.Wc), rem = m - b * (c.Hc * c.Wc);
    const int ih = rem / c.Wc, iw = rem - ih * c.Wc;
    const int hi = c.hi_first + ih * g.S, wi = c.wi_first + iw * g.S;
    float* out = dx + ((size_t)(b * g.H + hi) * g.W + wi) * g.Cin;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Cin) out[n] = acc[i][j];
    }
  }
}

// ============================================================================ wgrad
// dw[kk, n] = sum_m A[m, kk] dy[m, n]; tile 64 (kk) x 64 (n), reduction tiles of 16 pixels,
// pixel range split over blockIdx.z and accumulated with fp32 atomics (dw zeroed by the caller wrapper).
__global__ void __launch_bounds__(CV_THREADS)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, CG g,
                  int m_per_split, int vecA, int vecB) {
  constexpr int BK = CV_BK, BM = 64, BN = 64, TM = 4, TN = 4;
  __shared__ __align__(16) float As[BK][BM + 4];   // [pixel][kk]
  __shared__ __align__(16) float Bs[BK][BN + 4];   // [pixel][n]
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int kk0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int m_begin = blockIdx.z * m_per_split, m_end = min(g.M, m_begin + m_per_split);
  if (m_begin >= m_end) return;

  const int l_row = tid >> 4, l_col = (tid & 15) * 4;   // pixel row within the tile, column quad
  // (kh, kw, ci) of this thread's A columns are fixed for the whole kernel
  int a_kh[4], a_kw[4], a_ci[4];
  bool a_kok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int kk = kk0 + l_col + q;
    a_kok[q] = kk < g.K;
    const int k2 = a_kok[q] ? kk : 0;
    const int tap = k2 / g.Cin;
    a_ci[q] = k2 - tap * g.Cin;
    a_kh[q] = tap / g.KW;
    a_kw[q] = tap - a_kh[q] * g.KW;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 a_reg, b_reg;

  auto fetch = [&](int mt) {
    const int m = mt + l_row;
    float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
      const int b = m / (g.Ho * g.Wo), rem = m - b * (g.Ho * g.Wo);
      const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
      const int hi0 = ho * g.S - g.PT, wi0 = wo * g.S - g.PL;
      if (vecA) {
        if (a_kok[0]) {
          const int hi = hi0 + a_kh[0], wi = wi0 + a_kw[0];
          if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
            const float4 t = ld4(x + ((size_t)(b * g.H + hi) * g.W + wi) * g.Cin + a_ci[0]);
            va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (a_kok[q]) {
            const int hi = hi0 + a_kh[q], wi = wi0 + a_kw[q];
            if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W)
              va[q] = x[((size_t)(b * g.H + hi) * g.W + wi) * g.Cin + a_ci[q]];
          }
        }
      }
      const int n = n0 + l_col;
      if (vecB) {
        if (n < g.Cout) {
          const float4 t = ld4(dy + (size_t)m * g.Cout + n);
          vb[0] = t.x; vb[1] = t.y; vb[2] = t.z; vb[3] = t.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n + q < g.Cout) vb[q] = dy[(size_t)m * g.Cout + n + q];
      }
    }
    a_reg = make_float4(va[0], va[1], va[2], va[3]);
    b_reg = make_float4(vb[0], vb[1], vb[2], vb[3]);
  };

  fetch(m_begin);
  for (int mt = m_begin; mt < m_end; mt += BK) {
    *reinterpret_cast<float4*>(&As[l_row][l_col]) = a_reg;
    *reinterpret_cast<float4*>(&Bs[l_row][l_col]) = b_reg;
    __syncthreads();
    if (mt + BK < m_end) fetch(mt + BK);
#pragma unroll
    for (int p = 0; p < BK; ++p) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[p][ty * TM]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[p][tx * TN]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int kk = kk0 + ty * TM + i;
    if (kk >= g.K) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.Cout) atomicAdd(dw + (size_t)kk * g.Cout + n, acc[i][j]);
    }
  }
}

// column sums of dy [M, C] -> db [C] (atomics over row chunks; db zeroed by the caller wrapper)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, long long M, int C, int rows_per_block) {
  __shared__ float sm[256];
  const int cpt = min(C, 256);                 // channels handled per pass
  const int rl = 256 / cpt;                    // row lanes
  const int c_lane = threadIdx.x % cpt, r_lane = threadIdx.x / cpt;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += cpt) {
    const int c = c0 + c_lane;
    float acc = 0.f;
    if (c < C && r_lane < rl)
      for (long long r = r0 + r_lane; r < r1; r += rl) acc += dy[r * C + c];
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (r_lane == 0 && c < C) {
      for (int j = 1; j < rl; ++j) acc += sm[j * cpt + c_lane];
      atomicAdd(db + c, acc);
    }
    __syncthreads();
  }
}

// dz = dy * act'(y)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz, long long n,
               int act, float slope) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 yy = ld4(y + i), dd = ld4(dy + i);
      float4 o;
      if (act == SAGAN_ACT_LRELU) {
        o.x = yy.x > 0.f ? dd.x : dd.x * slope; o.y = yy.y > 0.f ? dd.y : dd.y * slope;
        o.z = yy.z > 0.f ? dd.z : dd.z * slope; o.w = yy.w > 0.f ? dd.w : dd.w * slope;
      } else if (act == SAGAN_ACT_TANH) {
        o.x = dd.x * (1.f - yy.x * yy.x); o.y = dd.y * (1.f - yy.y * yy.y);
        o.z = dd.z * (1.f - yy.z * yy.z); o.w = dd.w * (1.f - yy.w * yy.w);
      } else {
        o = dd;
      }
      st4(dz + i, o);
    } else {
      for (long long q = i; q < n; ++q) {
        const float yy = y[q], dd = dy[q];
        dz[q] = act == SAGAN_ACT_LRELU ? (yy > 0.f ? dd : dd * slope)
                : act == SAGAN_ACT_TANH ? dd * (1.f - yy * yy) : dd;
      }
    }
  }
}

static int make_geom(const sagan_conv_geom* s, CG* g, const char* who) {
  SAGAN_REQUIRE(s, "%s: null geometry", who);
  SAGAN_REQUIRE(s->B > 0 && s->H > 0 && s->W > 0 && s->Cin > 0 && s->Ho > 0 && s->Wo > 0 && s->Cout > 0 && s->kh > 0 &&
                    s->kw > 0 && s->stride > 0 && s->pad_t >= 0 && s->pad_l >= 0,
                "%s: non-positive size in geometry", who);
  g->B = s->B; g->H = s->H; g->W = s->W; g->Cin = s->Cin; g->Ho = s->Ho; g->Wo = s->Wo; g->Cout = s->Cout;
  g->KH = s->kh; g->KW = s->kw; g->S = s->stride; g->PT = s->pad_t; g->PL = s->pad_l;
  const long long M = (long long)s->B * s->Ho * s->Wo, K = (long long)s->kh * s->kw * s->Cin;
  SAGAN_REQUIRE(M < (1ll << 31) && K < (1ll << 31) && (long long)s->B * s->H * s->W * s->Cin < (1ll << 40),
                "%s: problem too large", who);
  g->M = (int)M; g->K = (int)K;
  return 0;
}

// y = act(a + b[c] (bias, may be null) + r (residual, may be null)): stand-alone ReLU / LeakyReLU (the pre-activation
// blocks of /root/reference/models/discriminator.py:24-36), bias of a Conv2DTranspose (models/generator.py:11) and
// `layers.add` of the residual blocks (models/generator.py:21, models/discriminator.py:17,38)
__global__ void __launch_bounds__(256)
ew_fwd_kernel(const float* __restrict__ a, const float* __restrict__ bias, const float* __restrict__ r,
              float* __restrict__ y, long long n, int C, int act, float slope) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = a[i];
    if (bias) v += bias[(int)(i % C)];
    if (r) v += r[i];
    if (act == SAGAN_ACT_LRELU) v = v > 0.f ? v : v * slope;
    else if (act == SAGAN_ACT_TANH) v = tanhf(v);
    y[i] = v;
  }
}

static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace sagan

using namespace sagan;

extern "C" int sagan_conv2d_fwd(const float* x, const float* w, const float* bias, float* y,
                                const sagan_conv_geom* geom, int act, float slope, int math_mode,
                                sagan_stream_t stream) {
  SAGAN_REQUIRE(x && w && y, "sagan_conv2d_fwd: null pointer");
  CG g;
  int rc = make_geom(geom, &g, "sagan_conv2d_fwd");
  if (rc) return rc;
  if (conv_small_fwd_ok(g, x, y)) return conv_small_fwd(x, w, bias, y, g, act, slope, (cudaStream_t)stream);
  if (math_mode == SAGAN_MATH_BF16_TC && conv_tc_fwd_ok(g, x, w))
    return conv_tc_fwd(x, w, bias, y, g, act, slope, (cudaStream_t)stream);
  const int vecA = (g.Cin % 4 == 0 && al16(x)) ? 1 : 0;
  const int vecB = (g.Cout % 4 == 0 && al16(w)) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (g.Cout > 16) {
    dim3 grid(ceil_div(g.M, 64), ceil_div(g.Cout, 64));
    conv_fwd_kernel<64, 64, 4, 4><<<grid, CV_THREADS, 0, st>>>(x, w, bias, y, g, act, slope, vecA, vecB);
  } else {
    dim3 grid(ceil_div(g.M, 128), ceil_div(g.Cout, 16));
    conv_fwd_kernel<128, 16, 4, 2><<<grid, CV_THREADS, 0, st>>>(x, w, bias, y, g, act, slope, vecA, vecB);
  }
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_conv2d_dgrad(const float* dy, const float* w, float* dx, const sagan_conv_geom* geom,
                                  int math_mode, sagan_stream_t stream) {
  SAGAN_REQUIRE(dy && w && dx, "sagan_conv2d_dgrad: null pointer");
  CG g;
  int rc = make_geom(geom, &g, "sagan_conv2d_dgrad");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // classes without taps (kernel smaller than stride) and pixels no tap reaches stay zero
  bool full_cover = true;
  for (int r = 0; r < g.S; ++r)
    if (r >= g.KH || r >= g.KW) full_cover = false;
  if (conv_small_ok(g, dy, dx)) return conv_small_dgrad(dy, w, dx, g, st);      // writes every element of dx
  if (!full_cover) SAGAN_CUDA(cudaMemsetAsync(dx, 0, (size_t)g.B * g.H * g.W * g.Cin * sizeof(float), st));
  if (math_mode == SAGAN_MATH_BF16_TC && conv_tc_dgrad_ok(g, dy, w)) return conv_tc_dgrad(dy, w, dx, g, st);
  const int vecA = (g.Cout % 4 == 0 && al16(dy)) ? 1 : 0;
  const int Hc = ceil_div(g.H, g.S), Wc = ceil_div(g.W, g.S);
  const int Mc = g.B * Hc * Wc;
  if (g.Cin > 16) {
    dim3 grid(ceil_div(Mc, 64), ceil_div(g.Cin, 64), g.S * g.S);
    conv_dgrad_kernel<64, 64, 4, 4><<<grid, CV_THREADS, 0, st>>>(dy, w, dx, g, vecA);
  } else {
    dim3 grid(ceil_div(Mc, 128), ceil_div(g.Cin, 16), g.S * g.S);
    conv_dgrad_kernel<128, 16, 4, 2><<<grid, CV_THREADS, 0, st>>>(dy, w, dx, g, vecA);
  }
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias,
                                  const sagan_conv_geom* geom, int math_mode, sagan_stream_t stream) {
  SAGAN_REQUIRE(x && dy && dw, "sagan_conv2d_wgrad: null pointer");
  CG g;
  int rc = make_geom(geom, &g, "sagan_conv2d_wgrad");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SAGAN_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.K * g.Cout * sizeof(float), st));
  if (conv_small_wgrad_ok(g, x, dy)) {
    if (dbias) SAGAN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)g.Cout * sizeof(float), st));
    return conv_small_wgrad(x, dy, dw, dbias, g, st);
  }
  if (math_mode == SAGAN_MATH_BF16_TC && conv_tc_wgrad_ok(g, x, dy)) {
    if (dbias) SAGAN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)g.Cout * sizeof(float), st));
    return conv_tc_wgrad(x, dy, dw, dbias, g, st);
  }
  const int vecA = (g.Cin % 4 == 0 && al16(x)) ? 1 : 0;
  const int vecB = (g.Cout % 4 == 0 && al16(dy)) ? 1 : 0;
  const int tiles = ceil_div(g.K, 64) * ceil_div(g.Cout, 64);
  int splits = std::max(1, std::min(ceil_div(g.M, 64), ceil_div(num_sms() * 4, tiles)));
  int m_per_split = ceil_div(ceil_div(g.M, splits), CV_BK) * CV_BK;
  splits = ceil_div(g.M, m_per_split);
  dim3 grid(ceil_div(g.K, 64), ceil_div(g.Cout, 64), splits);
  conv_wgrad_kernel<<<grid, CV_THREADS, 0, st>>>(x, dy, dw, g, m_per_split, vecA, vecB);
  SAGAN_LAUNCH_CHECK();
  if (dbias) {
    SAGAN_CUDA(cudaMemsetAsync(dbias, 0, (size_t)g.Cout * sizeof(float), st));
    const int rows_per_block = std::max(64, ceil_div(g.M, num_sms() * 2));
    colsum_kernel<<<ceil_div(g.M, rows_per_block), 256, 0, st>>>(dy, dbias, g.M, g.Cout, rows_per_block);
    SAGAN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int sagan_conv_tc_precision(int set) {
  if (set == SAGAN_CONV_TC_SPLIT_BF16 || set == SAGAN_CONV_TC_TF32) conv_tc_set_precision(set);
  return conv_tc_get_precision();
}

extern "C" int sagan_act_bwd(const float* y, const float* dy, float* dz, long long n, int act, float slope,
                             sagan_stream_t stream) {
  SAGAN_REQUIRE(y && dy && dz && n > 0, "sagan_act_bwd: bad argument");
  SAGAN_REQUIRE(al16(y) && al16(dy) && al16(dz), "sagan_act_bwd: pointers must be 16-byte aligned");
  const int blocks = (int)std::min<long long>(num_sms() * 8, ceil_div<long long>(n, 1024));
  act_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, dy, dz, n, act, slope);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_ew_fwd(const float* a, const float* bias, const float* residual, float* y, long long n, int C,
                            int act, float slope, sagan_stream_t stream) {
  SAGAN_REQUIRE(a && y && n > 0 && (!bias || C > 0), "sagan_ew_fwd: bad argument");
  const int blocks = (int)std::min<long long>(num_sms() * 8, ceil_div<long long>(n, 256));
  ew_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, bias, residual, y, n, bias ? C : 1, act, slope);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_colsum(const float* x, float* out, long long rows, int C, sagan_stream_t stream) {
  SAGAN_REQUIRE(x && out && rows > 0 && C > 0, "sagan_colsum: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  SAGAN_CUDA(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), st));
  const int rows_per_block = (int)std::max<long long>(64, ceil_div<long long>(rows, num_sms() * 2));
  colsum_kernel<<<(unsigned)ceil_div<long long>(rows, rows_per_block), 256, 0, st>>>(x, out, rows, C, rows_per_block);
  SAGAN_LAUNCH_CHECK();
  return 0;
}
