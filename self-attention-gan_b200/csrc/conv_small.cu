// Direct (no GEMM view) fp32 convolution kernels for the layers whose channel count on one side is tiny:
//   the discriminator's first layer   SN(Conv2D(df, 4, 2, 'same')) on the 3-channel image   discriminator.py:8,21-22
//   the generator's output layer      Conv2D(3, 4, 1, 'same', tanh)                          generator.py:36
// An implicit GEMM with K = 48 or N = 3 wastes a 128-wide tile on them and is bound by its im2col gather (every input
// pixel re-read 16 x through L2); here one thread owns one pixel, keeps the whole small side in registers, reads its
// 4 x 4 window straight through L1 (neighbouring threads share 12 of 16 taps) and the weights from shared memory.
// These layers move 3 - 17 MB each, so the kernels are HBM / L1 bound, not FMA bound.  fp32 throughout: used by both
// math modes (exact to fp32 rounding; the tensor-core kernels have nothing to gain at these shapes).
//
//   fwd     y[p, :] = act(sum_taps x[p @ tap, :] w[tap, :, :] + bias)        thread = output pixel
//   dgrad   dx[q, :] = sum_{taps reaching q} dy[p(q, tap), :] w[tap, :, :]^T  thread = input pixel
//   wgrad   dw[tap, ci, co] = sum_p x[p @ tap, ci] dy[p, co], db = sum_p dy  CTA = 256 pixels, thread = 3 outputs, atomics
#include <stdlib.h>

#include "common.cuh"

namespace sagan {

constexpr int CS_THREADS = 256;

// the CIN * COUT weights of one tap: broadcast 128-bit shared-memory loads (scalar ones would make the kernels LDS-bound)
template <int N>
__device__ __forceinline__ void cs_load_tap(const float* wp, float (&wv)[N]) {
  static_assert(N % 4 == 0, "tap size must be a multiple of 4 floats");
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(wp + i);
    wv[i] = t.x; wv[i + 1] = t.y; wv[i + 2] = t.z; wv[i + 3] = t.w;
  }
}

__device__ __forceinline__ float cs_act(float v, int act, float slope) {
  if (act == SAGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == SAGAN_ACT_TANH) return tanhf(v);
  return v;
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(CS_THREADS)
conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ y, CG g, int act, float slope) {
  extern __shared__ __align__(16) float sw[];        // [KH*KW*CIN][COUT] + bias[COUT]
  const int nw = g.K * COUT;
  for (int i = threadIdx.x; i < nw; i += CS_THREADS) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += CS_THREADS) sw[nw + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int m = blockIdx.x * CS_THREADS + threadIdx.x;
  if (m >= g.M) return;
  const int b = m / (g.Ho * g.Wo), rem = m - b * (g.Ho * g.Wo);
  const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
  const int h0 = ho * g.S - g.PT, w0 = wo * g.S - g.PL;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = sw[nw + c];
  for (int kh = 0; kh < g.KH; ++kh) {
    const int hi = h0 + kh;
    if (hi < 0 || hi >= g.H) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      const int wi = w0 + kw;
      if (wi < 0 || wi >= g.W) continue;
      const float* xp = x + ((size_t)(b * g.H + hi) * g.W + wi) * CIN;
      const float* wp = sw + (kh * g.KW + kw) * CIN * COUT;
      float xv[CIN];
      if (CIN % 4 == 0) {
#pragma unroll
        for (int c = 0; c < CIN; c += 4) {
          const float4 t = ld4(xp + c);
          xv[c] = t.x; xv[c + 1] = t.y; xv[c + 2] = t.z; xv[c + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[c] = __ldg(xp + c);
      }
      float wv[CIN * COUT];
      cs_load_tap(wp, wv);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(xv[ci], wv[ci * COUT + co], acc[co]);
    }
  }
  float* yp = y + (size_t)m * COUT;
  if (COUT % 4 == 0) {
#pragma unroll
    for (int c = 0; c < COUT; c += 4)
      st4(yp + c, make_float4(cs_act(acc[c], act, slope), cs_act(acc[c + 1], act, slope), cs_act(acc[c + 2], act, slope),
                              cs_act(acc[c + 3], act, slope)));
  } else {
#pragma unroll
    for (int c = 0; c < COUT; ++c) yp[c] = cs_act(acc[c], act, slope);
  }
}

// Stride-1 forward with a tiny output side (G's output layer 16 -> 3, generator.py:36): thread = PX consecutive output
// pixels of one row.  The weights of a tap (CIN x COUT floats) are read from shared memory once per PX pixels and the
// PX + KW - 1 input columns of a kernel row are loaded once and shared by the PX windows, which takes the kernel from
// shared-memory-bound (4 FMA per 128-bit LDS at PX = 1: 54 us) to FMA / L1 bound.
template <int CIN, int COUT, int PX, int KW>
__global__ void __launch_bounds__(128)
conv_small_fwd_s1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         float* __restrict__ y, CG g, int act, float slope) {
  extern __shared__ __align__(16) float sw[];        // [KH*KW*CIN][COUT] + bias[COUT]
  const int nw = g.K * COUT;
  for (int i = threadIdx.x; i < nw; i += 128) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += 128) sw[nw + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int wgroups = g.Wo / PX;                     // host guarantees Wo % PX == 0
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= (long long)g.B * g.Ho * wgroups) return;
  const int b = (int)(t / (g.Ho * wgroups)), rem = (int)(t - (long long)b * g.Ho * wgroups);
  const int ho = rem / wgroups, wo0 = (rem - ho * wgroups) * PX;
  float acc[PX][COUT];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[p][c] = sw[nw + c];
  for (int kh = 0; kh < g.KH; ++kh) {
    const int hi = ho - g.PT + kh;
    if (hi < 0 || hi >= g.H) continue;
    float xr[PX + KW - 1][CIN];                      // the input columns wo0 - PL .. wo0 - PL + PX + KW - 2 of this row
#pragma unroll
    for (int j = 0; j < PX + KW - 1; ++j) {
      const int wi = wo0 - g.PL + j;
      const bool ok = wi >= 0 && wi < g.W;
      const float* xp = x + ((size_t)(b * g.H + hi) * g.W + (ok ? wi : 0)) * CIN;
#pragma unroll
      for (int c = 0; c < CIN; c += 4) {
        const float4 v = ok ? ld4(xp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[j][c] = v.x; xr[j][c + 1] = v.y; xr[j][c + 2] = v.z; xr[j][c + 3] = v.w;
      }
    }
#pragma unroll
    for (int kw = 0; kw < KW; ++kw) {
      float wv[CIN * COUT];
      cs_load_tap(sw + (kh * KW + kw) * CIN * COUT, wv);
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[p][co] = fmaf(xr[p + kw][ci], wv[ci * COUT + co], acc[p][co]);
    }
  }
  float* yp = y + ((size_t)(b * g.Ho + ho) * g.Wo + wo0) * COUT;      // PX * COUT contiguous floats
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) yp[p * COUT + c] = cs_act(acc[p][c], act, slope);
}

// The same layer with the input tile staged through shared memory.  In the kernel above the lanes of a warp read 16-byte
// pieces 256 bytes apart (four pixels x 16 channels per thread): every load instruction touches 32 sectors and the kernel
// is bound by the L1 / load-store path (38 us for 17 MB of input).  Here a CTA owns TR output rows x the whole width
// (TR * Wo = 512 pixels, four per thread): the (TR + KH - 1) input rows are copied with fully coalesced 16-byte loads into
// a layout whose four-pixel groups are padded from 256 to 272 bytes, so the per-thread 16-byte reads of a warp fall on
// distinct bank groups (four wavefronts per instruction, the minimum for 512 bytes).
constexpr int CS_GRP = 272;                          // bytes per padded group of four pixels x 16 channels
template <int PX, int KW>
__global__ void __launch_bounds__(128)
conv_small_fwd_s1_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ y, CG g, int act, float slope, int TR) {
  constexpr int CIN = 16, COUT = 3;
  extern __shared__ __align__(16) float sw[];        // [KH*KW*CIN][COUT] + bias[COUT] (padded to 16 bytes), then the input tile
  const int nw = g.K * COUT;
  const int w_floats = (nw + COUT + 3) / 4 * 4;
  uint8_t* tile = reinterpret_cast<uint8_t*>(sw + w_floats);
  const int ngrp = (g.W + KW - 1 + 3) / 4;           // padded groups per input row (columns -PL .. W + KW - 2 - PL)
  const int pitch = ngrp * CS_GRP;
  const int rows = TR + g.KH - 1;
  for (int i = threadIdx.x; i < nw; i += 128) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += 128) sw[nw + i] = bias ? bias[i] : 0.f;
  const int bpr = g.Ho / TR;                          // CTAs per image (host guarantees Ho % TR == 0)
  const int b = blockIdx.x / bpr, ho0 = (blockIdx.x - b * bpr) * TR;
  // zero the tile (halo columns, rows outside the image), then copy the rows that exist: 4 float4 per pixel
  for (int i = threadIdx.x; i < rows * pitch / 16; i += 128) reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int f4_per_row = g.W * 4;
  for (int i = threadIdx.x; i < rows * f4_per_row; i += 128) {
    const int rr = i / f4_per_row, rem = i - rr * f4_per_row;
    const int hi = ho0 - g.PT + rr;
    if (hi < 0 || hi >= g.H) continue;
    const int px = rem >> 2, c4 = rem & 3, ci = px + g.PL;
    const float4 v = ld4(x + ((size_t)(b * g.H + hi) * g.W + px) * CIN + c4 * 4);
    *reinterpret_cast<float4*>(tile + rr * pitch + (ci >> 2) * CS_GRP + (ci & 3) * 64 + c4 * 16) = v;
  }
  __syncthreads();
  const int wgroups = g.Wo / PX;
  const int r = threadIdx.x / wgroups, wo0 = (threadIdx.x - r * wgroups) * PX;     // output row ho0 + r, columns wo0 .. wo0 + PX - 1
  float acc[PX][COUT];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[p][c] = sw[nw + c];
  for (int kh = 0; kh < g.KH; ++kh) {
    const uint8_t* trow = tile + (r + kh) * pitch;
    float xr[PX + KW - 1][CIN];                      // tile columns wo0 .. wo0 + PX + KW - 2 (= input columns wo0 - PL ..)
#pragma unroll
    for (int j = 0; j < PX + KW - 1; ++j) {
      const int ci = wo0 + j;
      const uint8_t* xp = trow + (ci >> 2) * CS_GRP + (ci & 3) * 64;
#pragma unroll
      for (int c = 0; c < CIN; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xp + c * 4);
        xr[j][c] = v.x; xr[j][c + 1] = v.y; xr[j][c + 2] = v.z; xr[j][c + 3] = v.w;
      }
    }
#pragma unroll
    for (int kw = 0; kw < KW; ++kw) {
      float wv[CIN * COUT];
      cs_load_tap(sw + (kh * KW + kw) * CIN * COUT, wv);
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[p][co] = fmaf(xr[p + kw][ci], wv[ci * COUT + co], acc[p][co]);
    }
  }
  float* yp = y + ((size_t)(b * g.Ho + ho0 + r) * g.Wo + wo0) * COUT;      // PX * COUT contiguous floats
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int c = 0; c < COUT; ++c) yp[p * COUT + c] = cs_act(acc[p][c], act, slope);
}

// dx[b, hi, wi, ci] = sum over (kh, kw) with (hi + PT - kh) % S == 0, ho = (hi + PT - kh) / S in range (same for w):
//                     sum_co dy[b, ho, wo, co] w[kh, kw, ci, co]
template <int CIN, int COUT>
__global__ void __launch_bounds__(CS_THREADS)
conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, CG g) {
  extern __shared__ __align__(16) float sw[];        // [KH*KW][CIN][COUT]
  const int nw = g.K * COUT;
  for (int i = threadIdx.x; i < nw; i += CS_THREADS) sw[i] = w[i];
  __syncthreads();
  const long long q = (long long)blockIdx.x * CS_THREADS + threadIdx.x;
  if (q >= (long long)g.B * g.H * g.W) return;
  const int b = (int)(q / (g.H * g.W)), rem = (int)(q - (long long)b * g.H * g.W);
  const int hi = rem / g.W, wi = rem - hi * g.W;
  float acc[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) acc[c] = 0.f;
  for (int kh = 0; kh < g.KH; ++kh) {
    const int th = hi + g.PT - kh;
    if (th < 0 || th % g.S) continue;
    const int ho = th / g.S;
    if (ho >= g.Ho) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      const int tw = wi + g.PL - kw;
      if (tw < 0 || tw % g.S) continue;
      const int wo = tw / g.S;
      if (wo >= g.Wo) continue;
      const float* dp = dy + ((size_t)(b * g.Ho + ho) * g.Wo + wo) * COUT;
      const float* wp = sw + (kh * g.KW + kw) * CIN * COUT;
      if (COUT % 4 == 0) {
        // four output channels at a time: one 128-bit global load of dy, CIN broadcast 128-bit loads of the weights
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 4) {
          const float4 t = ld4(dp + c0);
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * COUT + c0);
            acc[ci] = fmaf(t.x, w4.x, acc[ci]); acc[ci] = fmaf(t.y, w4.y, acc[ci]);
            acc[ci] = fmaf(t.z, w4.z, acc[ci]); acc[ci] = fmaf(t.w, w4.w, acc[ci]);
          }
        }
      } else {
        float dv[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) dv[c] = __ldg(dp + c);
        float wv[CIN * COUT];
        cs_load_tap(wp, wv);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[ci] = fmaf(dv[co], wv[ci * COUT + co], acc[ci]);
      }
    }
  }
  float* xp = dx + (size_t)q * CIN;
  if (CIN % 4 == 0) {
#pragma unroll
    for (int c = 0; c < CIN; c += 4) st4(xp + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
  } else {
#pragma unroll
    for (int c = 0; c < CIN; ++c) xp[c] = acc[c];
  }
}

// CTA = PIX consecutive output pixels, dy tile in shared memory.  Thread t owns ONE (tap, ci) row of dw -- all COUT
// columns of it in registers -- and every NSL-th pixel of the tile: per pixel one x load (through L1) feeds COUT FMAs
// against a broadcast dy row.  The NSL pixel slices are folded through shared memory, then one fp32 atomic per gradient
// element and CTA (dw / db zeroed by the caller).  Requires KH * KW * CIN <= CS_THREADS.
template <int CIN, int COUT, int PIX>
__global__ void __launch_bounds__(CS_THREADS)
conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ db, CG g) {
  __shared__ __align__(16) float sdy[PIX][COUT];
  __shared__ int s_base[PIX], s_h0[PIX], s_w0[PIX];
  __shared__ float red[CS_THREADS][COUT + 1];
  const int m0 = blockIdx.x * PIX;
  for (int i = threadIdx.x; i < PIX * COUT; i += CS_THREADS) {
    const int p = i / COUT, c = i - p * COUT;
    sdy[p][c] = (m0 + p < g.M) ? dy[(size_t)(m0 + p) * COUT + c] : 0.f;
  }
  for (int p = threadIdx.x; p < PIX; p += CS_THREADS) {
    const int m = min(m0 + p, g.M - 1);
    const int b = m / (g.Ho * g.Wo), rem = m - b * (g.Ho * g.Wo);
    const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
    s_base[p] = b; s_h0[p] = ho * g.S - g.PT; s_w0[p] = wo * g.S - g.PL;
  }
  __syncthreads();
  const int K = g.K;                         // rows of dw
  const int nsl = CS_THREADS / K;            // pixel slices
  const int kk = threadIdx.x % K, sl = threadIdx.x / K;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  if (sl < nsl) {
    const int tap = kk / CIN, ci = kk - tap * CIN;
    const int kh = tap / g.KW, kw = tap - kh * g.KW;
    for (int p = sl; p < PIX; p += nsl) {
      const int hi = s_h0[p] + kh, wi = s_w0[p] + kw;
      if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
        const float xv = __ldg(x + ((size_t)(s_base[p] * g.H + hi) * g.W + wi) * CIN + ci);
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(xv, sdy[p][c], acc[c]);
      }
    }
  } else if (db) {
    // the threads left over (CS_THREADS - nsl * K >= COUT is checked by the host) sum the dy columns
    const int c = threadIdx.x - nsl * K;
    if (c < COUT) {
      float t = 0.f;
      for (int p = 0; p < PIX; ++p) t += sdy[p][c];
      atomicAdd(db + c, t);
    }
  }
#pragma unroll
  for (int c = 0; c < COUT; ++c) red[threadIdx.x][c] = acc[c];
  __syncthreads();
  for (int o = threadIdx.x; o < K * COUT; o += CS_THREADS) {
    const int r = o / COUT, c = o - r * COUT;
    float t = 0.f;
    for (int q = 0; q < nsl; ++q) t += red[q * K + r][c];
    atomicAdd(dw + o, t);
  }
}

static inline bool cs_al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// which (Cin, Cout) pairs have a direct kernel: the image side of D (3 -> 16) and of G (16 -> 3).  (A direct
// backward-data kernel for 16 <- 32 was measured at 57 us against 31 us on the tensor-core kernel: these kernels only
// pay where one side has 3 channels.)
static int cs_pair(const CG& g) {
  if (g.KH * g.KW > 16) return 0;
  if (g.Cin == 3 && g.Cout == 16) return 1;
  if (g.Cin == 16 && g.Cout == 3) return 2;
  return 0;
}

bool conv_small_ok(const CG& g, const void* a, const void* b) {      // backward-data
  return cs_pair(g) != 0 && cs_al16(a) && cs_al16(b);
}
// forward: 3 -> 16 (one pixel per thread) and the stride-1 16 -> 3 output layer (four pixels per thread)
static bool cs_fwd_s1(const CG& g) { return cs_pair(g) == 2 && g.S == 1 && g.KW == 4 && g.Wo % 4 == 0 && g.Wo == g.W && g.Ho == g.H; }
bool conv_small_fwd_ok(const CG& g, const void* a, const void* b) {
  return (cs_pair(g) == 1 || cs_fwd_s1(g)) && cs_al16(a) && cs_al16(b);
}
// backward-filter: only where all rows of dw fit the CTA (K <= 240 leaves >= COUT threads for the bias column sums)
bool conv_small_wgrad_ok(const CG& g, const void* a, const void* b) {
  return cs_pair(g) == 1 && g.K + 16 <= CS_THREADS && cs_al16(a) && cs_al16(b);
}

int conv_small_fwd(const float* x, const float* w, const float* bias, float* y, const CG& g, int act, float slope,
                   cudaStream_t st) {
  const unsigned nb = (unsigned)ceil_div(g.M, CS_THREADS);
  if (cs_pair(g) == 1)
    conv_small_fwd_kernel<3, 16><<<nb, CS_THREADS, (g.K * 16 + 16) * sizeof(float), st>>>(x, w, bias, y, g, act, slope);
  else {
    // tiled form: TR output rows x the whole width = 512 pixels per CTA
    const int TR = g.Wo > 0 && 512 % g.Wo == 0 ? 512 / g.Wo : 0;
    static const bool untiled = getenv("SAGAN_CONV_SMALL_UNTILED") != nullptr;   // diagnostics only
    if (TR >= 1 && g.Ho % TR == 0 && g.KH == 4 && !untiled) {
      const int ngrp = (g.W + 4 - 1 + 3) / 4;
      const size_t smem = (size_t)((g.K * 3 + 3 + 3) / 4 * 4) * sizeof(float) + (size_t)(TR + g.KH - 1) * ngrp * CS_GRP;
      static size_t configured = 0;
      if (smem > configured) {
        SAGAN_CUDA(cudaFuncSetAttribute(conv_small_fwd_s1_tiled_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      conv_small_fwd_s1_tiled_kernel<4, 4><<<(unsigned)(g.B * (g.Ho / TR)), 128, smem, st>>>(x, w, bias, y, g, act, slope, TR);
    } else {
      conv_small_fwd_s1_kernel<16, 3, 4, 4><<<(unsigned)ceil_div(g.M / 4, 128), 128, (g.K * 3 + 3) * sizeof(float), st>>>(
          x, w, bias, y, g, act, slope);
    }
  }
  SAGAN_LAUNCH_CHECK();
  return 0;
}

int conv_small_dgrad(const float* dy, const float* w, float* dx, const CG& g, cudaStream_t st) {
  const unsigned nb = (unsigned)ceil_div<long long>((long long)g.B * g.H * g.W, CS_THREADS);
  if (cs_pair(g) == 1)
    conv_small_dgrad_kernel<3, 16><<<nb, CS_THREADS, g.K * 16 * sizeof(float), st>>>(dy, w, dx, g);
  else
    conv_small_dgrad_kernel<16, 3><<<nb, CS_THREADS, g.K * 3 * sizeof(float), st>>>(dy, w, dx, g);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// dw (and db) must be zero on entry
int conv_small_wgrad(const float* x, const float* dy, float* dw, float* db, const CG& g, cudaStream_t st) {
  constexpr int PIX = 256;
  const unsigned nb = (unsigned)ceil_div(g.M, PIX);
  conv_small_wgrad_kernel<3, 16, PIX><<<nb, CS_THREADS, 0, st>>>(x, dy, dw, db, g);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

}  // namespace sagan
