// Self-attention block, FP32_STRICT math mode (CUDA cores, fp32 everywhere): the parity anchor of
// the tensor-core path and the kernel used for the in-model channel counts C = 8..64
// (d = C/8 in {1,2,4,8}, dv = C/2 in {4,8,16,32}), where one thread can own a whole query/key row.
//
// Replaces Attention_Layer.call (/root/reference/layers.py:93-120).  The [B,N,N] map is never
// materialised: forward is a streaming (online-softmax) pass, backward recomputes P from the saved
// row log-sum-exp.
//
//   fwd:  proj (theta, phi, g)  ->  flash (softmax(theta phi^T) g, out-proj, gamma residual fused)
//   bwd:  pre  (dA = gamma dY Wo^T, D = rowsum(dA * A))
//         main (per key tile: dK, dV in registers; dQ reduced per warp, atomics per tile)
//         post (dX = dY + dQ Wq^T + dK Wk^T + dV Wv^T)
//         weight grads as 1x1-conv wgrads (X^T dQ ...), dgamma / dWo / dbo from A^T dY.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sagan {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int AT_THREADS = 128;

// ---------------------------------------------------------------------------- projections
// one thread per token: q = x Wq + bq, k = x Wk + bk, v = x Wv + bv
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_proj_kernel(const float* __restrict__ X, const float* __restrict__ Wq, const float* __restrict__ bq,
                 const float* __restrict__ Wk, const float* __restrict__ bk, const float* __restrict__ Wv,
                 const float* __restrict__ bv, float* __restrict__ Q, float* __restrict__ K, float* __restrict__ V,
                 long long T) {
  constexpr int D = C / 8, DV = C / 2;
  __shared__ float sWq[C * D], sWk[C * D], sWv[C * DV], sbq[D], sbk[D], sbv[DV];
  for (int i = threadIdx.x; i < C * D; i += AT_THREADS) { sWq[i] = Wq[i]; sWk[i] = Wk[i]; }
  for (int i = threadIdx.x; i < C * DV; i += AT_THREADS) sWv[i] = Wv[i];
  for (int i = threadIdx.x; i < D; i += AT_THREADS) { sbq[i] = bq[i]; sbk[i] = bk[i]; }
  for (int i = threadIdx.x; i < DV; i += AT_THREADS) sbv[i] = bv[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * AT_THREADS + threadIdx.x;
  if (t >= T) return;
  float x[C];
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    const float4 v = ld4(X + t * C + c);
    x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
  }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    float a = sbq[j], b = sbk[j];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a = fmaf(x[c], sWq[c * D + j], a);
      b = fmaf(x[c], sWk[c * D + j], b);
    }
    Q[t * D + j] = a;
    K[t * D + j] = b;
  }
#pragma unroll
  for (int j0 = 0; j0 < DV; j0 += 4) {
    float a[4] = {sbv[j0], sbv[j0 + 1], sbv[j0 + 2], sbv[j0 + 3]};
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 w = *reinterpret_cast<const float4*>(&sWv[c * DV + j0]);
      a[0] = fmaf(x[c], w.x, a[0]); a[1] = fmaf(x[c], w.y, a[1]);
      a[2] = fmaf(x[c], w.z, a[2]); a[3] = fmaf(x[c], w.w, a[3]);
    }
    st4(V + t * DV + j0, make_float4(a[0], a[1], a[2], a[3]));
  }
}

// ---------------------------------------------------------------------------- down-sampled keys / values
// SURVEY.md §8f row 2 (what /root/reference/layers.py:96,100,113 reaches for: `MaxPool2D` on phi and g): keys and values
// are max-pooled 2x2 / stride 2 over the [H, W] token grid, per channel, so a query attends to N/4 keys.  One thread per
// pooled position; the first maximum of a window wins (TF / cuDNN order: (0,0), (0,1), (1,0), (1,1)); its index
// (0..3) per channel is kept for the backward scatter.
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_pool_kernel(const float* __restrict__ K, const float* __restrict__ V, float* __restrict__ Kp, float* __restrict__ Vp,
                 uint8_t* __restrict__ idxK, uint8_t* __restrict__ idxV, int B, int H, int W) {
  constexpr int D = C / 8, DV = C / 2;
  const int Hp = H / 2, Wp = W / 2;
  const long long p = (long long)blockIdx.x * AT_THREADS + threadIdx.x;
  if (p >= (long long)B * Hp * Wp) return;
  const int b = (int)(p / (Hp * Wp)), r = (int)(p - (long long)b * Hp * Wp);
  const int ph = r / Wp, pw = r - ph * Wp;
  const long long t0 = (long long)b * H * W + (long long)(2 * ph) * W + 2 * pw;
  const long long tk[4] = {t0, t0 + 1, t0 + W, t0 + W + 1};
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float best = K[tk[0] * D + c];
    int arg = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = K[tk[k] * D + c];
      if (v > best) { best = v; arg = k; }
    }
    Kp[p * D + c] = best;
    if (idxK) idxK[p * D + c] = (uint8_t)arg;
  }
#pragma unroll
  for (int c = 0; c < DV; ++c) {
    float best = V[tk[0] * DV + c];
    int arg = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = V[tk[k] * DV + c];
      if (v > best) { best = v; arg = k; }
    }
    Vp[p * DV + c] = best;
    if (idxV) idxV[p * DV + c] = (uint8_t)arg;
  }
}

// backward of the pooling: the gradient of a pooled key / value channel goes to the window position that won, the
// other three get zero.  Every token-level element is written exactly once (no zero-fill needed).
template <int D, int DV>
__global__ void __launch_bounds__(AT_THREADS)
attn_unpool_kernel(const float* __restrict__ dKp, const float* __restrict__ dVp, const uint8_t* __restrict__ idxK,
                   const uint8_t* __restrict__ idxV, float* __restrict__ dK, float* __restrict__ dV, int B, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const long long p = (long long)blockIdx.x * AT_THREADS + threadIdx.x;
  if (p >= (long long)B * Hp * Wp) return;
  const int b = (int)(p / (Hp * Wp)), r = (int)(p - (long long)b * Hp * Wp);
  const int ph = r / Wp, pw = r - ph * Wp;
  const long long t0 = (long long)b * H * W + (long long)(2 * ph) * W + 2 * pw;
  const long long tk[4] = {t0, t0 + 1, t0 + W, t0 + W + 1};
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const float g = dKp[p * D + c];
    const int a = idxK[p * D + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) dK[tk[k] * D + c] = (a == k) ? g : 0.f;
  }
#pragma unroll
  for (int c = 0; c < DV; ++c) {
    const float g = dVp[p * DV + c];
    const int a = idxV[p * DV + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) dV[tk[k] * DV + c] = (a == k) ? g : 0.f;
  }
}

int attn_unpool_launch(const float* dKp, const float* dVp, const uint8_t* idxK, const uint8_t* idxV, float* dK, float* dV,
                       int B, int H, int W, int C, cudaStream_t st) {
  const unsigned nb = (unsigned)ceil_div<long long>((long long)B * (H / 2) * (W / 2), AT_THREADS);
  switch (C) {
    case 8: attn_unpool_kernel<1, 4><<<nb, AT_THREADS, 0, st>>>(dKp, dVp, idxK, idxV, dK, dV, B, H, W); break;
    case 16: attn_unpool_kernel<2, 8><<<nb, AT_THREADS, 0, st>>>(dKp, dVp, idxK, idxV, dK, dV, B, H, W); break;
    case 32: attn_unpool_kernel<4, 16><<<nb, AT_THREADS, 0, st>>>(dKp, dVp, idxK, idxV, dK, dV, B, H, W); break;
    case 64: attn_unpool_kernel<8, 32><<<nb, AT_THREADS, 0, st>>>(dKp, dVp, idxK, idxV, dK, dV, B, H, W); break;
    default: set_err("attention pooling supports C in {8,16,32,64} (C=%d)", C); return SAGAN_EUNSUPPORTED;
  }
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------- forward
// grid (ceil(N/128), B); thread i owns query row i: q[D], o[DV], running max m (log2 units), sum l.
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_strict_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                       const float* __restrict__ X, const float* __restrict__ Wo, const float* __restrict__ bo,
                       const float* __restrict__ gamma, float* __restrict__ Y, float* __restrict__ lse,
                       float* __restrict__ A, int N, int Nk) {
  constexpr int D = C / 8, DV = C / 2, KT = 128;       // N queries, Nk keys / values (Nk == N unless pooled)
  __shared__ __align__(16) float Ks[KT * D];
  __shared__ __align__(16) float Vs[KT * DV];
  __shared__ __align__(16) float sWo[DV * C];
  __shared__ float sbo[C];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int i = blockIdx.x * AT_THREADS + tid;
  const bool valid = i < N;
  const long long row = (long long)b * N + (valid ? i : 0);
  for (int e = tid; e < DV * C; e += AT_THREADS) sWo[e] = Wo[e];
  for (int e = tid; e < C; e += AT_THREADS) sbo[e] = bo[e];

  float q[D];
#pragma unroll
  for (int d = 0; d < D; ++d) q[d] = Q[row * D + d] * LOG2E;
  float o[DV];
#pragma unroll
  for (int v = 0; v < DV; ++v) o[v] = 0.f;
  float m = -INFINITY, l = 0.f;

  for (int kt = 0; kt < Nk; kt += KT) {
    const int nk = min(KT, Nk - kt);
    __syncthreads();
    const float* Kg = K + ((long long)b * Nk + kt) * D;
    const float* Vg = V + ((long long)b * Nk + kt) * DV;
    for (int e = tid; e < KT * D; e += AT_THREADS) Ks[e] = e < nk * D ? Kg[e] : 0.f;
    for (int e = tid * 4; e < KT * DV; e += AT_THREADS * 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < nk * DV) t = ld4(Vg + e);
      *reinterpret_cast<float4*>(&Vs[e]) = t;
    }
    __syncthreads();
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(q[d], Ks[(j0 + jj) * D + d], acc);
        s[jj] = (j0 + jj < nk) ? acc : -INFINITY;
        cmax = fmaxf(cmax, s[jj]);
      }
      if (cmax > m) {
        const float corr = exp2f(m - cmax);
        l *= corr;
#pragma unroll
        for (int v = 0; v < DV; ++v) o[v] *= corr;
        m = cmax;
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = exp2f(s[jj] - m);
        l += p;
#pragma unroll
        for (int v = 0; v < DV; v += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(&Vs[(j0 + jj) * DV + v]);
          o[v] = fmaf(p, vv.x, o[v]); o[v + 1] = fmaf(p, vv.y, o[v + 1]);
          o[v + 2] = fmaf(p, vv.z, o[v + 2]); o[v + 3] = fmaf(p, vv.w, o[v + 3]);
        }
      }
    }
  }
  if (!valid) return;
  const float inv = 1.0f / l;
#pragma unroll
  for (int v = 0; v < DV; ++v) o[v] *= inv;
#pragma unroll
  for (int v = 0; v < DV; v += 4) st4(A + row * DV + v, make_float4(o[v], o[v + 1], o[v + 2], o[v + 3]));
  lse[row] = (m + log2f(l)) * LN2;
  const float gm = *gamma;
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    float a[4] = {sbo[c], sbo[c + 1], sbo[c + 2], sbo[c + 3]};
#pragma unroll
    for (int v = 0; v < DV; ++v) {
      const float4 w = *reinterpret_cast<const float4*>(&sWo[v * C + c]);
      a[0] = fmaf(o[v], w.x, a[0]); a[1] = fmaf(o[v], w.y, a[1]);
      a[2] = fmaf(o[v], w.z, a[2]); a[3] = fmaf(o[v], w.w, a[3]);
    }
    const float4 xx = ld4(X + row * C + c);
    st4(Y + row * C + c, make_float4(fmaf(gm, a[0], xx.x), fmaf(gm, a[1], xx.y), fmaf(gm, a[2], xx.z),
                                     fmaf(gm, a[3], xx.w)));
  }
}

// ---------------------------------------------------------------------------- backward: pre
// dA[t,:] = gamma * dY[t,:] Wo^T, Dd[t] = dA[t,:] . A[t,:]
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_pre_kernel(const float* __restrict__ dY, const float* __restrict__ A, const float* __restrict__ Wo,
                    const float* __restrict__ gamma, float* __restrict__ dA, float* __restrict__ Dd, long long T) {
  constexpr int DV = C / 2;
  __shared__ __align__(16) float sWo[DV * C];
  for (int e = threadIdx.x; e < DV * C; e += AT_THREADS) sWo[e] = Wo[e];
  __syncthreads();
  const long long t = (long long)blockIdx.x * AT_THREADS + threadIdx.x;
  if (t >= T) return;
  const float gm = *gamma;
  float dy[C];
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    const float4 v = ld4(dY + t * C + c);
    dy[c] = v.x; dy[c + 1] = v.y; dy[c + 2] = v.z; dy[c + 3] = v.w;
  }
  float dd = 0.f;
#pragma unroll
  for (int j0 = 0; j0 < DV; j0 += 4) {
    float a[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&sWo[(j0 + jj) * C + c]);
        acc = fmaf(dy[c], w.x, acc); acc = fmaf(dy[c + 1], w.y, acc);
        acc = fmaf(dy[c + 2], w.z, acc); acc = fmaf(dy[c + 3], w.w, acc);
      }
      a[jj] = acc * gm;
    }
    const float4 av = ld4(A + t * DV + j0);
    dd = fmaf(a[0], av.x, dd); dd = fmaf(a[1], av.y, dd); dd = fmaf(a[2], av.z, dd); dd = fmaf(a[3], av.w, dd);
    st4(dA + t * DV + j0, make_float4(a[0], a[1], a[2], a[3]));
  }
  Dd[t] = dd;
}

// ---------------------------------------------------------------------------- backward: main
// grid (ceil(N/128), B): thread j owns key row j (k[D], v[DV], dk[D], dv[DV]); loops over query tiles.
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_strict_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                       const float* __restrict__ dA, const float* __restrict__ lse, const float* __restrict__ Dd,
                       float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV, int N, int Nk) {
  constexpr int D = C / 8, DV = C / 2, QT = 128, NW = AT_THREADS / 32;    // Nk key rows (K, V, dK, dV), N query rows
  __shared__ __align__(16) float Qs[QT * D];      // pre-scaled by log2(e)
  __shared__ __align__(16) float dAs[QT * DV];
  __shared__ float lses[QT], Dds[QT];
  __shared__ float dqs[NW][QT * D];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.y;
  const int j = blockIdx.x * AT_THREADS + tid;
  const bool valid = j < Nk;
  const long long krow = (long long)b * Nk + (valid ? j : 0);
  float k[D], v[DV], dk[D], dv[DV];
#pragma unroll
  for (int d = 0; d < D; ++d) { k[d] = K[krow * D + d]; dk[d] = 0.f; }
#pragma unroll
  for (int e = 0; e < DV; e += 4) {
    const float4 t = ld4(V + krow * DV + e);
    v[e] = t.x; v[e + 1] = t.y; v[e + 2] = t.z; v[e + 3] = t.w;
    dv[e] = dv[e + 1] = dv[e + 2] = dv[e + 3] = 0.f;
  }

  for (int qt = 0; qt < N; qt += QT) {
    const int nq = min(QT, N - qt);
    __syncthreads();
    const long long qrow0 = (long long)b * N + qt;
    for (int e = tid; e < QT * D; e += AT_THREADS) Qs[e] = e < nq * D ? Q[qrow0 * D + e] * LOG2E : 0.f;
    for (int e = tid * 4; e < QT * DV; e += AT_THREADS * 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < nq * DV) t = ld4(dA + qrow0 * DV + e);
      *reinterpret_cast<float4*>(&dAs[e]) = t;
    }
    for (int e = tid; e < QT; e += AT_THREADS) {
      lses[e] = e < nq ? lse[qrow0 + e] * LOG2E : 0.f;
      Dds[e] = e < nq ? Dd[qrow0 + e] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < nq; ++i) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(Qs[i * D + d], k[d], s);
      float p = exp2f(s - lses[i]);
      p = valid ? p : 0.f;
      float dp = 0.f;
#pragma unroll
      for (int e = 0; e < DV; e += 4) {
        const float4 g = *reinterpret_cast<const float4*>(&dAs[i * DV + e]);
        dp = fmaf(g.x, v[e], dp); dp = fmaf(g.y, v[e + 1], dp);
        dp = fmaf(g.z, v[e + 2], dp); dp = fmaf(g.w, v[e + 3], dp);
        dv[e] = fmaf(p, g.x, dv[e]); dv[e + 1] = fmaf(p, g.y, dv[e + 1]);
        dv[e + 2] = fmaf(p, g.z, dv[e + 2]); dv[e + 3] = fmaf(p, g.w, dv[e + 3]);
      }
      const float ds = p * (dp - Dds[i]);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        dk[d] = fmaf(ds, Qs[i * D + d], dk[d]);
        const float r = warp_sum(ds * k[d]);
        if (lane == 0) dqs[wid][i * D + d] = r;
      }
    }
    __syncthreads();
    for (int e = tid; e < nq * D; e += AT_THREADS) {
      float r = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) r += dqs[w][e];
      atomicAdd(dQ + qrow0 * D + e, r);
    }
  }
  if (!valid) return;
#pragma unroll
  for (int d = 0; d < D; ++d) dK[krow * D + d] = dk[d] * LN2;   // Qs carried a log2(e) factor
#pragma unroll
  for (int e = 0; e < DV; e += 4) st4(dV + krow * DV + e, make_float4(dv[e], dv[e + 1], dv[e + 2], dv[e + 3]));
}

// ---------------------------------------------------------------------------- backward: post
// dX = dY + dQ Wq^T + dK Wk^T + dV Wv^T
template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_post_kernel(const float* __restrict__ dY, const float* __restrict__ dQ, const float* __restrict__ dK,
                     const float* __restrict__ dV, const float* __restrict__ Wq, const float* __restrict__ Wk,
                     const float* __restrict__ Wv, float* __restrict__ dX, long long T) {
  constexpr int D = C / 8, DV = C / 2;
  __shared__ float sWq[C * D], sWk[C * D];
  __shared__ __align__(16) float sWv[C * DV];
  for (int i = threadIdx.x; i < C * D; i += AT_THREADS) { sWq[i] = Wq[i]; sWk[i] = Wk[i]; }
  for (int i = threadIdx.x; i < C * DV; i += AT_THREADS) sWv[i] = Wv[i];
  __syncthreads();
  const long long t = (long long)blockIdx.x * AT_THREADS + threadIdx.x;
  if (t >= T) return;
  float dq[D], dk[D], dv[DV];
#pragma unroll
  for (int d = 0; d < D; ++d) { dq[d] = dQ[t * D + d]; dk[d] = dK[t * D + d]; }
#pragma unroll
  for (int e = 0; e < DV; e += 4) {
    const float4 g = ld4(dV + t * DV + e);
    dv[e] = g.x; dv[e + 1] = g.y; dv[e + 2] = g.z; dv[e + 3] = g.w;
  }
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 4) {
    const float4 g = ld4(dY + t * C + c0);
    float a[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = c0 + cc;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        a[cc] = fmaf(dq[d], sWq[c * D + d], a[cc]);
        a[cc] = fmaf(dk[d], sWk[c * D + d], a[cc]);
      }
#pragma unroll
      for (int e = 0; e < DV; e += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&sWv[c * DV + e]);
        a[cc] = fmaf(dv[e], w.x, a[cc]); a[cc] = fmaf(dv[e + 1], w.y, a[cc]);
        a[cc] = fmaf(dv[e + 2], w.z, a[cc]); a[cc] = fmaf(dv[e + 3], w.w, a[cc]);
      }
    }
    st4(dX + t * C + c0, make_float4(a[0], a[1], a[2], a[3]));
  }
}

// ---------------------------------------------------------------------------- backward: parameter gradients
// All parameter gradients of the block in ONE launch (they are skinny GEMMs over the token axis):
//   G1 = [X | 1]^T [dQ dK dV]   ((C+1) x (2d+dv))   -> dWq, dWk, dWv, dbq, dbk, dbv
//   G2 = [A | 1]^T dY           ((dv+1) x C)        -> dWo' = A^T dY, dbo' = colsum(dY)   (scaled by gamma in finalize)
// Tokens are staged through shared memory in tiles.  A thread owns a 4 x 4 BLOCK of outputs (two 128-bit shared-memory
// loads feed 16 FMAs; one output per thread was shared-memory-bound at 2 loads per FMA: 67 us at T = 262 144) and, when
// there are fewer blocks than threads, one of TG interleaved token groups; the groups are folded through shared memory
// and each CTA adds its partial sums with one fp32 atomic per output (outputs zeroed by the caller).
template <int C>
__global__ void __launch_bounds__(256)
attn_wgrad_small_kernel(const float* __restrict__ X, const float* __restrict__ A, const float* __restrict__ dY,
                        const float* __restrict__ dQ, const float* __restrict__ dK, const float* __restrict__ dV,
                        float* __restrict__ dWq, float* __restrict__ dbq, float* __restrict__ dWk,
                        float* __restrict__ dbk, float* __restrict__ dWv, float* __restrict__ dbv,
                        float* __restrict__ dWo, float* __restrict__ dbo, long long T, int tokens_per_block) {
  constexpr int D = C / 8, DV = C / 2, TT = C <= 16 ? 128 : (C <= 32 ? 64 : 32);
  constexpr int LW = C + 1, RW = 2 * D + DV;          // G1: left width (X | 1), right width (dQ dK dV)
  constexpr int L2 = DV + 1, R2 = C;                  // G2: (A | 1), dY
  constexpr int LWP = (LW + 3) / 4 * 4, RWP = (RW + 3) / 4 * 4, L2P = (L2 + 3) / 4 * 4, R2P = R2;   // padded to float4
  constexpr int NB1 = (LWP / 4) * (RWP / 4), NB2 = (L2P / 4) * (R2P / 4), NBLK = NB1 + NB2;
  constexpr int TG = NBLK >= 256 ? 1 : 256 / NBLK;    // token groups when there are fewer blocks than threads
  constexpr int PER = (NBLK + 255) / 256;             // blocks per thread when there are more
  __shared__ __align__(16) float sL1[TT][LWP], sR1[TT][RWP], sL2[TT][L2P], sR2[TT][R2P];
  float acc[PER][16];
  int lo[PER], ro[PER], which[PER];
  const int tg = TG > 1 ? threadIdx.x / NBLK : 0;
#pragma unroll
  for (int p = 0; p < PER; ++p) {
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[p][e] = 0.f;
    const int blk = TG > 1 ? threadIdx.x % NBLK : threadIdx.x + p * 256;
    which[p] = (tg >= TG || blk >= NBLK) ? 0 : (blk < NB1 ? 1 : 2);
    const int bb = blk < NB1 ? blk : blk - NB1;
    const int rblocks = blk < NB1 ? RWP / 4 : R2P / 4;
    lo[p] = (bb / rblocks) * 4;
    ro[p] = (bb % rblocks) * 4;
  }
  const long long t0 = (long long)blockIdx.x * tokens_per_block;
  const long long t1 = min(T, t0 + tokens_per_block);
  for (long long tb = t0; tb < t1; tb += TT) {
    const int nt = (int)min((long long)TT, t1 - tb);
    __syncthreads();
    // tile fill with 128-bit (dQ / dK: d floats) global loads; rows beyond nt and the padding columns are zero, the
    // column after the data is the ones column that yields the bias gradients
    for (int e = threadIdx.x; e < TT * (C / 4); e += 256) {
      const int r = e / (C / 4), c = (e % (C / 4)) * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&sL1[r][c]) = r < nt ? ld4(X + (tb + r) * C + c) : z;
      *reinterpret_cast<float4*>(&sR2[r][c]) = r < nt ? ld4(dY + (tb + r) * C + c) : z;
    }
    for (int e = threadIdx.x; e < TT * (DV / 4); e += 256) {
      const int r = e / (DV / 4), c = (e % (DV / 4)) * 4;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&sL2[r][c]) = r < nt ? ld4(A + (tb + r) * DV + c) : z;
      const float4 v = r < nt ? ld4(dV + (tb + r) * DV + c) : z;
      sR1[r][2 * D + c] = v.x; sR1[r][2 * D + c + 1] = v.y; sR1[r][2 * D + c + 2] = v.z; sR1[r][2 * D + c + 3] = v.w;
    }
    for (int e = threadIdx.x; e < TT * D; e += 256) {
      const int r = e / D, c = e % D;
      sR1[r][c] = r < nt ? dQ[(tb + r) * D + c] : 0.f;
      sR1[r][D + c] = r < nt ? dK[(tb + r) * D + c] : 0.f;
    }
    for (int r = threadIdx.x; r < TT; r += 256) {
#pragma unroll
      for (int c = C; c < LWP; ++c) sL1[r][c] = (c == C && r < nt) ? 1.f : 0.f;
#pragma unroll
      for (int c = DV; c < L2P; ++c) sL2[r][c] = (c == DV && r < nt) ? 1.f : 0.f;
#pragma unroll
      for (int c = RW; c < RWP; ++c) sR1[r][c] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      if (which[p] == 0) continue;
#pragma unroll 4
      for (int r = tg; r < TT; r += TG) {
        const float4 l4 = which[p] == 1 ? *reinterpret_cast<const float4*>(&sL1[r][lo[p]]) : *reinterpret_cast<const float4*>(&sL2[r][lo[p]]);
        const float4 r4 = which[p] == 1 ? *reinterpret_cast<const float4*>(&sR1[r][ro[p]]) : *reinterpret_cast<const float4*>(&sR2[r][ro[p]]);
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[p][i * 4 + j] = fmaf(lv[i], rv[j], acc[p][i * 4 + j]);
      }
    }
  }
  // ---- fold the token groups (through the dead left tile), then one atomic per output and CTA
  __syncthreads();
  float* red = &sL1[0][0];
  static_assert(TG == 1 || NBLK * 16 <= TT * LWP, "the fold buffer (16 floats per output block) must fit the left G1 tile");
  auto emit = [&](int wh, int l0, int r0, const float* v) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = l0 + i, r = r0 + j;
        if (wh == 1) {
          if (l >= LW || r >= RW) continue;
          float* dst;
          if (l < C) dst = r < D ? dWq + l * D + r : (r < 2 * D ? dWk + l * D + (r - D) : dWv + l * DV + (r - 2 * D));
          else dst = r < D ? dbq + r : (r < 2 * D ? dbk + (r - D) : dbv + (r - 2 * D));
          atomicAdd(dst, v[i * 4 + j]);
        } else {
          if (l >= L2 || r >= R2) continue;
          atomicAdd(l < DV ? dWo + l * C + r : dbo + r, v[i * 4 + j]);
        }
      }
  };
  if (TG > 1) {
    // groups 1.. hand their partial sums to group 0 in rounds of 16 floats per thread
    for (int g = 1; g < TG; ++g) {
      if (tg == g && which[0])
#pragma unroll
        for (int e = 0; e < 16; ++e) red[(threadIdx.x % NBLK) * 16 + e] = acc[0][e];
      __syncthreads();
      if (tg == 0 && which[0])
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[0][e] += red[threadIdx.x * 16 + e];
      __syncthreads();
    }
    if (tg == 0 && which[0]) emit(which[0], lo[0], ro[0], acc[0]);
  } else {
#pragma unroll
    for (int p = 0; p < PER; ++p)
      if (which[p]) emit(which[p], lo[p], ro[p], acc[p]);
  }
}

// dgamma = sum_c bo[c] dbo'[c] + sum_{j,c} Wo[j,c] dWo'[j,c]  with dWo' = A^T dY, dbo' = colsum(dY);
// then dWo = gamma dWo', dbo = gamma dbo'.  One CTA.
__global__ void __launch_bounds__(256)
attn_bwd_finalize_kernel(const float* __restrict__ Wo, const float* __restrict__ bo, const float* __restrict__ gamma,
                         float* __restrict__ dWo, float* __restrict__ dbo, float* __restrict__ dgamma, int nW, int C) {
  __shared__ float red[32];
  const float gm = *gamma;
  float acc = 0.f;
  for (int i = threadIdx.x; i < nW; i += 256) {
    const float g = dWo[i];
    acc = fmaf(Wo[i], g, acc);
    dWo[i] = g * gm;
  }
  for (int i = threadIdx.x; i < C; i += 256) {
    const float g = dbo[i];
    acc = fmaf(bo[i], g, acc);
    dbo[i] = g * gm;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *dgamma = acc;
}

template <int C>
static int attn_fwd_strict_t(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                             const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma,
                             float* Y, float* lse, float* A, int B, int N, int PH, int PW, float* ws, cudaStream_t st) {
  constexpr int D = C / 8, DV = C / 2;
  const long long T = (long long)B * N;
  float* Q = ws;
  float* K = Q + T * D;
  float* V = K + T * D;
  attn_proj_kernel<C><<<(unsigned)ceil_div<long long>(T, AT_THREADS), AT_THREADS, 0, st>>>(X, Wq, bq, Wk, bk, Wv, bv, Q,
                                                                                            K, V, T);
  SAGAN_LAUNCH_CHECK();
  int Nk = N;
  if (PH > 0) {          // down-sampled keys / values: [PH, PW] token grid -> N / 4 pooled rows
    Nk = N / 4;
    float* Kp = V + T * DV;
    float* Vp = Kp + (long long)B * Nk * D;
    attn_pool_kernel<C><<<(unsigned)ceil_div<long long>((long long)B * Nk, AT_THREADS), AT_THREADS, 0, st>>>(
        K, V, Kp, Vp, nullptr, nullptr, B, PH, PW);
    SAGAN_LAUNCH_CHECK();
    K = Kp; V = Vp;
  }
  attn_fwd_strict_kernel<C><<<dim3(ceil_div(N, AT_THREADS), B), AT_THREADS, 0, st>>>(Q, K, V, X, Wo, bo, gamma, Y, lse,
                                                                                      A, N, Nk);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// common tail of the backward: dX from (dQ, dK, dV), parameter gradients as 1x1-conv wgrads
struct ZeroTab {
  float* p[8];
  int n[8];
};
__global__ void zero8_kernel(const ZeroTab t) {
  float* p = t.p[blockIdx.x];
  for (int i = threadIdx.x; i < t.n[blockIdx.x]; i += blockDim.x) p[i] = 0.f;
}

template <int C>
static int attn_bwd_tail_t(const float* dY, const float* X, const float* Wq, const float* Wk, const float* Wv,
                           const float* Wo, const float* bo, const float* gamma, const float* A, const float* dQ,
                           const float* dK, const float* dV, float* dX, float* dWq, float* dbq, float* dWk, float* dbk,
                           float* dWv, float* dbv, float* dWo, float* dbo, float* dgamma, int B, int N, cudaStream_t st) {
  constexpr int D = C / 8, DV = C / 2;
  const long long T = (long long)B * N;
  const unsigned tb = (unsigned)ceil_div<long long>(T, AT_THREADS);
  if (dX) {
    attn_bwd_post_kernel<C><<<tb, AT_THREADS, 0, st>>>(dY, dQ, dK, dV, Wq, Wk, Wv, dX, T);
    SAGAN_LAUNCH_CHECK();
  }
  if (dWq) {
    // the eight caller-owned gradient buffers are accumulated atomically: one launch zeroes them all
    ZeroTab zt;
    float* zp[8] = {dWq, dWk, dWv, dWo, dbq, dbk, dbv, dbo};
    const int zn[8] = {C * D, C * D, C * DV, DV * C, D, D, DV, C};
    for (int i = 0; i < 8; ++i) { zt.p[i] = zp[i]; zt.n[i] = zn[i]; }
    zero8_kernel<<<8, 256, 0, st>>>(zt);
    SAGAN_LAUNCH_CHECK();
    const int blocks = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(T, 256));
    const int tpb = (int)(ceil_div<long long>(ceil_div<long long>(T, blocks), 128) * 128);
    attn_wgrad_small_kernel<C><<<(unsigned)ceil_div<long long>(T, tpb), 256, 0, st>>>(X, A, dY, dQ, dK, dV, dWq, dbq, dWk,
                                                                                    dbk, dWv, dbv, dWo, dbo, T, tpb);
    SAGAN_LAUNCH_CHECK();
    attn_bwd_finalize_kernel<<<1, 256, 0, st>>>(Wo, bo, gamma, dWo, dbo, dgamma, DV * C, C);
    SAGAN_LAUNCH_CHECK();
  }
  return 0;
}

template <int C>
static int attn_bwd_strict_t(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                             const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                             const float* gamma, const float* lse, const float* A, float* dX, float* dWq, float* dbq,
                             float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo, float* dgamma,
                             int B, int N, int PH, int PW, float* ws, cudaStream_t st) {
  constexpr int D = C / 8, DV = C / 2;
  const long long T = (long long)B * N;
  float* dQ = ws;
  float* dK = dQ + T * D;
  float* dV = dK + T * D;
  float* Q = dV + T * DV;
  float* K = Q + T * D;
  float* V = K + T * D;
  float* dA = V + T * DV;
  float* Dd = dA + T * DV;
  const unsigned tb = (unsigned)ceil_div<long long>(T, AT_THREADS);
  attn_proj_kernel<C><<<tb, AT_THREADS, 0, st>>>(X, Wq, bq, Wk, bk, Wv, bv, Q, K, V, T);
  SAGAN_LAUNCH_CHECK();
  attn_bwd_pre_kernel<C><<<tb, AT_THREADS, 0, st>>>(dY, A, Wo, gamma, dA, Dd, T);
  SAGAN_LAUNCH_CHECK();
  SAGAN_CUDA(cudaMemsetAsync(dQ, 0, (size_t)T * D * sizeof(float), st));
  if (PH > 0) {
    // pooled keys / values: Kp, Vp, their gradients and the argmax codes live behind the token-level buffers
    const int Nk = N / 4;
    const long long Tk = (long long)B * Nk;
    float* Kp = Dd + T;
    float* Vp = Kp + Tk * D;
    float* dKp = Vp + Tk * DV;
    float* dVp = dKp + Tk * D;
    uint8_t* idxK = reinterpret_cast<uint8_t*>(dVp + Tk * DV);
    uint8_t* idxV = idxK + Tk * D;
    const unsigned pb = (unsigned)ceil_div<long long>(Tk, AT_THREADS);
    attn_pool_kernel<C><<<pb, AT_THREADS, 0, st>>>(K, V, Kp, Vp, idxK, idxV, B, PH, PW);
    SAGAN_LAUNCH_CHECK();
    attn_bwd_strict_kernel<C><<<dim3(ceil_div(Nk, AT_THREADS), B), AT_THREADS, 0, st>>>(Q, Kp, Vp, dA, lse, Dd, dQ, dKp,
                                                                                          dVp, N, Nk);
    SAGAN_LAUNCH_CHECK();
    int rc = attn_unpool_launch(dKp, dVp, idxK, idxV, dK, dV, B, PH, PW, C, st);
    if (rc) return rc;
  } else {
    attn_bwd_strict_kernel<C><<<dim3(ceil_div(N, AT_THREADS), B), AT_THREADS, 0, st>>>(Q, K, V, dA, lse, Dd, dQ, dK, dV, N, N);
    SAGAN_LAUNCH_CHECK();
  }
  return attn_bwd_tail_t<C>(dY, X, Wq, Wk, Wv, Wo, bo, gamma, A, dQ, dK, dV, dX, dWq, dbq, dWk, dbk, dWv, dbv, dWo, dbo,
                            dgamma, B, N, st);
}

// implemented in attn_tc_bwd.cu
size_t attn_tc_bwd_workspace_bytes(int B, int N, int C, bool pool);
int attn_tc_bwd_core(const float* X, const float* dY, const float* A, const float* lse, const float* Wq, const float* bq,
                     const float* Wk, const float* bk, const float* Wv, const float* bv, const float* Wo,
                     const float* gamma, float* dQ, float* dK, float* dV, int B, int N, int C, int PH, int PW, void* ws,
                     size_t ws_bytes, cudaStream_t st);

// BF16_TC backward: dQ / dK / dV on the tensor cores, then the common tail
template <int C>
static int attn_bwd_tc_t(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                         const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                         const float* gamma, const float* lse, const float* A, float* dX, float* dWq, float* dbq,
                         float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo, float* dgamma, int B,
                         int N, int PH, int PW, float* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int D = C / 8, DV = C / 2;
  const long long T = (long long)B * N;
  float* dQ = ws;
  float* dK = dQ + T * D;
  float* dV = dK + T * D;
  const size_t used = (size_t)(T * (2 * D + DV) + 64) * sizeof(float);
  SAGAN_CUDA(cudaMemsetAsync(dQ, 0, (size_t)T * D * sizeof(float), st));
  int rc = attn_tc_bwd_core(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, dQ, dK, dV, B, N, C, PH, PW,
                            reinterpret_cast<uint8_t*>(ws) + used, ws_bytes - used, st);
  if (rc) return rc;
  return attn_bwd_tail_t<C>(dY, X, Wq, Wk, Wv, Wo, bo, gamma, A, dQ, dK, dV, dX, dWq, dbq, dWk, dbk, dWv, dbv, dWo, dbo,
                            dgamma, B, N, st);
}

}  // namespace sagan

using namespace sagan;

// implemented in attn_tc.cu
namespace sagan {
size_t attn_tc_workspace_bytes(int B, int N, int C);
int attn_tc_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                const float* bv, const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A,
                int B, int N, int C, int PH, int PW, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace sagan

namespace sagan {
int attn_bwd_finalize_launch(const float* Wo, const float* bo, const float* gamma, float* dWo, float* dbo, float* dgamma,
                             int nW, int C, cudaStream_t st) {
  attn_bwd_finalize_kernel<<<1, 256, 0, st>>>(Wo, bo, gamma, dWo, dbo, dgamma, nW, C);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// attn_big_bwd.cu
size_t attn_big_bwd_workspace_bytes(int B, int N, int C);
int attn_tc_big_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                    const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma, const float* A,
                    const float* lse, float* dX, float* dWq, float* dbq, float* dWk, float* dbk, float* dWv, float* dbv,
                    float* dWo, float* dbo, float* dgamma, int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st);
bool attn_tc_big_supported(int N, int C);
}  // namespace sagan

static size_t attn_ws_bytes(int B, int N, int C, int math_mode, bool pool) {
  if (B <= 0 || N <= 0 || C <= 0) return 0;
  const long long T = (long long)B * N;
  // strict: dQ dK dV Q K V dA Dd (+ pooled K / V, their gradients and the argmax codes: < 2 C bytes per token)
  const size_t strict = (size_t)(T * (2 * C + 1) + 64) * sizeof(float) + (pool ? (size_t)T * C * 2 : 0);
  if (math_mode == SAGAN_MATH_BF16_TC) {
    if (C > 64)   // large-C path: fused forward, composed backward (attn_big_bwd.cu)
      return std::max(attn_tc_workspace_bytes(B, N, C), attn_big_bwd_workspace_bytes(B, N, C) + 256);
    const size_t small = (size_t)(T * (C / 4 + C / 2) + 64) * sizeof(float);   // dQ, dK, dV
    return std::max(strict, std::max(attn_tc_workspace_bytes(B, N, C), small + attn_tc_bwd_workspace_bytes(B, N, C, pool)));
  }
  return strict;
}

extern "C" size_t sagan_attn_workspace_bytes(int B, int N, int C, int math_mode) {
  return attn_ws_bytes(B, N, C, math_mode, false);
}

extern "C" size_t sagan_attn_pool_workspace_bytes(int B, int H, int W, int C, int math_mode) {
  if (H <= 0 || W <= 0) return 0;
  return attn_ws_bytes(B, H * W, C, math_mode, true);
}

#define SAGAN_ATTN_DISPATCH(FN, ...)                      \
  switch (C) {                                            \
    case 8: return FN<8>(__VA_ARGS__);                    \
    case 16: return FN<16>(__VA_ARGS__);                  \
    case 32: return FN<32>(__VA_ARGS__);                  \
    case 64: return FN<64>(__VA_ARGS__);                  \
    default: break;                                       \
  }

// PH, PW = token grid when keys / values are down-sampled (N == PH * PW), 0 otherwise
static int attn_fwd_impl(const char* who, const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                         const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma, float* Y,
                         float* lse, float* A_saved, int B, int N, int C, int PH, int PW, int math_mode, void* ws,
                         size_t ws_bytes, sagan_stream_t stream) {
  SAGAN_REQUIRE(X && Wq && bq && Wk && bk && Wv && bv && Wo && bo && gamma && Y && lse && A_saved && ws,
                "%s: null pointer", who);
  SAGAN_REQUIRE(B > 0 && N > 0 && C >= 8 && C % 8 == 0, "%s: need B,N > 0 and C a positive multiple of 8 (C=%d)", who, C);
  SAGAN_REQUIRE((((uintptr_t)X | (uintptr_t)Y | (uintptr_t)A_saved | (uintptr_t)ws) & 15) == 0,
                "%s: X, Y, A_saved, ws must be 16-byte aligned", who);
  const size_t need = attn_ws_bytes(B, N, C, math_mode, PH > 0);
  if (ws_bytes < need) {
    set_err("%s: workspace %zu < %zu bytes", who, ws_bytes, need);
    return SAGAN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode == SAGAN_MATH_BF16_TC)
    return attn_tc_fwd(X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, Y, lse, A_saved, B, N, C, PH, PW, ws, ws_bytes, st);
  if (math_mode != SAGAN_MATH_FP32_STRICT) {
    set_err("%s: unknown math_mode %d", who, math_mode);
    return SAGAN_EINVAL;
  }
  SAGAN_ATTN_DISPATCH(attn_fwd_strict_t, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, Y, lse, A_saved, B, N, PH, PW, (float*)ws, st);
  set_err("%s: FP32_STRICT supports C in {8,16,32,64} (C=%d); use SAGAN_MATH_BF16_TC", who, C);
  return SAGAN_EUNSUPPORTED;
}

static int attn_bwd_impl(const char* who, const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                         const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                         const float* gamma, const float* lse, const float* A_saved, float* dX, float* dWq, float* dbq,
                         float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo, float* dgamma, int B,
                         int N, int C, int PH, int PW, int math_mode, void* ws, size_t ws_bytes, sagan_stream_t stream) {
  SAGAN_REQUIRE(dY && X && Wq && bq && Wk && bk && Wv && bv && Wo && bo && gamma && lse && A_saved && ws,
                "%s: null pointer", who);
  SAGAN_REQUIRE(B > 0 && N > 0 && C >= 8 && C % 8 == 0, "%s: need B,N > 0 and C a positive multiple of 8 (C=%d)", who, C);
  const bool all_w = dWq && dbq && dWk && dbk && dWv && dbv && dWo && dbo && dgamma;
  const bool no_w = !dWq && !dbq && !dWk && !dbk && !dWv && !dbv && !dWo && !dbo && !dgamma;
  SAGAN_REQUIRE(all_w || no_w, "%s: parameter-gradient outputs must be all set or all NULL", who);
  SAGAN_REQUIRE(dX || all_w, "%s: nothing to compute", who);
  const size_t need = attn_ws_bytes(B, N, C, math_mode, PH > 0);
  if (ws_bytes < need) {
    set_err("%s: workspace %zu < %zu bytes", who, ws_bytes, need);
    return SAGAN_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode == SAGAN_MATH_BF16_TC && (C == 16 || C == 32 || C == 64)) {
    SAGAN_ATTN_DISPATCH(attn_bwd_tc_t, dY, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, lse, A_saved, dX, dWq, dbq, dWk, dbk,
                        dWv, dbv, dWo, dbo, dgamma, B, N, PH, PW, (float*)ws, ws_bytes, st);
  }
  if (math_mode == SAGAN_MATH_BF16_TC && PH == 0 && attn_tc_big_supported(N, C))
    return attn_tc_big_bwd(dY, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, A_saved, lse, dX, dWq, dbq, dWk, dbk, dWv, dbv,
                           dWo, dbo, dgamma, B, N, C, ws, ws_bytes, st);
  if (math_mode == SAGAN_MATH_FP32_STRICT || C == 8) {
    SAGAN_ATTN_DISPATCH(attn_bwd_strict_t, dY, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, lse, A_saved, dX, dWq, dbq, dWk,
                        dbk, dWv, dbv, dWo, dbo, dgamma, B, N, PH, PW, (float*)ws, st);
  }
  set_err("%s: supports C in {8,16,32,64} (FP32_STRICT, BF16_TC) and {128,256,512} with N %% 128 == 0 (BF16_TC, "
          "un-pooled); got C=%d N=%d", who, C, N);
  return SAGAN_EUNSUPPORTED;
}

extern "C" int sagan_attn_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                              const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma,
                              float* Y, float* lse, float* A_saved, int B, int N, int C, int math_mode, void* ws,
                              size_t ws_bytes, sagan_stream_t stream) {
  return attn_fwd_impl("sagan_attn_fwd", X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, Y, lse, A_saved, B, N, C, 0, 0, math_mode,
                       ws, ws_bytes, stream);
}

extern "C" int sagan_attn_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                              const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                              const float* gamma, const float* lse, const float* A_saved, float* dX, float* dWq,
                              float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo,
                              float* dgamma, int B, int N, int C, int math_mode, void* ws, size_t ws_bytes,
                              sagan_stream_t stream) {
  return attn_bwd_impl("sagan_attn_bwd", dY, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, lse, A_saved, dX, dWq, dbq, dWk, dbk,
                       dWv, dbv, dWo, dbo, dgamma, B, N, C, 0, 0, math_mode, ws, ws_bytes, stream);
}

static int check_pool_grid(const char* who, int H, int W, int C) {
  SAGAN_REQUIRE(H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "%s: the token grid must be even in both directions (H=%d, W=%d)",
                who, H, W);
  if (!(C == 8 || C == 16 || C == 32 || C == 64)) {
    set_err("%s: down-sampled keys / values are built for C in {8,16,32,64} (C=%d)", who, C);
    return SAGAN_EUNSUPPORTED;
  }
  return 0;
}

extern "C" int sagan_attn_pool_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                                   const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma,
                                   float* Y, float* lse, float* A_saved, int B, int H, int W, int C, int math_mode,
                                   void* ws, size_t ws_bytes, sagan_stream_t stream) {
  int rc = check_pool_grid("sagan_attn_pool_fwd", H, W, C);
  if (rc) return rc;
  if (C == 8) math_mode = SAGAN_MATH_FP32_STRICT;      // no tensor-core kernel at d = 1
  return attn_fwd_impl("sagan_attn_pool_fwd", X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, Y, lse, A_saved, B, H * W, C, H, W,
                       math_mode, ws, ws_bytes, stream);
}

extern "C" int sagan_attn_pool_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                                   const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                                   const float* gamma, const float* lse, const float* A_saved, float* dX, float* dWq,
                                   float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo,
                                   float* dgamma, int B, int H, int W, int C, int math_mode, void* ws, size_t ws_bytes,
                                   sagan_stream_t stream) {
  int rc = check_pool_grid("sagan_attn_pool_bwd", H, W, C);
  if (rc) return rc;
  if (C == 8) math_mode = SAGAN_MATH_FP32_STRICT;
  return attn_bwd_impl("sagan_attn_pool_bwd", dY, X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, lse, A_saved, dX, dWq, dbq, dWk,
                       dbk, dWv, dbv, dWo, dbo, dgamma, B, H * W, C, H, W, math_mode, ws, ws_bytes, stream);
}
