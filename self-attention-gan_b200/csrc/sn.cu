// Spectral-norm power iteration as ONE cooperative, multi-tensor, memory-bound kernel.
//
// Replaces SpectralNormalization.update_uv (/root/reference/layers.py:50-68): per wrapped kernel
// the reference issues ~12 TF micro-ops (reshape, 3 GEMVs, 2 norms, divides, reduce).  Here all
// spectrally-normalised kernels of a network are processed by one launch:
//
//   phase 1   part_s[rb][k] = sum_{r in row block rb} u[r] W[r,k]        (column pass, float4 along k)
//   phase 1b  s[k] = sum_rb part_s, partial ||s||^2                      (fixed order: deterministic)
//   phase 2   v[k] = s[k]/(||s||+eps);  t[r] = sum_k v[k] W[r,k]         (row pass, warp-shuffle reduce)
//   phase 2b  t[r] = sum_cb part_t, partial ||t||^2
//   phase 3   u[r] = t[r]/(||t||+eps); sigma = ||t||^2/(||t||+eps) [/factor]; W_bar = W / sigma
//
// sigma = sum((u W) * v) of layers.py:62 equals u . t = ||t||^2/(||t||+eps), so no third GEMV.
// Phases are separated by grid-wide barriers (cooperative launch); W is read three times but
// passes 2 and 3 hit the 126 MB L2 for every in-model matrix (5.6 MB in total) -- compulsory HBM
// traffic is one read of W and one write of W_bar (8 B/element).
#include <cooperative_groups.h>
#include <math.h>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace sagan {

constexpr int SN_THREADS = 256;
constexpr int SN_WARPS = SN_THREADS / 32;
constexpr int SN_MAX_MATS = 64;
constexpr float SN_EPS = 1e-12f;  // layers.py:4

struct SnDev {
  const float* W;
  float* u;
  float* v;
  float* Wbar;
  __nv_bfloat16* Wbar16;
  float* sigma;
  int R, K, Ip;
  float factor;
  int vec;        // K % 4 == 0 and 16-byte aligned bases
  int vecflat;    // numel % 4 == 0 and aligned: phase 3 may use float4 over the flat tensor
  int TR, nrb;    // phase 1: rows per row block, number of row blocks
  int ncb1;       // phase 1/1b: 128-column blocks
  int CB2, ncb2;  // phase 2: columns per block (multiple of 128), number of blocks
  int nrc;        // phase 2b/3: 128-row chunks
  float* part_s;  // [nrb][K]
  float* s;       // [K]
  float* part_t;  // [ncb2][R] (unused when ncb2 == 1)
  float* t;       // [R]
  float* nrm_s;   // [ncb1] partial squared norms
  float* nrm_t;   // [nrc]
  int off1, off1b, off2, off2b, off3;  // first work unit of this matrix in each phase
  long long numel;
};

struct SnTotals {
  int n;
  int tot1, tot1b, tot2, tot2b, tot3;
  int maxIp;
  unsigned long long* stamps;   // optional [8] %globaltimer values at the phase boundaries of the last iteration
};

__device__ __forceinline__ unsigned long long sn_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SN_STAMP(i)                                                                    \
  do {                                                                                 \
    if (tot.stamps && blockIdx.x == 0 && threadIdx.x == 0) tot.stamps[i] = sn_now();   \
  } while (0)

// cache-hinted accesses: W is read three times (passes 1-2: keep in L2; pass 3: last use), W_bar is write-once
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld4_keep(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 ld4_last(const float* p) {
  float4 r;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 sn_load4(const float* row, int k, int K, int vec) {
  if (vec) return ld4(row + k);
  float4 r;
  r.x = (k + 0 < K) ? row[k + 0] : 0.f;
  r.y = (k + 1 < K) ? row[k + 1] : 0.f;
  r.z = (k + 2 < K) ? row[k + 2] : 0.f;
  r.w = (k + 3 < K) ? row[k + 3] : 0.f;
  return r;
}
__device__ __forceinline__ void sn_store4(float* row, int k, int K, int vec, float4 a) {
  if (vec) {
    st4(row + k, a);
    return;
  }
  if (k + 0 < K) row[k + 0] = a.x;
  if (k + 1 < K) row[k + 1] = a.y;
  if (k + 2 < K) row[k + 2] = a.z;
  if (k + 3 < K) row[k + 3] = a.w;
}

// which matrix owns work unit `unit` of a phase (offsets cached in shared memory, ascending)
__device__ __forceinline__ int sn_find(const int* offs, int n, int unit) {
  int m = 0;
  while (m + 1 < n && offs[m + 1] <= unit) ++m;
  return m;
}

__device__ __forceinline__ float4 sn_load4_keep(const float* row, int k, int K, int vec, uint64_t pol) {
  return vec ? ld4_keep(row + k, pol) : sn_load4(row, k, K, 0);
}

__global__ void __launch_bounds__(SN_THREADS, 2)
sn_power_iter_kernel(const SnDev* __restrict__ tab, SnTotals tot) {
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_off[5][SN_MAX_MATS];
  __shared__ float s_inv[SN_MAX_MATS];   // 1/(||s||+eps) or 1/(||t||+eps) of the current phase
  __shared__ float s_sig[SN_MAX_MATS];
  __shared__ float4 s_red[SN_WARPS][32];

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * SN_WARPS + wid;
  const int nwarps = gridDim.x * SN_WARPS;
  const int n = tot.n;
  const uint64_t pol = l2_policy_evict_last();

  for (int i = threadIdx.x; i < n; i += SN_THREADS) {
    s_off[0][i] = tab[i].off1;
    s_off[1][i] = tab[i].off1b;
    s_off[2][i] = tab[i].off2;
    s_off[3][i] = tab[i].off2b;
    s_off[4][i] = tab[i].off3;
  }
  __syncthreads();
  SN_STAMP(0);

  for (int it = 0; it < tot.maxIp; ++it) {
    // ---------------------------------------------------------------- phase 1: s partials = u W
    // warp unit = TR rows x 128 columns; 8 independent 16-byte loads in flight per lane
    for (int unit = gwarp; unit < tot.tot1; unit += nwarps) {
      const int m = sn_find(s_off[0], n, unit);
      const SnDev d = tab[m];
      if (it >= d.Ip) continue;
      const int local = unit - d.off1;
      const int rb = local / d.ncb1, cb = local - rb * d.ncb1;
      const int k = cb * 128 + lane * 4;
      const int r0 = rb * d.TR, r1 = min(d.R, r0 + d.TR);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < d.K) {
        int r = r0;
        for (; r + 8 <= r1; r += 8) {
          float4 w[8];
          float uu[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w[j] = sn_load4_keep(d.W + (size_t)(r + j) * d.K, k, d.K, d.vec, pol);
            uu[j] = d.u[r + j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc.x = fmaf(uu[j], w[j].x, acc.x);
            acc.y = fmaf(uu[j], w[j].y, acc.y);
            acc.z = fmaf(uu[j], w[j].z, acc.z);
            acc.w = fmaf(uu[j], w[j].w, acc.w);
          }
        }
        for (; r < r1; ++r) {
          const float4 w = sn_load4_keep(d.W + (size_t)r * d.K, k, d.K, d.vec, pol);
          const float uu = d.u[r];
          acc.x = fmaf(uu, w.x, acc.x);
          acc.y = fmaf(uu, w.y, acc.y);
          acc.z = fmaf(uu, w.z, acc.z);
          acc.w = fmaf(uu, w.w, acc.w);
        }
        sn_store4(d.part_s + (size_t)rb * d.K, k, d.K, d.vec, acc);
      }
    }
    grid.sync();
    SN_STAMP(1);

    // ---------------------------------------------------------------- phase 1b: s, ||s||^2 partials
    // CTA unit = 128 columns: warp w folds row blocks w, w+8, ... (4 loads in flight), then warp 0 folds the 8 warp
    // sums -- a fixed order, so the result is deterministic
    for (int unit = blockIdx.x; unit < tot.tot1b; unit += gridDim.x) {
      const int m = sn_find(s_off[1], n, unit);
      const SnDev d = tab[m];
      if (it >= d.Ip) continue;                 // uniform per CTA
      const int cb = unit - d.off1b;
      const int k = cb * 128 + lane * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < d.K) {
        int rb = wid;
        for (; rb + 3 * SN_WARPS < d.nrb; rb += 4 * SN_WARPS) {
          float4 p[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) p[j] = sn_load4(d.part_s + (size_t)(rb + j * SN_WARPS) * d.K, k, d.K, d.vec);
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc.x += p[j].x; acc.y += p[j].y; acc.z += p[j].z; acc.w += p[j].w; }
        }
        for (; rb < d.nrb; rb += SN_WARPS) {
          const float4 p = sn_load4(d.part_s + (size_t)rb * d.K, k, d.K, d.vec);
          acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
      }
      s_red[wid][lane] = acc;
      __syncthreads();
      if (wid == 0) {
        acc = s_red[0][lane];
#pragma unroll
        for (int w = 1; w < SN_WARPS; ++w) {
          const float4 p = s_red[w][lane];
          acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        if (k < d.K) sn_store4(d.s, k, d.K, d.vec, acc);
        float q = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
        q = warp_sum(q);
        if (lane == 0) d.nrm_s[cb] = q;
      }
      __syncthreads();
    }
    grid.sync();
    SN_STAMP(2);

    // ---------------------------------------------------------------- phase 2: v, t partials = v W^T
    for (int m = wid; m < n; m += SN_WARPS) {
      const SnDev& d = tab[m];
      float q = 0.f;
      for (int i = lane; i < d.ncb1; i += 32) q += d.nrm_s[i];
      q = warp_sum(q);
      if (lane == 0) s_inv[m] = 1.0f / (sqrtf(q) + SN_EPS);
    }
    __syncthreads();
    // warp unit = 4 rows x CB2 columns, visited in REVERSE order: pass 1 ended on the last rows, which are the
    // lines most recently brought into L2
    for (int uu_ = gwarp; uu_ < tot.tot2; uu_ += nwarps) {
      const int unit = tot.tot2 - 1 - uu_;
      const int m = sn_find(s_off[2], n, unit);
      const SnDev d = tab[m];
      if (it >= d.Ip) continue;
      const int local = unit - d.off2;
      const int rg = local / d.ncb2, cb = local - rg * d.ncb2;
      const int r0 = rg * 4;
      const int k0 = cb * d.CB2, k1 = min(d.K, k0 + d.CB2);
      const float inv = s_inv[m];
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const bool full = (r0 + 4 <= d.R) && d.vec;
      int k = k0 + lane * 4;
      if (full) {
        const float* w0 = d.W + (size_t)r0 * d.K;
        for (; k + 128 < k1; k += 256) {          // two column steps: 8 independent W loads in flight
          float4 va = ld4(d.s + k), vb = ld4(d.s + k + 128);
          float4 wa[4], wb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wa[j] = ld4_keep(w0 + (size_t)j * d.K + k, pol);
            wb[j] = ld4_keep(w0 + (size_t)j * d.K + k + 128, pol);
          }
          va.x *= inv; va.y *= inv; va.z *= inv; va.w *= inv;
          vb.x *= inv; vb.y *= inv; vb.z *= inv; vb.w *= inv;
          if (rg == 0) { st4(d.v + k, va); st4(d.v + k + 128, vb); }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = fmaf(wa[j].x, va.x, acc[j]); acc[j] = fmaf(wa[j].y, va.y, acc[j]);
            acc[j] = fmaf(wa[j].z, va.z, acc[j]); acc[j] = fmaf(wa[j].w, va.w, acc[j]);
            acc[j] = fmaf(wb[j].x, vb.x, acc[j]); acc[j] = fmaf(wb[j].y, vb.y, acc[j]);
            acc[j] = fmaf(wb[j].z, vb.z, acc[j]); acc[j] = fmaf(wb[j].w, vb.w, acc[j]);
          }
        }
      }
      for (; k < k1; k += 128) {
        float4 vv = sn_load4(d.s, k, d.K, d.vec);
        vv.x *= inv; vv.y *= inv; vv.z *= inv; vv.w *= inv;
        if (rg == 0) sn_store4(d.v, k, d.K, d.vec, vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r0 + j < d.R) {
            const float4 w = sn_load4_keep(d.W + (size_t)(r0 + j) * d.K, k, d.K, d.vec, pol);
            acc[j] = fmaf(w.x, vv.x, acc[j]);
            acc[j] = fmaf(w.y, vv.y, acc[j]);
            acc[j] = fmaf(w.z, vv.z, acc[j]);
            acc[j] = fmaf(w.w, vv.w, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
      if (lane == 0) {
        float* dst = (d.ncb2 == 1) ? d.t : d.part_t + (size_t)cb * d.R;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (r0 + j < d.R) dst[r0 + j] = acc[j];
      }
    }
    grid.sync();
    SN_STAMP(3);

    // ---------------------------------------------------------------- phase 2b: t, ||t||^2 partials
    for (int unit = gwarp; unit < tot.tot2b; unit += nwarps) {
      const int m = sn_find(s_off[3], n, unit);
      const SnDev d = tab[m];
      if (it >= d.Ip) continue;
      const int rc = unit - d.off2b;
      const int r = rc * 128 + lane * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < d.R) {
        if (d.ncb2 == 1) {
          acc = sn_load4(d.t, r, d.R, 0);
        } else {
          for (int cb = 0; cb < d.ncb2; ++cb) {
            const float4 p = sn_load4(d.part_t + (size_t)cb * d.R, r, d.R, 0);
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
          }
          sn_store4(d.t, r, d.R, 0, acc);
        }
      }
      float q = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
      q = warp_sum(q);
      if (lane == 0) d.nrm_t[rc] = q;
    }
    grid.sync();
    SN_STAMP(4);

    // ---------------------------------------------------------------- phase 3: u, sigma, W_bar
    __syncthreads();
    for (int m = wid; m < n; m += SN_WARPS) {
      const SnDev& d = tab[m];
      float q = 0.f;
      for (int i = lane; i < d.nrc; i += 32) q += d.nrm_t[i];
      q = warp_sum(q);
      if (lane == 0) {
        const float nt = sqrtf(q);
        s_inv[m] = 1.0f / (nt + SN_EPS);
        float sg = nt * nt / (nt + SN_EPS);       // == sum((u W) * v), layers.py:62
        if (d.factor != 0.f) sg = sg / d.factor;  // layers.py:65-66
        s_sig[m] = sg;
      }
    }
    __syncthreads();
    // forward order again: pass 2 ended on the first rows
    for (int unit = gwarp; unit < tot.tot3; unit += nwarps) {
      const int m = sn_find(s_off[4], n, unit);
      const SnDev d = tab[m];
      if (it >= d.Ip) continue;
      const int local = unit - d.off3;
      if (local < d.nrc) {  // u update (+ sigma)
        const int r = local * 128 + lane * 4;
        if (r < d.R) {
          float4 tt = sn_load4(d.t, r, d.R, 0);
          const float inv = s_inv[m];
          tt.x *= inv; tt.y *= inv; tt.z *= inv; tt.w *= inv;
          sn_store4(d.u, r, d.R, 0, tt);
        }
        if (local == 0 && lane == 0) *d.sigma = s_sig[m];
      } else if (it == d.Ip - 1) {  // W_bar = W / sigma over a 2048-element chunk (layers.py:68)
        // W / sigma (layers.py:68) as a multiplication by the correctly-rounded reciprocal (<= 1.5 ulp from the quotient)
        const float rs = 1.0f / s_sig[m];
        const long long base = (long long)(local - d.nrc) * 2048;
        if (d.vecflat && !d.Wbar16 && base + 2048 <= d.numel) {
          // streaming fast path: 16 loads in flight, last use of W (evict-first), write-once W_bar
          float4 w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = ld4_last(d.W + base + j * 128 + lane * 4);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w[j].x *= rs; w[j].y *= rs; w[j].z *= rs; w[j].w *= rs;
            st4_stream(d.Wbar + base + j * 128 + lane * 4, w[j]);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const long long e = base + (long long)j * 128 + lane * 4;
          if (e < d.numel) {
            if (d.vecflat) {
              float4 w = ld4(d.W + e);
              w.x *= rs; w.y *= rs; w.z *= rs; w.w *= rs;
              st4(d.Wbar + e, w);
              if (d.Wbar16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(d.Wbar16 + e) = pk;
              }
            } else {
              for (int q = 0; q < 4 && e + q < d.numel; ++q) {
                const float w = d.W[e + q] * rs;
                d.Wbar[e + q] = w;
                if (d.Wbar16) d.Wbar16[e + q] = __float2bfloat16_rn(w);
              }
            }
          }
        }
      }
    }
    if (it + 1 < tot.maxIp) grid.sync();
  }
  SN_STAMP(5);
}

// ------------------------------------------------------------------------------------ backward
// c = sum dW_bar * W_bar (two-stage, fixed order), then dW = (dW_bar - c * u[r] v[k] / factor) / sigma
constexpr int SNB_THREADS = 256;
constexpr int SNB_MAX_BLOCKS = 1024;

__global__ void __launch_bounds__(SNB_THREADS)
sn_bwd_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ part) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * SNB_THREADS;
  for (long long i = (long long)blockIdx.x * SNB_THREADS + threadIdx.x; i < n; i += stride) acc = fmaf(a[i], b[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(SNB_THREADS)
sn_bwd_apply_kernel(const float* __restrict__ dWbar, const float* __restrict__ u, const float* __restrict__ v,
                    const float* __restrict__ sigma, float factor, const float* __restrict__ part, int nparts,
                    float* __restrict__ dW, int R, int K) {
  __shared__ float red[32];
  float c = 0.f;
  for (int i = threadIdx.x; i < nparts; i += SNB_THREADS) c += part[i];
  c = block_sum(c, red);
  const float sg = *sigma;
  const float cf = (factor != 0.f) ? c / factor : c;
  const long long n = (long long)R * K;
  const long long stride = (long long)gridDim.x * SNB_THREADS;
  for (long long i = (long long)blockIdx.x * SNB_THREADS + threadIdx.x; i < n; i += stride) {
    const int r = (int)(i / K), k = (int)(i - (long long)r * K);
    dW[i] = (dWbar[i] - cf * u[r] * v[k]) / sg;
  }
}

// ---- multi-tensor backward: every spectrally-normalised kernel of a network in two launches
constexpr int SNB_MAX_MATS = 16;
constexpr int SNB_PARTS = 64;            // partial sums per matrix
struct SnBwdTab {
  const float* dWbar[SNB_MAX_MATS];
  const float* Wbar[SNB_MAX_MATS];
  const float* u[SNB_MAX_MATS];
  const float* v[SNB_MAX_MATS];
  const float* sigma[SNB_MAX_MATS];
  float* dW[SNB_MAX_MATS];
  float factor[SNB_MAX_MATS];
  int R[SNB_MAX_MATS], K[SNB_MAX_MATS];
  int n;
};

// part[m][SNB_PARTS]: fixed-order partial sums of sum(dWbar * Wbar); grid (SNB_PARTS, n)
__global__ void __launch_bounds__(SNB_THREADS)
sn_bwd_dot_multi_kernel(const SnBwdTab t, float* __restrict__ part) {
  __shared__ float red[32];
  const int m = blockIdx.y;
  const long long n = (long long)t.R[m] * t.K[m];
  const float* a = t.dWbar[m];
  const float* b = t.Wbar[m];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * SNB_THREADS;
  for (long long i = (long long)blockIdx.x * SNB_THREADS + threadIdx.x; i < n; i += stride) acc = fmaf(a[i], b[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[m * SNB_PARTS + blockIdx.x] = acc;
}

// dW (+)= (dWbar - c u v^T / factor) / sigma; grid (blocks, n)
__global__ void __launch_bounds__(SNB_THREADS)
sn_bwd_apply_multi_kernel(const SnBwdTab t, const float* __restrict__ part, int accumulate) {
  __shared__ float red[32];
  const int m = blockIdx.y;
  float c = 0.f;
  for (int i = threadIdx.x; i < SNB_PARTS; i += SNB_THREADS) c += part[m * SNB_PARTS + i];
  c = block_sum(c, red);
  const float sg = *t.sigma[m];
  const float cf = (t.factor[m] != 0.f) ? c / t.factor[m] : c;
  const int K = t.K[m];
  const long long n = (long long)t.R[m] * K;
  const float* dWbar = t.dWbar[m];
  const float* u = t.u[m];
  const float* v = t.v[m];
  float* dW = t.dW[m];
  const long long stride = (long long)gridDim.x * SNB_THREADS;
  for (long long i = (long long)blockIdx.x * SNB_THREADS + threadIdx.x; i < n; i += stride) {
    const int r = (int)(i / K), k = (int)(i - (long long)r * K);
    const float g = (dWbar[i] - cf * u[r] * v[k]) / sg;
    dW[i] = accumulate ? dW[i] + g : g;
  }
}

}  // namespace sagan

using namespace sagan;

extern "C" int sagan_sn_backward_multi(const sagan_sn_bwd_desc* d, int n, int accumulate, void* ws, size_t ws_bytes,
                                       sagan_stream_t stream) {
  SAGAN_REQUIRE(d && ws, "sagan_sn_backward_multi: null pointer");
  SAGAN_REQUIRE(n >= 1 && n <= SNB_MAX_MATS, "sagan_sn_backward_multi: n=%d outside [1,%d]", n, SNB_MAX_MATS);
  if (ws_bytes < (size_t)SNB_MAX_MATS * SNB_PARTS * sizeof(float)) {
    set_err("sagan_sn_backward_multi: workspace %zu < %zu bytes", ws_bytes, (size_t)SNB_MAX_MATS * SNB_PARTS * sizeof(float));
    return SAGAN_EWORKSPACE;
  }
  SnBwdTab t{};
  t.n = n;
  long long biggest = 0;
  for (int i = 0; i < n; ++i) {
    SAGAN_REQUIRE(d[i].dW_bar && d[i].W_bar && d[i].u && d[i].v && d[i].sigma && d[i].dW && d[i].rows >= 1 && d[i].cols >= 1,
                  "sagan_sn_backward_multi: bad descriptor %d", i);
    t.dWbar[i] = d[i].dW_bar; t.Wbar[i] = d[i].W_bar; t.u[i] = d[i].u; t.v[i] = d[i].v; t.sigma[i] = d[i].sigma;
    t.dW[i] = d[i].dW; t.factor[i] = d[i].factor; t.R[i] = d[i].rows; t.K[i] = d[i].cols;
    biggest = std::max(biggest, (long long)d[i].rows * d[i].cols);
  }
  cudaStream_t st = (cudaStream_t)stream;
  sn_bwd_dot_multi_kernel<<<dim3(SNB_PARTS, n), SNB_THREADS, 0, st>>>(t, (float*)ws);
  SAGAN_LAUNCH_CHECK();
  const int blocks = (int)std::max<long long>(1, std::min<long long>(64, ceil_div<long long>(biggest, SNB_THREADS * 4)));
  sn_bwd_apply_multi_kernel<<<dim3(blocks, n), SNB_THREADS, 0, st>>>(t, (const float*)ws, accumulate);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------ refresh (training=False)
// The wrapped layer called with training=False (the sample dump, sagan/main.py:333): NO power iteration; sigma is
// recomputed from the STORED u, v and the CURRENT kernel, sigma = sum((u W_mat) * v) [/ factor] (layers.py:62-66), and
// W_bar = W / sigma (layers.py:68).  u and v are not written.  Off the training path: one CTA per matrix for sigma
// (fp64 accumulation over the CTA, fixed order), then a grid over the elements.
constexpr int SNR_THREADS = 512;

__global__ void __launch_bounds__(SNR_THREADS) sn_refresh_sigma_kernel(const SnDev* __restrict__ tab) {
  const SnDev d = tab[blockIdx.x];
  __shared__ double red[SNR_THREADS];
  double acc = 0.0;
  for (long long e = threadIdx.x; e < d.numel; e += SNR_THREADS) {
    const int r = (int)(e / d.K), k = (int)(e % d.K);
    acc += (double)d.u[r] * (double)d.W[e] * (double)d.v[k];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = SNR_THREADS / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float sg = (float)red[0];
    if (d.factor != 0.f) sg = sg / d.factor;
    *d.sigma = sg;
  }
}

__global__ void __launch_bounds__(256) sn_refresh_scale_kernel(const SnDev* __restrict__ tab) {
  const SnDev d = tab[blockIdx.y];
  const float rs = 1.0f / *d.sigma;      // same reciprocal form as phase 3 of the training kernel
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < d.numel; e += (long long)gridDim.x * 256) {
    const float w = d.W[e] * rs;
    d.Wbar[e] = w;
    if (d.Wbar16) d.Wbar16[e] = __float2bfloat16_rn(w);
  }
}

struct sagan_sn_plan {
  int n = 0;
  int device = 0;
  SnDev* tab_dev = nullptr;
  float* ws_dev = nullptr;
  unsigned long long* stamps_dev = nullptr;
  SnTotals tot{};
  int grid = 0;
  unsigned long long alg_bytes = 0;
};

extern "C" int sagan_sn_plan_create(const sagan_sn_desc* descs, int n, int device, sagan_sn_plan** plan_out) {
  SAGAN_REQUIRE(descs && plan_out, "sagan_sn_plan_create: null argument");
  SAGAN_REQUIRE(n >= 1 && n <= SN_MAX_MATS, "sagan_sn_plan_create: n=%d outside [1,%d]", n, SN_MAX_MATS);
  std::vector<SnDev> tab(n);
  size_t ws_floats = 0;
  SnTotals tot{};
  tot.n = n;
  unsigned long long alg = 0;
  long long total_numel = 0;
  for (int i = 0; i < n; ++i) total_numel += (long long)std::max(descs[i].rows, 0) * std::max(descs[i].cols, 0);
  if (total_numel <= 0) total_numel = 1;
  const int resident_warps = num_sms() * 2 * SN_WARPS;
  for (int i = 0; i < n; ++i) {
    const sagan_sn_desc& s = descs[i];
    SAGAN_REQUIRE(s.W && s.u && s.v && s.W_bar && s.sigma, "sagan_sn_plan_create: null pointer in descriptor %d", i);
    SAGAN_REQUIRE(s.rows >= 1 && s.cols >= 1, "sagan_sn_plan_create: bad shape [%d,%d] in descriptor %d", s.rows, s.cols, i);
    // layers.py:17-18
    SAGAN_REQUIRE(s.Ip >= 1, "The number of power iterations should be positive integer (descriptor %d: Ip=%d)", i, s.Ip);
    SnDev& d = tab[i];
    d.W = s.W; d.u = s.u; d.v = s.v; d.Wbar = s.W_bar; d.Wbar16 = (__nv_bfloat16*)s.W_bar_bf16; d.sigma = s.sigma;
    d.R = s.rows; d.K = s.cols; d.Ip = s.Ip; d.factor = s.factor;
    d.numel = (long long)s.rows * s.cols;
    const bool al = (((uintptr_t)s.W | (uintptr_t)s.W_bar) & 15) == 0 && (((uintptr_t)s.W_bar_bf16) & 7) == 0;
    d.vec = (al && (d.K % 4 == 0)) ? 1 : 0;
    d.vecflat = (al && (d.numel % 4 == 0)) ? 1 : 0;
    // work decomposition: ~3 warp units per resident warp for the matrices that dominate the plan, so that both
    // GEMV passes keep every SM's load queues full (2 CTAs x 8 warps x 8 x 16 B in flight per SM)
    const double share = (double)d.numel / (double)total_numel;
    const int want_units = std::max(1, (int)(3.0 * resident_warps * share));
    d.ncb1 = ceil_div(d.K, 128);
    int nrb = ceil_div(want_units, d.ncb1);
    nrb = std::max(1, std::min(nrb, std::min(128, ceil_div(d.R, 16))));   // >= 16 rows per unit, <= 128 partials to fold
    d.TR = ceil_div(ceil_div(d.R, nrb), 4) * 4;
    d.nrb = ceil_div(d.R, d.TR);
    int ncb2 = ceil_div(want_units, ceil_div(d.R, 4));
    ncb2 = std::max(1, std::min(ncb2, ceil_div(d.K, 512)));       // >= 512 columns per unit
    d.CB2 = ceil_div(ceil_div(d.K, ncb2), 256) * 256;
    d.ncb2 = ceil_div(d.K, d.CB2);
    d.nrc = ceil_div(d.R, 128);
    d.off1 = tot.tot1;   tot.tot1 += d.nrb * d.ncb1;
    d.off1b = tot.tot1b; tot.tot1b += d.ncb1;
    d.off2 = tot.tot2;   tot.tot2 += ceil_div(d.R, 4) * d.ncb2;
    d.off2b = tot.tot2b; tot.tot2b += d.nrc;
    d.off3 = tot.tot3;   tot.tot3 += d.nrc + (int)ceil_div(d.numel, (long long)2048);
    tot.maxIp = std::max(tot.maxIp, d.Ip);
    // workspace layout (float offsets, each region padded to 4 floats so the float4 path stays aligned)
    auto take = [&](size_t cnt) { size_t o = ws_floats; ws_floats += (cnt + 3) / 4 * 4; return o; };
    const size_t o_ps = take((size_t)d.nrb * d.K), o_s = take(d.K), o_pt = take((size_t)d.ncb2 * d.R), o_t = take(d.R);
    const size_t o_ns = take(d.ncb1), o_nt = take(d.nrc);
    d.part_s = (float*)o_ps; d.s = (float*)o_s; d.part_t = (float*)o_pt; d.t = (float*)o_t;
    d.nrm_s = (float*)o_ns; d.nrm_t = (float*)o_nt;
    const unsigned long long per = (s.W_bar_bf16 ? 10ull : 8ull) + ((d.numel * 4 > (64ll << 20)) ? 8ull : 0ull);
    alg += per * (unsigned long long)d.numel;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  SAGAN_CUDA(cudaSetDevice(device));
  sagan_sn_plan* p = new sagan_sn_plan();
  p->n = n; p->device = device; p->tot = tot; p->alg_bytes = alg;
  cudaError_t e = cudaMalloc(&p->ws_dev, ws_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&p->tab_dev, sizeof(SnDev) * n);
  if (e != cudaSuccess) {
    set_err("sagan_sn_plan_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    cudaFree(p->ws_dev);
    delete p;
    cudaSetDevice(prev);
    return (int)e;
  }
  for (auto& d : tab) {
    d.part_s = p->ws_dev + (size_t)d.part_s; d.s = p->ws_dev + (size_t)d.s;
    d.part_t = p->ws_dev + (size_t)d.part_t; d.t = p->ws_dev + (size_t)d.t;
    d.nrm_s = p->ws_dev + (size_t)d.nrm_s; d.nrm_t = p->ws_dev + (size_t)d.nrm_t;
  }
  e = cudaMemcpy(p->tab_dev, tab.data(), sizeof(SnDev) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&p->stamps_dev, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(p->stamps_dev, 0, 8 * sizeof(unsigned long long));
  p->tot.stamps = p->stamps_dev;
  int occ = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sn_power_iter_kernel, SN_THREADS, 0);
  if (e != cudaSuccess || occ < 1) {
    set_err("sagan_sn_plan_create: setup failed: %s", cudaGetErrorString(e));
    cudaFree(p->ws_dev); cudaFree(p->tab_dev);
    delete p;
    cudaSetDevice(prev);
    return e != cudaSuccess ? (int)e : SAGAN_EUNSUPPORTED;
  }
  const int max_units = std::max(std::max(tot.tot1, tot.tot2), tot.tot3);
  const int want = std::max(1, ceil_div(max_units, SN_WARPS));
  p->grid = std::min(num_sms() * std::min(occ, 2), want);
  cudaSetDevice(prev);
  *plan_out = p;
  return 0;
}

extern "C" int sagan_sn_plan_run(sagan_sn_plan* p, sagan_stream_t stream) {
  SAGAN_REQUIRE(p, "sagan_sn_plan_run: null plan");
  const SnDev* tab = p->tab_dev;
  SnTotals tot = p->tot;
  void* args[] = {(void*)&tab, (void*)&tot};
  SAGAN_CUDA(cudaLaunchCooperativeKernel((const void*)sn_power_iter_kernel, dim3(p->grid), dim3(SN_THREADS), args, 0,
                                         (cudaStream_t)stream));
  count_launch();
  return 0;
}

extern "C" int sagan_sn_plan_refresh(sagan_sn_plan* p, sagan_stream_t stream) {
  SAGAN_REQUIRE(p, "sagan_sn_plan_refresh: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  sn_refresh_sigma_kernel<<<p->n, SNR_THREADS, 0, st>>>(p->tab_dev);
  SAGAN_LAUNCH_CHECK();
  sn_refresh_scale_kernel<<<dim3(64, p->n), 256, 0, st>>>(p->tab_dev);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_sn_plan_destroy(sagan_sn_plan* p) {
  if (!p) return 0;
  cudaFree(p->ws_dev);
  cudaFree(p->tab_dev);
  cudaFree(p->stamps_dev);
  delete p;
  return 0;
}

extern "C" int sagan_sn_plan_phase_times(sagan_sn_plan* p, float* ms_host) {
  SAGAN_REQUIRE(p && ms_host, "sagan_sn_plan_phase_times: null argument");
  unsigned long long st[8];
  SAGAN_CUDA(cudaMemcpy(st, p->stamps_dev, sizeof(st), cudaMemcpyDeviceToHost));   // synchronises
  for (int i = 0; i < 5; ++i) ms_host[i] = st[i + 1] >= st[i] ? (float)((double)(st[i + 1] - st[i]) * 1e-6) : 0.f;
  return 0;
}

extern "C" unsigned long long sagan_sn_plan_algorithmic_bytes(const sagan_sn_plan* p) { return p ? p->alg_bytes : 0ull; }

extern "C" size_t sagan_sn_backward_workspace_bytes(long long numel) {
  (void)numel;
  return SNB_MAX_BLOCKS * sizeof(float);
}

extern "C" int sagan_sn_backward(const float* dW_bar, const float* W_bar, const float* u, const float* v,
                                 const float* sigma, float factor, float* dW, int rows, int cols, void* ws,
                                 size_t ws_bytes, sagan_stream_t stream) {
  SAGAN_REQUIRE(dW_bar && W_bar && u && v && sigma && dW && ws, "sagan_sn_backward: null pointer");
  SAGAN_REQUIRE(rows >= 1 && cols >= 1, "sagan_sn_backward: bad shape [%d,%d]", rows, cols);
  if (ws_bytes < SNB_MAX_BLOCKS * sizeof(float)) {
    set_err("sagan_sn_backward: workspace %zu < %zu bytes", ws_bytes, SNB_MAX_BLOCKS * sizeof(float));
    return SAGAN_EWORKSPACE;
  }
  const long long n = (long long)rows * cols;
  const int blocks = (int)std::min<long long>(SNB_MAX_BLOCKS, ceil_div<long long>(n, SNB_THREADS * 4));
  sn_bwd_dot_kernel<<<blocks, SNB_THREADS, 0, (cudaStream_t)stream>>>(dW_bar, W_bar, n, (float*)ws);
  SAGAN_LAUNCH_CHECK();
  sn_bwd_apply_kernel<<<blocks, SNB_THREADS, 0, (cudaStream_t)stream>>>(dW_bar, u, v, sigma, factor, (const float*)ws,
                                                                         blocks, dW, rows, cols);
  SAGAN_LAUNCH_CHECK();
  return 0;
}
