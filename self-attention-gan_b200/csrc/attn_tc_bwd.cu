// Self-attention block, BF16_TC math mode: flash-style BACKWARD on the 5th-gen tensor cores.
//
// The reference has no hand-written backward (tf.GradientTape differentiates /root/reference/layers.py:93-120);
// the formulas are SURVEY.md §8a row 2:  dP = dA g^T, dS = P * (dP - D), D = rowsum(dA * A),
// dg = P^T dA, dtheta = dS phi, dphi = dS^T theta.  P is recomputed from the saved row log-sum-exp.
//
// CTA = one 128-KEY tile j of one sample, looping over 128-query tiles i; TMEM lanes = keys, so the two
// score-shaped MMAs are computed transposed and every accumulation over queries stays inside the CTA:
//     S^T  = K_j Q_i^T           (M=keys, N=queries, K=16*kq)        [split-bf16 logits, log2 units]
//     dP^T = V_j dA_i^T          (M=keys, N=queries, K=16*kv)
//     P'^T = bf16(exp2(S^T - M_i)), dS^T = P'^T * (dP'^T - D'_i)     -> bf16 tiles [keys][queries] in smem (SW128)
//            M_i = ceil(lse'_i) integer, so P' reproduces the forward's bf16 weights exactly (attn_tc.cu, "consistent
//            rounding"); the per-query factor 2^(M_i - lse'_i) that normalises P' is folded into dA and D by the prep
//            kernel (dA' = f dA, D' = f D), so sum_j dS_ij = 0 holds to fp32 accuracy and nothing is amplified
//     dV_j += P^T  dA_i          (A = P^T  K-major, B = dA_i^T rows)   TMEM, accumulated over i
//     dK_j += dS^T Q_i           (A = dS^T K-major, B = Q_i^T rows)    TMEM, accumulated over i
//     dQ_i  = dS   K_j           (A = the SAME dS^T tile read MN-major, B = K_j^T rows) -> atomicAdd over key tiles
// 19 warps: 0-15 compute (four threads per key row, 32 query columns each), warp 16 = TMA producer,
// warps 17, 18 = MMA issuers (S^T / dP^T / dV / dK and dQ).  Compute warps never wait for each other: they signal the issuer through mbarriers
// (S/dP registers loaded -> next S/dP MMA may overwrite TMEM; P/dS tiles written -> gradient MMAs may start).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "attn_pool.cuh"

namespace sagan {

using namespace tc;

constexpr float TB_LOG2E = 1.4426950408889634f;
constexpr int TB_THREADS = 608;   // 16 compute warps + TMA producer warp + 2 MMA issuer warps
constexpr int TB_CWARPS = 16;
constexpr int TB_COLS = 64;   // widest bf16 row of the K-major operand buffers (128 B)
// Row lengths of the score-MMA operands: exactly the K extent the MMAs read (16 / 32 / 64 bf16 = 32 / 64 / 128 bytes,
// loaded in the matching TMA swizzle mode).  Every CTA re-reads all query tiles of its sample, so padded rows would be
// L2 -> SM traffic and shared memory spent on zeros.
__host__ __device__ constexpr int tb_kq(int C) { return ((3 * (C / 8) + 15) / 16) * 16; }
__host__ __device__ constexpr int tb_kv(int C) { return C <= 32 ? ((3 * (C / 2) + 2 + 15) / 16) * 16 : C; }

// ------------------------------------------------------------------------------------ dS as ONE fp16 term (C <= 32)
// dS = P' (dP - D) feeds two GEMMs that cancel (sum_j dS_ij = 0): bf16 (8 significant bits) needed a hi + lo pair of
// tiles -- an unpack, a subtraction, a second pack, a TMEM store, two shared-memory stores per pair of scores and 16
// more MMAs per tile (2.5 of the 15 warp instructions per score).  fp16 carries 11 significant bits in ONE term and
// measures the same gradient errors (bound by the bf16 rounding of P' either way); kind::f16 does not mix an fp16 A with
// a bf16 B (illegal instruction), so the transposed operands Q^T / K^T of those GEMMs are fp16 hi / lo pairs as well.
// fp16 has 5 exponent bits: the gradient is normalised per sample by a power of two s_b taken from max |dY| (one small
// reduction kernel), applied exactly to dA' and D' by the prep kernel and taken out of dQ / dK in the epilogues.
__device__ __forceinline__ __nv_bfloat16 tb_f16_bits(float x) {          // fp16 bits through a bf16-typed pointer
  const __half h = __float2half_rn(x);
  return *reinterpret_cast<const __nv_bfloat16*>(&h);
}
__device__ __forceinline__ float tb_f16_round(float x) { return __half2float(__float2half_rn(x)); }
template <bool F16>
__device__ __forceinline__ __nv_bfloat16 tb_t_hi(float x) { return F16 ? tb_f16_bits(x) : __float2bfloat16_rn(x); }
template <bool F16>
__device__ __forceinline__ __nv_bfloat16 tb_t_lo(float x) {
  return F16 ? tb_f16_bits(x - tb_f16_round(x)) : __float2bfloat16_rn(x - __bfloat162float(__float2bfloat16_rn(x)));
}

// gmax[b] = max |dY[b, :, :]| (as the bits of a non-negative float: atomicMax on unsigned keeps the order); zeroed by the caller
__global__ void __launch_bounds__(256)
attn_dy_absmax_kernel(const float* __restrict__ dY, unsigned* __restrict__ gmax, long long per_sample) {
  const int b = blockIdx.y;
  const float* p = dY + (size_t)b * per_sample;
  float m = 0.f;
  const long long n4 = per_sample >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = ld4(p + 4 * i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < per_sample; i += (long long)gridDim.x * 256)
    m = fmaxf(m, fabsf(p[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f && m <= 3.0e38f) atomicMax(gmax + b, __float_as_uint(m));
}
// s_b = 2^(5 - floor(log2(max |gamma dY|))): the scaled gradient peaks in [32, 64) (1 for an all-zero / non-finite sample)
__device__ __forceinline__ float tb_grad_scale(unsigned gmax_bits, float gamma) {
  const float m = __uint_as_float(gmax_bits) * fabsf(gamma);
  if (!(m > 0.f) || !(m <= 3.0e38f)) return 1.0f;
  int e = (int)((__float_as_uint(m) >> 23) & 255u) - 127;      // floor(log2 m) for normal m (denormal: -127)
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  return __uint_as_float((unsigned)(127 + 5 - e) << 23);
}

// ------------------------------------------------------------------------------------ prep (small C)
// one thread per PADDED token; recomputes theta/phi/g from X and forms dA = gamma dY Wo^T, D = dA . A
// KSIDE = false (down-sampled keys / values): the key-side rows (Kb, Kt, Vb) come from attn_bwd_prep_pool_tc_kernel.
template <int C, bool KSIDE>
__global__ void __launch_bounds__(128)
attn_bwd_prep_tc_kernel(const float* __restrict__ X, const float* __restrict__ dY, const float* __restrict__ A,
                        const float* __restrict__ lse, const float* __restrict__ Wq, const float* __restrict__ bq,
                        const float* __restrict__ Wk, const float* __restrict__ bk, const float* __restrict__ Wv,
                        const float* __restrict__ bv, const float* __restrict__ Wo, const float* __restrict__ gamma,
                        __nv_bfloat16* __restrict__ Qb, __nv_bfloat16* __restrict__ Kb, __nv_bfloat16* __restrict__ Vb,
                        __nv_bfloat16* __restrict__ dAb, __nv_bfloat16* __restrict__ dAt, __nv_bfloat16* __restrict__ Qt,
                        __nv_bfloat16* __restrict__ Kt, float* __restrict__ lse2, float* __restrict__ Dd,
                        const unsigned* __restrict__ gmax, float* __restrict__ sinv, int B, int N, int Npad) {
  constexpr int D = C / 8, DV = C / 2;
  constexpr bool F16 = C <= 32;                      // dS and the transposed operands of its GEMMs in fp16 (see above)
  constexpr bool SPLIT3 = DV <= 16;                  // 3-term split of the dP contraction fits a 64-column row
  constexpr bool SPLIT_DA = DV <= 16;                // dA^T rows [hi | lo] for the dV GEMM
  constexpr int DVP = SPLIT_DA ? (2 * DV < 16 ? 16 : 2 * DV) : DV;
  constexpr int KQ = ((3 * D + 15) / 16) * 16;
  // FOLD (C <= 32): the per-query shift M_i of the logits, the mask of padded keys and the row term D_i are folded
  // into spare K columns of the two score-shaped MMAs, which then deliver S - M_i and dP - D_i directly:
  //   Q row gets [-M_hi, -16384, -M_lo] against K row [1, key padded ? 1 : 0, 1]     (all bf16-exact)
  //   dA row gets [-D_hi, -D_lo]        against V row [1, 1]
  // -> one exp2 and one multiply per score are left on the CUDA cores (no subtractions, no mask select, no row vectors)
  constexpr bool FOLD = SPLIT3;
  constexpr int KV = SPLIT3 ? ((3 * DV + 2 + 15) / 16) * 16 : 2 * DV;
  static_assert(KV <= TB_COLS, "dP operand row must fit one 128-byte swizzle span");
  static_assert(KQ == tb_kq(C) && KV == tb_kv(C), "row lengths shared with the main kernel");
  static_assert(!FOLD || 3 * D + 3 <= KQ, "no spare logit columns for the folded shift");
  __shared__ float sWq[C * D], sWk[C * D], sWv[C * DV], sbq[D], sbk[D], sbv[DV];
  __shared__ __align__(16) float sWo[DV * C];
  for (int i = threadIdx.x; i < C * D; i += 128) { sWq[i] = Wq[i]; sWk[i] = Wk[i]; }
  for (int i = threadIdx.x; i < C * DV; i += 128) { sWv[i] = Wv[i]; sWo[i] = Wo[i]; }
  for (int i = threadIdx.x; i < D; i += 128) { sbq[i] = bq[i]; sbk[i] = bk[i]; }
  for (int i = threadIdx.x; i < DV; i += 128) sbv[i] = bv[i];
  __syncthreads();
  const long long tp = (long long)blockIdx.x * 128 + threadIdx.x;
  if (tp >= (long long)B * Npad) return;
  const int b = (int)(tp / Npad), n = (int)(tp - (long long)b * Npad);
  const bool valid = n < N;
  const long long t = (long long)b * N + (valid ? n : 0);
  float x[C], dy[C];
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = v;
    if (valid) { v = ld4(X + t * C + c); g = ld4(dY + t * C + c); }
    x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
    dy[c] = g.x; dy[c + 1] = g.y; dy[c + 2] = g.z; dy[c + 3] = g.w;
  }
  const float l2 = valid ? lse[t] * TB_LOG2E : 0.f;
  const float Mi = ceilf(l2);
  // ---- theta / phi: split-bf16 rows for the logits, plain hi/lo transposed rows for the dK / dQ GEMMs
  float q[KQ], k[KQ];
#pragma unroll
  for (int j = 0; j < KQ; ++j) { q[j] = 0.f; k[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < D; ++j) {
    float a = sbq[j], kk = sbk[j];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a = fmaf(x[c], sWq[c * D + j], a);
      kk = fmaf(x[c], sWk[c * D + j], kk);
    }
    a = valid ? a : 0.f;
    kk = valid ? kk : 0.f;
    // transposed (un-scaled) copies: rows [hi at 0..7 | lo at 8..15]
    const __nv_bfloat16 kh = __float2bfloat16_rn(kk);
    Qt[((long long)b * 16 + j) * Npad + n] = tb_t_hi<F16>(a);
    Qt[((long long)b * 16 + 8 + j) * Npad + n] = tb_t_lo<F16>(a);
    if (KSIDE) {
      Kt[((long long)b * 16 + j) * Npad + n] = tb_t_hi<F16>(kk);
      Kt[((long long)b * 16 + 8 + j) * Npad + n] = tb_t_lo<F16>(kk);
    }
    const float as = a * TB_LOG2E;
    const float a_hi = __bfloat162float(__float2bfloat16_rn(as)), k_hi = __bfloat162float(kh);
    q[j] = a_hi; q[D + j] = as - a_hi; q[2 * D + j] = a_hi;
    k[j] = k_hi; k[D + j] = k_hi;      k[2 * D + j] = kk - k_hi;
  }
  if (FOLD) {
    const float mq = valid ? Mi : 16384.f;
    const float m_hi = __bfloat162float(__float2bfloat16_rn(mq));
    q[3 * D] = -m_hi;          k[3 * D] = 1.f;
    q[3 * D + 1] = -16384.f;   k[3 * D + 1] = valid ? 0.f : 1.f;
    q[3 * D + 2] = -(mq - m_hi); k[3 * D + 2] = 1.f;
  }
#pragma unroll
  for (int j = D; j < 8; ++j) {      // rows [hi (0..7) | lo (8..15)]: unused rows are zero
    Qt[((long long)b * 16 + j) * Npad + n] = __float2bfloat16_rn(0.f);
    Qt[((long long)b * 16 + 8 + j) * Npad + n] = __float2bfloat16_rn(0.f);
    if (KSIDE) {
      Kt[((long long)b * 16 + j) * Npad + n] = __float2bfloat16_rn(0.f);
      Kt[((long long)b * 16 + 8 + j) * Npad + n] = __float2bfloat16_rn(0.f);
    }
  }
  uint4* qd = reinterpret_cast<uint4*>(Qb + tp * KQ);
  uint4* kd = reinterpret_cast<uint4*>(Kb + tp * KQ);
#pragma unroll
  for (int g = 0; g < KQ / 8; ++g) {
    qd[g] = make_uint4(pack_bf16x2(q[g * 8 + 0], q[g * 8 + 1]), pack_bf16x2(q[g * 8 + 2], q[g * 8 + 3]),
                       pack_bf16x2(q[g * 8 + 4], q[g * 8 + 5]), pack_bf16x2(q[g * 8 + 6], q[g * 8 + 7]));
    if (KSIDE)
      kd[g] = make_uint4(pack_bf16x2(k[g * 8 + 0], k[g * 8 + 1]), pack_bf16x2(k[g * 8 + 2], k[g * 8 + 3]),
                         pack_bf16x2(k[g * 8 + 4], k[g * 8 + 5]), pack_bf16x2(k[g * 8 + 6], k[g * 8 + 7]));
  }
  // ---- g (values) and dA' = f gamma dY Wo^T, D' = dA' . A, with f = 2^(M - lse') the normaliser of P' (see top)
  // (F16: times the sample's power-of-two gradient scale; dA^T for the dV GEMM stays un-scaled)
  const float sb = F16 ? tb_grad_scale(gmax[b], *gamma) : 1.0f;
  const float sb_inv = 1.0f / sb;
  if (F16 && n == 0) sinv[b] = sb_inv;
  const float gm = *gamma * exp2f(Mi - l2) * sb;
  float v[KV], da[KV];
#pragma unroll
  for (int j = 0; j < KV; ++j) { v[j] = 0.f; da[j] = 0.f; }
  float dd = 0.f;
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    float a = sbv[j], g = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a = fmaf(x[c], sWv[c * DV + j], a);
      g = fmaf(dy[c], sWo[j * C + c], g);
    }
    g *= gm;
    a = valid ? a : 0.f;
    g = valid ? g : 0.f;
    // split-bf16 operands of dP^T = V dA^T (dP - D cancels, so bf16 V / dA would dominate the gradient error):
    //   V row [v_hi | v_hi | v_lo], dA row [da_hi | da_lo | da_hi]  ->  v_hi da_hi + v_hi da_lo + v_lo da_hi
    // dv = 32: only two terms fit the 64-column row: [v_hi | v_lo] x [da_hi | da_hi] = v . bf16(dA), and D uses the
    // same rounded dA so that D_i = sum_j P'_ij dP_ij stays consistent
    const float vh = __bfloat162float(__float2bfloat16_rn(a)), gh = __bfloat162float(__float2bfloat16_rn(g));
    if (SPLIT3) {
      v[j] = vh; v[DV + j] = vh; v[2 * DV + j] = a - vh;
      da[j] = gh; da[DV + j] = g - gh; da[2 * DV + j] = gh;
      if (valid) dd = fmaf(g, A[t * DV + j], dd);
    } else {
      v[j] = vh; v[DV + j] = a - vh;
      da[j] = gh; da[DV + j] = gh;
      if (valid) dd = fmaf(gh, A[t * DV + j], dd);
    }
    // transposed dA rows for dV = P^T dA: [hi | lo] when they fit
    const float gu = g * sb_inv, guh = __bfloat162float(__float2bfloat16_rn(gu));
    dAt[((long long)b * DVP + j) * Npad + n] = __float2bfloat16_rn(gu);
    if (SPLIT_DA) dAt[((long long)b * DVP + DV + j) * Npad + n] = __float2bfloat16_rn(gu - guh);
  }
  if (FOLD) {
    const float d_hi = __bfloat162float(__float2bfloat16_rn(dd));
    v[3 * DV] = 1.f;      da[3 * DV] = -d_hi;
    v[3 * DV + 1] = 1.f;  da[3 * DV + 1] = -(dd - d_hi);
  }
  uint4* vd = reinterpret_cast<uint4*>(Vb + tp * KV);
  uint4* ad = reinterpret_cast<uint4*>(dAb + tp * KV);
#pragma unroll
  for (int g = 0; g < KV / 8; ++g) {
    if (KSIDE)
      vd[g] = make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                         pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
    ad[g] = make_uint4(pack_bf16x2(da[g * 8 + 0], da[g * 8 + 1]), pack_bf16x2(da[g * 8 + 2], da[g * 8 + 3]),
                       pack_bf16x2(da[g * 8 + 4], da[g * 8 + 5]), pack_bf16x2(da[g * 8 + 6], da[g * 8 + 7]));
  }
#pragma unroll
  for (int j = (SPLIT_DA ? 2 * DV : DV); j < DVP; ++j) dAt[((long long)b * DVP + j) * Npad + n] = __float2bfloat16_rn(0.f);
  lse2[tp] = valid ? Mi : INFINITY;                  // integer shift M_i; +inf => P' = 0 for padded queries
  Dd[tp] = valid ? dd : 0.f;
}

// ------------------------------------------------------------------------------------ prep, down-sampled keys / values
// one thread per PADDED pooled position: the key-side operand rows of the main kernel (Kb with the folded shift / mask
// columns, Kt, Vb with the folded ones columns) from the 2x2 max-pooled phi / g, plus the window position that supplied
// each channel (for the scatter of dK / dV back to the tokens).
template <int C>
__global__ void __launch_bounds__(128)
attn_bwd_prep_pool_tc_kernel(const float* __restrict__ X, const float* __restrict__ Wk, const float* __restrict__ bk,
                             const float* __restrict__ Wv, const float* __restrict__ bv, __nv_bfloat16* __restrict__ Kb,
                             __nv_bfloat16* __restrict__ Vb, __nv_bfloat16* __restrict__ Kt, uint8_t* __restrict__ idxK,
                             uint8_t* __restrict__ idxV, int B, int H, int W, int Nk, int Nkpad) {
  constexpr int D = C / 8, DV = C / 2;
  constexpr bool SPLIT3 = DV <= 16, FOLD = SPLIT3;
  constexpr int KQ = ((3 * D + 15) / 16) * 16;
  constexpr int KV = SPLIT3 ? ((3 * DV + 2 + 15) / 16) * 16 : 2 * DV;
  static_assert(KQ == tb_kq(C) && KV == tb_kv(C), "row lengths shared with the main kernel");
  __shared__ float sWk[C * D], sWv[C * DV], sbk[D], sbv[DV];
  for (int i = threadIdx.x; i < C * D; i += 128) sWk[i] = Wk[i];
  for (int i = threadIdx.x; i < C * DV; i += 128) sWv[i] = Wv[i];
  for (int i = threadIdx.x; i < D; i += 128) sbk[i] = bk[i];
  for (int i = threadIdx.x; i < DV; i += 128) sbv[i] = bv[i];
  __syncthreads();
  const long long tp = (long long)blockIdx.x * 128 + threadIdx.x;
  if (tp >= (long long)B * Nkpad) return;
  const int b = (int)(tp / Nkpad), n = (int)(tp - (long long)b * Nkpad);
  const bool valid = n < Nk;
  float kk[D], vv[DV];
  uint8_t ik[D], iv[DV];
#pragma unroll
  for (int j = 0; j < D; ++j) { kk[j] = 0.f; ik[j] = 0; }
#pragma unroll
  for (int j = 0; j < DV; ++j) { vv[j] = 0.f; iv[j] = 0; }
  if (valid) {
    pooled_kv<C>(X, sWk, sbk, sWv, sbv, b, H, W, n / (W / 2), n % (W / 2), kk, vv, ik, iv);
    const long long p = (long long)b * Nk + n;
#pragma unroll
    for (int j = 0; j < D; ++j) idxK[p * D + j] = ik[j];
#pragma unroll
    for (int j = 0; j < DV; ++j) idxV[p * DV + j] = iv[j];
  }
  float k[KQ], v[KV];
#pragma unroll
  for (int j = 0; j < KQ; ++j) k[j] = 0.f;
#pragma unroll
  for (int j = 0; j < KV; ++j) v[j] = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const __nv_bfloat16 kh = __float2bfloat16_rn(kk[j]);
    const float k_hi = __bfloat162float(kh);
    Kt[((long long)b * 16 + j) * Nkpad + n] = tb_t_hi<(C <= 32)>(kk[j]);
    Kt[((long long)b * 16 + 8 + j) * Nkpad + n] = tb_t_lo<(C <= 32)>(kk[j]);
    k[j] = k_hi; k[D + j] = k_hi; k[2 * D + j] = kk[j] - k_hi;
  }
#pragma unroll
  for (int j = D; j < 8; ++j) {
    Kt[((long long)b * 16 + j) * Nkpad + n] = __float2bfloat16_rn(0.f);
    Kt[((long long)b * 16 + 8 + j) * Nkpad + n] = __float2bfloat16_rn(0.f);
  }
  if (FOLD) { k[3 * D] = 1.f; k[3 * D + 1] = valid ? 0.f : 1.f; k[3 * D + 2] = 1.f; }
#pragma unroll
  for (int j = 0; j < DV; ++j) {
    const float vh = __bfloat162float(__float2bfloat16_rn(vv[j]));
    if (SPLIT3) { v[j] = vh; v[DV + j] = vh; v[2 * DV + j] = vv[j] - vh; }
    else { v[j] = vh; v[DV + j] = vv[j] - vh; }
  }
  if (FOLD) { v[3 * DV] = 1.f; v[3 * DV + 1] = 1.f; }
  uint4* kd = reinterpret_cast<uint4*>(Kb + tp * KQ);
#pragma unroll
  for (int g = 0; g < KQ / 8; ++g)
    kd[g] = make_uint4(pack_bf16x2(k[g * 8 + 0], k[g * 8 + 1]), pack_bf16x2(k[g * 8 + 2], k[g * 8 + 3]),
                       pack_bf16x2(k[g * 8 + 4], k[g * 8 + 5]), pack_bf16x2(k[g * 8 + 6], k[g * 8 + 7]));
  uint4* vd = reinterpret_cast<uint4*>(Vb + tp * KV);
#pragma unroll
  for (int g = 0; g < KV / 8; ++g)
    vd[g] = make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                       pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

// ------------------------------------------------------------------------------------ main
template <int DVP, int QKB, int VAB, int NDS>
struct BwdSmem {
  static constexpr int TILE = 128 * 128;                 // [128 rows][128 B]
  static constexpr int T16 = 2 * 16 * 128;               // two 64-column sub-tiles of [16 rows][128 B]
  static constexpr int TDV = 2 * DVP * 128;
  static constexpr int QK_TILE = 128 * QKB;              // Q / K rows of QKB bytes
  static constexpr int VA_TILE = 128 * VAB;              // V / dA rows of VAB bytes
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + QK_TILE;
  static constexpr int OFF_KT = OFF_V + VA_TILE;
  // query-side operands in two rings: A = what the score-shaped MMAs read (Q_i, dA_i rows; free as soon as dP^T_i has
  // been computed, so three stages let the loads run two tiles ahead), B = what the gradient MMAs read (dA_i^T, Q_i^T
  // rows and the lse / D vectors; free when the gradient MMAs of the tile are done)
  static constexpr int NA = 3, NB = 2;
  static constexpr int OFF_A = OFF_KT + T16;
  static constexpr int A_Q = 0, A_DA = QK_TILE, ASTAGE = QK_TILE + VA_TILE;
  static constexpr int OFF_B = OFF_A + NA * ASTAGE;
  static constexpr int B_DAT = 0, B_QT = TDV, B_VEC = TDV + T16, BSTAGE = TDV + T16 + 1024;   // vec: lse2[128], D[128]
  // dS^T tiles for the dQ MMAs (bf16 hi / lo terms), NDS buffers: with two, the stores of tile i do not wait for dQ_{i-1}
  static constexpr int DS_BUF = 4 * TILE;                 // [hi: 2 sub-tiles | lo: 2 sub-tiles]
  static constexpr int OFF_DS = OFF_B + NB * BSTAGE;
  static constexpr int OFF_BAR = OFF_DS + NDS * DS_BUF;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static constexpr int A_TX = QK_TILE + VA_TILE, B_TX = TDV + T16 + 1024;
  static_assert(ASTAGE % 1024 == 0 && BSTAGE % 1024 == 0 && OFF_A % 1024 == 0, "tiles must stay 1024-byte aligned");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
  // TMEM columns.  P'^T and the two bf16 terms of dS^T are the A operands of the dV / dK MMAs and live in TMEM
  // (tcgen05.mma with A from TMEM): the thread that owns dP^T columns [32 h, 32 h + 32) of a key row overwrites them
  // with its 32 P'^T (16 packed columns) and 32 dS^T_lo values once it holds dP^T in registers; dS^T_hi has its own
  // columns.  Small-N MMAs that accumulate into the SAME TMEM tile execute back to back at the full pipeline latency,
  // so the gradient GEMMs are spread over five independent accumulators issued round-robin: dV, dK from dS_hi, dK from
  // dS_lo, dQ from dS_hi, dQ from dS_lo (the hi / lo partial sums are added in the epilogue)
  static constexpr int ST_COL = 0, DP_COL = 128, DSH_COL = 256, DV_COL = 320, DKH_COL = 352, DKL_COL = 368,
                       DQ_COL = 384;   // DQ: 2 x [hi 16 | lo 16]
};

// D[tmem] (+)= A[tmem] * B[smem]^T : A is a K-major bf16 tile in TMEM (lane = row, two K elements per 32-bit column)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  mma_bf16_ts_g<1>(d_tmem, a_tmem, b_desc, idesc, accumulate);
}

template <int DVP, bool SPLIT_DA, int QKB, int VAB, int NDS>
__global__ void __launch_bounds__(TB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdA,
                   const __grid_constant__ CUtensorMap tmdAt, const __grid_constant__ CUtensorMap tmQt,
                   const __grid_constant__ CUtensorMap tmKt, const float* __restrict__ lse2,
                   const float* __restrict__ Dd, const float* __restrict__ sinv, float* __restrict__ dQ,
                   float* __restrict__ dK, float* __restrict__ dV, int N, int Npad, int Nk, int Nkpad, int d, int dv,
                   int kq_steps, int kv_steps) {
  constexpr bool F16 = SPLIT_DA;       // dS as one fp16 term, scaled by the sample's s_b (prep kernel): no lo tiles, no lo MMAs
  // N queries (Npad padded), Nk keys / values (Nkpad padded; grid.x = Nkpad / 128); Nk == N unless down-sampled
  using L = BwdSmem<DVP, QKB, VAB, NDS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem + L::OFF_K;
  uint8_t* sV = smem + L::OFF_V;
  uint8_t* sKt = smem + L::OFF_KT;
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sB = smem + L::OFF_B;
  uint8_t* sDS = smem + L::OFF_DS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* barKV = bars + 0;
  uint64_t* barAfull = bars + 1;     // [3] Q_i / dA_i rows landed
  uint64_t* barAfree = bars + 4;     // [3] S^T_i and dP^T_i computed: the A stage may be overwritten
  uint64_t* barBfull = bars + 7;     // [2] dA_i^T / Q_i^T rows and lse / D landed
  uint64_t* barBfree = bars + 9;     // [2] dV / dK MMAs of tile i done: the B stage may be overwritten
  uint64_t* barQd = bars + 11;       // [2] dQ_i computed: the shared-memory dS^T tiles are free, dQ_i may be flushed
  // The 16 compute warps work as two HALVES (query columns [0,64) / [64,128) of the tile, 8 warps each) with their own
  // barrier chains.  All warps of one chain run the phases MUFU (exp2) -> ALU (dS, bf16 splits) -> LSU (stores) in
  // lockstep; two chains served alternately by the MMA issuer settle out of phase, so the three pipes overlap.
  uint64_t* barS = bars + 13;        // [half] S^T columns of the half ready
  uint64_t* barDP = bars + 15;       // [half] dP^T columns ready
  uint64_t* barSfree = bars + 17;    // [half] 256 arrivals: S^T of tile i is in registers
  uint64_t* barTiles = bars + 19;    // [half] 256 arrivals: P'^T / dS^T of tile i are in TMEM and shared memory
  uint64_t* barGh = bars + 21;       // [half] dK-from-dS_hi MMAs of tile i done: the dS^T_hi columns may be overwritten
  uint64_t* barDone = bars + 23;     // both issuers: every MMA of the CTA has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, kt = blockIdx.x;
  const int nq = Npad / 128;

  if (threadIdx.x == 0) {
    mbar_init(barKV, 1);
    for (int i = 0; i < L::NA; ++i) { mbar_init(barAfull + i, 1); mbar_init(barAfree + i, 1); }
    for (int i = 0; i < L::NB; ++i) { mbar_init(barBfull + i, 1); mbar_init(barBfree + i, 1); mbar_init(barQd + i, 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(barS + x, 1); mbar_init(barDP + x, 1); mbar_init(barGh + x, 1);
      mbar_init(barSfree + x, TB_CWARPS * 16); mbar_init(barTiles + x, TB_CWARPS * 16);
    }
    mbar_init(barDone, 2);
    mbar_fence_init();
  }
  if (warp == TB_CWARPS) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == TB_CWARPS) {
    // ================================================================ TMA producer
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdA);
      tma_prefetch_desc(&tmdAt); tma_prefetch_desc(&tmQt); tma_prefetch_desc(&tmKt);
      mbar_expect_tx(barKV, L::QK_TILE + L::VA_TILE + L::T16);
      tma_load_2d(sK, &tmK, barKV, 0, b * Nkpad + kt * 128);
      tma_load_2d(sV, &tmV, barKV, 0, b * Nkpad + kt * 128);
      tma_load_2d(sKt, &tmKt, barKV, kt * 128, b * 16);
      tma_load_2d(sKt + 16 * 128, &tmKt, barKV, kt * 128 + 64, b * 16);
      // the two rings are refilled by polling, whichever stage frees first (no assumption on their relative order)
      int ia = 0, ib = 0;
      while (ia < nq || ib < nq) {
        const int before = ia + ib;
        if (ia < nq) {
          const int s = ia % L::NA;
          if (ia < L::NA || mbar_try_wait(barAfree + s, ((ia / L::NA) - 1) & 1)) {
            uint8_t* st = sA + s * L::ASTAGE;
            mbar_expect_tx(barAfull + s, L::A_TX);
            tma_load_2d(st + L::A_Q, &tmQ, barAfull + s, 0, b * Npad + ia * 128);
            tma_load_2d(st + L::A_DA, &tmdA, barAfull + s, 0, b * Npad + ia * 128);
            ++ia;
          }
        }
        if (ib < nq) {
          const int s = ib % L::NB;
          if (ib < L::NB || mbar_try_wait(barBfree + s, ((ib / L::NB) - 1) & 1)) {
            uint8_t* st = sB + s * L::BSTAGE;
            mbar_expect_tx(barBfull + s, L::B_TX);
            tma_load_2d(st + L::B_DAT, &tmdAt, barBfull + s, ib * 128, b * DVP);
            tma_load_2d(st + L::B_DAT + DVP * 128, &tmdAt, barBfull + s, ib * 128 + 64, b * DVP);
            tma_load_2d(st + L::B_QT, &tmQt, barBfull + s, ib * 128, b * 16);
            tma_load_2d(st + L::B_QT + 16 * 128, &tmQt, barBfull + s, ib * 128 + 64, b * 16);
            bulk_load_1d(st + L::B_VEC, lse2 + (size_t)b * Npad + ib * 128, 512, barBfull + s);
            bulk_load_1d(st + L::B_VEC + 512, Dd + (size_t)b * Npad + ib * 128, 512, barBfull + s);
            ++ib;
          }
        }
        if (ia + ib == before) __nanosleep(256);      // nothing freed: do not hog the scheduler's issue slots
      }
    }
  } else if (warp == TB_CWARPS + 1) {
    // ================================================================ MMA issuer
    if (elect_one_sync()) {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, 64);
      constexpr uint32_t IDESC_DV = make_idesc_bf16(128, DVP);
      constexpr uint32_t F16_FMT = F16 ? ~((7u << 7) | (7u << 10)) : ~0u;       // A, B format 0 = fp16
      constexpr uint32_t IDESC_DK = make_idesc_bf16(128, 16) & F16_FMT;
      const uint64_t dK_ = make_desc_rows<QKB>(smem_u32(sK)), dV_ = make_desc_rows<VAB>(smem_u32(sV));
      auto issue_s = [&](int i, int x) {       // S^T[:, 64 x .. 64 x + 64) = K Q_i^T (64 query rows of the stage)
        const uint64_t dQ_ = make_desc_rows<QKB>(smem_u32(sA + (i % L::NA) * L::ASTAGE + L::A_Q + x * (64 * QKB)));
        for (int ks = 0; ks < kq_steps; ++ks)
          mma_bf16_ss(tmem_base + L::ST_COL + x * 64, dK_ + (uint64_t)(ks * 2), dQ_ + (uint64_t)(ks * 2), IDESC_S, ks > 0);
        mma_commit(barS + x);
      };
      auto issue_dp = [&](int i, int x) {      // dP^T[:, 64 x .. 64 x + 64) = V dA_i^T
        const uint64_t dA_ = make_desc_rows<VAB>(smem_u32(sA + (i % L::NA) * L::ASTAGE + L::A_DA + x * (64 * VAB)));
        for (int ks = 0; ks < kv_steps; ++ks)
          mma_bf16_ss(tmem_base + L::DP_COL + x * 64, dV_ + (uint64_t)(ks * 2), dA_ + (uint64_t)(ks * 2), IDESC_S, ks > 0);
        mma_commit(barDP + x);
        if (x == 1) mma_commit(barAfree + (i % L::NA));     // last reader of the A stage
      };
      mbar_wait(barKV, 0);
      mbar_wait(barAfull, 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_dp(0, 0);
      mbar_wait(barSfree, 0);                  // start the second half a little later: the two chains should not run
      tc_fence_after();                        // their MUFU / ALU / LSU phases at the same time
      issue_s(0, 1);
      issue_dp(0, 1);
      for (int i = 0; i < nq; ++i) {
        const uint8_t* st = sB + (i & 1) * L::BSTAGE;
        const uint64_t dAt_ = make_desc_sw128(smem_u32(st + L::B_DAT)), dQt_ = make_desc_sw128(smem_u32(st + L::B_QT));
        const bool acc0 = i > 0;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (i + 1 < nq) {
            if (x == 0) mbar_wait(barAfull + ((i + 1) % L::NA), ((i + 1) / L::NA) & 1);
            mbar_wait(barSfree + x, i & 1);          // every compute thread of the half holds S^T_i in registers
            tc_fence_after();
            issue_s(i + 1, x);                       // runs under the exp / dS math of tile i
          }
          if (x == 0) mbar_wait(barBfull + (i & 1), (i >> 1) & 1);
          mbar_wait(barTiles + x, i & 1);            // P'^T_i / dS^T_i are in TMEM (and every thread holds dP^T_i)
          tc_fence_after();
          // dV += P'^T dA_i ; dK += dS^T Q_i : A operands from TMEM, 16 queries per step.  The tensor pipe executes in
          // issue order, so dP^T_{i+1} (issued right after the MMAs that read the aliased P'^T / dS^T_lo columns)
          // cannot overwrite them early; the dS^T_hi MMAs follow it.
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = x * 4 + k4;
            const uint32_t a_pt = tmem_base + L::DP_COL + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8);
            const uint64_t b16 = (uint64_t)((ks >> 2) * ((16 * 128) >> 4) + (ks & 3) * 2);
            mma_bf16_ts(tmem_base + L::DV_COL, a_pt, dAt_ + (uint64_t)((ks >> 2) * ((DVP * 128) >> 4) + (ks & 3) * 2),
                        IDESC_DV, acc0 || (ks > 0));
            if (!F16) mma_bf16_ts(tmem_base + L::DKL_COL, a_pt + 16, dQt_ + b16, IDESC_DK, acc0 || (ks > 0));
          }
          if (i + 1 < nq) issue_dp(i + 1, x);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = x * 4 + k4;
            const uint32_t a_hi = tmem_base + L::DSH_COL + (uint32_t)(ks * 8);
            const uint64_t b16 = (uint64_t)((ks >> 2) * ((16 * 128) >> 4) + (ks & 3) * 2);
            mma_bf16_ts(tmem_base + L::DKH_COL, a_hi, dQt_ + b16, IDESC_DK, acc0 || (ks > 0));
          }
          mma_commit(barGh + x);
          if (x == 1) mma_commit(barBfree + (i & 1));
        }
      }
      mma_commit(barDone);
    }
  } else if (warp == TB_CWARPS + 2) {
    // ================================================================ second MMA issuer: dQ_i = dS_i K_j
    // (issuing one tcgen05.mma costs the elected thread ~35 cycles; 43 per tile from one thread were the critical path)
    if (elect_one_sync()) {
      constexpr uint32_t IDESC_DQ = make_idesc_bf16(128, 16, /*a_mn_major=*/1, 0) & (F16 ? ~((7u << 7) | (7u << 10)) : ~0u);
      const uint64_t dKt_ = make_desc_sw128(smem_u32(sKt));
      mbar_wait(barKV, 0);
      for (int i = 0; i < nq; ++i) {
        mbar_wait(barTiles, i & 1);
        mbar_wait(barTiles + 1, i & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // K = 16 keys per step: the dS^T tiles read MN-major (M = queries contiguous)
          const uint64_t b16 = (uint64_t)((ks >> 2) * ((16 * 128) >> 4) + (ks & 3) * 2);
          const uint32_t ds = smem_u32(sDS + (i % NDS) * L::DS_BUF);
          mma_bf16_ss(tmem_base + L::DQ_COL + (i & 1) * 32, make_desc_sw128_mn(ds + ks * 16 * 128, L::TILE, 1024),
                      dKt_ + b16, IDESC_DQ, ks > 0);
          if (!F16)
            mma_bf16_ss(tmem_base + L::DQ_COL + (i & 1) * 32 + 16, make_desc_sw128_mn(ds + 2 * L::TILE + ks * 16 * 128, L::TILE, 1024),
                        dKt_ + b16, IDESC_DQ, ks > 0);
        }
        mma_commit(barQd + (i & 1));
      }
      mma_commit(barDone);
    }
  } else {
    // ================================================================ compute warps
    // thread <-> key row krow (TMEM lane) x 32 of the tile's 128 query columns: warp w owns lane quarter w & 3 and
    // query columns [32 h, 32 h + 32), h = w >> 2.  Four warps per scheduler hide the MUFU / LDS latencies of the chain
    // exp2 -> bf16 -> (dP - D) -> split, which two warps per scheduler could not (measured 0.4 IPC).
    const int qd = warp & 3, h = warp >> 2, x = h >> 1;
    const int krow = qd * 32 + lane;                                // key row inside the tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    const bool key_ok = kt * 128 + krow < Nk;
    const float unscale = F16 ? sinv[b] : 1.0f;              // 1 / s_b: dQ and dK come out of the MMAs scaled
    // dQ_i tile (TMEM lanes = queries) -> atomicAdd; the four column groups take turns so the atomics are spread evenly
    auto flush_dq = [&](int i) {
      if (h == (i & 3)) {
        uint32_t r[32];
        if (F16) {                                            // [K_hi (8) | K_lo (8)] from the one fp16 dS term
          tmem_ld16(t_row + L::DQ_COL + (i & 1) * 32, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
#pragma unroll
          for (int c = 16; c < 32; ++c) r[c] = 0u;
        } else {
          tmem_ld32(t_row + L::DQ_COL + (i & 1) * 32, r);    // [from dS_hi: K_hi (8) K_lo (8) | from dS_lo: K_hi K_lo]
        }
        tmem_wait_ld();
        const int qrow = i * 128 + krow;
        if (qrow < N) {
          float* dst = dQ + ((size_t)b * N + qrow) * d;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < d)
              atomicAdd(dst + c, ((__uint_as_float(r[c]) + __uint_as_float(r[8 + c])) +
                                  (__uint_as_float(r[16 + c]) + __uint_as_float(r[24 + c]))) * unscale);
        }
      }
    };

    for (int i = 0; i < nq; ++i) {
      const int s = i & 1;
      mbar_wait(barS + x, i & 1);
      tc_fence_after();
      uint32_t pp[16];
      float pf[32];       // P' (bf16-rounded, as fp32) -- the weights the forward used
      {
        uint32_t rs[32];
        tmem_ld32(t_row + L::ST_COL + h * 32, rs);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(barSfree + x);                              // this thread's S^T is in registers
        if (SPLIT_DA) {   // folded operands: the MMA delivered S - M_i (padded keys at -16384)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t pk = pack_bf16x2(ex2_approx(__uint_as_float(rs[2 * e])), ex2_approx(__uint_as_float(rs[2 * e + 1])));
            pp[e] = pk;
            pf[2 * e] = __uint_as_float(pk << 16);
            pf[2 * e + 1] = __uint_as_float(pk & 0xffff0000u);
          }
        } else {
          mbar_wait(barBfull + s, (i >> 1) & 1);                // lse2 / D of this query tile have landed
          const float* vec = reinterpret_cast<const float*>(sB + s * L::BSTAGE + L::B_VEC) + h * 32;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float2 ls = *reinterpret_cast<const float2*>(vec + 2 * e);
            const float p0 = key_ok ? ex2_approx(__uint_as_float(rs[2 * e]) - ls.x) : 0.f;
            const float p1 = key_ok ? ex2_approx(__uint_as_float(rs[2 * e + 1]) - ls.y) : 0.f;
            const uint32_t pk = pack_bf16x2(p0, p1);
            pp[e] = pk;
            pf[2 * e] = __uint_as_float(pk << 16);
            pf[2 * e + 1] = __uint_as_float(pk & 0xffff0000u);
          }
        }
      }
      // ---- dS^T = P' (dP - D) = hi + lo (two bf16 terms): sum_j dS_ij = 0, so the theta / phi gradients cancel and
      //      need the extra bits
      uint32_t hi[16], lo[16];
      {
        mbar_wait(barDP + x, i & 1);
        tc_fence_after();
        uint32_t rp[32];
        tmem_ld32(t_row + L::DP_COL + h * 32, rp);
        tmem_wait_ld();
        const float* dvec = reinterpret_cast<const float*>(sB + s * L::BSTAGE + L::B_VEC) + 128 + h * 32;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float g0, g1;
          if (SPLIT_DA) {          // folded: the MMA delivered dP - D_i
            g0 = pf[2 * e] * __uint_as_float(rp[2 * e]);
            g1 = pf[2 * e + 1] * __uint_as_float(rp[2 * e + 1]);
          } else {
            const float2 dd = *reinterpret_cast<const float2*>(dvec + 2 * e);
            g0 = pf[2 * e] * (__uint_as_float(rp[2 * e]) - dd.x);
            g1 = pf[2 * e + 1] * (__uint_as_float(rp[2 * e + 1]) - dd.y);
          }
          if (F16) {
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi[e]) : "f"(g1), "f"(g0));      // first source -> upper half
            lo[e] = 0u;
          } else {
            const uint32_t hk = pack_bf16x2(g0, g1);
            hi[e] = hk;
            lo[e] = pack_bf16x2(g0 - __uint_as_float(hk << 16), g1 - __uint_as_float(hk & 0xffff0000u));
          }
        }
      }
      // ---- A operands of the dV / dK MMAs -> TMEM.  P'^T and dS^T_lo go over this thread's own dP^T columns, which
      //      it has just loaded: dP^T_i complete implies that the MMAs of tile i-1 that read those columns (issued
      //      before it) are complete.  dS^T_hi waits for the MMAs issued after dP^T_i, the shared-memory dS^T tiles
      //      for the dQ MMAs of tile i-1.
      tmem_st16(t_row + L::DP_COL + h * 32, pp);
      if (!F16) tmem_st16(t_row + L::DP_COL + h * 32 + 16, lo);
      if (i >= 1) {
        mbar_wait(barGh + x, (i - 1) & 1);
        tc_fence_after();
      }
      tmem_st16(t_row + L::DSH_COL + h * 16, hi);
      if (i >= NDS) {                                         // this dS^T buffer was last read by the dQ MMAs of tile i - NDS
        mbar_wait(barQd + ((i - NDS) & 1), ((i - NDS) >> 1) & 1);
        tc_fence_after();
      }
      // query columns [32 h, 32 h + 32): 64-query sub-tile h >> 1, 16-byte chunks (h & 1) * 4 + g
      uint8_t* dst = sDS + (i % NDS) * L::DS_BUF + (h >> 1) * L::TILE;
      uint8_t* dsl = dst + 2 * L::TILE;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t off = sw128_offset(krow, (h & 1) * 4 + g);
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi[g * 4], hi[g * 4 + 1], hi[g * 4 + 2], hi[g * 4 + 3]);
        if (!F16) *reinterpret_cast<uint4*>(dsl + off) = make_uint4(lo[g * 4], lo[g * 4 + 1], lo[g * 4 + 2], lo[g * 4 + 3]);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(barTiles + x);                              // this thread's part of P'^T / dS^T is written
      if (i >= 1) {                                           // off the critical path: dQ_{i-1} -> global memory
        if (NDS > 1) mbar_wait(barQd + ((i - 1) & 1), ((i - 1) >> 1) & 1);
        tc_fence_after();
        flush_dq(i - 1);
      }
    }
    // ---- epilogue
    mbar_wait(barDone, 0);
    tc_fence_after();
    flush_dq(nq - 1);
    if (h == 0) {
      const int key = kt * 128 + krow;
      const size_t grow = (size_t)b * Nk + key;
      if (SPLIT_DA) {
        // columns [dV from dA_hi (dv) | dV from dA_lo (dv)]
        float acc[DVP];
#pragma unroll
        for (int c = 0; c < DVP / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(t_row + L::DV_COL + c * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[c * 16 + e] = __uint_as_float(r[e]);
        }
        constexpr int DVH = DVP / 2;   // == dv
        if (key < Nk) {
#pragma unroll
          for (int e = 0; e < DVH; e += 4)
            st4(dV + grow * DVH + e, make_float4(acc[e] + acc[DVH + e], acc[e + 1] + acc[DVH + e + 1],
                                                 acc[e + 2] + acc[DVH + e + 2], acc[e + 3] + acc[DVH + e + 3]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < DVP / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(t_row + L::DV_COL + c * 16, r);
          tmem_wait_ld();
          if (key < Nk) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              if (c * 16 + e < dv)
                st4(dV + grow * dv + c * 16 + e, make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                              __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])));
          }
        }
      }
      uint32_t r[32];
      if (F16) {                                              // [Q_hi (8) | Q_lo (8)] from the one fp16 dS term
        tmem_ld16(t_row + L::DKH_COL, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
#pragma unroll
        for (int c = 16; c < 32; ++c) r[c] = 0u;
      } else {
        tmem_ld32(t_row + L::DKH_COL, r);                     // [from dS_hi: Q_hi (8) Q_lo (8) | from dS_lo: Q_hi Q_lo]
      }
      tmem_wait_ld();
      if (key < Nk) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < d)
            dK[grow * d + c] = ((__uint_as_float(r[c]) + __uint_as_float(r[8 + c])) +
                                (__uint_as_float(r[16 + c]) + __uint_as_float(r[24 + c]))) * unscale;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == TB_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host
struct TbLayout {
  int Npad, Nk, Nkpad, DVP, kq_steps, kv_steps;
  size_t off_q, off_k, off_v, off_da, off_dat, off_qt, off_kt, off_lse, off_dd, off_gmax, off_sinv, off_dkp, off_dvp, off_ik, off_iv, total;
};

// Nk = number of keys / values (N, or N / 4 when they are down-sampled: then the pooled gradients and the argmax codes
// live in the workspace as well)
static TbLayout tb_layout(int B, int N, int Nk, int C) {
  TbLayout t;
  const int d = C / 8, dv = C / 2;
  t.Npad = (N + 127) / 128 * 128;
  t.Nk = Nk;
  t.Nkpad = (Nk + 127) / 128 * 128;
  t.DVP = C == 16 ? 16 : 32;                              // C <= 32: [dA_hi | dA_lo] rows; C = 64: dA rows unsplit
  t.kq_steps = (3 * d + 15) / 16;
  t.kv_steps = dv <= 16 ? (3 * dv + 2 + 15) / 16 : (2 * dv) / 16;   // split-bf16 dP contraction + folded D columns
  const size_t T = (size_t)B * t.Npad, Tk = (size_t)B * t.Nkpad;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 1023) / 1024 * 1024; return r; };
  t.off_q = take(T * tb_kq(C) * 2);
  t.off_k = take(Tk * tb_kq(C) * 2);
  t.off_v = take(Tk * tb_kv(C) * 2);
  t.off_da = take(T * tb_kv(C) * 2);
  t.off_dat = take((size_t)B * t.DVP * t.Npad * 2);
  t.off_qt = take((size_t)B * 16 * t.Npad * 2);
  t.off_kt = take((size_t)B * 16 * t.Nkpad * 2);
  t.off_lse = take(T * 4);
  t.off_dd = take(T * 4);
  t.off_gmax = take((size_t)B * 4);                       // per-sample max |dY| and 1 / s_b (fp16 dS path)
  t.off_sinv = take((size_t)B * 4);
  t.off_dkp = t.off_dvp = t.off_ik = t.off_iv = 0;
  if (Nk != N) {
    t.off_dkp = take((size_t)B * Nk * d * 4);
    t.off_dvp = take((size_t)B * Nk * dv * 4);
    t.off_ik = take((size_t)B * Nk * d);
    t.off_iv = take((size_t)B * Nk * dv);
  }
  t.total = o + 1024;
  return t;
}

size_t attn_tc_bwd_workspace_bytes(int B, int N, int C, bool pool) { return tb_layout(B, N, pool ? N / 4 : N, C).total; }

int attn_unpool_launch(const float* dKp, const float* dVp, const uint8_t* idxK, const uint8_t* idxV, float* dK, float* dV,
                       int B, int H, int W, int C, cudaStream_t st);      // attn_strict.cu

template <int C>
static int run_prep(const float* X, const float* dY, const float* A, const float* lse, const float* Wq, const float* bq,
                    const float* Wk, const float* bk, const float* Wv, const float* bv, const float* Wo,
                    const float* gamma, uint8_t* base, const TbLayout& t, int B, int N, int PH, int PW, cudaStream_t st) {
  const long long Tp = (long long)B * t.Npad;
  const unsigned nb = (unsigned)ceil_div<long long>(Tp, 128);
  __nv_bfloat16 *Qb = (__nv_bfloat16*)(base + t.off_q), *Kb = (__nv_bfloat16*)(base + t.off_k),
                *Vb = (__nv_bfloat16*)(base + t.off_v), *dAb = (__nv_bfloat16*)(base + t.off_da),
                *dAt = (__nv_bfloat16*)(base + t.off_dat), *Qt = (__nv_bfloat16*)(base + t.off_qt),
                *Kt = (__nv_bfloat16*)(base + t.off_kt);
  float *lse2 = (float*)(base + t.off_lse), *Dd = (float*)(base + t.off_dd);
  unsigned* gmax = (unsigned*)(base + t.off_gmax);
  float* sinv = (float*)(base + t.off_sinv);
  if (C <= 32) {       // fp16 dS: the sample's gradient scale comes from max |dY|
    SAGAN_CUDA(cudaMemsetAsync(gmax, 0, (size_t)B * 4, st));
    const long long per_sample = (long long)N * C;
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(ceil_div<long long>(per_sample, 256 * 16), (4 * num_sms() + B - 1) / B));
    attn_dy_absmax_kernel<<<dim3(gx, B), 256, 0, st>>>(dY, gmax, per_sample);
    SAGAN_LAUNCH_CHECK();
  }
  if (PH > 0) {
    attn_bwd_prep_tc_kernel<C, false><<<nb, 128, 0, st>>>(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, Qb, Kb, Vb, dAb,
                                                          dAt, Qt, Kt, lse2, Dd, gmax, sinv, B, N, t.Npad);
    SAGAN_LAUNCH_CHECK();
    attn_bwd_prep_pool_tc_kernel<C><<<(unsigned)ceil_div<long long>((long long)B * t.Nkpad, 128), 128, 0, st>>>(
        X, Wk, bk, Wv, bv, Kb, Vb, Kt, base + t.off_ik, base + t.off_iv, B, PH, PW, t.Nk, t.Nkpad);
  } else {
    attn_bwd_prep_tc_kernel<C, true><<<nb, 128, 0, st>>>(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, Qb, Kb, Vb, dAb,
                                                         dAt, Qt, Kt, lse2, Dd, gmax, sinv, B, N, t.Npad);
  }
  SAGAN_LAUNCH_CHECK();
  return 0;
}

template <int DVP, bool SPLIT_DA, int QKB, int VAB, int NDS>
static int launch_bwd(const CUtensorMap* m, const float* lse2, const float* Dd, const float* sinv, float* dQ, float* dK,
                      float* dV, int B, int N, const TbLayout& t, int d, int dv, cudaStream_t st) {
  using L = BwdSmem<DVP, QKB, VAB, NDS>;
  auto kern = attn_bwd_tc_kernel<DVP, SPLIT_DA, QKB, VAB, NDS>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<dim3(t.Nkpad / 128, B), TB_THREADS, L::TOTAL, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], lse2, Dd, sinv, dQ, dK, dV,
                                                             N, t.Npad, t.Nk, t.Nkpad, d, dv, t.kq_steps, t.kv_steps);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// dQ / dK / dV [B,N,.] fp32 from the tensor-core kernel (dQ must be zero on entry: it is accumulated atomically).
// PH, PW > 0: keys / values were max-pooled 2x2 / stride 2 over the [PH, PW] token grid; dK / dV are still TOKEN-level
// (the pooled gradients are scattered to the window positions that won).
int attn_tc_bwd_core(const float* X, const float* dY, const float* A, const float* lse, const float* Wq, const float* bq,
                     const float* Wk, const float* bk, const float* Wv, const float* bv, const float* Wo,
                     const float* gamma, float* dQ, float* dK, float* dV, int B, int N, int C, int PH, int PW, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  if (!(C == 16 || C == 32 || C == 64)) {
    set_err("sagan_attn_bwd: BF16_TC supports C in {16,32,64} (C=%d)", C);
    return SAGAN_EUNSUPPORTED;
  }
  const bool pool = PH > 0;
  const TbLayout t = tb_layout(B, N, pool ? N / 4 : N, C);
  if (ws_bytes < t.total) {
    set_err("sagan_attn_bwd: tensor-core workspace %zu < %zu bytes", ws_bytes, t.total);
    return SAGAN_EWORKSPACE;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  int rc = 0;
  switch (C) {
    case 16: rc = run_prep<16>(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, base, t, B, N, PH, PW, st); break;
    case 32: rc = run_prep<32>(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, base, t, B, N, PH, PW, st); break;
    case 64: rc = run_prep<64>(X, dY, A, lse, Wq, bq, Wk, bk, Wv, bv, Wo, gamma, base, t, B, N, PH, PW, st); break;
  }
  if (rc) return rc;
  const uint64_t Tp = (uint64_t)B * t.Npad, Tkp = (uint64_t)B * t.Nkpad;
  CUtensorMap m[7];
  const uint32_t kq = (uint32_t)tb_kq(C), kv = (uint32_t)tb_kv(C);
  if ((rc = make_tmap_bf16_2d(&m[0], base + t.off_q, Tp, kq, kq * 2, 128, kq, (int)kq * 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[1], base + t.off_k, Tkp, kq, kq * 2, 128, kq, (int)kq * 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[2], base + t.off_v, Tkp, kv, kv * 2, 128, kv, (int)kv * 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[3], base + t.off_da, Tp, kv, kv * 2, 128, kv, (int)kv * 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[4], base + t.off_dat, (uint64_t)B * t.DVP, t.Npad, (uint64_t)t.Npad * 2, t.DVP))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[5], base + t.off_qt, (uint64_t)B * 16, t.Npad, (uint64_t)t.Npad * 2, 16))) return rc;
  if ((rc = make_tmap_bf16_2d(&m[6], base + t.off_kt, (uint64_t)B * 16, t.Nkpad, (uint64_t)t.Nkpad * 2, 16))) return rc;
  const float* lse2 = (const float*)(base + t.off_lse);
  const float* Dd = (const float*)(base + t.off_dd);
  const float* sinv = (const float*)(base + t.off_sinv);
  const int d = C / 8, dv = C / 2;
  float* dKo = pool ? (float*)(base + t.off_dkp) : dK;
  float* dVo = pool ? (float*)(base + t.off_dvp) : dV;
  if (C == 16) rc = launch_bwd<16, true, 32, 64, 2>(m, lse2, Dd, sinv, dQ, dKo, dVo, B, N, t, d, dv, st);
  else if (C == 32) rc = launch_bwd<32, true, 32, 128, 1>(m, lse2, Dd, sinv, dQ, dKo, dVo, B, N, t, d, dv, st);
  else rc = launch_bwd<32, false, 64, 128, 1>(m, lse2, Dd, sinv, dQ, dKo, dVo, B, N, t, d, dv, st);
  if (rc || !pool) return rc;
  return attn_unpool_launch(dKo, dVo, base + t.off_ik, base + t.off_iv, dK, dV, B, PH, PW, C, st);
}

}  // namespace sagan
