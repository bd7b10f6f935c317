// Self-attention block, BF16_TC math mode, LARGE channel counts (C = 128 / 256 / 512): BACKWARD, host orchestration.
//
// The reference has no hand-written backward (tf.GradientTape differentiates /root/reference/layers.py:93-120); the
// formulas are SURVEY.md §8a row 2.  The score-shaped part is FUSED (attn_big_fbwd.cu: two flash launches, the [N, N]
// maps never leave the SM); around it every contraction over the channel axis runs on the tensor cores:
//
//   Q', K, V (bf16)     = X [Wtheta | Wphi | Wg] + b          CTA-pair tf32 GEMM, gemm_tc.cu (the forward's own launch,
//                                                             so S is recomputed from exactly the forward's operands)
//   dA' (bf16)          = dY Wo^T                             same GEMM (gamma is folded into the flash epilogues)
//   D', lse2            = rowsum(dA' * A), lse log2 e         one warp per token
//   dV, dK | dQ         flash backward, attn_big_fbwd.cu
//   dWo', dbo', dgamma  = A^T dY, colsum dY, finalize         transpose + conv-as-GEMM, column sums
//   dX = dY + dQ Wq^T + dK Wk^T + dV Wv^T                     3 x conv dgrad + adds
//   dWq, dbq, ...       = X^T [dQ dK dV], column sums         transpose of X + 3 x conv-as-GEMM (split-K over the tokens)
//
// Launch count is independent of the batch (round 1 ran ~14 launches PER SAMPLE over [N, N] fp32 tensors in HBM:
// 9.6 ms at B=16, N=4096, C=512 and 21 ms at B=256, N=256, C=128).
#include <algorithm>

#include <stdlib.h>

#include "common.cuh"

namespace sagan {

// gram_tc.cu
bool gram_split_supported(int l, int r, long long T);
int gram_split(const float* L, int l, const float* R, int r, float* G, float* colsum, long long T, cudaStream_t st);


int attn_bwd_finalize_launch(const float* Wo, const float* bo, const float* gamma, float* dWo, float* dbo, float* dgamma,
                             int nW, int C, cudaStream_t st);   // attn_strict.cu

namespace {

__global__ void bb_set_kernel(float* p, float v) { *p = v; }

// src [R][2 d + dv] -> q [R][d], k [R][d], v [R][dv]
__global__ void bb_split_cols_kernel(const float* __restrict__ src, float* __restrict__ q, float* __restrict__ k,
                                     float* __restrict__ v, int R, int d, int dv) {
  const int nn = 2 * d + dv;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * nn) return;
  const int r = i / nn, c = i - r * nn;
  if (c < d) q[r * d + c] = src[i];
  else if (c < 2 * d) k[r * d + c - d] = src[i];
  else v[r * dv + c - 2 * d] = src[i];
}

// wcat [C][2 d + dv] = [Wq | Wk | Wv] row by row: the K-major "transposed" operand of dX = dY + [dQ dK dV] Wcat^T
__global__ void bb_concat_w_kernel(const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
                                   float* __restrict__ wcat, int C, int d, int dv) {
  const int nn = 2 * d + dv;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * nn) return;
  const int r = i / nn, c = i - r * nn;
  wcat[i] = c < d ? Wq[r * d + c] : (c < 2 * d ? Wk[r * d + c - d] : Wv[r * dv + c - 2 * d]);
}

// dgamma += sum Wo * dWo' + sum bo * dbo'; dWo = gamma dWo', dbo = gamma dbo'   (many CTAs; dgamma zeroed by the caller)
__global__ void __launch_bounds__(256)
bb_finalize_kernel(const float* __restrict__ Wo, const float* __restrict__ bo, const float* __restrict__ gamma,
                   float* __restrict__ dWo, float* __restrict__ dbo, float* __restrict__ dgamma, int nW, int C) {
  __shared__ float red[32];
  const float gm = *gamma;
  float acc = 0.f;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < nW) {
    const float g = dWo[i];
    acc = Wo[i] * g;
    dWo[i] = g * gm;
  } else if (i - nW < C) {
    const float g = dbo[i - nW];
    acc = bo[i - nW] * g;
    dbo[i - nW] = g * gm;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(dgamma, acc);
}

sagan_conv_geom dense_geom(long long rows, int cin, int cout) {
  sagan_conv_geom g;
  g.B = (int32_t)rows; g.H = 1; g.W = 1; g.Cin = cin;
  g.Ho = 1; g.Wo = 1; g.Cout = cout;
  g.kh = 1; g.kw = 1; g.stride = 1; g.pad_t = 0; g.pad_l = 0;
  return g;
}

struct BbLayout {
  size_t dqkv, dwcat, dbcat, wcat, one, wt, bcat, wot, zero, qb, kb, vb, dab, lse2, dd, total;   // float offsets (256-byte aligned)
};

BbLayout bb_layout(int B, int N, int C) {
  const size_t T = (size_t)B * N, d = C / 8, dv = C / 2;
  BbLayout t;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  t.dqkv = take(T * (2 * d + dv));                       // [dQ | dK | dV] rows
  t.dwcat = take((size_t)C * (2 * d + dv)); t.dbcat = take(2 * d + dv); t.wcat = take((size_t)C * (2 * d + dv)); t.one = take(1);
  t.wt = take((size_t)C * (2 * d + dv)); t.bcat = take(2 * d + dv); t.wot = take((size_t)C * dv / 2); t.zero = take(C);
  t.qb = take(T * 32); t.kb = take(T * 32);               // bf16 [T][64]
  t.vb = take(T * dv / 2); t.dab = take(T * dv / 2);      // bf16 [T][dv]
  t.lse2 = take(T); t.dd = take(T);
  t.total = o + 64;
  return t;
}

}  // namespace

size_t attn_big_bwd_workspace_bytes(int B, int N, int C) { return bb_layout(B, N, C).total * sizeof(float); }

#define BB_CALL(expr)        \
  do {                       \
    int rc_ = (expr);        \
    if (rc_) return rc_;     \
  } while (0)

// gemm_tc.cu / attn_tc_big.cu / attn_big_fbwd.cu
int gemm_tf32_qkv(const float* x, const float* wt, const float* bcat, __nv_bfloat16* q, __nv_bfloat16* k,
                  __nv_bfloat16* v, long long M, int K, int d, int dv, float q_scale, cudaStream_t st);
int attn_big_weights_launch(const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                            const float* bv, const float* Wo, float* Wt, float* bcat, __nv_bfloat16* WoT, int C, cudaStream_t st);
int attn_big_fused_bwd_core(const __nv_bfloat16* Qb, const __nv_bfloat16* Kb, const __nv_bfloat16* Vb,
                            const __nv_bfloat16* dAb, const float* A, const float* lse, const float* gamma, float* lse2,
                            float* Dd, float* dQKV, int B, int N, int C, cudaStream_t st);
int gemm_tf32_residual(const float* a, const float* wt, const float* bias, const float* res, const float* res_scale,
                       float* y, long long M, int K, int N, cudaStream_t st);

int attn_tc_big_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                    const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma, const float* A,
                    const float* lse, float* dX, float* dWq, float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo,
                    float* dbo, float* dgamma, int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  const BbLayout t = bb_layout(B, N, C);
  if (ws_bytes < t.total * sizeof(float)) {
    set_err("sagan_attn_bwd: workspace %zu < %zu bytes", ws_bytes, t.total * sizeof(float));
    return SAGAN_EWORKSPACE;
  }
  const int d = C / 8, dv = C / 2;
  const long long T = (long long)B * N;
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int NN = 2 * d + dv;
  float *dQKV = base + t.dqkv, *dWcat = base + t.dwcat, *dbcat = base + t.dbcat, *Wcat = base + t.wcat, *one = base + t.one;
  float *Wt = base + t.wt, *bcat = base + t.bcat, *zero = base + t.zero, *lse2 = base + t.lse2, *Dd = base + t.dd;
  __nv_bfloat16 *WoT = (__nv_bfloat16*)(base + t.wot), *Qb = (__nv_bfloat16*)(base + t.qb), *Kb = (__nv_bfloat16*)(base + t.kb),
                *Vb = (__nv_bfloat16*)(base + t.vb), *dAb = (__nv_bfloat16*)(base + t.dab);
  const int TC = SAGAN_MATH_BF16_TC;
  const bool want_w = dWq != nullptr;

  // ---- bf16 operands of the flash kernels: the forward's own projection GEMM, and dA' = dY Wo^T through the same GEMM
  //      (Wo [dv, C] row-major IS the K-major transposed operand; no bias; all columns go to the "v" output)
  BB_CALL(attn_big_weights_launch(Wq, bq, Wk, bk, Wv, bv, Wo, Wt, bcat, WoT, C, st));
  if (d < 64) {   // rows of Q / K are padded to one 128-byte swizzle span
    SAGAN_CUDA(cudaMemsetAsync(Qb, 0, (size_t)T * 64 * 2, st));
    SAGAN_CUDA(cudaMemsetAsync(Kb, 0, (size_t)T * 64 * 2, st));
  }
  SAGAN_CUDA(cudaMemsetAsync(zero, 0, sizeof(float) * C, st));
  BB_CALL(gemm_tf32_qkv(X, Wt, bcat, Qb, Kb, Vb, T, C, d, dv, 1.4426950408889634f, st));
  BB_CALL(gemm_tf32_qkv(dY, Wo, zero, nullptr, nullptr, dAb, T, C, 0, dv, 1.0f, st));
  // G = L^T R and the column sums of R for token-major L [T, l], R [T, r]: the backward-filter form of the 1x1 conv
  // (split-bf16 tensor-core kernel: both operands read MN-major, reduction over the tokens split across CTAs, the
  // ones column yields colsum(R) for free) -- no transposed copies
  auto gram = [&](const float* Lm, int l, const float* Rm, int r, float* out, float* colsum_out) -> int {
    // gram_tc.cu: TMA-fed, fp32 -> split-bf16 conversion inside the CTA, reduction split over the tokens
    static const bool conv_form = getenv("SAGAN_GRAM_CONV") != nullptr;      // diagnostics only: the round-2a path
    if (gram_split_supported(l, r, T) && !conv_form) return gram_split(Lm, l, Rm, r, out, colsum_out, T, st);
    const sagan_conv_geom g = dense_geom(T, l, r);
    return sagan_conv2d_wgrad(Lm, Rm, out, colsum_out, &g, TC, st);
  };
  if (want_w) {
    BB_CALL(gram(A, dv, dY, C, dWo, dbo));                                    // dWo' = A^T dY, dbo' = colsum(dY)
    SAGAN_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float), st));
    bb_finalize_kernel<<<ceil_div(dv * C + C, 256), 256, 0, st>>>(Wo, bo, gamma, dWo, dbo, dgamma, dv * C, C);
    SAGAN_LAUNCH_CHECK();
  }

  // ---- the score-shaped part: two flash launches (dK, dV | dQ), gamma folded into their epilogues
  BB_CALL(attn_big_fused_bwd_core(Qb, Kb, Vb, dAb, A, lse, gamma, lse2, Dd, dQKV, B, N, C, st));

  // ---- back through the projections, on the concatenated [dQ | dK | dV] rows
  if (want_w) {
    // [dWq | dWk | dWv] = X^T dQKV (one GEMM whose reduction runs over the tokens), [dbq | dbk | dbv] = column sums
    BB_CALL(gram(X, C, dQKV, NN, dWcat, dbcat));
    bb_split_cols_kernel<<<ceil_div(C * NN, 256), 256, 0, st>>>(dWcat, dWq, dWk, dWv, C, d, dv);
    SAGAN_LAUNCH_CHECK();
    bb_split_cols_kernel<<<ceil_div(NN, 256), 256, 0, st>>>(dbcat, dbq, dbk, dbv, 1, d, dv);
    SAGAN_LAUNCH_CHECK();
  }
  if (dX) {
    // dX = dY + dQKV [Wq | Wk | Wv]^T: one CTA-pair tf32 GEMM with the residual epilogue
    bb_concat_w_kernel<<<ceil_div(C * NN, 256), 256, 0, st>>>(Wq, Wk, Wv, Wcat, C, d, dv);
    SAGAN_LAUNCH_CHECK();
    bb_set_kernel<<<1, 1, 0, st>>>(one, 1.0f);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(gemm_tf32_residual(dQKV, Wcat, zero, dY, one, dX, T, NN, C, st));
  }
  return 0;
}

}  // namespace sagan
