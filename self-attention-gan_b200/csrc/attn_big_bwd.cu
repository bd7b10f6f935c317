// Self-attention block, BF16_TC math mode, LARGE channel counts (C = 128 / 256 / 512): BACKWARD.
//
// The reference has no hand-written backward (tf.GradientTape differentiates /root/reference/layers.py:93-120); the
// formulas are SURVEY.md §8a row 2.  This first large-C backward is COMPOSED, not fused: every contraction runs on the
// tensor cores through the library's own implicit-GEMM kernels in their kind::tf32 forms (1x1 geometry, conv_tc.cu; the
// backward-filter form rounds both operands to bf16, which costs 5e-3 on every gradient here, so the transposed
// contractions are run as forward GEMMs on explicitly transposed operands instead), the score-shaped tensors S, P, dP,
// dS of ONE sample at a time live in the workspace ([N, N] fp32 each), and two row kernels do the softmax and its Jacobian:
//
//   theta, phi, g = X W + b                      3 x conv fwd   (the forward's bf16 copies are not saved: recomputed)
//   dA = gamma dY Wo^T                           conv dgrad + scale
//   dWo', dbo' = A^T dY, colsum dY               transpose + conv fwd, column sums (gamma, dgamma: finalize kernel)
//   per sample b:
//     S   = theta_b phi_b^T                      conv fwd against the transposed phi_b
//     P   = softmax_rows(S)                      row kernel (in place); self-consistent with the S computed HERE
//     dP  = dA_b g_b^T                           conv fwd against the transposed g_b
//     dS  = P * (dP - rowsum(P * dP))            row kernel (in place)
//     dg_b     = P^T dA_b                        transpose + conv fwd
//     dphi_b   = dS^T theta_b                    transpose + conv fwd
//     dtheta_b = dS phi_b                        conv fwd (sum_j dS_ij = 0: this contraction cancels and amplifies the tf32
//                                                truncation of dS; measured +0.6e-3 on dWtheta against an fp32 GEMM)
//   dX = dY + dtheta Wq^T + dphi Wk^T + dg Wv^T  3 x conv dgrad + adds
//   dWq, dbq, ... = X^T [dtheta dphi dg]          transpose of X + 3 x conv fwd (split-K over the tokens), column sums
//
// ~14 launches per sample: HBM-bound on the [N, N] tensors (about 14 passes of 4 N^2 bytes per sample), i.e. roughly
// 15-20 x the fused forward.  The fused plan (two flash kernels) replaces the per-sample part.
#include <algorithm>

#include "common.cuh"

namespace sagan {

int attn_bwd_finalize_launch(const float* Wo, const float* bo, const float* gamma, float* dWo, float* dbo, float* dgamma,
                             int nW, int C, cudaStream_t st);   // attn_strict.cu

namespace {

// [rows, cols] -> [cols, rows]
__global__ void bb_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// one CTA per row: P = softmax(S) in place
__global__ void __launch_bounds__(256) bb_softmax_rows_kernel(float* __restrict__ S, int n) {
  __shared__ float red[32];
  float* row = S + (size_t)blockIdx.x * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float e = __expf(row[i] - mx);
    row[i] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < n; i += 256) row[i] *= inv;
}

// one CTA per row: dP <- P * (dP - sum_j P_j dP_j)
__global__ void __launch_bounds__(256) bb_ds_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, int n) {
  __shared__ float red[32];
  const float* p = P + (size_t)blockIdx.x * n;
  float* g = dP + (size_t)blockIdx.x * n;
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) dot = fmaf(p[i], g[i], dot);
  dot = block_sum(dot, red);
  for (int i = threadIdx.x; i < n; i += 256) g[i] = p[i] * (g[i] - dot);
}

__global__ void bb_scale_kernel(float* __restrict__ x, const float* __restrict__ s, long long n) {
  const float f = *s;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= f;
}

// out = a + b  or  out += b
__global__ void bb_add_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (a ? a[i] : out[i]) + b[i];
}

// out[c] = sum_r x[r, c]   (out zeroed by the caller; fp32 partial sums per CTA, one atomic per column and CTA)
__global__ void __launch_bounds__(256) bb_colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows,
                                                        int cols, long long rows_per_block) {
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c = blockIdx.x * 256 + threadIdx.x; c < cols; c += gridDim.x * 256) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += x[r * cols + c];
    atomicAdd(out + c, acc);
  }
}

sagan_conv_geom dense_geom(long long rows, int cin, int cout) {
  sagan_conv_geom g;
  g.B = (int32_t)rows; g.H = 1; g.W = 1; g.Cin = cin;
  g.Ho = 1; g.Wo = 1; g.Cout = cout;
  g.kh = 1; g.kw = 1; g.stride = 1; g.pad_t = 0; g.pad_l = 0;
  return g;
}

struct BbLayout {
  size_t q, k, v, dq, dk, dv, da, kt, vt, s, dp, tr, tmp, total;   // float offsets
};

BbLayout bb_layout(int B, int N, int C) {
  const size_t T = (size_t)B * N, d = C / 8, dv = C / 2;
  BbLayout t;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  t.q = take(T * d); t.k = take(T * d); t.v = take(T * dv);
  t.dq = take(T * d); t.dk = take(T * d); t.dv = take(T * dv);
  t.da = take(T * dv);
  t.kt = take((size_t)N * d); t.vt = take((size_t)N * dv);
  t.s = take((size_t)N * N); t.dp = take((size_t)N * N); t.tr = take((size_t)N * N);
  t.tmp = take(T * C);
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

int attn_tc_big_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                    const float* Wv, const float* bv, const float* Wo, const float* bo, const float* gamma, const float* A,
                    float* dX, float* dWq, float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo,
                    float* dbo, float* dgamma, int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  const BbLayout t = bb_layout(B, N, C);
  if (ws_bytes < t.total * sizeof(float)) {
    set_err("sagan_attn_bwd: workspace %zu < %zu bytes", ws_bytes, t.total * sizeof(float));
    return SAGAN_EWORKSPACE;
  }
  const int d = C / 8, dv = C / 2;
  const long long T = (long long)B * N;
  float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float *Q = base + t.q, *K = base + t.k, *V = base + t.v, *dQ = base + t.dq, *dK = base + t.dk, *dV = base + t.dv;
  float *dA = base + t.da, *KT = base + t.kt, *VT = base + t.vt, *S = base + t.s, *dP = base + t.dp, *TR = base + t.tr;
  float* tmp = base + t.tmp;
  const int TC = SAGAN_MATH_BF16_TC;
  const bool want_w = dWq != nullptr;
  const int ew_blocks = num_sms() * 8;

  // ---- projections and the gradient of the attention output
  const sagan_conv_geom gq = dense_geom(T, C, d), gv = dense_geom(T, C, dv), go = dense_geom(T, dv, C);
  BB_CALL(sagan_conv2d_fwd(X, Wq, bq, Q, &gq, SAGAN_ACT_NONE, 0.f, TC, st));
  BB_CALL(sagan_conv2d_fwd(X, Wk, bk, K, &gq, SAGAN_ACT_NONE, 0.f, TC, st));
  BB_CALL(sagan_conv2d_fwd(X, Wv, bv, V, &gv, SAGAN_ACT_NONE, 0.f, TC, st));
  BB_CALL(sagan_conv2d_dgrad(dY, Wo, dA, &go, TC, st));                      // dA' = dY Wo^T
  const dim3 tb(32, 8);
  auto colsum = [&](const float* x, float* out, int cols) -> int {
    SAGAN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    const long long rpb = 512;
    bb_colsum_kernel<<<dim3(ceil_div(cols, 256), (unsigned)ceil_div<long long>(T, rpb)), 256, 0, st>>>(x, out, T, cols, rpb);
    SAGAN_LAUNCH_CHECK();
    return 0;
  };
  // G = L^T R for token-major L [T, l], R [T, r]: transpose L into tmp, then a forward GEMM whose reduction runs over T
  auto gram = [&](const float* Lm, int l, const float* Rm, int r, float* out) -> int {
    bb_transpose_kernel<<<dim3(ceil_div(l, 32), (unsigned)ceil_div<long long>(T, 32)), tb, 0, st>>>(Lm, tmp, (int)T, l);
    SAGAN_LAUNCH_CHECK();
    const sagan_conv_geom g = dense_geom(l, (int)T, r);
    return sagan_conv2d_fwd(tmp, Rm, nullptr, out, &g, SAGAN_ACT_NONE, 0.f, TC, st);
  };
  if (want_w) {
    BB_CALL(gram(A, dv, dY, C, dWo));                                         // dWo' = A^T dY
    BB_CALL(colsum(dY, dbo, C));                                              // dbo' = colsum(dY)
    BB_CALL(attn_bwd_finalize_launch(Wo, bo, gamma, dWo, dbo, dgamma, dv * C, C, st));
  }
  bb_scale_kernel<<<ew_blocks, 256, 0, st>>>(dA, gamma, T * dv);
  SAGAN_LAUNCH_CHECK();

  // ---- per sample: the score-shaped part
  const sagan_conv_geom gs = dense_geom(N, d, N), gp = dense_geom(N, dv, N);          // S = theta phi^T, dP = dA g^T
  const sagan_conv_geom gdv = dense_geom(N, N, dv), gdk = dense_geom(N, N, d);        // P^T dA; dS^T theta and dS phi
  const dim3 tgrid(ceil_div(N, 32), ceil_div(N, 32));
  for (int b = 0; b < B; ++b) {
    const size_t r0 = (size_t)b * N;
    bb_transpose_kernel<<<dim3(ceil_div(d, 32), ceil_div(N, 32)), tb, 0, st>>>(K + r0 * d, KT, N, d);
    SAGAN_LAUNCH_CHECK();
    bb_transpose_kernel<<<dim3(ceil_div(dv, 32), ceil_div(N, 32)), tb, 0, st>>>(V + r0 * dv, VT, N, dv);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_fwd(Q + r0 * d, KT, nullptr, S, &gs, SAGAN_ACT_NONE, 0.f, TC, st));
    bb_softmax_rows_kernel<<<N, 256, 0, st>>>(S, N);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_fwd(dA + r0 * dv, VT, nullptr, dP, &gp, SAGAN_ACT_NONE, 0.f, TC, st));
    bb_transpose_kernel<<<tgrid, tb, 0, st>>>(S, TR, N, N);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_fwd(TR, dA + r0 * dv, nullptr, dV + r0 * dv, &gdv, SAGAN_ACT_NONE, 0.f, TC, st));   // dg = P^T dA
    bb_ds_rows_kernel<<<N, 256, 0, st>>>(S, dP, N);                                      // dP <- dS
    SAGAN_LAUNCH_CHECK();
    bb_transpose_kernel<<<tgrid, tb, 0, st>>>(dP, TR, N, N);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_fwd(TR, Q + r0 * d, nullptr, dK + r0 * d, &gdk, SAGAN_ACT_NONE, 0.f, TC, st));      // dphi = dS^T theta
    BB_CALL(sagan_conv2d_fwd(dP, K + r0 * d, nullptr, dQ + r0 * d, &gdk, SAGAN_ACT_NONE, 0.f, TC, st));        // dtheta = dS phi
  }

  // ---- back through the projections (the weight gradients first: they use tmp for X^T)
  if (want_w) {
    bb_transpose_kernel<<<dim3(ceil_div(C, 32), (unsigned)ceil_div<long long>(T, 32)), tb, 0, st>>>(X, tmp, (int)T, C);
    SAGAN_LAUNCH_CHECK();
    const sagan_conv_geom gwq = dense_geom(C, (int)T, d), gwv = dense_geom(C, (int)T, dv);
    BB_CALL(sagan_conv2d_fwd(tmp, dQ, nullptr, dWq, &gwq, SAGAN_ACT_NONE, 0.f, TC, st));
    BB_CALL(sagan_conv2d_fwd(tmp, dK, nullptr, dWk, &gwq, SAGAN_ACT_NONE, 0.f, TC, st));
    BB_CALL(sagan_conv2d_fwd(tmp, dV, nullptr, dWv, &gwv, SAGAN_ACT_NONE, 0.f, TC, st));
    BB_CALL(colsum(dQ, dbq, d));
    BB_CALL(colsum(dK, dbk, d));
    BB_CALL(colsum(dV, dbv, dv));
  }
  if (dX) {
    BB_CALL(sagan_conv2d_dgrad(dQ, Wq, tmp, &gq, TC, st));
    bb_add_kernel<<<ew_blocks, 256, 0, st>>>(dX, dY, tmp, T * C);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_dgrad(dK, Wk, tmp, &gq, TC, st));
    bb_add_kernel<<<ew_blocks, 256, 0, st>>>(dX, nullptr, tmp, T * C);
    SAGAN_LAUNCH_CHECK();
    BB_CALL(sagan_conv2d_dgrad(dV, Wv, tmp, &gv, TC, st));
    bb_add_kernel<<<ew_blocks, 256, 0, st>>>(dX, nullptr, tmp, T * C);
    SAGAN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace sagan
