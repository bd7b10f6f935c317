// Self-attention block, BF16_TC math mode: flash-style forward on the 5th-gen tensor cores.
//
// Replaces Attention_Layer.call (/root/reference/layers.py:93-120): S = theta phi^T and A = softmax(S) g are
// tcgen05.mma (bf16 operands, fp32 accumulators in TMEM), K / V^T tiles arrive by TMA into 128-byte-swizzled
// shared memory, the online softmax runs on fp32 rows read back with tcgen05.ld (one thread per query row),
// and the output 1x1 conv + gamma residual are fused in the epilogue.  The [B,N,N] map is never written.
//
// CTA = one 128-query tile of one sample; 6 warps:
//   warps 0-3  softmax / epilogue (thread r <-> TMEM lane r <-> query row r)
//   warp  4    TMA producer (one elected lane), TMEM allocator
//   warp  5    MMA issuer (one elected lane)
// Per 128-key tile j:   S_j = Q K_j^T   (M=128, N=128, K=16*kq_steps)   -> TMEM columns [S]
//                       P_j = exp2(S_j - m)  -> bf16, written to shared memory in the UMMA K-major SW128 layout
//                       O  += P_j V_j  (M=128, N=DVP, K=128)            -> TMEM columns [O]
// The running-max rescale of O is lazy (only when the row max grows by > 2^32), so O stays in TMEM.
//
// Consistent rounding.  The softmax weights enter the PV MMA as bf16.  To keep the block and its backward an exact
// (fp32-accurate) function / gradient pair of ONE well-defined set of weights, the running max is kept INTEGER-valued
// (log2 units), so P'_ij = bf16(exp2(S_ij - m)) does not depend on which m was current (bf16 rounding commutes with
// powers of two); the row sum l' = sum_j P'_ij is accumulated by the same MMA through a ones row appended to V^T, and
// A = (P' V) / l'.  The backward (attn_tc_bwd.cu) regenerates exactly these P' from S and lse' = m + log2 l'.
// NS = number of S buffers: 2 lets QK_{j+1} run under softmax_j (ping-pong inside the CTA, one CTA per SM);
// NS = 1 is used for small value dims where two CTAs share an SM and overlap each other instead.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "attn_pool.cuh"

namespace sagan {

using namespace tc;

constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr float TC_LN2 = 0.6931471805599453f;
constexpr int TC_THREADS = 320;   // 8 softmax warps (two key-column halves per query row) + TMA producer + MMA issuer
// Q / K rows: 16 bf16 (32 B, SWIZZLE_32B tiles) when the split logits fit one MMA K step (C <= 32: 3 d <= 12),
// otherwise 64 bf16 (128 B, SWIZZLE_128B)
__host__ __device__ constexpr int qk_cols(int C) { return C <= 32 ? 16 : 64; }

// ------------------------------------------------------------------------------------ host: tensor maps
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled is not available from this driver");
    return SAGAN_EUNSUPPORTED;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                       : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu stride=%llu box=%ux%u)", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_rows, box_cols);
    return SAGAN_EINVAL;
  }
  return 0;
}

// ------------------------------------------------------------------------------------ projections (small C)
// one thread per PADDED token: Qb[b][n][64] = bf16(log2e * (x Wq + bq)), Kb[b][n][64] = bf16(x Wk + bk),
// Vt[b][v][n] = bf16(x Wv + bv); rows n >= N are zero (so masked keys contribute exactly 0 to P V).
// KV = false (down-sampled keys / values): only the query rows are written; K / V^T come from attn_pool_proj_tc_kernel.
template <int C, bool KV>
__global__ void __launch_bounds__(128)
attn_proj_tc_kernel(const float* __restrict__ X, const float* __restrict__ Wq, const float* __restrict__ bq,
                    const float* __restrict__ Wk, const float* __restrict__ bk, const float* __restrict__ Wv,
                    const float* __restrict__ bv, __nv_bfloat16* __restrict__ Qb, __nv_bfloat16* __restrict__ Kb,
                    __nv_bfloat16* __restrict__ Vt, int B, int N, int Npad) {
  constexpr int D = C / 8, DV = C / 2;
  constexpr int DVP = ((2 * DV + 1 + 15) / 16) * 16;   // V^T rows: [v_hi (DV) | v_lo (DV) | ones | 0 ...]
  __shared__ float sWq[C * D], sWk[C * D], sWv[C * DV], sbq[D], sbk[D], sbv[DV];
  for (int i = threadIdx.x; i < C * D; i += 128) { sWq[i] = Wq[i]; sWk[i] = Wk[i]; }
  for (int i = threadIdx.x; i < C * DV; i += 128) sWv[i] = Wv[i];
  for (int i = threadIdx.x; i < D; i += 128) { sbq[i] = bq[i]; sbk[i] = bk[i]; }
  for (int i = threadIdx.x; i < DV; i += 128) sbv[i] = bv[i];
  __syncthreads();
  const long long tp = (long long)blockIdx.x * 128 + threadIdx.x;
  if (tp >= (long long)B * Npad) return;
  const int b = (int)(tp / Npad), n = (int)(tp - (long long)b * Npad);
  const bool valid = n < N;
  float x[C];
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = ld4(X + ((long long)b * N + n) * C + c);
    x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
  }
  // Split-bf16 logits: q = q_hi + q_lo, k = k_hi + k_lo (each half a bf16).  The K dimension of the QK^T MMA is
  // padded to a multiple of 16 anyway, so the row is laid out as  Q: [q_hi | q_lo | q_hi | 0..]  K: [k_hi | k_hi | k_lo | 0..]
  // and one MMA yields q_hi.k_hi + q_lo.k_hi + q_hi.k_lo -- logits accurate to ~2^-16 instead of 2^-8 at no extra cost
  // for d <= 5 (the logits are NOT scaled by 1/sqrt(d), layers.py:108, so bf16 logits would dominate the error).
  constexpr int KQ = ((3 * D + 15) / 16) * 16;
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
    a = valid ? a * TC_LOG2E : 0.f;
    kk = valid ? kk : 0.f;
    const float a_hi = __bfloat162float(__float2bfloat16_rn(a)), k_hi = __bfloat162float(__float2bfloat16_rn(kk));
    q[j] = a_hi; q[D + j] = a - a_hi; q[2 * D + j] = a_hi;
    k[j] = k_hi; k[D + j] = k_hi;     k[2 * D + j] = kk - k_hi;
  }
  uint4* qd = reinterpret_cast<uint4*>(Qb + tp * qk_cols(C));
  uint4* kd = reinterpret_cast<uint4*>(Kb + tp * qk_cols(C));
#pragma unroll
  for (int g = 0; g < KQ / 8; ++g) {
    qd[g] = make_uint4(pack_bf16x2(q[g * 8 + 0], q[g * 8 + 1]), pack_bf16x2(q[g * 8 + 2], q[g * 8 + 3]),
                       pack_bf16x2(q[g * 8 + 4], q[g * 8 + 5]), pack_bf16x2(q[g * 8 + 6], q[g * 8 + 7]));
    if (KV)
      kd[g] = make_uint4(pack_bf16x2(k[g * 8 + 0], k[g * 8 + 1]), pack_bf16x2(k[g * 8 + 2], k[g * 8 + 3]),
                         pack_bf16x2(k[g * 8 + 4], k[g * 8 + 5]), pack_bf16x2(k[g * 8 + 6], k[g * 8 + 7]));
  }
  if (!KV) return;
  // values are split the same way (v = v_hi + v_lo, two bf16 rows of V^T), so O = P [v_hi | v_lo] carries V exactly
  // and the only bf16 rounding left in A is that of P itself
#pragma unroll
  for (int v = 0; v < DV; ++v) {
    float a = sbv[v];
#pragma unroll
    for (int c = 0; c < C; ++c) a = fmaf(x[c], sWv[c * DV + v], a);
    a = valid ? a : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(a);
    Vt[((long long)b * DVP + v) * Npad + n] = hi;
    Vt[((long long)b * DVP + DV + v) * Npad + n] = __float2bfloat16_rn(a - __bfloat162float(hi));
  }
  // ones row: column 2 DV of O accumulates l' = sum_j P'_ij (masked / padded keys contribute 0)
  Vt[((long long)b * DVP + 2 * DV) * Npad + n] = __float2bfloat16_rn(valid ? 1.0f : 0.0f);
#pragma unroll
  for (int v = 2 * DV + 1; v < DVP; ++v) Vt[((long long)b * DVP + v) * Npad + n] = __float2bfloat16_rn(0.0f);
}

// ------------------------------------------------------------------------------------ down-sampled keys / values
// SURVEY.md §8f row 2 (/root/reference/layers.py:96,100,113): phi and g max-pooled 2x2 / stride 2 over the [H, W] token
// grid, per channel.  pooled_kv computes the two projections of the four tokens of one window in fp32 (the SAME fmaf
// sequence in the forward and the backward, so both pick the same winners), the channel maxima and which window position
// (0..3, first maximum wins) supplied each of them.
// one thread per PADDED pooled position: Kb rows (split-bf16 layout of attn_proj_tc_kernel) and V^T columns
template <int C>
__global__ void __launch_bounds__(128)
attn_pool_proj_tc_kernel(const float* __restrict__ X, const float* __restrict__ Wk, const float* __restrict__ bk,
                         const float* __restrict__ Wv, const float* __restrict__ bv, __nv_bfloat16* __restrict__ Kb,
                         __nv_bfloat16* __restrict__ Vt, int B, int H, int W, int Nk, int Nkpad) {
  constexpr int D = C / 8, DV = C / 2;
  constexpr int DVP = ((2 * DV + 1 + 15) / 16) * 16;
  constexpr int KQ = ((3 * D + 15) / 16) * 16;
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
  for (int j = 0; j < D; ++j) kk[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DV; ++j) vv[j] = 0.f;
  if (valid) pooled_kv<C>(X, sWk, sbk, sWv, sbv, b, H, W, n / (W / 2), n % (W / 2), kk, vv, ik, iv);
  float k[KQ];
#pragma unroll
  for (int j = 0; j < KQ; ++j) k[j] = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const float k_hi = __bfloat162float(__float2bfloat16_rn(kk[j]));
    k[j] = k_hi; k[D + j] = k_hi; k[2 * D + j] = kk[j] - k_hi;
  }
  uint4* kd = reinterpret_cast<uint4*>(Kb + tp * qk_cols(C));
#pragma unroll
  for (int g = 0; g < KQ / 8; ++g)
    kd[g] = make_uint4(pack_bf16x2(k[g * 8 + 0], k[g * 8 + 1]), pack_bf16x2(k[g * 8 + 2], k[g * 8 + 3]),
                       pack_bf16x2(k[g * 8 + 4], k[g * 8 + 5]), pack_bf16x2(k[g * 8 + 6], k[g * 8 + 7]));
#pragma unroll
  for (int v = 0; v < DV; ++v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(vv[v]);
    Vt[((long long)b * DVP + v) * Nkpad + n] = hi;
    Vt[((long long)b * DVP + DV + v) * Nkpad + n] = __float2bfloat16_rn(vv[v] - __bfloat162float(hi));
  }
  Vt[((long long)b * DVP + 2 * DV) * Nkpad + n] = __float2bfloat16_rn(valid ? 1.0f : 0.0f);
#pragma unroll
  for (int v = 2 * DV + 1; v < DVP; ++v) Vt[((long long)b * DVP + v) * Nkpad + n] = __float2bfloat16_rn(0.0f);
}

// ------------------------------------------------------------------------------------ flash forward
#ifndef TC_POLY_EXP
#define TC_POLY_EXP 1
#endif
// 2^x on the FMA pipe for one key column in four: round-to-nearest split x = n + f (magic-number add), cubic minimax of
// 2^f on [-0.5, 0.5] (max relative error 7.5e-5, 50 x below the bf16 rounding of P'), n added into the exponent field.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;
  const float f = x - (r - 12582912.0f);
  const float p = fmaf(fmaf(fmaf(5.517132208e-02f, f, 2.426105440e-01f), f, 6.932609677e-01f), f, 9.999281168e-01f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// the same arithmetic on a pair of columns with packed fp32 instructions (bit-identical per lane to ex2_poly)
__device__ __forceinline__ uint32_t ex2_poly2_bf16(f2 x) {
  float a, b;
  f2_unpack(x, a, b);
  x = f2_pack(fmaxf(a, -125.0f), fmaxf(b, -125.0f));
  const f2 magic = f2_pack(12582912.0f, 12582912.0f);
  const f2 r = f2_add(x, magic);
  const f2 f = f2_sub(x, f2_sub(r, magic));
  f2 p = f2_fma(f2_pack(5.517132208e-02f, 5.517132208e-02f), f, f2_pack(2.426105440e-01f, 2.426105440e-01f));
  p = f2_fma(p, f, f2_pack(6.932609677e-01f, 6.932609677e-01f));
  p = f2_fma(p, f, f2_pack(9.999281168e-01f, 9.999281168e-01f));
  float p0, p1, r0, r1;
  f2_unpack(p, p0, p1);
  f2_unpack(r, r0, r1);
  return pack_bf16x2(__int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23)),
                     __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23)));
}

template <int DVP, int NS, int NP, int CEPI>
struct FwdSmem {
  static constexpr int QKB = qk_cols(CEPI) * 2;       // bytes per Q / K row
  static constexpr int Q_BYTES = 128 * QKB;
  static constexpr int K_BYTES = 128 * QKB;
  static constexpr int V_BYTES = 2 * DVP * 128;      // two 64-key sub-tiles of [DVP rows][128 B]
  static constexpr int P_BYTES = 2 * 128 * 128;      // two 64-key sub-tiles of [128 rows][128 B]
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * K_BYTES;
  static constexpr int OFF_P = OFF_V + 2 * V_BYTES;
  static constexpr int OFF_W = OFF_P + NP * P_BYTES;         // fp32 Wo [dv][C] + bo [C] for the fused epilogue
  static constexpr int W_BYTES = (((CEPI / 2) * CEPI + CEPI) * 4 + 127) / 128 * 128;
  static constexpr int OFF_X = OFF_W + W_BYTES;               // row-max exchange between the two column halves
  static constexpr int OFF_BAR = OFF_X + 2 * 2 * 128 * 4;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;          // + alignment slack
  // O is spread over NACC accumulators (key steps ks % NACC): small-N MMAs that accumulate into the SAME TMEM tile run
  // back to back at the full pipeline latency, independent accumulators pipeline; they are summed in the epilogue
  static constexpr int NACC = (NS * 128 + 4 * DVP <= 256 || DVP > 64) ? 4 : 2;
  static constexpr int TMEM_COLS = (NS * 128 + NACC * DVP) <= 256 ? 256 : 512;
  static constexpr int OCOL = NS * 128;
};

// CEPI: channel count C when the out-projection + residual are fused on the CUDA cores (C <= 64), 0 otherwise
template <int DVP, int NS, int NP, int CEPI>
__global__ void __launch_bounds__(TC_THREADS, NS == 1 ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const float* __restrict__ X, const float* __restrict__ Wo,
                   const float* __restrict__ bo, const float* __restrict__ gamma, float* __restrict__ Y,
                   float* __restrict__ lse, float* __restrict__ A_saved, __nv_bfloat16* __restrict__ A_bf16, int N,
                   int Npad, int Nk, int Nkpad, int dv, int kq_steps) {
  // N queries (Npad padded), Nk keys / values (Nkpad padded); Nk == N unless the keys are down-sampled
  using L = FwdSmem<DVP, NS, NP, CEPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + L::OFF_Q;
  uint8_t* sK = smem + L::OFF_K;
  uint8_t* sV = smem + L::OFF_V;
  uint8_t* sP = smem + L::OFF_P;
  float* sW = reinterpret_cast<float*>(smem + L::OFF_W);
  float* sX = reinterpret_cast<float*>(smem + L::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* barQ = bars + 0;
  uint64_t* barKV = bars + 1;   // [2]
  uint64_t* barS = bars + 3;    // [2]
  uint64_t* barPV = bars + 5;   // [2]
  uint64_t* barP = bars + 7;    // [2] 256 arrivals each: P_j is in shared memory (and S_j has been read)
  uint64_t* barSfree = bars + 9;   // 256 arrivals: S_j is in registers (single S buffer: QK_{j+1} may overwrite it)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, qt = blockIdx.x;
  const int nt = Nkpad / 128;

  if (threadIdx.x == 0) {
    mbar_init(barQ, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(barKV + i, 1); mbar_init(barS + i, 1); mbar_init(barPV + i, 1); }
    mbar_init(barP, 256); mbar_init(barP + 1, 256);
    mbar_init(barSfree, 256);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ================================================================ TMA producer
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      mbar_expect_tx(barQ, L::Q_BYTES);
      tma_load_2d(sQ, &tmQ, barQ, 0, b * Npad + qt * 128);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(barPV + s, ((j - 2) >> 1) & 1);   // PV_{j-2} done (=> QK_{j-2} done): stage s is free
        mbar_expect_tx(barKV + s, L::K_BYTES + L::V_BYTES);
        tma_load_2d(sK + s * L::K_BYTES, &tmK, barKV + s, 0, b * Nkpad + j * 128);
        tma_load_2d(sV + s * L::V_BYTES, &tmV, barKV + s, j * 128, b * DVP);
        tma_load_2d(sV + s * L::V_BYTES + DVP * 128, &tmV, barKV + s, j * 128 + 64, b * DVP);
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issuer (one elected thread)
    if (elect_one_sync()) {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128);
      constexpr uint32_t IDESC_O = make_idesc_bf16(128, DVP);
      const uint64_t descQ = L::QKB == 32 ? make_desc_sw32(smem_u32(sQ)) : make_desc_sw128(smem_u32(sQ));
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(barKV + s, (j >> 1) & 1);
        tc_fence_after();
        const uint64_t descK = L::QKB == 32 ? make_desc_sw32(smem_u32(sK + s * L::K_BYTES)) : make_desc_sw128(smem_u32(sK + s * L::K_BYTES));
        const uint32_t d = tmem_base + (uint32_t)((j % NS) * 128);
        for (int ks = 0; ks < kq_steps; ++ks) mma_bf16_ss(d, descQ + (uint64_t)(ks * 2), descK + (uint64_t)(ks * 2), IDESC_S, ks > 0);
        mma_commit(barS + (j % NS));
      };
      auto issue_pv = [&](int j) {
        const int s = j & 1, pb = j % NP;
        const uint64_t descP = make_desc_sw128(smem_u32(sP + pb * L::P_BYTES));
        const uint64_t descV = make_desc_sw128(smem_u32(sV + s * L::V_BYTES));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t a = descP + (uint64_t)((ks >> 2) * ((128 * 128) >> 4) + (ks & 3) * 2);
          const uint64_t bb = descV + (uint64_t)((ks >> 2) * ((DVP * 128) >> 4) + (ks & 3) * 2);
          mma_bf16_ss(tmem_base + L::OCOL + (ks % L::NACC) * DVP, a, bb, IDESC_O, (j > 0) || (ks >= L::NACC));
        }
        mma_commit(barPV + s);
      };
      mbar_wait(barQ, 0);
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        if (NS == 2 && j + 1 < nt) issue_qk(j + 1);   // second S buffer: runs under softmax_j
        if (NS == 1 && j + 1 < nt) {                  // single S buffer: free as soon as every softmax thread holds S_j in
          mbar_wait(barSfree, j & 1);                 // registers; QK_{j+1} then runs under the exponentials of tile j
          tc_fence_after();
          issue_qk(j + 1);
        }
        mbar_wait(barP + (j & 1), (j >> 1) & 1);      // P_j written
        tc_fence_after();
        issue_pv(j);
      }
    }
  } else {
    // ================================================================ softmax warps
    // thread <-> query row (TMEM lane) x one HALF of the tile's 128 key columns: warps w and w + 4 share the lane
    // quarter w & 3.  Eight softmax warps per CTA (four per scheduler with the co-resident CTA) keep the MUFU pipe fed
    // where four left the schedulers idle three cycles out of four.
    const int h = warp >> 2;
    const int row = threadIdx.x & 127;                             // query row inside the tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

    if (CEPI > 0) {
      const int nW = dv * CEPI;
      for (int e = threadIdx.x; e < nW; e += 256) sW[e] = Wo[e];
      for (int e = threadIdx.x; e < CEPI; e += 256) sW[nW + e] = bo[e];
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the epilogue of every softmax thread reads all of sW
    }

    float m_used = -INFINITY;      // integer-valued (log2 units) once set: see "consistent rounding" above
    const bool ragged = (Nk % 128) != 0;

    for (int j = 0; j < nt; ++j) {
      mbar_wait(barS + (j % NS), (j / NS) & 1);
      tc_fence_after();
      const uint32_t t_s = t_row + (uint32_t)((j % NS) * 128);
      const int kvalid = (ragged && j == nt - 1) ? (Nk - j * 128) : 128;   // keys of this tile that exist

      // ---- the whole S row (128 fp32) comes to registers with ONE exposed TMEM round trip
      //      (logits are already in log2 units: Q carries log2(e))
      uint32_t r[64];
      tmem_ld32(t_s + h * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(t_s + h * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_wait_ld();
      if (NS == 1) {
        tc_fence_before();
        mbar_arrive(barSfree);
      }
      if (kvalid < 128) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (h * 64 + i >= kvalid) r[i] = 0xff800000u;   // -inf: masked keys of the ragged last tile
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i]));
        mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // both halves of a row must use the same shift: exchange the partial maxima (double-buffered, one barrier a tile)
      {
        float* xb = sX + (j & 1) * 256;
        xb[h * 128 + row] = mx;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        mx = fmaxf(mx, xb[(h ^ 1) * 128 + row]);
      }
      // ---- lazy rescale of the running accumulators
      const bool need = mx > m_used + 32.0f;   // P and the fp32 accumulators share the exponent range: a stale max costs no precision
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(barPV + ((j - 1) & 1), ((j - 1) >> 1) & 1);       // every PV issued so far has completed
          tc_fence_after();
          const float scale = need ? exp2f(m_used - ceilf(mx)) : 1.0f;   // exact power of two
          static_assert((L::NACC * DVP / 16) % 2 == 0, "the two halves share the O columns evenly");
#pragma unroll
          for (int c = 0; c < L::NACC * DVP / 32; ++c) {     // this half's share of the O columns
            const uint32_t col = L::OCOL + (h * (L::NACC * DVP / 32) + c) * 16;
            uint32_t o[16];
            tmem_ld16(t_row + col, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
            tmem_st16(t_row + col, o);
          }
          tmem_wait_st();
        }
        if (need) m_used = ceilf(mx);
      }
      // ---- P buffer free?  (PV_{j-NP} has finished reading it)
      const int pb = j % NP;
      if (j >= NP) mbar_wait(barPV + ((j - NP) & 1), ((j - NP) >> 1) & 1);
      uint8_t* sPj = sP + pb * L::P_BYTES;
      // ---- P' = bf16(exp2(S - m)), swizzled store (16 B = 8 keys per store); the row sum comes out of the PV MMA
      const f2 neg_m = f2_pack(-m_used, -m_used);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {     // S - m for two columns per FADD2; one column pair in four takes the FMA-pipe 2^x
          const f2 x = f2_add(f2_pack(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1])), neg_m);
          // TC_POLY_EXP = 1: 2 column pairs of 8 on the FMA pipe, 2: 3 of 8, 3: 4 of 8
          if ((TC_POLY_EXP >= 1 && i == 3) || (TC_POLY_EXP == 2 && i == 1 && (g & 1)) || (TC_POLY_EXP >= 3 && i == 1)) {
            pk[i] = ex2_poly2_bf16(x);
          } else {
            float x0, x1;
            f2_unpack(x, x0, x1);
            pk[i] = pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
          }
        }
        // key columns 64 h + [8g, 8g+8): 64-key sub-tile h, 16-byte chunk g
        *reinterpret_cast<uint4*>(sPj + h * (128 * 128) + sw128_offset(row, g)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();      // st.shared of P -> visible to the tensor core (async proxy)
      tc_fence_before();             // orders this thread's tcgen05.ld / st before the arrive
      mbar_arrive(barP + (j & 1));
    }

    // ---- epilogue: A = (P' V) / l', saved tensors, fused output conv + gamma residual
    mbar_wait(barPV + ((nt - 1) & 1), ((nt - 1) >> 1) & 1);
    tc_fence_after();
    const int i_tok = qt * 128 + row;
    const bool valid = i_tok < N;
    const long long grow = (long long)b * N + (valid ? i_tok : 0);
    {
      constexpr int C = CEPI;
      constexpr int DV = C / 2;
      static_assert(DVP >= 2 * DV + 1, "V^T rows are [v_hi | v_lo | ones]");
      float a[DVP];
#pragma unroll
      for (int i = 0; i < DVP; ++i) a[i] = 0.f;
#pragma unroll
      for (int acc = 0; acc < L::NACC; ++acc) {
#pragma unroll
        for (int c = 0; c < DVP / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(t_row + L::OCOL + acc * DVP + c * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) a[c * 16 + i] += __uint_as_float(r[i]);
        }
      }
      const float l = a[2 * DV];           // l' = sum_j P'_ij, accumulated by the MMA through the ones row
      const float inv = 1.0f / l;
#pragma unroll
      for (int v = 0; v < DV; ++v) a[v] = (a[v] + a[DV + v]) * inv;
      if (valid) {
        if (h == 0) {
#pragma unroll
          for (int v = 0; v < DV; v += 4) st4(A_saved + grow * DV + v, make_float4(a[v], a[v + 1], a[v + 2], a[v + 3]));
          lse[grow] = (m_used + log2f(l)) * TC_LN2;
        }
        const float gm = *gamma;
#pragma unroll
        for (int cc = 0; cc < C / 2; cc += 4) {              // each half of the row's threads writes half of the channels
          const int c = h * (C / 2) + cc;
          float o[4] = {sW[DV * C + c], sW[DV * C + c + 1], sW[DV * C + c + 2], sW[DV * C + c + 3]};
#pragma unroll
          for (int v = 0; v < DV; ++v) {
            const float4 w = *reinterpret_cast<const float4*>(&sW[v * C + c]);
            o[0] = fmaf(a[v], w.x, o[0]); o[1] = fmaf(a[v], w.y, o[1]);
            o[2] = fmaf(a[v], w.z, o[2]); o[3] = fmaf(a[v], w.w, o[3]);
          }
          const float4 xx = ld4(X + grow * C + c);
          st4(Y + grow * C + c,
              make_float4(fmaf(gm, o[0], xx.x), fmaf(gm, o[1], xx.y), fmaf(gm, o[2], xx.z), fmaf(gm, o[3], xx.w)));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ flash forward, four chains
// C <= 32 (in-model shapes), PERSISTENT: one CTA per SM walks the (sample, 128-query tile) work items w = blockIdx.x,
// blockIdx.x + gridDim.x, ...
//
// The kernel above runs every softmax thread of a CTA through the same phases at the same time (one S barrier, one
// row-max exchange, one P barrier per tile) and pays ~5 us of prologue / epilogue per CTA.  Here the 128 key columns of
// a tile are four CHAINS of 32 columns, each with its own S slice, running max and O accumulator:
//   16 softmax warps: warp w <-> chain w >> 2, TMEM lane quarter w & 3; every scheduler hosts one warp of each chain
//   S_q(g) = Q K_g[32 q .. 32 q + 32)^T        (M = 128, N = 32)  -> TMEM, NS buffers per chain
//   P'_q(g) = bf16(exp2(S - m_q)) overwrites the first 16 columns of S_q(g) IN TMEM and is the A operand of
//   O_q += P'_q(g) V_g[32 q .. )               (tcgen05.mma, A from TMEM): no shared-memory P tile, no proxy fence
// No thread exchanges anything with another thread inside the key loop, and each waits on ONE barrier per tile.
// A thread keeps ONE 32-column chunk in registers: as soon as a 16-column half has been turned into P', the same
// registers take that half of S_q(g+1) (with NS = 3 -- C = 16: 3 x 128 + 4 x 32 TMEM columns -- that tile has been
// complete since the previous one), so the TMEM round trip runs under the other half's exponentials.  (Two full
// register sets, with the next row maximum computed under the current exponentials, measured the same per tile but
// spilled at the 96 registers 608 threads leave.)
// g counts key tiles across work items: the K / V rings, the S buffers and every barrier phase run on, so the TMA
// producer and the MMA issuers are already loading and multiplying the next item's first tiles while the softmax
// warps run the epilogue of the current one (fit of measured times over tiles-per-CTA: 5.3 us per non-persistent CTA
// besides 0.63 us per tile).
// The tensor pipe executes in issue order: QK_q(g+NS), issued after PV_q(g), cannot overwrite the P' columns early,
// and "S_q(g+NS) ready" implies "PV_q(g) complete" (used by the lazy rescale and the epilogue).  m_q is integer-valued
// per chain ("consistent rounding" holds per chain: bf16 rounding commutes with powers of two); the epilogue
// combines the four accumulators with exact power-of-two factors.
// What bounds it (tools/ubench/softmax_mix.cu, softmax_tmem.cu; 4 warps x 32 columns per scheduler = one tile):
// MUFU alone 768 cycles at 3/4 of the exponentials, the whole register-only mix 866, with the TMEM load / store and
// barrier operations of the tile body ~1000; the kernel runs at ~1200 cycles per tile plus 2.2 us per work item
// (removing every MMA changes that by 7 %: it is the softmax warps' own instruction stream, not the tensor pipe).
constexpr int TC4_THREADS = 608;     // 16 softmax warps + TMA producer + 2 MMA issuers (chains {0,1} and {2,3})
constexpr int TC4_NK = 6, TC4_NV = 3;

template <int DVP, int CEPI>
struct Fwd4Smem {
  static constexpr int QKB = qk_cols(CEPI) * 2;
  static_assert(QKB == 32, "the chained kernel is built for C <= 32 (one MMA K step of split logits)");
  static constexpr int NS = (3 * 128 + 4 * DVP <= 512) ? 3 : 2;      // S buffers per chain
  static constexpr int Q_BYTES = 128 * QKB;
  static constexpr int K_BYTES = 128 * QKB;
  static constexpr int V_BYTES = 2 * DVP * 128;
  static constexpr int OFF_Q = 0;                             // two buffers (work items k, k + 1)
  static constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_V = (OFF_K + TC4_NK * K_BYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_W = OFF_V + TC4_NV * V_BYTES;
  static constexpr int W_BYTES = (((CEPI / 2) * CEPI + CEPI) * 4 + 127) / 128 * 128;
  static constexpr int OFF_X = OFF_W + W_BYTES;               // m_q of the four chains, [item parity][4][128]
  static constexpr int OFF_BAR = OFF_X + 2 * 4 * 128 * 4;
  static constexpr int TOTAL = OFF_BAR + 512 + 1024;
  static constexpr int OCOL = NS * 128;                       // S buffers, then one accumulator per chain
  static_assert(OCOL + 4 * DVP <= 512, "TMEM columns");
  static_assert(V_BYTES % 1024 == 0, "V tiles must stay 1024-byte aligned");
  static_assert(TC4_NK >= NS + 1, "K ring");
};

template <int DVP, int CEPI>
__global__ void __launch_bounds__(TC4_THREADS, 1)
attn_fwd_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const float* __restrict__ X, const float* __restrict__ Wo,
                    const float* __restrict__ bo, const float* __restrict__ gamma, float* __restrict__ Y,
                    float* __restrict__ lse, float* __restrict__ A_saved, int N, int Npad, int Nk, int Nkpad, int dv,
                    int nitems) {
  using L = Fwd4Smem<DVP, CEPI>;
  constexpr int NS = L::NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + L::OFF_Q;
  uint8_t* sK = smem + L::OFF_K;
  uint8_t* sV = smem + L::OFF_V;
  float* sW = reinterpret_cast<float*>(smem + L::OFF_W);
  float* sX = reinterpret_cast<float*>(smem + L::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* barQfull = bars + 0;                 // [2]
  uint64_t* barQfree = bars + 2;                 // [2] both issuers: the last QK MMA of the item has been issued and is done
  uint64_t* barKfull = bars + 4;                 // [NK]
  uint64_t* barKfree = barKfull + TC4_NK;        // [NK] both issuers: the QK MMAs that read the stage are done
  uint64_t* barVfull = barKfree + TC4_NK;        // [NV]
  uint64_t* barVfree = barVfull + TC4_NV;        // [NV] both issuers: the PV MMAs that read the stage are done
  uint64_t* barS = barVfree + TC4_NV;            // [chain][NS] S_q(g) ready (=> PV_q(g - NS) complete), buffer g % NS
  uint64_t* barP = barS + 4 * NS;                // [chain][NS] 128 arrivals: P'_q(g) is in TMEM
  uint64_t* barOfree = barP + 4 * NS;            // 512 arrivals: the item's accumulators are in registers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(barOfree + 1);

  const int warp = threadIdx.x >> 5;
  const int nt = Nkpad / 128;                    // key tiles per item
  const int nq = Npad / 128;                     // query tiles per sample
  const int my_items = blockIdx.x < nitems ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(barQfull + i, 1); mbar_init(barQfree + i, 2); }
    for (int i = 0; i < TC4_NK; ++i) { mbar_init(barKfull + i, 1); mbar_init(barKfree + i, 2); }
    for (int i = 0; i < TC4_NV; ++i) { mbar_init(barVfull + i, 1); mbar_init(barVfree + i, 2); }
    for (int i = 0; i < 4 * NS; ++i) { mbar_init(barS + i, 1); mbar_init(barP + i, 128); }
    mbar_init(barOfree, 512);
    mbar_fence_init();
  }
  if (warp == 16) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 16) {
    // ================================================================ TMA producer: Q per item, K ring and V ring
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      // positions of the next K tile and the next V tile to load: (item index k, tile j inside it), ring stage and use count
      int kk = 0, kj = 0, ks = 0, ku = 0;
      int vk = 0, vj = 0, vs = 0, vu = 0;
      while (kk < my_items || vk < my_items) {
        bool moved = false;
        if (kk < my_items) {
          bool ok = ku == 0 || mbar_try_wait(barKfree + ks, (ku - 1) & 1);
          // the item's Q tile goes first, into buffer kk & 1 (free once the QK MMAs of item kk - 2 are done)
          if (ok && kj == 0 && kk >= 2) ok = mbar_try_wait(barQfree + (kk & 1), ((kk >> 1) - 1) & 1);
          if (ok) {
            const int w = (int)blockIdx.x + kk * (int)gridDim.x;
            const int b = w / nq, qt = w - b * nq;
            if (kj == 0) {
              mbar_expect_tx(barQfull + (kk & 1), L::Q_BYTES);
              tma_load_2d(sQ + (kk & 1) * L::Q_BYTES, &tmQ, barQfull + (kk & 1), 0, b * Npad + qt * 128);
            }
            mbar_expect_tx(barKfull + ks, L::K_BYTES);
            tma_load_2d(sK + ks * L::K_BYTES, &tmK, barKfull + ks, 0, b * Nkpad + kj * 128);
            if (++kj == nt) { kj = 0; ++kk; }
            if (++ks == TC4_NK) { ks = 0; ++ku; }
            moved = true;
          }
        }
        if (vk < my_items) {
          if (vu == 0 || mbar_try_wait(barVfree + vs, (vu - 1) & 1)) {
            const int w = (int)blockIdx.x + vk * (int)gridDim.x;
            const int b = w / nq;
            mbar_expect_tx(barVfull + vs, L::V_BYTES);
            tma_load_2d(sV + vs * L::V_BYTES, &tmV, barVfull + vs, vj * 128, b * DVP);
            tma_load_2d(sV + vs * L::V_BYTES + DVP * 128, &tmV, barVfull + vs, vj * 128 + 64, b * DVP);
            if (++vj == nt) { vj = 0; ++vk; }
            if (++vs == TC4_NV) { vs = 0; ++vu; }
            moved = true;
          }
        }
        if (!moved) __nanosleep(64);
      }
    }
  } else if (warp >= 17) {
    // ================================================================ MMA issuers: warp 17 chains {0, 1}, warp 18 chains {2, 3}
    // The WHOLE warp runs the loop (uniform control flow, ring positions kept as wrapping counters, addresses in the
    // uniform datapath); only the tcgen05 instructions are issued by the elected lane.  With the loop inside the
    // elected thread the address arithmetic (divisions by the ring sizes, R2UR moves) cost ~70 cycles per MMA and the
    // issuer, not the softmax, set the tile time (clock64 timeline: 565 cycles of issue per tile and issuer).
    {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, 32);
      constexpr uint32_t IDESC_O = make_idesc_bf16(128, DVP);
      const bool leader = elect_one_sync();
      const int q0 = (warp - 17) * 2;
      const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK) + q0 * (32 * L::QKB), sVa = smem_u32(sV);
      const int total = my_items * nt;
      // look-ahead position: the tile whose QK MMAs are issued next (item nk, tile nj, K stage sk, S buffer nb)
      int nk = 0, nj = 0, sk = 0, nb = 0;
      uint32_t pk = 0;
      int gn = 0;
      auto issue_qk_next = [&]() {       // S_q(gn) = Q_nk K_gn[32 q ..)^T for this issuer's two chains; commits NOT included
        if (nj == 0) mbar_wait(barQfull + (nk & 1), (nk >> 1) & 1);
        mbar_wait(barKfull + sk, pk);
        tc_fence_after();
      };
      auto advance_next = [&]() {
        if (leader) {
          mma_commit(barKfree + sk);
          if (nj == nt - 1) mma_commit(barQfree + (nk & 1));
        }
        if (++nj == nt) { nj = 0; ++nk; }
        if (++sk == TC4_NK) { sk = 0; pk ^= 1; }
        if (++nb == NS) nb = 0;
        ++gn;
      };
      for (; gn < NS && gn < total;) {   // prologue: the first NS tiles
        issue_qk_next();
        if (leader) {
          const uint64_t descQ = make_desc_sw32(sQa + (nk & 1) * L::Q_BYTES);
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {
            mma_bf16_ss(tmem_base + (uint32_t)(nb * 128 + (q0 + qq) * 32), descQ,
                        make_desc_sw32(sKa + sk * L::K_BYTES + qq * (32 * L::QKB)), IDESC_S, false);
            mma_commit(barS + (q0 + qq) * NS + nb);
          }
        }
        advance_next();
      }
      // position of the tile whose PV MMAs are issued: item k, tile j, V stage sv, S buffer sb (phases pv, pb)
      int k = 0, j = 0, sv = 0, sb = 0;
      uint32_t pv = 0, pb = 0;
      for (int g = 0; g < total; ++g) {
        const bool has_next = gn < total;
        mbar_wait(barVfull + sv, pv);
        if (has_next) issue_qk_next();
        if (j == 0 && k > 0) mbar_wait(barOfree, (k - 1) & 1);     // the previous item's accumulators have been read
        tc_fence_after();
        const uint32_t vbase = sVa + sv * L::V_BYTES;
        const uint64_t descQ = make_desc_sw32(sQa + (nk & 1) * L::Q_BYTES);
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = q0 + qq;
          mbar_wait(barP + q * NS + sb, pb);                   // P'_q(g) is in TMEM
          tc_fence_after();
          if (leader) {
            const uint32_t t_s = tmem_base + (uint32_t)(sb * 128 + q * 32);
            // key steps 2q, 2q+1 of the tile: 64-key sub-tile q >> 1, 32-byte column blocks 2 (q & 1) and 2 (q & 1) + 1
            const uint32_t vq = vbase + (q >> 1) * (DVP * 128) + (q & 1) * 64;
            mma_bf16_ts_g<1>(tmem_base + L::OCOL + q * DVP, t_s, make_desc_sw128(vq), IDESC_O, j > 0);
            mma_bf16_ts_g<1>(tmem_base + L::OCOL + q * DVP, t_s + 8, make_desc_sw128(vq + 32), IDESC_O, true);
            if (has_next)      // S_q(g + NS) into the same buffer: ordered behind the two MMAs that read P' from it
              mma_bf16_ss(t_s, descQ, make_desc_sw32(sKa + sk * L::K_BYTES + qq * (32 * L::QKB)), IDESC_S, false);
            mma_commit(barS + q * NS + sb);                    // "S_q(g+NS) ready" == "PV_q(g) complete" when there is no such tile
          }
        }
        if (leader) mma_commit(barVfree + sv);
        if (has_next) advance_next();
        if (++j == nt) { j = 0; ++k; }
        if (++sv == TC4_NV) { sv = 0; pv ^= 1; }
        if (++sb == NS) { sb = 0; pb ^= 1; }
      }
    }
  } else {
    // ================================================================ softmax warps
    // (state that lives across the key loop is kept to a handful of 32-bit registers: two 32-column S chunks plus the
    //  packed P' half already take 72 of the 104 a thread can have)
    const int q = warp >> 2;
    const uint32_t t_chain = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(q * 32);   // S_q, buffer 0, this lane quarter
    const uint32_t bS = smem_u32(barS + q * NS);                    // barS[q][0]; barP[q][0] is 4 NS barriers further
    constexpr uint32_t BP = 4 * NS * 8;
    {
      const int nW = dv * CEPI;
      for (int e = threadIdx.x; e < nW; e += 512) sW[e] = Wo[e];
      for (int e = threadIdx.x; e < CEPI; e += 512) sW[nW + e] = bo[e];
    }
    const int kv_last = Nk - (nt - 1) * 128 - q * 32;                // columns of this chunk that exist in the last key tile
    uint32_t r[32];                // S_q of the current tile; each half is refilled with the next tile's as soon as it is used
    uint32_t sb = 0, pb = 0;       // S buffer of the current tile (tiles are counted across items) and its phase
    float m_used;                  // integer-valued (log2 units) once set

    // chunk maximum; `last`: the tile is the item's last one, whose missing keys (ragged Nk) count as -inf
    auto row_max = [&](bool last) -> float {
      if (last && kv_last < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= kv_last) r[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
      }
      return fmaxf(mx0, mx1);
    };
    auto landed = [&]() {
      tmem_wait_ld();
      // the loaded values exist from here on: keep the compiler from touching the registers before the wait
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
    };
    auto exps = [&](int e, float nm, uint32_t (&pk)[8]) {   // columns 2e, 2e+1: S - m with one FADD2; one pair in four on the FMA pipe
      const f2 x = f2_add(f2_pack(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), f2_pack(nm, nm));
      if (TC_POLY_EXP && (e & 3) == 3) {
        pk[e & 7] = ex2_poly2_bf16(x);
      } else {
        float x0, x1;
        f2_unpack(x, x0, x1);
        pk[e & 7] = pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
      }
    };

    for (int k = 0; k < my_items; ++k) {
      m_used = -1.0e30f;           // finite, so that all-masked chunks give P' = 0
      mbar_wait_a(bS + sb * 8, pb);
      tc_fence_after();
      tmem_ld32(t_chain + sb * 128, r);
      landed();
      float mx = row_max(nt == 1);
      for (int rem = nt; rem > 0; --rem) {        // `rem` key tiles of the item are left including this one
        const uint32_t sbn = sb + 1 == NS ? 0 : sb + 1;
        const uint32_t pbn = sb + 1 == NS ? pb ^ 1 : pb;
        // ---- lazy rescale of this chain's accumulator row (P' and the fp32 accumulators share the exponent range)
        const bool need = mx > m_used + 32.0f;
        if (__any_sync(0xffffffffu, need)) {
          if (rem < nt) {            // not the item's first tile: there is something to rescale
            // "S of the tile NS after the previous one is ready" => PV_q of the previous tile and all before it are complete
            mbar_wait_a(bS + (sb == 0 ? NS - 1 : sb - 1) * 8, (sb == 0 ? pb ^ 1 : pb) ^ 1);
            tc_fence_after();
            const float scale = need ? exp2f(m_used - ceilf(mx)) : 1.0f;     // exact power of two (0 from the initial value)
            const uint32_t t_acc = t_chain - (uint32_t)(q * 32) + L::OCOL + q * DVP;
#pragma unroll
            for (int c = 0; c < DVP / 16; ++c) {
              uint32_t o[16];
              tmem_ld16(t_acc + c * 16, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
              tmem_st16(t_acc + c * 16, o);
            }
            tmem_wait_st();
          }
          if (need) m_used = ceilf(mx);
        }
        const float nm = -m_used;
        const bool more = rem > 1;
        {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) exps(e, nm, pk);
          tmem_st8(t_chain + sb * 128, pk);              // P' over the first 16 columns of S_q(g), stored in halves
        }
        // ---- the registers of the half just used take the same half of S_q(g+1) (with three buffers it has been complete
        //      since the previous tile; with two it follows PV_q(g-1) and is ready about now)
        if (more) {
          mbar_wait_a(bS + sbn * 8, pbn);
          tc_fence_after();
          tmem_ld16(t_chain + sbn * 128, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
        }
        {
          uint32_t pk[8];
#pragma unroll
          for (int e = 8; e < 16; ++e) exps(e, nm, pk);
          tmem_st8(t_chain + sb * 128 + 8, pk);
        }
        if (more) tmem_ld16(t_chain + sbn * 128 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_a(bS + BP + sb * 8);
        if (more) {
          landed();
          mx = row_max(rem == 2);
        }
        sb = sbn;
        pb = pbn;
      }

      // ---- epilogue: the four chains' (m_q, O_q) of a row -> A = sum_q 2^(m_q - m) O_q / l', saved tensors,
      //      fused output conv + gamma residual (each of the row's four threads writes a quarter of the channels)
      constexpr int C = CEPI;
      constexpr int DV = C / 2;
      static_assert(DVP >= 2 * DV + 1, "V^T rows are [v_hi | v_lo | ones]");
      const int row = (warp & 3) * 32 + (threadIdx.x & 31);         // query row inside the tile == TMEM lane
      const uint32_t t_row = t_chain - (uint32_t)(q * 32);
      const int w = (int)blockIdx.x + k * (int)gridDim.x;
      const int b = w / nq, qt = w - b * nq;
      const int i_tok = qt * 128 + row;
      const bool valid = i_tok < N;
      const long long grow = (long long)b * N + (valid ? i_tok : 0);
      float4 xx[C / 16];
#pragma unroll
      for (int cc = 0; cc < C / 16; ++cc) xx[cc] = ld4(X + grow * C + q * (C / 4) + cc * 4);   // in flight under the TMEM reads
      float* sXk = sX + (k & 1) * 512;
      sXk[q * 128 + row] = m_used;
      mbar_wait_a(bS + (sb == 0 ? NS - 1 : sb - 1) * 8, (sb == 0 ? pb ^ 1 : pb) ^ 1);   // PV_q of the item's last tile is complete
      tc_fence_before();
      asm volatile("bar.sync 1, 512;" ::: "memory");                     // also orders the sW fill above
      tc_fence_after();
      const float m0 = sXk[row], m1 = sXk[128 + row], m2 = sXk[256 + row], m3 = sXk[384 + row];
      const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      const float sc[4] = {exp2f(m0 - m), exp2f(m1 - m), exp2f(m2 - m), exp2f(m3 - m)};
      float a[DVP];
#pragma unroll
      for (int i = 0; i < DVP; ++i) a[i] = 0.f;
#pragma unroll
      for (int acc = 0; acc < 4; ++acc) {
#pragma unroll
        for (int c = 0; c < DVP / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(t_row + L::OCOL + acc * DVP + c * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) a[c * 16 + i] = fmaf(__uint_as_float(r[i]), sc[acc], a[c * 16 + i]);
        }
      }
      tc_fence_before();
      mbar_arrive(barOfree);               // the next item's first PV MMAs may overwrite the accumulators
      const float l = a[2 * DV];           // l' = sum_j P'_ij, accumulated by the MMA through the ones row
      const float inv = 1.0f / l;
#pragma unroll
      for (int v = 0; v < DV; ++v) a[v] = (a[v] + a[DV + v]) * inv;
      if (valid) {
        if (q == 0) {
#pragma unroll
          for (int v = 0; v < DV; v += 4) st4(A_saved + grow * DV + v, make_float4(a[v], a[v + 1], a[v + 2], a[v + 3]));
          lse[grow] = (m + log2f(l)) * TC_LN2;
        }
        const float gm = *gamma;
#pragma unroll
        for (int cc = 0; cc < C / 4; cc += 4) {
          const int c = q * (C / 4) + cc;
          float o[4] = {sW[DV * C + c], sW[DV * C + c + 1], sW[DV * C + c + 2], sW[DV * C + c + 3]};
#pragma unroll
          for (int v = 0; v < DV; ++v) {
            const float4 wv = *reinterpret_cast<const float4*>(&sW[v * C + c]);
            o[0] = fmaf(a[v], wv.x, o[0]); o[1] = fmaf(a[v], wv.y, o[1]);
            o[2] = fmaf(a[v], wv.z, o[2]); o[3] = fmaf(a[v], wv.w, o[3]);
          }
          const float4 x4 = xx[cc / 4];
          st4(Y + grow * C + c,
              make_float4(fmaf(gm, o[0], x4.x), fmaf(gm, o[1], x4.y), fmaf(gm, o[2], x4.z), fmaf(gm, o[3], x4.w)));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host side
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct TcLayout {
  int Npad, Nk, Nkpad, DVP, kq_steps;
  size_t off_q, off_k, off_v, total;
};

// Nk = number of keys / values (N, or N / 4 when they are down-sampled)
static TcLayout tc_layout(int B, int N, int Nk, int C) {
  TcLayout t;
  const int d = C / 8, dv = C / 2;
  t.Npad = round_up(N, 128);
  t.Nk = Nk;
  t.Nkpad = round_up(Nk, 128);
  t.DVP = round_up(2 * dv + 1, 16);   // V^T rows [v_hi | v_lo | ones | 0 ...]
  t.kq_steps = (3 * d + 15) / 16;   // split-bf16 logits: [hi|lo|hi] x [hi|hi|lo]
  const size_t T = (size_t)B * t.Npad, Tk = (size_t)B * t.Nkpad;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 1023) / 1024 * 1024; return r; };
  t.off_q = take(T * qk_cols(C) * 2);
  t.off_k = take(Tk * qk_cols(C) * 2);
  t.off_v = take((size_t)B * t.DVP * t.Nkpad * 2);
  t.total = o + 1024;
  return t;
}

// attn_tc_big.cu: C in {128, 256, 512}
bool attn_tc_big_supported(int N, int C);
size_t attn_tc_big_workspace_bytes(int B, int N, int C);
int attn_tc_big_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                    const float* bv, const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A,
                    int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st);

size_t attn_tc_workspace_bytes(int B, int N, int C) {
  if (C > 64) return attn_tc_big_workspace_bytes(B, N, C);
  return tc_layout(B, N, N, C).total;      // upper bound of the down-sampled layout as well
}

template <int DVP, int NS, int NP, int CEPI>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const float* X,
                      const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A,
                      __nv_bfloat16* Ab, int B, int N, int Npad, int Nk, int Nkpad, int dv, int kq_steps, cudaStream_t st) {
  using L = FwdSmem<DVP, NS, NP, CEPI>;
  auto kern = attn_fwd_tc_kernel<DVP, NS, NP, CEPI>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<dim3(Npad / 128, B), TC_THREADS, L::TOTAL, st>>>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, Ab, N, Npad, Nk, Nkpad,
                                                          dv, kq_steps);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

template <int DVP, int CEPI>
static int launch_fwd4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const float* X,
                       const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A, int B, int N,
                       int Npad, int Nk, int Nkpad, int dv, cudaStream_t st) {
  using L = Fwd4Smem<DVP, CEPI>;
  auto kern = attn_fwd_tc4_kernel<DVP, CEPI>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int nitems = B * (Npad / 128);
  const int grid = nitems < num_sms() ? nitems : num_sms();      // persistent: one CTA per SM
  kern<<<grid, TC4_THREADS, L::TOTAL, st>>>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, N, Npad, Nk, Nkpad, dv, nitems);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// PH, PW > 0: keys / values max-pooled 2x2 / stride 2 over the [PH, PW] token grid (N == PH * PW), else PH = PW = 0
int attn_tc_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                const float* bv, const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A,
                int B, int N, int C, int PH, int PW, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool pool = PH > 0;
  if (C > 64) {
    if (pool) {
      set_err("sagan_attn_pool_fwd: down-sampled keys / values are built for C in {8,16,32,64} (C=%d)", C);
      return SAGAN_EUNSUPPORTED;
    }
    return attn_tc_big_fwd(X, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, Y, lse, A, B, N, C, ws, ws_bytes, st);
  }
  if (!(C == 16 || C == 32 || C == 64)) {
    set_err("sagan_attn_fwd: BF16_TC supports C in {16,32,64} (small-d kernel) and {128,256,512} (C=%d)", C);
    return SAGAN_EUNSUPPORTED;
  }
  const TcLayout t = tc_layout(B, N, pool ? N / 4 : N, C);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* Qb = reinterpret_cast<__nv_bfloat16*>(base + t.off_q);
  __nv_bfloat16* Kb = reinterpret_cast<__nv_bfloat16*>(base + t.off_k);
  __nv_bfloat16* Vt = reinterpret_cast<__nv_bfloat16*>(base + t.off_v);
  const long long Tp = (long long)B * t.Npad, Tkp = (long long)B * t.Nkpad;
  const unsigned pb = (unsigned)ceil_div<long long>(Tp, 128), pkb = (unsigned)ceil_div<long long>(Tkp, 128);
#define SAGAN_PROJ_CASE(CC)                                                                                              \
  case CC:                                                                                                              \
    if (pool) {                                                                                                         \
      attn_proj_tc_kernel<CC, false><<<pb, 128, 0, st>>>(X, Wq, bq, Wk, bk, Wv, bv, Qb, Kb, Vt, B, N, t.Npad);           \
      SAGAN_LAUNCH_CHECK();                                                                                             \
      attn_pool_proj_tc_kernel<CC><<<pkb, 128, 0, st>>>(X, Wk, bk, Wv, bv, Kb, Vt, B, PH, PW, t.Nk, t.Nkpad);            \
    } else {                                                                                                            \
      attn_proj_tc_kernel<CC, true><<<pb, 128, 0, st>>>(X, Wq, bq, Wk, bk, Wv, bv, Qb, Kb, Vt, B, N, t.Npad);            \
    }                                                                                                                   \
    break;
  switch (C) {
    SAGAN_PROJ_CASE(16)
    SAGAN_PROJ_CASE(32)
    SAGAN_PROJ_CASE(64)
  }
#undef SAGAN_PROJ_CASE
  SAGAN_LAUNCH_CHECK();
  CUtensorMap tq, tk, tv;
  int rc;
  const int qkc = qk_cols(C);
  if ((rc = make_tmap_bf16_2d(&tq, Qb, (uint64_t)Tp, qkc, qkc * 2, 128, qkc, qkc == 16 ? 32 : 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, Kb, (uint64_t)Tkp, qkc, qkc * 2, 128, qkc, qkc == 16 ? 32 : 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, Vt, (uint64_t)B * t.DVP, (uint64_t)t.Nkpad, (uint64_t)t.Nkpad * 2, (uint32_t)t.DVP))) return rc;
  const int dv = C / 2;
  static const bool chained = getenv("SAGAN_FWD_UNCHAINED") == nullptr;      // diagnostics only: the round-1 kernel for C <= 32
  if (chained && C == 16) return launch_fwd4<32, 16>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, B, N, t.Npad, t.Nk, t.Nkpad, dv, st);
  if (chained && C == 32) return launch_fwd4<48, 32>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, B, N, t.Npad, t.Nk, t.Nkpad, dv, st);
  switch (C) {
    case 16: return launch_fwd<32, 1, 2, 16>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, nullptr, B, N, t.Npad, t.Nk, t.Nkpad, dv, t.kq_steps, st);
    case 32: return launch_fwd<48, 1, 2, 32>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, nullptr, B, N, t.Npad, t.Nk, t.Nkpad, dv, t.kq_steps, st);
    case 64: return launch_fwd<80, 1, 1, 64>(tq, tk, tv, X, Wo, bo, gamma, Y, lse, A, nullptr, B, N, t.Npad, t.Nk, t.Nkpad, dv, t.kq_steps, st);
  }
  return SAGAN_EUNSUPPORTED;
}

}  // namespace sagan
