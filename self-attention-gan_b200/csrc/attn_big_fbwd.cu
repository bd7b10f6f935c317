// Self-attention block, BF16_TC math mode, LARGE channel counts (C = 128 / 256 / 512): FUSED flash backward.
//
// Replaces the score-shaped part of the composed backward (attn_big_bwd.cu, round 1: S, P, dP, dS of one sample at a
// time as [N, N] fp32 tensors in HBM, ~14 launches per sample) by TWO launches of one kernel template; the [N, N] maps
// never leave the SM (formulas: SURVEY.md §8a row 2; the reference differentiates /root/reference/layers.py:108-117
// with tf.GradientTape):
//
//   DKV = true   CTA = 128 KEYS of one sample (TMEM lanes = keys), streams 64-QUERY tiles:
//                  S^T = K Q_i^T, dP^T = V dA_i^T                    -> TMEM (fp32)
//                  P^T = exp2(S^T - lse2[query]), dS^T = P^T (dP^T - D[query])   -> bf16, back into TMEM
//                  dV += P^T dA_i, dK += dS^T Q_i                    (A operands from TMEM, accumulators in TMEM)
//   DKV = false  CTA = 128 QUERIES (TMEM lanes = queries), streams 64-KEY tiles:
//                  S = Q K_j^T, dP = dA V_j^T, P, dS as above (row-wise lse2 / D), dQ += dS K_j
//
// Both are the same program: "resident" operand tiles R1 [128][64] (K or Q rows) and R2 [128][dv] (V or dA rows),
// "streamed" tiles S1 [64][64] and S2 [64][dv] (the other side), two score-shaped MMAs (A = R, B = S, both K-major),
// an elementwise stage on 8 warps, and one or two accumulating MMAs whose A operand is the bf16 P / dS tile in TMEM
// and whose B operand is the SAME streamed tile read MN-major (no transposed copies in HBM or shared memory).
// S is recomputed twice (once per launch): 2 N^2 exp2 and (5 d + 3 dv) MMA columns per score against the minimal
// (3 d + 2 dv), which buys accumulators that need no atomics and fit the 512 TMEM columns at dv = 256:
//   S 64 | dP 64 | P 32 | dS 32 | acc1 (dK or dQ) 64 | acc2 (dV) 256.
// P and dS have their own columns, so the score MMAs of tile i+1 are issued as soon as the elementwise warps hold tile
// i in registers and run under its exponentials; the tensor pipe executes in issue order.
//
// Warps: 0-7 elementwise (lane quarter w & 3, score columns [32 h, 32 h + 32), h = w >> 2), 8 = TMA producer,
// 9 = MMA issuer / TMEM owner.  Operands are bf16: Q (pre-scaled by log2 e) / K rows padded to 64 columns, V / dA rows of
// dv columns, all written by the CTA-pair GEMMs of gemm_tc.cu; lse2 = lse log2 e and D = rowsum(dA * A) per token.
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace sagan {

using namespace tc;

constexpr int FB_THREADS = 320;
constexpr int FB_NS = 3;                   // stages of the streamed-tile ring
constexpr float FB_LN2 = 0.6931471805599453f;

template <int DV>
struct FbSmem {
  static constexpr int R1_BYTES = 128 * 128;
  static constexpr int R2_BYTES = (DV / 64) * 128 * 128;
  static constexpr int S1_BYTES = 64 * 128;
  static constexpr int S2_BYTES = (DV / 64) * 64 * 128;
  static constexpr int VEC_BYTES = 1024;                      // lse2[64] | D[64] (DKV only), padded to the tile alignment
  static constexpr int STAGE = S1_BYTES + S2_BYTES + VEC_BYTES;
  static constexpr int OFF_R1 = 0;
  static constexpr int OFF_R2 = OFF_R1 + R1_BYTES;
  static constexpr int OFF_ST = OFF_R2 + R2_BYTES;
  static constexpr int OFF_BAR = OFF_ST + FB_NS * STAGE;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static constexpr int S_COL = 0, DP_COL = 64, P_COL = 128, DS_COL = 160, ACC1_COL = 192, ACC2_COL = 256;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(ACC2_COL + DV <= 512, "TMEM budget");
};

template <int DV, int D, bool DKV>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_big_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                    const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmS2,
                    const float* __restrict__ lse2, const float* __restrict__ Dd, const float* __restrict__ gamma,
                    float* __restrict__ out1, float* __restrict__ out2, int ld, int N) {
  // out1 (dK or dQ: D columns) and out2 (dV: DV columns) are column slices of one [T][ld] fp32 matrix [dQ | dK | dV]
  using L = FbSmem<DV>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sR1 = smem + L::OFF_R1;
  uint8_t* sR2 = smem + L::OFF_R2;
  uint8_t* sST = smem + L::OFF_ST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* barR = bars + 0;               // resident tiles landed
  uint64_t* barFull = bars + 1;            // [FB_NS] streamed stage landed
  uint64_t* barFree = barFull + FB_NS;     // [FB_NS] accumulating MMAs of the stage's tile done
  uint64_t* barScore = barFree + FB_NS;    // S / dP of tile i ready
  uint64_t* barSfree = barScore + 1;       // 8 warp arrivals: S / dP of tile i are in registers
  uint64_t* barPD = barSfree + 1;          // 8 warp arrivals: P / dS of tile i are in TMEM
  uint64_t* barAcc = barPD + 1;            // accumulating MMAs of tile i done (P / dS columns free)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(barAcc + 1);

  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y, rt = blockIdx.x;
  const int nst = N / 64;
  const long long row0 = (long long)b * N;

  if (threadIdx.x == 0) {
    mbar_init(barR, 1);
    for (int i = 0; i < FB_NS; ++i) { mbar_init(barFull + i, 1); mbar_init(barFree + i, 1); }
    mbar_init(barScore, 1); mbar_init(barSfree, 8); mbar_init(barPD, 8); mbar_init(barAcc, 1);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ================================================================ TMA producer
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmR1); tma_prefetch_desc(&tmR2); tma_prefetch_desc(&tmS1); tma_prefetch_desc(&tmS2);
      mbar_expect_tx(barR, L::R1_BYTES + L::R2_BYTES);
      tma_load_2d(sR1, &tmR1, barR, 0, (int)(row0 + rt * 128));
#pragma unroll
      for (int sl = 0; sl < DV / 64; ++sl) tma_load_2d(sR2 + sl * (128 * 128), &tmR2, barR, sl * 64, (int)(row0 + rt * 128));
      for (int i = 0; i < nst; ++i) {
        const int s = i % FB_NS;
        if (i >= FB_NS) mbar_wait(barFree + s, ((i / FB_NS) - 1) & 1);
        uint8_t* st = sST + s * L::STAGE;
        mbar_expect_tx(barFull + s, L::S1_BYTES + L::S2_BYTES + (DKV ? 512 : 0));
        tma_load_2d(st, &tmS1, barFull + s, 0, (int)(row0 + i * 64));
#pragma unroll
        for (int sl = 0; sl < DV / 64; ++sl)
          tma_load_2d(st + L::S1_BYTES + sl * (64 * 128), &tmS2, barFull + s, sl * 64, (int)(row0 + i * 64));
        if (DKV) {
          bulk_load_1d(st + L::S1_BYTES + L::S2_BYTES, lse2 + row0 + i * 64, 256, barFull + s);
          bulk_load_1d(st + L::S1_BYTES + L::S2_BYTES + 256, Dd + row0 + i * 64, 256, barFull + s);
        }
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issuer
    if (elect_one_sync()) {
      constexpr uint32_t IDESC_SC = make_idesc_bf16(128, 64);                 // score-shaped: K-major x K-major
      // acc1: B = S1 read MN-major over its whole 64-column atom (Q / K rows are zero-padded beyond d, so the extra
      // accumulator columns stay zero; the epilogue reads the first d)
      constexpr uint32_t IDESC_A1 = make_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t IDESC_A2 = make_idesc_bf16(128, DV, 0, 1);           // acc2: B = S2 read MN-major
      const uint64_t dR1 = make_desc_sw128(smem_u32(sR1));
      auto issue_score = [&](int i) {
        const uint32_t st = smem_u32(sST + (i % FB_NS) * L::STAGE);
        const uint64_t dS1 = make_desc_sw128(st);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          mma_bf16_ss(tmem_base + L::S_COL, dR1 + (uint64_t)(ks * 2), dS1 + (uint64_t)(ks * 2), IDESC_SC, ks > 0);
#pragma unroll
        for (int kk = 0; kk < DV / 16; ++kk) {
          const uint64_t a = make_desc_sw128(smem_u32(sR2) + (kk >> 2) * (128 * 128)) + (uint64_t)((kk & 3) * 2);
          const uint64_t bb = make_desc_sw128(st + L::S1_BYTES + (kk >> 2) * (64 * 128)) + (uint64_t)((kk & 3) * 2);
          mma_bf16_ss(tmem_base + L::DP_COL, a, bb, IDESC_SC, kk > 0);
        }
        mma_commit(barScore);
      };
      mbar_wait(barR, 0);
      mbar_wait(barFull, 0);
      tc_fence_after();
      issue_score(0);
      for (int i = 0; i < nst; ++i) {
        if (i + 1 < nst) {
          mbar_wait(barFull + ((i + 1) % FB_NS), ((i + 1) / FB_NS) & 1);
          mbar_wait(barSfree, i & 1);              // every elementwise thread holds S_i / dP_i in registers
          tc_fence_after();
          issue_score(i + 1);                      // runs under the exponentials of tile i
        }
        mbar_wait(barPD, i & 1);                   // P_i / dS_i are in TMEM
        tc_fence_after();
        const uint32_t st = smem_u32(sST + (i % FB_NS) * L::STAGE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {           // 16 streamed rows per K step
          if (DKV)
            mma_bf16_ts_g<1>(tmem_base + L::ACC2_COL, tmem_base + L::P_COL + ks * 8,
                             make_desc_sw128_mn(st + L::S1_BYTES + ks * 2048, 64 * 128, 1024), IDESC_A2, (i > 0) || (ks > 0));
          mma_bf16_ts_g<1>(tmem_base + L::ACC1_COL, tmem_base + L::DS_COL + ks * 8,
                           make_desc_sw128_mn(st + ks * 2048, 64 * 128, 1024), IDESC_A1, (i > 0) || (ks > 0));
        }
        mma_commit(barFree + (i % FB_NS));
        mma_commit(barAcc);
      }
    }
  } else {
    // ================================================================ elementwise warps
    const int qd = warp & 3, h = warp >> 2;
    const int lrow = qd * 32 + (threadIdx.x & 31);                  // row inside the resident tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    const long long grow = row0 + rt * 128 + lrow;
    float lse_r = 0.f, d_r = 0.f;
    if (!DKV) { lse_r = lse2[grow]; d_r = Dd[grow]; }

    for (int i = 0; i < nst; ++i) {
      const int s = i % FB_NS;
      mbar_wait(barScore, i & 1);
      tc_fence_after();
      uint32_t rs[32], rp[32];
      tmem_ld32(t_row + L::S_COL + h * 32, rs);
      tmem_ld32(t_row + L::DP_COL + h * 32, rp);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (elect_one_sync()) mbar_arrive(barSfree);
      uint32_t pk[16], dk[16];
      if (DKV) {
        mbar_wait(barFull + s, (i / FB_NS) & 1);                    // makes the producer's lse2 / D writes visible here
        const float* vec = reinterpret_cast<const float*>(sST + s * L::STAGE + L::S1_BYTES + L::S2_BYTES);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float2 ls = *reinterpret_cast<const float2*>(vec + h * 32 + 2 * e);
          const float2 dd = *reinterpret_cast<const float2*>(vec + 64 + h * 32 + 2 * e);
          const float p0 = ex2_approx(__uint_as_float(rs[2 * e]) - ls.x);
          const float p1 = ex2_approx(__uint_as_float(rs[2 * e + 1]) - ls.y);
          pk[e] = pack_bf16x2(p0, p1);
          dk[e] = pack_bf16x2(p0 * (__uint_as_float(rp[2 * e]) - dd.x), p1 * (__uint_as_float(rp[2 * e + 1]) - dd.y));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(__uint_as_float(rs[2 * e]) - lse_r);
          const float p1 = ex2_approx(__uint_as_float(rs[2 * e + 1]) - lse_r);
          pk[e] = pack_bf16x2(p0, p1);
          dk[e] = pack_bf16x2(p0 * (__uint_as_float(rp[2 * e]) - d_r), p1 * (__uint_as_float(rp[2 * e + 1]) - d_r));
        }
      }
      if (i >= 1) {                                                 // the MMAs of tile i-1 have read the P / dS columns
        mbar_wait(barAcc, (i - 1) & 1);
        tc_fence_after();
      }
      if (DKV) tmem_st16(t_row + L::P_COL + h * 16, pk);
      tmem_st16(t_row + L::DS_COL + h * 16, dk);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (elect_one_sync()) mbar_arrive(barPD);
    }

    // ---- epilogue: accumulators -> global memory (every row of the outputs is written by exactly one CTA)
    mbar_wait(barAcc, (nst - 1) & 1);
    tc_fence_after();
    const float gm = *gamma;
    const float sc1 = DKV ? gm * FB_LN2 : gm;                       // Q rows carry log2 e: dK = ln 2 * dS^T Q'
    if (h == 0) {
#pragma unroll
      for (int c = 0; c < D; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + L::ACC1_COL + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          st4(out1 + grow * ld + c + e, make_float4(__uint_as_float(r[e]) * sc1, __uint_as_float(r[e + 1]) * sc1,
                                                   __uint_as_float(r[e + 2]) * sc1, __uint_as_float(r[e + 3]) * sc1));
      }
    }
    if (DKV) {
#pragma unroll 1
      for (int c = h * (DV / 2); c < (h + 1) * (DV / 2); c += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + L::ACC2_COL + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          st4(out2 + grow * ld + c + e, make_float4(__uint_as_float(r[e]) * gm, __uint_as_float(r[e + 1]) * gm,
                                                    __uint_as_float(r[e + 2]) * gm, __uint_as_float(r[e + 3]) * gm));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D'[t] = sum_v bf16(dA')[t, v] * A[t, v] (the SAME rounded dA' the dP MMA sees, so sum_j P_ij (dP_ij - D_i) = 0 up
// to the rounding of P), lse2[t] = lse[t] log2 e.  One warp per token.
__global__ void __launch_bounds__(256)
attn_big_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dAb, const float* __restrict__ A, const float* __restrict__ lse,
                         float* __restrict__ Dd, float* __restrict__ lse2, long long T, int dv) {
  const long long t = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int v = lane * 2; v < dv; v += 64) {
    const __nv_bfloat162 g = *reinterpret_cast<const __nv_bfloat162*>(dAb + t * dv + v);
    const float2 a = *reinterpret_cast<const float2*>(A + t * dv + v);
    acc = fmaf(__bfloat162float(g.x), a.x, acc);
    acc = fmaf(__bfloat162float(g.y), a.y, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    Dd[t] = acc;
    lse2[t] = lse[t] * 1.4426950408889634f;
  }
}

template <int DV, int D, bool DKV>
static int launch_fb(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& s1, const CUtensorMap& s2,
                     const float* lse2, const float* Dd, const float* gamma, float* out1, float* out2, int ld, int B, int N,
                     cudaStream_t st) {
  using L = FbSmem<DV>;
  auto kern = attn_bwd_big_kernel<DV, D, DKV>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<dim3(N / 128, B), FB_THREADS, L::TOTAL, st>>>(r1, r2, s1, s2, lse2, Dd, gamma, out1, out2, ld, N);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// dQKV [T, 2 d + dv] = [dQ | dK | dV] (fp32, gamma folded in) from the bf16 operands Qb / Kb [T, 64], Vb / dAb [T, dv]
// (dAb = dY Wo^T WITHOUT gamma), the forward's row log-sum-exp and the saved attention output A.
int attn_big_fused_bwd_core(const __nv_bfloat16* Qb, const __nv_bfloat16* Kb, const __nv_bfloat16* Vb,
                            const __nv_bfloat16* dAb, const float* A, const float* lse, const float* gamma, float* lse2,
                            float* Dd, float* dQKV, int B, int N, int C, cudaStream_t st) {
  const int d = C / 8, dv = C / 2, ld = 2 * d + dv;
  float *dQ = dQKV, *dK = dQKV + d, *dV = dQKV + 2 * d;
  const long long T = (long long)B * N;
  attn_big_bwd_prep_kernel<<<(unsigned)ceil_div<long long>(T, 8), 256, 0, st>>>(dAb, A, lse, Dd, lse2, T, dv);
  SAGAN_LAUNCH_CHECK();
  CUtensorMap q128, q64, k128, k64, v128, v64, a128, a64;
  int rc;
  if ((rc = make_tmap_bf16_2d(&q128, Qb, (uint64_t)T, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&q64, Qb, (uint64_t)T, 64, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&k128, Kb, (uint64_t)T, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&k64, Kb, (uint64_t)T, 64, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&v128, Vb, (uint64_t)T, (uint64_t)dv, (uint64_t)dv * 2, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&v64, Vb, (uint64_t)T, (uint64_t)dv, (uint64_t)dv * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&a128, dAb, (uint64_t)T, (uint64_t)dv, (uint64_t)dv * 2, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&a64, dAb, (uint64_t)T, (uint64_t)dv, (uint64_t)dv * 2, 64, 64))) return rc;
#define SAGAN_FB_CASE(DVV, DD)                                                                                            \
  case DVV:                                                                                                              \
    if ((rc = launch_fb<DVV, DD, true>(k128, v128, q64, a64, lse2, Dd, gamma, dK, dV, ld, B, N, st))) return rc;          \
    return launch_fb<DVV, DD, false>(q128, a128, k64, v64, lse2, Dd, gamma, dQ, nullptr, ld, B, N, st);
  switch (dv) {
    SAGAN_FB_CASE(64, 16)
    SAGAN_FB_CASE(128, 32)
    SAGAN_FB_CASE(256, 64)
  }
#undef SAGAN_FB_CASE
  set_err("attn_big_fused_bwd_core: dv = %d not built", dv);
  return SAGAN_EUNSUPPORTED;
}

}  // namespace sagan
