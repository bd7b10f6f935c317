// Self-attention block, BF16_TC math mode, LARGE channel counts (C = 128 / 256 / 512: d = 16 / 32 / 64 logit
// dims, dv = 64 / 128 / 256 value dims) -- the regime of the stand-alone kernel sweep (BASELINE.json configs[4]),
// where the two contractions S = theta phi^T and A = softmax(S) g are real tensor-core work.
//
// Replaces Attention_Layer.call (/root/reference/layers.py:93-120) as four launches:
//   1. gemm_tf32_qkv     X [T,C] x [Wtheta | Wphi | Wg] -> bf16 Q (pre-scaled by log2 e), K (rows of 64), V [T,dv]
//                        (TMA-fed tcgen05 kind::tf32 GEMM over CTA pairs, gemm_tc.cu, projection epilogue)
//   2. attn_fwd_big      flash forward, this file
//   3. gemm_bf16_residual Y = X + gamma (A Wo + bo)  (same GEMM core on the bf16 copy of A, residual epilogue)
//
// attn_fwd_big: CTA = one 128-query tile of one sample, 6 warps:
//   warps 0-3  softmax + epilogue: thread r <-> TMEM lane r <-> query row r; the whole S row (128 fp32) is pulled
//              into registers with one exposed tcgen05.ld round trip, P = exp2(S - m) goes back to shared memory as
//              bf16 in the UMMA K-major SW128 layout
//   warp 4     TMA producer: K tile [128 keys][64] and V tile as dv/64 slabs of [128 keys][64 values] (V is used in its
//              natural [token][dv] layout, as the MN-major B operand of the PV MMA -- no transposed copy)
//   warp 5     MMA issuer: S_{j+1} = Q K_{j+1}^T is issued BEFORE it waits for P_j, so the tensor core computes the
//              next logits under the softmax of the current tile (two S buffers in TMEM); O += P_j V_j
// TMEM: S0 [0,128) S1 [128,256) O [256, 256+dv) -> all 512 columns at dv = 256, one CTA per SM.
// The O rescale is lazy (row max grows by > 2^32), so O stays in TMEM across key tiles.
//
// Measured per 128-key tile at dv = 256 (clock64 timeline, round 1): softmax thread ~3200 cycles (wait S 990, exp 1430,
// P store 430), PV issue + completion ~1900 against a 1024-cycle MMA floor.  The SS-mode PV MMA reads P (4 KB) and V (8 KB)
// from shared memory for every K step (96 B/clk) while TMA writes the next 80 KB tile and the softmax warps store P: the
// kernel is shared-memory-bandwidth bound.  Next steps: P as the TMEM A operand (no P store / read), a 3-deep K ring
// (S ready earlier), cluster multicast of K / V.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace sagan {

using namespace tc;

// gemm_tc.cu
int gemm_tf32_qkv(const float* x, const float* wt, const float* bcat, __nv_bfloat16* q, __nv_bfloat16* k,
                  __nv_bfloat16* v, long long M, int K, int d, int dv, float q_scale, cudaStream_t st);
int gemm_bf16_residual(const __nv_bfloat16* a, const __nv_bfloat16* wt, const float* bias, const float* res,
                       const float* res_scale, float* y, long long M, int K, int N, cudaStream_t st);

constexpr float BG_LOG2E = 1.4426950408889634f;
constexpr float BG_LN2 = 0.6931471805599453f;
#ifdef SAGAN_TIMELINE
__device__ unsigned long long g_big_tl[32];
#define TL_DECL unsigned long long tl_t0 = clock64(), tl_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TL_MARK(i) do { const unsigned long long t_ = clock64(); tl_acc[i] += t_ - tl_t0; tl_t0 = t_; } while (0)
#define TL_DUMP(base, n) do { if (blockIdx.x == 0 && blockIdx.y == 0) for (int i_ = 0; i_ < (n); ++i_) g_big_tl[(base) + i_] = tl_acc[i_]; } while (0)
#else
#define TL_DECL
#define TL_MARK(i)
#define TL_DUMP(base, n)
#endif
// 2^x on the FMA pipe: round-to-nearest split x = n + f (magic-number add), cubic minimax of 2^f on [-0.5, 0.5]
// (max relative error 7.5e-5, far below the bf16 rounding of P), n added into the exponent field
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;
  const float f = x - (r - 12582912.0f);
  const float p = fmaf(fmaf(fmaf(5.517132208e-02f, f, 2.426105440e-01f), f, 6.932609677e-01f), f, 9.999281168e-01f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

constexpr int BG_THREADS = 320;   // 8 softmax warps + TMA producer warp + MMA issuer warp

// NCTA = 2: a CTA pair (cluster of two adjacent 128-query tiles) runs every MMA with cta_group::2 (M = 256); each CTA
// loads only HALF of every K tile (64 keys) and of every V tile (dv/2 value columns), which halves the L2 -> SM traffic
// per FLOP and the shared-memory footprint per stage (4 stages instead of 2).
template <int DV, int NCTA>
struct BigSmem {
  static constexpr int Q_BYTES = 128 * 128;
  static constexpr int K_BYTES = (128 / NCTA) * 128;               // this CTA's keys of the tile, rows of 64 bf16
  static constexpr int V_SLABS = DV / 64 / NCTA;                   // this CTA's slabs of [128 keys][64 values = 128 B]
  static constexpr int V_BYTES = V_SLABS * 128 * 128;
  static constexpr int NS = 2 * NCTA;                              // K / V stages
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + NS * K_BYTES;
  static constexpr int OFF_X = OFF_V + NS * V_BYTES;               // row-max / row-sum exchange between the two halves
  static constexpr int OFF_BAR = OFF_X + 2 * 2 * 128 * 4;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static constexpr int OCOL = 256;
  static constexpr uint32_t STAGE_TX = NCTA * (K_BYTES + V_BYTES); // bytes landing per stage over the whole pair
  static_assert(V_SLABS >= 1, "a CTA pair needs dv >= 128");
};

// CTA (or CTA pair) = one 128-query (256-query) tile of one sample, 10 warps per CTA:
//   warps 0-3 / 4-7  softmax halves: thread <-> TMEM lane <-> query row, key columns [0,64) / [64,128) of the tile.  The
//              half row is pulled into registers with tcgen05.ld, the row max is exchanged with the other half through
//              shared memory, and P = bf16(exp2(S - m)) goes BACK INTO TMEM over the logits it came from (tcgen05.st,
//              two keys per 32-bit column): the PV MMA takes P as its TMEM A operand, so P never touches shared memory
//   warp 8     TMA producer (K tile + V slabs per stage; V in its natural [token][dv] layout = MN-major B operand)
//   warp 9     MMA issuer (leader CTA only): S_{j+1} = Q K_{j+1}^T is issued BEFORE the wait for P_j (two S buffers),
//              then O += P_j V_j.  The tensor pipe executes in issue order, which is what makes the aliasing safe:
//              S_{j+2} overwrites buffer j & 1 only after PV_j has read P_j from it.
// TMEM: S0/P0 [0,128)  S1/P1 [128,256)  O [256, 256+dv).  The O rescale is lazy (row max grows by > 2^32).
template <int DV, int NCTA>
__global__ void __launch_bounds__(BG_THREADS, 1)
attn_fwd_big_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, float* __restrict__ lse, float* __restrict__ A_saved,
                    __nv_bfloat16* __restrict__ A_bf16, int N, int kq_steps) {
  using L = BigSmem<DV, NCTA>;
  extern __shared__ uint8_t smem_raw[];
  // dynamic shared memory starts at the same window offset in both CTAs of a pair, so the aligned layout matches too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + L::OFF_Q;
  uint8_t* sK = smem + L::OFF_K;
  uint8_t* sV = smem + L::OFF_V;
  float* sX = reinterpret_cast<float*>(smem + L::OFF_X);   // [2 buffers][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* barQ = bars + 0;             // leader: Q tiles of the pair landed
  uint64_t* barFull = bars + 1;          // [NS] leader: K / V stage landed in both CTAs
  uint64_t* barEmpty = barFull + L::NS;  // [NS] every CTA: PV of the stage's tile done -> stage free
  uint64_t* barS = barEmpty + L::NS;     // [2]  every CTA: S buffer ready
  uint64_t* barP = barS + 2;             // [2]  leader: 8 * NCTA warp arrivals: P_j is in TMEM (and S_j has been read)
  uint64_t* barO = barP + 2;             //      every CTA: one phase per key tile: PV_j done, O up to date
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(barO + 1);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = NCTA == 1 ? 0u : cluster_ctarank();
  const bool leader = rank == 0;
  const int b = blockIdx.y, qt = blockIdx.x;
  const int nt = N / 128;

  if (threadIdx.x == 0) {
    mbar_init(barQ, 1);
    for (int i = 0; i < L::NS; ++i) { mbar_init(barFull + i, 1); mbar_init(barEmpty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(barS + i, 1); mbar_init(barP + i, 8 * NCTA); }
    mbar_init(barO, 1);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc_g<NCTA>(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ================================================================ TMA producer (both CTAs: own halves)
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      if (leader) mbar_expect_tx(barQ, NCTA * L::Q_BYTES);
      tma_load_2d_g<NCTA>(sQ, &tmQ, barQ, 0, b * N + qt * 128);
      for (int j = 0; j < nt; ++j) {
        const int s = j % L::NS;
        if (j >= L::NS) mbar_wait(barEmpty + s, ((j / L::NS) - 1) & 1);
        if (leader) mbar_expect_tx(barFull + s, L::STAGE_TX);
        tma_load_2d_g<NCTA>(sK + s * L::K_BYTES, &tmK, barFull + s, 0, b * N + j * 128 + (int)rank * (128 / NCTA));
#pragma unroll
        for (int sl = 0; sl < L::V_SLABS; ++sl)
          tma_load_2d_g<NCTA>(sV + s * L::V_BYTES + sl * (128 * 128), &tmV, barFull + s,
                              ((int)rank * L::V_SLABS + sl) * 64, b * N + j * 128);
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issuer (leader CTA)
    if (leader && elect_one_sync()) {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128 * NCTA, 128);
      constexpr uint32_t IDESC_O = make_idesc_bf16(128 * NCTA, DV, 0, /*b_mn_major=*/1);
      const uint64_t descQ = make_desc_sw128(smem_u32(sQ));
      TL_DECL;
      auto issue_qk = [&](int j) {
        const int s = j % L::NS;
        TL_MARK(0);
        mbar_wait(barFull + s, (j / L::NS) & 1);
        TL_MARK(1);
        tc_fence_after();
        const uint64_t descK = make_desc_sw128(smem_u32(sK + s * L::K_BYTES));
        for (int ks = 0; ks < kq_steps; ++ks)
          mma_bf16_ss_g<NCTA>(tmem_base + (uint32_t)((j & 1) * 128), descQ + (uint64_t)(ks * 2), descK + (uint64_t)(ks * 2),
                              IDESC_S, ks > 0);
        mma_commit_g<NCTA>(barS + (j & 1));
      };
      mbar_wait(barQ, 0);
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_qk(j + 1);
        TL_MARK(0);
        mbar_wait(barP + (j & 1), (j >> 1) & 1);
        TL_MARK(2);
        tc_fence_after();
        const int s = j % L::NS;
        const uint32_t pbase = tmem_base + (uint32_t)((j & 1) * 128);
        const uint32_t vbase = smem_u32(sV + s * L::V_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 keys per step; P of keys [64 h, 64 h + 64) sits at columns 64 h + [0, 32)
          const uint32_t a = pbase + (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8);
          const uint64_t bb = make_desc_sw128_mn(vbase + ks * 2048, 128 * 128, 1024);
          mma_bf16_ts_g<NCTA>(tmem_base + L::OCOL, a, bb, IDESC_O, (j > 0) || (ks > 0));
        }
        mma_commit_g<NCTA>(barEmpty + s);
        mma_commit_g<NCTA>(barO);
#ifdef SAGAN_TIMELINE_PV
        TL_MARK(0);
        mbar_wait(barO, j & 1);
        TL_MARK(3);
#endif
      }
      TL_MARK(0);
      TL_DUMP(16, 4);
    }
  } else {
    // ================================================================ softmax warps
    const int h = warp >> 2;                      // key-column half
    const int row = threadIdx.x & 127;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float m_used = -INFINITY, l = 0.f;
    TL_DECL;

    for (int j = 0; j < nt; ++j) {
      TL_MARK(0);
      mbar_wait(barS + (j & 1), (j >> 1) & 1);
      TL_MARK(1);
      tc_fence_after();
      const uint32_t t_s = t_row + (uint32_t)((j & 1) * 128 + h * 64);
      uint32_t r[64];
      tmem_ld32(t_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_wait_ld();
      TL_MARK(2);
      // ---- P = exp2(S - m) with the shift m of the PREVIOUS tiles (logits are in log2 units: Q carries log2 e).  The
      // shift only has to keep P inside the fp32 / bf16 exponent range, so the exponentials do not wait for this
      // tile's row max; the max (computed alongside, exchanged with the other half of the row) is checked afterwards
      // and only when it exceeds the shift by more than 2^32 -- in practice in the first tiles only -- are O and l
      // rescaled (which has to wait for the tensor core and touch all of O) and this tile's P recomputed.
      uint32_t pk[32];
      float l0, l1;
      auto exps = [&](float m) {
        l0 = 0.f; l1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          // one exponential in five runs as a polynomial on the FMA pipe: the MUFU unit (16 ex2 / clk / SM) is the
          // busiest pipe of this kernel
          const float p0 = (i % 5 == 2) ? ex2_poly(__uint_as_float(r[2 * i]) - m) : ex2_approx(__uint_as_float(r[2 * i]) - m);
          const float p1 = (i % 5 == 4) ? ex2_poly(__uint_as_float(r[2 * i + 1]) - m) : ex2_approx(__uint_as_float(r[2 * i + 1]) - m);
          l0 += p0;
          l1 += p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
      };
      if (j > 0) exps(m_used);
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i]));
        mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      TL_MARK(5);
      // both halves of a row must use the same shift: exchange the partial maxima (double-buffered, one barrier a tile)
      float* xb = sX + (j & 1) * 256;
      xb[h * 128 + row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, xb[(h ^ 1) * 128 + row]);
      TL_MARK(3);
      const bool need = mx > m_used + 32.0f;
      if (__any_sync(0xffffffffu, need)) {
        if (j > 0) {
          mbar_wait(barO, (j - 1) & 1);                    // every PV issued so far has completed
          tc_fence_after();
          const float scale = need ? exp2f(m_used - mx) : 1.0f;
          l *= scale;
#pragma unroll
          for (int c = 0; c < DV / 64; ++c) {                // this half's share of the O columns
            uint32_t o[32];
            tmem_ld32(t_row + L::OCOL + h * (DV / 2) + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
            tmem_st32(t_row + L::OCOL + h * (DV / 2) + c * 32, o);
          }
        }
        if (need) m_used = mx;
        if (need || j == 0) exps(m_used);                  // warp-uniform in the first tile; rare afterwards
      }
      l += l0 + l1;
      TL_MARK(4);
      tmem_st32(t_s, pk);
      tmem_wait_st();
      TL_MARK(7);
      tc_fence_before();             // orders this thread's tcgen05.ld / st before the arrive
      __syncwarp();
      if (elect_one_sync()) mbar_arrive_leader(barP + (j & 1));
      TL_MARK(6);
    }
    if (threadIdx.x == 0) TL_DUMP(0, 8);
    if (threadIdx.x == 128) TL_DUMP(8, 8);

    // ---- epilogue: A = O / l (fp32, saved for the backward and read by the output-conv GEMM), lse
    // barO completes one phase per key tile and a parity wait can only tell the current phase from the one before:
    // having read S_{nt-1}, this thread knows that PV_{nt-3} is complete (the pipe runs in issue order), no more --
    // so it waits for PV_{nt-2} first; asking for the last phase directly could alias with phase nt-3 and return
    // while two PV MMAs are still in flight (seen as a wrong A on a cold first launch).
    if (nt >= 2) mbar_wait(barO, (nt - 2) & 1);
    mbar_wait(barO, (nt - 1) & 1);
    tc_fence_after();
    float* xb = sX + (nt & 1) * 256;
    xb[h * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xb[(h ^ 1) * 128 + row];
    const long long grow = (long long)b * N + qt * 128 + row;
    const float inv = 1.0f / l;
    // TMEM hands each thread one ROW of O (32 columns at a time).  Every MMA of the pair has completed, so the K / V
    // ring is dead: a per-warp transpose through it lets the warp store 4 rows x 128 contiguous bytes per instruction
    // instead of 32 rows x 16 bytes.
    const int lane = threadIdx.x & 31;
    float* sT = reinterpret_cast<float*>(sK) + warp * (32 * 36);       // [32 rows][36]: conflict-free 16-byte accesses
    const long long wrow0 = (long long)b * N + qt * 128 + (warp & 3) * 32;
#pragma unroll 1
    for (int c = 0; c < DV / 64; ++c) {
      uint32_t o[32];
      tmem_ld32(t_row + L::OCOL + h * (DV / 2) + c * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(sT + lane * 36 + q * 4) =
            make_float4(__uint_as_float(o[4 * q]) * inv, __uint_as_float(o[4 * q + 1]) * inv,
                        __uint_as_float(o[4 * q + 2]) * inv, __uint_as_float(o[4 * q + 3]) * inv);
      __syncwarp();
      const int col = h * (DV / 2) + c * 32 + (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {          // 4 rows x (8 lanes x 4 columns) per pass
        const int rr = it * 4 + (lane >> 3);
        const float4 a = *reinterpret_cast<const float4*>(sT + rr * 36 + (lane & 7) * 4);
        st4(A_saved + (wrow0 + rr) * DV + col, a);
        // bf16 copy: the A operand of the output-conv GEMM
        *reinterpret_cast<uint2*>(A_bf16 + (wrow0 + rr) * DV + col) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
      }
      __syncwarp();
    }
    if (h == 0) lse[grow] = (m_used + log2f(l)) * BG_LN2;
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();     // no CTA of the pair retires while the other may still signal it
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_g<NCTA>(tmem_base, 512);
  }
}

// K-major (transposed) weight operands of the two GEMMs: Wt [2d+dv][C] = [Wq | Wk | Wv]^T (fp32), bcat, WoT [C][dv] = Wo^T (bf16)
__global__ void attn_big_weights_kernel(const float* __restrict__ Wq, const float* __restrict__ bq,
                                        const float* __restrict__ Wk, const float* __restrict__ bk,
                                        const float* __restrict__ Wv, const float* __restrict__ bv,
                                        const float* __restrict__ Wo, float* __restrict__ Wt, float* __restrict__ bcat,
                                        __nv_bfloat16* __restrict__ WoT, int C, int d, int dv) {
  const int NN = 2 * d + dv;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * NN) {
    const int n = i / C, c = i - n * C;
    Wt[i] = n < d ? Wq[c * d + n] : (n < 2 * d ? Wk[c * d + n - d] : Wv[c * dv + n - 2 * d]);
  }
  if (i < C * dv) {
    const int c = i / dv, v = i - c * dv;
    WoT[i] = __float2bfloat16_rn(Wo[v * C + c]);
  }
  if (i < NN) bcat[i] = i < d ? bq[i] : (i < 2 * d ? bk[i - d] : bv[i - 2 * d]);
}

int attn_big_weights_launch(const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                            const float* bv, const float* Wo, float* Wt, float* bcat, __nv_bfloat16* WoT, int C, cudaStream_t st) {
  const int d = C / 8, dv = C / 2, nn = C * (2 * d + dv);
  attn_big_weights_kernel<<<ceil_div(nn, 256), 256, 0, st>>>(Wq, bq, Wk, bk, Wv, bv, Wo, Wt, bcat, WoT, C, d, dv);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

struct BigLayout {
  size_t off_w, off_b, off_wo, off_q, off_k, off_v, off_ab, total;
};

static BigLayout big_layout(int B, int N, int C) {
  const int d = C / 8, dv = C / 2;
  const size_t T = (size_t)B * N;
  BigLayout t;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 1023) / 1024 * 1024; return r; };
  t.off_w = take((size_t)C * (2 * d + dv) * 4);
  t.off_b = take((size_t)(2 * d + dv) * 4);
  t.off_wo = take((size_t)C * dv * 2);
  t.off_q = take(T * 64 * 2);
  t.off_k = take(T * 64 * 2);
  t.off_v = take(T * dv * 2);
  t.off_ab = take(T * dv * 2);
  t.total = o + 1024;
  return t;
}

#ifdef SAGAN_TIMELINE
extern "C" int sagan_debug_big_timeline(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_big_tl, sizeof(g_big_tl));
}
#endif

bool attn_tc_big_supported(int N, int C) { return (C == 128 || C == 256 || C == 512) && N % 128 == 0; }
size_t attn_tc_big_workspace_bytes(int B, int N, int C) { return big_layout(B, N, C).total; }

template <int DV, int NCTA>
static int launch_big(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, float* lse, float* A,
                      __nv_bfloat16* Ab, int B, int N, int kq_steps, cudaStream_t st) {
  using L = BigSmem<DV, NCTA>;
  auto kern = attn_fwd_big_kernel<DV, NCTA>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(N / 128, B);
  cfg.blockDim = dim3(BG_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SAGAN_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, lse, A, Ab, N, kq_steps));
  SAGAN_LAUNCH_CHECK();
  return 0;
}

int attn_tc_big_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                    const float* bv, const float* Wo, const float* bo, const float* gamma, float* Y, float* lse, float* A,
                    int B, int N, int C, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!attn_tc_big_supported(N, C)) {
    set_err("sagan_attn_fwd: large-C tensor-core path needs C in {128,256,512} and N %% 128 == 0 (C=%d, N=%d)", C, N);
    return SAGAN_EUNSUPPORTED;
  }
  const BigLayout t = big_layout(B, N, C);
  if (ws_bytes < t.total) {
    set_err("sagan_attn_fwd: workspace %zu < %zu bytes", ws_bytes, t.total);
    return SAGAN_EWORKSPACE;
  }
  const int d = C / 8, dv = C / 2;
  const long long T = (long long)B * N;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  float* Wt = reinterpret_cast<float*>(base + t.off_w);
  float* bcat = reinterpret_cast<float*>(base + t.off_b);
  __nv_bfloat16* WoT = reinterpret_cast<__nv_bfloat16*>(base + t.off_wo);
  __nv_bfloat16* Ab = reinterpret_cast<__nv_bfloat16*>(base + t.off_ab);
  __nv_bfloat16* Qb = reinterpret_cast<__nv_bfloat16*>(base + t.off_q);
  __nv_bfloat16* Kb = reinterpret_cast<__nv_bfloat16*>(base + t.off_k);
  __nv_bfloat16* Vb = reinterpret_cast<__nv_bfloat16*>(base + t.off_v);
  const int nn = C * (2 * d + dv);
  attn_big_weights_kernel<<<ceil_div(nn, 256), 256, 0, st>>>(Wq, bq, Wk, bk, Wv, bv, Wo, Wt, bcat, WoT, C, d, dv);
  SAGAN_LAUNCH_CHECK();
  if (d < 64) {   // rows of Q / K are padded to one 128-byte swizzle span
    SAGAN_CUDA(cudaMemsetAsync(Qb, 0, (size_t)T * 64 * 2, st));
    SAGAN_CUDA(cudaMemsetAsync(Kb, 0, (size_t)T * 64 * 2, st));
  }
  int rc = gemm_tf32_qkv(X, Wt, bcat, Qb, Kb, Vb, T, C, d, dv, BG_LOG2E, st);
  if (rc) return rc;
  CUtensorMap tq, tk, tv;
  if ((rc = make_tmap_bf16_2d(&tq, Qb, (uint64_t)T, 64, 128, 128))) return rc;
  // CTA pairs (cta_group::2) whenever a pair has two whole query tiles and each CTA at least one 64-value slab of V
  static const bool force_single = getenv("SAGAN_ATTN_BIG_SINGLE_CTA") != nullptr;   // diagnostics only
  const bool pair = dv >= 128 && (N / 128) % 2 == 0 && !force_single;
  if ((rc = make_tmap_bf16_2d(&tk, Kb, (uint64_t)T, 64, 128, pair ? 64 : 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, Vb, (uint64_t)T, (uint64_t)dv, (uint64_t)dv * 2, 128, 64))) return rc;
  const int kq = d / 16;
  switch (dv) {
    case 64: rc = launch_big<64, 1>(tq, tk, tv, lse, A, Ab, B, N, kq, st); break;
    case 128: rc = pair ? launch_big<128, 2>(tq, tk, tv, lse, A, Ab, B, N, kq, st) : launch_big<128, 1>(tq, tk, tv, lse, A, Ab, B, N, kq, st); break;
    default: rc = pair ? launch_big<256, 2>(tq, tk, tv, lse, A, Ab, B, N, kq, st) : launch_big<256, 1>(tq, tk, tv, lse, A, Ab, B, N, kq, st); break;
  }
  if (rc) return rc;
  return gemm_bf16_residual(Ab, WoT, bo, X, gamma, Y, T, dv, C, st);
}

}  // namespace sagan
