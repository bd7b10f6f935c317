// Implicit-GEMM convolution on the 5th-gen tensor cores (BF16_TC math mode; FWD / DGRAD run kind::tf32 on the fp32
// operands, WGRAD kind::f16 on bf16-converted operands -- see the TF32 note at the kernel): forward, backward-data
// (= Conv2DTranspose forward) and backward-filter of the spectrally-normalised Conv2D / Conv2DTranspose /
// Dense layers (/root/reference/sagan/models/generator.py:8-9,25,36, discriminator.py:8,35).
//
// One kernel template, three gather modes.  fp32 NHWC activations and fp32 Keras-layout kernels are read
// directly: the producer warps gather 8-element chunks, convert to bf16 and store them into the 128-byte-
// swizzled UMMA operand layouts (software im2col -- no intermediate tensors in HBM); accumulation is fp32 in TMEM.
//
//   FWD    D[m][n] = sum_k A[m][k] W[k][n]      A = im2col(x) (K-major tile), W = HWIO kernel read as an MN-major B tile
//   DGRAD  per stride-parity class: D[m][n] = sum_k dY[m][k] Wc[n][k]   both K-major (W[kh,kw,n,:] rows are contiguous in co)
//   WGRAD  D[kk][n] = sum_m A[m][kk] dY[m][n]   both operands MN-major (the same im2col rows, dY rows); split over
//          pixel ranges with fp32 atomics; an extra im2col column of ones yields the bias gradient for free
//
// Operand precision (template PREC): 0 = plain bf16, 1 = tf32 on the fp32 operands, 2 = SPLIT bf16: every operand is
// carried as hi + lo bf16 tiles (x = hi + lo to 16 mantissa bits) and each K step issues the three MMAs
// hi*hi + hi*lo + lo*hi -- fp32-grade products (dropped term 2^-16) at the shared-memory footprint of the fp32 / tf32
// tiles and 0.75x the tensor time of tf32.  PREC 2 is what BF16_TC mode uses (sagan_conv_tc_set_precision): with it
// the conv / deconv / dense layers no longer move LeakyReLU pre-activations across zero, which is what kept the
// model-level gradients of the tensor-core mode out of the 2e-3 tier.
//
// CTA tile 128 x NT (NT in {16,32,64,128}), K blocks of 64, 2- or 3-stage shared-memory ring.
// 9 warps: 0-7 producers + epilogue, warp 8 = MMA issuer / TMEM owner.
#include "common.cuh"
#include "tc_common.cuh"

namespace sagan {

using namespace tc;

enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };
constexpr int CT_THREADS = 288;
enum { PREC_BF16 = 0, PREC_TF32 = 1, PREC_SPLIT = 2 };

struct ConvTcP {
  const float* a_src;   // FWD: x     DGRAD: dy    WGRAD: x
  const float* b_src;   // FWD: w     DGRAD: w     WGRAD: dy
  const float* bias;    // FWD (may be null)
  float* out;           // FWD: y     DGRAD: dx    WGRAD: dw (zeroed by the caller)
  float* dbias;         // WGRAD (may be null; zeroed by the caller)
  CG g;
  int act;
  float slope;
  int m_per_split;      // WGRAD: pixels per blockIdx.z
  int k_splits;         // FWD / DGRAD: split-K factor (blockIdx.z = class * k_splits + split); > 1 => atomic accumulation
                        //              into a zeroed output, bias / activation applied afterwards by conv_tc_finish_kernel
  // FWD epilogue variants used by the large-C attention block (attn_tc_big.cu):
  const float* res;         // out = res + (*res_scale) * (acc + bias)   (gamma residual, layers.py:120)
  const float* res_scale;
  __nv_bfloat16* q_out;     // when set: columns [0,d) -> q_out[m][64] * q_scale, [d,2d) -> k_out[m][64], [2d,2d+dv) -> v_out[m][dv]
  __nv_bfloat16* k_out;
  __nv_bfloat16* v_out;
  int qk_d, v_dv;
  float q_scale;
};

__device__ __forceinline__ float ct_act(float v, int act, float slope) {
  if (act == SAGAN_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == SAGAN_ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ void st_chunk(uint8_t* dst, const float4& a, const float4& b) {
  *reinterpret_cast<uint4*>(dst) =
      make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

__device__ __forceinline__ void st_chunk_split(uint8_t* dst_hi, uint8_t* dst_lo, const float4& a, const float4& b) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) lo[e] = v[e] - __bfloat162float(__float2bfloat16_rn(v[e]));
  *reinterpret_cast<uint4*>(dst_hi) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(dst_lo) =
      make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]), pack_bf16x2(lo[6], lo[7]));
}

template <int MODE, int NT, int PREC>
struct ConvTcSmem {
  static constexpr bool TF32 = PREC == PREC_TF32, SPLIT = PREC == PREC_SPLIT;
  static constexpr int A_BYTES = 128 * 128;                                              // 16 KB either major
  static constexpr int B_RAW = (MODE == TC_DGRAD || TF32) ? NT * 128 : 8192 * ((NT + 63) / 64);  // K-major rows / MN-major sub-tiles
  static constexpr int B_BYTES = B_RAW < 1024 ? 1024 : B_RAW;
  static constexpr int STAGE = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);    // SPLIT: [A_hi | B_hi | A_lo | B_lo]
  static constexpr int LO_OFF = A_BYTES + B_BYTES;                       // SPLIT: offset of the lo tiles inside a stage
  static constexpr int STAGES = SPLIT ? 2 : 3;
  static constexpr int OFF_BAR = STAGES * STAGE;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;
  static constexpr int TMEM_COLS = NT < 32 ? 32 : NT;
};

// TF32 (FWD / DGRAD only): the operands stay fp32 in shared memory and the MMA is kind::tf32 (10 mantissa bits instead
// of bf16's 7, K = 8 per instruction), K blocks of 32 elements, both operands K-major (FWD gathers the HWIO kernel
// transposed).  Five bf16-operand layers in a row cost 6e-3 on G(z); tf32 keeps the model inside the 2e-3 tier at the
// same shared-memory footprint, and the producers copy instead of converting.
__host__ __device__ constexpr uint32_t make_idesc_tf32_conv(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss_conv(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

template <int MODE, int NT, int PREC>
__global__ void __launch_bounds__(CT_THREADS, (PREC == PREC_SPLIT && NT > 64) ? 1 : 2)
conv_tc_kernel(const ConvTcP p) {
  constexpr bool TF32 = PREC == PREC_TF32, SPLIT = PREC == PREC_SPLIT;
  static_assert(!(TF32 && MODE == TC_WGRAD), "the tf32 variant covers FWD and DGRAD");
  constexpr int KB = TF32 ? 32 : 64;      // K elements per block (128 bytes per row either way)
  constexpr int CE = TF32 ? 4 : 8;        // elements per 16-byte chunk
  using L = ConvTcSmem<MODE, NT, PREC>;
  constexpr int CT_STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full = bars;                 // [CT_STAGES], 256 arrivals
  uint64_t* empty = bars + CT_STAGES;    // [CT_STAGES], tcgen05.commit
  uint64_t* accum = bars + 2 * CT_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * CT_STAGES + 1);
  const CG& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  // ---- GEMM view of this CTA
  int Mg, Ng, k_begin, k_end;            // rows, cols, reduction range
  int rh = 0, rw = 0, hi_first = 0, wi_first = 0, Hc = 1, Wc = 1, ntw = 1;   // DGRAD class geometry
  if (MODE == TC_FWD) {
    Mg = g.M; Ng = g.Cout;
    const int kper = (((g.K + p.k_splits - 1) / p.k_splits) + 63) / 64 * 64;
    k_begin = (int)blockIdx.z * kper; k_end = min(g.K, k_begin + kper);
  } else if (MODE == TC_DGRAD) {
    const int cls = blockIdx.z / p.k_splits, split = blockIdx.z - cls * p.k_splits;
    rh = cls / g.S; rw = cls - rh * g.S;
    hi_first = ((rh - g.PT) % g.S + g.S) % g.S;
    wi_first = ((rw - g.PL) % g.S + g.S) % g.S;
    Hc = hi_first < g.H ? (g.H - hi_first + g.S - 1) / g.S : 0;
    Wc = wi_first < g.W ? (g.W - wi_first + g.S - 1) / g.S : 0;
    const int nth = rh < g.KH ? (g.KH - rh + g.S - 1) / g.S : 0;
    ntw = rw < g.KW ? (g.KW - rw + g.S - 1) / g.S : 0;
    Mg = g.B * Hc * Wc; Ng = g.Cin;
    const int Kc = nth * ntw * g.Cout;
    const int kper = (((Kc + p.k_splits - 1) / p.k_splits) + 63) / 64 * 64;
    k_begin = split * kper; k_end = min(Kc, k_begin + kper);
  } else {
    Mg = g.K + (p.dbias ? 1 : 0); Ng = g.Cout;
    k_begin = blockIdx.z * p.m_per_split; k_end = min(g.M, k_begin + p.m_per_split);
  }
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * NT;
  if (m0 >= Mg || k_begin >= k_end) return;      // uniform per CTA (before any barrier / TMEM use)
  const int nkb = (k_end - k_begin + KB - 1) / KB;

  if (tid == 0) {
    for (int i = 0; i < CT_STAGES; ++i) { mbar_init(full + i, 256); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ================================================================ MMA issuer
    if (elect_one_sync()) {
      constexpr uint32_t IDESC = TF32 ? make_idesc_tf32_conv(128, NT)
                                      : make_idesc_bf16(128, NT, MODE == TC_WGRAD ? 1 : 0, MODE == TC_DGRAD ? 0 : 1);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % CT_STAGES;
        mbar_wait(full + s, (kb / CT_STAGES) & 1);
        tc_fence_after();
        const uint32_t aA = smem_u32(smem + s * L::STAGE), aB = aA + L::A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (TF32) {     // 8 tf32 = 32 bytes of K per instruction, both operands K-major SW128
            mma_tf32_ss_conv(tmem_base, make_desc_sw128(aA) + (uint64_t)(ks * 2), make_desc_sw128(aB) + (uint64_t)(ks * 2),
                             IDESC, (kb > 0) || (ks > 0));
            continue;
          }
          auto desc_a = [&](uint32_t a) {
            return MODE == TC_WGRAD ? make_desc_sw128_mn(a + ks * 2048, 8192, 1024) : make_desc_sw128(a) + (uint64_t)(ks * 2);
          };
          auto desc_b = [&](uint32_t b) {
            return MODE == TC_DGRAD ? make_desc_sw128(b) + (uint64_t)(ks * 2) : make_desc_sw128_mn(b + ks * 2048, 8192, 1024);
          };
          const uint64_t da = desc_a(aA), db = desc_b(aB);
          mma_bf16_ss(tmem_base, da, db, IDESC, (kb > 0) || (ks > 0));
          if (SPLIT) {      // + hi * lo + lo * hi (the lo * lo term is below 2^-16 of the product)
            mma_bf16_ss(tmem_base, da, desc_b(aB + L::LO_OFF), IDESC, true);
            mma_bf16_ss(tmem_base, desc_a(aA + L::LO_OFF), db, IDESC, true);
          }
        }
        mma_commit(empty + s);
      }
      mma_commit(accum);
    }
  } else {
    // ================================================================ producers
    // A tile: 1024 chunks of 8 elements, 4 per thread.  K-major modes: thread -> rows (tid>>3) + 32 j, chunk tid & 7.
    //                                                   WGRAD:         thread -> m-rows (tid>>4) + 16 j, kk-chunk tid & 15.
    const int a_chunk = MODE == TC_WGRAD ? (tid & 15) : (tid & 7);
    int a_base[4];        // element offset of the row's pixel origin in a_src (FWD / DGRAD), -1 = row out of range
    int a_h0[4], a_w0[4];
    if (MODE != TC_WGRAD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + (tid >> 3) + 32 * j;
        if (m >= Mg) { a_base[j] = -1; a_h0[j] = 0; a_w0[j] = 0; continue; }
        if (MODE == TC_FWD) {
          const int b = m / (g.Ho * g.Wo), rem = m - b * (g.Ho * g.Wo);
          const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
          a_base[j] = b; a_h0[j] = ho * g.S - g.PT; a_w0[j] = wo * g.S - g.PL;
        } else {
          const int b = m / (Hc * Wc), rem = m - b * (Hc * Wc);
          const int ih = rem / Wc, iw = rem - ih * Wc;
          const int hi = hi_first + ih * g.S, wi = wi_first + iw * g.S;
          a_base[j] = b; a_h0[j] = (hi + g.PT - rh) / g.S; a_w0[j] = (wi + g.PL - rw) / g.S;   // ho = h0 - th
        }
      }
    }
    // WGRAD: the kk columns of this thread's chunk are fixed: decode (kh, kw, ci0) once
    int wg_kh = 0, wg_kw = 0, wg_ci = 0, wg_kind = 0;   // kind: 0 = im2col, 1 = ones column (bias), 2 = beyond range
    if (MODE == TC_WGRAD) {
      const int kk = m0 + a_chunk * 8;
      if (kk >= g.K) wg_kind = (kk == g.K && p.dbias) ? 1 : 2;
      else { const int tap = kk / g.Cin; wg_ci = kk - tap * g.Cin; wg_kh = tap / g.KW; wg_kw = tap - wg_kh * g.KW; }
    }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % CT_STAGES;
      if (kb >= CT_STAGES) mbar_wait(empty + s, ((kb / CT_STAGES) - 1) & 1);
      uint8_t* sA = smem + s * L::STAGE;
      uint8_t* sB = sA + L::A_BYTES;
      const int kbase = k_begin + kb * KB;
      // ------------------------------------------------ A tile
      float4 va0[4], va1[4];
      if (MODE == TC_FWD || MODE == TC_DGRAD) {
        const int k0 = kbase + a_chunk * CE;
        const int cch = MODE == TC_FWD ? g.Cin : g.Cout;          // channels per tap
        const int tap = k0 / cch, c0 = k0 - tap * cch;
        int dh, dw;
        if (MODE == TC_FWD) { dh = tap / g.KW; dw = tap - dh * g.KW; }
        else { dh = tap / ntw; dw = tap - dh * ntw; }
        // all global loads of the block are issued before the first shared-memory store: the gather is latency-bound
        // (L2 round trips), so the loads of a thread must be in flight together
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          va0[j] = z4; va1[j] = z4;
          if (a_base[j] >= 0 && k0 < k_end) {
            if (MODE == TC_FWD) {
              const int hi = a_h0[j] + dh, wi = a_w0[j] + dw;
              if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
                const float* src = p.a_src + ((size_t)(a_base[j] * g.H + hi) * g.W + wi) * g.Cin + c0;
                va0[j] = ld4(src);
                if (!TF32) va1[j] = ld4(src + 4);
              }
            } else {
              const int ho = a_h0[j] - dh, wo = a_w0[j] - dw;
              if (ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo) {
                const float* src = p.a_src + ((size_t)(a_base[j] * g.Ho + ho) * g.Wo + wo) * g.Cout + c0;
                va0[j] = ld4(src);
                if (!TF32) va1[j] = ld4(src + 4);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mr = (tid >> 4) + 16 * j;          // pixel row inside the 64-pixel block
          const int m = kbase + mr;
          va0[j] = z4; va1[j] = z4;
          if (m < k_end) {
            if (wg_kind == 0) {
              const int b = m / (g.Ho * g.Wo), rem = m - b * (g.Ho * g.Wo);
              const int ho = rem / g.Wo, wo = rem - ho * g.Wo;
              const int hi = ho * g.S - g.PT + wg_kh, wi = wo * g.S - g.PL + wg_kw;
              if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
                const float* src = p.a_src + ((size_t)(b * g.H + hi) * g.W + wi) * g.Cin + wg_ci;
                va0[j] = ld4(src); va1[j] = ld4(src + 4);
              }
            } else if (wg_kind == 1) {
              va0[j].x = 1.0f;                          // ones column -> bias gradient row
            }
          }
        }
      }
      auto store_a = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint8_t* dst;
          if (MODE == TC_WGRAD) dst = sA + (a_chunk >> 3) * 8192 + sw128_offset((tid >> 4) + 16 * j, a_chunk & 7);
          else dst = sA + sw128_offset((tid >> 3) + 32 * j, a_chunk);
          if (TF32) *reinterpret_cast<float4*>(dst) = va0[j];
          else if (SPLIT) st_chunk_split(dst, dst + L::LO_OFF, va0[j], va1[j]);
          else st_chunk(dst, va0[j], va1[j]);
        }
      };
      // ------------------------------------------------ B tile
      if (TF32 && MODE == TC_FWD) {
        // K-major rows n (NT), 8 chunks of 4 k: the HWIO kernel w[k][n] gathered transposed (4 strided scalar loads;
        // the kernel is small and L2-resident)
        constexpr int NCH = NT * 8;
        constexpr int NIT = (NCH + 255) / 256;
        float4 vb[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          const int n = c >> 3, ch = c & 7;
          const int k0 = kbase + ch * 4;
          vb[it] = z4;
          if (c < NCH && n0 + n < Ng) {
            const float* src = p.b_src + (size_t)k0 * g.Cout + n0 + n;
            if (k0 + 0 < k_end) vb[it].x = src[0];
            if (k0 + 1 < k_end) vb[it].y = src[g.Cout];
            if (k0 + 2 < k_end) vb[it].z = src[2 * (size_t)g.Cout];
            if (k0 + 3 < k_end) vb[it].w = src[3 * (size_t)g.Cout];
          }
        }
        store_a();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          if (c < NCH) *reinterpret_cast<float4*>(sB + sw128_offset(c >> 3, c & 7)) = vb[it];
        }
      } else if (MODE == TC_DGRAD) {
        // K-major rows n (NT), 8 chunks of k: source w[(tap * Cin + n) * Cout + co0 ..]
        constexpr int NCH = NT * 8;
        constexpr int NIT = (NCH + 255) / 256;
        float4 vb0[NIT], vb1[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          const int n = c >> 3, ch = c & 7;
          const int k0 = kbase + ch * CE;
          vb0[it] = z4; vb1[it] = z4;
          if (c < NCH && n0 + n < Ng && k0 < k_end) {
            const int tap = k0 / g.Cout, co0 = k0 - tap * g.Cout;
            const int th = tap / ntw, tw = tap - th * ntw;
            const int kh = rh + th * g.S, kw = rw + tw * g.S;
            const float* src = p.b_src + ((size_t)(kh * g.KW + kw) * g.Cin + n0 + n) * g.Cout + co0;
            vb0[it] = ld4(src);
            if (!TF32) vb1[it] = ld4(src + 4);
          }
        }
        store_a();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          if (c < NCH) {
            uint8_t* dst = sB + sw128_offset(c >> 3, c & 7);
            if (TF32) *reinterpret_cast<float4*>(dst) = vb0[it];
            else if (SPLIT) st_chunk_split(dst, dst + L::LO_OFF, vb0[it], vb1[it]);
            else st_chunk(dst, vb0[it], vb1[it]);
          }
        }
      } else {
        // MN-major: 64 reduction rows, NT/8 chunks of n: source rows are contiguous in n (w[k][:] or dy[m][:])
        constexpr int CPR = NT / 8;                    // chunks per row
        constexpr int NCH = 64 * CPR;
        constexpr int NIT = (NCH + 255) / 256;
        float4 vb0[NIT], vb1[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          const int r = c / CPR, ch = c - r * CPR;
          const int kk = kbase + r, n = n0 + ch * 8;
          vb0[it] = z4; vb1[it] = z4;
          if (c < NCH && kk < k_end && n < Ng) {
            const float* src = p.b_src + (size_t)kk * g.Cout + n;
            if ((g.Cout & 7) == 0) {
              vb0[it] = ld4(src); vb1[it] = ld4(src + 4);
            } else {   // ragged channel count (the 3-channel image head): guarded scalar loads
              float t8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) t8[e] = (n + e < Ng) ? src[e] : 0.f;
              vb0[it] = make_float4(t8[0], t8[1], t8[2], t8[3]); vb1[it] = make_float4(t8[4], t8[5], t8[6], t8[7]);
            }
          }
        }
        store_a();
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int c = tid + it * 256;
          if (c < NCH) {
            const int r = c / CPR, ch = c - r * CPR;
            uint8_t* dst = sB + (ch >> 3) * 8192 + sw128_offset(r, ch & 7);
            if (SPLIT) st_chunk_split(dst, dst + L::LO_OFF, vb0[it], vb1[it]);
            else st_chunk(dst, vb0[it], vb1[it]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(full + s);
    }

    // ================================================================ epilogue
    mbar_wait(accum, 0);
    tc_fence_after();
    const int qd = warp & 3, hh = warp >> 2;           // TMEM lane quarter, column half
    const int row = qd * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(qd * 32) << 16);
    constexpr int HALF = NT >= 32 ? NT / 2 : NT;       // NT = 16: only the hh == 0 warps work
    if (NT >= 32 || hh == 0) {
      const int cbase = NT >= 32 ? hh * HALF : 0;
      const int m = m0 + row;
      float* orow = nullptr;
      if (m < Mg) {
        if (MODE == TC_FWD) orow = p.out + (size_t)m * g.Cout;
        else if (MODE == TC_DGRAD) {
          const int b = m / (Hc * Wc), rem = m - b * (Hc * Wc);
          const int ih = rem / Wc, iw = rem - ih * Wc;
          orow = p.out + ((size_t)(b * g.H + hi_first + ih * g.S) * g.W + wi_first + iw * g.S) * g.Cin;
        } else orow = (m < g.K) ? p.out + (size_t)m * g.Cout : p.dbias;
      }
      const bool row_ok = m < Mg;
      const float res_gm = (MODE == TC_FWD && p.res) ? *p.res_scale : 0.f;
#pragma unroll
      for (int c = 0; c < HALF; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + cbase + c, r);
        tmem_wait_ld();
        if (row_ok) {
          const int nb = n0 + cbase + c;
          if (MODE == TC_FWD && p.q_out) {
            // fused q / k / v projection epilogue: bf16 rows in the layouts the flash kernel's TMA boxes read
            if (nb < Ng) {
              float v[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + (p.bias ? p.bias[nb + e] : 0.f);
              __nv_bfloat16* dst;
              if (nb < p.qk_d) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] *= p.q_scale;
                dst = p.q_out + (size_t)m * 64 + nb;
              } else if (nb < 2 * p.qk_d) {
                dst = p.k_out + (size_t)m * 64 + (nb - p.qk_d);
              } else {
                dst = p.v_out + (size_t)m * p.v_dv + (nb - 2 * p.qk_d);
              }
              uint4* d4 = reinterpret_cast<uint4*>(dst);
              d4[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              d4[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
            }
          } else if (MODE == TC_FWD && p.res) {
            if (nb + 16 <= Ng) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 x4 = ld4(p.res + (size_t)m * g.Cout + nb + e);
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __uint_as_float(r[e + q]) + (p.bias ? p.bias[nb + e + q] : 0.f);
                st4(orow + nb + e, make_float4(fmaf(res_gm, v[0], x4.x), fmaf(res_gm, v[1], x4.y), fmaf(res_gm, v[2], x4.z),
                                               fmaf(res_gm, v[3], x4.w)));
              }
            } else {
              for (int e = 0; e < 16; ++e)
                if (nb + e < Ng)
                  orow[nb + e] = fmaf(res_gm, __uint_as_float(r[e]) + (p.bias ? p.bias[nb + e] : 0.f),
                                      p.res[(size_t)m * g.Cout + nb + e]);
            }
          } else if (MODE != TC_WGRAD && p.k_splits > 1) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (nb + e < Ng) atomicAdd(orow + nb + e, __uint_as_float(r[e]));
          } else if (MODE == TC_WGRAD) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (nb + e < Ng) atomicAdd(orow + nb + e, __uint_as_float(r[e]));
          } else if (nb + 16 <= Ng) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[q] = __uint_as_float(r[e + q]);
                if (MODE == TC_FWD) v[q] = ct_act(v[q] + (p.bias ? p.bias[nb + e + q] : 0.f), p.act, p.slope);
              }
              st4(orow + nb + e, make_float4(v[0], v[1], v[2], v[3]));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (nb + e < Ng) {
                float v = __uint_as_float(r[e]);
                if (MODE == TC_FWD) v = ct_act(v + (p.bias ? p.bias[nb + e] : 0.f), p.act, p.slope);
                orow[nb + e] = v;
              }
          }
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

template <int MODE, int NT, int PREC>
static int launch_conv_tc(const ConvTcP& p, dim3 grid, cudaStream_t st) {
  using L = ConvTcSmem<MODE, NT, PREC>;
  auto kern = conv_tc_kernel<MODE, NT, PREC>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  kern<<<grid, CT_THREADS, L::TOTAL, st>>>(p);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

template <int MODE, int PREC = PREC_BF16>
static int dispatch_nt(const ConvTcP& p, int Ng, int gx, int gz, cudaStream_t st) {
  if (Ng <= 16) return launch_conv_tc<MODE, 16, PREC>(p, dim3(gx, 1, gz), st);
  if (Ng <= 32) return launch_conv_tc<MODE, 32, PREC>(p, dim3(gx, 1, gz), st);
  if (Ng <= 64) return launch_conv_tc<MODE, 64, PREC>(p, dim3(gx, 1, gz), st);
  return launch_conv_tc<MODE, 128, PREC>(p, dim3(gx, ceil_div(Ng, 128), gz), st);
}

// 2 (default): split-bf16 everywhere; 1: round-1 arithmetic (tf32 forward / backward-data, plain bf16 backward-filter)
static std::atomic<int> g_conv_prec{PREC_SPLIT};
int conv_tc_get_precision() { return g_conv_prec.load(); }
void conv_tc_set_precision(int p) { g_conv_prec.store(p == PREC_TF32 ? PREC_TF32 : PREC_SPLIT); }

static inline bool al16(const void* q) { return (((uintptr_t)q) & 15) == 0; }

// eligibility: 8-element chunks must stay inside one tap and be 16-byte aligned
bool conv_tc_fwd_ok(const CG& g, const float* x, const float* w) { return g.Cin % 8 == 0 && al16(x) && al16(w); }
bool conv_tc_dgrad_ok(const CG& g, const float* dy, const float* w) { return g.Cout % 8 == 0 && g.Cin % 4 == 0 && al16(dy) && al16(w); }
bool conv_tc_wgrad_ok(const CG& g, const float* x, const float* dy) { return g.Cin % 8 == 0 && al16(x) && al16(dy); }

// y = act(y + bias) after a split-K accumulation
__global__ void conv_tc_finish_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n, int C, int act,
                                      float slope) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = ct_act(y[i] + (bias ? bias[(int)(i % C)] : 0.f), act, slope);
}

// Layers with few output pixels (the 4x4 / 8x8 maps: 8-16 CTAs walking 16-32 K blocks each) are latency-bound on the
// producers' gather; splitting K over more CTAs fills the machine.  Returns 1 when splitting does not pay.
static int pick_k_splits(int ctas, int K) {
  if (g_deterministic_forward.load()) return 1;      // split-K accumulates with fp32 atomics: run-to-run rounding noise
  const int nkb = ceil_div(K, 64);
  if (ctas * 2 > num_sms() || nkb < 8) return 1;
  return std::max(1, std::min(nkb / 4, num_sms() / ctas));
}

int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, const CG& g, int act, float slope,
                cudaStream_t st) {
  ConvTcP p{x, w, bias, y, nullptr, g, act, slope, 0, 1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f};
  const int gx = ceil_div(g.M, 128);
  p.k_splits = pick_k_splits(gx * ceil_div(g.Cout, 128), g.K);
  const bool split = g_conv_prec.load() == PREC_SPLIT;
  if (p.k_splits == 1)
    return split ? dispatch_nt<TC_FWD, PREC_SPLIT>(p, g.Cout, gx, 1, st) : dispatch_nt<TC_FWD, PREC_TF32>(p, g.Cout, gx, 1, st);
  const long long n = (long long)g.M * g.Cout;
  SAGAN_CUDA(cudaMemsetAsync(y, 0, (size_t)n * sizeof(float), st));
  int rc = split ? dispatch_nt<TC_FWD, PREC_SPLIT>(p, g.Cout, gx, p.k_splits, st)
                 : dispatch_nt<TC_FWD, PREC_TF32>(p, g.Cout, gx, p.k_splits, st);
  if (rc) return rc;
  if (bias || act != SAGAN_ACT_NONE) {
    conv_tc_finish_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, st>>>(y, bias, n, g.Cout, act, slope);
    SAGAN_LAUNCH_CHECK();
  }
  return 0;
}

static CG gemm_geom(long long M, int K, int N) {
  CG g{};
  g.B = (int)M; g.H = 1; g.W = 1; g.Cin = K; g.Ho = 1; g.Wo = 1; g.Cout = N; g.KH = 1; g.KW = 1; g.S = 1; g.PT = 0; g.PL = 0;
  g.M = (int)M; g.K = K;
  return g;
}

// [M, K] fp32 x [K, 2d + dv] fp32 (+ bias) -> bf16 q (scaled) / k rows of 64 columns and v rows of dv columns
int gemm_tc_qkv(const float* x, const float* wcat, const float* bcat, __nv_bfloat16* q, __nv_bfloat16* k,
                __nv_bfloat16* v, long long M, int K, int d, int dv, float q_scale, cudaStream_t st) {
  const CG g = gemm_geom(M, K, 2 * d + dv);
  ConvTcP p{x, wcat, bcat, nullptr, nullptr, g, 0, 0.f, 0, 1, nullptr, nullptr, q, k, v, d, dv, q_scale};
  return dispatch_nt<TC_FWD>(p, g.Cout, ceil_div(g.M, 128), 1, st);
}

// y = res + (*res_scale) * (x [M, K] * w [K, N] + bias)
int gemm_tc_residual(const float* x, const float* w, const float* bias, const float* res, const float* res_scale, float* y,
                     long long M, int K, int N, cudaStream_t st) {
  const CG g = gemm_geom(M, K, N);
  ConvTcP p{x, w, bias, y, nullptr, g, 0, 0.f, 0, 1, res, res_scale, nullptr, nullptr, nullptr, 0, 0, 0.f};
  return dispatch_nt<TC_FWD>(p, g.Cout, ceil_div(g.M, 128), 1, st);
}

int conv_tc_dgrad(const float* dy, const float* w, float* dx, const CG& g, cudaStream_t st) {
  ConvTcP p{dy, w, nullptr, dx, nullptr, g, 0, 0.f, 0, 1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f};
  const int Hc = ceil_div(g.H, g.S), Wc = ceil_div(g.W, g.S);
  const int gx = ceil_div(g.B * Hc * Wc, 128), classes = g.S * g.S;
  const int Kc = ceil_div(g.KH, g.S) * ceil_div(g.KW, g.S) * g.Cout;      // reduction length of the largest class
  p.k_splits = pick_k_splits(gx * ceil_div(g.Cin, 128) * classes, Kc);
  if (p.k_splits > 1) SAGAN_CUDA(cudaMemsetAsync(dx, 0, (size_t)g.B * g.H * g.W * g.Cin * sizeof(float), st));
  if (g_conv_prec.load() == PREC_SPLIT) return dispatch_nt<TC_DGRAD, PREC_SPLIT>(p, g.Cin, gx, classes * p.k_splits, st);
  return dispatch_nt<TC_DGRAD, PREC_TF32>(p, g.Cin, gx, classes * p.k_splits, st);
}

int conv_tc_wgrad(const float* x, const float* dy, float* dw, float* dbias, const CG& g, cudaStream_t st) {
  const int Mg = g.K + (dbias ? 1 : 0);
  const int tiles = ceil_div(Mg, 128) * ceil_div(g.Cout, 128);
  int splits = std::max(1, std::min(ceil_div(g.M, 256), ceil_div(num_sms() * 2, tiles)));
  int mps = ceil_div(ceil_div(g.M, splits), 64) * 64;
  splits = ceil_div(g.M, mps);
  ConvTcP p{x, dy, nullptr, dw, dbias, g, 0, 0.f, mps, 1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.f};
  if (g_conv_prec.load() == PREC_SPLIT) return dispatch_nt<TC_WGRAD, PREC_SPLIT>(p, g.Cout, ceil_div(Mg, 128), splits, st);
  return dispatch_nt<TC_WGRAD>(p, g.Cout, ceil_div(Mg, 128), splits, st);
}

}  // namespace sagan
