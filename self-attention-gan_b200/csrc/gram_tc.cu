// Gram GEMM over the token axis for the large-C attention backward:  G[l, r] = sum_t L[t, l] R[t, r]  (+ column sums of R)
//
// The two weight-gradient products of the block ( dWo' = A^T dY  and  [dWq | dWk | dWv] = X^T [dQ | dK | dV],
// /root/reference/layers.py:93-120 differentiated by tf.GradientTape ) reduce over T = B N tokens of token-major fp32
// activations.  tcgen05 reads such operands MN-major only as bf16 (kind::tf32 with an MN-major operand returns zeros,
// DESIGN.md §9), so the fp32 tiles are brought in by TMA and split into bf16 hi + lo tiles INSIDE the CTA:
//   warp 16     TMA producer: per 32-token stage, L[32 tok][128 ch] and R[32 tok][128 ch] fp32 as 4 + 4 SWIZZLE_128B boxes
//   warps 0-15  converters: fp32 tile -> bf16 hi / lo tiles in the MN-major SWIZZLE_128B operand layout (token rows of
//               64 channels = 128 B); the same threads keep the column sums of R and run the epilogue
//   warp 17     MMA issuer: hi*hi + hi*lo + lo*hi (the lo*lo term is below 2^-16 of the product), fp32 accumulator in TMEM
// Reduction split over CTAs along the tokens (grid.z), tiles of the same token range adjacent in launch order so the
// operands are read from HBM once and re-read from L2; partial tiles added to G with 16-byte fp32 atomics.
// Replaces the backward-filter form of the conv kernel for these two products (software im2col): 535 -> 288 us of the
// backward at B=16, N=4096, C=512 (1.50 -> 1.23 ms).  What bounds it now is L2 -> SM operand traffic: 128 x 128 tiles
// re-read every operand tile 3-4 times, 1.34 GB per backward at the ~4.7 TB/s this part delivers from L2 to the SMs.
#include "common.cuh"
#include "tc_common.cuh"

namespace sagan {

using namespace tc;

constexpr int GR_CONV = 512;                  // converter threads (16 warps: four per scheduler hide the shared-memory latency)
constexpr int GR_THREADS = GR_CONV + 64;
constexpr int GR_TOK = 32;                    // tokens per stage
constexpr int GR_BOX = GR_TOK * 128;          // one fp32 box: [32 tokens][32 channels]
constexpr int GR_SUB = GR_TOK * 128;          // one bf16 sub-tile: [32 tokens][64 channels]
// NT = width of the right operand's tile (128 or 256 channels): the wider tile re-reads the left operand half as often
template <int NT>
struct GrSmem {
  static constexpr int NF = NT == 256 ? 2 : 3, NB = 2;          // fp32 stages, bf16 buffers
  static constexpr int RBOX = NT / 32;                          // fp32 boxes of the right operand per stage
  static constexpr int FSTAGE = (4 + RBOX) * GR_BOX;            // [L: 4 boxes | R: RBOX boxes]
  static constexpr int LT = 2 * GR_SUB, RT = (NT / 64) * GR_SUB;   // bf16 tiles
  static constexpr int BBUF = 2 * (LT + RT);                    // [L_hi | R_hi | L_lo | R_lo]
  static constexpr int LO = LT + RT;
  static constexpr int OFF_B = NF * FSTAGE;
  static constexpr int OFF_BAR = OFF_B + NB * BBUF;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

template <int NT>
__global__ void __launch_bounds__(GR_THREADS, 1)
gram_split_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, float* __restrict__ G,
                  float* __restrict__ colsum, int l, int r, long long T, int tok_per_split) {
  using S = GrSmem<NT>;
  constexpr int GR_NF = S::NF, GR_NB = S::NB, GR_FSTAGE = S::FSTAGE, GR_BBUF = S::BBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sF = smem;
  uint8_t* sB = smem + S::OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* fullF = bars;                 // [NF] TMA bytes landed
  uint64_t* freeF = bars + GR_NF;         // [NF] 128 arrivals: the converters have read the stage
  uint64_t* fullB = bars + 2 * GR_NF;     // [NB] 128 arrivals: bf16 tiles written
  uint64_t* freeB = fullB + GR_NB;        // [NB] tcgen05.commit: the MMAs that read the buffer are done
  uint64_t* accum = freeB + GR_NB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5;
  const int l0 = blockIdx.x * 128, r0 = blockIdx.y * NT;
  const long long t_begin = (long long)blockIdx.z * tok_per_split;
  const long long t_end = t_begin + tok_per_split < T ? t_begin + tok_per_split : T;
  if (t_begin >= t_end) return;                         // uniform per CTA, before any barrier / TMEM use
  const int nst = (int)((t_end - t_begin + GR_TOK - 1) / GR_TOK);

  if (threadIdx.x == 0) {
    for (int i = 0; i < GR_NF; ++i) { mbar_init(fullF + i, 1); mbar_init(freeF + i, GR_CONV); }
    for (int i = 0; i < GR_NB; ++i) { mbar_init(fullB + i, GR_CONV); mbar_init(freeB + i, 1); }
    mbar_init(accum, 1);
    mbar_fence_init();
  }
  if (warp == 16) tmem_alloc(tmem_ptr, NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 16) {
    // ================================================================ TMA producer
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmL); tma_prefetch_desc(&tmR);
      int s = 0, u = 0;
      for (int i = 0; i < nst; ++i) {
        if (u > 0) mbar_wait(freeF + s, (u - 1) & 1);
        uint8_t* st = sF + s * GR_FSTAGE;
        const int t = (int)(t_begin + (long long)i * GR_TOK);      // rows beyond T (and beyond this split) read as zero or are
        mbar_expect_tx(fullF + s, GR_FSTAGE);                      // whole stages of the next split: see the host's split size
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(st + c * GR_BOX, &tmL, fullF + s, l0 + c * 32, t);
#pragma unroll
        for (int c = 0; c < S::RBOX; ++c) tma_load_2d(st + (4 + c) * GR_BOX, &tmR, fullF + s, r0 + c * 32, t);
        if (++s == GR_NF) { s = 0; ++u; }
      }
    }
  } else if (warp == 17) {
    // ================================================================ MMA issuer
    if (elect_one_sync()) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, NT, /*a_mn_major=*/1, /*b_mn_major=*/1);
      int b = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nst; ++i) {
        mbar_wait(fullB + b, ph);
        tc_fence_after();
        const uint32_t base = smem_u32(sB + b * GR_BBUF);
#pragma unroll
        for (int ks = 0; ks < GR_TOK / 16; ++ks) {     // 16 tokens per step; sub-tiles of 64 channels are 4 KB apart
          const uint64_t lh = make_desc_sw128_mn(base + ks * 2048, 4096, 1024);
          const uint64_t rh = make_desc_sw128_mn(base + S::LT + ks * 2048, 4096, 1024);
          const uint64_t ll = make_desc_sw128_mn(base + S::LO + ks * 2048, 4096, 1024);
          const uint64_t rl = make_desc_sw128_mn(base + S::LO + S::LT + ks * 2048, 4096, 1024);
          mma_bf16_ss(tmem_base, lh, rh, IDESC, (i > 0) || (ks > 0));
          mma_bf16_ss(tmem_base, lh, rl, IDESC, true);
          mma_bf16_ss(tmem_base, ll, rh, IDESC, true);
        }
        mma_commit(freeB + b);
        if (++b == GR_NB) { b = 0; ph ^= 1; }
      }
      mma_commit(accum);
    }
  } else {
    // ================================================================ converters (512 threads), then epilogue
    // thread -> token row (tid >> 2) & 31 of every stage, channels [32 k + 8 m, + 8) of BOTH operands (k = tid >> 7,
    // m = tid & 3): per warp instruction 8 consecutive token rows x 4 of their 8 sixteen-byte positions, which the 128-byte
    // swizzle spreads over all banks.  Measured on the way: 4 converter warps with 16 channel groups x 2 rows per warp
    // (4-way bank conflicts on the fp32 reads) 172 us per product, conflict-free 157 us -- one warp per scheduler is
    // bound by the latency of its own load -> convert -> store chain; sixteen warps hide it.
    const int tid = threadIdx.x;
    const int tok = (tid >> 2) & 31, m = tid & 3, k = tid >> 7;
    const uint32_t sw = (uint32_t)(tok & 7);
    const uint32_t src_off = (uint32_t)(k * GR_BOX + tok * 128) + (((uint32_t)(2 * m) ^ sw) << 4);        // logical chunks 2m, 2m+1
    const uint32_t src_off1 = (uint32_t)(k * GR_BOX + tok * 128) + (((uint32_t)(2 * m + 1) ^ sw) << 4);
    // channel group 4 k + m of the tile: 64-channel sub-tile k >> 1, 16-byte position (4 (k & 1) + m) ^ swizzle
    const uint32_t dst_off = (uint32_t)((k >> 1) * GR_SUB + tok * 128) + ((((uint32_t)(4 * (k & 1) + m)) ^ sw) << 4);
    const bool want_sum = colsum != nullptr && blockIdx.x == 0;
    constexpr int NR = NT / 128;           // right-operand boxes per thread: k, k + 4
    float csum[NR][8];
#pragma unroll
    for (int j = 0; j < NR; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) csum[j][e] = 0.f;
    int s = 0, b = 0, ub = 0;
    uint32_t phF = 0;
    for (int i = 0; i < nst; ++i) {
      mbar_wait(fullF + s, phF);
      if (ub > 0) {
        mbar_wait(freeB + b, (ub - 1) & 1);
        tc_fence_after();
      }
      const uint8_t* st = sF + s * GR_FSTAGE;
      uint8_t* bb = sB + b * GR_BBUF;
#pragma unroll
      for (int u = 0; u < 1 + NR; ++u) {   // u = 0: left operand box k; u >= 1: right operand box k + 4 (u - 1)
        const uint8_t* src = st + (u == 0 ? 0 : (4 + 4 * (u - 1)) * GR_BOX);
        const float4 a = *reinterpret_cast<const float4*>(src + src_off);
        const float4 c = *reinterpret_cast<const float4*>(src + src_off1);
        const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        if (u >= 1 && want_sum) {
#pragma unroll
          for (int e = 0; e < 8; ++e) csum[u >= 1 ? u - 1 : 0][e] += v[e];
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t h = pack_bf16x2(v[2 * e], v[2 * e + 1]);
          hi[e] = h;
          lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(h << 16), v[2 * e + 1] - __uint_as_float(h & 0xffff0000u));
        }
        // right operand: channel 128 (u - 1) + 32 k + 8 m -> sub-tiles 2 (u - 1) + (k >> 1)
        uint8_t* dst = bb + (u == 0 ? 0 : S::LT + 2 * (u - 1) * GR_SUB) + dst_off;
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + S::LO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();          // st.shared of the bf16 tiles -> visible to the tensor core (async proxy)
      mbar_arrive(fullB + b);
      mbar_arrive(freeF + s);
      if (++s == GR_NF) { s = 0; phF ^= 1; }
      if (++b == GR_NB) { b = 0; ++ub; }
    }
    if (want_sum) {      // fold the 8 token rows of the warp (lanes with the same m), then one atomic per channel and warp
#pragma unroll
      for (int j = 0; j < NR; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float q = csum[j][e];
          q += __shfl_xor_sync(0xffffffffu, q, 4);
          q += __shfl_xor_sync(0xffffffffu, q, 8);
          q += __shfl_xor_sync(0xffffffffu, q, 16);
          const int col = r0 + 128 * j + 32 * k + 8 * m + e;
          if ((tid & 31) < 4 && col < r) atomicAdd(colsum + col, q);
        }
    }
    // ---- epilogue: warp w <-> TMEM lane quarter w & 3 (rows) and columns [32 (w >> 2), + 32); 16-byte atomics along the row
    mbar_wait(accum, 0);
    tc_fence_after();
    const int row = l0 + (warp & 3) * 32 + (tid & 31);
#pragma unroll
    for (int j = 0; j < NT / 128; ++j) {
      const int c0 = 128 * j + (warp >> 2) * 32;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
      tmem_wait_ld();
      if (row < l) {
        float* dst = G + (size_t)row * r + r0 + c0;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          if (r0 + c0 + e < r)
            atomicAdd(reinterpret_cast<float4*>(dst + e), make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                     __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NT);
  }
}

// 2D fp32 tensor [rows][cols] (cols contiguous), box [32 rows][32 cols = 128 B], SWIZZLE_128B, OOB -> 0
static int gram_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled is not available from this driver");
    return SAGAN_EUNSUPPORTED;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)GR_TOK};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled(gram) failed with CUresult %d (rows=%llu cols=%llu)", (int)rc, (unsigned long long)rows,
            (unsigned long long)cols);
    return SAGAN_EINVAL;
  }
  return 0;
}

// (below 128 channels on either side the 128-row tile is half empty and the conv form measured faster: C = 128)
bool gram_split_supported(int l, int r, long long T) {
  return l % 4 == 0 && r % 4 == 0 && l >= 128 && r >= 128 && T >= GR_TOK && T < (1ll << 31);
}

// G [l, r] (row-major) = L^T R, colsum [r] = column sums of R (or nullptr); L [T, l], R [T, r] fp32 token-major
int gram_split(const float* L, int l, const float* R, int r, float* G, float* colsum, long long T, cudaStream_t st) {
  SAGAN_REQUIRE(gram_split_supported(l, r, T), "gram_split: unsupported shape l=%d r=%d T=%lld", l, r, T);
  CUtensorMap tl, tr;
  int rc;
  if ((rc = gram_tmap(&tl, L, (uint64_t)T, (uint64_t)l))) return rc;
  if ((rc = gram_tmap(&tr, R, (uint64_t)T, (uint64_t)r))) return rc;
  SAGAN_CUDA(cudaMemsetAsync(G, 0, (size_t)l * r * sizeof(float), st));
  if (colsum) SAGAN_CUDA(cudaMemsetAsync(colsum, 0, (size_t)r * sizeof(float), st));
  const int NT = r > 128 ? 256 : 128;
  const int tiles = ceil_div(l, 128) * ceil_div(r, NT);
  const long long stages = ceil_div<long long>(T, GR_TOK);
  long long splits = ceil_div<long long>(2 * num_sms(), tiles);          // ~2 CTAs' worth of work per SM in all
  if (splits > stages) splits = stages;
  if (splits < 1) splits = 1;
  const long long st_per_split = ceil_div<long long>(stages, splits);
  splits = ceil_div<long long>(stages, st_per_split);
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(gram_split_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GrSmem<128>::TOTAL));
    SAGAN_CUDA(cudaFuncSetAttribute(gram_split_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GrSmem<256>::TOTAL));
    configured = true;
  }
  const dim3 grid(ceil_div(l, 128), ceil_div(r, NT), (unsigned)splits);
  if (NT == 256)
    gram_split_kernel<256><<<grid, GR_THREADS, GrSmem<256>::TOTAL, st>>>(tl, tr, G, colsum, l, r, T, (int)(st_per_split * GR_TOK));
  else
    gram_split_kernel<128><<<grid, GR_THREADS, GrSmem<128>::TOTAL, st>>>(tl, tr, G, colsum, l, r, T, (int)(st_per_split * GR_TOK));
  SAGAN_LAUNCH_CHECK();
  return 0;
}

}  // namespace sagan
