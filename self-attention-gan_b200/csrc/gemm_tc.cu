// TMA-fed tcgen05 GEMM over CTA PAIRS (cta_group::2) for the 1x1 convolutions of the large-C self-attention block
// (/root/reference/layers.py:82-85,99,104-105,112,119):
//
//   D [M, N] = A [M, K] * Bt [N, K]^T        A = activations (NHWC rows, K contiguous), Bt = transposed Keras kernel
//
// Two operand kinds share one kernel:
//   kind::tf32  fp32 operands read straight through TMA (the q/k/v projection: the activations X are fp32 in HBM and a
//               separate fp32 -> bf16 pass would cost more traffic than the GEMM itself; 10 mantissa bits kept)
//   kind::f16   bf16 operands (the output conv: the flash kernel emits the attention output A as bf16 next to fp32)
// These GEMMs are bound by L2 -> SM operand traffic, not by the tensor pipe (K <= 512 against 128-wide tiles), so the
// tiling minimises re-reads: a pair of CTAs computes a 256 x NT tile with ONE tcgen05.mma per K step (M = 256), each
// CTA loading its own 128 rows of A and only HALF of the B tile (NT / 2 rows of Bt); NT up to 256 columns keeps the
// re-reads of A across N tiles at one or two.  3-stage TMA ring, 6 warps per CTA (0-3 epilogue: TMEM lane = output
// row; 4 TMA producer; 5 MMA issuer, leader CTA only); two CTAs of different pairs are resident per SM (<= 96 KB of
// shared memory, <= 256 TMEM columns each), so one's epilogue runs under the other's main loop.
//
// Epilogues:
//   QKV       columns [0,d) -> bf16 q rows (64 wide) scaled by log2 e, [d,2d) -> k rows, [2d,2d+dv) -> v rows [T,dv]
//   RESIDUAL  y = res + gamma * (acc + bias)                                   (layers.py:119-120)
#include "common.cuh"
#include "tc_common.cuh"

namespace sagan {

using namespace tc;

constexpr int GT_THREADS = 192;
constexpr int GT_STAGES = 3;
constexpr int GT_A_BYTES = 128 * 128;     // [128 rows][128 B of K]

enum { GT_EPI_QKV = 0, GT_EPI_RESIDUAL = 1 };
enum { GT_TF32 = 0, GT_BF16 = 1 };

template <int NT>
struct GemmSmem {
  static constexpr int B_BYTES = (NT / 2) * 128;          // this CTA's half of the Bt tile
  static constexpr int STAGE = GT_A_BYTES + B_BYTES;
  static constexpr int TOTAL = GT_STAGES * STAGE + 256 + 1024;
  static constexpr int TMEM_COLS = NT <= 128 ? 128 : 256;
  static_assert(NT % 32 == 0 && NT <= 256, "pair MMA: N multiple of 32, at most 256");
  static_assert(4 * 32 * 36 * 4 <= GT_STAGES * STAGE, "epilogue transpose buffers live in the dead operand ring");
};

struct GemmP {
  int M, N, K;
  const float* bias;        // [N]
  // QKV
  __nv_bfloat16* q_out;
  __nv_bfloat16* k_out;
  __nv_bfloat16* v_out;
  int qk_d, v_dv;
  float q_scale;
  // RESIDUAL
  const float* res;         // [M, N]
  const float* res_scale;   // device scalar (gamma)
  float* out;               // [M, N]
};

// instruction descriptors: D = f32; A, B = tf32 (format 2) or bf16 (format 1), both K-major
__host__ __device__ constexpr uint32_t make_idesc_kind(int kind, int M, int N) {
  return (1u << 4) | ((kind == GT_TF32 ? 2u : 1u) << 7) | ((kind == GT_TF32 ? 2u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int KIND>
__device__ __forceinline__ void mma_pair_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  if constexpr (KIND == GT_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
  } else {
    mma_bf16_ss_g<2>(d_tmem, a_desc, b_desc, idesc, accumulate);
  }
}

template <int KIND, int EPI, int NT>
__global__ void __launch_bounds__(GT_THREADS, 2)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmP p) {
  using L = GemmSmem<NT>;
  constexpr int KB = KIND == GT_TF32 ? 32 : 64;            // K elements per 128-byte swizzle span
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * L::STAGE);
  uint64_t* full = bars;                  // [GT_STAGES] leader: the stage landed in both CTAs
  uint64_t* empty = bars + GT_STAGES;     // [GT_STAGES] every CTA: MMAs of the stage done
  uint64_t* accum = bars + 2 * GT_STAGES; //             every CTA: accumulator complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GT_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid: block = (m_pair * n_tiles + n_tile) * 2 + rank -- the N tiles of one 256-row slab are adjacent, so A is
  // re-read from L2, not from HBM; cluster = the two 128-row halves of the slab
  const uint32_t rank = cluster_ctarank();                 // == blockIdx.x & 1
  const bool leader = rank == 0;
  const int n_tiles = (p.N + NT - 1) / NT;
  const int tile = blockIdx.x >> 1;
  const int n0 = (tile % n_tiles) * NT, m0 = (tile / n_tiles) * 256 + (int)rank * 128;
  const int nkb = (p.K + KB - 1) / KB;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc_g<2>(tmem_ptr, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        if (kb >= GT_STAGES) mbar_wait(empty + s, ((kb / GT_STAGES) - 1) & 1);
        if (leader) mbar_expect_tx(full + s, 2 * L::STAGE);
        uint8_t* st = smem + s * L::STAGE;
        tma_load_2d_g<2>(st, &tmA, full + s, kb * KB, m0);                           // OOB rows / columns are zero-filled
        tma_load_2d_g<2>(st + GT_A_BYTES, &tmB, full + s, kb * KB, n0 + (int)rank * (NT / 2));
      }
    }
  } else if (warp == 5) {
    if (leader && elect_one_sync()) {
      constexpr uint32_t IDESC = make_idesc_kind(KIND, 256, NT);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        mbar_wait(full + s, (kb / GT_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = make_desc_sw128(smem_u32(smem + s * L::STAGE));
        const uint64_t db = make_desc_sw128(smem_u32(smem + s * L::STAGE + GT_A_BYTES));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)      // 32 bytes of K per instruction (8 tf32 / 16 bf16)
          mma_pair_ss<KIND>(tmem_base, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), IDESC, (kb > 0) || (ks > 0));
        mma_commit_g<2>(empty + s);
      }
      mma_commit_g<2>(accum);
    }
  } else {
    mbar_wait(accum, 0);
    tc_fence_after();
    // Every MMA of the pair has completed, so the operand ring is dead: it becomes the transpose buffer of the epilogue.
    // TMEM hands each thread one output ROW (32 columns at a time); going through shared memory lets the warp touch
    // global memory with 128 contiguous bytes per row instead of 32 rows x 16 bytes.
    float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 36);     // [32 rows][36]: conflict-free 16-byte accesses
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float gm = (EPI == GT_EPI_RESIDUAL) ? *p.res_scale : 0.f;
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
      if (n0 + c >= p.N) break;                                         // uniform
      uint32_t r[32];
      tmem_ld32(t_row + c, r);
      // residual / bias operands of this chunk are fetched while the TMEM load is in flight
      float4 x4[8], b4[8];
      if (EPI == GT_EPI_RESIDUAL) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), cc = (lane & 7) * 4;
          const int m = m0 + warp * 32 + row, n = n0 + c + cc;
          if (m < p.M && n < p.N) { x4[it] = ld4(p.res + (size_t)m * p.N + n); b4[it] = ld4(p.bias + n); }
        }
      }
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(sT + lane * 36 + q * 4) =
            make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                        __uint_as_float(r[4 * q + 3]));
      __syncwarp();
      if (EPI == GT_EPI_QKV) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {        // 8 rows x (4 lanes x 8 columns) per pass
          const int row = it * 8 + (lane >> 2), cc = (lane & 3) * 8;
          const int m = m0 + warp * 32 + row, n = n0 + c + cc;
          if (m < p.M && n < p.N) {             // 8-column groups never straddle the q / k / v regions
            const float4 a0 = *reinterpret_cast<const float4*>(sT + row * 36 + cc);
            const float4 a1 = *reinterpret_cast<const float4*>(sT + row * 36 + cc + 4);
            const float4 b0 = ld4(p.bias + n), b1 = ld4(p.bias + n + 4);
            float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
            __nv_bfloat16* dst;
            if (n < p.qk_d) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] *= p.q_scale;
              dst = p.q_out + (size_t)m * 64 + n;
            } else if (n < 2 * p.qk_d) {
              dst = p.k_out + (size_t)m * 64 + (n - p.qk_d);
            } else {
              dst = p.v_out + (size_t)m * p.v_dv + (n - 2 * p.qk_d);
            }
            *reinterpret_cast<uint4*>(dst) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
        }
      } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) {        // 4 rows x (8 lanes x 4 columns) per pass
          const int row = it * 4 + (lane >> 3), cc = (lane & 7) * 4;
          const int m = m0 + warp * 32 + row, n = n0 + c + cc;
          if (m < p.M && n < p.N) {
            const float4 a = *reinterpret_cast<const float4*>(sT + row * 36 + cc);
            st4(p.out + (size_t)m * p.N + n,
                make_float4(fmaf(gm, a.x + b4[it].x, x4[it].x), fmaf(gm, a.y + b4[it].y, x4[it].y),
                            fmaf(gm, a.z + b4[it].z, x4[it].z), fmaf(gm, a.w + b4[it].w, x4[it].w)));
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();                    // the peer may still be reading this CTA's accumulator barrier / operand ring
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_g<2>(tmem_base, L::TMEM_COLS);
  }
}

// 2D fp32 tensor [rows][cols] (cols contiguous), box [box_rows][32 cols = 128 B], SWIZZLE_128B, OOB -> 0
static int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled is not available from this driver");
    return SAGAN_EUNSUPPORTED;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled(fp32) failed with CUresult %d (rows=%llu cols=%llu)", (int)r, (unsigned long long)rows,
            (unsigned long long)cols);
    return SAGAN_EINVAL;
  }
  return 0;
}

template <int KIND, int EPI, int NT>
static int launch_gemm_nt(const void* A, const void* Bt, const GemmP& p, cudaStream_t st) {
  using L = GemmSmem<NT>;
  CUtensorMap ta, tb;
  int rc;
  if (KIND == GT_TF32) {
    if ((rc = make_tmap_f32_2d(&ta, A, (uint64_t)p.M, (uint64_t)p.K, 128))) return rc;
    if ((rc = make_tmap_f32_2d(&tb, Bt, (uint64_t)p.N, (uint64_t)p.K, NT / 2))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)p.K * 2, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tb, Bt, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.K * 2, NT / 2))) return rc;
  }
  auto kern = gemm_pair_kernel<KIND, EPI, NT>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ceil_div(p.N, NT) * ceil_div(p.M, 256));
  cfg.blockDim = dim3(GT_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SAGAN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// column tile: the widest of 128 / 192 / 256 that wastes the fewest padded columns (fewer N tiles = fewer re-reads of A)
template <int KIND, int EPI>
static int launch_gemm(const void* A, const void* Bt, const GemmP& p, cudaStream_t st) {
  const int align = KIND == GT_TF32 ? 4 : 8;
  if ((p.K % align) != 0 || (((uintptr_t)A | (uintptr_t)Bt) & 15) != 0) {
    set_err("gemm_tc: K must be a multiple of %d and operands 16-byte aligned (K=%d)", align, p.K);
    return SAGAN_EUNSUPPORTED;
  }
  int best = 128, waste_best = 1 << 30;
  for (int nt : {256, 192, 128}) {
    const int tiles = ceil_div(p.N, nt), waste = tiles * nt - p.N;
    if (waste < waste_best) { best = nt; waste_best = waste; }
  }
  switch (best) {
    case 256: return launch_gemm_nt<KIND, EPI, 256>(A, Bt, p, st);
    case 192: return launch_gemm_nt<KIND, EPI, 192>(A, Bt, p, st);
    default: return launch_gemm_nt<KIND, EPI, 128>(A, Bt, p, st);
  }
}

// x [M, K] fp32, wt [2d + dv, K] fp32 (transposed concatenation of the three 1x1 kernels), bias [2d + dv]
int gemm_tf32_qkv(const float* x, const float* wt, const float* bcat, __nv_bfloat16* q, __nv_bfloat16* k,
                  __nv_bfloat16* v, long long M, int K, int d, int dv, float q_scale, cudaStream_t st) {
  GemmP p{};
  p.M = (int)M; p.N = 2 * d + dv; p.K = K; p.bias = bcat;
  p.q_out = q; p.k_out = k; p.v_out = v; p.qk_d = d; p.v_dv = dv; p.q_scale = q_scale;
  return launch_gemm<GT_TF32, GT_EPI_QKV>(x, wt, p, st);
}

// y = res + (*res_scale) * (a [M, K] bf16 * wt [N, K]^T bf16 + bias)
int gemm_bf16_residual(const __nv_bfloat16* a, const __nv_bfloat16* wt, const float* bias, const float* res,
                       const float* res_scale, float* y, long long M, int K, int N, cudaStream_t st) {
  GemmP p{};
  p.M = (int)M; p.N = N; p.K = K; p.bias = bias; p.res = res; p.res_scale = res_scale; p.out = y;
  if (N % 4) {
    set_err("gemm_bf16_residual: N must be a multiple of 4 (N=%d)", N);
    return SAGAN_EUNSUPPORTED;
  }
  return launch_gemm<GT_BF16, GT_EPI_RESIDUAL>(a, wt, p, st);
}

// y = res + (*res_scale) * (a [M, K] fp32 * wt [N, K]^T fp32 + bias)   (kind::tf32 on the fp32 operands)
int gemm_tf32_residual(const float* a, const float* wt, const float* bias, const float* res, const float* res_scale,
                       float* y, long long M, int K, int N, cudaStream_t st) {
  GemmP p{};
  p.M = (int)M; p.N = N; p.K = K; p.bias = bias; p.res = res; p.res_scale = res_scale; p.out = y;
  if (N % 4) {
    set_err("gemm_tf32_residual: N must be a multiple of 4 (N=%d)", N);
    return SAGAN_EUNSUPPORTED;
  }
  return launch_gemm<GT_TF32, GT_EPI_RESIDUAL>(a, wt, p, st);
}

}  // namespace sagan
