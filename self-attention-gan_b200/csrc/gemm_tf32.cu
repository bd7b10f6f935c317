// TMA-fed tcgen05 GEMM on fp32 operands (kind::tf32, fp32 accumulate in TMEM) for the 1x1 convolutions of the
// large-C self-attention block (/root/reference/layers.py:82-85,99,104-105,112,119):
//
//   D [M, N] = A [M, K] * Bt [N, K]^T        A = activations (NHWC rows, K contiguous), Bt = transposed Keras kernel
//
// The activations are fp32 in HBM; reading them as TF32 straight through TMA avoids a separate fp32 -> bf16
// conversion pass (which would cost more HBM traffic than the GEMM itself at these shapes) and keeps 10 mantissa bits.
// CTA tile 128 x 128, K blocks of 32 fp32 (one 128-byte swizzle span), 3-stage TMA ring, 6 warps:
//   warps 0-3 epilogue (TMEM lane = output row), warp 4 TMA producer, warp 5 MMA issuer.
// Two CTAs are resident per SM (96 KB of shared memory, 128 TMEM columns each), so one CTA's epilogue runs under the
// other's main loop.  The N tiles of one M tile are adjacent in the grid, so A is re-read from L2, not from HBM.
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
constexpr int GT_A_BYTES = 128 * 128;     // [128 rows][32 fp32]
constexpr int GT_B_BYTES = 128 * 128;
constexpr int GT_STAGE = GT_A_BYTES + GT_B_BYTES;
constexpr int GT_SMEM = GT_STAGES * GT_STAGE + 256 + 1024;

enum { GT_EPI_QKV = 0, GT_EPI_RESIDUAL = 1 };

struct GemmTf32P {
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

// instruction descriptor, kind::tf32: A, B = tf32 (format 2), D = f32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(GT_THREADS, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTf32P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE);
  uint64_t* full = bars;                  // [GT_STAGES] TMA landed
  uint64_t* empty = bars + GT_STAGES;     // [GT_STAGES] MMAs of the stage done
  uint64_t* accum = bars + 2 * GT_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GT_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * 128;
  const int nkb = (p.K + 31) / 32;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(accum, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    if (elect_one_sync()) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        if (kb >= GT_STAGES) mbar_wait(empty + s, ((kb / GT_STAGES) - 1) & 1);
        mbar_expect_tx(full + s, GT_STAGE);
        tma_load_2d(smem + s * GT_STAGE, &tmA, full + s, kb * 32, m0);               // OOB rows / columns are zero-filled
        tma_load_2d(smem + s * GT_STAGE + GT_A_BYTES, &tmB, full + s, kb * 32, n0);
      }
    }
  } else if (warp == 5) {
    if (elect_one_sync()) {
      constexpr uint32_t IDESC = make_idesc_tf32(128, 128);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GT_STAGES;
        mbar_wait(full + s, (kb / GT_STAGES) & 1);
        tc_fence_after();
        const uint64_t da = make_desc_sw128(smem_u32(smem + s * GT_STAGE));
        const uint64_t db = make_desc_sw128(smem_u32(smem + s * GT_STAGE + GT_A_BYTES));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)      // 8 tf32 = 32 bytes of K per instruction
          mma_tf32_ss(tmem_base, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), IDESC, (kb > 0) || (ks > 0));
        mma_commit(empty + s);
      }
      mma_commit(accum);
    }
  } else {
    mbar_wait(accum, 0);
    tc_fence_after();
    // Every MMA has completed, so the operand ring is dead: stage 0 becomes the transpose buffer of the epilogue.
    // TMEM hands each thread one output ROW (32 columns at a time); going through shared memory lets the warp touch
    // global memory with 128 contiguous bytes per row instead of 32 rows x 16 bytes.
    float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 36);     // [32 rows][36]: conflict-free 16-byte accesses
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float gm = (EPI == GT_EPI_RESIDUAL) ? *p.res_scale : 0.f;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      tmem_ld32(t_row + c, r);
      tmem_wait_ld();
      if (n0 + c >= p.N) break;                                         // uniform
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
            const float4 x4 = ld4(p.res + (size_t)m * p.N + n);
            const float4 b4 = ld4(p.bias + n);
            st4(p.out + (size_t)m * p.N + n, make_float4(fmaf(gm, a.x + b4.x, x4.x), fmaf(gm, a.y + b4.y, x4.y),
                                                         fmaf(gm, a.z + b4.z, x4.z), fmaf(gm, a.w + b4.w, x4.w)));
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
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

template <int EPI>
static int launch_gemm_tf32(const float* A, const float* Bt, const GemmTf32P& p, cudaStream_t st) {
  if ((p.K % 4) != 0 || (((uintptr_t)A | (uintptr_t)Bt) & 15) != 0) {
    set_err("gemm_tf32: K must be a multiple of 4 and operands 16-byte aligned (K=%d)", p.K);
    return SAGAN_EUNSUPPORTED;
  }
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_f32_2d(&ta, A, (uint64_t)p.M, (uint64_t)p.K, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&tb, Bt, (uint64_t)p.N, (uint64_t)p.K, 128))) return rc;
  auto kern = gemm_tf32_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    SAGAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
    configured = true;
  }
  kern<<<dim3(ceil_div(p.N, 128), ceil_div(p.M, 128)), GT_THREADS, GT_SMEM, st>>>(ta, tb, p);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

// x [M, K] fp32, wt [2d + dv, K] fp32 (transposed concatenation of the three 1x1 kernels), bias [2d + dv]
int gemm_tf32_qkv(const float* x, const float* wt, const float* bcat, __nv_bfloat16* q, __nv_bfloat16* k,
                  __nv_bfloat16* v, long long M, int K, int d, int dv, float q_scale, cudaStream_t st) {
  GemmTf32P p{};
  p.M = (int)M; p.N = 2 * d + dv; p.K = K; p.bias = bcat;
  p.q_out = q; p.k_out = k; p.v_out = v; p.qk_d = d; p.v_dv = dv; p.q_scale = q_scale;
  return launch_gemm_tf32<GT_EPI_QKV>(x, wt, p, st);
}

// y = res + (*res_scale) * (x [M, K] * wt [N, K]^T + bias)
int gemm_tf32_residual(const float* x, const float* wt, const float* bias, const float* res, const float* res_scale,
                       float* y, long long M, int K, int N, cudaStream_t st) {
  GemmTf32P p{};
  p.M = (int)M; p.N = N; p.K = K; p.bias = bias; p.res = res; p.res_scale = res_scale; p.out = y;
  if (N % 4) {
    set_err("gemm_tf32_residual: N must be a multiple of 4 (N=%d)", N);
    return SAGAN_EUNSUPPORTED;
  }
  return launch_gemm_tf32<GT_EPI_RESIDUAL>(x, wt, p, st);
}

}  // namespace sagan
