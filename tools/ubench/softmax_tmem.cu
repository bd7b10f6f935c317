// Microbenchmark (diagnostics, not part of the library): the per-tile body of attn_fwd_tc4_kernel's softmax warps with
// its TMEM loads / stores and barrier operations switched on one at a time (no MMA: the barriers are always complete).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I self-attention-gan_b200/csrc -o tools/ubench/softmax_tmem tools/ubench/softmax_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace sagan::tc;

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

__device__ __forceinline__ void mbar_wait_timeout(uint64_t* bar) {   // one try_wait with the suspend-time hint used by mbar_wait
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t}"
      ::"r"(smem_u32(bar)), "r"(0u), "r"(0x989680u)
      : "memory");
}

// MODE bits: 1 LDTM of the next chunk, 2 STTM of P', 4 wait::st + fence + arrive right after the store,
//            8 mbarrier wait (already complete) before the LDTM, 16 rescale check (vote + branch), 32 row maximum,
//            64 exponentials (else XOR packing)
//            128 three more warps (17-19 of the CTA) spin on an mbarrier that never completes, as waiting issuers do
template <int MODE>
__global__ void __launch_bounds__(608, 1) tile_kernel(uint32_t* out, int iters, long long* cycles) {
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  __shared__ uint64_t tok[4][4];     // [scheduler][chain]: "the MUFU burst before yours on this scheduler is issued"
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) stop = 0;
  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1);            // never arrived on: waiting for parity 1 succeeds at once
    mbar_init(bars + 1, (1 << 20) - 1);
    mbar_init(bars + 2, 1);
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) mbar_init(&tok[a][b], 32);      // sink for the arrivals
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp >= 16) {
    if (MODE & 128) {
      while (!stop) {
        if (MODE & 256) mbar_wait_timeout(bars + 2);
        else mbar_try_wait(bars + 2, 0);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (false) out[0] = 0;
    return;
  }
  const int q = warp >> 2;
  const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t ra[32], rb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { ra[i] = __float_as_uint(-0.01f * (float)((threadIdx.x + i) % 97)); rb[i] = ra[i] ^ 0x100u; }
  if (MODE & 1) {      // defined TMEM contents
    for (int c = 0; c < 3; ++c) tmem_st32(t_row + c * 128 + q * 32, ra);
    tmem_wait_st();
  }
  float m_used = 3.0f;
  uint32_t acc = 0;
  auto load_next = [&](int t, uint32_t (&r)[32]) {
    if (MODE & 8) { mbar_wait(bars + 0, 1); tc_fence_after(); }
    if (MODE & 1) tmem_ld32(t_row + (uint32_t)((t % 3) * 128 + q * 32), r);
  };
  auto landed = [&](uint32_t (&r)[32]) {
    if (MODE & 1) {
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(r[i]));
    }
  };
  auto row_max = [&](uint32_t (&r)[32]) -> float {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
      mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
    }
    return fmaxf(mx0, mx1);
  };
  auto tile = [&](int j, uint32_t (&cur)[32], uint32_t (&nxt)[32], float& mx) {
    if (MODE & 16) {
      const bool need = mx > m_used + 32.0f;
      if (__any_sync(0xffffffffu, need)) {
        if (need) m_used = ceilf(mx);
      }
    }
    const f2 neg_m = f2_pack(-m_used, -m_used);
    uint32_t pk[16];
    auto exps = [&](int e) {
      const f2 x = f2_add(f2_pack(__uint_as_float(cur[2 * e]), __uint_as_float(cur[2 * e + 1])), neg_m);
      if (!(MODE & 64)) {
        float x0, x1;
        f2_unpack(x, x0, x1);
        pk[e] = __float_as_uint(x0) ^ __float_as_uint(x1);
      } else if ((e & 3) == 3) {
        pk[e] = ex2_poly2_bf16(x);
      } else {
        float x0, x1;
        f2_unpack(x, x0, x1);
        pk[e] = pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
      }
    };
    load_next(j + 1, nxt);
    if (MODE & 2048) {
      // token ring: the four warps of a scheduler take turns issuing their 12-MUFU bursts
      const int sm = warp & 3;
      // burst index of this warp: b = 2 j (+1); ring position: chain q waits for chain q-1's same burst (chain 0 for chain 3's previous)
      auto burst = [&](int b, int e0) {
        f2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = f2_add(f2_pack(__uint_as_float(cur[2 * (e0 + e)]), __uint_as_float(cur[2 * (e0 + e) + 1])), neg_m);
        // wait for the token: chain 0 waits for chain 3's burst b-1, chain q for chain q-1's burst b
        if (q == 0) { if (b > 0) mbar_wait(&tok[sm][3], (b - 1) & 1); }
        else mbar_wait(&tok[sm][q - 1], b & 1);
        float y[16];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (((e0 + e) & 3) == 3) continue;
          float x0, x1;
          f2_unpack(x[e], x0, x1);
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[2 * e]) : "f"(x0));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[2 * e + 1]) : "f"(x1));
        }
        mbar_arrive(&tok[sm][q]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (((e0 + e) & 3) == 3) pk[e0 + e] = ex2_poly2_bf16(x[e]);
          else pk[e0 + e] = pack_bf16x2(y[2 * e], y[2 * e + 1]);
        }
      };
      burst(2 * j, 0);
      landed(nxt);
      if (MODE & 32) mx = row_max(nxt);
      burst(2 * j + 1, 8);
    } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) exps(e);
    landed(nxt);
    if (MODE & 32) mx = row_max(nxt);
    else cur[0] ^= nxt[0] & 1u;
#pragma unroll
    for (int e = 8; e < 16; ++e) exps(e);
    }
    if (MODE & 2) {
      tmem_st16(t_row + (uint32_t)((j % 3) * 128 + q * 32), pk);
      if (MODE & 4) {
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bars + 1);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc ^= pk[e];
    }
    if (!(MODE & 1)) {      // without the TMEM load the next chunk is the current one, perturbed
#pragma unroll
      for (int i = 0; i < 32; i += 8) nxt[i] = cur[i] + (acc & 1u);
#pragma unroll
      for (int i = 0; i < 32; ++i) if (i & 7) nxt[i] = cur[i];
    }
  };
  float mx = 0.f;
  asm volatile("bar.sync 1, 512;" ::: "memory");
  const long long t0 = clock64();
  if (MODE & 512) {       // chains start a quarter of a tile apart
    const long long until = t0 + q * 240;
    while (clock64() < until) { }
  }
  for (int j = 0; j < iters; j += 2) {
    tile(j, ra, rb, mx);
    tile(j + 1, rb, ra, mx);
  }
  if (MODE & 2) tmem_wait_st();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ ra[3] ^ __float_as_uint(mx);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  stop = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int MODE>
static void run(const char* name, uint32_t* out, long long* cyc) {
  const int iters = 2000;
  tile_kernel<MODE><<<148, 608>>>(out, 20, cyc);
  cudaDeviceSynchronize();
  cudaMemset(cyc, 0, 148 * 8);
  tile_kernel<MODE><<<148, 608>>>(out, iters, cyc);
  cudaError_t e = cudaGetLastError();
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = e2;
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-72s %8.1f cycles / tile  (%s)\n", name, s / 148 / iters, cudaGetErrorString(e));
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 640 * 4); cudaMalloc(&cyc, 148 * 8);
  run<64 + 32>("math only (max + exps)", out, cyc);
  run<64 + 32 + 16>("+ rescale check", out, cyc);
  run<64 + 32 + 16 + 1>("+ LDTM next chunk", out, cyc);
  run<64 + 32 + 16 + 1 + 8>("+ mbarrier wait before the LDTM", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2>("+ STTM of P' (no wait)", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2 + 4>("+ wait::st, fence, arrive  (= the kernel's tile body)", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2 + 4 + 128>("tile body + 3 warps spinning on try_wait (no hint)", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2 + 4 + 128 + 256>("tile body + 3 warps spinning on try_wait with the suspend hint", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2 + 4 + 512>("tile body, chains started a quarter tile apart", out, cyc);
  run<64 + 32 + 16 + 1 + 8 + 2 + 4 + 2048>("tile body, MUFU bursts passed round the scheduler's four warps", out, cyc);
  run<64 + 32 + 16 + 2 + 4>("math + check + STTM + wait::st + arrive (no LDTM)", out, cyc);
  run<1 + 8 + 2 + 4>("skeleton: LDTM + wait + STTM + arrive, XOR packing, no max", out, cyc);
  run<1>("LDTM only + XOR", out, cyc);
  run<2 + 4>("STTM + wait::st + arrive only + XOR", out, cyc);
  return 0;
}
