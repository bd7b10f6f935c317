// Microbenchmark (diagnostics, not part of the library): the register-only instruction mix of the flash-forward
// softmax inner loop (attn_tc.cu), 16 warps per SM as in attn_fwd_tc4_kernel, no TMEM / barriers / MMA.
// Prints cycles per 32-column chunk and warp for several subsets of the mix, to tell pipe limits from sync limits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/softmax_mix tools/ubench/softmax_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 f2_pack(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f2 x, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
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

// MODE bits: 1 = row maximum (FMNMX3), 2 = MUFU exponentials (else FMUL), 4 = bf16 packing (else XOR),
//            8 = scalar FADD instead of FADD2 for S - m;  POLY = polynomial pairs per 8 pairs (0, 2, 3, 4)
template <int MODE, int POLY>
__global__ void __launch_bounds__(512, 1) mix_kernel(const float* in, uint32_t* out, int iters, long long* cycles) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = in[(threadIdx.x * 32 + i) & 1023];
  uint32_t acc = 0;
  float m = 3.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const f2 neg_m = f2_pack(-m, -m);
    float x[32];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      if (MODE & 8) {
        x[2 * e] = r[2 * e] - m;
        x[2 * e + 1] = r[2 * e + 1] - m;
      } else {
        f2_unpack(f2_add(f2_pack(r[2 * e], r[2 * e + 1]), neg_m), x[2 * e], x[2 * e + 1]);
      }
    }
    if (MODE & 1) {
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, fmaxf(x[i], x[i + 1]));
        mx1 = fmaxf(mx1, fmaxf(x[i + 2], x[i + 3]));
      }
      m += fmaxf(mx0, mx1) * 1e-30f;
    } else {
      m += 1e-3f;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      uint32_t pk;
      const bool poly = (POLY == 2 && (e & 3) == 3) || (POLY == 3 && ((e & 3) == 3 || (e & 7) == 1)) || (POLY == 4 && (e & 1));
      float p0, p1;
      if (poly) {
        pk = ex2_poly2_bf16(f2_pack(x[2 * e], x[2 * e + 1]));
      } else {
        if (MODE & 2) { p0 = ex2_approx(x[2 * e]); p1 = ex2_approx(x[2 * e + 1]); }
        else { p0 = x[2 * e] * 0.5f; p1 = x[2 * e + 1] * 0.25f; }
        pk = (MODE & 4) ? pack_bf16x2(p0, p1) : (__float_as_uint(p0) ^ __float_as_uint(p1));
      }
      acc ^= pk;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int POLY>
static void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
  const int iters = 2000;
  mix_kernel<MODE, POLY><<<148, 512>>>(in, out, 10, cyc);
  cudaDeviceSynchronize();
  mix_kernel<MODE, POLY><<<148, 512>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  // 4 warps per scheduler: cycles per iteration == scheduler cycles per 4 warp-chunks == per 128-key tile of the kernel
  printf("%-58s %8.1f cycles / tile-equivalent (4 warps x 32 columns per scheduler)\n", name, s / 148 / iters);
}

int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = -0.01f * (float)(i % 97);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<7, 2>("full mix: max + FADD2 + MUFU 3/4 + poly 1/4 + F2FP", in, out, cyc);
  run<7, 0>("all MUFU: max + FADD2 + MUFU + F2FP", in, out, cyc);
  run<7, 3>("poly 3/8", in, out, cyc);
  run<7, 4>("poly 1/2", in, out, cyc);
  run<6, 2>("no max", in, out, cyc);
  run<6, 0>("no max, all MUFU", in, out, cyc);
  run<2, 0>("FADD2 + MUFU + XOR (no F2FP, no max)", in, out, cyc);
  run<4, 0>("FADD2 + FMUL + F2FP (no MUFU, no max)", in, out, cyc);
  run<5, 0>("max + FADD2 + FMUL + F2FP (no MUFU)", in, out, cyc);
  run<0, 0>("FADD2 + FMUL + XOR", in, out, cyc);
  run<15, 2>("full mix with scalar FADD", in, out, cyc);
  run<14, 0>("scalar FADD + MUFU + F2FP, no max", in, out, cyc);
  return 0;
}
