// Shared host/device helpers for libsagan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/sagan_b200.h"

namespace sagan {

// thread-local error text behind sagan_last_error()
char* err_buf();
void set_err(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
extern std::atomic<int> g_deterministic_forward;   // sagan_deterministic_forward(): no atomically-accumulated activations

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define SAGAN_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::sagan::set_err(__VA_ARGS__);      \
      return SAGAN_EINVAL;                \
    }                                     \
  } while (0)

#define SAGAN_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::sagan::set_err("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                 \
    }                                                                                 \
  } while (0)

// after a <<<>>> launch: pick up launch-configuration errors without synchronising
#define SAGAN_LAUNCH_CHECK()                                                          \
  do {                                                                                \
    ::sagan::count_launch();                                                          \
    cudaError_t _e = cudaPeekAtLastError();                                           \
    if (_e != cudaSuccess) {                                                          \
      ::sagan::set_err("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                 \
    }                                                                                 \
  } while (0)

inline int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// convolution geometry shared by the CUDA-core and tensor-core conv kernels
struct CG {
  int B, H, W, Cin, Ho, Wo, Cout, KH, KW, S, PT, PL;
  int M;  // B*Ho*Wo
  int K;  // KH*KW*Cin
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` = 32 floats of shared memory; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace sagan
