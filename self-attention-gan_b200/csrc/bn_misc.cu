// BatchNorm(training)+LeakyReLU (generator.py:10-11), hinge losses (sagan/main.py:21-27),
// Keras Adam (sagan/main.py:119-120) -- the elementwise glue between the conv / attention kernels.
#include "common.cuh"

namespace sagan {

// ------------------------------------------------------------------------------------ BatchNorm
// x [rows, C] NHWC-flattened.  Stage 1: per-CTA partial (sum, sumsq) in fp32 -> ws[nblk][2][C];
// stage 2 (apply): every CTA folds the partials in fp64 in a fixed order (deterministic), then
// y = lrelu((x - mean) * invstd * gamma + beta).
constexpr int BN_THREADS = 256;
constexpr int BN_MAX_BLOCKS = 296;
constexpr int BN_MAX_C = 1024;

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float* __restrict__ x, float* __restrict__ part, long long rows, int C, int rows_per_block) {
  extern __shared__ float sm[];   // [2][BN_THREADS]
  const int cpt = min(C, BN_THREADS);
  const int rl = BN_THREADS / cpt;
  const int c_lane = threadIdx.x % cpt, r_lane = threadIdx.x / cpt;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += cpt) {
    const int c = c0 + c_lane;
    float s = 0.f, q = 0.f;
    if (c < C && r_lane < rl)
      for (long long r = r0 + r_lane; r < r1; r += rl) {
        const float v = x[r * C + c];
        s += v;
        q = fmaf(v, v, q);
      }
    sm[threadIdx.x] = s;
    sm[BN_THREADS + threadIdx.x] = q;
    __syncthreads();
    if (r_lane == 0 && c < C) {
      for (int j = 1; j < rl; ++j) {
        s += sm[j * cpt + c_lane];
        q += sm[BN_THREADS + j * cpt + c_lane];
      }
      part[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
      part[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
    }
    __syncthreads();
  }
}

// Folds part[nparts][2][C] into (sum, sumsq) per channel in fp64, in a FIXED order (deterministic), using the whole
// CTA: thread t takes channel t % cpt and every (BN_THREADS / cpt)-th partial, then the slices are folded in order.
// (A single thread per channel walking all partials serialised ~600 dependent L2 loads in EVERY CTA: 40-60 us.)
__device__ __forceinline__ void bn_fold_partials(const float* __restrict__ part, int nparts, int C, int c0, double* red,
                                                 double& s_out, double& q_out, bool& owner, int& c_out) {
  const int cpt = min(C, BN_THREADS);
  const int sl = BN_THREADS / cpt;
  const int c_lane = threadIdx.x % cpt, slice = threadIdx.x / cpt;
  const int c = c0 + c_lane;
  double s = 0.0, q = 0.0;
  if (c < C && slice < sl)
    for (int p = slice; p < nparts; p += sl) {
      s += (double)part[((size_t)p * 2 + 0) * C + c];
      q += (double)part[((size_t)p * 2 + 1) * C + c];
    }
  red[threadIdx.x] = s;
  red[BN_THREADS + threadIdx.x] = q;
  __syncthreads();
  owner = (slice == 0 && c < C);
  if (owner)
    for (int j = 1; j < sl; ++j) {
      s += red[j * cpt + c_lane];
      q += red[BN_THREADS + j * cpt + c_lane];
    }
  __syncthreads();
  s_out = s; q_out = q; c_out = c;
}

__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ y, const float* __restrict__ part, int nparts, float* __restrict__ save_mean,
                float* __restrict__ save_invstd, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                long long rows, int C, float eps, float momentum, float slope) {
  extern __shared__ float sm[];   // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + C;
  __shared__ double red[2 * BN_THREADS];
  for (int c0 = 0; c0 < C; c0 += min(C, BN_THREADS)) {
    double s, q;
    bool owner;
    int c;
    bn_fold_partials(part, nparts, C, c0, red, s, q, owner, c);
    if (!owner) continue;
    const double mean = s / (double)rows;
    double var = q / (double)rows - mean * mean;   // biased variance (Keras training mode)
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (blockIdx.x == 0) {
      save_mean[c] = (float)mean;
      save_invstd[c] = invstd;
      if (moving_mean) moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)mean;
      if (moving_var) moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)var;
    }
  }
  __syncthreads();
  const long long n = rows * C;
  if (C & 3) {   // channel counts that are not a multiple of 4 (narrow test models): one element per thread
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
      const int c = (int)(i % C);
      const float o = fmaf(x[i], scale[c], shift[c]);
      y[i] = o > 0.f ? o : o * slope;
    }
    return;
  }
  const long long stride = (long long)gridDim.x * BN_THREADS * 4;
  for (long long i = ((long long)blockIdx.x * BN_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const int c = (int)(i % C);   // C % 4 == 0 here
    const float4 v = ld4(x + i);
    float4 o;
    o.x = fmaf(v.x, scale[c + 0], shift[c + 0]); o.y = fmaf(v.y, scale[c + 1], shift[c + 1]);
    o.z = fmaf(v.z, scale[c + 2], shift[c + 2]); o.w = fmaf(v.w, scale[c + 3], shift[c + 3]);
    o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
    o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
    st4(y + i, o);
  }
}

// Inference-mode BatchNormalization + LeakyReLU (generator(..., training=False): the sample dumps of
// sagan/main.py:333): y = lrelu((x - moving_mean) * rsqrt(moving_var + eps) * gamma + beta).  One pass, no statistics.
__global__ void __launch_bounds__(BN_THREADS)
bn_infer_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ moving_mean, const float* __restrict__ moving_var, float* __restrict__ y,
                long long rows, int C, float eps, float slope) {
  extern __shared__ float sm[];   // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + C;
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    const float sc = gamma[c] * (1.0f / sqrtf(moving_var[c] + eps));
    scale[c] = sc;
    shift[c] = beta[c] - moving_mean[c] * sc;
  }
  __syncthreads();
  const long long n = rows * C;
  if (C & 3) {
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
      const int c = (int)(i % C);
      const float o = fmaf(x[i], scale[c], shift[c]);
      y[i] = o > 0.f ? o : o * slope;
    }
    return;
  }
  const long long stride = (long long)gridDim.x * BN_THREADS * 4;
  for (long long i = ((long long)blockIdx.x * BN_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const int c = (int)(i % C);
    const float4 v = ld4(x + i);
    float4 o;
    o.x = fmaf(v.x, scale[c + 0], shift[c + 0]); o.y = fmaf(v.y, scale[c + 1], shift[c + 1]);
    o.z = fmaf(v.z, scale[c + 2], shift[c + 2]); o.w = fmaf(v.w, scale[c + 3], shift[c + 3]);
    o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
    o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
    st4(y + i, o);
  }
}

// backward stage 1: partial dbeta = sum dz, dgamma = sum dz * xhat, with dz = dy * lrelu'(y)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                    const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ part,
                    long long rows, int C, int rows_per_block, float slope) {
  extern __shared__ float sm[];
  const int cpt = min(C, BN_THREADS);
  const int rl = BN_THREADS / cpt;
  const int c_lane = threadIdx.x % cpt, r_lane = threadIdx.x / cpt;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += cpt) {
    const int c = c0 + c_lane;
    float s = 0.f, q = 0.f;
    if (c < C && r_lane < rl) {
      const float mu = mean[c], is = invstd[c];
      for (long long r = r0 + r_lane; r < r1; r += rl) {
        const float yy = y[r * C + c];
        float dz = dy[r * C + c];
        dz = yy > 0.f ? dz : dz * slope;
        s += dz;
        q = fmaf(dz, (x[r * C + c] - mu) * is, q);
      }
    }
    sm[threadIdx.x] = s;
    sm[BN_THREADS + threadIdx.x] = q;
    __syncthreads();
    if (r_lane == 0 && c < C) {
      for (int j = 1; j < rl; ++j) {
        s += sm[j * cpt + c_lane];
        q += sm[BN_THREADS + j * cpt + c_lane];
      }
      part[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
      part[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
    }
    __syncthreads();
  }
}

// backward stage 2: dx = gamma * invstd * (dz - dbeta/M - xhat * dgamma/M)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ part, int nparts, float* __restrict__ dx, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, long long rows, int C, float slope) {
  extern __shared__ float sm[];   // a[C], b[C], c0[C], mu[C], is[C]
  float* ka = sm;            // gamma * invstd
  float* kb = sm + C;        // dbeta / M
  float* kc = sm + 2 * C;    // dgamma / M
  float* kmu = sm + 3 * C;
  float* kis = sm + 4 * C;
  __shared__ double red[2 * BN_THREADS];
  for (int c0 = 0; c0 < C; c0 += min(C, BN_THREADS)) {
    double s, q;
    bool owner;
    int c;
    bn_fold_partials(part, nparts, C, c0, red, s, q, owner, c);
    if (!owner) continue;
    if (blockIdx.x == 0) {
      dbeta[c] = (float)s;
      dgamma[c] = (float)q;
    }
    ka[c] = gamma[c] * invstd[c];
    kb[c] = (float)(s / (double)rows);
    kc[c] = (float)(q / (double)rows);
    kmu[c] = mean[c];
    kis[c] = invstd[c];
  }
  __syncthreads();
  const long long n = rows * C;
  if (C & 3) {   // scalar form, see bn_apply_kernel
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
      const int c = (int)(i % C);
      const float dz = y[i] > 0.f ? dy[i] : dy[i] * slope;
      const float xh = (x[i] - kmu[c]) * kis[c];
      dx[i] = ka[c] * (dz - kb[c] - xh * kc[c]);
    }
    return;
  }
  const long long stride = (long long)gridDim.x * BN_THREADS * 4;
  for (long long i = ((long long)blockIdx.x * BN_THREADS + threadIdx.x) * 4; i < n; i += stride) {
    const int c = (int)(i % C);
    const float4 d = ld4(dy + i), xx = ld4(x + i), yy = ld4(y + i);
    const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {xx.x, xx.y, xx.z, xx.w}, yv[4] = {yy.x, yy.y, yy.z, yy.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dz = yv[j] > 0.f ? dv[j] : dv[j] * slope;
      const float xh = (xv[j] - kmu[c + j]) * kis[c + j];
      o[j] = ka[c + j] * (dz - kb[c + j] - xh * kc[c + j]);
    }
    st4(dx + i, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------ hinge
__global__ void __launch_bounds__(256)
hinge_d_kernel(const float* __restrict__ dr, const float* __restrict__ df, long long n, float scale,
               float* __restrict__ loss_sum, float* __restrict__ gr, float* __restrict__ gf) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = 1.f - dr[i], b = 1.f + df[i];       // main.py:25-26
    acc += fmaxf(a, 0.f) + fmaxf(b, 0.f);
    gr[i] = a > 0.f ? -scale : 0.f;
    gf[i] = b > 0.f ? scale : 0.f;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_sum, acc);
}

__global__ void __launch_bounds__(256)
hinge_g_kernel(const float* __restrict__ df, long long n, float scale, float* __restrict__ loss_sum,
               float* __restrict__ gf) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    acc += -df[i];                                       // main.py:22
    gf[i] = -scale;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_sum, acc);
}

// ------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, const float* __restrict__ hyper, float grad_scale) {
  const float lr_t = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gg = g[i] * grad_scale;
    float mm = gg;
    if (m) {
      mm = b1 * m[i] + (1.f - b1) * gg;
      m[i] = mm;
    }
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    v[i] = vv;
    p[i] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}

// Keras ExponentialDecay(staircase) + Adam bias correction evaluated ON THE DEVICE from a device-resident step counter
// (sagan/main.py:111-120): one thread, double precision.  Runs right before each Adam application inside the step
// (graph), so a replayed CUDA graph follows the schedule without any per-step host -> device traffic.
__global__ void adam_schedule_kernel(float* __restrict__ hyper, long long* __restrict__ iterations, double lr0,
                                     double decay_rate, long long decay_steps, double b1, double b2, double eps) {
  const long long it = *iterations;
  const double lr = lr0 * pow(decay_rate, (double)(it / decay_steps));        // staircase=True
  const double t = (double)(it + 1);
  const double corr = sqrt(1.0 - pow(b2, t)) / (1.0 - (b1 > 0.0 ? pow(b1, t) : 0.0));
  hyper[0] = (float)(lr * corr);
  hyper[1] = (float)b1;
  hyper[2] = (float)b2;
  hyper[3] = (float)eps;
  *iterations = it + 1;                                                        // optimizer.iterations
}

// ------------------------------------------------------------------------------------ multi-tensor accumulate
// dst[i] += src[i] for up to 64 small tensors in ONE launch: the gradients of the parameters that are not spectrally
// normalised (biases, BatchNorm gamma / beta, attention gamma, head kernels) go into the network's flat gradient bucket
// this way instead of one autograd accumulation kernel per parameter (42 launches per step at church64).
constexpr int ACC_MAX = 64;
struct AccTab {
  float* dst[ACC_MAX];
  const float* src[ACC_MAX];
  int n[ACC_MAX];
};
__global__ void __launch_bounds__(256) accumulate_multi_kernel(const AccTab t) {
  float* d = t.dst[blockIdx.y];
  const float* s = t.src[blockIdx.y];
  const int n = t.n[blockIdx.y];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) d[i] += s[i];
}

// ------------------------------------------------------------------------------------ weight normalisation
// /root/reference/sagan/layers.py:75-135 (the TF-Addons WeightNormalization wrapper the `sagan/` tree wraps its layers
// with): kernel = l2_normalize(v, all axes but the last) * g.  v is the [rows, cols] row-major view of the Keras kernel
// (cols = the last, filter axis).  Two phases: per-column partial sums over row ranges (coalesced along the columns,
// fp32 atomics into a zeroed [cols] vector), then one elementwise pass.
constexpr int WN_COLS = 32, WN_ROWS_PER_CTA = 256;

template <bool DOT>   // DOT = false: sum v^2 ; DOT = true: sum a * b
__global__ void __launch_bounds__(256)
wn_colreduce_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int rows, int cols) {
  __shared__ float red[8][WN_COLS];
  const int c = blockIdx.x * WN_COLS + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * WN_ROWS_PER_CTA, r1 = min(rows, r0 + WN_ROWS_PER_CTA);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + rl; r < r1; r += 8) {
      const float x = a[(size_t)r * cols + c];
      acc = fmaf(x, DOT ? b[(size_t)r * cols + c] : x, acc);
    }
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// w = v * rsqrt(max(sumsq, 1e-12)) * g   (tf.nn.l2_normalize: epsilon = 1e-12 under the square root); inv_norm <- rsqrt(..)
__global__ void __launch_bounds__(256)
wn_apply_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ sumsq_to_inv,
                float* __restrict__ w, long long n, int cols, const float* __restrict__ sumsq_ro) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % cols);
    const float inv = rsqrtf(fmaxf(sumsq_ro[c], 1e-12f));
    w[i] = v[i] * inv * g[c];
  }
  (void)sumsq_to_inv;
}

__global__ void wn_inv_kernel(float* __restrict__ sumsq, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) sumsq[c] = rsqrtf(fmaxf(sumsq[c], 1e-12f));
}

// dv = g inv (dw - v dot inv^2), dg = dot inv     with dot[c] = sum_r dw[r,c] v[r,c]
__global__ void __launch_bounds__(256)
wn_bwd_apply_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                    const float* __restrict__ inv_norm, const float* __restrict__ dot, float* __restrict__ dv,
                    float* __restrict__ dg, long long n, int cols) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % cols);
    const float inv = inv_norm[c];
    dv[i] = g[c] * inv * (dw[i] - v[i] * dot[c] * inv * inv);
    if (i < cols) dg[i] = dot[i] * inv_norm[i];
  }
}

// ------------------------------------------------------------------------------------ uint8 records -> [-1, 1] floats
// sagan/dataset.py:31-34: image = cast(decode_raw(uint8), float32) * (2. / 255) - 1.  (two separately rounded fp32 ops,
// as the un-fused TF graph computes them: no FMA contraction, so the result is bit-identical to numpy float32)
__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n, float scale, float shift) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const uchar4 u = *reinterpret_cast<const uchar4*>(src + i);
      st4(dst + i, make_float4(__fadd_rn(__fmul_rn((float)u.x, scale), shift), __fadd_rn(__fmul_rn((float)u.y, scale), shift),
                               __fadd_rn(__fmul_rn((float)u.z, scale), shift), __fadd_rn(__fmul_rn((float)u.w, scale), shift)));
    } else {
      for (long long j = i; j < n; ++j) dst[j] = __fadd_rn(__fmul_rn((float)src[j], scale), shift);
    }
  }
}

static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace sagan

using namespace sagan;

extern "C" size_t sagan_bn_workspace_bytes(int C) { return (size_t)BN_MAX_BLOCKS * 2 * (size_t)(C > 0 ? C : 1) * sizeof(float); }

static int bn_blocks(long long rows, int* rows_per_block) {
  int nblk = (int)std::min<long long>(BN_MAX_BLOCKS, std::max<long long>(1, rows / 64));
  *rows_per_block = (int)ceil_div<long long>(rows, nblk);
  return (int)ceil_div<long long>(rows, *rows_per_block);
}

extern "C" int sagan_bn_lrelu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean,
                                  float* save_invstd, float* moving_mean, float* moving_var, long long rows, int C,
                                  float eps, float momentum, float slope, void* ws, size_t ws_bytes,
                                  sagan_stream_t stream) {
  SAGAN_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && ws, "sagan_bn_lrelu_fwd: null pointer");
  SAGAN_REQUIRE(rows > 0 && C > 0 && C <= BN_MAX_C, "sagan_bn_lrelu_fwd: need rows>0, 0<C<=%d (C=%d)", BN_MAX_C, C);
  SAGAN_REQUIRE(al16(x) && al16(y), "sagan_bn_lrelu_fwd: x/y must be 16-byte aligned");
  if (ws_bytes < sagan_bn_workspace_bytes(C)) {
    set_err("sagan_bn_lrelu_fwd: workspace %zu < %zu", ws_bytes, sagan_bn_workspace_bytes(C));
    return SAGAN_EWORKSPACE;
  }
  int rpb;
  const int nblk = bn_blocks(rows, &rpb);
  cudaStream_t st = (cudaStream_t)stream;
  bn_stats_kernel<<<nblk, BN_THREADS, 2 * BN_THREADS * sizeof(float), st>>>(x, (float*)ws, rows, C, rpb);
  SAGAN_LAUNCH_CHECK();
  const int ablk = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(rows * C, BN_THREADS * 4));
  bn_apply_kernel<<<ablk, BN_THREADS, 2 * C * sizeof(float), st>>>(x, gamma, beta, y, (const float*)ws, nblk, save_mean,
                                                                   save_invstd, moving_mean, moving_var, rows, C, eps,
                                                                   momentum, slope);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_bn_lrelu_infer(const float* x, const float* gamma, const float* beta, const float* moving_mean,
                                    const float* moving_var, float* y, long long rows, int C, float eps, float slope,
                                    sagan_stream_t stream) {
  SAGAN_REQUIRE(x && gamma && beta && moving_mean && moving_var && y, "sagan_bn_lrelu_infer: null pointer");
  SAGAN_REQUIRE(rows > 0 && C > 0 && C <= BN_MAX_C, "sagan_bn_lrelu_infer: need rows>0, 0<C<=%d (C=%d)", BN_MAX_C, C);
  SAGAN_REQUIRE(al16(x) && al16(y), "sagan_bn_lrelu_infer: x/y must be 16-byte aligned");
  const int blk = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(rows * C, BN_THREADS * 4));
  bn_infer_kernel<<<blk, BN_THREADS, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, gamma, beta, moving_mean, moving_var,
                                                                                     y, rows, C, eps, slope);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_bn_lrelu_bwd(const float* dy, const float* x, const float* y, const float* gamma,
                                  const float* save_mean, const float* save_invstd, float* dx, float* dgamma,
                                  float* dbeta, long long rows, int C, float slope, void* ws, size_t ws_bytes,
                                  sagan_stream_t stream) {
  SAGAN_REQUIRE(dy && x && y && gamma && save_mean && save_invstd && dx && dgamma && dbeta && ws,
                "sagan_bn_lrelu_bwd: null pointer");
  SAGAN_REQUIRE(rows > 0 && C > 0 && C <= BN_MAX_C, "sagan_bn_lrelu_bwd: need rows>0, 0<C<=%d (C=%d)", BN_MAX_C, C);
  SAGAN_REQUIRE(al16(dy) && al16(x) && al16(y) && al16(dx), "sagan_bn_lrelu_bwd: tensors must be 16-byte aligned");
  if (ws_bytes < sagan_bn_workspace_bytes(C)) {
    set_err("sagan_bn_lrelu_bwd: workspace %zu < %zu", ws_bytes, sagan_bn_workspace_bytes(C));
    return SAGAN_EWORKSPACE;
  }
  int rpb;
  const int nblk = bn_blocks(rows, &rpb);
  cudaStream_t st = (cudaStream_t)stream;
  bn_bwd_stats_kernel<<<nblk, BN_THREADS, 2 * BN_THREADS * sizeof(float), st>>>(dy, x, y, save_mean, save_invstd,
                                                                                 (float*)ws, rows, C, rpb, slope);
  SAGAN_LAUNCH_CHECK();
  const int ablk = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(rows * C, BN_THREADS * 4));
  bn_bwd_apply_kernel<<<ablk, BN_THREADS, 5 * C * sizeof(float), st>>>(dy, x, y, gamma, save_mean, save_invstd,
                                                                       (const float*)ws, nblk, dx, dgamma, dbeta, rows, C,
                                                                       slope);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_hinge_d(const float* d_real, const float* d_fake, long long n, float scale, float* loss_sum,
                             float* g_real, float* g_fake, sagan_stream_t stream) {
  SAGAN_REQUIRE(d_real && d_fake && loss_sum && g_real && g_fake && n > 0, "sagan_hinge_d: bad argument");
  const int blocks = (int)std::min<long long>(num_sms(), ceil_div<long long>(n, 256));
  hinge_d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_real, d_fake, n, scale, loss_sum, g_real, g_fake);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_hinge_g(const float* d_fake, long long n, float scale, float* loss_sum, float* g_fake,
                             sagan_stream_t stream) {
  SAGAN_REQUIRE(d_fake && loss_sum && g_fake && n > 0, "sagan_hinge_g: bad argument");
  const int blocks = (int)std::min<long long>(num_sms(), ceil_div<long long>(n, 256));
  hinge_g_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_fake, n, scale, loss_sum, g_fake);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_adam_step(float* param, const float* grad, float* m, float* v, long long n, const float* hyper,
                               float grad_scale, sagan_stream_t stream) {
  SAGAN_REQUIRE(param && grad && v && hyper && n > 0, "sagan_adam_step: bad argument");
  const int blocks = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(n, 256));
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, hyper, grad_scale);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_adam_schedule(float* hyper, long long* iterations, double lr0, double decay_rate,
                                   long long decay_steps, double b1, double b2, double eps, sagan_stream_t stream) {
  SAGAN_REQUIRE(hyper && iterations && decay_steps > 0 && lr0 > 0 && b2 > 0 && b2 < 1 && b1 >= 0 && b1 < 1,
                "sagan_adam_schedule: bad argument");
  adam_schedule_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper, iterations, lr0, decay_rate, decay_steps, b1, b2, eps);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_wn_fwd(const float* v, const float* g, float* w, float* inv_norm, int rows, int cols,
                            sagan_stream_t stream) {
  SAGAN_REQUIRE(v && g && w && inv_norm && rows > 0 && cols > 0, "sagan_wn_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  SAGAN_CUDA(cudaMemsetAsync(inv_norm, 0, sizeof(float) * cols, st));
  wn_colreduce_kernel<false><<<dim3(ceil_div(cols, WN_COLS), ceil_div(rows, WN_ROWS_PER_CTA)), 256, 0, st>>>(v, nullptr, inv_norm, rows, cols);
  SAGAN_LAUNCH_CHECK();
  const long long n = (long long)rows * cols;
  const int blocks = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(n, 256));
  wn_apply_kernel<<<blocks, 256, 0, st>>>(v, g, nullptr, w, n, cols, inv_norm);
  SAGAN_LAUNCH_CHECK();
  wn_inv_kernel<<<ceil_div(cols, 256), 256, 0, st>>>(inv_norm, cols);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_wn_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg,
                            float* dot_ws, int rows, int cols, sagan_stream_t stream) {
  SAGAN_REQUIRE(dw && v && g && inv_norm && dv && dg && dot_ws && rows > 0 && cols > 0, "sagan_wn_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  SAGAN_CUDA(cudaMemsetAsync(dot_ws, 0, sizeof(float) * cols, st));
  wn_colreduce_kernel<true><<<dim3(ceil_div(cols, WN_COLS), ceil_div(rows, WN_ROWS_PER_CTA)), 256, 0, st>>>(dw, v, dot_ws, rows, cols);
  SAGAN_LAUNCH_CHECK();
  const long long n = (long long)rows * cols;
  const int blocks = (int)std::min<long long>(num_sms() * 4, ceil_div<long long>(n, 256));
  wn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(dw, v, g, inv_norm, dot_ws, dv, dg, n, cols);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_u8_to_f32(const uint8_t* src, float* dst, long long n, float scale, float shift,
                               sagan_stream_t stream) {
  SAGAN_REQUIRE(src && dst && n > 0, "sagan_u8_to_f32: bad argument");
  SAGAN_REQUIRE((((uintptr_t)src) & 3) == 0 && al16(dst), "sagan_u8_to_f32: src must be 4-byte, dst 16-byte aligned");
  const int blocks = (int)std::min<long long>(num_sms() * 8, ceil_div<long long>(n, 1024));
  u8_to_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n, scale, shift);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_accumulate_multi(const sagan_acc_desc* descs, int n, sagan_stream_t stream) {
  SAGAN_REQUIRE(descs && n >= 1 && n <= ACC_MAX, "sagan_accumulate_multi: n=%d outside [1,%d]", n, ACC_MAX);
  AccTab t{};
  long long biggest = 0;
  for (int i = 0; i < n; ++i) {
    SAGAN_REQUIRE(descs[i].dst && descs[i].src && descs[i].n > 0 && descs[i].n < (1ll << 31), "sagan_accumulate_multi: bad descriptor %d", i);
    t.dst[i] = descs[i].dst; t.src[i] = descs[i].src; t.n[i] = (int)descs[i].n;
    biggest = std::max(biggest, descs[i].n);
  }
  const int bx = (int)std::max<long long>(1, std::min<long long>(32, ceil_div<long long>(biggest, 1024)));
  accumulate_multi_kernel<<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>(t);
  SAGAN_LAUNCH_CHECK();
  return 0;
}
