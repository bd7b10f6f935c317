/*
 * sagan_b200.h -- C ABI of libsagan_b200.so: the SAGAN generator/discriminator hot path
 * (self-attention block + spectrally-normalised conv / deconv / dense layers + the
 * elementwise glue of one training step) as hand-written sm_100a CUDA.
 *
 * The reference (jimmYA-1995/Self-Attention-GAN) has no FFI / plugin registry: its
 * boundary for this path is the Keras Layer protocol (`SpectralNormalization(layer)(x)`,
 * `AttentionLayer()(x)`) and every FLOP runs inside TensorFlow library kernels.  Each
 * entry point below cites the reference code whose arithmetic it replaces; the
 * reference-side binding (ctypes from `layers.py`, or a TF REGISTER_OP shim) is shown
 * in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / TF types.
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - activations are NHWC fp32, kernels use the Keras layouts
 *       Conv2D [kh,kw,cin,cout], Conv2DTranspose [kh,kw,cout,cin], Dense [in,out].
 *   - every call is asynchronous on the caller's stream (`stream` = cudaStream_t), never
 *     allocates or synchronises, and is re-entrant across streams and devices.  The caller
 *     owns all buffers, including workspaces sized by the `*_workspace_bytes` queries.
 *   - return value: 0 = ok, < 0 = bad argument (SAGAN_E*), > 0 = cudaError_t.
 *     `sagan_last_error()` returns a thread-local description of the last failure.
 */
#ifndef SAGAN_B200_H_
#define SAGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAGAN_B200_ABI_VERSION 1

#define SAGAN_EINVAL (-1)        /* bad argument (null pointer, non-positive size, Ip < 1 ...) */
#define SAGAN_EUNSUPPORTED (-2)  /* shape / mode combination not built */
#define SAGAN_EWORKSPACE (-3)    /* workspace too small */

/* math_mode of the attention / conv kernels (the two tolerance tiers of BASELINE.json) */
#define SAGAN_MATH_FP32_STRICT 0 /* fp32 operands, fp32 accumulate, CUDA cores           */
#define SAGAN_MATH_BF16_TC 1     /* bf16 operands, fp32 accumulate, tcgen05 tensor cores */

/* activation fused in conv / deconv epilogues */
#define SAGAN_ACT_NONE 0
#define SAGAN_ACT_LRELU 1 /* LeakyReLU(alpha = slope); generator.py:11, discriminator.py:10 */
#define SAGAN_ACT_TANH 2  /* generator.py:36 */

typedef void* sagan_stream_t; /* cudaStream_t */

int sagan_abi_version(void);
const char* sagan_last_error(void);
/* number of kernels this library has launched in this process (all threads); bench.py's gpu_launches */
unsigned long long sagan_launch_count(void);

/* Process-wide switch (0 / 1; anything else only queries; returns the value in force; default 0).  With 1 no
 * ACTIVATION is accumulated with floating-point atomics (the conv kernels stop splitting their reduction over CTAs
 * for layers with few output pixels), so every forward pass -- and with it every LeakyReLU mask -- is bit-reproducible
 * from run to run.  Weight-gradient reductions keep their atomics (summation-order noise of ~1e-7, no mask involved).
 * Costs a few percent on small batches; used by the graph-vs-eager parity test. */
int sagan_deterministic_forward(int set);

/* ------------------------------------------------------------------------------------------
 * Spectral normalisation.  Replaces SpectralNormalization.update_uv, layers.py:50-68
 * (l2normalize layers.py:4-5).  W is the wrapped layer's kernel viewed as the RAW row-major
 * matrix [rows = last kernel axis, cols = numel / rows] (layers.py:56).  One call performs,
 * `Ip` times, v <- l2n(u W), u <- l2n(v W^T); then sigma = sum((u W) * v) [/ factor] and
 * W_bar = W / sigma.  u [rows] is read and updated in place, v [cols] is written.
 * ------------------------------------------------------------------------------------------ */
typedef struct sagan_sn_desc {
  const float* W;     /* [rows, cols] */
  float* u;           /* [rows] in/out */
  float* v;           /* [cols] out */
  float* W_bar;       /* [rows, cols] out (may alias nothing else) */
  void* W_bar_bf16;   /* optional [rows, cols] bf16 copy of W_bar, or NULL */
  float* sigma;       /* [1] out */
  int32_t rows, cols;
  int32_t Ip;         /* >= 1 (layers.py:17-18) */
  float factor;       /* 0 = none (layers.py:65-66) */
} sagan_sn_desc;

typedef struct sagan_sn_plan sagan_sn_plan; /* opaque: device descriptor table + workspace */

/* Builds a multi-tensor plan: ONE cooperative launch normalises all n matrices
 * (the 13 / 8 spectrally-normalised kernels of G / D).  Allocates device memory (table +
 * workspace) -- call outside the step loop.  `descs_host` is copied. */
int sagan_sn_plan_create(const sagan_sn_desc* descs_host, int n, int device, sagan_sn_plan** plan_out);
int sagan_sn_plan_run(sagan_sn_plan* plan, sagan_stream_t stream);
/* The wrapped layer called with training=False (sample dump, sagan/main.py:333): no power iteration.  sigma is
 * recomputed from the stored u, v and the CURRENT kernels, sum((u W_mat) * v) [/ factor] (layers.py:62-66), and
 * W_bar = W / sigma rewritten; u and v are left alone. */
int sagan_sn_plan_refresh(sagan_sn_plan* plan, sagan_stream_t stream);
int sagan_sn_plan_destroy(sagan_sn_plan* plan);
/* Diagnostics: device-side durations (ms, %globaltimer) of the five phases of the plan's most recent run
 * [u W partials, s + ||s||, v W^T partials, t + ||t||, u / sigma / W_bar].  Synchronises the device. */
int sagan_sn_plan_phase_times(sagan_sn_plan* plan, float* ms_host /* [5] */);
/* algorithmic HBM bytes of one run of the plan (8 B/element, 10 with the bf16 copy,
 * 16/18 when the matrix exceeds 64 MB) */
unsigned long long sagan_sn_plan_algorithmic_bytes(const sagan_sn_plan* plan);

/* d L / d W given d L / d W_bar, with u, v constants:
 *   dW = (dW_bar - (sum dW_bar * W_bar) * reshape(u^T v) / factor) / sigma
 * `ws` holds >= sagan_sn_backward_workspace_bytes(rows*cols) bytes. */
size_t sagan_sn_backward_workspace_bytes(long long numel);
int sagan_sn_backward(const float* dW_bar, const float* W_bar, const float* u, const float* v,
                      const float* sigma, float factor, float* dW, int rows, int cols,
                      void* ws, size_t ws_bytes, sagan_stream_t stream);

/* The same for up to 16 kernels (all spectrally-normalised kernels of one network) in two launches.
 * accumulate != 0: dW += ... (gradient buckets that collect several backward passes, sagan/main.py:181-189).
 * ws: >= 16 * 64 floats. */
typedef struct sagan_sn_bwd_desc {
  const float* dW_bar;
  const float* W_bar;
  const float* u;
  const float* v;
  const float* sigma;
  float* dW;
  float factor;      /* 0 = none */
  int32_t rows, cols;
} sagan_sn_bwd_desc;
int sagan_sn_backward_multi(const sagan_sn_bwd_desc* descs_host, int n, int accumulate, void* ws, size_t ws_bytes,
                            sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Self-attention block.  Replaces Attention_Layer.call, layers.py:93-120 (paper form):
 *   theta = X Wq + bq, phi = X Wk + bk, g = X Wv + bv, P = softmax(theta phi^T), A = P g,
 *   Y = X + gamma * (A Wo + bo).
 * X, Y: [B, N, C] (NHWC, N = H*W).  d = C/8, dv = C/2 (layers.py:82-85).  Wq, Wk: [C, d];
 * Wv: [C, dv]; Wo: [dv, C]  (Keras [1,1,cin,cout] kernels, already spectrally normalised).
 * gamma: device scalar (`sigma`, layers.py:76-79).  Saved for backward: lse [B, N]
 * (row-wise log-sum-exp of the logits) and A [B, N, dv].  The [B,N,N] map is never written.
 * Supported shapes: FP32_STRICT C in {8,16,32,64}, any N; BF16_TC C in {16,32,64}, any N (fused
 * tcgen05 forward and backward) and C in {128,256,512} with N % 128 == 0 (fused forward; the
 * backward of this regime is composed of GEMM launches and keeps the [N,N] maps of ONE sample
 * at a time in the workspace).  Anything else returns SAGAN_EUNSUPPORTED.
 * ------------------------------------------------------------------------------------------ */
size_t sagan_attn_workspace_bytes(int B, int N, int C, int math_mode);
int sagan_attn_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                   const float* Wv, const float* bv, const float* Wo, const float* bo,
                   const float* gamma, float* Y, float* lse, float* A_saved, int B, int N, int C,
                   int math_mode, void* ws, size_t ws_bytes, sagan_stream_t stream);
/* Gradients of the block.  Parameter gradients are OVERWRITTEN (not accumulated).
 * Any of the d* parameter outputs may be NULL together with dX to skip it. */
int sagan_attn_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                   const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                   const float* gamma, const float* lse, const float* A_saved, float* dX, float* dWq,
                   float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo,
                   float* dgamma, int B, int N, int C, int math_mode, void* ws, size_t ws_bytes,
                   sagan_stream_t stream);

/* Down-sampled keys / values: what layers.py:96,100,113 reaches for -- `MaxPool2D` on
 * phi and g, so that a query attends to N/4 keys (the literal `MaxPool2D(2, 1)` + reshape is ill-formed, SURVEY.md
 * appendix A.5; the reading built here is the SAGAN paper's: pool 2 / stride 2 / 'valid' over the [H, W] token grid,
 * per channel, first maximum of a window wins as in TF).  Same arguments as sagan_attn_fwd / sagan_attn_bwd with the
 * token grid given explicitly: X, Y [B, H*W, C]; H, W even; C in {8,16,32,64} (C = 8 always runs the fp32 kernels).
 * lse [B, H*W] and A_saved [B, H*W, C/2] as before; the backward scatters the pooled-key / pooled-value gradients to
 * the window positions that won.  Workspace: sagan_attn_pool_workspace_bytes. */
size_t sagan_attn_pool_workspace_bytes(int B, int H, int W, int C, int math_mode);
int sagan_attn_pool_fwd(const float* X, const float* Wq, const float* bq, const float* Wk, const float* bk,
                        const float* Wv, const float* bv, const float* Wo, const float* bo,
                        const float* gamma, float* Y, float* lse, float* A_saved, int B, int H, int W, int C,
                        int math_mode, void* ws, size_t ws_bytes, sagan_stream_t stream);
int sagan_attn_pool_bwd(const float* dY, const float* X, const float* Wq, const float* bq, const float* Wk,
                        const float* bk, const float* Wv, const float* bv, const float* Wo, const float* bo,
                        const float* gamma, const float* lse, const float* A_saved, float* dX, float* dWq,
                        float* dbq, float* dWk, float* dbk, float* dWv, float* dbv, float* dWo, float* dbo,
                        float* dgamma, int B, int H, int W, int C, int math_mode, void* ws, size_t ws_bytes,
                        sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Conv2D (generator.py:36, discriminator.py:8,35; the 1x1 convs of layers.py:82-85; Dense as
 * the H=W=k=1 case, generator.py:25).  Geometry is explicit so TF 'same' padding asymmetry
 * (k=4,s=1: 1 before / 2 after) is the caller's to state: pad_t / pad_l = padding before.
 *   fwd   : y[b,ho,wo,co] = act(sum x[b,ho*s-pad_t+kh, wo*s-pad_l+kw, ci] w[kh,kw,ci,co] + bias[co])
 *   dgrad : dx = conv^T(dy, w)      (also the forward of Conv2DTranspose, generator.py:8-9)
 *   wgrad : dw[kh,kw,ci,co] = sum_m x[...] dy[m,co]; dbias[co] = sum_m dy[m,co] (dbias may be NULL)
 * ------------------------------------------------------------------------------------------ */
typedef struct sagan_conv_geom {
  int32_t B, H, W, Cin;    /* input  [B,H,W,Cin]  */
  int32_t Ho, Wo, Cout;    /* output [B,Ho,Wo,Cout] */
  int32_t kh, kw, stride, pad_t, pad_l;
} sagan_conv_geom;

int sagan_conv2d_fwd(const float* x, const float* w, const float* bias, float* y,
                     const sagan_conv_geom* g, int act, float slope, int math_mode, sagan_stream_t stream);
int sagan_conv2d_dgrad(const float* dy, const float* w, float* dx, const sagan_conv_geom* g,
                       int math_mode, sagan_stream_t stream);
/* dw must hold kh*kw*Cin*Cout floats; it is overwritten. */
int sagan_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias,
                       const sagan_conv_geom* g, int math_mode, sagan_stream_t stream);

/* Operand precision of the tensor-core conv kernels (process-wide; BF16_TC math mode only):
 *   SAGAN_CONV_TC_SPLIT_BF16 (default)  every operand as hi + lo bf16 tiles, three MMAs per K step (hi*hi + hi*lo +
 *                                       lo*hi): fp32-grade products, forward / backward-data / backward-filter alike
 *   SAGAN_CONV_TC_TF32                  round-1 arithmetic: kind::tf32 forward / backward-data, plain bf16 backward-filter
 * Pass one of the two to select it; anything else (e.g. -1) only queries.  Returns the precision in force. */
#define SAGAN_CONV_TC_TF32 1
#define SAGAN_CONV_TC_SPLIT_BF16 2
int sagan_conv_tc_precision(int set);

/* dz = dy * act'(y) from the activation OUTPUT y (LeakyReLU: y>0 ? 1 : slope; tanh: 1 - y^2). */
int sagan_act_bwd(const float* y, const float* dy, float* dz, long long n, int act, float slope,
                  sagan_stream_t stream);

/* Elementwise glue of the residual topologies (models/generator.py:6-21,
 * models/discriminator.py:6-38): y[i] = act(a[i] + bias[i % C] + residual[i]); bias and residual may be NULL.
 * Covers the stand-alone ReLU / LeakyReLU in front of a conv, the bias of a Conv2DTranspose and `layers.add`.
 * Backward: sagan_act_bwd on y, then the gradient passes to a / residual unchanged and sagan_colsum gives d bias. */
int sagan_ew_fwd(const float* a, const float* bias, const float* residual, float* y, long long n, int C,
                 int act, float slope, sagan_stream_t stream);
/* out[c] = sum_r x[r, c]  (x [rows, C]; out is overwritten) */
int sagan_colsum(const float* x, float* out, long long rows, int C, sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * BatchNormalization(training) + LeakyReLU, generator.py:10-11.  x, y: [rows, C] (NHWC flattened).
 * Batch statistics are per replica (plain BatchNormalization, not SyncBN).  save_mean/save_invstd [C]
 * are kept for the backward; moving_mean / moving_var (may be NULL) get the Keras update
 * m <- momentum*m + (1-momentum)*batch (biased variance).  ws: sagan_bn_workspace_bytes(C).
 * ------------------------------------------------------------------------------------------ */
size_t sagan_bn_workspace_bytes(int C);
int sagan_bn_lrelu_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean,
                       float* save_invstd, float* moving_mean, float* moving_var, long long rows, int C,
                       float eps, float momentum, float slope, void* ws, size_t ws_bytes,
                       sagan_stream_t stream);
int sagan_bn_lrelu_bwd(const float* dy, const float* x, const float* y, const float* gamma,
                       const float* save_mean, const float* save_invstd, float* dx, float* dgamma,
                       float* dbeta, long long rows, int C, float slope, void* ws, size_t ws_bytes,
                       sagan_stream_t stream);
/* Inference-mode BatchNormalization + LeakyReLU -- generator(..., training=False), the sample dumps of
 * sagan/main.py:333 (Keras: (x - moving_mean) / sqrt(moving_var + eps) * gamma + beta).  No statistics, no
 * workspace, nothing saved. */
int sagan_bn_lrelu_infer(const float* x, const float* gamma, const float* beta, const float* moving_mean,
                         const float* moving_var, float* y, long long rows, int C, float eps, float slope,
                         sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Weight normalisation: the wrapper the `sagan/` tree actually wraps its layers with
 * (sagan/layers.py:75-135, TF-Addons WeightNormalization under the name `SpectralNormalization`):
 *   kernel = l2_normalize(v, all axes but the last) * g = v * rsqrt(max(sum_rows v^2, 1e-12)) * g      sagan/layers.py:124
 * v, w, dw, dv: [rows, cols] row-major views of the Keras kernel (cols = its last, filter axis); g, dg: [cols].
 * fwd writes w and inv_norm [cols] (kept for the backward); bwd: dot = sum_rows dw * v,
 *   dv = g inv_norm (dw - v dot inv_norm^2),  dg = dot inv_norm.   dot_ws: [cols] scratch.
 * The data-dependent initialisation (sagan/layers.py:159-194) is host logic over sagan_bn_lrelu_fwd's batch moments.
 * ------------------------------------------------------------------------------------------ */
int sagan_wn_fwd(const float* v, const float* g, float* w, float* inv_norm, int rows, int cols, sagan_stream_t stream);
int sagan_wn_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg,
                 float* dot_ws, int rows, int cols, sagan_stream_t stream);

/* Input contract of the reference's record reader (sagan/dataset.py:27-40): a record is the raw uint8 HWC image (+ an
 * int64 label); the reader feeds float32 NHWC `image * (2. / 255) - 1.`.  dst[i] = float(src[i]) * scale + shift as two
 * separately rounded fp32 operations (bit-identical to the un-fused TF / numpy float32 arithmetic).  Lets a step take
 * its batch as uint8 (4x fewer host -> device bytes) and decode it on the device. */
int sagan_u8_to_f32(const uint8_t* src, float* dst, long long n, float scale, float shift, sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hinge losses, sagan/main.py:21-27, with the scaling of main.py:184,201 folded in:
 *   D: L = relu(1 - d_real) + relu(1 + d_fake);   G: L = -d_fake
 * loss_sum[0] += sum(L) (caller zeroes it); gradients are d(scale * sum L)/d logits with
 * scale = 1 / (n * global_batch).  g_real / g_fake may alias nothing.
 * ------------------------------------------------------------------------------------------ */
int sagan_hinge_d(const float* d_real, const float* d_fake, long long n, float scale, float* loss_sum,
                  float* g_real, float* g_fake, sagan_stream_t stream);
int sagan_hinge_g(const float* d_fake, long long n, float scale, float* loss_sum, float* g_fake,
                  sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Keras Adam over one flat fp32 bucket (sagan/main.py:119-120: beta_1 = 0, epsilon = 1e-7):
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t * m / (sqrt(v) + eps)
 * hyper [4] device floats = {lr_t, b1, b2, eps} (device-resident so a captured CUDA graph sees the
 * per-step learning rate).  m may be NULL when b1 == 0 (then m == g).  grad_scale multiplies g
 * first (1 for the SUM all-reduce convention of main.py:184).
 * ------------------------------------------------------------------------------------------ */
int sagan_adam_step(float* param, const float* grad, float* m, float* v, long long n,
                    const float* hyper, float grad_scale, sagan_stream_t stream);
/* dst[i] += src[i] for n <= 64 tensors in one launch (gradient accumulation of the small parameters into a network's
 * flat gradient bucket: what `tape.gradient` + `apply_gradients` leave to one accumulation op per variable). */
typedef struct sagan_acc_desc {
  float* dst;
  const float* src;
  long long n;
} sagan_acc_desc;
int sagan_accumulate_multi(const sagan_acc_desc* descs_host, int n, sagan_stream_t stream);

/* Learning-rate schedule + Adam bias correction of sagan/main.py:111-120 evaluated on the device:
 *   lr   = lr0 * decay_rate ^ (iterations / decay_steps)          ExponentialDecay(..., staircase=True)
 *   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),  t = iterations + 1   (Keras Adam)
 * writes hyper = {lr_t, b1, b2, eps} and increments the DEVICE counter `iterations` (optimizer.iterations).
 * One launch per apply_gradients; a captured CUDA graph that contains it follows the schedule on replay. */
int sagan_adam_schedule(float* hyper, long long* iterations, double lr0, double decay_rate,
                        long long decay_steps, double b1, double b2, double eps, sagan_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange fused with Keras Adam over NVLink peer memory.  Replaces the
 * replica gradient SUM that MirroredStrategy performs inside optimizer.apply_gradients
 * (sagan/main.py:190,205) plus the Adam update that follows it (main.py:119-120, beta_1 = 0).
 * `peers` holds, for each of the `world` replicas, the PEER-MAPPED device addresses (as seen from the
 * calling replica) of its flat gradient bucket, its flat parameter buffer (both `n` floats, n a
 * multiple of 4 * world) and its flag pad (sagan_dp_flag_bytes() bytes, zero-initialised once).
 * Replica `rank` sums slice `rank` of all gradient buckets (fixed order), updates that slice with its
 * shard `v_shard` [n / world] of the second-moment state and writes the new weights into every
 * replica's parameter buffer.  Two inter-replica barriers (flags in peer memory) bracket the exchange;
 * the call is asynchronous on `stream`; all replicas must issue it in the same order.  `epoch` is a
 * device counter owned by the caller (zero-initialised, one per peers set); `status[0]` becomes 1 if a
 * peer never arrived (the kernel gives up instead of hanging).  hyper = {lr_t, b1, b2, eps} as for
 * sagan_adam_step.
 * ------------------------------------------------------------------------------------------ */
#define SAGAN_DP_MAX_WORLD 8
typedef struct sagan_dp_peers {
  const void* grads[SAGAN_DP_MAX_WORLD];
  void* params[SAGAN_DP_MAX_WORLD];
  void* flags[SAGAN_DP_MAX_WORLD];
} sagan_dp_peers;
/* A replica that waits longer than the timeout (default 30 s) for a peer raises status[0], skips its update and
 * returns; the caller must poll `status` (Trainer.losses() does) -- the replicas have diverged at that point. */
int sagan_dp_max_world(void);
size_t sagan_dp_flag_bytes(void);
int sagan_dp_set_timeout_ms(long long ms);
int sagan_dp_sum_adam(const sagan_dp_peers* peers, int rank, int world, long long n, float* v_shard,
                      const float* hyper, unsigned int* epoch, unsigned int* status, sagan_stream_t stream);
/* The same exchange, which in addition SUMS the replicas' loss sums (strategy.reduce(SUM, per-example loss),
 * sagan/main.py:216-220): loss_peers[q] = peer-mapped address of replica q's [2] fp32 {sum L_D, sum L_G};
 * loss_global [2] (local) receives the totals, identical on every replica.  Rides on the exchange's first barrier
 * (every replica's backward, hence its loss kernels, has completed by then): no extra collective per step. */
int sagan_dp_sum_adam_losses(const sagan_dp_peers* peers, int rank, int world, long long n, float* v_shard,
                             const float* hyper, unsigned int* epoch, unsigned int* status,
                             const void* const* loss_peers, float* loss_global, sagan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAGAN_B200_H_ */
