// Data-parallel gradient exchange FUSED with the optimiser, over NVLink peer memory (no NCCL on the data path).
//
// Replaces the replica gradient sum that tf.distribute.MirroredStrategy performs inside
// `optimizer.apply_gradients` (/root/reference/sagan/main.py:190,205: NcclAllReduce of every gradient, then Adam on
// every replica) with ONE kernel per network per update:
//
//   barrier A   every replica's flat gradient bucket is complete                       (flags in peer memory)
//   shard r     replica r reads slice r of ALL replicas' buckets through NVLink (P2P loads, fixed order 0..W-1, so
//               the sum is deterministic and every replica ends up with bit-identical weights), applies Keras Adam
//               (beta_1 = 0: m == g) to slice r with ITS slice of the second-moment state (the optimiser state is
//               sharded, ZeRO-1 style), and writes the updated weights of slice r into EVERY replica's flat
//               parameter buffer (P2P stores)
//   barrier B   all slices have been written everywhere; gradient buckets may be reused
//
// Per replica and update 2 (W-1)/W x bucket bytes cross NVLink (G: 4.9 MB, D: 0.7 MB at church64) instead of an
// all-reduce followed by a full-bucket Adam on every replica.  The buffers are symmetric-memory allocations
// (torch.distributed._symmetric_memory) whose peer-mapped addresses the host passes in `sagan_dp_peers`.
// Every wait is bounded by a wall-clock timeout (sagan_dp_set_timeout_ms, default 30 s): on a lost peer the kernel
// raises `status[0]`, skips the update and still completes its barrier bookkeeping instead of hanging the GPU.
#include "common.cuh"

namespace sagan {

constexpr int DP_MAX_WORLD = 8;
constexpr int DP_THREADS = 256;
// how long a replica waits for its peers at a barrier before it gives up (wall clock, %globaltimer); host-settable
static std::atomic<long long> g_dp_timeout_ns{30ll * 1000 * 1000 * 1000};

struct DpPeers {
  const float* grads[DP_MAX_WORLD];
  float* params[DP_MAX_WORLD];
  unsigned int* flags[DP_MAX_WORLD];   // per replica: [2][DP_MAX_WORLD] uint32 (barrier A row, barrier B row)
  const float* losses[DP_MAX_WORLD];   // optional: per replica [2] loss sums (peer-mapped), null = no loss reduction
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// epoch counter bumped by its own 1-thread launch, so that every CTA of the main kernel reads the same value however
// late it is scheduled, and a captured CUDA graph replays with fresh epochs
__global__ void dp_bump_epoch_kernel(unsigned int* epoch) { *epoch += 1u; }

__device__ __forceinline__ unsigned long long dp_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool dp_barrier(const DpPeers& pr, int rank, int world, int row, unsigned int epoch,
                                           unsigned int* status, bool signal, long long timeout_ns) {
  __syncthreads();
  if (signal && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pr.flags[threadIdx.x] + row * DP_MAX_WORLD + rank, epoch);      // "replica `rank` has arrived"
  }
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const unsigned int* mine = pr.flags[rank] + row * DP_MAX_WORLD + threadIdx.x;
    const unsigned long long t0 = dp_now_ns();
    unsigned int spins = 0;
    // epochs only grow; signed distance handles wrap-around
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 1023u) == 0 && (long long)(dp_now_ns() - t0) > timeout_ns) {
        ok = 0;
        atomicExch(status, 1u);
        break;
      }
    }
  }
  __syncthreads();
  return ok != 0;
}

__global__ void __launch_bounds__(DP_THREADS)
dp_sum_adam_kernel(const DpPeers pr, int rank, int world, long long n, float* __restrict__ v_shard,
                   const float* __restrict__ hyper, const unsigned int* __restrict__ epoch_ptr, unsigned int* status,
                   long long timeout_ns, float* __restrict__ loss_global) {
  const unsigned int epoch = *epoch_ptr;
  // barrier A: only CTA 0 signals, every CTA waits on the local flags.  A CTA whose wait timed out (status raised)
  // skips the update but still takes part in the CTA count and barrier B below, so the counter never goes stale and
  // the peers are not made to wait for a signal that would never come.
  const bool arrived = dp_barrier(pr, rank, world, 0, epoch, status, blockIdx.x == 0, timeout_ns);

  // strategy.reduce(SUM) of the per-replica loss sums (sagan/main.py:216-220) rides on the same barrier: every replica's
  // backward -- hence its loss kernels -- is complete once barrier A has passed; fixed order => identical everywhere
  if (loss_global && arrived && blockIdx.x == 0 && threadIdx.x < 2) {
    float t = 0.f;
    for (int q = 0; q < world; ++q) t += pr.losses[q][threadIdx.x];
    loss_global[threadIdx.x] = t;
  }
  const float lr_t = hyper[0], b2 = hyper[2], eps = hyper[3];     // beta_1 = 0 (sagan/main.py:119-120): m == g
  const long long per = n / world;                                 // n is a multiple of 4 * world (host pads)
  const long long base = (long long)rank * per;
  const long long n4 = per / 4;
  const long long stride = (long long)gridDim.x * DP_THREADS;
  for (long long i = (long long)blockIdx.x * DP_THREADS + threadIdx.x; arrived && i < n4; i += stride) {
    const long long e = base + i * 4;
    float4 g = ld4(pr.grads[0] + e);
    for (int q = 1; q < world; ++q) {
      const float4 t = ld4(pr.grads[q] + e);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 vv = ld4(v_shard + i * 4);
    vv.x = b2 * vv.x + (1.f - b2) * g.x * g.x; vv.y = b2 * vv.y + (1.f - b2) * g.y * g.y;
    vv.z = b2 * vv.z + (1.f - b2) * g.z * g.z; vv.w = b2 * vv.w + (1.f - b2) * g.w * g.w;
    st4(v_shard + i * 4, vv);
    float4 p = ld4(pr.params[rank] + e);
    p.x -= lr_t * g.x / (sqrtf(vv.x) + eps); p.y -= lr_t * g.y / (sqrtf(vv.y) + eps);
    p.z -= lr_t * g.z / (sqrtf(vv.z) + eps); p.w -= lr_t * g.w / (sqrtf(vv.w) + eps);
    for (int q = 0; q < world; ++q) st4(pr.params[q] + e, p);
  }
  // barrier B: every CTA's stores must be visible before CTA 0 signals -> count CTAs on a local counter first
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    unsigned int* done = pr.flags[rank] + 2 * DP_MAX_WORLD;       // local CTA counter (third row of the pad)
    const unsigned int prev = atomicAdd(done, 1u);
    last = (prev == gridDim.x - 1);
    if (last) *done = 0u;
  }
  __syncthreads();
  // the last CTA signals AND waits: the kernel (hence the stream) completes on this replica only after every replica's
  // slice has landed in this replica's parameter buffer and every replica has stopped reading this replica's gradients
  if (last) {
    __threadfence_system();
    dp_barrier(pr, rank, world, 1, epoch, status, true, timeout_ns);
  }
}

}  // namespace sagan

using namespace sagan;

extern "C" int sagan_dp_max_world(void) { return DP_MAX_WORLD; }
extern "C" size_t sagan_dp_flag_bytes(void) { return 4 * DP_MAX_WORLD * sizeof(unsigned int); }

extern "C" int sagan_dp_set_timeout_ms(long long ms) {
  SAGAN_REQUIRE(ms > 0, "sagan_dp_set_timeout_ms: timeout must be positive");
  g_dp_timeout_ns.store(ms * 1000000ll);
  return 0;
}

static int dp_sum_adam_impl(const sagan_dp_peers* peers, int rank, int world, long long n, float* v_shard,
                            const float* hyper, unsigned int* epoch, unsigned int* status, const void* const* loss_peers,
                            float* loss_global, sagan_stream_t stream) {
  SAGAN_REQUIRE(peers && v_shard && hyper && epoch && status, "sagan_dp_sum_adam: null pointer");
  SAGAN_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world,
                "sagan_dp_sum_adam: bad rank / world (%d / %d, at most %d replicas)", rank, world, DP_MAX_WORLD);
  SAGAN_REQUIRE(n > 0 && n % (4ll * world) == 0, "sagan_dp_sum_adam: bucket length %lld must be a multiple of 4 * world", n);
  DpPeers pr{};
  for (int q = 0; q < world; ++q) {
    SAGAN_REQUIRE(peers->grads[q] && peers->params[q] && peers->flags[q], "sagan_dp_sum_adam: null peer pointer (replica %d)", q);
    pr.grads[q] = (const float*)peers->grads[q];
    pr.params[q] = (float*)peers->params[q];
    pr.flags[q] = (unsigned int*)peers->flags[q];
    if (loss_global) {
      SAGAN_REQUIRE(loss_peers && loss_peers[q], "sagan_dp_sum_adam_losses: null loss pointer (replica %d)", q);
      pr.losses[q] = (const float*)loss_peers[q];
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  dp_bump_epoch_kernel<<<1, 1, 0, st>>>(epoch);
  SAGAN_LAUNCH_CHECK();
  const long long n4 = n / world / 4;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(num_sms(), ceil_div<long long>(n4, DP_THREADS)));
  dp_sum_adam_kernel<<<blocks, DP_THREADS, 0, st>>>(pr, rank, world, n, v_shard, hyper, epoch, status,
                                                       g_dp_timeout_ns.load(), loss_global);
  SAGAN_LAUNCH_CHECK();
  return 0;
}

extern "C" int sagan_dp_sum_adam(const sagan_dp_peers* peers, int rank, int world, long long n, float* v_shard,
                                 const float* hyper, unsigned int* epoch, unsigned int* status,
                                 sagan_stream_t stream) {
  return dp_sum_adam_impl(peers, rank, world, n, v_shard, hyper, epoch, status, nullptr, nullptr, stream);
}

extern "C" int sagan_dp_sum_adam_losses(const sagan_dp_peers* peers, int rank, int world, long long n, float* v_shard,
                                        const float* hyper, unsigned int* epoch, unsigned int* status,
                                        const void* const* loss_peers, float* loss_global, sagan_stream_t stream) {
  SAGAN_REQUIRE(loss_peers && loss_global, "sagan_dp_sum_adam_losses: null loss pointer");
  return dp_sum_adam_impl(peers, rank, world, n, v_shard, hyper, epoch, status, loss_peers, loss_global, stream);
}
