// Data-parallel update fused with its collective over NVLink / NVSwitch peer memory (new functionality: the reference
// is single-GPU).  One process per GPU; every rank maps every other rank's flat gradient and parameter
// vectors through CUDA IPC.  ONE kernel per rank does what "all-reduce(grads) ; clamp ; RMSprop" does, as
//
//     reduce-scatter : rank r sums shard r of the gradient over all ranks, reading the peers' shards through NVLink
//                      in a fixed rank order (deterministic, and every element is reduced by exactly one rank, so all
//                      replicas stay bit-identical)
//     update         : scale 1/N -> clamp(-10,10) -> optim.rmsprop on shard r (002_train_baseline.lua:329,408;
//                      misc/rmsprop_lrscale.lua:26-34); the RMSprop state is thereby sharded N ways
//     all-gather     : the updated parameter shard is stored straight into every rank's parameter vector
//
// so per step each GPU moves 2 * (N-1)/N * P * 4 bytes over NVLink and the optimizer's HBM traffic drops N-fold.
// The exchange runs per RANGE of parameter blocks -- arch 1: {encoder + embedding} and {multimodal} -- each as soon as its
// gradient is final (nvqa_dp_train_step); a rank owns 1/N of every range.  Cross-GPU ordering uses two monotonically increasing flag words per (block, rank, peer), written
// remotely with st.release.sys and polled locally with ld.acquire.sys:
//     ready[r] >= 2k+1 : rank r's gradients of step k are complete (signalled by the first CTA of rank r's kernel, which
//                        follows the backward pass on the stream)
//     done[r]  >= 2k+2 : rank r has finished reading everybody's gradients and writing its parameter shard everywhere;
//                        a rank's kernel does not exit before it has seen done from ALL ranks, so neither its
//                        gradients nor its parameters are touched by a peer once the next kernel on its stream starts.
#include "model.cuh"

#include <cstdlib>

namespace nvqa {

constexpr int DP_MAX = 8;                       // ranks of one NVSwitch domain
// Flag page of a rank (128 words): per parameter block k = 0..2  ready[k][16] at 32 k, done[k][16] at 32 k + 16 (both
// indexed by the WRITER's rank); [96 + k] = CTA completion counter of block k's kernel; [100] = "a wait timed out".
constexpr int DP_FLAG_WORDS = 128, DP_CTR = 96, DP_ERR = 100;
struct DpPeers {
  const float* g[DP_MAX];
  float* x[DP_MAX];
  unsigned int* flags[DP_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dp_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Waits for a peer's flag.  A peer that lags (validation pass, checkpoint, data stall) is waited for up to timeout_ns of
// WALL-CLOCK time (NVQA_DP_TIMEOUT_S, default 600 s); after that the wait gives up, raises the error word of this rank's
// flag page -- which nvqa_sync / nvqa_loss turn into an error return -- and the kernel runs to completion instead of
// trapping (a trap would take the CUDA context, i.e. the whole process, with it).
__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target, unsigned long long timeout_ns,
                                           unsigned int* err) {
  const unsigned long long t0 = dp_now_ns();
  while (ld_acquire_sys(p) < target) {
    __nanosleep(100);
    if (dp_now_ns() - t0 > timeout_ns) {
      *err = 1u;
      __threadfence_system();
      return;
    }
  }
}
// peer memory is read exactly once per step: bypass L1 (lines of a remote GPU must never be served stale)
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One parameter block's shard [lo4, lo4 + n4) (float4 units of the flat vector) of this rank: grid-stride, so the same
// kernel runs wide on the main stream or with a handful of CTAs on the side stream beside the 128-CTA LSTM backward.
__global__ void __launch_bounds__(256)
dp_fused_rmsprop_kernel(DpPeers p, float* __restrict__ rms, unsigned int* my_flags, int block, int rank, int world,
                        long long lo4, long long n4, unsigned int step, float lr, float alpha, float oma, float eps, float wd,
                        float clampv, float gscale, unsigned long long timeout_ns) {
  unsigned int* ready = my_flags + 32 * block;
  unsigned int* done = ready + 16;
  // "my gradients of this range are final": everything earlier on this stream -- the backward kernels -- has completed
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + 32 * block + rank, 2 * step + 1);
  }
  if ((int)threadIdx.x < world) spin_until(ready + threadIdx.x, 2 * step + 1, timeout_ns, my_flags + DP_ERR);
  __syncthreads();
  // U items per thread and iteration with all their loads issued before the first use: the side-stream launch has only a
  // handful of CTAs and every load crosses NVLink (~2 us), so a thread must keep U x world loads in flight (one item per
  // iteration made the 2-GPU exchange of the multimodal block 224 dependent round trips long: measured +0.1 ms per step)
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += U * stride) {
    float4 g[U], xv[U], mv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) {
        const long long e = (lo4 + i) * 4;
#pragma unroll
        for (int r = 0; r < DP_MAX; ++r) {
          if (r < world) {
            const float4 v = ld_peer(p.g[r] + e);
            g[u].x += v.x; g[u].y += v.y; g[u].z += v.z; g[u].w += v.w;
          }
        }
        xv[u] = *reinterpret_cast<const float4*>(p.x[rank] + e);
        mv[u] = *reinterpret_cast<const float4*>(rms + e);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) continue;
      const long long e = (lo4 + i) * 4;
#define UP(k)                                                       \
      { float gg = fminf(fmaxf(g[u].k * gscale, -clampv), clampv);  \
        gg += wd * xv[u].k;                                         \
        mv[u].k = alpha * mv[u].k + oma * gg * gg;                  \
        xv[u].k -= lr * (gg / (sqrtf(mv[u].k) + eps)); }
      UP(x) UP(y) UP(z) UP(w)
#undef UP
      *reinterpret_cast<float4*>(rms + e) = mv[u];
#pragma unroll
      for (int r = 0; r < DP_MAX; ++r)
        if (r < world) *reinterpret_cast<float4*>(p.x[r] + e) = xv[u];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(my_flags + DP_CTR + block, 1u);
    if (prev == gridDim.x - 1) {                 // last CTA of this rank: everything this rank reads / writes is done
      my_flags[DP_CTR + block] = 0;
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(p.flags[r] + 32 * block + 16 + rank, 2 * step + 2);
      for (int r = 0; r < world; ++r) spin_until(done + r, 2 * step + 2, timeout_ns, my_flags + DP_ERR);
    }
  }
}

}  // namespace nvqa

using namespace nvqa;

struct DpBlob {
  cudaIpcMemHandle_t grads, params, flags;
  int64_t P;
  int32_t device, pad;
};

extern "C" int nvqa_dp_blob_size(void) { return (int)sizeof(DpBlob); }

extern "C" int nvqa_dp_export(nvqa_model* m, void* blob_out) {
  NVQA_CHECK(m && blob_out, "null argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  if (!m->dp_flags) {
    NVQA_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->dp_flags), DP_FLAG_WORDS * sizeof(unsigned int)));
    m->allocs.push_back(m->dp_flags);
    NVQA_CUDA(cudaMemset(m->dp_flags, 0, DP_FLAG_WORDS * sizeof(unsigned int)));
  }
  DpBlob b;
  memset(&b, 0, sizeof(b));
  NVQA_CUDA(cudaIpcGetMemHandle(&b.grads, m->grads));
  NVQA_CUDA(cudaIpcGetMemHandle(&b.params, m->params));
  NVQA_CUDA(cudaIpcGetMemHandle(&b.flags, m->dp_flags));
  b.P = m->P;
  b.device = m->cfg.device;
  memcpy(blob_out, &b, sizeof(b));
  return 0;
}

extern "C" int nvqa_dp_connect(nvqa_model* m, int32_t rank, int32_t world, const void* blobs) {
  NVQA_CHECK(m && blobs && m->dp_flags, "nvqa_dp_connect: call nvqa_dp_export on every rank first");
  NVQA_CHECK(world >= 1 && world <= DP_MAX && rank >= 0 && rank < world, "rank / world out of range (one NVSwitch domain: <= 8)");
  NVQA_CHECK(m->dp_world == 0, "already connected");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  const DpBlob* b = static_cast<const DpBlob*>(blobs);
  for (int r = 0; r < world; ++r) {
    NVQA_CHECK(b[r].P == m->P, "peer model has a different parameter count");
    if (r == rank) {
      m->dp_peer_grads[r] = m->grads; m->dp_peer_params[r] = m->params; m->dp_peer_flags[r] = m->dp_flags;
      continue;
    }
    void *pg = nullptr, *px = nullptr, *pf = nullptr;
    NVQA_CUDA(cudaIpcOpenMemHandle(&pg, b[r].grads, cudaIpcMemLazyEnablePeerAccess));
    NVQA_CUDA(cudaIpcOpenMemHandle(&px, b[r].params, cudaIpcMemLazyEnablePeerAccess));
    NVQA_CUDA(cudaIpcOpenMemHandle(&pf, b[r].flags, cudaIpcMemLazyEnablePeerAccess));
    m->dp_opened.push_back(pg); m->dp_opened.push_back(px); m->dp_opened.push_back(pf);
    m->dp_peer_grads[r] = static_cast<float*>(pg);
    m->dp_peer_params[r] = static_cast<float*>(px);
    m->dp_peer_flags[r] = static_cast<unsigned int*>(pf);
  }
  if (!m->dp_fork) {
    NVQA_CUDA(cudaEventCreateWithFlags(&m->dp_fork, cudaEventDisableTiming));
    NVQA_CUDA(cudaEventCreateWithFlags(&m->dp_join, cudaEventDisableTiming));
  }
  m->dp_rank = rank; m->dp_world = world;
  for (int k = 0; k < 3; ++k) m->dp_steps[k] = 0;
  return 0;
}

extern "C" int nvqa_dp_disconnect(nvqa_model* m) {
  NVQA_CHECK(m, "null model");
  if (m->dp_world) NVQA_CUDA(cudaSetDevice(m->cfg.device));
  if (m->aux_stream) NVQA_CUDA(cudaStreamSynchronize(m->aux_stream));
  for (void* p : m->dp_opened) cudaIpcCloseMemHandle(p);
  m->dp_opened.clear();
  m->dp_world = 0;
  return 0;
}

// Has a peer wait of this rank's fused kernels timed out since the connect?  (blocks: reads one word from the device)
extern "C" int nvqa_dp_status(nvqa_model* m, int32_t* timed_out) {
  NVQA_CHECK(m && timed_out, "null argument");
  *timed_out = 0;
  if (!m->dp_flags) return 0;
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  unsigned int v = 0;
  NVQA_CUDA(cudaMemcpy(&v, m->dp_flags + DP_ERR, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = v != 0;
  return 0;
}

static unsigned long long dp_timeout_ns() {
  static unsigned long long ns = 0;
  if (!ns) {
    const char* e = getenv("NVQA_DP_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (!(s > 0)) s = 600.0;
    ns = (unsigned long long)(s * 1e9);
  }
  return ns;
}

// One exchange = a contiguous range of parameter blocks [blk0, blk1] of the flat vector with one gradient scale (channel =
// blk0 selects the flag words): the fused reduce-scatter + update + all-gather kernel over this rank's 1/N of the range.
// side = true: on the side stream with few CTAs (beside the LSTM backward); the unused dynamic shared memory keeps these
// CTAs off the SMs of the persistent LSTM kernel (which owns all of its SM's shared memory), i.e. on the idle ones.
static int dp_range(nvqa_model* m, int blk0, int blk1, float lr, float alpha, float eps, float wd, float clamp, bool side) {
  DpPeers p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < m->dp_world; ++r) { p.g[r] = m->dp_peer_grads[r]; p.x[r] = m->dp_peer_params[r]; p.flags[r] = m->dp_peer_flags[r]; }
  const long long b0 = m->off_blk[blk0] / 4, b1 = m->off_blk[blk1 + 1] / 4;              // block offsets are multiples of 4
  const long long per = (b1 - b0 + m->dp_world - 1) / m->dp_world;
  const long long lo4 = std::min(b1, b0 + per * m->dp_rank), hi4 = std::min(b1, lo4 + per);
  const long long n4 = hi4 - lo4;
  const unsigned int step = m->dp_steps[blk0]++;
  // -lr_scale of the arch 1 trainer variants multiplies the encoder and embedding gradients before the clamp
  // (003_train_ae_based_wp.lua:344-346), exactly as nvqa_rmsprop_step does
  float gscale = 1.0f / (float)m->dp_world;
  if (m->cfg.arch == 1 && blk1 < NVQA_BLOCK_MULTIMODAL) gscale *= m->lr_scale;
  cudaStream_t s = side ? m->aux_stream : m->stream;
  static int side_ctas = -1;
  if (side_ctas < 0) { const char* e = getenv("NVQA_DP_SIDE_CTAS"); side_ctas = e ? std::max(1, atoi(e)) : 16; }
  static int main_ctas = -1;
  if (main_ctas < 0) { const char* e = getenv("NVQA_DP_MAIN_CTAS"); main_ctas = e ? std::max(1, atoi(e)) : 148 * 8; }
  const int full = std::max(1, std::min(ceil_div(n4, 256), main_ctas));
  const int grid = side ? std::min(full, side_ctas) : full;
  const size_t smem = side ? 64 * 1024 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    NVQA_CUDA(cudaFuncSetAttribute(dp_fused_rmsprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  dp_fused_rmsprop_kernel<<<grid, 256, smem, s>>>(p, m->rms, m->dp_flags, blk0, m->dp_rank, m->dp_world, lo4, n4, step, lr, alpha,
                                                 (float)(1.0 - (double)alpha), eps, wd, clamp, gscale, dp_timeout_ns());
  NVQA_LAUNCHED();
  return 0;
}
// the exchanges of one step: arch 1 = {encoder + embedding (one range, one lr_scale), multimodal}; other architectures:
// the whole flat vector at once
static int dp_tail_ranges(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp, bool with_multimodal) {
  if (m->cfg.arch != 1) return dp_range(m, 0, 2, lr, alpha, eps, wd, clamp, false);
  NVQA_TRY(dp_range(m, NVQA_BLOCK_ENCODER, NVQA_BLOCK_EMBEDDING, lr, alpha, eps, wd, clamp, false));
  if (with_multimodal) NVQA_TRY(dp_range(m, NVQA_BLOCK_MULTIMODAL, NVQA_BLOCK_MULTIMODAL, lr, alpha, eps, wd, clamp, false));
  return 0;
}

static int drop_lookup_grad_dp(nvqa_model* m) {
  // the literal reference's gradient-less LookupTable (nvqa_set_lookup_grad_literal), as in nvqa_rmsprop_step
  if (m->lookup_grad_literal && m->glookup)
    NVQA_CUDA(cudaMemsetAsync(m->glookup, 0, (size_t)(m->cfg.V + 1) * m->cfg.E * 4, m->stream));
  return 0;
}

// gradients of every rank -> (sum / world) -> clamp -> RMSprop -> parameters of every rank; see the header comment
extern "C" int nvqa_dp_rmsprop_step(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp) {
  NVQA_CHECK(m && m->dp_world >= 1, "nvqa_dp_rmsprop_step: not connected (nvqa_dp_export / nvqa_dp_connect)");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  umma_workspace_invalidate(m->ws);
  NVQA_TRY(drop_lookup_grad_dp(m));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (m->profiling) {
    NVQA_CUDA(cudaEventCreate(&e0)); NVQA_CUDA(cudaEventCreate(&e1));
    NVQA_CUDA(cudaEventRecord(e0, m->stream));
  }
  NVQA_TRY(dp_tail_ranges(m, lr, alpha, eps, wd, clamp, true));
  if (m->profiling) {
    NVQA_CUDA(cudaEventRecord(e1, m->stream));
    m->prof[CAT_OPT].pending.emplace_back(e0, e1);
    m->prof[CAT_OPT].launches += m->cfg.arch == 1 ? 2 : 1;
  }
  return 0;
}

// One whole data-parallel training step on the batch already set: JdJ with the exchange of each parameter block started as
// soon as its gradient is final.  arch 1: the multimodal block (53 % of the gradient) is final after the head backward and
// none of its weights is read again in this step, so its reduce-scatter + update + all-gather runs on a side stream with
// a few CTAs on the 20 SMs the 128-CTA persistent LSTM backward leaves free; encoder and embedding follow on the main
// stream after their gradients (the LSTM backward / dgrad GEMMs still read the encoder weights until then).
extern "C" int nvqa_dp_train_step(nvqa_model* m, float lr, uint64_t seed, float alpha, float eps, float wd, float clamp) {
  NVQA_CHECK(m && m->dp_world >= 1, "nvqa_dp_train_step: not connected (nvqa_dp_export / nvqa_dp_connect)");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  // overlapping the multimodal exchange pays from 4 ranks on (B200, round 2: 8 ranks 1.777 -> 1.691 ms); with 2 ranks the
  // 16-CTA side kernel is slower than the tail kernel it saves (1.622 -> 1.645 ms), so the default is by world size
  static int overlap_env = -2;
  if (overlap_env == -2) { const char* e = getenv("NVQA_DP_OVERLAP"); overlap_env = e ? atoi(e) : -1; }
  const int overlap = overlap_env >= 0 ? overlap_env : (m->dp_world >= 4 ? 1 : 0);
  m->fused_step = true;            // a backward follows: its gradient slices are cleared beside the forward (model.cu)
  const int frc = nvqa_forward(m, NVQA_MODE_TRAIN, seed);
  m->fused_step = false;
  NVQA_TRY(frc);
  if (m->cfg.arch != 1 || !overlap || m->profiling) {
    NVQA_TRY(nvqa_backward(m, NVQA_PHASE_ALL));
    return nvqa_dp_rmsprop_step(m, lr, alpha, eps, wd, clamp);
  }
  // head backward; its two AxB weight-gradient GEMMs are deferred to the side stream (model.cu: aux_launch_bwd), so the
  // multimodal gradient becomes final ON THE SIDE STREAM and its exchange is enqueued there, behind them; both run beside
  // the persistent LSTM backward of the main stream
  m->defer_head = true;
  const int rc_head = nvqa_backward(m, NVQA_PHASE_HEAD);
  m->defer_head = false;
  NVQA_TRY(rc_head);
  NVQA_TRY(aux_launch_bwd(m, true));
  NVQA_TRY(dp_range(m, NVQA_BLOCK_MULTIMODAL, NVQA_BLOCK_MULTIMODAL, lr, alpha, eps, wd, clamp, true));
  NVQA_CUDA(cudaEventRecord(m->dp_join, m->aux_stream));
  NVQA_TRY(nvqa_backward(m, NVQA_PHASE_LSTM));
  NVQA_TRY(nvqa_backward(m, NVQA_PHASE_EMBED));
  NVQA_TRY(dp_tail_ranges(m, lr, alpha, eps, wd, clamp, false));
  NVQA_CUDA(cudaStreamWaitEvent(m->stream, m->dp_join, 0));      // the next forward reads the multimodal weights
  umma_workspace_invalidate(m->ws);
  return 0;
}
