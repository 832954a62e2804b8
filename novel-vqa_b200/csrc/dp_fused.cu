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
// Cross-GPU ordering uses two monotonically increasing flag words per (rank, peer) pair, written remotely with
// st.release.sys and polled locally with ld.acquire.sys:
//     ready[r] >= 2k+1 : rank r's gradients of step k are complete (signalled by a 1-thread kernel that follows the
//                        backward pass on the stream)
//     done[r]  >= 2k+2 : rank r has finished reading everybody's gradients and writing its parameter shard everywhere;
//                        a rank's kernel does not exit before it has seen done from ALL ranks, so neither its
//                        gradients nor its parameters are touched by a peer once the next kernel on its stream starts.
#include "model.cuh"

namespace nvqa {

constexpr int DP_MAX = 8;                       // ranks of one NVSwitch domain
struct DpPeers {
  const float* g[DP_MAX];
  float* x[DP_MAX];
  unsigned int* flags[DP_MAX];                  // [0..15] ready counters, [16..31] done counters, indexed by writer rank
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < target) {
    __nanosleep(100);
    if (clock64() - t0 > 20000000000LL) {       // ~10 s: a peer died or never launched -> fail loudly instead of hanging
      printf("nvqa dp: timed out waiting for a peer flag (want %u)\n", target);
      __trap();
    }
  }
}
// peer memory is read exactly once per step: bypass L1 (lines of a remote GPU must never be served stale)
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void dp_signal_ready_kernel(DpPeers p, int rank, int world, unsigned int value) {
  __threadfence_system();
  if ((int)threadIdx.x < world) st_release_sys(p.flags[threadIdx.x] + rank, value);
}

__global__ void __launch_bounds__(256)
dp_fused_rmsprop_kernel(DpPeers p, float* __restrict__ rms, const unsigned int* my_flags, unsigned int* done_ctr, int rank,
                        int world, long long lo4, long long n4, unsigned int step, float lr, float alpha, float oma, float eps,
                        float wd, float clampv, float gscale) {
  if ((int)threadIdx.x < world) spin_until(my_flags + threadIdx.x, 2 * step + 1);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const long long e = (lo4 + i) * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < DP_MAX; ++r) {
      if (r < world) {
        const float4 v = ld_peer(p.g[r] + e);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    }
    float4 xv = *reinterpret_cast<const float4*>(p.x[rank] + e), mv = *reinterpret_cast<const float4*>(rms + e);
#define UP(k)                                                    \
    { float gg = fminf(fmaxf(g.k * gscale, -clampv), clampv);    \
      gg += wd * xv.k;                                           \
      mv.k = alpha * mv.k + oma * gg * gg;                       \
      xv.k -= lr * (gg / (sqrtf(mv.k) + eps)); }
    UP(x) UP(y) UP(z) UP(w)
#undef UP
    *reinterpret_cast<float4*>(rms + e) = mv;
#pragma unroll
    for (int r = 0; r < DP_MAX; ++r)
      if (r < world) *reinterpret_cast<float4*>(p.x[r] + e) = xv;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done_ctr, 1u);
    if (prev == gridDim.x - 1) {                 // last CTA of this rank: everything this rank reads / writes is done
      *done_ctr = 0;
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(p.flags[r] + 16 + rank, 2 * step + 2);
      for (int r = 0; r < world; ++r) spin_until(my_flags + 16 + r, 2 * step + 2);
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
    NVQA_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->dp_flags), 64 * sizeof(unsigned int)));
    m->allocs.push_back(m->dp_flags);
    NVQA_CUDA(cudaMemset(m->dp_flags, 0, 64 * sizeof(unsigned int)));
    m->dp_done = m->dp_flags + 48;
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
  m->dp_rank = rank; m->dp_world = world; m->dp_step = 0;
  return 0;
}

extern "C" int nvqa_dp_disconnect(nvqa_model* m) {
  NVQA_CHECK(m, "null model");
  if (m->dp_world) NVQA_CUDA(cudaSetDevice(m->cfg.device));
  for (void* p : m->dp_opened) cudaIpcCloseMemHandle(p);
  m->dp_opened.clear();
  m->dp_world = 0;
  return 0;
}

// gradients of every rank -> (sum / world) -> clamp -> RMSprop -> parameters of every rank; see the header comment
extern "C" int nvqa_dp_rmsprop_step(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp) {
  NVQA_CHECK(m && m->dp_world >= 1, "nvqa_dp_rmsprop_step: not connected (nvqa_dp_export / nvqa_dp_connect)");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  umma_workspace_invalidate(m->ws);
  DpPeers p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < m->dp_world; ++r) { p.g[r] = m->dp_peer_grads[r]; p.x[r] = m->dp_peer_params[r]; p.flags[r] = m->dp_peer_flags[r]; }
  const long long n4_all = m->P / 4, per = (n4_all + m->dp_world - 1) / m->dp_world;
  const long long lo4 = std::min<long long>(n4_all, per * m->dp_rank), hi4 = std::min<long long>(n4_all, lo4 + per);
  const long long n4 = hi4 - lo4;
  const unsigned int step = m->dp_step++;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (m->profiling) {
    NVQA_CUDA(cudaEventCreate(&e0)); NVQA_CUDA(cudaEventCreate(&e1));
    NVQA_CUDA(cudaEventRecord(e0, m->stream));
  }
  dp_signal_ready_kernel<<<1, 32, 0, m->stream>>>(p, m->dp_rank, m->dp_world, 2 * step + 1);
  NVQA_LAUNCHED();
  const int grid = std::max(1, ceil_div(n4, 256));
  dp_fused_rmsprop_kernel<<<grid, 256, 0, m->stream>>>(p, m->rms, m->dp_flags, m->dp_done, m->dp_rank, m->dp_world, lo4, n4, step, lr,
                                                      alpha, (float)(1.0 - (double)alpha), eps, wd, clamp, 1.0f / (float)m->dp_world);
  NVQA_LAUNCHED();
  if (m->profiling) {
    NVQA_CUDA(cudaEventRecord(e1, m->stream));
    m->prof[CAT_OPT].pending.emplace_back(e0, e1);
    m->prof[CAT_OPT].launches += 2;
  }
  return 0;
}
