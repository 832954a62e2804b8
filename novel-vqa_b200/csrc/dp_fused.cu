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
__device__ __forceinline__ float4 ld_peer(const float* p, int mode) {
  float4 v;
  if (mode == 0)
    asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (mode == 1)
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One parameter block's shard [lo4, lo4 + n4) (float4 units of the flat vector) of this rank: grid-stride, so the same
// kernel runs wide on the main stream or with a handful of CTAs on the side stream beside the 128-CTA LSTM backward.
__global__ void __launch_bounds__(256)
dp_fused_rmsprop_kernel(DpPeers p, float* __restrict__ rms, unsigned int* my_flags, int block, int rank, int world,
                        long long lo4, long long n4, unsigned int step, float lr, float alpha, float oma, float eps, float wd,
                        float clampv, float gscale, unsigned long long timeout_ns, int ldmode) {
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
            const float4 v = ld_peer(p.g[r] + e, ldmode);
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

// ---- the same exchange with BULK ASYNC COPIES (default; NVQA_DP_BULK=0 restores the ld / st kernel above) ----------------
// Measured on two B200s (tools/dp_probe.py): the ld.global / st.global kernel moves a rank's 27.7 MB each way in ~70 us per
// direction-pair -- 190-280 GB/s, whatever its grid size or load flavour -- where the copy engine does 560-750 GB/s.  Here a
// persistent CTA works through 8 KB chunks of the rank's shard in a ring of stages: ONE thread asks the TMA unit for the
// chunk of every rank's gradient (peers through NVLink), of the parameters and of the RMSprop state
// (cp.async.bulk.shared::cluster.global, one mbarrier per stage counts the bytes), 256 threads reduce in the fixed rank
// order / scale / clamp / update in shared memory, and one thread stores the state back and the updated parameters into
// EVERY rank's vector (cp.async.bulk.global.shared::cta, bulk groups).  A few CTAs keep megabytes in flight, so the same
// kernel saturates the links from the 16 side-stream CTAs beside the LSTM backward.  Same flags, same summation order,
// same arithmetic as the kernel above (replicas stay bit-identical; the two kernels agree bit for bit).
constexpr int DPB_CHUNK4_DEFAULT = 512;         // float4 per chunk (8 KB); NVQA_DP_CHUNK_KB overrides
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dpb_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dpb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// false: gave up after timeout_ns of wall-clock time (the error word is raised by the caller)
__device__ __forceinline__ bool dpb_wait(uint32_t bar, uint32_t parity, unsigned long long timeout_ns) {
  uint32_t ok = 0;
  const unsigned long long t0 = dp_now_ns();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (dp_now_ns() - t0 > timeout_ns) return false;
  }
}

// 1024 threads: the update is ~60 dependent instructions per element (IEEE sqrt and divide); with 8 warps per CTA it was
// latency-bound at 2.6 us per 8 KB chunk (72 us for a 27.7 MB shard with no NVLink traffic at all, tools/dp_probe2.sh)
constexpr int DPB_THREADS = 1024;
__global__ void __launch_bounds__(DPB_THREADS, 1)
dp_bulk_rmsprop_kernel(DpPeers p, float* __restrict__ rms, unsigned int* my_flags, int block, int rank, int world,
                       long long lo4, long long n4, unsigned int step, float lr, float alpha, float oma, float eps, float wd,
                       float clampv, float gscale, unsigned long long timeout_ns, int stages, int DPB_CHUNK4, int dbg) {
  extern __shared__ __align__(128) uint8_t dpb_smem[];
  const uint32_t DPB_CHUNK = (uint32_t)DPB_CHUNK4 * 16u;
  // stage s: [world gradient chunks][parameter chunk][state chunk]; the `stages` mbarriers follow the last stage
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dpb_smem);
  const uint32_t stage_bytes = (uint32_t)(world + 2) * DPB_CHUNK;
  const uint32_t bar0 = sbase + (uint32_t)stages * stage_bytes;
  unsigned int* ready = my_flags + 32 * block;
  unsigned int* done = ready + 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) dpb_mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + 32 * block + rank, 2 * step + 1);
  }
  if ((int)threadIdx.x < world) spin_until(ready + threadIdx.x, 2 * step + 1, timeout_ns, my_flags + DP_ERR);
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");         // the acquired gradients are read through the async proxy below

  const long long nchunks = (n4 + DPB_CHUNK4 - 1) / DPB_CHUNK4;
  const long long mine = blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // chunks c = blockIdx.x + k * gridDim.x
  auto issue_loads = [&](long long k) {                    // thread 0: chunk k of this CTA into stage k % stages
    const long long c = blockIdx.x + k * gridDim.x;
    const long long f0 = lo4 + c * DPB_CHUNK4;
    const uint32_t bytes = (uint32_t)(min((long long)DPB_CHUNK4, n4 - c * DPB_CHUNK4) * 16);
    const int s = (int)(k % stages);
    const uint32_t st = sbase + (uint32_t)s * stage_bytes, bar = bar0 + 8 * s;
    dpb_expect_tx(bar, bytes * (uint32_t)(world + 2));
    for (int r = 0; r < world; ++r) bulk_g2s(st + (uint32_t)r * DPB_CHUNK, p.g[(dbg & 1) ? rank : r] + f0 * 4, bytes, bar);
    bulk_g2s(st + (uint32_t)world * DPB_CHUNK, p.x[rank] + f0 * 4, bytes, bar);
    bulk_g2s(st + (uint32_t)(world + 1) * DPB_CHUNK, rms + f0 * 4, bytes, bar);
  };
  if (threadIdx.x == 0)
    for (long long k = 0; k < mine && k < stages; ++k) issue_loads(k);
  bool ok = true;       // a wait that gave up raises the error word; the kernel still runs to completion (no divergent exit)
  for (long long k = 0; k < mine; ++k) {
    const long long c = blockIdx.x + k * gridDim.x;
    const int cnt4 = (int)min((long long)DPB_CHUNK4, n4 - c * DPB_CHUNK4);
    const int s = (int)(k % stages);
    uint8_t* st = dpb_smem + (size_t)s * stage_bytes;
    if (!dpb_wait(bar0 + 8 * s, (uint32_t)(k / stages) & 1u, timeout_ns)) ok = false;
    float4* xs = reinterpret_cast<float4*>(st + (size_t)world * DPB_CHUNK);
    float4* ms = reinterpret_cast<float4*>(st + (size_t)(world + 1) * DPB_CHUNK);
    for (int j = threadIdx.x; j < cnt4; j += DPB_THREADS) {
      {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
          const float4 v = reinterpret_cast<const float4*>(st + (size_t)r * DPB_CHUNK)[j];
          g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
        }
        float4 xv = xs[j], mv = ms[j];
#define UP(q)                                                      \
        { float gg = fminf(fmaxf(g.q * gscale, -clampv), clampv);  \
          gg += wd * xv.q;                                         \
          mv.q = alpha * mv.q + oma * gg * gg;                     \
          xv.q -= lr * (gg / (sqrtf(mv.q) + eps)); }
        UP(x) UP(y) UP(z) UP(w)
#undef UP
        xs[j] = xv; ms[j] = mv;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the updated chunks are read by the bulk stores
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long f0 = lo4 + c * DPB_CHUNK4;
      const uint32_t bytes = (uint32_t)cnt4 * 16u;
      const uint32_t sx = sbase + (uint32_t)s * stage_bytes + (uint32_t)world * DPB_CHUNK;
      bulk_s2g(rms + f0 * 4, sx + DPB_CHUNK, bytes);
      for (int r = 0; r < world; ++r)
        if (!(dbg & 2) || r == rank) bulk_s2g(p.x[r] + f0 * 4, sx, bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // refill the stage of the PREVIOUS chunk once its stores have read it (one store group stays in flight)
      if (k >= 1 && k - 1 + stages < mine) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        issue_loads(k - 1 + stages);
      }
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");         // every store of this CTA has been performed
  }
  if (!ok) my_flags[DP_ERR] = 1u;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(my_flags + DP_CTR + block, 1u);
    if (prev == gridDim.x - 1) {
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
  // one exchange kernel for the whole vector when the multimodal block is not exchanged early (see nvqa_dp_train_step) and
  // one gradient scale covers everything; NVQA_DP_WHOLE=0 keeps the two ranges
  {
    const char* eo = getenv("NVQA_DP_OVERLAP");
    const char* ew = getenv("NVQA_DP_WHOLE");
    const int overlap = eo ? atoi(eo) : (world >= 4 ? 1 : 0);
    m->dp_whole_vector = m->cfg.arch == 1 && !overlap && m->lr_scale == 1.0f && !(ew && atoi(ew) == 0);
  }
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

// How the flat vector is partitioned among the ranks (which also shards the RMSprop state): *whole_vector = 1 -> one range,
// every rank owns the r-th 1/N of the whole vector; 0 -> two ranges {encoder + embedding}, {multimodal}, a rank owns the
// r-th 1/N of each (arch 1 with the multimodal exchange overlapped, or with -lr_scale)
extern "C" int nvqa_dp_layout(nvqa_model* m, int32_t* whole_vector) {
  NVQA_CHECK(m && whole_vector, "null argument");
  *whole_vector = (m->dp_whole_vector || m->cfg.arch != 1) ? 1 : 0;
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
  static int bulk = -1;
  if (bulk < 0) { const char* e = getenv("NVQA_DP_BULK"); bulk = e ? atoi(e) : 1; }
  if (bulk) {
    // persistent CTAs, one per SM (side: side_ctas), ring of 2..6 stages of (world + 2) x 8 KB
    static int chunk4 = -1;
    if (chunk4 < 0) { const char* e = getenv("NVQA_DP_CHUNK_KB"); chunk4 = e ? std::max(1, std::min(32, atoi(e))) * 64 : 0; }
    // default: 16 KB chunks (one float4 per thread) while three stages of (world + 2) chunks fit, else 8 KB
    const int DPB_CHUNK4 = chunk4 ? chunk4 : ((size_t)(m->dp_world + 2) * 16384 * 3 <= 200 * 1024 ? 1024 : DPB_CHUNK4_DEFAULT);
    const size_t stage_bytes = (size_t)(m->dp_world + 2) * DPB_CHUNK4 * 16;
    int stages = (int)std::min<size_t>(8, (200 * 1024) / stage_bytes);
    if (stages < 2) stages = 2;
    const size_t bsmem = stages * stage_bytes + 64;
    static bool battr = false;
    if (!battr) {
      NVQA_CUDA(cudaFuncSetAttribute(dp_bulk_rmsprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      battr = true;
    }
    static int bulk_ctas = -1;
    if (bulk_ctas < 0) { const char* e = getenv("NVQA_DP_BULK_CTAS"); bulk_ctas = e ? std::max(1, atoi(e)) : 148; }
    static int dbgmode = -1;         // timing experiments only (tools/dp_probe.py): 1 no peer loads, 2 no peer stores
    if (dbgmode < 0) { const char* e = getenv("NVQA_DP_DBG"); dbgmode = e ? atoi(e) : 0; }
    const long long nchunks = ceil_div(n4, (long long)DPB_CHUNK4);
    const int bgrid = (int)std::max<long long>(1, std::min<long long>(nchunks, side ? side_ctas : bulk_ctas));
    dp_bulk_rmsprop_kernel<<<bgrid, DPB_THREADS, bsmem, s>>>(p, m->rms, m->dp_flags, blk0, m->dp_rank, m->dp_world, lo4, n4, step, lr,
                                                    alpha, (float)(1.0 - (double)alpha), eps, wd, clamp, gscale,
                                                    dp_timeout_ns(), stages, DPB_CHUNK4, dbgmode);
    NVQA_LAUNCHED();
    return 0;
  }
  static int ldmode = -1;
  if (ldmode < 0) { const char* e = getenv("NVQA_DP_LD"); ldmode = e ? atoi(e) : 0; }
  dp_fused_rmsprop_kernel<<<grid, 256, smem, s>>>(p, m->rms, m->dp_flags, blk0, m->dp_rank, m->dp_world, lo4, n4, step, lr, alpha,
                                                 (float)(1.0 - (double)alpha), eps, wd, clamp, gscale, dp_timeout_ns(), ldmode);
  NVQA_LAUNCHED();
  return 0;
}
// the exchanges of one step: arch 1 = {encoder + embedding (one range, one lr_scale), multimodal}; other architectures:
// the whole flat vector at once
static int dp_tail_ranges(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp, bool with_multimodal) {
  if (m->cfg.arch != 1) return dp_range(m, 0, 2, lr, alpha, eps, wd, clamp, false);
  // all three blocks are due and share one gradient scale: ONE kernel (one ready / done flag round instead of two).  The
  // flag channel of a range is its first block, and the per-channel step counters advance independently, so a run may mix
  // this with the two-range form only if it always uses the same form: decided once per connection
  if (m->dp_whole_vector) {
    NVQA_CHECK(with_multimodal, "whole-vector exchange: the multimodal block cannot be exchanged separately");
    NVQA_CHECK(m->lr_scale == 1.0f, "-lr_scale must be set before nvqa_dp_connect (it splits the exchange into two ranges)");
    return dp_range(m, 0, 2, lr, alpha, eps, wd, clamp, false);
  }
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
  const int overlap = m->dp_whole_vector ? 0 : (overlap_env >= 0 ? overlap_env : (m->dp_world >= 4 ? 1 : 0));
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
