// Probe: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 per CTA pair = 128 rows per SM, K = 16) as a function of N
// and of the A operand's home (shared memory "SS" / tensor memory "TS").  In pair mode every CTA supplies its own 128 rows
// of A and HALF of the B rows (N/2); the leader CTA issues for both.  Companion of probe_mma_rate.cu (cta_group::1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I novel-vqa_b200/csrc -o tools/probes/probe_mma_rate2.bin tools/probes/probe_mma_rate2.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include "umma_ptx.cuh"
namespace nvqa { void set_error(const std::string&) {} int64_t g_launches = 0; }
using namespace nvqa;

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe(long long* out, int N, int ts, int count) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t a0 = base, b0 = base + 16384, bar = base + 16384 + 16384;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 32768 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  {
    uint32_t w[32];
    for (int j = 0; j < 32; ++j) w[j] = 0x3c003c00u;
    tmem_st32(tb + ((uint32_t)(32 * warp) << 16) + 384, w);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, N, false, false);
      const uint64_t da = make_kmajor_sw128_desc(a0), db = make_kmajor_sw128_desc(b0);
      __syncwarp();
      t0 = clock64();
      if (elect_one_sync()) {
        for (int i = 0; i < count; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ts) umma2_f16_ts(tb, tb + 384 + k * 8, db + (uint64_t)(k * 2), idesc);
            else umma2_f16(tb, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc);
          }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
      t1 = clock64();
    }
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int count = 192;
  printf("cycles per tcgen05.mma.cta_group::2 kind::f16 M=256 (128 rows per SM) K=16 (chain of %d, one accumulator)\n", count);
  for (int grid : {2, 128})
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {32, 64, 128, 256}) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<grid, 128, 40000>>>(d, N, ts, count);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("grid=%d ts=%d N=%d: CUDA error %s\n", grid, ts, N, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("grid=%3d %s N=%3d: issue %6.1f  complete %6.1f cycles/MMA (per-SM tensor floor N/2 = %d)\n", grid, ts ? "TS" : "SS", N,
               (double)h[0] / count, (double)h[1] / count, N / 2);
      }
  return 0;
}
