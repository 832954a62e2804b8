// Probe: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of the A operand's home (shared memory
// "SS" / tensor memory "TS") and of the number of independent accumulators, measured from the first issue to the
// tcgen05.commit arrival of a chain of 192 MMAs issued back to back by one elected lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I novel-vqa_b200/csrc -o tools/probes/probe_mma_rate.bin tools/probes/probe_mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include "umma_ptx.cuh"
namespace nvqa { void set_error(const std::string&) {} int64_t g_launches = 0; }
using namespace nvqa;

__global__ void __launch_bounds__(128) probe(long long* out, int N, int ts, int nacc, int count, int bstack) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t a0 = base, b0 = base + 16384, bar = base + 16384 + 32768;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 16384 + 32768 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(slot), 512);
  }
  fence_proxy_async();
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
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t da = make_kmajor_sw128_desc(a0), db = make_kmajor_sw128_desc(b0);
    __syncwarp();
    const long long t0 = clock64();
    if (elect_one_sync()) {
      for (int i = 0; i < count; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = tb + (uint32_t)(((i + k) % nacc) * (nacc > 1 ? N : 0));
          if (ts) umma_f16_ts(acc, tb + 384 + k * 8, db + (uint64_t)(k * 2), idesc, 1u);
          else umma_f16(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int count = 192;
  printf("cycles per tcgen05.mma kind::f16 M=128 K=16 (chain of %d, one accumulator unless stated)\n", count);
  for (int grid : {1, 128})
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {16, 32, 64, 128, 256})
        for (int nacc : {1, 3}) {
          if (nacc * N > 256) continue;
          long long h[2] = {0, 0};
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 128, 60000>>>(d, N, ts, nacc, count, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid=%3d %s N=%3d accumulators=%d: issue %6.1f  complete %6.1f cycles/MMA (floor N/2 = %d)\n", grid, ts ? "TS" : "SS", N, nacc,
                 (double)h[0] / count, (double)h[1] / count, N / 2);
        }
  return 0;
}
