// Probe: does tcgen05.mma with the A operand in TMEM (lane = M row, 32-bit column = 2 consecutive K elements) give the
// same D as the shared-memory A operand?  Builds both, prints max |diff| vs a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I novel-vqa_b200/csrc -o tools/probes/probe_tmem_a.bin tools/probes/probe_tmem_a.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include "umma_ptx.cuh"
namespace nvqa { void set_error(const std::string&) {} int64_t g_launches = 0; }
using namespace nvqa;

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* Dss, float* Dts, int swap_halves) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t a0 = base, b0 = base + 16384, bar = base + 16384 + 8192;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 16384 + 8192 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill smem tiles, K-major SWIZZLE_128B: row r at r*128 B, 16-byte chunk c stored at chunk (c ^ (r & 7))
  for (int i = threadIdx.x; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<__nv_bfloat16*>(sm + off) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<__nv_bfloat16*>(sm + 16384 + off) = B[i];
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(slot), 256);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  // A row (32*warp + lane) -> TMEM lanes, columns 128..159: word j = (A[r][2j], A[r][2j+1])
  {
    uint32_t w[32];
    const int r = 32 * warp + lane;
    for (int j = 0; j < 32; ++j) {
      uint32_t lo = __bfloat16_as_ushort(A[r * K + 2 * j]), hi = __bfloat16_as_ushort(A[r * K + 2 * j + 1]);
      w[j] = swap_halves ? (hi | (lo << 16)) : (lo | (hi << 16));
    }
    tmem_st32(tb + ((uint32_t)(32 * warp) << 16) + 128, w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    for (int k = 0; k < 4; ++k)
      umma_f16(tb + 0, make_kmajor_sw128_desc(a0 + k * 32), make_kmajor_sw128_desc(b0 + k * 32), idesc, k > 0);
    for (int k = 0; k < 4; ++k)
      umma_f16_ts(tb + 64, tb + 128 + k * 8, make_kmajor_sw128_desc(b0 + k * 32), idesc, k > 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  float acc[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + half * 32, acc);
    for (int j = 0; j < 32; ++j) Dss[(32 * warp + lane) * N + half * 32 + j] = acc[j];
    tmem_ld32(tb + ((uint32_t)(32 * warp) << 16) + 64 + half * 32, acc);
    for (int j = 0; j < 32; ++j) Dts[(32 * warp + lane) * N + half * 32 + j] = acc[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float *dss, *dts;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dss, M * N * 4); cudaMalloc(&dts, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(dss, 0, M * N * 4); cudaMemset(dts, 0, M * N * 4);
    probe<<<1, 128, 40000>>>(dA, dB, dss, dts, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
    std::vector<float> hss(M * N), hts(M * N);
    cudaMemcpy(hss.data(), dss, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hts.data(), dts, M * N * 4, cudaMemcpyDeviceToHost);
    double ess = 0, ets = 0;
    for (int i = 0; i < M * N; ++i) { ess = fmax(ess, fabs(hss[i] - ref[i])); ets = fmax(ets, fabs(hts[i] - ref[i])); }
    printf("swap_halves=%d: max|D_ss - ref| = %.3e   max|D_ts - ref| = %.3e   (ref max %.3f)\n", swap, ess, ets, 30.0);
  }
  return 0;
}
