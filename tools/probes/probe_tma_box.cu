// Probe: what does a cp.async.bulk.tensor box cost the issuing warp, and
// how long until a half h-tile has landed, as a function of the box shape?  The persistent LSTM kernels measured ~316
// cycles per issued box whatever its size (8 boxes = 2,500 cycles before the last one is even issued).
// Variants over the SAME 32 KB of global memory ([2 planes][16 rows][512 bf16], row pitch 1 KB, as the h planes):
//   A  8 x 3-D boxes (64 cols, 16 rows, 2 planes)           = what lstm_fwd_v2_kernel<PAIR> issues per sub-tile
//   B  2 x 4-D boxes (64 cols, 16 rows, 2 planes, 4 k-blocks) over a 4-D view (k-block stride 128 B)
//   C  1 x 4-D box  (64 cols, 16 rows, 2 planes, 8 k-blocks)
// All land as [k-block][plane][row][128 B], SWIZZLE_128B, i.e. the layout the MMA descriptors of the kernel expect.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I novel-vqa_b200/csrc -o tools/probes/probe_tma_box.bin tools/probes/probe_tma_box.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "umma_ptx.cuh"
namespace nvqa { void set_error(const std::string&) {} int64_t g_launches = 0; }
using namespace nvqa;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int ROWS = 16, COLS = 512, PLANES = 2, KB = 8;
constexpr uint32_t TILE_BYTES = ROWS * COLS * 2 * PLANES;        // 32 KB

__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap map3, const __grid_constant__ CUtensorMap map4,
                                            long long* out, unsigned int* check, int variant, int row0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base + TILE_BYTES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map4) : "memory");
  }
  __syncthreads();
  if (warp == 0) {
    long long t_issue = 0;
    const long long t0 = clock64();
    if (elect_one_sync()) {
      mbar_expect_tx(bar, TILE_BYTES);
      if (variant == 0) {
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(base + kb * 4096, &map3, bar, kb * 64, row0, 0);
      } else if (variant == 1) {
        for (int h = 0; h < 2; ++h) tma_load_4d(base + h * 16384, &map4, bar, 0, row0, 0, 4 * h);
      } else {
        tma_load_4d(base, &map4, bar, 0, row0, 0, 0);
      }
      t_issue = clock64() - t0;
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64() - t0;
    t_issue = __shfl_sync(0xffffffffu, t_issue, __ffs(__activemask()) - 1);     // (the elected lane is lane 0 in practice)
    if (lane == 0 && blockIdx.x == 0) { out[0] = t_issue; out[1] = t1; }
    // checksum of the landed tile: the variants must agree (same bytes at the same shared-memory addresses)
    unsigned int s = 0;
    for (uint32_t i = lane; i < TILE_BYTES / 4; i += 32) s = s * 31u + reinterpret_cast<const unsigned int*>(smem_raw + (base - raw))[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && blockIdx.x == 0) *check = s;
  }
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  const int total_rows = 27 * 512;                                  // (T + 1) x B rows of one plane, as the h planes
  const size_t plane = (size_t)total_rows * COLS;
  __nv_bfloat16* d;
  cudaMalloc(&d, plane * PLANES * 2);
  std::vector<unsigned short> h(plane * PLANES);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned short)(i * 2654435761u >> 16);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap m3, m4;
  {
    cuuint64_t dims[3] = {COLS, (cuuint64_t)total_rows, PLANES};
    cuuint64_t strides[2] = {COLS * 2, plane * 2};
    cuuint32_t box[3] = {64, ROWS, PLANES}, es[3] = {1, 1, 1};
    CUresult r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("3-D encode failed: %d\n", (int)r); return 1; }
  }
  {
    // 4-D view of the same memory: (col within k-block, row, plane, k-block); the k-block stride (128 B) is SMALLER than
    // the row stride -- whether the driver accepts that is the first thing this probe answers
    cuuint64_t dims[4] = {64, (cuuint64_t)total_rows, PLANES, KB};
    cuuint64_t strides[3] = {COLS * 2, plane * 2, 128};
    cuuint32_t box[4] = {64, ROWS, PLANES, 4}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("4-D encode (4 k-blocks per box) failed: %d -- the 4-D view is not encodable\n", (int)r); return 2; }
  }
  CUtensorMap m4full = m4;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)total_rows, PLANES, KB};
    cuuint64_t strides[3] = {COLS * 2, plane * 2, 128};
    cuuint32_t box[4] = {64, ROWS, PLANES, KB}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m4full, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("4-D encode (8 k-blocks per box) failed: %d\n", (int)r); return 2; }
  }
  long long* out;
  unsigned int* chk;
  cudaMalloc(&out, 16);
  cudaMalloc(&chk, 4);
  const int smem = TILE_BYTES + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"8 x 3-D box (4 KB)", "2 x 4-D box (16 KB)", "1 x 4-D box (32 KB)"};
  for (int grid : {1, 128})
    for (int v = 0; v < 3; ++v) {
      long long hres[2] = {0, 0};
      unsigned int hc = 0;
      for (int rep = 0; rep < 3; ++rep) {                             // the last repetition reads L2-resident data
        probe<<<grid, 64, smem>>>(m3, v == 2 ? m4full : m4, out, chk, v, 1000);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", v, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(hres, out, 16, cudaMemcpyDeviceToHost);
      cudaMemcpy(&hc, chk, 4, cudaMemcpyDeviceToHost);
      printf("grid=%3d %-20s: issue %6lld cycles, all 32 KB landed after %6lld cycles, checksum %08x\n", grid, names[v], hres[0], hres[1], hc);
    }
  return 0;
}
