// Shared device/host helpers of libnvqa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

namespace nvqa {

// ---- error plumbing: every C-ABI entry point returns int and never throws ---------------------
void set_error(const std::string& msg);
extern int64_t g_launches;                       // kernels launched since load (nvqa_launch_count)

#define NVQA_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      ::nvqa::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")");                                   \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

#define NVQA_CHECK(cond, msg)                                                  \
  do {                                                                         \
    if (!(cond)) {                                                             \
      ::nvqa::set_error(std::string(msg) + " [" #cond "]");                    \
      return 1;                                                                \
    }                                                                          \
  } while (0)

#define NVQA_TRY(expr)            \
  do {                            \
    int r__ = (expr);             \
    if (r__ != 0) return r__;     \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------
// A kernel launched through launch_pdl may be SCHEDULED while its predecessor in the stream is still running (its CTAs
// become resident as the predecessor's exit, block prologues run, launch latency is hidden); it must not touch memory
// before pdl_entry(), which waits until the predecessor grid has completed and its writes are visible, and which then lets
// the next kernel of the stream be scheduled in turn.  pdl_entry() is the FIRST statement of every such kernel (a kernel
// that returned without it could complete before its own predecessor and break the chain); in a kernel launched without
// the attribute both instructions are no-ops.  NVQA_PDL=0 launches everything fully serialised.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_wait(); pdl_trigger(); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// launch-error check used after every kernel launch
#define NVQA_LAUNCHED()                 \
  do {                                  \
    ++::nvqa::g_launches;               \
    NVQA_CUDA(cudaGetLastError());      \
  } while (0)

// ---- counter-hash dropout (bit-identical twin of oracle/rng.py) -----------------------------
enum : uint32_t {
  STREAM_EMB = 1, STREAM_LSTM0 = 2, STREAM_AXB_Q = 16, STREAM_AXB_I = 17, STREAM_HEAD = 18
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t stream_key(uint64_t seed, uint32_t stream) {
  uint32_t lo = (uint32_t)(seed & 0xFFFFFFFFu), hi = (uint32_t)(seed >> 32);
  return mix32(lo ^ mix32(hi + stream * 0x9E3779B1u));
}
// hash word covering elements [4w, 4w+4)
__host__ __device__ __forceinline__ uint32_t keep_word(uint32_t key, uint64_t w) {
  uint32_t wlo = (uint32_t)(w & 0xFFFFFFFFu), whi = (uint32_t)(w >> 32);
  return mix32(wlo * 0x9E3779B1u + mix32(whi ^ key));
}

// How a kernel obtains the Dropout multiplier of element idx.
struct Drop {
  const float* mask;   // explicit multipliers (parity tests) or nullptr
  uint32_t key;        // stream key of the hash generator
  uint32_t thresh;     // round(p*256)
  float scale;         // 1/(1-p)
  int mode;            // 0: identity (evaluate), 1: explicit mask, 2: hash
};
__device__ __forceinline__ float drop_at(const Drop& d, uint64_t idx) {
  if (d.mode == 0) return 1.0f;
  if (d.mode == 1) return d.mask[idx];
  uint32_t h = keep_word(d.key, idx >> 2);
  uint32_t byte = (h >> (8u * (uint32_t)(idx & 3u))) & 0xFFu;
  return byte >= d.thresh ? d.scale : 0.0f;
}
// four consecutive elements, idx % 4 == 0
__device__ __forceinline__ float4 drop_at4(const Drop& d, uint64_t idx) {
  if (d.mode == 0) return make_float4(1.f, 1.f, 1.f, 1.f);
  if (d.mode == 1) return *reinterpret_cast<const float4*>(d.mask + idx);
  uint32_t h = keep_word(d.key, idx >> 2);
  float4 r;
  r.x = ((h) & 0xFFu) >= d.thresh ? d.scale : 0.0f;
  r.y = ((h >> 8) & 0xFFu) >= d.thresh ? d.scale : 0.0f;
  r.z = ((h >> 16) & 0xFFu) >= d.thresh ? d.scale : 0.0f;
  r.w = ((h >> 24) & 0xFFu) >= d.thresh ? d.scale : 0.0f;
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- SIMT fp32 GEMM (simt_gemm.cu) -----------------------------------------------------------
// C[M x N] (ldc) = (beta ? C : 0) + A (.) B + bias0[n] + bias1[n]
//   A stored [M x K] row-major (a_kmajor) or [K x M];  B stored [N x K] (b_kmajor) or [K x N].
int simt_gemm(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const float* A, int lda,
              const float* B, int ldb, float* C, int ldc, bool beta, const float* bias0, const float* bias1);

// ---- tcgen05 GEMM engine (umma_gemm.cu) -------------------------------------------------------
struct UmmaWorkspace;   // bf16 operand planes + tensor maps
// a_static / b_static: cache class of the operand's bf16 planes -- 0 none, 1 weight (until umma_workspace_invalidate()),
// 2 activation written once per forward (until umma_workspace_new_forward())
int umma_gemm(cudaStream_t s, int planes, bool a_kmajor, bool b_kmajor, int M, int N, int K, const float* A,
              int lda, const float* B, int ldb, float* C, int ldc, bool beta, const float* bias0,
              const float* bias1, UmmaWorkspace* ws, int a_static, int b_static);
int umma_workspace_create(UmmaWorkspace** ws, size_t transient_bytes, size_t static_bytes, size_t act_bytes = 0);
void umma_workspace_new_forward(UmmaWorkspace* ws);
void umma_workspace_destroy(UmmaWorkspace* ws);
void umma_workspace_invalidate(UmmaWorkspace* ws);

}  // namespace nvqa
