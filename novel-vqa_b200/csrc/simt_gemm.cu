// fp32 FFMA GEMM (precision mode NVQA_PREC_FP32_SIMT): the exact-fp32 CUDA path that the tcgen05
// engine is cross-checked against on the device.  Register-tiled, shared-memory staged, 128-bit
// loads/stores.  Replaces the cuBLAS sgemm calls behind nn.Linear in the reference
// (002_train_vqa_arch1/misc/LSTM.lua:41-42, misc/netdef.lua:10-11, 002_train_baseline.lua:154).
#include "common.cuh"

namespace nvqa {

constexpr int SG_BK = 16;

template <int BM, int BN, int TM, int TN, bool AK, bool BKM>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, int beta, const float* __restrict__ bias0,
             const float* __restrict__ bias1, int vecA, int vecB, int vecC) {
  __shared__ __align__(16) float As[SG_BK][BM + 4];
  __shared__ __align__(16) float Bs[SG_BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // ---- stage A tile [BM x 16] as As[k][m] ----
    if (AK) {
#pragma unroll
      for (int r = 0; r < BM * SG_BK / 4 / 256; ++r) {
        int i = tid + r * 256;
        int row = i / 4, kq = (i % 4) * 4;
        int gm = m0 + row, gk = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gm < M) {
          const float* p = A + (size_t)gm * lda + gk;
          if (vecA && gk < K) {
            float4 t = *reinterpret_cast<const float4*>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (gk + j < K) v[j] = p[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[kq + j][row] = v[j];
      }
    } else {
#pragma unroll
      for (int r = 0; r < BM * SG_BK / 4 / 256; ++r) {
        int i = tid + r * 256;
        int k = i / (BM / 4), mq = (i % (BM / 4)) * 4;
        int gm = m0 + mq, gk = k0 + k;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gk < K) {
          const float* p = A + (size_t)gk * lda + gm;
          if (vecA && gm < M) t = *reinterpret_cast<const float4*>(p);
          else {
            if (gm + 0 < M) t.x = p[0];
            if (gm + 1 < M) t.y = p[1];
            if (gm + 2 < M) t.z = p[2];
            if (gm + 3 < M) t.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&As[k][mq]) = t;
      }
    }
    // ---- stage B tile [BN x 16] as Bs[k][n] ----
    if (BKM) {
#pragma unroll
      for (int r = 0; r < BN * SG_BK / 4 / 256; ++r) {
        int i = tid + r * 256;
        int row = i / 4, kq = (i % 4) * 4;
        int gn = n0 + row, gk = k0 + kq;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gn < N) {
          const float* p = B + (size_t)gn * ldb + gk;
          if (vecB && gk < K) {
            float4 t = *reinterpret_cast<const float4*>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (gk + j < K) v[j] = p[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[kq + j][row] = v[j];
      }
    } else {
#pragma unroll
      for (int r = 0; r < BN * SG_BK / 4 / 256; ++r) {
        int i = tid + r * 256;
        int k = i / (BN / 4), nq = (i % (BN / 4)) * 4;
        int gn = n0 + nq, gk = k0 + k;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gk < K) {
          const float* p = B + (size_t)gk * ldb + gn;
          if (vecB && gn < N) t = *reinterpret_cast<const float4*>(p);
          else {
            if (gn + 0 < N) t.x = p[0];
            if (gn + 1 < N) t.y = p[1];
            if (gn + 2 < N) t.z = p[2];
            if (gn + 3 < N) t.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[k][nq]) = t;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: + bias, (+ C), 128-bit stores where aligned ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float* cp = C + (size_t)gm * ldc + gn;
      float r[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (gn + q < N) {
          if (bias0) r[q] += bias0[gn + q];
          if (bias1) r[q] += bias1[gn + q];
        }
      }
      if (vecC && gn + 3 < N) {
        if (beta) {
          float4 o = *reinterpret_cast<const float4*>(cp);
          r[0] += o.x; r[1] += o.y; r[2] += o.z; r[3] += o.w;
        }
        *reinterpret_cast<float4*>(cp) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (gn + q < N) cp[q] = (beta ? cp[q] : 0.f) + r[q];
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
static int launch_sgemm(cudaStream_t s, bool ak, bool bk, int M, int N, int K, const float* A, int lda,
                        const float* B, int ldb, float* C, int ldc, bool beta, const float* b0, const float* b1) {
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  int vecA = al16(A) && lda % 4 == 0 && (ak ? K % 4 == 0 : M % 4 == 0);
  int vecB = al16(B) && ldb % 4 == 0 && (bk ? K % 4 == 0 : N % 4 == 0);
  int vecC = al16(C) && ldc % 4 == 0;
#define NVQA_SG(AK, BKM)                                                                               \
  sgemm_kernel<BM, BN, TM, TN, AK, BKM><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, beta ? 1 : 0, \
                                                             b0, b1, vecA, vecB, vecC)
  if (ak && bk) NVQA_SG(true, true);
  else if (ak && !bk) NVQA_SG(true, false);
  else if (!ak && bk) NVQA_SG(false, true);
  else NVQA_SG(false, false);
#undef NVQA_SG
  NVQA_LAUNCHED();
  return 0;
}

int simt_gemm(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const float* A, int lda,
              const float* B, int ldb, float* C, int ldc, bool beta, const float* bias0, const float* bias1) {
  if (M <= 0 || N <= 0) return 0;
  // big problems: 128x128 tiles (8x8 per thread); small ones: 64x64 so that >= 1 wave of 148 SMs exists
  long tiles128 = (long)ceil_div(M, 128) * ceil_div(N, 128);
  if (tiles128 >= 2 * 148)
    return launch_sgemm<128, 128, 8, 8>(s, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, beta, bias0, bias1);
  return launch_sgemm<64, 64, 4, 4>(s, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, beta, bias0, bias1);
}

}  // namespace nvqa
