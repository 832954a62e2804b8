// Bandwidth-bound kernels of the text-autoencoder step (001_train_autoencoder/): shared LookupTable
// -> Dropout(0.5) -> Tanh gather / scatter for the encoder and decoder clones, nn.LanguageModelCriterion target
// selection, LogSoftMax + criterion over the (V+1)-wide vocabulary rows, and the clamp + weight-decay + Adam update.
// Same conventions as pointwise.cu: coalesced, 128-bit vectorised, warp-shuffle reduced, deterministic reductions.
//
// Row layout of the step buffers: rows [0, tmax*B) are the encoder steps, rows [tmax*B, (2*tmax+1)*B) the decoder steps
// (time-major [t][b]); seq is [B x T], zero-padded on the right.
#include <algorithm>

#include "pointwise.cuh"
#include "umma_ptx.cuh"

namespace nvqa {

#define LD4(p) (*reinterpret_cast<const float4*>(p))
#define ST4(p, v) (*reinterpret_cast<float4*>(p) = (v))

// token fed at row n and the index of its Dropout element 0 (misc/AutoEncoder_text_nostart.lua:244-266, 296-327)
__device__ __forceinline__ int ae_token(const int32_t* __restrict__ seq, int64_t n, int B, int T, int V, int tmax, bool* is_dec,
                                        int64_t* mrow) {
  const int b = (int)(n % B), t = (int)(n / B);
  int tok;
  if (t < tmax) {
    *is_dec = false; *mrow = n;
    tok = seq[(int64_t)b * T + t];
  } else {
    const int td = t - tmax;
    *is_dec = true; *mrow = (int64_t)td * B + b;
    tok = td == 0 ? V + 1 : seq[(int64_t)b * T + td - 1];         // START, then the sequence shifted by one
  }
  return (tok < 1 || tok > V + 1) ? 1 : tok;                      // nulls -> token 1 (:258-266)
}

__global__ void __launch_bounds__(256)
ae_embed_fwd_kernel(const int32_t* __restrict__ seq, const float* __restrict__ table, float* __restrict__ y, Drop denc,
                    Drop ddec, int B, int T, int E, int V, int tmax) {
  pdl_entry();
  const int E4 = E >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(2 * tmax + 1) * B * E4) return;
  const int e = (int)(i % E4) * 4;
  const int64_t n = i / E4;
  bool dec; int64_t mrow;
  const int tok = ae_token(seq, n, B, T, V, tmax, &dec, &mrow);
  const float4 w = LD4(table + (int64_t)(tok - 1) * E + e);
  const float4 m = drop_at4(dec ? ddec : denc, (uint64_t)mrow * E + e);
  ST4(y + n * E + e, make_float4(tanhf(m.x * w.x), tanhf(m.y * w.y), tanhf(m.z * w.z), tanhf(m.w * w.w)));
}

int ae_embed_fwd(cudaStream_t s, const int32_t* seq, const float* table, float* y, Drop denc, Drop ddec, int B, int T, int E,
                 int V, int tmax) {
  const int64_t total = (int64_t)(2 * tmax + 1) * B * (E / 4);
  NVQA_CUDA(launch_pdl(ae_embed_fwd_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, seq, table, y, denc, ddec, B, T, E, V, tmax));
  NVQA_LAUNCHED();
  return 0;
}

// Tanh / Dropout backward + LookupTable accGradParameters of every encoder and decoder clone (:362-364, 384-386)
__global__ void __launch_bounds__(256)
ae_embed_bwd_kernel(const int32_t* __restrict__ seq, const float* __restrict__ y, const float* __restrict__ dx,
                    float* __restrict__ dtable, Drop denc, Drop ddec, int B, int T, int E, int V, int tmax) {
  pdl_entry();
  const int E4 = E >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(2 * tmax + 1) * B * E4) return;
  const int e = (int)(i % E4) * 4;
  const int64_t n = i / E4;
  bool dec; int64_t mrow;
  const int tok = ae_token(seq, n, B, T, V, tmax, &dec, &mrow);
  const float4 g = LD4(dx + n * E + e), yy = LD4(y + n * E + e), m = drop_at4(dec ? ddec : denc, (uint64_t)mrow * E + e);
  if (tok == 1 || tok == V + 1) return;          // the two HOT rows are summed by ae_embed_bwd_hot_kernel
  float* dst = dtable + (int64_t)(tok - 1) * E + e;
  atomicAdd(dst + 0, g.x * (1.0f - yy.x * yy.x) * m.x);
  atomicAdd(dst + 1, g.y * (1.0f - yy.y * yy.y) * m.y);
  atomicAdd(dst + 2, g.z * (1.0f - yy.z * yy.z) * m.z);
  atomicAdd(dst + 3, g.w * (1.0f - yy.w * yy.w) * m.w);
}

// Token 1 (every null position is fed as token 1, :258-266: ~37 % of the positions of a book-corpus batch) and the START
// token V + 1 (all B rows of the first decoder step) are the targets of thousands of scatter-adds each: as atomics they
// serialise on E addresses (measured: the scatter kernel took 570 us at config 5, the rest of the vocabulary ~50 us).  Their
// gradient rows are column sums over the rows that feed them: CTA = 32 columns x 8 row lanes over a chunk of rows, one
// atomic per (column, chunk, hot token).
__global__ void __launch_bounds__(256)
ae_embed_bwd_hot_kernel(const int32_t* __restrict__ seq, const float* __restrict__ y, const float* __restrict__ dx,
                        float* __restrict__ dtable, Drop denc, Drop ddec, int B, int T, int E, int V, int tmax,
                        int rows_per_chunk) {
  pdl_entry();
  __shared__ float red[2][8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int64_t rows = (int64_t)(2 * tmax + 1) * B;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float a1 = 0.f, as = 0.f;
  if (c < E) {
    for (int64_t n = r0 + ry; n < r1; n += 8) {
      bool dec; int64_t mrow;
      const int tok = ae_token(seq, n, B, T, V, tmax, &dec, &mrow);
      if (tok != 1 && tok != V + 1) continue;
      const float g = dx[n * E + c], yy = y[n * E + c], m = drop_at(dec ? ddec : denc, (uint64_t)mrow * E + c);
      const float v = g * (1.0f - yy * yy) * m;
      if (tok == 1) a1 += v; else as += v;
    }
  }
  red[0][ry][threadIdx.x & 31] = a1;
  red[1][ry][threadIdx.x & 31] = as;
  __syncthreads();
  if (ry == 0 && c < E) {
    float t1 = 0.f, ts = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t1 += red[0][k][threadIdx.x & 31]; ts += red[1][k][threadIdx.x & 31]; }
    if (t1 != 0.f) atomicAdd(dtable + c, t1);
    if (ts != 0.f) atomicAdd(dtable + (int64_t)V * E + c, ts);
  }
}

int ae_embed_bwd(cudaStream_t s, const int32_t* seq, const float* y, const float* dx, float* dtable, Drop denc, Drop ddec,
                 int B, int T, int E, int V, int tmax) {
  const int64_t total = (int64_t)(2 * tmax + 1) * B * (E / 4);
  NVQA_CUDA(launch_pdl(ae_embed_bwd_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, seq, y, dx, dtable, denc, ddec, B, T, E, V, tmax));
  NVQA_LAUNCHED();
  const int64_t rows = (int64_t)(2 * tmax + 1) * B;
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(64, rows / 64));
  dim3 grid(ceil_div(E, 32), chunks);
  NVQA_CUDA(launch_pdl(ae_embed_bwd_hot_kernel, dim3(grid), dim3(256), 0, s, seq, y, dx, dtable, denc, ddec, B, T, E, V, tmax, ceil_div(rows, chunks)));
  NVQA_LAUNCHED();
  return 0;
}

// nn.LanguageModelCriterion target selection (:427-447): targets [(T+1) x B] (0 = no prediction), n_pred += count
__global__ void __launch_bounds__(256)
lm_targets_kernel(const int32_t* __restrict__ seq, int32_t* __restrict__ targets, int32_t* __restrict__ n_pred, int B, int T, int V) {
  pdl_entry();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (b < B) {
    bool first = true;
    for (int t = 0; t <= T; ++t) {
      int ti = t < T ? seq[(int64_t)b * T + t] : 0;
      if (ti == 0 && first) { ti = V + 1; first = false; }
      targets[(int64_t)t * B + b] = ti;
      cnt += ti != 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_pred, cnt);
}

int lm_targets(cudaStream_t s, const int32_t* seq, int32_t* targets, int32_t* n_pred, int B, int T, int V) {
  NVQA_CUDA(cudaMemsetAsync(n_pred, 0, sizeof(int32_t), s));
  NVQA_CUDA(launch_pdl(lm_targets_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, s, seq, targets, n_pred, B, T, V));
  NVQA_LAUNCHED();
  return 0;
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// nn.LogSoftMax over one vocabulary row per CTA, in place (003_train_vqa_arch2/misc/LSTM_decoder.lua:58-59), and the
// criterion's term of that row: rowloss = -logp[target] (0 when the row predicts nothing).
__global__ void __launch_bounds__(256)
logsoftmax_lm_kernel(float* __restrict__ x, int ld, int ncols, const int32_t* __restrict__ targets, float* __restrict__ rowloss) {
  pdl_entry();
  __shared__ float red[8];
  float* r = x + (int64_t)blockIdx.x * ld;
  const int n4 = ncols >> 2;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n4; j += 256) {
    const float4 v = LD4(r + 4 * j);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int j = 4 * n4 + threadIdx.x; j < ncols; j += 256) mx = fmaxf(mx, r[j]);
  mx = block_reduce(mx, true, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < n4; j += 256) {
    const float4 v = LD4(r + 4 * j);
    se += (expf(v.x - mx) + expf(v.y - mx)) + (expf(v.z - mx) + expf(v.w - mx));
  }
  for (int j = 4 * n4 + threadIdx.x; j < ncols; j += 256) se += expf(r[j] - mx);
  se = block_reduce(se, false, red);
  const float lse = mx + logf(se);
  const int tg = targets ? targets[blockIdx.x] : 0;
  if (threadIdx.x == 0 && rowloss) rowloss[blockIdx.x] = (tg >= 1 && tg <= ncols) ? -(r[tg - 1] - lse) : 0.f;
  __syncthreads();                                   // the target's logit was read before it is overwritten
  for (int j = threadIdx.x; j < n4; j += 256) {
    float4 v = LD4(r + 4 * j);
    v.x -= lse; v.y -= lse; v.z -= lse; v.w -= lse;
    ST4(r + 4 * j, v);
  }
  for (int j = 4 * n4 + threadIdx.x; j < ncols; j += 256) r[j] -= lse;
}

int logsoftmax_lm(cudaStream_t s, float* x, int rows, int ld, int ncols, const int32_t* targets, float* rowloss) {
  if (rows <= 0) return 0;
  NVQA_CUDA(launch_pdl(logsoftmax_lm_kernel, dim3(rows), dim3(256), 0, s, x, ld, ncols, targets, rowloss));
  NVQA_LAUNCHED();
  return 0;
}

// The same WITHOUT rewriting the row (tensor-core modes): lse[row] = log sum exp and the criterion's term; the 1.36 GB of
// logits stay as they are, log-probs are formed where they are consumed (lm_grad_planes, lm_logprobs).
__global__ void __launch_bounds__(256)
lm_row_stats_kernel(const float* __restrict__ x, int ld, int ncols, const int32_t* __restrict__ targets, float* __restrict__ lse_out,
                    float* __restrict__ rowloss) {
  pdl_entry();
  __shared__ float red[8];
  const float* r = x + (int64_t)blockIdx.x * ld;
  const int n4 = ncols >> 2;
  // one pass: running max and rescaled sum per thread (online softmax), combined across the CTA
  float mx = -INFINITY, se = 0.f;
  for (int j = threadIdx.x; j < n4; j += 256) {
    const float4 v = LD4(r + 4 * j);
    const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (m4 > mx) { se *= expf(mx - m4); mx = m4; }
    se += (expf(v.x - mx) + expf(v.y - mx)) + (expf(v.z - mx) + expf(v.w - mx));
  }
  for (int j = 4 * n4 + threadIdx.x; j < ncols; j += 256) {
    const float v = r[j];
    if (v > mx) { se *= expf(mx - v); mx = v; }
    se += expf(v - mx);
  }
  const float gmx = block_reduce(mx, true, red);
  se = mx == -INFINITY ? 0.f : se * expf(mx - gmx);
  se = block_reduce(se, false, red);
  const float lse = gmx + logf(se);
  if (threadIdx.x == 0) {
    lse_out[blockIdx.x] = lse;
    const int tg = targets ? targets[blockIdx.x] : 0;
    if (rowloss) rowloss[blockIdx.x] = (tg >= 1 && tg <= ncols) ? -(r[tg - 1] - lse) : 0.f;
  }
}

int lm_row_stats(cudaStream_t s, const float* x, int rows, int ld, int ncols, const int32_t* targets, float* lse, float* rowloss) {
  if (rows <= 0) return 0;
  NVQA_CUDA(launch_pdl(lm_row_stats_kernel, dim3(rows), dim3(256), 0, s, x, ld, ncols, targets, lse, rowloss));
  NVQA_LAUNCHED();
  return 0;
}

// log-probs of nrows rows (logits - lse) into a dense [nrows x ncols] buffer (nvqa_logprobs_get)
__global__ void __launch_bounds__(256)
lm_logprobs_kernel(const float* __restrict__ x, const float* __restrict__ lse, int ld, int ncols, float* __restrict__ out) {
  pdl_entry();
  const int row = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncols) out[(int64_t)row * ncols + j] = x[(int64_t)row * ld + j] - lse[row];
}

int lm_logprobs(cudaStream_t s, const float* x, const float* lse, int nrows, int ld, int ncols, float* out) {
  if (nrows <= 0) return 0;
  dim3 grid(ceil_div(ncols, 256), nrows);
  NVQA_CUDA(launch_pdl(lm_logprobs_kernel, dim3(grid), dim3(256), 0, s, x, lse, ld, ncols, out));
  NVQA_LAUNCHED();
  return 0;
}

// criterion + LogSoftMax backward straight into the operand format of the two vocabulary GEMMs of the backward pass:
// d logits = (softmax - onehot(target)) / n as bf16 PLANES [P][rows][pitch] (pitch = ncols rounded up to 8, zero padded) --
// the fp32 gradient tensor, its separate split pass and the separate column-sum pass of the bias gradient (3 x 1.36 GB
// read, 2 x 1.36 GB written at config 5) do not exist.  CTA = 1024 columns x RPC rows: a thread keeps its 4 columns,
// accumulates their sums over the rows in registers and adds them to gbias with one atomic per column and CTA.
constexpr int LMG_RPC = 64;
__global__ void __launch_bounds__(256)
lm_grad_planes_kernel(const float* __restrict__ x, const float* __restrict__ lse, int rows, int ld, int ncols,
                      const int32_t* __restrict__ targets, const int32_t* __restrict__ n_pred, float gscale,
                      __nv_bfloat16* __restrict__ planes, int pitch, long long plane_stride, int P, float* __restrict__ gbias) {
  pdl_entry();
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= pitch) return;
  const int r0 = blockIdx.y * LMG_RPC, r1 = min(rows, r0 + LMG_RPC);
  const float inv = gscale / (float)max(*n_pred, 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool full = j + 3 < ncols;
  constexpr int RU = 4;                                   // rows in flight per thread: the loop is latency-bound otherwise
  for (int rb = r0; rb < r1; rb += RU) {
    float4 a[RU];
    int tg[RU];
    float l[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int row = rb + u;
      tg[u] = row < r1 ? targets[row] - 1 : -1;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      l[u] = 0.f;
      if (tg[u] >= 0) {
        l[u] = lse[row];
        const float* r = x + (int64_t)row * ld + j;
        if (full) a[u] = LD4(r);
        else {
          if (j + 0 < ncols) a[u].x = r[0];
          if (j + 1 < ncols) a[u].y = r[1];
          if (j + 2 < ncols) a[u].z = r[2];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int row = rb + u;
      if (row >= r1) break;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tg[u] >= 0) {
        if (j + 0 < ncols) v.x = (expf(a[u].x - l[u]) - (tg[u] == j + 0 ? 1.f : 0.f)) * inv;
        if (j + 1 < ncols) v.y = (expf(a[u].y - l[u]) - (tg[u] == j + 1 ? 1.f : 0.f)) * inv;
        if (j + 2 < ncols) v.z = (expf(a[u].z - l[u]) - (tg[u] == j + 2 ? 1.f : 0.f)) * inv;
        if (j + 3 < ncols) v.w = (expf(a[u].w - l[u]) - (tg[u] == j + 3 ? 1.f : 0.f)) * inv;
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const float xs[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 pl[3][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split3(xs[k], pl[0][k], pl[1][k], pl[2][k]);
      for (int q = 0; q < P; ++q) {
        uint2 o;
        o.x = (uint32_t)__bfloat16_as_ushort(pl[q][0]) | ((uint32_t)__bfloat16_as_ushort(pl[q][1]) << 16);
        o.y = (uint32_t)__bfloat16_as_ushort(pl[q][2]) | ((uint32_t)__bfloat16_as_ushort(pl[q][3]) << 16);
        *reinterpret_cast<uint2*>(planes + (size_t)q * plane_stride + (size_t)row * pitch + j) = o;
      }
    }
  }
  if (gbias) {
    if (j + 0 < ncols && acc.x != 0.f) atomicAdd(gbias + j + 0, acc.x);
    if (j + 1 < ncols && acc.y != 0.f) atomicAdd(gbias + j + 1, acc.y);
    if (j + 2 < ncols && acc.z != 0.f) atomicAdd(gbias + j + 2, acc.z);
    if (j + 3 < ncols && acc.w != 0.f) atomicAdd(gbias + j + 3, acc.w);
  }
}

int lm_grad_planes(cudaStream_t s, const float* x, const float* lse, int rows, int ld, int ncols, const int32_t* targets,
                   const int32_t* n_pred, float gscale, __nv_bfloat16* planes, int pitch, long long plane_stride, int P,
                   float* gbias) {
  if (rows <= 0) return 0;
  dim3 grid(ceil_div(pitch / 4, 256), ceil_div(rows, LMG_RPC));
  NVQA_CUDA(launch_pdl(lm_grad_planes_kernel, dim3(grid), dim3(256), 0, s, x, lse, rows, ld, ncols, targets, n_pred, gscale, planes, pitch, plane_stride, P, gbias));
  NVQA_LAUNCHED();
  return 0;
}

// loss = sum(rowloss) / n_pred in a fixed order (deterministic)   (:449)
__global__ void __launch_bounds__(256)
lm_loss_reduce_kernel(const float* __restrict__ rowloss, int rows, const int32_t* __restrict__ n_pred, float* __restrict__ loss) {
  pdl_entry();
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < rows; i += 256) acc += rowloss[i];
  acc = block_reduce(acc, false, red);
  if (threadIdx.x == 0) loss[0] = acc / (float)max(*n_pred, 1);
}

int lm_loss_reduce(cudaStream_t s, const float* rowloss, int rows, const int32_t* n_pred, float* loss) {
  NVQA_CUDA(launch_pdl(lm_loss_reduce_kernel, dim3(1), dim3(256), 0, s, rowloss, rows, n_pred, loss));
  NVQA_LAUNCHED();
  return 0;
}

// criterion backward (-1/n at the target, :444-450) folded through LogSoftMax backward, in place over the log-probs:
// d logits = (softmax - onehot(target)) / n on predicting rows, 0 on the others.
__global__ void __launch_bounds__(256)
lm_grad_kernel(float* __restrict__ lp, int ld, int ncols, const int32_t* __restrict__ targets, const int32_t* __restrict__ n_pred,
               float gscale) {
  pdl_entry();
  const int row = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= ld) return;
  float* r = lp + (int64_t)row * ld + j;
  const int tg = targets[row] - 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tg >= 0) {
    const float inv = gscale / (float)max(*n_pred, 1);
    const float4 l = LD4(r);
    v.x = j + 0 < ncols ? (expf(l.x) - (tg == j + 0 ? 1.f : 0.f)) * inv : 0.f;
    v.y = j + 1 < ncols ? (expf(l.y) - (tg == j + 1 ? 1.f : 0.f)) * inv : 0.f;
    v.z = j + 2 < ncols ? (expf(l.z) - (tg == j + 2 ? 1.f : 0.f)) * inv : 0.f;
    v.w = j + 3 < ncols ? (expf(l.w) - (tg == j + 3 ? 1.f : 0.f)) * inv : 0.f;
  }
  ST4(r, v);
}

int lm_grad(cudaStream_t s, float* lp, int rows, int ld, int ncols, const int32_t* targets, const int32_t* n_pred, float gscale) {
  if (rows <= 0) return 0;
  dim3 grid(ceil_div(ld / 4, 256), rows);
  NVQA_CUDA(launch_pdl(lm_grad_kernel, dim3(grid), dim3(256), 0, s, lp, ld, ncols, targets, n_pred, gscale));
  NVQA_LAUNCHED();
  return 0;
}

// grad_params:clamp(-c, c); grad_params:add(wd, params); adam()   (001_train_arch1_text_autoencoder.lua:237-243,
// misc/optim_updates.lua:78-111: eps outside the sqrt, bias corrections folded into the step size)
__global__ void __launch_bounds__(256)
clamp_adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                  float step, float b1, float omb1, float b2, float omb2, float eps, float wd, float clampv, float gscale) {
  pdl_entry();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 xv = LD4(x + i), gv = LD4(g + i), mv = LD4(m + i), vv = LD4(v + i);
#define UP(k)                                                    \
    { float gg = fminf(fmaxf(gv.k * gscale, -clampv), clampv);   \
      gg += wd * xv.k;                                           \
      mv.k = b1 * mv.k + omb1 * gg;                              \
      vv.k = b2 * vv.k + omb2 * gg * gg;                         \
      xv.k -= step * (mv.k / (sqrtf(vv.k) + eps)); }
    UP(x) UP(y) UP(z) UP(w)
#undef UP
    ST4(x + i, xv); ST4(m + i, mv); ST4(v + i, vv);
  } else {
    for (int64_t k = i; k < n; ++k) {
      float gg = fminf(fmaxf(g[k] * gscale, -clampv), clampv);
      gg += wd * x[k];
      const float mm = b1 * m[k] + omb1 * gg, vk = b2 * v[k] + omb2 * gg * gg;
      m[k] = mm; v[k] = vk;
      x[k] -= step * (mm / (sqrtf(vk) + eps));
    }
  }
}

int clamp_adam(cudaStream_t s, float* x, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
               float eps, float wd, float clamp, float gscale, int64_t t) {
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step = (float)((double)lr * sqrt(bc2) / bc1);
  NVQA_CUDA(launch_pdl(clamp_adam_kernel, dim3(ceil_div(ceil_div(n, 4), 256)), dim3(256), 0, s, x, g, m, v, n, step, beta1, (float)(1.0 - (double)beta1), beta2,
                                                                 (float)(1.0 - (double)beta2), eps, wd, clamp, gscale));
  NVQA_LAUNCHED();
  return 0;
}

}  // namespace nvqa
