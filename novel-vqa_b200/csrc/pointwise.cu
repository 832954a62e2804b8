// Bandwidth-bound kernels of the arch1 step: coalesced, 128-bit vectorised, warp-shuffle reduced.
// Grids are sized to cover the data once (element-wise kernels) with 256-thread CTAs.
// Layout: activations are time-major padded [T][B][feat]; a row (t,b) is ACTIVE iff
// t >= T - len[b] (questions are right-aligned, misc/RNNUtils.lua:54-61).  Inactive rows keep the
// zero initial state and produce zero gradients, which reproduces the reference's length-sorted
// packed recurrence (misc/RNNUtils.lua:128-211) without sorting.
#include <algorithm>

#include "pointwise.cuh"
#include "umma_ptx.cuh"

namespace nvqa {

#define LD4(p) (*reinterpret_cast<const float4*>(p))
#define ST4(p, v) (*reinterpret_cast<float4*>(p) = (v))

// the PlaneOut twin of a store of four consecutive fp32 values at element index idx (idx % 4 == 0) / of one value
__device__ __forceinline__ void planes_st4(const PlaneOut& po, size_t idx, float4 v) {
  if (!po.p) return;
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 pl[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split3(x[j], pl[0][j], pl[1][j], pl[2][j]);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (q >= po.P) break;
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(pl[q][0]) | ((uint32_t)__bfloat16_as_ushort(pl[q][1]) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(pl[q][2]) | ((uint32_t)__bfloat16_as_ushort(pl[q][3]) << 16);
    *reinterpret_cast<uint2*>(po.p + (size_t)q * po.stride + idx) = o;
  }
}
__device__ __forceinline__ void planes_st1(const PlaneOut& po, size_t idx, float v) {
  if (!po.p) return;
  __nv_bfloat16 pl[3];
  split3(v, pl[0], pl[1], pl[2]);
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (q >= po.P) break;
    po.p[(size_t)q * po.stride + idx] = pl[q];
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int32_t* __restrict__ q, const int32_t* __restrict__ len, const float* __restrict__ WeT,
                 const float* __restrict__ be, float* __restrict__ y, Drop d, int B, int T, int E, int V, PlaneOut yp) {
  pdl_entry();
  const int E4 = E >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)T * B * E4;
  if (i >= total) return;
  int e = (int)(i % E4) * 4;
  int64_t n = i / E4;
  int b = (int)(n % B), t = (int)(n / B);
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  int w = q[(int64_t)b * T + t];
  if ((!len || t >= T - len[b]) && w >= 1 && w <= V) {
    float4 wv = LD4(WeT + (int64_t)(w - 1) * E + e);
    float4 bv = LD4(be + e);
    float4 m = drop_at4(d, (uint64_t)n * E + e);
    out.x = tanhf(m.x * (wv.x + bv.x));
    out.y = tanhf(m.y * (wv.y + bv.y));
    out.z = tanhf(m.z * (wv.z + bv.z));
    out.w = tanhf(m.w * (wv.w + bv.w));
  }
  ST4(y + n * E + e, out);
  planes_st4(yp, (size_t)n * E + e, out);
}

int embed_fwd(cudaStream_t s, const int32_t* q, const int32_t* len, const float* WeT, const float* be, float* y,
              Drop d, int B, int T, int E, int V, PlaneOut yp) {
  int64_t total = (int64_t)T * B * (E / 4);
  NVQA_CUDA(launch_pdl(embed_fwd_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, q, len, WeT, be, y, d, B, T, E, V, yp));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one CTA per image row: sum of squares by warp shuffles, then scale + dropout
__global__ void __launch_bounds__(256)
imgnorm_drop_kernel(const float* __restrict__ fc7, float* __restrict__ vd, Drop d, int I, int img_norm, int split,
                    PlaneOut vp) {
  pdl_entry();
  // split > 0: columns [0, split) and [split, I) are two feature blocks normalised separately
  // (early fusion, 003_train_ae_based_ef.lua:116-124); split is a multiple of 4
  __shared__ float red[2][8];
  const int b = blockIdx.x;
  const float* row = fc7 + (int64_t)b * I;
  float inv0 = 1.0f, inv1 = 1.0f;
  const int cut = split > 0 ? split : I;
  if (img_norm) {
    float s0 = 0.f, s1 = 0.f;
    for (int j = threadIdx.x * 4; j < I; j += 256 * 4) {
      float4 v = LD4(row + j);
      const float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (j < cut) s0 += q; else s1 += q;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    inv0 = 1.0f / sqrtf(t0);          // no epsilon, as in the reference (:118-119)
    inv1 = split > 0 ? 1.0f / sqrtf(t1) : inv0;
  }
  for (int j = threadIdx.x * 4; j < I; j += 256 * 4) {
    float4 v = LD4(row + j);
    float4 m = drop_at4(d, (uint64_t)b * I + j);
    const float inv = j < cut ? inv0 : inv1;
    v.x = v.x * inv * m.x; v.y = v.y * inv * m.y; v.z = v.z * inv * m.z; v.w = v.w * inv * m.w;
    ST4(vd + (int64_t)b * I + j, v);
    planes_st4(vp, (size_t)b * I + j, v);
  }
}

int imgnorm_drop(cudaStream_t s, const float* fc7, float* vd, Drop d, int B, int I, int img_norm, int split, PlaneOut vp) {
  NVQA_CUDA(launch_pdl(imgnorm_drop_kernel, dim3(B), dim3(256), 0, s, fc7, vd, d, I, img_norm, split, vp));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lstm_gates_fwd_kernel(float* __restrict__ pre, const float* __restrict__ c_prev, int ldp, float* __restrict__ c_new,
                      float* __restrict__ h_new, int ldn, float* __restrict__ xdrop, const int32_t* __restrict__ len,
                      Drop d, int t, int T, int B, int H) {
  pdl_entry();
  const int H4 = H >> 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H4) return;
  int j = (i % H4) * 4, b = i / H4;
  float* a = pre + (int64_t)b * 4 * H;
  int64_t o = (int64_t)b * H + j, on = (int64_t)b * ldn + j;
  float4 gi = make_float4(0.f, 0.f, 0.f, 0.f), gf = gi, go = gi, gg = gi, c = gi, h = gi;
  if (!len || t >= T - len[b]) {
    float4 ai = LD4(a + j), af = LD4(a + H + j), ao = LD4(a + 2 * H + j), ag = LD4(a + 3 * H + j);
    float4 cp = LD4(c_prev + (int64_t)b * ldp + j);
#define GATE(k)                                                  \
    gi.k = sigmoidf_(ai.k); gf.k = sigmoidf_(af.k); go.k = sigmoidf_(ao.k); gg.k = tanhf(ag.k); \
    c.k = gf.k * cp.k + gi.k * gg.k; h.k = go.k * tanhf(c.k);
    GATE(x) GATE(y) GATE(z) GATE(w)
#undef GATE
  }
  ST4(a + j, gi); ST4(a + H + j, gf); ST4(a + 2 * H + j, go); ST4(a + 3 * H + j, gg);
  ST4(c_new + on, c); ST4(h_new + on, h);
  if (xdrop) {
    float4 m = drop_at4(d, ((uint64_t)t * B + b) * H + j);
    ST4(xdrop + o, make_float4(h.x * m.x, h.y * m.y, h.z * m.z, h.w * m.w));
  }
}

int lstm_gates_fwd(cudaStream_t s, float* pre_t, const float* c_prev, int ldp, float* c_new, float* h_new, int ldn,
                   float* xdrop_next_t, const int32_t* len, Drop d, int t, int T, int B, int H) {
  NVQA_CUDA(launch_pdl(lstm_gates_fwd_kernel, dim3(ceil_div((int64_t)B * H / 4, 256)), dim3(256), 0, s, pre_t, c_prev, ldp, c_new, h_new, ldn,
                                                                        xdrop_next_t, len, d, t, T, B, H));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
struct StatePtrs { const float* c[4]; const float* h[4]; };

__global__ void __launch_bounds__(256)
qvec_fwd_kernel(StatePtrs sp, float* __restrict__ state, float* __restrict__ qd, Drop d, int B, int H, int L, PlaneOut qp) {
  pdl_entry();
  const int S = 2 * L * H, S4 = S >> 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S4) return;
  int j = (i % S4) * 4, b = i / S4;
  int blk = j / H, jj = j % H, l = blk >> 1;
  const float* src = (blk & 1) ? sp.h[l] : sp.c[l];
  float4 v = LD4(src + (int64_t)b * H + jj);
  ST4(state + (int64_t)b * S + j, v);
  float4 m = drop_at4(d, (uint64_t)b * S + j);
  const float4 o = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
  ST4(qd + (int64_t)b * S + j, o);
  planes_st4(qp, (size_t)b * S + j, o);
}

int qvec_fwd(cudaStream_t s, const float* const* c_fin, const float* const* h_fin, float* state, float* qd,
             Drop d, int B, int H, int L, PlaneOut qp) {
  StatePtrs sp;
  for (int l = 0; l < 4; ++l) { sp.c[l] = l < L ? c_fin[l] : nullptr; sp.h[l] = l < L ? h_fin[l] : nullptr; }
  NVQA_CUDA(launch_pdl(qvec_fwd_kernel, dim3(ceil_div((int64_t)B * 2 * L * H / 4, 256)), dim3(256), 0, s, sp, state, qd, d, B, H, L, qp));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fuse_fwd_kernel(float* __restrict__ qc, float* __restrict__ ic, float* __restrict__ zd, Drop d, int64_t n4, int skip,
                PlaneOut zp) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = LD4(qc + i * 4), b = LD4(ic + i * 4);
  a.x = tanhf(a.x); a.y = tanhf(a.y); a.z = tanhf(a.z); a.w = tanhf(a.w);
  b.x = tanhf(b.x); b.y = tanhf(b.y); b.z = tanhf(b.z); b.w = tanhf(b.w);
  float4 m = drop_at4(d, (uint64_t)i * 4);
  ST4(qc + i * 4, a); ST4(ic + i * 4, b);
  if (skip) {   // netdef.AskipB (misc/netdef.lua:16-25): output = qc + qc * ic
    b.x += 1.0f; b.y += 1.0f; b.z += 1.0f; b.w += 1.0f;
  }
  const float4 o = make_float4(a.x * b.x * m.x, a.y * b.y * m.y, a.z * b.z * m.z, a.w * b.w * m.w);
  ST4(zd + i * 4, o);
  planes_st4(zp, (size_t)i * 4, o);
}

int fuse_fwd(cudaStream_t s, float* qc, float* ic, float* zd, Drop d, int B, int C, int skip, PlaneOut zp) {
  int64_t n4 = (int64_t)B * C / 4;
  NVQA_CUDA(launch_pdl(fuse_fwd_kernel, dim3(ceil_div(n4, 256)), dim3(256), 0, s, qc, ic, zd, d, n4, skip, zp));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one warp per row: max / sum by shuffles; first-max argmax (strict '>' scan, torch.max semantics)
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ scores, const int32_t* __restrict__ labels, float* __restrict__ dscores,
                  float* __restrict__ rowloss, int32_t* __restrict__ argmax, int n, int O, float inv_n, PlaneOut dp) {
  pdl_entry();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* sr = scores + (int64_t)row * O;
  float mx = -INFINITY;
  int am = 0x7fffffff;
  for (int j = lane; j < O; j += 32) {
    float v = sr[j];
    if (v > mx) { mx = v; am = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    int oam = __shfl_xor_sync(0xffffffffu, am, o);
    if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
  }
  if (lane == 0 && argmax) argmax[row] = am + 1;
  if (!labels) return;
  float se = 0.f;
  for (int j = lane; j < O; j += 32) se += expf(sr[j] - mx);
  se = warp_sum(se);
  int y = labels[row] - 1;
  float inv = 1.0f / se;
  if (dscores) {
    float* dr = dscores + (int64_t)row * O;
    for (int j = lane; j < O; j += 32) {
      float p = expf(sr[j] - mx) * inv;
      const float g = (p - (j == y ? 1.0f : 0.0f)) * inv_n;
      dr[j] = g;
      planes_st1(dp, (size_t)row * O + j, g);
    }
  }
  if (lane == 0 && rowloss) rowloss[row] = (y >= 0 && y < O) ? -(sr[y] - mx - logf(se)) : 0.f;
}

int softmax_ce(cudaStream_t s, const float* scores, const int32_t* labels, float* dscores, float* rowloss,
               int32_t* argmax, int n, int O, float inv_n, PlaneOut dp) {
  NVQA_CUDA(launch_pdl(softmax_ce_kernel, dim3(ceil_div(n, 8)), dim3(256), 0, s, scores, labels, dscores, rowloss, argmax, n, O, inv_n, dp));
  NVQA_LAUNCHED();
  return 0;
}

// deterministic final reduction: a single CTA walks the rows in a fixed order
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ rowloss, float* __restrict__ loss, int n) {
  pdl_entry();
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += rowloss[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[0] = t / (float)n;
  }
}

int loss_reduce(cudaStream_t s, const float* rowloss, float* loss, int n) {
  NVQA_CUDA(launch_pdl(loss_reduce_kernel, dim3(1), dim3(256), 0, s, rowloss, loss, n));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fuse_bwd_kernel(const float* __restrict__ dzd, const float* __restrict__ qc, const float* __restrict__ ic,
                float* __restrict__ dqpre, float* __restrict__ dipre, Drop d, int64_t n4, int skip, PlaneOut qp, PlaneOut ip) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = LD4(dzd + i * 4), a = LD4(qc + i * 4), b = LD4(ic + i * 4), m = drop_at4(d, (uint64_t)i * 4);
  float4 dq, di;
#define FB(k)                                   \
  { float dz = g.k * m.k;                                          \
    dq.k = dz * (skip ? b.k + 1.0f : b.k) * (1.0f - a.k * a.k);    \
    di.k = dz * a.k * (1.0f - b.k * b.k); }
  FB(x) FB(y) FB(z) FB(w)
#undef FB
  ST4(dqpre + i * 4, dq); ST4(dipre + i * 4, di);
  planes_st4(qp, (size_t)i * 4, dq); planes_st4(ip, (size_t)i * 4, di);
}

int fuse_bwd(cudaStream_t s, const float* dzd, const float* qc, const float* ic, float* dqpre, float* dipre,
             Drop d, int B, int C, int skip, PlaneOut qp, PlaneOut ip) {
  int64_t n4 = (int64_t)B * C / 4;
  NVQA_CUDA(launch_pdl(fuse_bwd_kernel, dim3(ceil_div(n4, 256)), dim3(256), 0, s, dzd, qc, ic, dqpre, dipre, d, n4, skip, qp, ip));
  NVQA_LAUNCHED();
  return 0;
}

__global__ void __launch_bounds__(256) mask_inplace_kernel(float* __restrict__ x, Drop d, int64_t n4) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = LD4(x + i * 4), m = drop_at4(d, (uint64_t)i * 4);
  ST4(x + i * 4, make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w));
}

int mask_inplace(cudaStream_t s, float* x, Drop d, int64_t n) {
  if (d.mode == 0) return 0;
  NVQA_CUDA(launch_pdl(mask_inplace_kernel, dim3(ceil_div(n / 4, 256)), dim3(256), 0, s, x, d, n / 4));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lstm_gates_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_new,
                      const float* __restrict__ dh_in, int dh_ld, const float* __restrict__ dh_above,
                      const float* __restrict__ dc_in, int dc_ld, float* __restrict__ da, float* __restrict__ dc_out,
                      const int32_t* __restrict__ len, Drop d, int t, int T, int B, int H) {
  pdl_entry();
  const int H4 = H >> 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H4) return;
  int j = (i % H4) * 4, b = i / H4;
  int64_t o = (int64_t)b * H + j;
  float* a = da + (int64_t)b * 4 * H;
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dai = z, daf = z, dao = z, dag = z, dcp = z;
  if (!len || t >= T - len[b]) {
    const float* g = gates + (int64_t)b * 4 * H;
    float4 gi = LD4(g + j), gf = LD4(g + H + j), go = LD4(g + 2 * H + j), gg = LD4(g + 3 * H + j);
    float4 cp = LD4(c_prev + o), cn = LD4(c_new + o);
    float4 dh = LD4(dh_in + (int64_t)b * dh_ld + j), dc = LD4(dc_in + (int64_t)b * dc_ld + j);
    if (dh_above) {
      float4 u = LD4(dh_above + o), m = drop_at4(d, ((uint64_t)t * B + b) * H + j);
      dh.x += u.x * m.x; dh.y += u.y * m.y; dh.z += u.z * m.z; dh.w += u.w * m.w;
    }
#define GB(k)                                                  \
    { float tc = tanhf(cn.k);                                  \
      float dct = dc.k + dh.k * go.k * (1.0f - tc * tc);       \
      dao.k = dh.k * tc * go.k * (1.0f - go.k);                \
      dai.k = dct * gg.k * gi.k * (1.0f - gi.k);               \
      daf.k = dct * cp.k * gf.k * (1.0f - gf.k);               \
      dag.k = dct * gi.k * (1.0f - gg.k * gg.k);               \
      dcp.k = dct * gf.k; }
    GB(x) GB(y) GB(z) GB(w)
#undef GB
  }
  ST4(a + j, dai); ST4(a + H + j, daf); ST4(a + 2 * H + j, dao); ST4(a + 3 * H + j, dag);
  ST4(dc_out + o, dcp);
}

int lstm_gates_bwd(cudaStream_t s, const float* gates_t, const float* c_prev, const float* c_new,
                   const float* dh_in, int dh_ld, const float* dh_above_t, const float* dc_in, int dc_ld,
                   float* da_t, float* dc_out, const int32_t* len, Drop d_above, int t, int T, int B, int H) {
  NVQA_CUDA(launch_pdl(lstm_gates_bwd_kernel, dim3(ceil_div((int64_t)B * H / 4, 256)), dim3(256), 0, s, 
      gates_t, c_prev, c_new, dh_in, dh_ld, dh_above_t, dc_in, dc_ld, da_t, dc_out, len, d_above, t, T, B, H));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// column sums: CTA = 32 columns x 8 row-lanes; row chunks across blockIdx.y; one atomicAdd per
// (column, chunk).  Outputs must be zeroed by the caller.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, int rows, int cols, int lda, int rows_per_chunk, float* __restrict__ out0,
              float* __restrict__ out1) {
  pdl_entry();
  __shared__ float red[8][33];
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int ry = threadIdx.x >> 5;
  int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float acc = 0.f;
  if (c < cols) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;          // four independent chains: the loop is latency-bound
    int r = r0 + ry;
    for (; r + 24 < r1; r += 32) {
      a0 += A[(int64_t)r * lda + c];
      a1 += A[(int64_t)(r + 8) * lda + c];
      a2 += A[(int64_t)(r + 16) * lda + c];
      a3 += A[(int64_t)(r + 24) * lda + c];
    }
    for (; r < r1; r += 8) a0 += A[(int64_t)r * lda + c];
    acc = (a0 + a1) + (a2 + a3);
  }
  red[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    atomicAdd(out0 + c, t);
    if (out1) atomicAdd(out1 + c, t);
  }
}

int colsum(cudaStream_t s, const float* A, int rows, int cols, int lda, float* out0, float* out1) {
  int chunks = rows >= 2048 ? 16 : rows >= 128 ? std::min(16, rows / 64) : 1;   // enough CTAs to hide the load latency
  int rpc = ceil_div(rows, chunks);
  dim3 grid(ceil_div(cols, 32), chunks);
  NVQA_CUDA(launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, s, A, rows, cols, lda, rpc, out0, out1));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Grid-stride over (row, 4 columns) items with a CTA of rows_per_cta * E/4 threads: a thread keeps the SAME four columns in
// every iteration, so the bias gradient (column sums of dpre) accumulates in registers and leaves the CTA as E atomics.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int32_t* __restrict__ q, const int32_t* __restrict__ len, const float* __restrict__ y,
                 float* __restrict__ dx, float* __restrict__ dWeT, Drop d, int B, int T, int E, int V,
                 float* __restrict__ dbias) {
  pdl_entry();
  extern __shared__ float bsum[];                // [E] column sums of this CTA (dbias != nullptr)
  if (dbias) {
    for (int j = threadIdx.x; j < E; j += blockDim.x) bsum[j] = 0.f;
    __syncthreads();
  }
  const int E4 = E >> 2;
  const int e = (int)(threadIdx.x % E4) * 4;     // blockDim.x is a multiple of E/4
  const int rows_per_cta = blockDim.x / E4;
  const int64_t rows = (int64_t)T * B;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t n = (int64_t)blockIdx.x * rows_per_cta + threadIdx.x / E4; n < rows; n += (int64_t)gridDim.x * rows_per_cta) {
    int b = (int)(n % B), t = (int)(n / B);
    int w = q[(int64_t)b * T + t];
    float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((!len || t >= T - len[b]) && w >= 1 && w <= V) {
      float4 g = LD4(dx + n * E + e), yy = LD4(y + n * E + e), m = drop_at4(d, (uint64_t)n * E + e);
      dp.x = g.x * (1.0f - yy.x * yy.x) * m.x;
      dp.y = g.y * (1.0f - yy.y * yy.y) * m.y;
      dp.z = g.z * (1.0f - yy.z * yy.z) * m.z;
      dp.w = g.w * (1.0f - yy.w * yy.w) * m.w;
      float* dst = dWeT + (int64_t)(w - 1) * E + e;
      atomicAdd(dst + 0, dp.x); atomicAdd(dst + 1, dp.y); atomicAdd(dst + 2, dp.z); atomicAdd(dst + 3, dp.w);
      acc.x += dp.x; acc.y += dp.y; acc.z += dp.z; acc.w += dp.w;
    }
    ST4(dx + n * E + e, dp);    // dpre kept (module-level callers sum the bias gradient from it)
  }
  if (dbias) {
    atomicAdd(bsum + e, acc.x); atomicAdd(bsum + e + 1, acc.y); atomicAdd(bsum + e + 2, acc.z); atomicAdd(bsum + e + 3, acc.w);
    __syncthreads();
    for (int j = threadIdx.x; j < E; j += blockDim.x)
      if (bsum[j] != 0.f) atomicAdd(dbias + j, bsum[j]);
  }
}

int embed_bwd(cudaStream_t s, const int32_t* q, const int32_t* len, const float* y, float* dx, float* dWeT,
              Drop d, int B, int T, int E, int V, float* dbias) {
  const int E4 = E / 4;
  if (E4 > 256) {               // wider than a CTA: plain one-item-per-thread layout is not worth a second code path
    NVQA_CHECK(E4 <= 256, "embed_bwd: E > 1024 is not supported");
  }
  const int threads = (256 / E4) * E4, rows_per_cta = threads / E4;
  const int64_t rows = (int64_t)T * B;
  const int grid = (int)std::min<int64_t>(ceil_div(rows, rows_per_cta), 148 * 4);
  NVQA_CUDA(launch_pdl(embed_bwd_kernel, dim3(grid), dim3(threads), dbias ? (size_t)E * 4 : 0, s, q, len, y, dx, dWeT, d, B, T, E, V, dbias));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// arch2 (003_train_vqa_arch2/misc/Encoder_lstm.lua:177-203): x[t] = LookupTable row of the START token (t = 1) or of
// word t-2 (zeros -> token 1); t = 0 is the projected image and is written by a GEMM.  seq is [B x T], NOT right-aligned.
__global__ void __launch_bounds__(256)
lookup_fwd_kernel(const int32_t* __restrict__ seq, const float* __restrict__ table, float* __restrict__ x, int B, int T,
                  int E, int V, int steps) {
  pdl_entry();
  const int E4 = E >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(steps - 1) * B * E4) return;
  int e = (int)(i % E4) * 4;
  int64_t n = i / E4 + B;                        // row t*B + b, t >= 1
  int b = (int)(n % B), t = (int)(n / B);
  int tok = t == 1 ? V + 1 : seq[(int64_t)b * T + (t - 2)];
  if (tok < 1 || tok > V + 1) tok = 1;
  ST4(x + n * E + e, LD4(table + (int64_t)(tok - 1) * E + e));
}

int lookup_fwd(cudaStream_t s, const int32_t* seq, const float* table, float* x, int B, int T, int E, int V, int steps) {
  int64_t total = (int64_t)(steps - 1) * B * (E / 4);
  if (total <= 0) return 0;
  NVQA_CUDA(launch_pdl(lookup_fwd_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, seq, table, x, B, T, E, V, steps));
  NVQA_LAUNCHED();
  return 0;
}

// LookupTable accGradParameters for steps 1.. (Encoder_lstm.lua:256): dtable[token] += dx[t,b]
__global__ void __launch_bounds__(256)
lookup_bwd_kernel(const int32_t* __restrict__ seq, const float* __restrict__ dx, float* __restrict__ dtable, int B, int T,
                  int E, int V, int steps) {
  pdl_entry();
  const int E4 = E >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)(steps - 1) * B * E4) return;
  int e = (int)(i % E4) * 4;
  int64_t n = i / E4 + B;
  int b = (int)(n % B), t = (int)(n / B);
  int tok = t == 1 ? V + 1 : seq[(int64_t)b * T + (t - 2)];
  if (tok < 1 || tok > V + 1) tok = 1;
  if (tok == 1 || tok == V + 1) return;          // the two HOT rows are summed by lookup_bwd_hot_kernel
  float4 g = LD4(dx + n * E + e);
  float* dst = dtable + (int64_t)(tok - 1) * E + e;
  atomicAdd(dst + 0, g.x); atomicAdd(dst + 1, g.y); atomicAdd(dst + 2, g.z); atomicAdd(dst + 3, g.w);
}

// Token 1 (every padded position is fed as token 1 and processed unmasked, Encoder_lstm.lua:197) and the START token V + 1
// (all B rows of step 2) receive B .. T*B scatter-adds each on the same E addresses; as atomics they serialise (measured on
// the autoencoder's twin of this kernel: 570 us instead of ~50 us at config 5).  Their gradient rows are column sums:
// CTA = 32 columns x 8 row lanes over a chunk of rows, one atomic per (column, chunk, hot token).
__global__ void __launch_bounds__(256)
lookup_bwd_hot_kernel(const int32_t* __restrict__ seq, const float* __restrict__ dx, float* __restrict__ dtable, int B, int T,
                      int E, int V, int steps, int rows_per_chunk) {
  pdl_entry();
  __shared__ float red[2][8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int64_t rows = (int64_t)(steps - 1) * B;                   // rows t*B + b, t >= 1
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float a1 = 0.f, as = 0.f;
  if (c < E) {
    for (int64_t k = r0 + ry; k < r1; k += 8) {
      const int64_t n = k + B;
      const int b = (int)(n % B), t = (int)(n / B);
      int tok = t == 1 ? V + 1 : seq[(int64_t)b * T + (t - 2)];
      if (tok < 1 || tok > V + 1) tok = 1;
      if (tok != 1 && tok != V + 1) continue;
      const float g = dx[n * E + c];
      if (tok == 1) a1 += g; else as += g;
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

int lookup_bwd(cudaStream_t s, const int32_t* seq, const float* dx, float* dtable, int B, int T, int E, int V, int steps) {
  int64_t total = (int64_t)(steps - 1) * B * (E / 4);
  if (total <= 0) return 0;
  NVQA_CUDA(launch_pdl(lookup_bwd_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, seq, dx, dtable, B, T, E, V, steps));
  NVQA_LAUNCHED();
  const int64_t rows = (int64_t)(steps - 1) * B;
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(64, rows / 64));
  dim3 grid(ceil_div(E, 32), chunks);
  NVQA_CUDA(launch_pdl(lookup_bwd_hot_kernel, dim3(grid), dim3(256), 0, s, seq, dx, dtable, B, T, E, V, steps, ceil_div(rows, chunks)));
  NVQA_LAUNCHED();
  return 0;
}

// dst[b][j] = drop(b*W + j) * src[b*ld + j]   (head Dropout on the encoder output; also saves the raw state)
__global__ void __launch_bounds__(256)
mask_copy_kernel(const float* __restrict__ src, int ld, float* __restrict__ raw, float* __restrict__ dst, Drop d, int B, int W) {
  pdl_entry();
  const int W4 = W >> 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W4) return;
  int j = (i % W4) * 4, b = i / W4;
  float4 v = LD4(src + (int64_t)b * ld + j), m = drop_at4(d, (uint64_t)b * W + j);
  if (raw) ST4(raw + (int64_t)b * W + j, v);
  ST4(dst + (int64_t)b * W + j, make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w));
}

int mask_copy(cudaStream_t s, const float* src, int ld, float* raw, float* dst, Drop d, int B, int W) {
  NVQA_CUDA(launch_pdl(mask_copy_kernel, dim3(ceil_div((int64_t)B * W / 4, 256)), dim3(256), 0, s, src, ld, raw, dst, d, B, W));
  NVQA_LAUNCHED();
  return 0;
}

// fp32 rows [n x W] -> the same rows as fp32 (dst32, optional) and as P bf16 planes (planes + p * plane_stride): writes an
// initial-state slot of the recurrence (h and its planes) from an arbitrary tensor
__global__ void __launch_bounds__(256)
rows_to_planes_kernel(const float* __restrict__ src, float* __restrict__ dst32, __nv_bfloat16* __restrict__ planes,
                      long long plane_stride, int P, int64_t n4) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = LD4(src + i * 4);
  if (dst32) ST4(dst32 + i * 4, v);
  if (!planes) return;
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 p[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split3(x[j], p[0][j], p[1][j], p[2][j]);
  for (int q = 0; q < P; ++q) {
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(p[q][0]) | ((uint32_t)__bfloat16_as_ushort(p[q][1]) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(p[q][2]) | ((uint32_t)__bfloat16_as_ushort(p[q][3]) << 16);
    *reinterpret_cast<uint2*>(planes + (size_t)q * plane_stride + i * 4) = o;
  }
}

int rows_to_planes(cudaStream_t s, const float* src, float* dst32, __nv_bfloat16* planes, long long plane_stride, int P,
                   int64_t n) {
  NVQA_CUDA(launch_pdl(rows_to_planes_kernel, dim3(ceil_div(n / 4, 256)), dim3(256), 0, s, src, dst32, planes, plane_stride, P, n / 4));
  NVQA_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
clamp_rmsprop_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, int64_t n, float lr,
                     float alpha, float oma, float eps, float wd, float clampv, float gscale) {
  pdl_entry();
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 xv = LD4(x + i), gv = LD4(g + i), mv = LD4(m + i);
#define UP(k)                                                    \
    { float gg = fminf(fmaxf(gv.k * gscale, -clampv), clampv);   \
      gg += wd * xv.k;                                           \
      mv.k = alpha * mv.k + oma * gg * gg;                       \
      xv.k -= lr * (gg / (sqrtf(mv.k) + eps)); }
    UP(x) UP(y) UP(z) UP(w)
#undef UP
    ST4(x + i, xv); ST4(m + i, mv);
  } else {
    for (int64_t k = i; k < n; ++k) {
      float gg = fminf(fmaxf(g[k] * gscale, -clampv), clampv);
      gg += wd * x[k];
      float mm = alpha * m[k] + oma * gg * gg;
      m[k] = mm;
      x[k] -= lr * (gg / (sqrtf(mm) + eps));
    }
  }
}

int clamp_rmsprop(cudaStream_t s, float* x, float* g, float* m, int64_t n, float lr, float alpha, float eps,
                  float wd, float clamp, float gscale) {
  float oma = (float)(1.0 - (double)alpha);
  NVQA_CUDA(launch_pdl(clamp_rmsprop_kernel, dim3(ceil_div(ceil_div(n, 4), 256)), dim3(256), 0, s, x, g, m, n, lr, alpha, oma, eps, wd, clamp, gscale));
  NVQA_LAUNCHED();
  return 0;
}


__global__ void __launch_bounds__(256) zero_segments_kernel(ZeroSegs z) {
  pdl_entry();
  uint32_t* p = reinterpret_cast<uint32_t*>(z.p[blockIdx.y]);
  const long long n = z.n[blockIdx.y];
  const long long head = min(n, (long long)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / 4);   // words up to 16-byte alignment
  const long long n4 = (n - head) / 4;
  uint4* p4 = reinterpret_cast<uint4*>(p + head);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    p4[i] = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < head; i += blockDim.x) p[i] = 0u;
    for (long long i = head + 4 * n4 + threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
  }
}

int zero_segments(cudaStream_t s, const ZeroSegs& z) {
  if (z.count == 0) return 0;
  long long most = 0;
  for (int i = 0; i < z.count; ++i) most = std::max(most, z.n[i]);
  dim3 grid((unsigned)std::min<long long>(ceil_div(ceil_div(most, 4), 256), 592), z.count);
  NVQA_CUDA(launch_pdl(zero_segments_kernel, grid, dim3(256), 0, s, z));
  NVQA_LAUNCHED();
  return 0;
}
}  // namespace nvqa
