// Launchers of the bandwidth-bound kernels of the arch1 step (pointwise.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace nvqa {

// Optional second output of a producer kernel: the same values as bf16 planes [P][rows][K] (K % 8 == 0, no padding), i.e.
// the operand format of the tcgen05 GEMM engine, at the place reserve_planes (umma_engine.cuh) registered for the fp32
// matrix -- the separate fp32 -> planes pass (split_planes_kernel) of every consumer GEMM then disappears.
struct PlaneOut {
  __nv_bfloat16* p = nullptr;
  long long stride = 0;        // elements between planes
  int P = 0;
};

// K2: one-hot Linear + Dropout + Tanh as a gather (002_train_baseline.lua:141-144).  y [T x B x E]
int embed_fwd(cudaStream_t s, const int32_t* q, const int32_t* len, const float* WeT, const float* be, float* y,
              Drop d, int B, int T, int E, int V, PlaneOut yp = PlaneOut());
// fc7 L2 row norm (002_train_baseline.lua:117-123) fused with AxB's Dropout on i (misc/netdef.lua:11)
// split > 0: [0, split) and [split, I) normalised separately (early fusion, 003_train_ae_based_ef.lua:116-124)
int imgnorm_drop(cudaStream_t s, const float* fc7, float* vd, Drop d, int B, int I, int img_norm, int split = 0,
                 PlaneOut vp = PlaneOut());
// gate math + cell update of one LSTM layer at time t (misc/LSTM.lua:45-59); pre -> gates in place
int lstm_gates_fwd(cudaStream_t s, float* pre_t, const float* c_prev, int ldp, float* c_new, float* h_new, int ldn,
                   float* xdrop_next_t, const int32_t* len, Drop d, int t, int T, int B, int H);
// final state -> question vector with AxB's Dropout on q  (002_train_baseline.lua:306, misc/netdef.lua:10)
int qvec_fwd(cudaStream_t s, const float* const* c_fin, const float* const* h_fin, float* state, float* qd,
             Drop d, int B, int H, int L, PlaneOut qp = PlaneOut());
// qc = tanh(qpre), ic = tanh(ipre) (in place), zd = mz * qc * ic   (misc/netdef.lua:10-12, 002_train_baseline.lua:153)
// skip = 1: netdef.AskipB (misc/netdef.lua:16-25), output = qc + qc * ic
int fuse_fwd(cudaStream_t s, float* qc, float* ic, float* zd, Drop d, int B, int C, int skip = 0, PlaneOut zp = PlaneOut());
// nn.CrossEntropyCriterion fwd+bwd + torch.max argmax (002_train_baseline.lua:308-310, 004_eval_model.lua:233)
int softmax_ce(cudaStream_t s, const float* scores, const int32_t* labels, float* dscores, float* rowloss,
               int32_t* argmax, int n, int O, float inv_n, PlaneOut dp = PlaneOut());
int loss_reduce(cudaStream_t s, const float* rowloss, float* loss, int n);
int fuse_bwd(cudaStream_t s, const float* dzd, const float* qc, const float* ic, float* dqpre, float* dipre,
             Drop d, int B, int C, int skip = 0, PlaneOut qp = PlaneOut(), PlaneOut ip = PlaneOut());
int mask_inplace(cudaStream_t s, float* x, Drop d, int64_t n);
// cell backward at time t: writes da_t [B x 4H] and the dc carry
int lstm_gates_bwd(cudaStream_t s, const float* gates_t, const float* c_prev, const float* c_new,
                   const float* dh_in, int dh_ld, const float* dh_above_t, const float* dc_in, int dc_ld,
                   float* da_t, float* dc_out, const int32_t* len, Drop d_above, int t, int T, int B, int H);
// out0[n] (and out1[n]) = sum_r A[r][n]   (bias gradients)
int colsum(cudaStream_t s, const float* A, int rows, int cols, int lda, float* out0, float* out1);
// Tanh/Dropout backward + scatter-add into dWeT [V x E]  (002_train_baseline.lua:320)
// dbias (optional, zeroed by the caller): += column sums of dpre (the Linear's bias gradient), accumulated per CTA in shared
// memory and added with one atomic per column and CTA
int embed_bwd(cudaStream_t s, const int32_t* q, const int32_t* len, const float* y, float* dx, float* dWeT,
              Drop d, int B, int T, int E, int V, float* dbias = nullptr);
// arch2 LookupTable gather / scatter-add (003_train_vqa_arch2/misc/Encoder_lstm.lua:177-203,256) and head Dropout
int lookup_fwd(cudaStream_t s, const int32_t* seq, const float* table, float* x, int B, int T, int E, int V, int steps);
int lookup_bwd(cudaStream_t s, const int32_t* seq, const float* dx, float* dtable, int B, int T, int E, int V, int steps);
int mask_copy(cudaStream_t s, const float* src, int ld, float* raw, float* dst, Drop d, int B, int W);
// n contiguous floats -> a copy (dst32, may be NULL) and P bf16 planes (planes + p * plane_stride, may be NULL)
int rows_to_planes(cudaStream_t s, const float* src, float* dst32, __nv_bfloat16* planes, long long plane_stride, int P,
                   int64_t n);
// gradients*scale -> clamp -> optim.rmsprop, one pass (002_train_baseline.lua:329,408; misc/rmsprop_lrscale.lua:26-34)
int clamp_rmsprop(cudaStream_t s, float* x, float* g, float* m, int64_t n, float lr, float alpha, float eps,
                  float wd, float clamp, float gscale);


// ---- text autoencoder (ae_kernels.cu; 001_train_autoencoder/misc/AutoEncoder_text_nostart.lua) ----
// rows [0, tmax*B) = encoder steps, rows [tmax*B, (2 tmax+1)*B) = decoder steps; seq [B x T] zero-padded right
int ae_embed_fwd(cudaStream_t s, const int32_t* seq, const float* table, float* y, Drop denc, Drop ddec, int B, int T, int E,
                 int V, int tmax);
int ae_embed_bwd(cudaStream_t s, const int32_t* seq, const float* y, const float* dx, float* dtable, Drop denc, Drop ddec,
                 int B, int T, int E, int V, int tmax);
// nn.LanguageModelCriterion (:414-455): targets [(T+1) x B] (0 = none), n_pred = number of predictions
int lm_targets(cudaStream_t s, const int32_t* seq, int32_t* targets, int32_t* n_pred, int B, int T, int V);
int logsoftmax_lm(cudaStream_t s, float* x, int rows, int ld, int ncols, const int32_t* targets, float* rowloss);
int lm_loss_reduce(cudaStream_t s, const float* rowloss, int rows, const int32_t* n_pred, float* loss);
// tensor-core modes: the logits are never rewritten -- per-row log-sum-exp + criterion term; log-probs on demand; the
// backward of criterion + LogSoftMax written directly as bf16 planes with the bias column sums (ae_kernels.cu)
int lm_row_stats(cudaStream_t s, const float* x, int rows, int ld, int ncols, const int32_t* targets, float* lse, float* rowloss);
int lm_logprobs(cudaStream_t s, const float* x, const float* lse, int nrows, int ld, int ncols, float* out);
int lm_grad_planes(cudaStream_t s, const float* x, const float* lse, int rows, int ld, int ncols, const int32_t* targets,
                   const int32_t* n_pred, float gscale, __nv_bfloat16* planes, int pitch, long long plane_stride, int P,
                   float* gbias);
int lm_grad(cudaStream_t s, float* lp, int rows, int ld, int ncols, const int32_t* targets, const int32_t* n_pred, float gscale);
// clamp -> += wd * x -> adam (001_train_arch1_text_autoencoder.lua:237-243, misc/optim_updates.lua:78-111); t = step count (1-based)
int clamp_adam(cudaStream_t s, float* x, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
               float eps, float wd, float clamp, float gscale, int64_t t);


// Clears up to 16 device segments (32-bit words) with ONE launch: the atomically accumulated gradient slices (biases,
// embedding scatter) and the step-barrier counters of a backward pass, instead of a dozen cudaMemsetAsync nodes that each
// cost a stream round trip and break the programmatic-launch chain (common.cuh).
struct ZeroSegs {
  void* p[16];
  long long n[16];      // 32-bit words
  int count = 0;
  void add(void* ptr, long long words) { if (ptr && words > 0 && count < 16) { p[count] = ptr; n[count] = words; ++count; } }
};
int zero_segments(cudaStream_t s, const ZeroSegs& z);
}  // namespace nvqa
