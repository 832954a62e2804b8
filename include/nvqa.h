/*
 * nvqa.h -- C ABI of the B200-native (sm_100a) arch1 VQA training / eval step.
 *
 * This is the drop-in boundary for the ONE hot path of srama2512/novel-vqa: the body of
 * JdJ() + optim.rmsprop in 002_train_vqa_arch1/002_train_baseline.lua:272-335,408 and of
 * forward() in 002_train_vqa_arch1/004_eval_model.lua:202-218.  The reference has no FFI of
 * its own (it is pure Torch7 Lua); these are the entry points a LuaJIT `ffi.cdef` of this very
 * file binds (see INTEGRATION.md and lua/nvqa_ffi.lua).  The header is restricted to the C
 * subset accepted by LuaJIT's ffi.cdef and by Python cffi/ctypes: no macros in declarations,
 * plain pointers and sizes, no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; it never throws and never
 *     aborts.  nvqa_last_error() returns the message of the last failure on this thread.
 *   - all device work is enqueued on the model's stream; the host blocks only in *_get,
 *     nvqa_loss, nvqa_train_step_host and nvqa_sync.
 *   - "Torch flat layout" = the layout of the reference's getParameters() vectors
 *     (SURVEY.md App. B): encoder = per layer i2h.weight[4H x in], i2h.bias, h2h.weight[4H x H],
 *     h2h.bias; embedding = Linear.weight[E x V], bias[E]; multimodal = Wq[C x 2LH], bq,
 *     Wi[C x I], bi, Wc[O x C], bc.  These are exactly the three tensors of the reference's
 *     .t7 checkpoint {encoder_w_q, embedding_w_q, multimodal_w} (002_train_baseline.lua:401-402).
 *   - token ids and labels are 1-based as in the reference's HDF5 files; 0 = padding.
 *   - there is NO CPU fallback: every compute entry point fails if no sm_100 device is present.
 */
#ifndef NVQA_H
#define NVQA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nvqa_model nvqa_model;

/* precision of the dense contractions (element-wise math is always fp32) */
enum {
  NVQA_PREC_FP32_SIMT = 0, /* fp32 FFMA kernels (exact-fp32 cross-check path)                        */
  NVQA_PREC_BF16X3    = 1, /* tcgen05, fp32 operands split into 3 bf16 planes, 6 MMAs, fp32 accum:
                              fp32-equivalent (1e-4 parity mode on tensor cores)                    */
  NVQA_PREC_BF16      = 2, /* tcgen05, single bf16 plane (the 1e-2 "bf16-operand" mode)             */
  NVQA_PREC_BF16X2    = 3  /* tcgen05, 2 bf16 planes, 3 MMAs (~1e-5 relative per product)            */
};

enum { NVQA_BLOCK_ENCODER = 0, NVQA_BLOCK_EMBEDDING = 1, NVQA_BLOCK_MULTIMODAL = 2 };
enum { NVQA_MODE_EVAL = 0, NVQA_MODE_TRAIN = 1 };
/* fusion block of the arch 1 trainers: netdef.AxB (misc/netdef.lua:6-14) or netdef.AskipB (:16-25, 003_train_ae_based_wp.lua:151) */
enum { NVQA_FUSION_AXB = 0, NVQA_FUSION_ASKIPB = 1 };
/* backward phases, in gradient-readiness order (what the data-parallel host overlaps with NCCL) */
enum { NVQA_PHASE_HEAD = 0, NVQA_PHASE_LSTM = 1, NVQA_PHASE_EMBED = 2, NVQA_PHASE_ALL = 3 };

/* mirrors the cmd:option block of 002_train_baseline.lua:22-48 */
typedef struct nvqa_config {
  int32_t arch;       /* 3 = 001_train_autoencoder text autoencoder (misc/AutoEncoder_text_nostart.lua): blocks
                         0 = encoder LSTM core, 1 = decoder LSTM core then Linear(H, V+1), 2 = LookupTable [(V+1) x E]
                         (nn.AutoEncoder:parameters() order, :86-105); q = seq [B x T] zero-padded right; I, C, O unused;
                         1 = 002_train_vqa_arch1, 2 = 003_train_vqa_arch2 (image + START + words through a
                         LookupTable LSTM encoder; E = -input_encoding_size, L = -num_layers, C unused;
                         blocks: 0 = cnn_w, 1 = encoder_w_q (LSTM core then LookupTable), 2 = multimodal_w;
                         q is the question matrix as stored, NOT right-aligned)                 */
  int32_t V;          /* vocabulary size (count of ix_to_word)            :126-127         */
  int32_t E;          /* -input_encoding_size  (200)                      :34              */
  int32_t H;          /* -rnn_size             (512)                      :35              */
  int32_t L;          /* -rnn_layer            (2)                        :36              */
  int32_t I;          /* -nhimage              (4096)                     :33              */
  int32_t C;          /* -common_embedding_size(1024)                     :37              */
  int32_t O;          /* -num_output           (1000)                     :38              */
  int32_t T;          /* question matrix width (buffer_size_q)            :136             */
  int32_t B;          /* max rows per step (-batch_size, 500)             :31              */
  int32_t precision;  /* NVQA_PREC_*                                                       */
  int32_t img_norm;   /* -img_norm: 1 = L2-normalise fc7 rows inside the step  :117-123    */
  int32_t device;     /* CUDA device ordinal (-gpuid)                                      */
  float   dropout;    /* 0.5 at every Dropout site of the path                             */
} nvqa_config;

const char* nvqa_last_error(void);
int nvqa_version(void);
/* number of CUDA devices of compute capability 10.x; 0 if none (then every compute call fails) */
int nvqa_device_count(void);

/* ---- lifetime ---------------------------------------------------------------------------- */
int nvqa_model_create(const nvqa_config* cfg, nvqa_model** out);
int nvqa_model_destroy(nvqa_model* m);
/* cudaStream_t; NULL = the model's own (non-blocking) stream.  To run on the legacy default stream pass cudaStreamLegacy
 * ((cudaStream_t)0x1), not 0 -- a host that mixes the library with NCCL / torch collectives on the default stream must do
 * so, or the collectives are not ordered against the library's kernels (novel-vqa_b200/dp.py bind_current_stream). */
int nvqa_set_stream(nvqa_model* m, void* cuda_stream);
int nvqa_sync(nvqa_model* m);

/* ---- parameters: module:getParameters() / torch.save table (002_train_baseline.lua:174-181,401) */
int nvqa_param_count(const nvqa_model* m, int block, int64_t* n);
int nvqa_params_set(nvqa_model* m, int block, const float* host_src);   /* Torch flat layout     */
int nvqa_params_get(nvqa_model* m, int block, float* host_dst);
int nvqa_grads_get(nvqa_model* m, int block, float* host_dst);          /* raw (unclamped) grads */
int nvqa_rms_get(nvqa_model* m, int block, float* host_dst);            /* optim state.m         */
int nvqa_rms_set(nvqa_model* m, int block, const float* host_src);
/* flat device vectors [encoder | embedding | multimodal] in the optimiser's order (:183,190), for
 * the host's all-reduce plumbing.  Element order inside a block is the library's internal one
 * (embedding weight is kept [V x E]); every op applied through these views must be element-wise. */
int nvqa_device_views(nvqa_model* m, float** params, float** grads, int64_t* block_offsets4);

/* ---- integer preprocessing on the host: misc/RNNUtils.lua:54-61 and :84-125 (bit-exact) ---- */
int nvqa_right_align(const int32_t* seq, const int32_t* lengths, int32_t nq, int32_t T, int32_t* out);
/* words[sum(len)], batch_sizes[max len], 1-based sort_index[B] / sort_index_inverse[B] (stable
 * descending sort), as returned by sort_encoding_onehot_right_align minus the one-hot expansion */
int nvqa_pack_batch(const int32_t* q_right_aligned, const int32_t* lengths, int32_t B, int32_t T,
                    int32_t* words, int32_t* batch_sizes, int32_t* sort_index,
                    int32_t* sort_index_inverse, int32_t* n_words, int32_t* n_steps);

/* multiple-choice answer selection (004_eval_model.lua:257-271): out[i] = the candidate id (1-based, 0 = padding in
 * mc_ids [n x K]) with the highest score in scores [n x O]; first candidate wins ties; 0 if a row has no candidate */
int nvqa_mc_select(const float* scores, const int32_t* mc_ids, int32_t n, int32_t O, int32_t K, int32_t* out);

/* ---- one step, piecewise (device pointers; rows in ORIGINAL batch order, no sort needed) --- */
/* q: [B x T] right-aligned token ids, len: [B], fc7: [B x I], labels: [B] 1-based (may be NULL
 * for eval).  Pointers must stay valid until the step's last kernel has run. */
int nvqa_set_batch(nvqa_model* m, const int32_t* q_dev, const int32_t* len_dev, const float* fc7_dev,
                   const int32_t* labels_dev, int32_t B);
/* same from host memory (pinned for async): copies into the model's staging buffers */
int nvqa_set_batch_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                        const int32_t* labels, int32_t B);
/* explicit Dropout multipliers (0 or 1/(1-p)) in the padded layout, device pointers, any NULL:
 * emb [T x B x E], lstm [(L-1) x T x B x H], q [B x 2LH], i [B x I], z [B x C].  While unset,
 * training mode draws masks from the counter hash shared with oracle/rng.py. */
/* arch 2 with device-resident batches: number of LSTM steps to execute = 2 + longest question of the batch
 * (misc/Encoder_lstm.lua:185-189,219); nvqa_set_batch resets it to T + 2, nvqa_set_batch_host derives it from len.
 * arch 3: steps = tmax = longest sequence of the batch (AutoEncoder_text_nostart.lua:281); reset to T / derived from len. */
int nvqa_set_steps(nvqa_model* m, int32_t steps);
int nvqa_set_masks(nvqa_model* m, const float* emb, const float* lstm, const float* q, const float* i,
                   const float* z);
int nvqa_forward(nvqa_model* m, int mode, uint64_t seed);        /* embedding .. scores           */
int nvqa_loss(nvqa_model* m, float* loss_host);                  /* criterion:forward (blocks)    */
int nvqa_backward(nvqa_model* m, int phase);                     /* criterion/multimodal/rnn/embedding backward */
/* gradients *= grad_scale; clamp(-clamp, clamp); optim.rmsprop (misc/rmsprop_lrscale.lua:14-34) */
int nvqa_rmsprop_step(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp,
                      float grad_scale);
/* arch 3: grad_params:clamp(-clamp, clamp); grad_params:add(wd, params); adam(...)
 * (001_train_autoencoder/001_train_arch1_text_autoencoder.lua:237-243, misc/optim_updates.lua:78-111) */
int nvqa_adam_step(nvqa_model* m, float lr, float beta1, float beta2, float eps, float wd, float clamp,
                   float grad_scale);
/* arch 3: log-probabilities [B x (V+1)] of decoder step `step` (0-based; self.output_dec[step+1],
 * AutoEncoder_text_nostart.lua:334); valid after nvqa_forward until nvqa_backward consumes them in place */
int nvqa_logprobs_get(nvqa_model* m, int32_t step, float* host_dst);
/* arch 1 trainer variants (002_train_vqa_arch1/003_train_ae_based*.lua): fusion = NVQA_FUSION_*; lr_scale multiplies the
 * encoder and embedding gradients before the clamp (-lr_scale, 003_train_ae_based_wp.lua:30,344); norm_split > 0: with img_norm = 1 the feature columns [0, norm_split) and [norm_split, I) are
 * L2-normalised separately (early fusion of two CNN features, 003_train_ae_based_ef.lua:74,116-124). */
int nvqa_set_variant(nvqa_model* m, int32_t fusion, float lr_scale, int32_t norm_split);
/* arch 2 / 3: on = 1 reproduces the literal reference, whose per-step lookup-table clones share `weight` but not
 * `gradWeight` with the module that parameters() exposes (misc/AutoEncoder_text_nostart.lua:64-66,
 * misc/Encoder_lstm.lua:53): the LookupTable block of the gradient is zeroed before the optimizer (it then only sees
 * weight decay).  Default 0 = the evident intent (gradient accumulated over all steps). */
int nvqa_set_lookup_grad_literal(nvqa_model* m, int32_t on);
/* arch 2: on = 1 reproduces the literal reference's initial state: nn.Encoder:updateGradInput stores
 * the head's gradInput tensor into self.init_state_enc[num_state] (misc/Encoder_lstm.lua:238-239) and _createInitState
 * re-zeroes it only when the batch size changes (:37-40), so from the second training step on the top LSTM layer starts
 * from h0 = the PREVIOUS step's d loss / d h_T (every forward, including validation passes in between).  Default 0 = zero
 * initial state.  Switching it (either way) restarts from zeros like a freshly built nn.Encoder. */
int nvqa_set_stale_h0_literal(nvqa_model* m, int32_t on);
int nvqa_scores_get(nvqa_model* m, float* host_dst);             /* [B x O]                       */
int nvqa_argmax_get(nvqa_model* m, int32_t* host_dst);           /* torch.max(scores,2), 1-based  */
int nvqa_state_get(nvqa_model* m, float* host_dst);              /* final LSTM state tv_q [B x 2LH] */

/* ---- one step, fused convenience (what bench.py's e2e number calls) ------------------------ */
/* JdJ + clamp + rmsprop from HOST buffers: H2D copies, all kernels, D2H of the loss. */
int nvqa_train_step_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                         const int32_t* labels, int32_t B, float lr, uint64_t seed, float* loss_out);
/* the same on the batch already set (nvqa_set_batch / nvqa_set_batch_host): nvqa_forward(TRAIN, seed) ; nvqa_backward(ALL) ;
 * clamp + optimizer with the reference's constants.  Knowing lr up front, the multimodal block's update follows its weight
 * gradients on the side stream beside the LSTM backward (same result as the three separate calls). */
int nvqa_train_step(nvqa_model* m, float lr, uint64_t seed);
/* forward() + argmax from HOST buffers */
int nvqa_eval_step_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                        int32_t B, int32_t* answers_out);

/* ---- module-level pieces behind the reference's nn.Module calls (device pointers, fp32) ---- */
/* LSTM.lstm_conventional():forward({state,x}) for one timestep (misc/LSTM.lua:12-73):
 * state/state_out [n x 2LH], x [n x E]; masks [(L-1) x n x H] or NULL (evaluate mode). */
int nvqa_lstm_cell_forward(nvqa_model* m, const float* state, const float* x, const float* masks,
                           int32_t n, float* state_out);
/* netdef.AxB():forward({q, i}) (misc/netdef.lua:6-14): q [n x 2LH], i [n x I] -> out [n x C]; masks NULL = evaluate */
int nvqa_axb_forward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                     int32_t n, float* out);
/* ---- module-level backward: what the reference's JdJ calls on its nn modules (002_train_baseline.lua:283-326).
 * Every *_backward ACCUMULATES the parameter gradients into the model's gradient block (accGradParameters semantics:
 * the sum over the unrolled timestep clones of :323-326 falls out of calling it once per clone); nvqa_grads_zero is
 * module:zeroGradParameters / `encoder_dw_q:zero()` (:283-285); nvqa_grads_get / nvqa_rmsprop_step then see the sums.
 * The entry points are stateless: they recompute the module's forward internals from the same inputs and masks
 * (a Torch7 clone keeps them in its own buffers).  rows n <= cfg.B (cell / AxB / multimodal), n <= B*T (embedding). */
int nvqa_grads_zero(nvqa_model* m, int block);
/* embedding_net_q = Sequential{Linear(V,E), Dropout, Tanh} on one-hot rows (:141-144,300,320): words [n] = 1-based
 * column of the 1 in each one-hot row (sort_encoding_onehot_right_align's packed vector); mask [n x E] or NULL;
 * backward discards the [n x V] gradInput exactly like the reference */
int nvqa_embedding_forward(nvqa_model* m, const int32_t* words, const float* mask, int32_t n, float* y);
int nvqa_embedding_backward(nvqa_model* m, const int32_t* words, const float* y, const float* dy, const float* mask,
                            int32_t n);
/* clone:backward({state, x}, dstate_out) of LSTM.lstm_conventional (misc/LSTM.lua:12-73, misc/RNNUtils.lua:195-196):
 * dstate_out, dstate [n x 2LH], dx [n x E] */
int nvqa_lstm_cell_backward(nvqa_model* m, const float* state, const float* x, const float* masks,
                            const float* dstate_out, int32_t n, float* dstate, float* dx);
/* netdef.AxB():backward({q, i}, dout) (misc/netdef.lua:6-14): dq [n x 2LH]; di [n x I] or NULL (the reference computes
 * and discards it, 002_train_baseline.lua:312-313) */
int nvqa_axb_backward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                      const float* dout, int32_t n, float* dq, float* di);
/* multimodal_net = Sequential{AxB, Dropout, Linear(C,O)} (002_train_baseline.lua:151-154,307,312); masks_z [n x C] */
int nvqa_multimodal_forward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                            const float* masks_z, int32_t n, float* scores);
int nvqa_multimodal_backward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                             const float* masks_z, const float* dscores, int32_t n, float* dq, float* di);
/* optim.rmsprop update (misc/rmsprop_lrscale.lua:26-34) on arbitrary device vectors of length n */
int nvqa_rmsprop_vector(nvqa_model* m, float* x, const float* g, float* state_m, int64_t n, float lr, float alpha,
                        float eps, float wd, float clamp, float grad_scale);
/* nn.CrossEntropyCriterion forward+backward on device scores [n x O] (labels 1-based) */
int nvqa_cross_entropy(nvqa_model* m, const float* scores, const int32_t* labels, int32_t n,
                       float* loss_host, float* dscores);

/* ---- data-parallel update fused with its collective over NVLink peer memory (new functionality; the reference is
 * single-GPU).  One process per GPU.  Every rank calls nvqa_dp_export, the host exchanges the blobs (any transport,
 * e.g. torch.distributed all_gather), every rank calls nvqa_dp_connect with all blobs in rank order.  Then, instead of
 * "all-reduce ; nvqa_rmsprop_step", each step calls nvqa_dp_rmsprop_step after nvqa_backward: ONE kernel reduce-scatters
 * the gradient over NVLink (fixed rank order), applies scale 1/world -> clamp -> optim.rmsprop
 * (002_train_baseline.lua:329,408) to this rank's shard and stores the updated shard into every rank's parameters. */
int nvqa_dp_blob_size(void);
int nvqa_dp_export(nvqa_model* m, void* blob_out);
int nvqa_dp_connect(nvqa_model* m, int32_t rank, int32_t world, const void* blobs);
int nvqa_dp_disconnect(nvqa_model* m);
int nvqa_dp_rmsprop_step(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp);
/* the whole data-parallel iteration on the batch already set (nvqa_set_batch*): nvqa_forward(TRAIN, seed), nvqa_backward
 * with the exchange + update of every parameter block started as soon as its gradient is final -- the multimodal block
 * (53 % of the gradient, final after the head backward) on a side stream beside the LSTM backward, encoder and embedding
 * after theirs.  Same result as nvqa_forward ; nvqa_backward ; nvqa_dp_rmsprop_step.  -lr_scale (nvqa_set_variant) and the
 * literal lookup gradient (nvqa_set_lookup_grad_literal) are honoured by both. */
int nvqa_dp_train_step(nvqa_model* m, float lr, uint64_t seed, float alpha, float eps, float wd, float clamp);
/* A lagging peer is waited for up to NVQA_DP_TIMEOUT_S seconds (environment, default 600); then the wait gives up and
 * *timed_out = 1 from here on (nvqa_sync returns an error): the replica's parameters are invalid, the process survives. */
int nvqa_dp_status(nvqa_model* m, int32_t* timed_out);
/* partition of the flat vector among the ranks (it also shards the RMSprop state): *whole_vector = 1: rank r owns the r-th
 * 1/N of the whole vector; 0: of each of the two ranges {encoder + embedding} and {multimodal}.  Fixed at nvqa_dp_connect
 * (set -lr_scale before connecting). */
int nvqa_dp_layout(nvqa_model* m, int32_t* whole_vector);

/* ---- utilities ------------------------------------------------------------------------------ */
int nvqa_host_alloc(void** p, int64_t bytes);     /* pinned host memory */
int nvqa_host_free(void* p);
int nvqa_device_alloc(void** p, int64_t bytes);
int nvqa_device_free(void* p);
int nvqa_memcpy_h2d(nvqa_model* m, void* dst, const void* src, int64_t bytes);
int nvqa_memcpy_d2h(nvqa_model* m, void* dst, const void* src, int64_t bytes);
/* live timing of the GEMM kernel classes with CUDA events on the model's stream: enable, run steps,
 * then fetch a JSON array [{"kernel","launches","ms","flops"}...] (blocks). */
int nvqa_profile(nvqa_model* m, int enable);
int nvqa_profile_report(nvqa_model* m, char* json_out, int32_t capacity);
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t nvqa_launch_count(void);
/* stand-alone GEMM for tests: C[M x N] = A (.) B with A stored [M x K] (a_kmajor) or [K x M],
 * B stored [N x K] (b_kmajor) or [K x N]; fp32 device pointers; precision = NVQA_PREC_* */
int nvqa_gemm_test(int precision, int a_kmajor, int b_kmajor, int32_t M, int32_t N, int32_t K,
                   const float* A, const float* B, float* C, void* cuda_stream);
/* the same with the rest of nn.Linear's contract: C (row pitch ldc >= N) = (beta ? C : 0) + A (.) B + bias0[n] + bias1[n]
 * (biases nullable) -- accGradParameters accumulates (beta), LSTM.lua:24-25 adds two Linear biases */
int nvqa_gemm_test_ex(int precision, int a_kmajor, int b_kmajor, int32_t M, int32_t N, int32_t K,
                      const float* A, const float* B, float* C, int32_t ldc, int beta, const float* bias0,
                      const float* bias1, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NVQA_H */
