// Internal state of a libnvqa model (shared by model.cu and dp_fused.cu).  Not part of the C ABI.
#pragma once
#include <vector>

#include "../../include/nvqa.h"
#include "lstm_persistent.cuh"
#include "pointwise.cuh"

using namespace nvqa;

struct LayerPtrs { float *Wi, *bi, *Wh, *bh; };

// live per-kernel-class timing with CUDA events on the model's stream (bench.py's roofline leg)
enum GemmCat { CAT_INPROJ = 0, CAT_REC_FWD, CAT_HEAD_FWD, CAT_HEAD_BWD, CAT_REC_BWD, CAT_WGRAD, CAT_DGRAD, CAT_OTHER,
               CAT_PW_FWD, CAT_PW_BWD, CAT_OPT, CAT_COUNT };
static const char* const kCatName[CAT_COUNT] = {"lstm_inproj_gemm", "lstm_recurrent_fwd", "head_fwd_gemm", "head_bwd_gemm",
                                          "lstm_recurrent_bwd", "lstm_wgrad_gemm", "lstm_dgrad_gemm", "other_gemm",
                                          "pointwise_fwd", "pointwise_bwd", "clamp_rmsprop"};
struct ProfCat {
  double ms = 0, flops = 0;
  int64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct nvqa_model {
  nvqa_config cfg;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  // host-buffer batches: the large fc7 copy runs on its own stream and is joined only where the head first reads it
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t fc7_ready = nullptr, fc7_consumed = nullptr;
  bool fc7_pending = false;
  int S = 0;
  int64_t n_blk[3] = {0, 0, 0};
  int64_t off_blk[4] = {0, 0, 0, 0};
  int64_t P = 0;
  float *params = nullptr, *grads = nullptr, *rms = nullptr;
  LayerPtrs lw[4], lg[4];
  float *WeT, *be, *gWeT, *gbe;
  float *Wq, *bq, *Wv, *bv, *Wc, *bc, *gWq, *gbq, *gWv, *gbv, *gWc, *gbc;
  // activations (sized for cfg.B rows)
  float* y = nullptr;
  float *pre[4] = {}, *c[4] = {}, *h[4] = {}, *xdrop[4] = {};
  float *state = nullptr, *qd = nullptr, *vd = nullptr, *qc = nullptr, *ic = nullptr, *zd = nullptr;
  float *scores = nullptr, *dscores = nullptr, *rowloss = nullptr, *loss = nullptr;
  int32_t* argmax = nullptr;
  float *dzd = nullptr, *dqpre = nullptr, *dipre = nullptr, *dqd = nullptr;
  float *da = nullptr, *dxbuf = nullptr, *dh_carry = nullptr, *dc_carry = nullptr;
  __nv_bfloat16* hp[4] = {};        // bf16 planes of h per layer [P][(T+1)B][H] (persistent recurrent kernels)
  __nv_bfloat16* dap = nullptr;     // bf16 planes of da [P][T*B][4H] (persistent backward kernel)
  float* dhbuf = nullptr;           // [2][4][B][H] split-K partials of dh
  unsigned int* grid_counter = nullptr;    // 8 slots of 512 words: forward layer l -> slot l, backward layer l -> slot 4 + l
  uint32_t ctr_zero_mask = 0;              // slots cleared at the start of the pass and not used since
  uint32_t prezero_mask = 0;               // backward phases (1 head, 2 LSTM, 4 embedding) whose atomically accumulated gradient
                                           // slices were already cleared by backward_prezero (one launch instead of 8 memsets)
  bool fused_step = false;                 // inside nvqa_train_step / nvqa_dp_train_step: a backward follows this forward
  int planes = 0;                   // bf16 planes per operand of the tensor-core modes (0 = SIMT)
  bool use_persistent = true;
  // arch2 (003_train_vqa_arch2): image projection, LookupTable, head on the top-layer h
  int TS = 0;                       // time-step capacity of the activation buffers (T for arch1, T + 2 for arch2)
  int steps = 0;                    // arch2: executed steps tmax = 2 + longest question of the batch
  float *Wcnn = nullptr, *bcnn = nullptr, *gWcnn = nullptr, *gbcnn = nullptr, *lookup = nullptr, *glookup = nullptr;
  float* zeros = nullptr;           // [B x H] zeros (initial dc / dh of the arch2 backward)
  // arch3 (001_train_autoencoder text autoencoder): encoder core = lw/lg, decoder core = lw2/lg2, decoder projection, lookup
  LayerPtrs lw2[4], lg2[4];
  float *Wd = nullptr, *bd = nullptr, *gWd = nullptr, *gbd = nullptr;
  float *logits = nullptr;          // [(T+1) B x ldl] logits -> log-probs -> d logits, in place
  int ldl = 0;                      // row pitch of logits: V+1 rounded up to 4
  float *hd = nullptr, *dhd = nullptr;   // Dropout(top h) of the decoder steps and its gradient [(T+1) B x H]
  float *dh_init = nullptr, *dc_init = nullptr;   // d(decoder initial state) = d(encoder final state) [B x H]
  int32_t *targets = nullptr, *n_pred = nullptr;
  float* lse = nullptr;             // tensor-core modes: per-row log-sum-exp of the logits (they stay raw: ae_kernels.cu)
  float* lp_scratch = nullptr;      // [B x (V+1)] log-probs of one decoder step for nvqa_logprobs_get (allocated on first use)
  bool lp_raw = false;              // logits holds raw logits + lse (true) or in-place log-probs (false)
  float *adam_m = nullptr;          // Adam first moment (second moment lives in rms)
  int64_t adam_t = 0;
  // trainer variants of 002_train_vqa_arch1 (nvqa_set_variant): AskipB fusion, lr_scale on encoder + embedding gradients,
  // two-block image norm (early fusion)
  int fusion_skip = 0, norm_split = 0;
  float lr_scale = 1.0f;
  // arch 2, literal reference: the top layer's initial h of a training step is the PREVIOUS step's
  // d loss / d h_T (Encoder_lstm.lua:238-239 stores the head's gradInput into init_state_enc; :37-40 only re-zeroes it on
  // a batch-size change)
  bool stale_h0_literal = false;
  int stale_h0_B = 0;               // batch size of the backward that left its gradInput in h0_stale (0 = none yet)
  bool h0_dirty = false;            // slot 0 of the top layer currently holds a non-zero initial state
  float* h0_stale = nullptr;        // [B x H] copy of that gradInput
  bool lookup_grad_literal = false;   // arch 2 / 3: drop the LookupTable gradient like the literal reference
  bool logp_valid = false;          // logits holds this forward's log-probs (backward overwrites them with d logits)
  bool hp_valid[4] = {false, false, false, false};   // hp[l] holds this step's h planes (persistent forward ran)
  bool dap_valid = false;                            // dap holds the current layer's da planes
  // scratch of the module-level entry points (nvqa_lstm_cell_backward ...), allocated on first use
  float *mod_gates[4] = {}, *mod_c[4] = {}, *mod_h[4] = {}, *mod_x[4] = {};
  float *mod_da = nullptr, *mod_dx = nullptr, *mod_cprev = nullptr, *mod_dc = nullptr, *mod_dy = nullptr;
  // batch
  const int32_t *q = nullptr, *len = nullptr, *labels = nullptr;
  const float* fc7 = nullptr;
  int B = 0;
  int B_layout = 0;                // batch size the time-major buffers were last laid out for (row stride of a time slot)
  int32_t *q_stage = nullptr, *len_stage = nullptr, *lab_stage = nullptr;
  float* fc7_stage = nullptr;
  float* loss_host = nullptr;      // pinned
  int32_t* ans_host = nullptr;     // pinned
  // dropout
  const float *mk_emb = nullptr, *mk_lstm = nullptr, *mk_q = nullptr, *mk_i = nullptr, *mk_z = nullptr;
  int mode = NVQA_MODE_EVAL;
  uint64_t seed = 0;
  bool fwd_done = false;
  UmmaWorkspace* ws = nullptr;
  std::vector<void*> allocs;
  std::vector<std::pair<const float*, const float*>> act_ranges;   // forward activations whose bf16 planes are cached per forward
  // fused data-parallel update over NVLink peer memory (dp_fused.cu)
  int dp_rank = 0, dp_world = 0;
  float* dp_peer_grads[16] = {};     // every rank's flat gradient vector (own entry = grads)
  float* dp_peer_params[16] = {};    // every rank's flat parameter vector
  unsigned int* dp_flags = nullptr;  // flag page written by the peers (layout: dp_fused.cu)
  unsigned int* dp_peer_flags[16] = {};
  unsigned int dp_steps[3] = {0, 0, 0};   // exchanges done so far, per parameter block
  cudaEvent_t dp_fork = nullptr, dp_join = nullptr;
  bool dp_whole_vector = false;      // the exchange is ONE range (whole flat vector) instead of {encoder + embedding}, {multimodal}:
                                     // fixed at nvqa_dp_connect, because the partition also shards the RMSprop state
  // Side stream: work that does not depend on the LSTM runs beside the 128-CTA persistent recurrent kernels, on the ~20 SMs
  // they leave free -- forward: fc7 norm + Dropout + the image Linear of AxB; backward: the weight gradients of the two AxB
  // Linears; data parallel: the exchange of the multimodal block (dp_fused.cu).  aux_fwd / aux_bwd: that work is due and
  // is launched by the hook in front of the first recurrent kernel (or inline on the main stream if none follows).
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  bool aux_fwd = false, aux_bwd = false, aux_fwd_inflight = false, aux_bwd_inflight = false, aux_reduce_used = false;
  bool loss_pending = false;               // fused step: the batch-mean reduction of the row losses rides with aux_launch_bwd
  int aux_enabled = -1;              // NVQA_AUX_STREAM (default 1); 0: everything on the main stream
  bool defer_head = false;           // the head backward may leave the AxB weight gradients to the side stream
  // nvqa_train_step: the multimodal block's clamp + RMSprop follows its weight gradients on the side stream (none of its
  // weights is read again in the step); side_opt_done: that happened, nvqa_rmsprop_step skips the block once
  struct { bool armed = false; float lr = 0, alpha = 0, eps = 0, wd = 0, clamp = 0, gscale = 1; } side_opt;
  bool side_opt_done = false;
  std::vector<void*> dp_opened;      // cudaIpcOpenMemHandle mappings to close
  bool profiling = false;
  ProfCat prof[CAT_COUNT];
};


// Side-stream plumbing (model.cu).  aux_launch_bwd: enqueue the deferred AxB weight-gradient GEMMs of the head backward on
// the side stream (force_stream: even if nothing is pending, make the side stream wait for the main stream's work so far).
int aux_launch_bwd(nvqa_model* m, bool force_stream);
