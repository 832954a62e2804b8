// libnvqa: the arch1 VQA training / eval step of srama2512/novel-vqa as hand-written sm_100a CUDA
// behind a C ABI (include/nvqa.h).  This file owns device memory, the parameter layout and the
// per-step kernel schedule; kernels live in pointwise.cu, simt_gemm.cu and umma_gemm.cu.
//
// Path restated: JdJ + optim.rmsprop (002_train_vqa_arch1/002_train_baseline.lua:272-335,408) and
// forward() (004_eval_model.lua:202-218).  Differences in *schedule* (never in math):
//   - the one-hot nn.Linear embedding (:141-144) is a gather (K2) / scatter-add (K11);
//   - the LSTM is processed layer by layer: the input projection x.Wi^T of ALL timesteps is one
//     batched GEMM, only h.Wh^T is serial in t (the reference runs 4 small GEMMs per timestep clone,
//     misc/LSTM.lua:41-43); weights are shared across timesteps instead of 26 cloned copies
//     (misc/RNNUtils.lua:66-81), dW is accumulated by one GEMM over all (t,b) rows instead of a sum
//     over clones (:323-326);
//   - rows stay in original batch order; activity masks replace the length sort (see pointwise.cu);
//   - d fc7 (computed and discarded by the reference) is not computed.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/nvqa.h"
#include "lstm_persistent.cuh"
#include "pointwise.cuh"

namespace nvqa {
static thread_local std::string g_err;
int64_t g_launches = 0;
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NVQA_PDL"); on = e ? atoi(e) : 1; }
  return on != 0;
}
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace nvqa

using namespace nvqa;

#include "model.cuh"

static int dalloc(nvqa_model* m, void** p, size_t bytes) {
  NVQA_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  m->allocs.push_back(*p);
  return 0;
}
template <typename T>
static int dallocT(nvqa_model* m, T** p, size_t count) { return dalloc(m, reinterpret_cast<void**>(p), count * sizeof(T)); }

// The hash generator compares 8 bits per element: the keep probability is quantised to (256 - thresh) / 256, and the
// multiplier of kept elements is its exact reciprocal, so that E[mask] = 1 for every p (for the reference's p = 0.5:
// thresh 128, scale 2 = 1/(1-p) exactly; oracle/rng.py keep_scale is the bit-identical twin).
static void drop_quantise(float p, uint32_t* thresh, float* scale) {
  uint32_t t = (uint32_t)(p * 256.0f + 0.5f);
  if (t > 255u) t = 255u;
  *thresh = t;
  *scale = 256.0f / (float)(256u - t);
}
static Drop make_drop(const nvqa_model* m, const float* mask, uint32_t stream) {
  Drop d;
  d.mask = mask;
  d.key = stream_key(m->seed, stream);
  drop_quantise(m->cfg.dropout, &d.thresh, &d.scale);
  d.mode = (m->mode == NVQA_MODE_TRAIN && m->cfg.dropout > 0.f) ? (mask ? 1 : 2) : 0;
  return d;
}

static int gemm_raw(nvqa_model* m, bool ak, bool bk, int M, int N, int K, const float* A, int lda, const float* B,
                    int ldb, float* C, int ldc, bool beta, const float* b0, const float* b1);
// cache class of a GEMM operand's bf16 planes: 1 = weight (inside the flat parameter vector: cached for the whole step),
// 2 = forward activation that is written once per forward and read again by the backward GEMMs, 0 = transient
static int drop_lookup_grad(nvqa_model* m);
static int cache_class(const nvqa_model* m, const float* p) {
  if (p >= m->params && p < m->params + m->P) return 1;
  for (const auto& r : m->act_ranges)
    if (p >= r.first && p < r.second) return 2;
  return 0;
}

// event bracket around any stretch of work on the model's stream (no-ops unless profiling)
struct ProfScope {
  nvqa_model* m; int cat; cudaEvent_t e0 = nullptr, e1 = nullptr;
  ProfScope(nvqa_model* m_, int cat_, double flops) : m(m_), cat(cat_) {
    if (!m->profiling) return;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, m->stream);
    m->prof[cat].flops += flops; m->prof[cat].launches += 1;
  }
  ~ProfScope() {
    if (!e0) return;
    cudaEventRecord(e1, m->stream);
    m->prof[cat].pending.emplace_back(e0, e1);
  }
};

static int gemm(nvqa_model* m, int cat, bool ak, bool bk, int M, int N, int K, const float* A, int lda, const float* B,
                int ldb, float* C, int ldc, bool beta, const float* b0 = nullptr, const float* b1 = nullptr) {
  if (!m->profiling) return gemm_raw(m, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1);
  cudaEvent_t e0, e1;
  NVQA_CUDA(cudaEventCreate(&e0));
  NVQA_CUDA(cudaEventCreate(&e1));
  NVQA_CUDA(cudaEventRecord(e0, m->stream));
  int r = gemm_raw(m, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1);
  NVQA_CUDA(cudaEventRecord(e1, m->stream));
  ProfCat& pc = m->prof[cat];
  pc.pending.emplace_back(e0, e1);
  pc.flops += 2.0 * M * N * K;
  pc.launches += 1;
  return r;
}

static int gemm_raw(nvqa_model* m, bool ak, bool bk, int M, int N, int K, const float* A, int lda, const float* B,
                    int ldb, float* C, int ldc, bool beta, const float* b0, const float* b1) {
  const int as = cache_class(m, A), bs = cache_class(m, B);
  switch (m->cfg.precision) {
    case NVQA_PREC_FP32_SIMT:
      return simt_gemm(m->stream, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1);
    case NVQA_PREC_BF16X3:
      return umma_gemm(m->stream, 3, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1, m->ws, as, bs);
    case NVQA_PREC_BF16X2:
      return umma_gemm(m->stream, 2, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1, m->ws, as, bs);
    case NVQA_PREC_BF16:
      return umma_gemm(m->stream, 1, ak, bk, M, N, K, A, lda, B, ldb, C, ldc, beta, b0, b1, m->ws, as, bs);
  }
  set_error("unknown precision");
  return 1;
}

// tensor-core GEMM whose operands may be ready-made bf16 planes (written by the persistent LSTM kernels)
static int gemm_ops(nvqa_model* m, int cat, const UmmaOperand& A, const UmmaOperand& B, int M, int N, int K, float* C,
                    int ldc, bool beta, const float* b0 = nullptr, const float* b1 = nullptr) {
  ProfScope ps(m, cat, 2.0 * M * N * K);
  return umma_gemm_ops(m->stream, m->planes, A, B, M, N, K, C, ldc, beta, b0, b1, m->ws);
}
static UmmaOperand op_f32(const nvqa_model* m, const float* src, int ld, bool kmajor) {
  UmmaOperand o;
  o.src = src; o.ld = ld; o.kmajor = kmajor;
  o.is_static = cache_class(m, src);
  return o;
}
static UmmaOperand op_planes(const __nv_bfloat16* planes, int plane_rows, int pitch, int row_offset, bool kmajor) {
  UmmaOperand o;
  o.planes = planes; o.plane_rows = plane_rows; o.pitch = pitch; o.row_offset = row_offset; o.kmajor = kmajor;
  return o;
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* nvqa_last_error(void) { return g_err.c_str(); }
extern "C" int nvqa_version(void) { return 100; }
extern "C" int64_t nvqa_launch_count(void) { return g_launches; }

extern "C" int nvqa_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

static int require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device: libnvqa has no CPU fallback (sm_100a kernels only)");
    return 1;
  }
  NVQA_CHECK(device >= 0 && device < n, "device ordinal out of range");
  cudaDeviceProp p;
  NVQA_CUDA(cudaGetDeviceProperties(&p, device));
  NVQA_CHECK(p.major == 10, "libnvqa is built for sm_100a only; this device is not compute capability 10.x");
  NVQA_CUDA(cudaSetDevice(device));
  return 0;
}

extern "C" int nvqa_model_destroy(nvqa_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->cfg.device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->aux_stream) cudaStreamSynchronize(m->aux_stream);
  for (void* p : m->dp_opened) cudaIpcCloseMemHandle(p);
  for (void* p : m->allocs) cudaFree(p);
  if (m->loss_host) cudaFreeHost(m->loss_host);
  if (m->ans_host) cudaFreeHost(m->ans_host);
  if (m->ws) umma_workspace_destroy(m->ws);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->fc7_ready) cudaEventDestroy(m->fc7_ready);
  if (m->fc7_consumed) cudaEventDestroy(m->fc7_consumed);
  if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
  if (m->aux_fork) cudaEventDestroy(m->aux_fork);
  if (m->aux_join) cudaEventDestroy(m->aux_join);
  if (m->dp_fork) cudaEventDestroy(m->dp_fork);
  if (m->dp_join) cudaEventDestroy(m->dp_join);
  delete m;
  return 0;
}

static int model_create_impl(const nvqa_config* cfg, nvqa_model* m) {
  m->cfg = *cfg;
  const int V = cfg->V, E = cfg->E, H = cfg->H, L = cfg->L, C = cfg->C, T = cfg->T, B = cfg->B;
  const int I = cfg->arch == 3 ? 4 : cfg->I, O = cfg->arch == 3 ? 4 : cfg->O;
  NVQA_CHECK(cfg->arch >= 1 && cfg->arch <= 3,
             "arch must be 1 (002_train_vqa_arch1), 2 (003_train_vqa_arch2) or 3 (001_train_autoencoder text autoencoder)");
  const bool a3 = cfg->arch == 3;
  const bool a2 = cfg->arch == 2 || a3;   // arch3 shares arch2's "no multimodal block" buffer shapes
  if (a2) m->cfg.C = cfg->H;        // arch2 has no common embedding: the head reads the H-wide encoder output
  if (a3) {                         // no image, no answer classifier: keep the shared buffers minimal
    m->cfg.I = 4; m->cfg.O = 4;
    NVQA_CHECK(cfg->L == 1, "arch 3 restates the reference default num_layers = 1");
  }
  NVQA_CHECK(V > 0 && E > 0 && H > 0 && L >= 1 && L <= 4 && I > 0 && (a2 || C > 0) && O > 0 && T > 0 && B > 0, "bad config");
  NVQA_CHECK(E % 4 == 0 && H % 4 == 0 && I % 4 == 0 && C % 4 == 0, "E, H, I, C must be multiples of 4 (128-bit rows)");
  NVQA_CHECK(cfg->dropout >= 0.f && cfg->dropout < 1.f, "dropout must be in [0,1)");
  NVQA_CHECK(cfg->precision >= 0 && cfg->precision <= 3, "unknown precision");
  NVQA_TRY(require_device(cfg->device));
  NVQA_CUDA(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  m->stream = m->own_stream;
  NVQA_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  NVQA_CUDA(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
  NVQA_CUDA(cudaEventCreateWithFlags(&m->aux_fork, cudaEventDisableTiming));
  NVQA_CUDA(cudaEventCreateWithFlags(&m->aux_join, cudaEventDisableTiming));
  { const char* e = getenv("NVQA_AUX_STREAM"); m->aux_enabled = e ? atoi(e) : 1; }
  NVQA_CUDA(cudaEventCreateWithFlags(&m->fc7_ready, cudaEventDisableTiming));
  NVQA_CUDA(cudaEventCreateWithFlags(&m->fc7_consumed, cudaEventDisableTiming));
  const int S = (a2 && !a3) ? H : 2 * L * H;
  m->S = S;
  m->TS = a3 ? 2 * T + 1 : a2 ? T + 2 : T;
  m->steps = a3 ? T : m->TS;
  int64_t n_enc = 0;
  for (int l = 0; l < L; ++l) n_enc += (int64_t)4 * H * (l == 0 ? E : H) + 4 * H + (int64_t)4 * H * H + 4 * H;
  if (a3) {
    // getParameters order: encoder, decoder (LSTM core then Linear(H, V+1)), lookup_table
    // (001_train_autoencoder/misc/AutoEncoder_text_nostart.lua:86-105)
    m->n_blk[0] = n_enc;
    m->n_blk[1] = n_enc + (int64_t)(V + 1) * H + (V + 1);
    m->n_blk[2] = (int64_t)(V + 1) * E;
  } else if (!a2) {
    // optimiser order: encoder, embedding, multimodal (002_train_baseline.lua:183,190)
    m->n_blk[0] = n_enc;
    m->n_blk[1] = (int64_t)V * E + E;
    m->n_blk[2] = (int64_t)C * S + C + (int64_t)C * I + C + (int64_t)O * C + O;
  } else {
    // optimiser / checkpoint order: cnn_w, encoder_w_q (LSTM then LookupTable), multimodal_w
    // (003_train_vqa_arch2/002_train_baseline.lua:192,198; misc/Encoder_lstm.lua:66-83)
    m->n_blk[0] = (int64_t)E * I + E;
    m->n_blk[1] = n_enc + (int64_t)(V + 1) * E;
    m->n_blk[2] = (int64_t)O * H + O;
  }
  for (int i = 0; i < 3; ++i) m->off_blk[i + 1] = m->off_blk[i] + ((m->n_blk[i] + 3) / 4) * 4;
  m->P = m->off_blk[3];
  NVQA_TRY(dallocT(m, &m->params, m->P));
  NVQA_TRY(dallocT(m, &m->grads, m->P));
  NVQA_TRY(dallocT(m, &m->rms, m->P));
  NVQA_CUDA(cudaMemsetAsync(m->params, 0, m->P * 4, m->stream));
  NVQA_CUDA(cudaMemsetAsync(m->grads, 0, m->P * 4, m->stream));
  NVQA_CUDA(cudaMemsetAsync(m->rms, 0, m->P * 4, m->stream));   // optim.rmsprop: state.m = 0
  auto carve = [&](float* base, LayerPtrs* lp, float** We, float** be_, float** q6) {
    float* p = base + m->off_blk[(a2 && !a3) ? 1 : 0];
    for (int l = 0; l < L; ++l) {
      int in = l == 0 ? E : H;
      lp[l].Wi = p; p += (int64_t)4 * H * in;
      lp[l].bi = p; p += 4 * H;
      lp[l].Wh = p; p += (int64_t)4 * H * H;
      lp[l].bh = p; p += 4 * H;
    }
    if (a3) {
      *We = base + m->off_blk[2];
      *be_ = nullptr;
      for (int k = 0; k < 6; ++k) q6[k] = nullptr;
      return;
    }
    if (a2) {
      *We = p;                                   // LookupTable.weight [(V+1) x E] follows the LSTM core
      *be_ = nullptr;
      q6[0] = base + m->off_blk[0];              // cnn_projection: W [E x I], b [E]
      q6[1] = q6[0] + (int64_t)E * I;
      q6[2] = q6[3] = nullptr;
      q6[4] = base + m->off_blk[2];              // head: W [O x H], b [O]
      q6[5] = q6[4] + (int64_t)O * H;
      return;
    }
    p = base + m->off_blk[1];
    *We = p; p += (int64_t)V * E;
    *be_ = p;
    p = base + m->off_blk[2];
    q6[0] = p; p += (int64_t)C * S;
    q6[1] = p; p += C;
    q6[2] = p; p += (int64_t)C * I;
    q6[3] = p; p += C;
    q6[4] = p; p += (int64_t)O * C;
    q6[5] = p;
  };
  float* w6[6];
  float* g6[6];
  carve(m->params, m->lw, &m->WeT, &m->be, w6);
  carve(m->grads, m->lg, &m->gWeT, &m->gbe, g6);
  m->Wq = w6[0]; m->bq = w6[1]; m->Wv = w6[2]; m->bv = w6[3]; m->Wc = w6[4]; m->bc = w6[5];
  m->gWq = g6[0]; m->gbq = g6[1]; m->gWv = g6[2]; m->gbv = g6[3]; m->gWc = g6[4]; m->gbc = g6[5];
  if (a3) {
    auto carve2 = [&](float* base, LayerPtrs* lp, float** W, float** b) {
      float* p = base + m->off_blk[1];
      for (int l = 0; l < L; ++l) {
        int in = l == 0 ? E : H;
        lp[l].Wi = p; p += (int64_t)4 * H * in;
        lp[l].bi = p; p += 4 * H;
        lp[l].Wh = p; p += (int64_t)4 * H * H;
        lp[l].bh = p; p += 4 * H;
      }
      *W = p; p += (int64_t)(V + 1) * H;
      *b = p;
    };
    carve2(m->params, m->lw2, &m->Wd, &m->bd);
    carve2(m->grads, m->lg2, &m->gWd, &m->gbd);
    m->ldl = ((V + 1 + 3) / 4) * 4;
    const int64_t rows = (int64_t)(T + 1) * B;
    NVQA_TRY(dallocT(m, &m->logits, rows * m->ldl));
    NVQA_TRY(dallocT(m, &m->hd, rows * H));
    NVQA_TRY(dallocT(m, &m->dhd, rows * H));
    NVQA_TRY(dallocT(m, &m->dh_init, (size_t)B * H));
    NVQA_TRY(dallocT(m, &m->dc_init, (size_t)B * H));
    NVQA_TRY(dallocT(m, &m->targets, rows));
    NVQA_TRY(dallocT(m, &m->lse, rows));
    NVQA_TRY(dallocT(m, &m->n_pred, 4));
    NVQA_TRY(dallocT(m, &m->adam_m, m->P));
    NVQA_CUDA(cudaMemsetAsync(m->adam_m, 0, m->P * 4, m->stream));
  }
  if (a2) {
    m->Wcnn = w6[0]; m->bcnn = w6[1]; m->gWcnn = g6[0]; m->gbcnn = g6[1];
    m->lookup = m->WeT; m->glookup = m->gWeT;
    NVQA_TRY(dallocT(m, &m->h0_stale, (size_t)B * H));
    NVQA_TRY(dallocT(m, &m->zeros, (size_t)B * H));
    NVQA_CUDA(cudaMemsetAsync(m->zeros, 0, (size_t)B * H * 4, m->stream));
  }
  const int C2 = m->cfg.C;

  const int64_t N = (int64_t)m->TS * B;
  NVQA_TRY(dallocT(m, &m->y, N * E));
  for (int l = 0; l < L; ++l) {
    NVQA_TRY(dallocT(m, &m->pre[l], N * 4 * H));
    NVQA_TRY(dallocT(m, &m->c[l], (N + B) * H));
    NVQA_TRY(dallocT(m, &m->h[l], (N + B) * H));
    NVQA_CUDA(cudaMemsetAsync(m->c[l], 0, (N + B) * H * 4, m->stream));
    NVQA_CUDA(cudaMemsetAsync(m->h[l], 0, (N + B) * H * 4, m->stream));
    if (l > 0) NVQA_TRY(dallocT(m, &m->xdrop[l], N * H));
  }
  NVQA_TRY(dallocT(m, &m->state, (int64_t)B * S));
  NVQA_TRY(dallocT(m, &m->qd, (int64_t)B * S));
  NVQA_TRY(dallocT(m, &m->vd, (int64_t)B * I));
  NVQA_TRY(dallocT(m, &m->qc, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->ic, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->zd, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->scores, (int64_t)B * O));
  NVQA_TRY(dallocT(m, &m->dscores, (int64_t)B * O));
  NVQA_TRY(dallocT(m, &m->rowloss, a3 ? (int64_t)(T + 1) * B : (int64_t)B));
  NVQA_TRY(dallocT(m, &m->loss, 4));
  NVQA_TRY(dallocT(m, &m->argmax, (int64_t)B));
  NVQA_TRY(dallocT(m, &m->dzd, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->dqpre, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->dipre, (int64_t)B * C2));
  NVQA_TRY(dallocT(m, &m->dqd, (int64_t)B * S));
  NVQA_TRY(dallocT(m, &m->da, N * 4 * H));
  NVQA_TRY(dallocT(m, &m->dxbuf, N * (H > E ? H : E)));
  NVQA_TRY(dallocT(m, &m->dh_carry, (int64_t)B * H));
  NVQA_TRY(dallocT(m, &m->dc_carry, (int64_t)B * H));
  NVQA_TRY(dallocT(m, &m->q_stage, (int64_t)B * T));
  NVQA_TRY(dallocT(m, &m->len_stage, (int64_t)B));
  NVQA_TRY(dallocT(m, &m->lab_stage, (int64_t)B));
  NVQA_TRY(dallocT(m, &m->fc7_stage, (int64_t)B * I));
  NVQA_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->loss_host), 64));
  NVQA_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->ans_host), (size_t)B * 4));
  m->planes = cfg->precision == NVQA_PREC_BF16X3 ? 3 : cfg->precision == NVQA_PREC_BF16X2 ? 2
              : cfg->precision == NVQA_PREC_BF16 ? 1 : 0;
  NVQA_TRY(dallocT(m, &m->grid_counter, 8 * 512)); // step-barrier counters of the persistent kernels (model.cuh)
  if (m->planes) {
    for (int l = 0; l < L; ++l) {
      NVQA_TRY(dallocT(m, &m->hp[l], (size_t)m->planes * (N + B) * H));
      NVQA_CUDA(cudaMemsetAsync(m->hp[l], 0, (size_t)m->planes * (N + B) * H * 2, m->stream));
    }
  }
  if (m->planes) {
    NVQA_TRY(dallocT(m, &m->dap, (size_t)m->planes * N * 4 * H));
    NVQA_TRY(dallocT(m, &m->dhbuf, (size_t)2 * 4 * B * H));
  }
  if (getenv("NVQA_NO_PERSISTENT")) m->use_persistent = false;
  if (cfg->precision != NVQA_PREC_FP32_SIMT) {
    // transient operand planes of the largest GEMM (wgrad: da^T [4H x TB] and x^T [H x TB], 3 bf16 planes each);
    // static region: every weight matrix in both orientations
    size_t elems = (size_t)(N + 64) * 4 * H + (size_t)(N + 64) * (H > E ? H : E) + (size_t)(B + 64) * (I + S + 2 * C2);
    if (a3) elems = std::max(elems, (size_t)((T + 1) * (size_t)B + 64) * (m->ldl + 8 + H + 8));   // d logits + hd planes
    size_t stat = (size_t)(n_enc + m->n_blk[2] + (a2 ? m->n_blk[0] : 0) + (a3 ? m->n_blk[1] : 0)) * 2 * 6 + (16 << 20);
    // per-forward activation planes (inputs of the forward GEMMs that the wgrad GEMMs read again)
    size_t act = 0;
    auto reg = [&](const float* p, size_t n) { if (p) { m->act_ranges.emplace_back(p, p + n); act += n * 2 * m->planes + 4096; } };
    reg(m->y, (size_t)N * E);
    for (int l = 1; l < L; ++l) reg(m->xdrop[l], (size_t)N * H);
    reg(m->qd, (size_t)B * S); reg(m->vd, (size_t)B * I); reg(m->zd, (size_t)B * C2);
    // written once per step before their GEMMs and read by two of them (weight gradient and input gradient)
    if (!a2) { reg(m->dscores, (size_t)B * O); reg(m->dqpre, (size_t)B * C2); reg(m->dipre, (size_t)B * C2); }
    if (a3) {
      reg(m->hd, (size_t)(T + 1) * B * H);
      reg(m->logits, (size_t)(T + 1) * B * m->ldl);     // d logits: operand of both the wgrad and the dgrad vocabulary GEMM
    }
    NVQA_TRY(umma_workspace_create(&m->ws, elems * 6 + (96 << 20), stat, act + (act >> 3) + (1 << 20)));
  }
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

extern "C" int nvqa_model_create(const nvqa_config* cfg, nvqa_model** out) {
  if (!cfg || !out) { set_error("nvqa_model_create: null argument"); return 1; }
  *out = nullptr;
  nvqa_model* m = new (std::nothrow) nvqa_model();
  if (!m) { set_error("out of host memory"); return 1; }
  int r = model_create_impl(cfg, m);
  if (r != 0) {
    std::string keep = g_err;
    nvqa_model_destroy(m);
    g_err = keep;
    return r;
  }
  *out = m;
  return 0;
}

extern "C" int nvqa_set_stream(nvqa_model* m, void* s) {
  NVQA_CHECK(m, "null model");
  m->stream = s ? reinterpret_cast<cudaStream_t>(s) : m->own_stream;
  return 0;
}

extern "C" int nvqa_sync(nvqa_model* m) {
  NVQA_CHECK(m, "null model");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  if (m->copy_stream) NVQA_CUDA(cudaStreamSynchronize(m->copy_stream));   // an fc7 copy of nvqa_set_batch_host may be in flight
  if (m->aux_stream) NVQA_CUDA(cudaStreamSynchronize(m->aux_stream));
  if (m->dp_world > 0) {
    int32_t bad = 0;
    NVQA_TRY(nvqa_dp_status(m, &bad));
    NVQA_CHECK(!bad, "data-parallel exchange: a peer did not arrive within NVQA_DP_TIMEOUT_S; this replica's parameters are invalid");
  }
  return 0;
}

// ---- parameters ----------------------------------------------------------------------------------
extern "C" int nvqa_param_count(const nvqa_model* m, int block, int64_t* n) {
  NVQA_CHECK(m && n && block >= 0 && block < 3, "bad argument");
  *n = m->n_blk[block];
  return 0;
}

// Torch flat layout <-> internal layout.  Only the embedding differs: Linear.weight is [E x V] in the
// checkpoint (002_train_baseline.lua:142,174) and [V x E] on the device (coalesced row gather).
static int block_copy(nvqa_model* m, float* dev_base, int block, float* host, bool to_device) {
  NVQA_CHECK(m && host && block >= 0 && block < 3, "bad argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  float* d = dev_base + m->off_blk[block];
  const int64_t n = m->n_blk[block];
  if (block != NVQA_BLOCK_EMBEDDING || m->cfg.arch != 1) {
    if (to_device) NVQA_CUDA(cudaMemcpyAsync(d, host, n * 4, cudaMemcpyHostToDevice, m->stream));
    else NVQA_CUDA(cudaMemcpyAsync(host, d, n * 4, cudaMemcpyDeviceToHost, m->stream));
    NVQA_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
  }
  const int V = m->cfg.V, E = m->cfg.E;
  std::vector<float> tmp((size_t)n);
  if (to_device) {
    for (int e = 0; e < E; ++e)
      for (int v = 0; v < V; ++v) tmp[(size_t)v * E + e] = host[(size_t)e * V + v];
    std::memcpy(tmp.data() + (size_t)V * E, host + (size_t)V * E, (size_t)E * 4);
    NVQA_CUDA(cudaMemcpyAsync(d, tmp.data(), n * 4, cudaMemcpyHostToDevice, m->stream));
    NVQA_CUDA(cudaStreamSynchronize(m->stream));
  } else {
    NVQA_CUDA(cudaMemcpyAsync(tmp.data(), d, n * 4, cudaMemcpyDeviceToHost, m->stream));
    NVQA_CUDA(cudaStreamSynchronize(m->stream));
    for (int v = 0; v < V; ++v)
      for (int e = 0; e < E; ++e) host[(size_t)e * V + v] = tmp[(size_t)v * E + e];
    std::memcpy(host + (size_t)V * E, tmp.data() + (size_t)V * E, (size_t)E * 4);
  }
  return 0;
}

extern "C" int nvqa_params_set(nvqa_model* m, int block, const float* src) {
  if (m) umma_workspace_invalidate(m->ws);
  return block_copy(m, m ? m->params : nullptr, block, const_cast<float*>(src), true);
}
extern "C" int nvqa_params_get(nvqa_model* m, int block, float* dst) { return block_copy(m, m ? m->params : nullptr, block, dst, false); }
extern "C" int nvqa_grads_get(nvqa_model* m, int block, float* dst) { return block_copy(m, m ? m->grads : nullptr, block, dst, false); }
extern "C" int nvqa_rms_get(nvqa_model* m, int block, float* dst) { return block_copy(m, m ? m->rms : nullptr, block, dst, false); }
extern "C" int nvqa_rms_set(nvqa_model* m, int block, const float* src) {
  return block_copy(m, m ? m->rms : nullptr, block, const_cast<float*>(src), true);
}

extern "C" int nvqa_device_views(nvqa_model* m, float** params, float** grads, int64_t* off4) {
  NVQA_CHECK(m, "null model");
  if (params) *params = m->params;
  if (grads) *grads = m->grads;
  if (off4) for (int i = 0; i < 4; ++i) off4[i] = m->off_blk[i];
  return 0;
}

// ---- batch ---------------------------------------------------------------------------------------
extern "C" int nvqa_set_batch(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                              const int32_t* labels, int32_t B) {
  NVQA_CHECK(m && q && (m->cfg.arch == 3 || (len && fc7)), "null argument");
  NVQA_CHECK(B > 0 && B <= m->cfg.B, "batch size out of range");
  m->q = q; m->len = len; m->fc7 = fc7; m->labels = labels; m->B = B;
  if (B != m->B_layout) {
    // A time slot is B rows, so slot 0 -- the zero initial state (c, h and the h planes), which no kernel ever writes --
    // now covers rows that were LATER slots of the previous, differently sized batch (e.g. the short last batch of a
    // validation pass followed by full training batches): make it zero again.
    NVQA_CUDA(cudaSetDevice(m->cfg.device));
    const size_t n = (size_t)B * m->cfg.H;
    const size_t plane = (size_t)(m->TS + 1) * m->cfg.B * m->cfg.H;
    for (int l = 0; l < m->cfg.L; ++l) {
      NVQA_CUDA(cudaMemsetAsync(m->c[l], 0, n * 4, m->stream));
      NVQA_CUDA(cudaMemsetAsync(m->h[l], 0, n * 4, m->stream));
      if (m->planes && m->hp[l])
        for (int p = 0; p < m->planes; ++p) NVQA_CUDA(cudaMemsetAsync(m->hp[l] + p * plane, 0, n * 2, m->stream));
    }
    m->B_layout = B;
  }
  m->fwd_done = false;
  m->fc7_pending = false;
  m->steps = m->cfg.arch == 3 ? m->cfg.T : m->TS;
  return 0;
}

extern "C" int nvqa_set_steps(nvqa_model* m, int32_t steps) {
  NVQA_CHECK(m && m->cfg.arch >= 2, "nvqa_set_steps applies to arch 2 and arch 3 models");
  if (m->cfg.arch == 3) NVQA_CHECK(steps >= 1 && steps <= m->cfg.T, "steps (tmax) out of range");
  else NVQA_CHECK(steps >= 2 && steps <= m->TS, "steps out of range");
  m->steps = steps;
  return 0;
}

extern "C" int nvqa_set_batch_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                                   const int32_t* labels, int32_t B) {
  NVQA_CHECK(m && q && len && (fc7 || m->cfg.arch == 3), "null argument");
  NVQA_CHECK(B > 0 && B <= m->cfg.B, "batch size out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaMemcpyAsync(m->q_stage, q, (size_t)B * m->cfg.T * 4, cudaMemcpyHostToDevice, m->stream));
  NVQA_CUDA(cudaMemcpyAsync(m->len_stage, len, (size_t)B * 4, cudaMemcpyHostToDevice, m->stream));
  if (labels) NVQA_CUDA(cudaMemcpyAsync(m->lab_stage, labels, (size_t)B * 4, cudaMemcpyHostToDevice, m->stream));
  NVQA_TRY(nvqa_set_batch(m, m->q_stage, m->len_stage, m->fc7_stage, labels ? m->lab_stage : nullptr, B));
  if (m->cfg.arch != 3) {
    // the 8 MB feature copy overlaps the embedding and the LSTM forward: it runs on the copy stream (after the previous
    // step's last read of the staging buffer) and the compute stream joins it only in front of the image branch
    NVQA_CUDA(cudaStreamWaitEvent(m->copy_stream, m->fc7_consumed, 0));
    NVQA_CUDA(cudaMemcpyAsync(m->fc7_stage, fc7, (size_t)B * m->cfg.I * 4, cudaMemcpyHostToDevice, m->copy_stream));
    NVQA_CUDA(cudaEventRecord(m->fc7_ready, m->copy_stream));
    m->fc7_pending = true;
  }
  if (m->cfg.arch == 2) {   // executed steps tmax = image + START + longest question (Encoder_lstm.lua:185-189,219)
    int mx = 0;
    for (int b = 0; b < B; ++b) mx = len[b] > mx ? len[b] : mx;
    NVQA_CHECK(mx >= 0 && mx <= m->cfg.T, "question length out of range");
    m->steps = 2 + mx;
  }
  if (m->cfg.arch == 3) {   // tmax = longest sequence of the batch (AutoEncoder_text_nostart.lua:249-256,281)
    int mx = 0;
    for (int b = 0; b < B; ++b) mx = len[b] > mx ? len[b] : mx;
    NVQA_CHECK(mx >= 1 && mx <= m->cfg.T, "sequence length out of range");
    m->steps = mx;
  }
  return 0;
}

extern "C" int nvqa_set_masks(nvqa_model* m, const float* emb, const float* lstm, const float* q, const float* i,
                              const float* z) {
  NVQA_CHECK(m, "null model");
  m->mk_emb = emb; m->mk_lstm = lstm; m->mk_q = q; m->mk_i = i; m->mk_z = z;
  return 0;
}

// ---- forward -------------------------------------------------------------------------------------
// the compute stream waits for the asynchronous fc7 host-to-device copy of nvqa_set_batch_host (if one is in flight)
static int join_fc7_copy(nvqa_model* m) {
  if (m->fc7_pending) {
    NVQA_CUDA(cudaStreamWaitEvent(m->stream, m->fc7_ready, 0));
    m->fc7_pending = false;
  }
  return 0;
}

// Planes of an fp32 activation written by its PRODUCER kernel (pointwise.cuh PlaneOut): reserves / registers them in the
// workspace's per-forward cache so that every GEMM naming `src` [rows x K] (ld = K) as an operand finds them.  Empty (the
// consumer GEMM splits as before) in the SIMT mode, for K % 8 != 0, or when `src` is not a registered activation.
static PlaneOut producer_planes(nvqa_model* m, const float* src, int rows, int K) {
  PlaneOut po;
  static int on = -1;
  if (on < 0) { const char* e = getenv("NVQA_PRODUCER_PLANES"); on = e ? atoi(e) : 1; }
  if (!on || !m->planes || !m->ws || (K & 7) || cache_class(m, src) != 2) return po;
  __nv_bfloat16* dst = nullptr;
  int pitch = 0;
  if (reserve_planes(m->ws, m->planes, src, rows, K, K, 2, &dst, &pitch) != 0 || pitch != K) return po;
  po.p = dst; po.stride = (long long)rows * K; po.P = m->planes;
  return po;
}

static int presplit_step_weights(nvqa_model* m, int part = 0);

// ---- side stream (see model.cuh) ---------------------------------------------------------------------------------
// Routes everything enqueued through m->stream to the side stream for the lifetime of the object; GEMMs then take their
// transients from the side arena of the workspace and run with a capped persistent grid.
struct AuxScope {
  nvqa_model* m; cudaStream_t saved; bool on;
  AuxScope(nvqa_model* m_, bool on_) : m(m_), saved(m_->stream), on(on_) {
    if (!on) return;
    m->stream = m->aux_stream;
    if (m->ws) {
      static int cap = -1;
      if (cap < 0) { const char* e = getenv("NVQA_AUX_CTAS"); cap = e ? atoi(e) : 16; }   // B200 sweep: 8..18 equal (1.734 ms), 20 collides with the 4-CTA clusters (1.772), 6 too few (1.785)
      m->ws->side = true; m->ws->cta_cap = cap;
    }
  }
  ~AuxScope() {
    if (!on) return;
    m->stream = saved;
    if (m->ws) { m->ws->side = false; m->ws->cta_cap = 0; }
  }
};
static bool aux_usable(const nvqa_model* m) { return m->aux_enabled && !m->profiling && m->cfg.arch == 1 && m->planes > 0; }
// arch 2 rides the side stream with less: the classifier's weight gradient, the deferred split-K reductions of the LSTM
// weight gradients and the clearing launch of a fused step
static bool aux2_usable(const nvqa_model* m) {
  return m->aux_enabled && !m->profiling && m->cfg.arch == 2 && m->planes > 0 && m->use_persistent;
}
// the autoencoder: only the deferred split-K reductions of its four LSTM weight-gradient GEMMs
static bool aux3_usable(const nvqa_model* m) {
  return m->aux_enabled && !m->profiling && m->cfg.arch == 3 && m->planes > 0 && m->use_persistent;
}

// Everything a backward pass accumulates into with atomics -- the bias gradients (column sums), the
// embedding / LookupTable scatter -- and the step-barrier counters of its recurrent kernels, cleared by ONE launch on m->stream
static int backward_prezero(nvqa_model* m) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NVQA_PREZERO"); on = e ? atoi(e) : 1; }
  if (!on) return 0;                                  // every phase clears its own slices with cudaMemsetAsync (round 2 start)
  const nvqa_config& c = m->cfg;
  ZeroSegs z;
  for (int l = 0; l < c.L; ++l) { z.add(m->lg[l].bi, 4 * c.H); z.add(m->lg[l].bh, 4 * c.H); }
  if (c.arch == 3) {            // decoder vocabulary bias, decoder LSTM biases, the shared LookupTable gradient
    z.add(m->gbd, c.V + 1);
    for (int l = 0; l < c.L; ++l) { z.add(m->lg2[l].bi, 4 * c.H); z.add(m->lg2[l].bh, 4 * c.H); }
    z.add(m->glookup, (long long)(c.V + 1) * c.E);
  } else if (c.arch == 2) {
    z.add(m->gbc, c.O);
    z.add(m->gbcnn, c.E);
    z.add(m->glookup, (long long)(c.V + 1) * c.E);
  } else {
    z.add(m->gbc, c.O); z.add(m->gbq, c.C); z.add(m->gbv, c.C);
    z.add(m->gWeT, m->n_blk[1]);
  }
  z.add(m->grid_counter + 4 * 512, 4 * 512);
  NVQA_TRY(zero_segments(m->stream, z));
  m->prezero_mask = 7u;
  m->ctr_zero_mask |= 0xF0u;
  return 0;
}

// forward image branch of arch 1: fc7 L2 norm + AxB's Dropout on i + Linear(I, C) -> ic (pre-tanh).  to_side: called from
// the hook in front of the first recurrent kernel; otherwise (no persistent kernel ran) inline on the main stream.
static int aux_launch_fwd(nvqa_model* m, bool to_side) {
  if (!m->aux_fwd) return 0;
  m->aux_fwd = false;
  const nvqa_config& c = m->cfg;
  const bool side = to_side && aux_usable(m);
  if (side) {
    NVQA_CUDA(cudaEventRecord(m->aux_fork, m->stream));
    NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
  }
  {
    AuxScope as(m, side);
    NVQA_TRY(presplit_step_weights(m, 2));     // Wq, Wi, Wc planes: needed by the head only
    {
      ProfScope ps(m, CAT_PW_FWD, 0);
      NVQA_TRY(join_fc7_copy(m));
      NVQA_TRY(imgnorm_drop(m->stream, m->fc7, m->vd, make_drop(m, m->mk_i, STREAM_AXB_I), m->B, c.I, c.img_norm, m->norm_split,
                            producer_planes(m, m->vd, m->B, c.I)));
      NVQA_CUDA(cudaEventRecord(m->fc7_consumed, m->stream));
    }
    NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, m->B, c.C, c.I, m->vd, c.I, m->Wv, c.I, m->ic, c.C, false, m->bv));
    // a fused training step: the gradient slices of the coming backward are cleared here, beside the recurrent kernel
    // (the previous step's optimizer and exchange have finished with the gradient vector: they precede this forward)
    if (side && m->fused_step && m->mode == NVQA_MODE_TRAIN) NVQA_TRY(backward_prezero(m));
    if (side) NVQA_CUDA(cudaEventRecord(m->aux_join, m->stream));
  }
  m->aux_fwd_inflight = side;
  return 0;
}
// the main stream waits for side-stream work in flight (forward image branch or backward weight gradients)
static int aux_join_main(nvqa_model* m) {
  if (m->aux_fwd_inflight || m->aux_bwd_inflight) NVQA_CUDA(cudaStreamWaitEvent(m->stream, m->aux_join, 0));
  m->aux_fwd_inflight = m->aux_bwd_inflight = false;
  return 0;
}

// the step-barrier counters of a persistent launch: its own slot, flagged to the launcher if still clear
static unsigned int* ctr_slot(nvqa_model* m, int slot) {
  if (m->ws) m->ws->ctr_zeroed = (m->ctr_zero_mask >> slot) & 1u;
  m->ctr_zero_mask &= ~(1u << slot);
  return m->grid_counter + 512 * slot;
}

static Drop lstm_drop(const nvqa_model* m, int l /* between layer l and l+1 */) {
  const int64_t per = (int64_t)(m->cfg.arch == 2 ? m->steps : m->cfg.T) * m->B * m->cfg.H;
  return make_drop(m, m->mk_lstm ? m->mk_lstm + per * l : nullptr, STREAM_LSTM0 + l);
}

// A run of consecutive time slots [t0, t0 + T) of the activation buffers processed with one set of LSTM weights.
// arch1 / arch2 use a single segment starting at slot 0; the autoencoder runs the encoder in [0, tmax) and the decoder
// in [tmax, 2 tmax + 1): the decoder's slot 0 IS the encoder's last slot, i.e. its initial state (has_init).
struct LstmSeg {
  int t0 = 0, T = 0;
  const LayerPtrs* w = nullptr;
  LayerPtrs* g = nullptr;
  bool has_init = false;
};

static int lstm_layers_backward(nvqa_model* m, const LstmSeg& sg, const int32_t* len, const float* const* dh0,
                                const float* const* dc0, int ld0, const float* dh_top, Drop dh_top_drop, bool want_init);

// All LSTM layers over the segment's steps, layer-major: batched input projection (K3), then the recurrence (K4 persistent
// kernel, or per-step GEMM + gate kernel as the generic fallback).  len == nullptr: every row is active at every step.
static int lstm_layers_forward(nvqa_model* m, const LstmSeg& sg, const int32_t* len) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, E = c.E, H = c.H, L = c.L, T = sg.T;
  const int64_t BH = (int64_t)B * H, r0 = (int64_t)sg.t0 * B;
  cudaStream_t s = m->stream;
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? E : H;
    const float* X = (l == 0 ? m->y : m->xdrop[l]) + r0 * in;
    float* pre = m->pre[l] + r0 * 4 * H;
    float* cb = m->c[l] + r0 * H;
    float* hb = m->h[l] + r0 * H;
    float* xnext = l + 1 < L ? m->xdrop[l + 1] + r0 * H : nullptr;
    NVQA_TRY(gemm(m, CAT_INPROJ, true, true, T * B, 4 * H, in, X, in, sg.w[l].Wi, in, pre, 4 * H, false, sg.w[l].bi,
                  sg.w[l].bh));
    if (m->planes && m->use_persistent) {
      if (l == 0) NVQA_TRY(aux_launch_fwd(m, true));     // the image branch runs beside the recurrent kernels
      // K4: all T steps in one persistent cooperative kernel (W_hh slice resident in shared memory)
      ProfScope ps(m, CAT_REC_FWD, 2.0 * (T - (sg.has_init ? 0 : 1)) * B * 4.0 * H * H);
      // Dropout(h_t) for the layer above leaves the recurrent kernel as bf16 planes (it is only ever a GEMM operand)
      PlaneOut xp;
      if (xnext && lstm_fwd_v2_supported(m->planes, H)) xp = producer_planes(m, xnext, T * B, H);
      unsigned int* ctr = ctr_slot(m, l);
      int rc = lstm_fwd_persistent_v2(s, m->ws, m->planes, sg.w[l].Wh, pre, cb, hb, m->hp[l] + r0 * H,
                                      (long long)(m->TS + 1) * c.B, xnext, len, lstm_drop(m, l), T, B, H, ctr, xp.p,
                                      xp.stride);
      if (rc < 0)
        rc = lstm_fwd_persistent(s, m->ws, m->planes, sg.w[l].Wh, pre, cb, hb, m->hp[l] + r0 * H,
                                 (long long)(m->TS + 1) * c.B, xnext, len, lstm_drop(m, l), T, B, H, ctr);
      if (rc > 0) return rc;
      m->hp_valid[l] = rc == 0;
      if (rc == 0) continue;
    }
    m->hp_valid[l] = false;
    for (int t = 0; t < T; ++t) {
      float* pre_t = pre + (int64_t)t * B * 4 * H;
      if (t > 0 || sg.has_init)   // h_0 == 0 without an initial state: the recurrent term of the first step vanishes
        NVQA_TRY(gemm(m, CAT_REC_FWD, true, true, B, 4 * H, H, hb + t * BH, H, sg.w[l].Wh, H, pre_t, 4 * H, true));
      ProfScope ps(m, CAT_PW_FWD, 0);
      NVQA_TRY(lstm_gates_fwd(s, pre_t, cb + t * BH, H, cb + (t + 1) * BH, hb + (t + 1) * BH, H,
                              xnext ? xnext + t * BH : nullptr, len, lstm_drop(m, l), t, T, B, H));
    }
  }
  return 0;
}
static int lstm_layers_forward(nvqa_model* m, int T, const int32_t* len) {
  LstmSeg sg;
  sg.T = T; sg.w = m->lw; sg.g = m->lg;
  return lstm_layers_forward(m, sg, len);
}

// arch2 forward: 003_train_vqa_arch2/002_train_baseline.lua:308-314 + misc/Encoder_lstm.lua:152-227
static int forward_arch2(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, T = c.T, E = c.E, H = c.H, L = c.L, steps = m->steps;
  cudaStream_t s = m->stream;
  Drop none = make_drop(m, nullptr, 0);
  none.mode = 0;
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(lookup_fwd(s, m->q, m->lookup, m->y, B, T, E, c.V, steps));
    NVQA_TRY(join_fc7_copy(m));
    NVQA_TRY(imgnorm_drop(s, m->fc7, m->vd, none, B, c.I, c.img_norm));
    NVQA_CUDA(cudaEventRecord(m->fc7_consumed, s));
  }
  // step 1: cnn_projection = Linear(I, E) on the (normalised) image feature, written as rows [0, B) of the LSTM input
  NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, B, E, c.I, m->vd, c.I, m->Wcnn, c.I, m->y, E, false, m->bcnn));
  // literal reference (App. C-5): from the second training step on (same batch size), the top layer starts from
  // h0 = the previous step's d loss / d h_T; otherwise slot 0 is the zero state (restored if a literal step dirtied it)
  const bool stale = m->stale_h0_literal && m->stale_h0_B == B;
  if (stale || m->h0_dirty) {
    const long long plane = (long long)(m->TS + 1) * c.B * H;
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(rows_to_planes(s, stale ? m->h0_stale : m->zeros, m->h[L - 1], m->planes ? m->hp[L - 1] : nullptr, plane, m->planes,
                            (int64_t)B * H));
    m->h0_dirty = stale;
  }
  if (m->fused_step && m->mode == NVQA_MODE_TRAIN && aux2_usable(m)) {
    // a fused training step: the gradient slices the coming backward accumulates into (30 MB LookupTable gradient, bias
    // vectors) and its step counters are cleared on the side stream beside the recurrence (see backward_prezero)
    NVQA_CUDA(cudaEventRecord(m->aux_fork, s));
    NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
    {
      AuxScope as(m, true);
      NVQA_TRY(backward_prezero(m));
      NVQA_CUDA(cudaEventRecord(m->aux_join, m->stream));
    }
    m->aux_fwd_inflight = true;
  }
  LstmSeg sg;
  sg.T = steps; sg.w = m->lw; sg.g = m->lg; sg.has_init = stale;
  NVQA_TRY(lstm_layers_forward(m, sg, nullptr));
  NVQA_TRY(aux_join_main(m));
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(mask_copy(s, m->h[L - 1] + (int64_t)steps * B * H, H, m->state, m->zd, make_drop(m, m->mk_z, STREAM_HEAD), B, H));
  }
  NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, B, c.O, H, m->zd, H, m->Wc, H, m->scores, c.O, false, m->bc));
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(softmax_ce(s, m->scores, m->labels, m->labels ? m->dscores : nullptr, m->rowloss, m->argmax, B, c.O,
                        1.0f / (float)B));
    if (m->labels) NVQA_TRY(loss_reduce(s, m->rowloss, m->loss, B));
  }
  m->fwd_done = true;
  return 0;
}

// arch3 forward: nn.AutoEncoder:updateOutput + nn.LanguageModelCriterion:updateOutput
// (001_train_autoencoder/misc/AutoEncoder_text_nostart.lua:222-339, 427-449)
enum : uint32_t { STREAM_AE_ENC_EMB = 32, STREAM_AE_DEC_EMB = 33, STREAM_AE_OUT = 34 };
static Drop ae_drop(const nvqa_model* m, uint32_t stream, float p) {
  Drop d = make_drop(m, nullptr, stream);
  drop_quantise(p, &d.thresh, &d.scale);
  d.mode = (m->mode == NVQA_MODE_TRAIN && p > 0.f) ? 2 : 0;
  return d;
}
static void ae_segments(nvqa_model* m, LstmSeg* enc, LstmSeg* dec) {
  enc->t0 = 0; enc->T = m->steps; enc->w = m->lw; enc->g = m->lg; enc->has_init = false;
  dec->t0 = m->steps; dec->T = m->steps + 1; dec->w = m->lw2; dec->g = m->lg2; dec->has_init = true;
}
static int forward_arch3(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, T = c.T, E = c.E, H = c.H, L = c.L, tmax = m->steps, V1 = c.V + 1;
  const int64_t BH = (int64_t)B * H;
  cudaStream_t s = m->stream;
  LstmSeg enc, dec;
  ae_segments(m, &enc, &dec);
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    // the lookup Dropout is a hard-coded 0.5 (:31); one mask per timestep clone
    NVQA_TRY(ae_embed_fwd(s, m->q, m->lookup, m->y, ae_drop(m, STREAM_AE_ENC_EMB, 0.5f), ae_drop(m, STREAM_AE_DEC_EMB, 0.5f), B, T,
                          E, c.V, tmax));
    NVQA_TRY(lm_targets(s, m->q, m->targets, m->n_pred, B, T, c.V));
  }
  NVQA_TRY(lstm_layers_forward(m, enc, nullptr));
  NVQA_TRY(lstm_layers_forward(m, dec, nullptr));          // initial state = encoder state at tmax (:287-288)
  const int rows = dec.T * B;
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    // encoder output state [c | h] (self.output_enc, :286) for nvqa_state_get
    NVQA_CUDA(cudaMemcpy2DAsync(m->state, (size_t)2 * H * 4, m->c[L - 1] + (int64_t)tmax * BH, (size_t)H * 4, (size_t)H * 4, B,
                                cudaMemcpyDeviceToDevice, s));
    NVQA_CUDA(cudaMemcpy2DAsync(m->state + H, (size_t)2 * H * 4, m->h[L - 1] + (int64_t)tmax * BH, (size_t)H * 4, (size_t)H * 4, B,
                                cudaMemcpyDeviceToDevice, s));
    // 'drop_final' on the decoder's top h of every step (003_train_vqa_arch2/misc/LSTM_decoder.lua:56)
    NVQA_TRY(mask_copy(s, m->h[L - 1] + (int64_t)(dec.t0 + 1) * BH, H, nullptr, m->hd, ae_drop(m, STREAM_AE_OUT, c.dropout), rows, H));
  }
  // 'decoder' Linear(H, V+1) (:57) for all steps at once
  NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, rows, V1, H, m->hd, H, m->Wd, H, m->logits, m->ldl, false, m->bd));
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    static int fused = -1;
    if (fused < 0) { const char* e = getenv("NVQA_AE_FUSED"); fused = e ? atoi(e) : 1; }
    m->lp_raw = fused && m->planes > 0 && m->ws;
    if (m->lp_raw) NVQA_TRY(lm_row_stats(s, m->logits, rows, m->ldl, V1, m->targets, m->lse, m->rowloss));
    else NVQA_TRY(logsoftmax_lm(s, m->logits, rows, m->ldl, V1, m->targets, m->rowloss));
    NVQA_TRY(lm_loss_reduce(s, m->rowloss, rows, m->n_pred, m->loss));
  }
  m->fwd_done = true;
  m->logp_valid = true;
  return 0;
}

// Every weight matrix a training / eval step multiplies with, for the one-launch plane split at the start of a forward.
// part: 0 = all, 1 = everything but arch 1's multimodal weights, 2 = only those (arch 1: they are first needed after the
// LSTM forward, so their split rides on the side stream with the image branch)
static int presplit_step_weights(nvqa_model* m, int part) {
  if (!m->planes || !m->ws) return 0;
  const nvqa_config& c = m->cfg;
  const float* src[16]; int rows[16], K[16], n = 0;
  auto add = [&](const float* p, int r, int k) { if (p && n < 16) { src[n] = p; rows[n] = r; K[n] = k; ++n; } };
  if (part != 2) {
    for (int l = 0; l < c.L; ++l) { add(m->lw[l].Wi, 4 * c.H, l == 0 ? c.E : c.H); add(m->lw[l].Wh, 4 * c.H, c.H); }
    if (c.arch == 3) {
      for (int l = 0; l < c.L; ++l) { add(m->lw2[l].Wi, 4 * c.H, l == 0 ? c.E : c.H); add(m->lw2[l].Wh, 4 * c.H, c.H); }
      add(m->Wd, c.V + 1, c.H);
    } else if (c.arch == 2) {
      add(m->Wcnn, c.E, c.I); add(m->Wc, c.O, c.H);
    }
  }
  if (c.arch == 1 && part != 1) { add(m->Wq, c.C, m->S); add(m->Wv, c.C, c.I); add(m->Wc, c.O, c.C); }
  return presplit_weights(m->ws, m->stream, m->planes, src, rows, K, n);
}

extern "C" int nvqa_forward(nvqa_model* m, int mode, uint64_t seed) {
  NVQA_CHECK(m && m->q, "nvqa_forward: no batch set");
  NVQA_CHECK(mode == NVQA_MODE_EVAL || mode == NVQA_MODE_TRAIN, "bad mode");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  m->mode = mode; m->seed = seed;
  umma_workspace_new_forward(m->ws);        // this forward rewrites the activations: their cached planes are stale
  m->prezero_mask = 0;
  // step-barrier counters of the forward recurrent kernels: one clear here instead of one in front of every launch
  NVQA_CUDA(cudaMemsetAsync(m->grid_counter, 0, sizeof(unsigned int) * 4 * 512, m->stream));
  m->ctr_zero_mask = (m->ctr_zero_mask & ~0xFu) | 0xFu;
  // weight matrices -> bf16 planes in one launch (no-op while they are cached); arch 1's multimodal weights follow on the
  // side stream (aux_launch_fwd)
  NVQA_TRY(presplit_step_weights(m, m->cfg.arch == 1 ? 1 : 0));
  if (m->cfg.arch == 3) return forward_arch3(m);
  if (m->cfg.arch == 2) return forward_arch2(m);
  const nvqa_config& c = m->cfg;
  const int B = m->B, T = c.T, E = c.E, H = c.H, L = c.L, S = m->S;
  const int64_t BH = (int64_t)B * H;
  cudaStream_t s = m->stream;
  // embedding_net_q:forward   (:300)
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(embed_fwd(s, m->q, m->len, m->WeT, m->be, m->y, make_drop(m, m->mk_emb, STREAM_EMB), B, T, E, c.V,
                       producer_planes(m, m->y, T * B, E)));
  }
  m->aux_fwd = true;                        // the image branch is due: see aux_launch_fwd
  // rnn_forward (:303), layer-major
  NVQA_TRY(lstm_layers_forward(m, T, m->len));
  // tv_q (:306) and multimodal_net:forward (:307)
  const float* cf[4];
  const float* hf[4];
  for (int l = 0; l < L; ++l) { cf[l] = m->c[l] + T * BH; hf[l] = m->h[l] + T * BH; }
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(qvec_fwd(s, cf, hf, m->state, m->qd, make_drop(m, m->mk_q, STREAM_AXB_Q), B, H, L, producer_planes(m, m->qd, B, S)));
  }
  // the image branch (head weight planes, fc7 norm + Dropout + Linear(I, C)) was launched on the side stream beside the
  // recurrent kernels (aux_launch_fwd, hooked in lstm_layers_forward), or runs here if no persistent kernel did
  NVQA_TRY(aux_launch_fwd(m, false));
  NVQA_TRY(aux_join_main(m));
  NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, B, c.C, S, m->qd, S, m->Wq, S, m->qc, c.C, false, m->bq));
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(fuse_fwd(s, m->qc, m->ic, m->zd, make_drop(m, m->mk_z, STREAM_HEAD), B, c.C, m->fusion_skip,
                      producer_planes(m, m->zd, B, c.C)));
  }
  NVQA_TRY(gemm(m, CAT_HEAD_FWD, true, true, B, c.O, c.C, m->zd, c.C, m->Wc, c.C, m->scores, c.O, false, m->bc));
  // criterion forward/backward (:308-310) + torch.max (004_eval_model.lua:233)
  {
    ProfScope ps(m, CAT_PW_FWD, 0);
    NVQA_TRY(softmax_ce(s, m->scores, m->labels, m->labels ? m->dscores : nullptr, m->rowloss, m->argmax, B, c.O,
                        1.0f / (float)B, m->labels ? producer_planes(m, m->dscores, B, c.O) : PlaneOut()));
    // the scalar loss is not an input of the backward: in a fused step its reduction leaves the critical path and rides
    // with the deferred weight gradients (aux_launch_bwd)
    m->loss_pending = m->labels && m->fused_step && aux_usable(m) && m->use_persistent;
    if (m->labels && !m->loss_pending) NVQA_TRY(loss_reduce(s, m->rowloss, m->loss, B));
  }
  m->fwd_done = true;
  return 0;
}

extern "C" int nvqa_loss(nvqa_model* m, float* out) {
  NVQA_CHECK(m && out && m->fwd_done && (m->labels || m->cfg.arch == 3), "nvqa_loss: forward with labels has not run");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  if (m->loss_pending) {             // a fused forward whose backward never came: reduce the row losses now
    m->loss_pending = false;
    NVQA_TRY(loss_reduce(m->stream, m->rowloss, m->loss, m->B));
  }
  NVQA_CUDA(cudaMemcpyAsync(m->loss_host, m->loss, 4, cudaMemcpyDeviceToHost, m->stream));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  *out = m->loss_host[0];
  return 0;
}

// ---- backward ------------------------------------------------------------------------------------
// arch2: Linear(H,O) + Dropout backward (003_train_vqa_arch2/002_train_baseline.lua:316-318)
static int backward_head_arch2(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, H = c.H, O = c.O;
  cudaStream_t s = m->stream;
  if (m->prezero_mask & 1u) m->prezero_mask &= ~1u;
  else NVQA_CUDA(cudaMemsetAsync(m->gbc, 0, (size_t)O * 4, s));
  // the input gradient the LSTM backward waits for first ...
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, true, false, B, H, O, m->dscores, O, m->Wc, H, m->dzd, H, false));
  NVQA_TRY(mask_inplace(s, m->dzd, make_drop(m, m->mk_z, STREAM_HEAD), (int64_t)B * H));
  if (m->stale_h0_literal) {   // what the literal reference leaves in init_state_enc[num_state] (Encoder_lstm.lua:238-239)
    NVQA_CUDA(cudaMemcpyAsync(m->h0_stale, m->dzd, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, s));
    m->stale_h0_B = B;
  }
  // ... then the classifier's weight gradient nobody waits for: beside the recurrent kernel when the LSTM phase follows
  // in the same call (defer_head, nvqa_backward(ALL)), else right here
  const bool side = m->defer_head && aux2_usable(m);
  if (side) {
    NVQA_CUDA(cudaEventRecord(m->aux_fork, s));
    NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
  }
  {
    AuxScope as(m, side);
    NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, O, H, B, m->dscores, O, m->zd, H, m->gWc, H, false));
    NVQA_TRY(colsum(m->stream, m->dscores, B, O, O, m->gbc, nullptr));
    if (side) NVQA_CUDA(cudaEventRecord(m->aux_join, m->stream));
  }
  if (side) m->aux_bwd_inflight = true;
  return 0;
}

// arch2: d(step-1 input) -> cnn_projection:backward (:322); d(steps 2..) -> LookupTable accGradParameters
static int backward_embed_arch2(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, E = c.E;
  cudaStream_t s = m->stream;
  ProfScope ps(m, CAT_PW_BWD, 0);
  if (m->prezero_mask & 4u) m->prezero_mask &= ~4u;
  else {
    NVQA_CUDA(cudaMemsetAsync(m->glookup, 0, (size_t)(c.V + 1) * E * 4, s));
    NVQA_CUDA(cudaMemsetAsync(m->gbcnn, 0, (size_t)E * 4, s));
  }
  NVQA_TRY(lookup_bwd(s, m->q, m->dxbuf, m->glookup, B, c.T, E, c.V, m->steps));
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, E, c.I, B, m->dxbuf, E, m->vd, c.I, m->gWcnn, c.I, false));
  NVQA_TRY(colsum(s, m->dxbuf, B, E, E, m->gbcnn, nullptr));
  return 0;
}

// arch3: criterion + LogSoftMax + 'decoder' Linear + 'drop_final' backward for all decoder steps
// (AutoEncoder_text_nostart.lua:444-450, 351-356; 003_train_vqa_arch2/misc/LSTM_decoder.lua:56-59)
static int backward_head_arch3(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, H = c.H, V1 = c.V + 1, rows = (m->steps + 1) * B;
  cudaStream_t s = m->stream;
  {
    ProfScope ps(m, CAT_PW_BWD, 0);
    if (m->prezero_mask & 1u) m->prezero_mask &= ~1u;
    else NVQA_CUDA(cudaMemsetAsync(m->gbd, 0, (size_t)V1 * 4, s));
    bool done = false;
    if (m->lp_raw) {
      // d logits straight as the bf16 planes both vocabulary GEMMs below read (registered under the key they will ask for),
      // with the bias column sums; no fp32 gradient tensor, no split pass, no column-sum pass
      __nv_bfloat16* dst = nullptr;
      int pitch = 0;
      if (reserve_planes(m->ws, m->planes, m->logits, rows, V1, m->ldl, 2, &dst, &pitch) == 0) {
        NVQA_TRY(lm_grad_planes(s, m->logits, m->lse, rows, m->ldl, V1, m->targets, m->n_pred, 1.0f, dst, pitch,
                                (long long)rows * pitch, m->planes, m->gbd));
        done = true;
      }
    }
    if (!done) {
      if (m->lp_raw) {   // (no room for the planes: form the log-probs in place first)
        NVQA_TRY(logsoftmax_lm(s, m->logits, rows, m->ldl, V1, nullptr, nullptr));
        m->lp_raw = false;
      }
      NVQA_TRY(lm_grad(s, m->logits, rows, m->ldl, V1, m->targets, m->n_pred, 1.0f));
      NVQA_TRY(colsum(s, m->logits, rows, V1, m->ldl, m->gbd, nullptr));
    }
  }
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, V1, H, rows, m->logits, m->ldl, m->hd, H, m->gWd, H, false));
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, true, false, rows, H, V1, m->logits, m->ldl, m->Wd, H, m->dhd, H, false));
  return 0;
}
// arch3: decoder steps tmax+1..1 from zero state gradients, then encoder steps tmax..1 from d(decoder initial state)
static int backward_lstm_arch3(nvqa_model* m) {
  const int H = m->cfg.H;
  LstmSeg enc, dec;
  ae_segments(m, &enc, &dec);
  Drop none = make_drop(m, nullptr, 0);
  none.mode = 0;
  const float* z[4] = {m->zeros, m->zeros, m->zeros, m->zeros};
  NVQA_TRY(lstm_layers_backward(m, dec, nullptr, z, z, H, m->dhd, ae_drop(m, STREAM_AE_OUT, m->cfg.dropout), true));
  const float* dh0[4] = {m->dh_init, nullptr, nullptr, nullptr};
  const float* dc0[4] = {m->dc_init, nullptr, nullptr, nullptr};
  NVQA_TRY(lstm_layers_backward(m, enc, nullptr, dh0, dc0, H, nullptr, none, false));
  m->prezero_mask &= ~2u;
  if (m->aux_reduce_used) {       // deferred split-K reductions of the weight gradients: the only work on the side stream
    NVQA_CUDA(cudaEventRecord(m->aux_join, m->aux_stream));
    m->aux_bwd_inflight = true;
    m->aux_reduce_used = false;
  }
  return aux_join_main(m);
}
static int backward_embed_arch3(nvqa_model* m) {
  const nvqa_config& c = m->cfg;
  cudaStream_t s = m->stream;
  ProfScope ps(m, CAT_PW_BWD, 0);
  if (m->prezero_mask & 4u) m->prezero_mask &= ~4u;
  else NVQA_CUDA(cudaMemsetAsync(m->glookup, 0, (size_t)(c.V + 1) * c.E * 4, s));
  return ae_embed_bwd(s, m->q, m->y, m->dxbuf, m->glookup, ae_drop(m, STREAM_AE_ENC_EMB, 0.5f), ae_drop(m, STREAM_AE_DEC_EMB, 0.5f),
                      m->B, c.T, c.E, c.V, m->steps);
}

int aux_launch_bwd(nvqa_model* m, bool to_side);
static int backward_head(nvqa_model* m) {
  if (m->cfg.arch == 3) return backward_head_arch3(m);
  if (m->cfg.arch == 2) return backward_head_arch2(m);
  const nvqa_config& c = m->cfg;
  const int B = m->B, S = m->S, C = c.C, O = c.O, I = c.I;
  cudaStream_t s = m->stream;
  // bias gradients and the embedding scatter accumulate with atomics: clear them (weight gradients
  // are written with beta = 0 by exactly one GEMM each) -- unless backward_prezero already has
  if (m->prezero_mask & 1u) m->prezero_mask &= ~1u;
  else {
    NVQA_CUDA(cudaMemsetAsync(m->gbc, 0, (size_t)O * 4, s));
    NVQA_CUDA(cudaMemsetAsync(m->gbq, 0, (size_t)C * 4, s));
    NVQA_CUDA(cudaMemsetAsync(m->gbv, 0, (size_t)C * 4, s));
  }
  // Linear(C,O) backward: the input gradient first; the weight gradient joins the deferred ones below
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, true, false, B, C, O, m->dscores, O, m->Wc, C, m->dzd, C, false));
  // Dropout, CMulTable, Tanh backward
  NVQA_TRY(fuse_bwd(s, m->dzd, m->qc, m->ic, m->dqpre, m->dipre, make_drop(m, m->mk_z, STREAM_HEAD), B, C, m->fusion_skip,
                    producer_planes(m, m->dqpre, B, C), producer_planes(m, m->dipre, B, C)));
  // AxB Linear backward (no d fc7): the input gradient the LSTM backward waits for first ...
  NVQA_TRY(gemm(m, CAT_HEAD_BWD, true, false, B, S, C, m->dqpre, C, m->Wq, S, m->dqd, S, false));
  NVQA_TRY(mask_inplace(s, m->dqd, make_drop(m, m->mk_q, STREAM_AXB_Q), (int64_t)B * S));
  // ... the two weight gradients nobody waits for are due (aux_launch_bwd): beside the recurrent kernels on the side stream
  // when the LSTM phase follows in the same call (nvqa_backward(ALL), nvqa_dp_train_step), else right here
  m->aux_bwd = true;
  (void)I;
  if (!m->defer_head) NVQA_TRY(aux_launch_bwd(m, false));
  return 0;
}

// weight gradients of the classifier and of the two AxB Linears (002_train_baseline.lua:154, misc/netdef.lua:10-11):
// dWc = dscores^T zd, dWq = dqpre^T qd, dWi = dipre^T vd, and the bias gradients
int aux_launch_bwd(nvqa_model* m, bool to_side) {
  const bool side = to_side && aux_usable(m) && m->cfg.arch == 1;
  if (!m->aux_bwd) {
    // nothing deferred: a caller that is about to enqueue on the side stream still needs it ordered behind the main stream
    if (to_side && m->aux_stream) {
      NVQA_CUDA(cudaEventRecord(m->aux_fork, m->stream));
      NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
    }
    return 0;
  }
  m->aux_bwd = false;
  const nvqa_config& c = m->cfg;
  const int B = m->B, S = m->S, C = c.C, I = c.I;
  if (side || to_side) {
    NVQA_CUDA(cudaEventRecord(m->aux_fork, m->stream));
    NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
  }
  {
    AuxScope as(m, side);
    if (m->loss_pending) {
      m->loss_pending = false;
      NVQA_TRY(loss_reduce(m->stream, m->rowloss, m->loss, B));
    }
    NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, c.O, C, B, m->dscores, c.O, m->zd, C, m->gWc, C, false));
    NVQA_TRY(colsum(m->stream, m->dscores, B, c.O, c.O, m->gbc, nullptr));
    NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, C, S, B, m->dqpre, C, m->qd, S, m->gWq, S, false));
    NVQA_TRY(colsum(m->stream, m->dqpre, B, C, C, m->gbq, nullptr));
    NVQA_TRY(gemm(m, CAT_HEAD_BWD, false, false, C, I, B, m->dipre, C, m->vd, I, m->gWv, I, false));
    NVQA_TRY(colsum(m->stream, m->dipre, B, C, C, m->gbv, nullptr));
    if (side && m->side_opt.armed) {
      // the multimodal block is final and its weights are not read again in this step: clamp + RMSprop right here,
      // beside the LSTM backward (nvqa_train_step)
      const int64_t o2 = m->off_blk[NVQA_BLOCK_MULTIMODAL], n2 = m->off_blk[3] - o2;
      NVQA_TRY(clamp_rmsprop(m->stream, m->params + o2, m->grads + o2, m->rms + o2, n2, m->side_opt.lr, m->side_opt.alpha,
                             m->side_opt.eps, m->side_opt.wd, m->side_opt.clamp, m->side_opt.gscale));
      m->side_opt_done = true;
    }
    if (side) NVQA_CUDA(cudaEventRecord(m->aux_join, m->stream));
  }
  m->aux_bwd_inflight = side;
  if (!side && to_side) {
    // ran inline on the main stream although the caller continues on the side stream: order the side stream behind it
    NVQA_CUDA(cudaEventRecord(m->aux_fork, m->stream));
    NVQA_CUDA(cudaStreamWaitEvent(m->aux_stream, m->aux_fork, 0));
  }
  return 0;
}

// rnn_backward for all layers (top first) over one segment.  dh0[l] / dc0[l]: gradient of the loss w.r.t. layer l's
// final h / c (leading dimension ld0); dh_top [T][B][H] (optional): extra gradient of every step's top-layer h, multiplied
// by dh_top_drop (the consumer's input Dropout); len == nullptr: every row active at every step; want_init: also produce
// the gradient w.r.t. the segment's initial state of layer 0 in m->dh_init / m->dc_init.
static int lstm_layers_backward(nvqa_model* m, const LstmSeg& sg, const int32_t* len, const float* const* dh0,
                                const float* const* dc0, int ld0, const float* dh_top, Drop dh_top_drop, bool want_init) {
  const nvqa_config& c = m->cfg;
  const int B = m->B, E = c.E, H = c.H, L = c.L, T = sg.T;
  const int64_t BH = (int64_t)B * H, r0 = (int64_t)sg.t0 * B;
  cudaStream_t s = m->stream;
  NVQA_CHECK(!want_init || L == 1, "initial-state gradients are implemented for single-layer segments");
  for (int l = L - 1; l >= 0; --l) {
    const float* dh_in = dh0[l];
    const float* dc_in = dc0[l];
    int ld = ld0;
    int rc = -1;
    const int in = l == 0 ? E : H;
    const float* gates = m->pre[l] + r0 * 4 * H;
    const float* cb = m->c[l] + r0 * H;
    float* da = m->da + r0 * 4 * H;
    float* dx = m->dxbuf + r0 * in;
    const float* dh_above = l + 1 < L ? m->dxbuf + r0 * H : dh_top;
    const Drop dabove = l + 1 < L ? lstm_drop(m, l) : dh_top_drop;
    if (m->planes && m->use_persistent) {
      // K9: the whole backward recurrence of this layer in one persistent cooperative kernel
      ProfScope ps(m, CAT_REC_BWD, 2.0 * (T - (want_init ? 0 : 1)) * B * 4.0 * H * H);
      unsigned int* ctr = ctr_slot(m, 4 + l);
      rc = lstm_bwd_persistent_v2(s, m->ws, m->planes, sg.w[l].Wh, gates, cb, dh_in, dc_in, ld, dh_above, dabove, da,
                                  m->dap + r0 * 4 * H, (long long)m->TS * c.B, m->dhbuf, want_init ? m->dh_init : nullptr,
                                  want_init ? m->dc_init : nullptr, len, T, B, H, ctr);
      if (rc < 0)
        rc = lstm_bwd_persistent(s, m->ws, m->planes, sg.w[l].Wh, gates, cb, dh_in, dc_in, ld, dh_above, dabove, da,
                                 m->dap + r0 * 4 * H, (long long)m->TS * c.B, m->dhbuf, want_init ? m->dh_init : nullptr,
                                 want_init ? m->dc_init : nullptr, len, T, B, H, ctr);
      if (rc > 0) return rc;
    }
    m->dap_valid = rc == 0;
    for (int t = T - 1; t >= 0 && rc != 0; --t) {
      NVQA_TRY(lstm_gates_bwd(s, gates + (int64_t)t * B * 4 * H, cb + t * BH, cb + (t + 1) * BH, dh_in, ld,
                              dh_above ? dh_above + t * BH : nullptr, dc_in, ld, da + (int64_t)t * B * 4 * H,
                              m->dc_carry, len, dabove, t, T, B, H));
      if (t > 0 || want_init)   // dh_{t-1} = da_t . Wh
        NVQA_TRY(gemm(m, CAT_REC_BWD, true, false, B, H, 4 * H, da + (int64_t)t * B * 4 * H, 4 * H, sg.w[l].Wh, H, m->dh_carry, H,
                      false));
      dh_in = m->dh_carry; dc_in = m->dc_carry; ld = H;
    }
    if (want_init && rc != 0) {
      NVQA_CUDA(cudaMemcpyAsync(m->dh_init, m->dh_carry, (size_t)BH * 4, cudaMemcpyDeviceToDevice, s));
      NVQA_CUDA(cudaMemcpyAsync(m->dc_init, m->dc_carry, (size_t)BH * 4, cudaMemcpyDeviceToDevice, s));
    }
    const float* X = (l == 0 ? m->y : m->xdrop[l]) + r0 * in;
    if (!(m->prezero_mask & 2u)) {
      NVQA_CUDA(cudaMemsetAsync(sg.g[l].bi, 0, (size_t)4 * H * 4, s));
      NVQA_CUDA(cudaMemsetAsync(sg.g[l].bh, 0, (size_t)4 * H * 4, s));
    }
    // sum over timestep clones of accGradParameters (:323-326) as one GEMM over all (t,b) rows.  Nobody on this stream
    // waits for these weight gradients: their split-K reductions run on the side stream beside the next GEMM.
    static int defer_red = -1;
    if (defer_red < 0) { const char* e = getenv("NVQA_DEFER_REDUCE"); defer_red = e ? atoi(e) : 1; }
    const bool defer = defer_red && m->ws && (aux_usable(m) || aux2_usable(m) || aux3_usable(m));
    if (defer) { m->ws->reduce_stream = m->aux_stream; m->aux_reduce_used = true; }
    if (m->dap_valid) {
      // da (and h_prev) already exist as bf16 planes: no split passes, the GEMMs read them through MN-major TMA maps
      UmmaOperand da_mn = op_planes(m->dap, m->TS * c.B, 4 * H, (int)r0, false);
      NVQA_TRY(gemm_ops(m, CAT_WGRAD, da_mn, op_f32(m, X, in, false), 4 * H, in, T * B, sg.g[l].Wi, in, false));
      UmmaOperand hprev = m->hp_valid[l] ? op_planes(m->hp[l], (m->TS + 1) * c.B, H, (int)r0, false)
                                         : op_f32(m, m->h[l] + r0 * H, H, false);
      NVQA_TRY(gemm_ops(m, CAT_WGRAD, da_mn, hprev, 4 * H, H, T * B, sg.g[l].Wh, H, false));
    } else {
      NVQA_TRY(gemm(m, CAT_WGRAD, false, false, 4 * H, in, T * B, da, 4 * H, X, in, sg.g[l].Wi, in, false));
      NVQA_TRY(gemm(m, CAT_WGRAD, false, false, 4 * H, H, T * B, da, 4 * H, m->h[l] + r0 * H, H, sg.g[l].Wh, H, false));
    }
    if (m->ws) m->ws->reduce_stream = nullptr;
    {
      ProfScope ps(m, CAT_PW_BWD, 0);
      // persistent kernel: da (here) holds the per-row sums over t [B x 4H]; fallback: da_t of every step [T*B x 4H]
      NVQA_TRY(colsum(s, da, m->dap_valid ? B : T * B, 4 * H, 4 * H, sg.g[l].bi, sg.g[l].bh));
    }
    // dX = da . Wi  (layer l-1's dh contribution, or the embedding gradient for l = 0)
    if (m->dap_valid)
      NVQA_TRY(gemm_ops(m, CAT_DGRAD, op_planes(m->dap, m->TS * c.B, 4 * H, (int)r0, true), op_f32(m, sg.w[l].Wi, in, false), T * B, in,
                        4 * H, dx, in, false));
    else
      NVQA_TRY(gemm(m, CAT_DGRAD, true, false, T * B, in, 4 * H, da, 4 * H, sg.w[l].Wi, in, dx, in, false));
  }
  return 0;
}
static int lstm_layers_backward(nvqa_model* m, int T, const int32_t* len, const float* const* dh0, const float* const* dc0, int ld0) {
  LstmSeg sg;
  sg.T = T; sg.w = m->lw; sg.g = m->lg;
  Drop none = make_drop(m, nullptr, 0);
  none.mode = 0;
  return lstm_layers_backward(m, sg, len, dh0, dc0, ld0, nullptr, none, false);
}

static int backward_lstm(nvqa_model* m) {
  if (m->cfg.arch == 3) return backward_lstm_arch3(m);
  const int H = m->cfg.H, L = m->cfg.L;
  const float *dh0[4], *dc0[4];
  if (m->cfg.arch == 2) {   // only the top layer's final h receives a gradient (Encoder_lstm.lua:238-239)
    for (int l = 0; l < L; ++l) { dh0[l] = l == L - 1 ? m->dzd : m->zeros; dc0[l] = m->zeros; }
    NVQA_TRY(lstm_layers_backward(m, m->steps, nullptr, dh0, dc0, H));
    m->prezero_mask &= ~2u;
    if (m->aux_reduce_used) {     // deferred split-K reductions: the last work on the side stream
      NVQA_CUDA(cudaEventRecord(m->aux_join, m->aux_stream));
      m->aux_bwd_inflight = true;
      m->aux_reduce_used = false;
    }
    return aux_join_main(m);
  }
  // arch1: d tv_q = d[c1 h1 c2 h2 ...] (002_train_baseline.lua:306,313)
  for (int l = 0; l < L; ++l) { dh0[l] = m->dqd + (2 * l + 1) * H; dc0[l] = m->dqd + (2 * l) * H; }
  NVQA_TRY(aux_launch_bwd(m, m->planes && m->use_persistent));      // the AxB weight gradients run beside the recurrent kernels
  NVQA_TRY(lstm_layers_backward(m, m->cfg.T, m->len, dh0, dc0, m->S));
  m->prezero_mask &= ~2u;
  if (m->aux_reduce_used) {       // deferred split-K reductions of the weight gradients are the last work on the side stream
    NVQA_CUDA(cudaEventRecord(m->aux_join, m->aux_stream));
    m->aux_bwd_inflight = true;
    m->aux_reduce_used = false;
  }
  return aux_join_main(m);
}

static int backward_embed(nvqa_model* m) {
  if (m->cfg.arch == 3) return backward_embed_arch3(m);
  if (m->cfg.arch == 2) return backward_embed_arch2(m);
  const nvqa_config& c = m->cfg;
  cudaStream_t s = m->stream;
  ProfScope ps(m, CAT_PW_BWD, 0);
  if (m->prezero_mask & 4u) m->prezero_mask &= ~4u;
  else NVQA_CUDA(cudaMemsetAsync(m->gWeT, 0, (size_t)m->n_blk[1] * 4, s));
  // (gWeT and the bias gradient gbe are one contiguous block: both were cleared above)
  NVQA_TRY(embed_bwd(s, m->q, m->len, m->y, m->dxbuf, m->gWeT, make_drop(m, m->mk_emb, STREAM_EMB), m->B, c.T, c.E, c.V, m->gbe));
  return 0;
}

extern "C" int nvqa_backward(nvqa_model* m, int phase) {
  NVQA_CHECK(m && m->fwd_done && (m->labels || m->cfg.arch == 3), "nvqa_backward: forward with labels has not run");
  if (m->cfg.arch == 3 && (phase == NVQA_PHASE_HEAD || phase == NVQA_PHASE_ALL)) {
    NVQA_CHECK(m->logp_valid, "nvqa_backward: the log-probs of this forward were already consumed (they are differentiated in place)");
    m->logp_valid = false;
  }
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  // a head-only call (NCCL bucket overlap, callers reading the gradients between phases) must leave the multimodal block
  // final: the AxB weight gradients are deferred to the side stream only when the LSTM phase follows in this very call
  const bool defer_saved = m->defer_head;
  if (phase == NVQA_PHASE_ALL) m->defer_head = true;
  if (phase == NVQA_PHASE_ALL && m->prezero_mask != 7u) NVQA_TRY(backward_prezero(m));
  int rc = 0;
  if (phase == NVQA_PHASE_HEAD || phase == NVQA_PHASE_ALL) rc = backward_head(m);
  m->defer_head = defer_saved;
  NVQA_TRY(rc);
  if (phase == NVQA_PHASE_LSTM || phase == NVQA_PHASE_ALL) NVQA_TRY(backward_lstm(m));
  if (phase == NVQA_PHASE_EMBED || phase == NVQA_PHASE_ALL) NVQA_TRY(backward_embed(m));
  NVQA_CHECK(phase >= 0 && phase <= 3, "bad phase");
  return 0;
}

extern "C" int nvqa_rmsprop_step(nvqa_model* m, float lr, float alpha, float eps, float wd, float clamp, float gscale) {
  NVQA_CHECK(m, "null model");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  umma_workspace_invalidate(m->ws);     // the weights change: their cached bf16 planes are stale
  ProfScope ps(m, CAT_OPT, 0);
  NVQA_TRY(drop_lookup_grad(m));
  if (m->side_opt_done) {
    // nvqa_train_step already updated the multimodal block on the side stream: encoder + embedding remain
    m->side_opt_done = false;
    const int64_t n01 = m->off_blk[2];
    return clamp_rmsprop(m->stream, m->params, m->grads, m->rms, n01, lr, alpha, eps, wd, clamp,
                         gscale * (m->cfg.arch == 1 ? m->lr_scale : 1.0f));
  }
  if (m->cfg.arch == 1 && m->lr_scale != 1.0f) {
    // gradients = join{encoder_dw * lr_scale, embedding_dw * lr_scale, multimodal_dw}, then clamp, then rmsprop
    // (003_train_ae_based_wp.lua:344-346): the scale is folded into the pre-clamp gradient scale of blocks 0 and 1
    const int64_t n01 = m->off_blk[2];
    NVQA_TRY(clamp_rmsprop(m->stream, m->params, m->grads, m->rms, n01, lr, alpha, eps, wd, clamp, gscale * m->lr_scale));
    return clamp_rmsprop(m->stream, m->params + n01, m->grads + n01, m->rms + n01, m->P - n01, lr, alpha, eps, wd, clamp, gscale);
  }
  return clamp_rmsprop(m->stream, m->params, m->grads, m->rms, m->P, lr, alpha, eps, wd, clamp, gscale);
}

// The reference's createClones shares the LookupTable weight, but not its gradWeight, with the module whose gradient
// parameters() returns (AutoEncoder_text_nostart.lua:64-66, Encoder_lstm.lua:53): literally, that gradient block stays zero.
extern "C" int nvqa_set_lookup_grad_literal(nvqa_model* m, int32_t on) {
  NVQA_CHECK(m && (m->cfg.arch == 2 || m->cfg.arch == 3), "nvqa_set_lookup_grad_literal applies to arch 2 / 3 models");
  m->lookup_grad_literal = on != 0;
  return 0;
}
extern "C" int nvqa_set_stale_h0_literal(nvqa_model* m, int32_t on) {
  NVQA_CHECK(m && m->cfg.arch == 2, "nvqa_set_stale_h0_literal applies to arch 2 models");
  m->stale_h0_literal = on != 0;
  m->stale_h0_B = 0;                 // like a freshly constructed nn.Encoder: the first step starts from zeros
  return 0;
}
static int drop_lookup_grad(nvqa_model* m) {
  if (m->lookup_grad_literal && m->glookup)
    NVQA_CUDA(cudaMemsetAsync(m->glookup, 0, (size_t)(m->cfg.V + 1) * m->cfg.E * 4, m->stream));
  return 0;
}

extern "C" int nvqa_set_variant(nvqa_model* m, int32_t fusion, float lr_scale, int32_t norm_split) {
  NVQA_CHECK(m && m->cfg.arch == 1, "nvqa_set_variant applies to arch 1 models");
  NVQA_CHECK(fusion == NVQA_FUSION_AXB || fusion == NVQA_FUSION_ASKIPB, "unknown fusion");
  NVQA_CHECK(norm_split >= 0 && norm_split < m->cfg.I && norm_split % 4 == 0, "norm_split must be a multiple of 4 inside [0, I)");
  NVQA_CHECK(lr_scale > 0.f, "lr_scale must be positive");
  m->fusion_skip = fusion == NVQA_FUSION_ASKIPB;
  m->lr_scale = lr_scale;
  m->norm_split = norm_split;
  return 0;
}

extern "C" int nvqa_adam_step(nvqa_model* m, float lr, float beta1, float beta2, float eps, float wd, float clamp, float gscale) {
  NVQA_CHECK(m && m->cfg.arch == 3 && m->adam_m, "nvqa_adam_step applies to arch 3 (text autoencoder) models");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  umma_workspace_invalidate(m->ws);
  ProfScope ps(m, CAT_OPT, 0);
  NVQA_TRY(drop_lookup_grad(m));
  ++m->adam_t;
  return clamp_adam(m->stream, m->params, m->grads, m->adam_m, m->rms, m->P, lr, beta1, beta2, eps, wd, clamp, gscale, m->adam_t);
}

static int d2h(nvqa_model* m, void* dst, const void* src, size_t bytes);
extern "C" int nvqa_logprobs_get(nvqa_model* m, int32_t step, float* dst) {
  NVQA_CHECK(m && dst && m->cfg.arch == 3 && m->fwd_done && m->logp_valid,
             "nvqa_logprobs_get: arch 3 forward has not run (or backward consumed the log-probs)");
  NVQA_CHECK(step >= 0 && step <= m->steps, "decoder step out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  if (m->lp_raw) {
    // the logits were not rewritten: log-probs of this step's rows = logits - lse, formed on demand
    const int V1 = m->cfg.V + 1;
    if (!m->lp_scratch) NVQA_TRY(dallocT(m, &m->lp_scratch, (size_t)m->cfg.B * V1));
    NVQA_TRY(lm_logprobs(m->stream, m->logits + (int64_t)step * m->B * m->ldl, m->lse + (int64_t)step * m->B, m->B, m->ldl, V1,
                         m->lp_scratch));
    return d2h(m, dst, m->lp_scratch, (size_t)m->B * V1 * 4);
  }
  NVQA_CUDA(cudaMemcpy2DAsync(dst, (size_t)(m->cfg.V + 1) * 4, m->logits + (int64_t)step * m->B * m->ldl, (size_t)m->ldl * 4,
                              (size_t)(m->cfg.V + 1) * 4, m->B, cudaMemcpyDeviceToHost, m->stream));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}

// ---- results -------------------------------------------------------------------------------------
static int d2h(nvqa_model* m, void* dst, const void* src, size_t bytes) {
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, m->stream));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}
extern "C" int nvqa_scores_get(nvqa_model* m, float* dst) {
  NVQA_CHECK(m && dst && m->fwd_done, "forward has not run");
  return d2h(m, dst, m->scores, (size_t)m->B * m->cfg.O * 4);
}
extern "C" int nvqa_argmax_get(nvqa_model* m, int32_t* dst) {
  NVQA_CHECK(m && dst && m->fwd_done, "forward has not run");
  return d2h(m, dst, m->argmax, (size_t)m->B * 4);
}
extern "C" int nvqa_state_get(nvqa_model* m, float* dst) {
  NVQA_CHECK(m && dst && m->fwd_done, "forward has not run");
  return d2h(m, dst, m->state, (size_t)m->B * m->S * 4);
}

// ---- fused convenience -----------------------------------------------------------------------------
// JdJ + clamp + optimizer on the batch already set (nvqa_set_batch*), with the reference's optimizer constants.  Same result
// as nvqa_forward ; nvqa_backward(ALL) ; nvqa_rmsprop_step -- but knowing the learning rate up front lets the multimodal
// block's clamp + RMSprop follow its weight gradients on the side stream, beside the LSTM backward.
extern "C" int nvqa_train_step(nvqa_model* m, float lr, uint64_t seed) {
  NVQA_CHECK(m, "null model");
  m->fused_step = true;
  const int frc = nvqa_forward(m, NVQA_MODE_TRAIN, seed);
  m->fused_step = false;
  NVQA_TRY(frc);
  if (m->cfg.arch == 3) {
    NVQA_TRY(nvqa_backward(m, NVQA_PHASE_ALL));
    // grad_clip 0.1, weight_decay 1e-6, adam(alpha .8, beta .999, eps 1e-8)  (001_train_arch1_text_autoencoder.lua:35,40-45,237-243)
    return nvqa_adam_step(m, lr, 0.8f, 0.999f, 1e-8f, 1e-6f, 0.1f, 1.f);
  }
  // clamp(-10,10) (:329) and optim.rmsprop defaults alpha=.99, eps=1e-8 (:408)
  // arch2 trains with optimize.weightDecay = 1e-4 (003_train_vqa_arch2/002_train_baseline.lua:197)
  const float wd = m->cfg.arch == 2 ? 1e-4f : 0.f;
  static int side_opt = -1;
  if (side_opt < 0) { const char* e = getenv("NVQA_SIDE_OPT"); side_opt = e ? atoi(e) : 1; }
  m->side_opt_done = false;
  if (side_opt && m->cfg.arch == 1 && aux_usable(m) && m->use_persistent) {
    m->side_opt.armed = true;
    m->side_opt.lr = lr; m->side_opt.alpha = 0.99f; m->side_opt.eps = 1e-8f; m->side_opt.wd = wd; m->side_opt.clamp = 10.f;
    m->side_opt.gscale = 1.f;
  }
  const int rc = nvqa_backward(m, NVQA_PHASE_ALL);
  m->side_opt.armed = false;
  NVQA_TRY(rc);
  return nvqa_rmsprop_step(m, lr, 0.99f, 1e-8f, wd, 10.f, 1.f);
}

extern "C" int nvqa_train_step_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7,
                                    const int32_t* labels, int32_t B, float lr, uint64_t seed, float* loss_out) {
  NVQA_CHECK(m, "null model");
  NVQA_CHECK(labels || m->cfg.arch == 3, "labels required");
  NVQA_TRY(nvqa_set_batch_host(m, q, len, fc7, labels, B));
  NVQA_TRY(nvqa_train_step(m, lr, seed));
  if (loss_out) NVQA_TRY(nvqa_loss(m, loss_out));
  return 0;
}

extern "C" int nvqa_eval_step_host(nvqa_model* m, const int32_t* q, const int32_t* len, const float* fc7, int32_t B,
                                   int32_t* answers_out) {
  NVQA_CHECK(answers_out, "null output");
  NVQA_TRY(nvqa_set_batch_host(m, q, len, fc7, nullptr, B));
  NVQA_TRY(nvqa_forward(m, NVQA_MODE_EVAL, 0));
  NVQA_CUDA(cudaMemcpyAsync(m->ans_host, m->argmax, (size_t)B * 4, cudaMemcpyDeviceToHost, m->stream));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  std::memcpy(answers_out, m->ans_host, (size_t)B * 4);
  return 0;
}

// ---- module-level pieces ---------------------------------------------------------------------------
extern "C" int nvqa_lstm_cell_forward(nvqa_model* m, const float* state, const float* x, const float* masks, int32_t n,
                                      float* state_out) {
  NVQA_CHECK(m && state && x && state_out, "null argument");
  const nvqa_config& c = m->cfg;
  NVQA_CHECK(n > 0 && (int64_t)n <= (int64_t)c.B * c.T, "row count out of range");
  NVQA_CUDA(cudaSetDevice(c.device));
  const int E = c.E, H = c.H, L = c.L, S = m->S;
  cudaStream_t s = m->stream;
  umma_workspace_new_forward(m->ws);      // scratch operands are rewritten by every call: no stale cached planes
  m->fwd_done = false;                    // ... which also drops the planes a pending nvqa_backward would read
  float* pre = m->da;                     // scratch [n x 4H]
  float* xin = m->dxbuf;                  // scratch [n x H] (dropped h of the layer below)
  Drop d;
  d.mask = masks; d.key = 0; d.thresh = 0; d.scale = 1.f; d.mode = masks ? 1 : 0;
  for (int l = 0; l < L; ++l) {
    const float* X = l == 0 ? x : xin;
    const int in = l == 0 ? E : H;
    NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, 4 * H, in, X, in, m->lw[l].Wi, in, pre, 4 * H, false, m->lw[l].bi, m->lw[l].bh));
    NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, 4 * H, H, state + (2 * l + 1) * H, S, m->lw[l].Wh, H, pre, 4 * H, true));
    Drop dl = d;
    if (masks) dl.mask = masks + (int64_t)l * n * H;
    NVQA_TRY(lstm_gates_fwd(s, pre, state + 2 * l * H, S, state_out + 2 * l * H, state_out + (2 * l + 1) * H, S,
                            l + 1 < L ? xin : nullptr, nullptr, dl, 0, 1, n, H));
  }
  return 0;
}

// netdef.AxB():forward({q, i}) (misc/netdef.lua:6-14) on device pointers: q [n x 2LH], i [n x I] -> out [n x C].
// masks_q / masks_i: explicit Dropout multipliers or NULL (evaluate mode).
extern "C" int nvqa_axb_forward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                                int32_t n, float* out) {
  NVQA_CHECK(m && q && i && out, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "AxB belongs to arch 1");
  NVQA_CHECK(n > 0 && n <= m->cfg.B, "row count out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  const nvqa_config& c = m->cfg;
  const int S = m->S;
  cudaStream_t s = m->stream;
  umma_workspace_new_forward(m->ws);       // qd / vd are rewritten here: their cached bf16 planes (class 2) are stale
  m->fwd_done = false;
  Drop dq, di, none;
  dq.mask = masks_q; dq.key = 0; dq.thresh = 0; dq.scale = 1.f; dq.mode = masks_q ? 1 : 0;
  di = dq; di.mask = masks_i; di.mode = masks_i ? 1 : 0;
  none = dq; none.mask = nullptr; none.mode = 0;
  NVQA_TRY(mask_copy(s, q, S, nullptr, m->qd, dq, n, S));
  NVQA_TRY(mask_copy(s, i, c.I, nullptr, m->vd, di, n, c.I));
  NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, c.C, S, m->qd, S, m->Wq, S, m->qc, c.C, false, m->bq));
  NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, c.C, c.I, m->vd, c.I, m->Wv, c.I, m->ic, c.C, false, m->bv));
  NVQA_TRY(fuse_fwd(s, m->qc, m->ic, out, none, n, c.C, m->fusion_skip));
  return 0;
}

// ---- module-level backward: the reference's nn.Module protocol, one call per module:backward ----------------------
// Scratch of the module-level calls (allocated on first use; rows <= cfg.B): per layer the gates [n x 4H], c' / h'
// [n x H] and the (dropped) input of the layer [n x max(E,H)]; da [n x 4H]; dc / dh carries; a dy copy for the embedding.
static int mod_scratch(nvqa_model* m) {
  if (m->mod_gates[0]) return 0;
  const nvqa_config& c = m->cfg;
  const size_t n = (size_t)c.B, W = (size_t)std::max(c.E, c.H);
  for (int l = 0; l < c.L; ++l) {
    NVQA_TRY(dallocT(m, &m->mod_gates[l], n * 4 * c.H));
    NVQA_TRY(dallocT(m, &m->mod_c[l], n * c.H));
    NVQA_TRY(dallocT(m, &m->mod_h[l], n * c.H));
    NVQA_TRY(dallocT(m, &m->mod_x[l], n * W));
  }
  NVQA_TRY(dallocT(m, &m->mod_da, n * 4 * c.H));
  NVQA_TRY(dallocT(m, &m->mod_dx, n * W));
  NVQA_TRY(dallocT(m, &m->mod_cprev, n * c.H));
  NVQA_TRY(dallocT(m, &m->mod_dc, n * c.H));
  NVQA_TRY(dallocT(m, &m->mod_dy, (size_t)c.B * c.T * c.E));
  return 0;
}
static Drop explicit_drop(const float* mask) {
  Drop d;
  d.mask = mask; d.key = 0; d.thresh = 0; d.scale = 1.f; d.mode = mask ? 1 : 0;
  return d;
}

// zeroGradParameters of one flat block (encoder_dw_q:zero() etc., 002_train_baseline.lua:283-285)
extern "C" int nvqa_grads_zero(nvqa_model* m, int block) {
  NVQA_CHECK(m && block >= 0 && block < 3, "bad argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaMemsetAsync(m->grads + m->off_blk[block], 0, (size_t)m->n_blk[block] * 4, m->stream));
  return 0;
}

// embedding_net_q:forward(onehot) (002_train_baseline.lua:141-144,300): Linear(V,E) on one-hot rows + Dropout + Tanh.
// words [n] = the 1-based column of the 1 in every one-hot row (the packed vector of sort_encoding_onehot_right_align).
extern "C" int nvqa_embedding_forward(nvqa_model* m, const int32_t* words, const float* mask, int32_t n, float* y) {
  NVQA_CHECK(m && words && y, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "the one-hot embedding Sequential belongs to arch 1");
  NVQA_CHECK(n > 0 && (int64_t)n <= (int64_t)m->cfg.B * m->cfg.T, "row count out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  return embed_fwd(m->stream, words, nullptr, m->WeT, m->be, y, explicit_drop(mask), n, 1, m->cfg.E, m->cfg.V);
}

// embedding_net_q:backward(onehot, dy) (:320): Tanh / Dropout backward, then accGradParameters of the one-hot Linear
// (ACCUMULATES into the embedding gradient block; its gradInput [n x V] is discarded by the reference and not computed)
extern "C" int nvqa_embedding_backward(nvqa_model* m, const int32_t* words, const float* y, const float* dy, const float* mask,
                                       int32_t n) {
  NVQA_CHECK(m && words && y && dy, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "the one-hot embedding Sequential belongs to arch 1");
  NVQA_CHECK(n > 0 && (int64_t)n <= (int64_t)m->cfg.B * m->cfg.T, "row count out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_TRY(mod_scratch(m));
  const int E = m->cfg.E;
  NVQA_CUDA(cudaMemcpyAsync(m->mod_dy, dy, (size_t)n * E * 4, cudaMemcpyDeviceToDevice, m->stream));
  NVQA_TRY(embed_bwd(m->stream, words, nullptr, y, m->mod_dy, m->gWeT, explicit_drop(mask), n, 1, E, m->cfg.V));
  return colsum(m->stream, m->mod_dy, n, E, E, m->gbe, nullptr);
}

// LSTM.lstm_conventional():backward({state, x}, dstate_out) of one timestep clone (misc/LSTM.lua:12-73; called by
// rnn_backward, misc/RNNUtils.lua:195-196).  The nngraph clone keeps its forward internals; this entry point is
// stateless and recomputes them from (state, x, masks).  dstate [n x 2LH] and dx [n x E] are the gradInputs;
// the parameter gradients are ACCUMULATED into the encoder block (sum over clones, :323-326).
extern "C" int nvqa_lstm_cell_backward(nvqa_model* m, const float* state, const float* x, const float* masks,
                                       const float* dstate_out, int32_t n, float* dstate, float* dx) {
  NVQA_CHECK(m && state && x && dstate_out && dstate && dx, "null argument");
  const nvqa_config& c = m->cfg;
  NVQA_CHECK(c.arch == 1, "the module-level cell belongs to arch 1");
  NVQA_CHECK(n > 0 && n <= c.B, "row count out of range");
  NVQA_CUDA(cudaSetDevice(c.device));
  NVQA_TRY(mod_scratch(m));
  umma_workspace_new_forward(m->ws);       // scratch operands are rewritten by every call: no stale cached planes
  m->fwd_done = false;
  const int E = c.E, H = c.H, L = c.L, S = m->S;
  cudaStream_t s = m->stream;
  // forward internals
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? E : H;
    const float* X = l == 0 ? x : m->mod_x[l];
    NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, 4 * H, in, X, in, m->lw[l].Wi, in, m->mod_gates[l], 4 * H, false, m->lw[l].bi,
                  m->lw[l].bh));
    NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, 4 * H, H, state + (2 * l + 1) * H, S, m->lw[l].Wh, H, m->mod_gates[l], 4 * H, true));
    Drop dl = explicit_drop(masks ? masks + (int64_t)l * n * H : nullptr);
    NVQA_TRY(lstm_gates_fwd(s, m->mod_gates[l], state + 2 * l * H, S, m->mod_c[l], m->mod_h[l], H,
                            l + 1 < L ? m->mod_x[l + 1] : nullptr, nullptr, dl, 0, 1, n, H));
  }
  // backward, top layer first; h'_l feeds the output state and layer l+1 (gradients summed at the fan-out)
  for (int l = L - 1; l >= 0; --l) {
    const int in = l == 0 ? E : H;
    const float* X = l == 0 ? x : m->mod_x[l];
    NVQA_CUDA(cudaMemcpy2DAsync(m->mod_cprev, (size_t)H * 4, state + 2 * l * H, (size_t)S * 4, (size_t)H * 4, n,
                                cudaMemcpyDeviceToDevice, s));
    const float* dh_above = l + 1 < L ? m->mod_dx : nullptr;
    Drop dab = explicit_drop(masks && l + 1 < L ? masks + (int64_t)l * n * H : nullptr);
    NVQA_TRY(lstm_gates_bwd(s, m->mod_gates[l], m->mod_cprev, m->mod_c[l], dstate_out + (2 * l + 1) * H, S, dh_above,
                            dstate_out + 2 * l * H, S, m->mod_da, m->mod_dc, nullptr, dab, 0, 1, n, H));
    NVQA_CUDA(cudaMemcpy2DAsync(dstate + 2 * l * H, (size_t)S * 4, m->mod_dc, (size_t)H * 4, (size_t)H * 4, n,
                                cudaMemcpyDeviceToDevice, s));
    // accGradParameters: dWi += da^T x, dWh += da^T h_prev, both biases += sum(da)
    NVQA_TRY(gemm(m, CAT_OTHER, false, false, 4 * H, in, n, m->mod_da, 4 * H, X, in, m->lg[l].Wi, in, true));
    NVQA_TRY(gemm(m, CAT_OTHER, false, false, 4 * H, H, n, m->mod_da, 4 * H, state + (2 * l + 1) * H, S, m->lg[l].Wh, H, true));
    NVQA_TRY(colsum(s, m->mod_da, n, 4 * H, 4 * H, m->lg[l].bi, m->lg[l].bh));
    // gradInputs: dh_prev = da . Wh (into the packed dstate), dx_l = da . Wi
    NVQA_TRY(gemm(m, CAT_OTHER, true, false, n, H, 4 * H, m->mod_da, 4 * H, m->lw[l].Wh, H, dstate + (2 * l + 1) * H, S, false));
    NVQA_TRY(gemm(m, CAT_OTHER, true, false, n, in, 4 * H, m->mod_da, 4 * H, m->lw[l].Wi, in, l == 0 ? dx : m->mod_dx, in, false));
  }
  return 0;
}

// netdef.AxB():backward({q, i}, dout) (misc/netdef.lua:6-14): recomputes the forward internals, returns dq [n x 2LH]
// (and di [n x I] when di != NULL -- the reference computes and discards it, 002_train_baseline.lua:312-313) and
// ACCUMULATES the gradients of Wq, bq, Wi, bi into the multimodal block.
extern "C" int nvqa_axb_backward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                                 const float* dout, int32_t n, float* dq, float* di) {
  NVQA_CHECK(m && q && i && dout && dq, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "AxB belongs to arch 1");
  NVQA_CHECK(n > 0 && n <= m->cfg.B, "row count out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  const nvqa_config& c = m->cfg;
  const int S = m->S, C = c.C, I = c.I;
  cudaStream_t s = m->stream;
  umma_workspace_new_forward(m->ws);
  m->fwd_done = false;
  Drop dq_ = explicit_drop(masks_q), di_ = explicit_drop(masks_i), none = explicit_drop(nullptr);
  NVQA_TRY(mask_copy(s, q, S, nullptr, m->qd, dq_, n, S));
  NVQA_TRY(mask_copy(s, i, I, nullptr, m->vd, di_, n, I));
  NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, C, S, m->qd, S, m->Wq, S, m->qc, C, false, m->bq));
  NVQA_TRY(gemm(m, CAT_OTHER, true, true, n, C, I, m->vd, I, m->Wv, I, m->ic, C, false, m->bv));
  NVQA_TRY(fuse_fwd(s, m->qc, m->ic, m->zd, none, n, C, m->fusion_skip));
  // CMulTable / Tanh backward
  NVQA_TRY(fuse_bwd(s, dout, m->qc, m->ic, m->dqpre, m->dipre, none, n, C, m->fusion_skip));
  NVQA_TRY(gemm(m, CAT_OTHER, false, false, C, S, n, m->dqpre, C, m->qd, S, m->gWq, S, true));
  NVQA_TRY(colsum(s, m->dqpre, n, C, C, m->gbq, nullptr));
  NVQA_TRY(gemm(m, CAT_OTHER, false, false, C, I, n, m->dipre, C, m->vd, I, m->gWv, I, true));
  NVQA_TRY(colsum(s, m->dipre, n, C, C, m->gbv, nullptr));
  NVQA_TRY(gemm(m, CAT_OTHER, true, false, n, S, C, m->dqpre, C, m->Wq, S, dq, S, false));
  NVQA_TRY(mask_inplace(s, dq, dq_, (int64_t)n * S));
  if (di) {
    NVQA_TRY(gemm(m, CAT_OTHER, true, false, n, I, C, m->dipre, C, m->Wv, I, di, I, false));
    NVQA_TRY(mask_inplace(s, di, di_, (int64_t)n * I));
  }
  return 0;
}

// multimodal_net = Sequential{AxB, Dropout, Linear(C, O)} (002_train_baseline.lua:151-154): forward({tv_q, fv_im})
extern "C" int nvqa_multimodal_forward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                                       const float* masks_z, int32_t n, float* scores) {
  NVQA_CHECK(m && q && i && scores, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "multimodal_net belongs to arch 1");
  NVQA_CHECK(n > 0 && n <= m->cfg.B, "row count out of range");
  const nvqa_config& c = m->cfg;
  NVQA_TRY(nvqa_axb_forward(m, q, i, masks_q, masks_i, n, m->zd));
  NVQA_TRY(mask_inplace(m->stream, m->zd, explicit_drop(masks_z), (int64_t)n * c.C));
  return gemm(m, CAT_OTHER, true, true, n, c.O, c.C, m->zd, c.C, m->Wc, c.C, scores, c.O, false, m->bc);
}

// multimodal_net:backward({tv_q, fv_im}, dscores) (:312): dq [n x 2LH] (and di when non-NULL); ACCUMULATES into the whole
// multimodal gradient block
extern "C" int nvqa_multimodal_backward(nvqa_model* m, const float* q, const float* i, const float* masks_q, const float* masks_i,
                                        const float* masks_z, const float* dscores, int32_t n, float* dq, float* di) {
  NVQA_CHECK(m && q && i && dscores && dq, "null argument");
  NVQA_CHECK(m->cfg.arch == 1, "multimodal_net belongs to arch 1");
  NVQA_CHECK(n > 0 && n <= m->cfg.B, "row count out of range");
  const nvqa_config& c = m->cfg;
  const int C = c.C, O = c.O;
  cudaStream_t s = m->stream;
  // forward internals up to the classifier input zd
  NVQA_TRY(nvqa_axb_forward(m, q, i, masks_q, masks_i, n, m->zd));
  NVQA_TRY(mask_inplace(s, m->zd, explicit_drop(masks_z), (int64_t)n * C));
  // Linear(C, O) backward
  NVQA_TRY(gemm(m, CAT_OTHER, false, false, O, C, n, dscores, O, m->zd, C, m->gWc, C, true));
  NVQA_TRY(colsum(s, dscores, n, O, O, m->gbc, nullptr));
  NVQA_TRY(gemm(m, CAT_OTHER, true, false, n, C, O, dscores, O, m->Wc, C, m->dzd, C, false));
  NVQA_TRY(mask_inplace(s, m->dzd, explicit_drop(masks_z), (int64_t)n * C));
  return nvqa_axb_backward(m, q, i, masks_q, masks_i, m->dzd, n, dq, di);
}

// optim.rmsprop's update (misc/rmsprop_lrscale.lua:26-34, lrs = 1) on arbitrary device vectors:
// g' = clamp(g * grad_scale) + wd * x;  m = alpha m + (1 - alpha) g'^2;  x -= lr g' / (sqrt(m) + eps)
extern "C" int nvqa_rmsprop_vector(nvqa_model* m, float* x, const float* g, float* state_m, int64_t n, float lr, float alpha,
                                   float eps, float wd, float clamp, float grad_scale) {
  NVQA_CHECK(m && x && g && state_m && n > 0, "bad argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  return clamp_rmsprop(m->stream, x, const_cast<float*>(g), state_m, n, lr, alpha, eps, wd, clamp, grad_scale);
}

extern "C" int nvqa_cross_entropy(nvqa_model* m, const float* scores, const int32_t* labels, int32_t n, float* loss_host,
                                  float* dscores) {
  NVQA_CHECK(m && scores && labels, "null argument");
  NVQA_CHECK(n > 0 && n <= m->cfg.B, "row count out of range");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_TRY(softmax_ce(m->stream, scores, labels, dscores, m->rowloss, nullptr, n, m->cfg.O, 1.0f / (float)n));
  NVQA_TRY(loss_reduce(m->stream, m->rowloss, m->loss, n));
  if (loss_host) {
    NVQA_CUDA(cudaMemcpyAsync(m->loss_host, m->loss, 4, cudaMemcpyDeviceToHost, m->stream));
    NVQA_CUDA(cudaStreamSynchronize(m->stream));
    *loss_host = m->loss_host[0];
  }
  return 0;
}

// ---- live profile ------------------------------------------------------------------------------------
extern "C" int nvqa_profile(nvqa_model* m, int enable) {
  NVQA_CHECK(m, "null model");
  m->profiling = enable != 0;
  if (enable) for (auto& pc : m->prof) { pc.ms = 0; pc.flops = 0; pc.launches = 0; }
  return 0;
}

extern "C" int nvqa_profile_report(nvqa_model* m, char* buf, int32_t cap) {
  NVQA_CHECK(m && buf && cap > 0, "bad argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  std::string out = "[";
  for (int c = 0; c < CAT_COUNT; ++c) {
    ProfCat& pc = m->prof[c];
    for (auto& ev : pc.pending) {
      float ms = 0.f;
      NVQA_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
      pc.ms += ms;
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
    pc.pending.clear();
    char line[256];
    snprintf(line, sizeof line, "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e}",
             c ? ", " : "", kCatName[c], (long long)pc.launches, pc.ms, pc.flops);
    out += line;
  }
  out += "]";
  NVQA_CHECK((int)out.size() + 1 <= cap, "report buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

// ---- utilities ---------------------------------------------------------------------------------------
extern "C" int nvqa_host_alloc(void** p, int64_t bytes) {
  NVQA_CHECK(p && bytes >= 0, "bad argument");
  NVQA_CUDA(cudaMallocHost(p, (size_t)(bytes ? bytes : 16)));
  return 0;
}
extern "C" int nvqa_host_free(void* p) { if (p) NVQA_CUDA(cudaFreeHost(p)); return 0; }
extern "C" int nvqa_device_alloc(void** p, int64_t bytes) {
  NVQA_CHECK(p && bytes >= 0, "bad argument");
  NVQA_CUDA(cudaMalloc(p, (size_t)(bytes ? bytes : 16)));
  return 0;
}
extern "C" int nvqa_device_free(void* p) { if (p) NVQA_CUDA(cudaFree(p)); return 0; }
extern "C" int nvqa_memcpy_h2d(nvqa_model* m, void* dst, const void* src, int64_t bytes) {
  NVQA_CHECK(m && dst && src, "null argument");
  NVQA_CUDA(cudaSetDevice(m->cfg.device));
  NVQA_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, m->stream));
  NVQA_CUDA(cudaStreamSynchronize(m->stream));
  return 0;
}
extern "C" int nvqa_memcpy_d2h(nvqa_model* m, void* dst, const void* src, int64_t bytes) {
  NVQA_CHECK(m && dst && src, "null argument");
  return d2h(m, dst, src, (size_t)bytes);
}

extern "C" int nvqa_gemm_test_ex(int precision, int a_kmajor, int b_kmajor, int32_t M, int32_t N, int32_t K, const float* A,
                                 const float* B, float* C, int32_t ldc, int beta, const float* bias0, const float* bias1,
                                 void* stream) {
  NVQA_CHECK(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int lda = a_kmajor ? K : M, ldb = b_kmajor ? K : N;
  if (precision == NVQA_PREC_FP32_SIMT) {
    NVQA_TRY(simt_gemm(s, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, beta != 0, bias0, bias1));
    NVQA_CUDA(cudaStreamSynchronize(s));
    return 0;
  }
  int planes = precision == NVQA_PREC_BF16X3 ? 3 : precision == NVQA_PREC_BF16X2 ? 2 : 1;
  UmmaWorkspace* ws = nullptr;
  NVQA_TRY(umma_workspace_create(&ws, ((size_t)(M + 64) * (K + 64) + (size_t)(N + 64) * (K + 64)) * 2 * 3 + (8 << 20), 0));
  int r = umma_gemm(s, planes, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, beta != 0, bias0, bias1, ws, false, false);
  cudaStreamSynchronize(s);
  umma_workspace_destroy(ws);
  return r;
}

extern "C" int nvqa_gemm_test(int precision, int a_kmajor, int b_kmajor, int32_t M, int32_t N, int32_t K, const float* A,
                              const float* B, float* C, void* stream) {
  return nvqa_gemm_test_ex(precision, a_kmajor, b_kmajor, M, N, K, A, B, C, N, 0, nullptr, nullptr, stream);
}
