// Host-side state of the tcgen05 GEMM engine: bf16 operand planes (arena + per-step cache of weight planes) and
// cached TMA tensor maps.  Shared by umma_gemm.cu and the persistent LSTM kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <unordered_map>

#include "common.cuh"

namespace nvqa {

struct PlaneKey {
  const void* src; int rows, K, ld, kmajor, P;
  bool operator==(const PlaneKey& o) const {
    return src == o.src && rows == o.rows && K == o.K && ld == o.ld && kmajor == o.kmajor && P == o.P;
  }
};
struct PlaneKeyHash {
  size_t operator()(const PlaneKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.src);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.K; h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)(k.kmajor * 8 + k.P);
    return h;
  }
};
struct MapKey {
  const void* planes; int rows, Kp, P, box; long long plane_stride; int box_planes;
  bool operator==(const MapKey& o) const {
    return planes == o.planes && rows == o.rows && Kp == o.Kp && P == o.P && box == o.box && plane_stride == o.plane_stride &&
           box_planes == o.box_planes;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.planes);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.Kp; h = h * 1000003u ^ (size_t)(k.P * 1024 + k.box + k.box_planes * 65536);
    return h;
  }
};

struct UmmaWorkspace {
  uint8_t* base = nullptr;
  size_t bytes = 0;
  size_t static_bytes = 0;      // [0, static_bytes): cached planes of static operands (weights)
  size_t static_top = 0;
  size_t act_bytes = 0;         // [static_bytes, static_bytes + act_bytes): planes of this forward's activations, so
  size_t act_top = 0;           // that the backward GEMMs (wgrad operands) reuse them instead of splitting again
  size_t trans_top = 0;         // transient planes live in [static_bytes + act_bytes, bytes - side_bytes), reset per GEMM
  // GEMMs enqueued on a SECOND stream (the image branch beside the persistent LSTM kernels, model.cu) must not share the
  // transient arena with the main stream's GEMMs: while `side` is set, transients come from [bytes - side_bytes, bytes)
  // and the persistent grid is capped at cta_cap CTAs (the SMs the 128-CTA recurrent kernel leaves free)
  size_t side_bytes = 0, side_top = 0;
  bool side = false;
  int cta_cap = 0;
  // Split-K reductions whose result nobody on the launching stream waits for (the weight-gradient GEMMs of the LSTM
  // layers) can run on a second stream beside the NEXT GEMM: while reduce_stream is set, the partials of such a GEMM go
  // to one of two dedicated slots at the end of the buffer and the reduction kernel is enqueued on reduce_stream
  // (event-ordered behind the GEMM; a slot is reused only after its previous reduction has finished).
  // hint for the persistent LSTM launchers: the step-barrier counters they were handed are already zero (cleared together
  // with everything else at the start of the pass); consumed (reset) by the launcher, valid for its first batch window
  bool ctr_zeroed = false;
  cudaStream_t reduce_stream = nullptr;
  cudaEvent_t reduce_gemm_done[2] = {nullptr, nullptr}, reduce_done[2] = {nullptr, nullptr};
  bool reduce_pending[2] = {false, false};
  int reduce_slot = 0;
  size_t defer_bytes = 0;       // 2 slots of defer_bytes / 2 in [bytes, bytes + defer_bytes)
  size_t& ttop() { return side ? side_top : trans_top; }
  size_t tbase() const { return side ? bytes - side_bytes : static_bytes + act_bytes; }
  size_t tlimit() const { return side ? bytes : bytes - side_bytes; }
  std::unordered_map<PlaneKey, __nv_bfloat16*, PlaneKeyHash> cache;
  std::unordered_map<PlaneKey, __nv_bfloat16*, PlaneKeyHash> act_cache;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
};


// fp32 row-major [rows x cols] (ld) -> bf16 planes [P][rows][pitch], pitch = cols rounded up to 8
// cache_class: 0 = transient, 1 = weight (cached until umma_workspace_invalidate), 2 = activation written once per
// forward (cached until umma_workspace_new_forward)
int prepare_planes(UmmaWorkspace* ws, cudaStream_t s, int P, const float* src, int rows, int cols, int ld,
                   int cache_class, __nv_bfloat16** out, int* pitch_out);
void umma_workspace_new_forward(UmmaWorkspace* ws);
int reserve_planes(UmmaWorkspace* ws, int P, const float* src, int rows, int K, int ld, int cache_class, __nv_bfloat16** out,
                   int* pitch_out);
int presplit_weights(UmmaWorkspace* ws, cudaStream_t s, int P, const float* const* src, const int* rows, const int* K, int n);
// tensor map over planes [P][rows][pitch]: box = 64 columns (128 B, SWIZZLE_128B) x box_rows x 1 plane
// (rows = bound of the row coordinate; plane_stride in elements, 0 = rows * pitch)
// box_planes > 1: one TMA instruction fetches the same [box_rows x 64] tile of that many consecutive planes (they land
// back to back in shared memory)
int get_map(UmmaWorkspace* ws, const __nv_bfloat16* planes, int rows, int pitch, int P, int box_rows, CUtensorMap* out,
            long long plane_stride = 0, int box_planes = 1);

// 4-D view (64 columns of a k-block, rows, planes, k-blocks) of the same planes: ONE TMA instruction fetches box_kb
// consecutive k-blocks of a [box_rows x 64] tile of all P planes; they land as [k-block][plane][row][128 B], i.e. as
// box_kb consecutive stages of the 3-D boxes above (tools/probes/probe_tma_box.cu: the driver accepts the k-block stride
// of 128 B and the bytes land identically).  pitch must be a multiple of 64.
int get_map_kb(UmmaWorkspace* ws, const __nv_bfloat16* planes, int rows, int pitch, int P, int box_rows, int box_kb,
               CUtensorMap* out, long long plane_stride = 0);

// A GEMM operand: either an fp32 row-major source (split into planes by the engine) or planes made upstream.
struct UmmaOperand {
  const float* src = nullptr;             // fp32 [rows x cols], leading dimension ld
  int ld = 0;
  int is_static = 0;                      // cache class of the planes: 0 none, 1 weight, 2 activation of this forward
  const __nv_bfloat16* planes = nullptr;  // ready-made planes [P][plane_rows][pitch] ...
  int plane_rows = 0, pitch = 0;
  int row_offset = 0;                     // ... of which this operand starts at row row_offset
  bool kmajor = true;                     // contraction index is the contiguous one
};
int umma_gemm_ops(cudaStream_t s, int planes, const UmmaOperand& A, const UmmaOperand& B, int M, int N, int K, float* C,
                  int ldc, bool beta, const float* bias0, const float* bias1, UmmaWorkspace* ws);

}  // namespace nvqa
