// Persistent multi-timestep LSTM layer kernels (lstm_persistent.cu).
#pragma once
#include "umma_engine.cuh"

namespace nvqa {

// All T steps of one LSTM layer's recurrence in one cooperative launch (see lstm_persistent.cu).
//   pre  [T][B][4H]   in: x.Wi^T + bi + bh (batched input projection)   out: gates i,f,o,g (post-activation)
//   c,h  [(T+1)][B][H] fp32 states, slot 0 = zeros, slot t+1 written by step t
//   hp   [P][(T+1)B][H] bf16 planes of h (slot 0 zeros): A operand of the next step and of the wgrad GEMM
//   xdrop_next [T][B][H] = Dropout(h_t) for the layer above (nullptr for the top layer)
// Returns 0 on success, -1 when the shape/precision is not supported (caller falls back to per-step kernels),
// > 0 on error.
// c slot 0 / hp slot 0 hold the INITIAL state (zeros, or the last slot of a preceding segment of the same buffers);
// hp_plane_rows = rows between consecutive planes of hp (0: (T+1)*B).
int lstm_fwd_persistent(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, float* pre, float* c, float* h,
                        __nv_bfloat16* hp, long long hp_plane_rows, float* xdrop_next, const int32_t* len, Drop d, int T,
                        int B, int H, unsigned int* counter);

// Second generation (lstm_persistent_v2.cu): W_hh slice resident in shared memory (plane 0) AND tensor memory (plane 1),
// 128 gate rows per CTA, 64-row h tiles; same contract, returns -1 for shapes it does not cover.
// xdrop_planes (optional): Dropout(h_t) for the layer above is written as bf16 planes [P][T*B][H] (xdrop_plane_stride
// elements between planes) INSTEAD of the fp32 xdrop_next (which must still be non-null to request that output at all).
// lstm_fwd_v2_supported: will this generation take the shape (so that the caller may reserve the planes beforehand)?
bool lstm_fwd_v2_supported(int P, int H);
int lstm_fwd_persistent_v2(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, float* pre, float* c, float* h,
                           __nv_bfloat16* hp, long long hp_plane_rows, float* xdrop_next, const int32_t* len, Drop d, int T,
                           int B, int H, unsigned int* counter, __nv_bfloat16* xdrop_planes = nullptr,
                           long long xdrop_plane_stride = 0);

int lstm_bwd_persistent_v2(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, const float* gates, const float* c,
                           const float* dh0, const float* dc0, int ld0, const float* dh_above, Drop d, float* dasum,
                           __nv_bfloat16* dap, long long dap_plane_rows, float* dhbuf, float* dh_init, float* dc_init,
                           const int32_t* len, int T, int B, int H, unsigned int* counter);

// All T steps of one layer's backward recurrence (see lstm_persistent.cu).
//   gates [T][B][4H] post-activation, c [(T+1)][B][H]; dh0/dc0 (leading dim ld0): d(final state) of this layer;
//   dh_above [T][B][H]: dX of the layer above (masked by its input Dropout) or nullptr;
//   out: dap [P][T*B][4H] bf16 planes of da (dap_plane_rows rows between planes, 0: T*B) and dasum [B][4H] fp32 =
//   sum over t of da_t per batch row (its column sums are the bias gradients);
//   scratch dhbuf [2][4][B][H]; dh_init / dc_init [B][H] (both or neither): gradient w.r.t. the initial state.
int lstm_bwd_persistent(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, const float* gates, const float* c,
                        const float* dh0, const float* dc0, int ld0, const float* dh_above, Drop d, float* dasum,
                        __nv_bfloat16* dap, long long dap_plane_rows, float* dhbuf, float* dh_init, float* dc_init,
                        const int32_t* len, int T, int B, int H, unsigned int* counter);

}  // namespace nvqa
