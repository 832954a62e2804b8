// Persistent multi-timestep LSTM layer kernels (sm_100a): ONE cooperative launch runs all T steps of a layer.
//
// Forward (K4): replaces, per layer, T x { h2h nn.Linear + CAddTable + 2 Narrow/Sigmoid/Tanh chains + 3 CMulTable +
// CAddTable } of the reference cell (002_train_vqa_arch1/misc/LSTM.lua:42-59) and the rnn_forward driver loop
// (misc/RNNUtils.lua:128-154).
//
//   grid  = (H/16 hidden-unit slices) x (ceil(B/128) batch tiles), one CTA per SM, all co-resident
//   smem  = the CTA's slice of W_hh (4 gates x 16 units = 64 rows x H, P bf16 planes) RESIDENT for all T steps
//           + a ring of A stages (h_{t-1} tile: 128 rows x 64 k x P planes) filled by TMA
//   TMEM  = one 128 x 64 fp32 accumulator: columns ordered [i f o g] x 8 units, twice -> every epilogue thread
//           (= one batch row) holds all four gates of its units: the gate math is thread-local
//   per step: TMA(h_{t-1}) -> tcgen05.mma (P=2: 3 MMAs per k16) -> tcgen05.ld -> + input projection -> sigmoid/tanh,
//           cell update -> gates, c_t, h_t (fp32), h_t as bf16 planes (next step's A operand, and the wgrad operand),
//           Dropout(h_t) for the layer above -> grid barrier (one release-add per CTA, acquire-spin by the
//           TMA producer thread only).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "lstm_persistent.cuh"
#include "umma_ptx.cuh"

namespace nvqa {

constexpr int LP_THREADS = 320;          // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr int LP_EPI_THREADS = 256;
constexpr int LP_MAX_STAGES = 4;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// (debug) per-CTA wall-clock stamps at steps 10 and 11: dbg2[cta*4 + {0: barrier seen open t=10, 1: arrived t=10, 2: open t=11, 3: arrived t=11}]
#define LP_GSTAMP(t_, k_) \
  do { if (dbg && ((t_) == 10 || (t_) == 11)) dbg[T * 8 + (blockIdx.y * gridDim.x + blockIdx.x) * 4 + ((t_) - 10) * 2 + (k_)] = (long long)gtimer(); } while (0)
#define LP_XSTAMP(t_, k_, v_) \
  do { if (dbg && (t_) == 10) dbg[T * 8 + 1024 + (blockIdx.y * gridDim.x + blockIdx.x) * 4 + (k_)] = (long long)(v_); } while (0)
#define LP_STAMP(t_, k_) \
  do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) dbg[(t_) * 8 + (k_)] = clock64(); } while (0)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Grid-wide counter barrier (backward kernel).  See group_arrive / group_wait below for why the arrive is a blocking
// atom.release and the poll an atom.acquire rather than plain loads + fences.
__device__ int g_poll_ns = 64;
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  unsigned int old;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
  if (old == 0xFFFFFFFFu) __trap();
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("atom.acquire.gpu.global.add.u32 %0, [%1], 0;" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) break;
    __nanosleep(g_poll_ns);
    if (clock64() - t0 > 4000000000LL) {
      printf("lstm_persistent: grid barrier timed out (block %d,%d,%d, have %u want %u)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, v, target);
      __trap();
    }
  }
}

// exp-based activations for the fused epilogue: ex2.approx + fast division, absolute error ~1e-7
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// Flag barrier among the CTAs of one group (here: the CTAs that share a batch tile).  Each member publishes its step
// count with a release store to its OWN word (no same-address atomic serialisation: measured 5 us of skew with a
// single counter and 128 arrivals); one warp polls all members' words with coalesced relaxed loads.
__device__ __forceinline__ void flag_arrive(unsigned int* flag, unsigned int value) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// Group counter barrier.  Arrive = one release-add per CTA; the poller reads the counter WITH AN ATOMIC: plain (even
// .acquire.gpu) loads were measured to observe a remote SM's store up to 4 us late on some SMs, atomics execute at the
// line's home L2 slice and see it within ~0.3 us.
// returns only after the add has been performed at L2 (the returned value is consumed), so that the caller can then
// let other warps flood the memory pipeline without delaying the release
__device__ __forceinline__ void group_arrive(unsigned int* ctr) {
  unsigned int old;
  asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
  if (old == 0xFFFFFFFFu) __trap();
}
__device__ __forceinline__ void group_wait(unsigned int* ctr, unsigned int target) {
  long long t0 = clock64();
  while (true) {
    unsigned int v;
    // acquire on the polling atomic itself: a trailing fence.acq_rel would also carry release semantics and was
    // measured to wait ~4 us for the SM's queued (unrelated) epilogue stores to drain
    asm volatile("atom.acquire.gpu.global.add.u32 %0, [%1], 0;" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) {
      printf("lstm_persistent: group barrier timed out (block %d,%d,%d have %u want %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, v, target);
      __trap();
    }
  }
}
// called by a full warp; returns when every member's flag >= target
__device__ __forceinline__ void flag_wait_warp(const unsigned int* flags, int members, unsigned int target) {
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  while (true) {
    bool ok = true;
    for (int i = lane; i < members; i += 32) {
      unsigned int v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      ok = ok && (v >= target);
    }
    if (__all_sync(0xffffffffu, ok)) break;
    if (clock64() - t0 > 4000000000LL) {
      if (lane == 0) printf("lstm_persistent: flag barrier timed out (block %d,%d,%d want %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, target);
      __trap();
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  __syncwarp();
}

template <int P, int CL>
__global__ void __launch_bounds__(LP_THREADS, 1)
lstm_fwd_persistent_kernel(const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapW,
                           float* __restrict__ pre, float* __restrict__ c, float* __restrict__ h,
                           __nv_bfloat16* __restrict__ hp, long long hp_plane, float* __restrict__ xdrop,
                           const int32_t* __restrict__ len, Drop drop, int T, int B, int H, int KB, int S,
                           unsigned int* counter, long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t W_TILE = 64 * 128;                 // 64 rows x 128 B, one plane, one k-block
  constexpr uint32_t A_PLANE = 128 * 128;               // 128 rows x 128 B
  const uint32_t w0 = base;                             // W: [(kb * P + p)] tiles
  const uint32_t a0 = w0 + (uint32_t)KB * P * W_TILE;   // A ring: [s][p]
  const uint32_t bar0 = a0 + (uint32_t)S * P * A_PLANE;
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * LP_MAX_STAGES, wfull = bar0 + 16 * LP_MAX_STAGES,
                 tfull = wfull + 8, gobar = wfull + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar0 - raw) + 16 * LP_MAX_STAGES + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * 16, m0 = blockIdx.y * 128;
  const unsigned int G = gridDim.x * gridDim.y;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      // with a cluster of CL CTAs sharing the A tile, a stage is free only when all CL consumers have released it
      for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CL); }
      mbar_init(wfull, 1);
      mbar_init(tfull, 1);
      mbar_init(gobar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 64);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();           // peers' barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    // ===== TMA producer (lane 0 issues; the whole warp polls the batch tile's flags) =====
    if (lane == 0) {
      // resident W_hh slice: smem row (cc*32 + g*8 + j) <- W row g*H + u0 + 8*cc + j
      mbar_expect_tx(wfull, (uint32_t)KB * P * W_TILE);
      for (int kb = 0; kb < KB; ++kb)
        for (int p = 0; p < P; ++p)
          for (int cc = 0; cc < 2; ++cc)
            for (int g = 0; g < 4; ++g)
              tma_load_3d(w0 + (uint32_t)(kb * P + p) * W_TILE + (uint32_t)(cc * 4 + g) * 1024, &mapW, wfull, kb * 64,
                          g * H + u0 + 8 * cc, p);
    }
    int it = 0;
    const int gokb = (S < KB ? S : KB) - 1;
    for (int t = 0; t < T; ++t) {
      if (t > 0) {
        // h_{t-1} rows of THIS batch tile are complete once the gridDim.x CTAs sharing it have published step t
        if (lane == 0) {
          group_wait(counter + 32 * blockIdx.y, (unsigned int)t * gridDim.x);      // one counter (own 128 B line) per batch tile
          fence_proxy_async();
        }
        __syncwarp();
      }
      if (lane == 0) {
        LP_STAMP(t, 1);
        LP_GSTAMP(t, 0);
        for (int kb = 0; kb < KB; ++kb) {
          const int s = (it + kb) % S;
          const uint32_t ph = (uint32_t)((it + kb) / S) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          mbar_expect_tx(full0 + 8 * s, P * A_PLANE);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            if (CL == 1) {
              tma_load_3d(a0 + (uint32_t)(s * P + p) * A_PLANE, &mapH, full0 + 8 * s, kb * 64, t * B + m0, p);
            } else {   // this CTA fetches rows [128/CL * rank, +128/CL) of the shared tile for the whole cluster
              constexpr uint32_t SL = 128 / CL;
              tma_load_3d_mc(a0 + (uint32_t)(s * P + p) * A_PLANE + crank * SL * 128, &mapH, full0 + 8 * s, kb * 64,
                             t * B + m0 + (int)(crank * SL), p, kMask);
            }
          }
          if (kb == gokb) mbar_arrive(gobar);     // this step's first loads are out: the epilogue may use the memory pipe
        }
      }
      it += KB;
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      mbar_wait(wfull, 0);
      int it = 0;
      for (int t = 0; t < T; ++t) {
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (kb == 0) { LP_STAMP(t, 3); LP_XSTAMP(t, 0, gtimer()); }
          if (kb == KB - 1) { LP_STAMP(t, 0); LP_XSTAMP(t, 1, gtimer()); }   // last k-block's data has landed
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t da[P], db[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
              da[p] = make_kmajor_sw128_desc(a0 + (uint32_t)(s * P + p) * A_PLANE + k * 32);
              db[p] = make_kmajor_sw128_desc(w0 + (uint32_t)(kb * P + p) * W_TILE + k * 32);
            }
            if (P >= 2) {
              umma_f16(tmem_base, da[0], db[1], idesc, acc); acc = 1;
              umma_f16(tmem_base, da[1], db[0], idesc, acc);
            }
            umma_f16(tmem_base, da[0], db[0], idesc, acc); acc = 1;
          }
          if (CL == 1) umma_commit(empty0 + 8 * s); else umma_commit_mc(empty0 + 8 * s, kMask);
        }
        umma_commit(tfull);
        LP_STAMP(t, 4);
        { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); LP_XSTAMP(t, 2, smid); }
      }
    }
  } else {
    // ===== epilogue: 8 warps; thread = (batch row, half of the CTA's 16 hidden units) =====
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int cc = (warp - 2) >> 2;               // which 32-column half: units u0 + 8 cc .. + 8
    const int b = m0 + q * 32 + lane;
    const bool rowok = b < B;
    const int mylen = rowok ? (len ? len[b] : T) : 0;
    const int uo = u0 + 8 * cc;
    float ccarry[8];                               // c_{t-1} of this thread's (row, 8 units): never leaves registers
#pragma unroll
    for (int j = 0; j < 8; ++j) ccarry[j] = 0.f;
    if (rowok) {                                   // slot 0 of c: zeros, or the initial state handed over by a previous segment
      const float4 c0a = *reinterpret_cast<const float4*>(c + (size_t)b * H + uo);
      const float4 c0b = *reinterpret_cast<const float4*>(c + (size_t)b * H + uo + 4);
      ccarry[0] = c0a.x; ccarry[1] = c0a.y; ccarry[2] = c0a.z; ccarry[3] = c0a.w;
      ccarry[4] = c0b.x; ccarry[5] = c0b.y; ccarry[6] = c0b.z; ccarry[7] = c0b.w;
    }
    for (int t = 0; t < T; ++t) {
      const bool active = rowok && (t >= T - mylen);
      const size_t rin = (size_t)t * B + b, rout = (size_t)(t + 1) * B + b;
      // prefetch the input projection while the MMAs run
      float4 pv[4][2];
      if (active) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float* src = pre + rin * 4 * H + (size_t)g * H + uo;
          pv[g][0] = *reinterpret_cast<const float4*>(src);
          pv[g][1] = *reinterpret_cast<const float4*>(src + 4);
        }
      }
      mbar_wait(tfull, (uint32_t)t & 1u);
      tc_fence_after();
      if (threadIdx.x == 64) { LP_STAMP(t, 5); LP_XSTAMP(t, 3, gtimer()); }
      float acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), acc);
      float gi[8], gf[8], go[8], gg[8], cn[8], hn[8];
      if (active) {
        const float* pi = reinterpret_cast<const float*>(&pv[0][0]);
        const float* pf = reinterpret_cast<const float*>(&pv[1][0]);
        const float* po = reinterpret_cast<const float*>(&pv[2][0]);
        const float* pg = reinterpret_cast<const float*>(&pv[3][0]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          gi[j] = fast_sigmoid(acc[j] + pi[j]);
          gf[j] = fast_sigmoid(acc[8 + j] + pf[j]);
          go[j] = fast_sigmoid(acc[16 + j] + po[j]);
          gg[j] = fast_tanh(acc[24 + j] + pg[j]);
          cn[j] = gf[j] * ccarry[j] + gi[j] * gg[j];
          hn[j] = go[j] * fast_tanh(cn[j]);
          ccarry[j] = cn[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { gi[j] = gf[j] = go[j] = gg[j] = cn[j] = hn[j] = 0.f; }
      }
      // (1) the only output the NEXT step depends on: h_t as bf16 planes, row (t+1)*B + b of [P][(T+1)B][H]
      if (rowok) {
        __nv_bfloat16 pl[3][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split3(hn[j], pl[0][j], pl[1][j], pl[2][j]);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          uint4 o;
          o.x = (uint32_t)__bfloat16_as_ushort(pl[p][0]) | ((uint32_t)__bfloat16_as_ushort(pl[p][1]) << 16);
          o.y = (uint32_t)__bfloat16_as_ushort(pl[p][2]) | ((uint32_t)__bfloat16_as_ushort(pl[p][3]) << 16);
          o.z = (uint32_t)__bfloat16_as_ushort(pl[p][4]) | ((uint32_t)__bfloat16_as_ushort(pl[p][5]) << 16);
          o.w = (uint32_t)__bfloat16_as_ushort(pl[p][6]) | ((uint32_t)__bfloat16_as_ushort(pl[p][7]) << 16);
          *reinterpret_cast<uint4*>(hp + (size_t)p * hp_plane + rout * H + uo) = o;
        }
      }
      // (2) publish it.  Every writer orders its generic-proxy stores before later async-proxy (TMA) reads; the CTA
      // barrier then makes one thread's gpu-scope release cumulative over all.
      if (threadIdx.x == 64) LP_STAMP(t, 6);
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, LP_EPI_THREADS);
      if (threadIdx.x == 64) {
        LP_STAMP(t, 2);      // (debug) reuse slot 2: all epilogue warps done
        group_arrive(counter + 32 * blockIdx.y);
        LP_STAMP(t, 7);
        LP_GSTAMP(t, 1);
      }
      // hold the other warps until the release is out: their stores / prefetch loads would otherwise sit in front of
      // the membar in the SM's memory pipeline (measured: +3.4k cycles per step)
      named_bar_sync(3, LP_EPI_THREADS);
      // ... and until the producer thread has seen the barrier open and issued the next step's first TMA loads: its
      // polling atomics share the SM's memory pipeline with the stores below (measured: each poll took ~2.5 us)
      if (t + 1 < T) mbar_wait(gobar, (uint32_t)(t + 1) & 1u);
      // (3) everything only the backward pass needs goes out OFF the critical path, overlapping the next step's TMA
      // loads and MMAs
      if (rowok) {
        float* gdst = pre + rin * 4 * H + uo;
#define ST8(ptr, a)                                                                          \
        *reinterpret_cast<float4*>(ptr) = make_float4(a[0], a[1], a[2], a[3]);               \
        *reinterpret_cast<float4*>((ptr) + 4) = make_float4(a[4], a[5], a[6], a[7]);
        ST8(gdst, gi) ST8(gdst + H, gf) ST8(gdst + 2 * H, go) ST8(gdst + 3 * H, gg)
        ST8(c + rout * H + uo, cn) ST8(h + rout * H + uo, hn)
        if (xdrop) {
          float xd[8];
          const uint64_t mi = (uint64_t)rin * H + uo;
          float4 ma = drop_at4(drop, mi), mb = drop_at4(drop, mi + 4);
          xd[0] = hn[0] * ma.x; xd[1] = hn[1] * ma.y; xd[2] = hn[2] * ma.z; xd[3] = hn[3] * ma.w;
          xd[4] = hn[4] * mb.x; xd[5] = hn[5] * mb.y; xd[6] = hn[6] * mb.z; xd[7] = hn[7] * mb.w;
          ST8(xdrop + rin * H + uo, xd)
        }
#undef ST8
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();           // no CTA leaves while a peer can still multicast into it / signal it
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------
// Backward (K9): all T steps of rnn_backward for one layer (misc/RNNUtils.lua:181-210; cell math SURVEY App. A).
//
//   grid = (H/64 column tiles of dh) x (4 K-splits = the four gate blocks of da) x (ceil(B/128) batch tiles)
//   per step t = T-1 .. 0, two phases separated by grid barriers:
//     A. element-wise, spread over all epilogue threads of the grid: da_t = f(dh_t, dc_t, gates_t, c_{t-1}, c_t),
//        written as bf16 planes (the TMA / wgrad / dgrad operand); dc carry and the per-row sum over t of da (bias
//        gradients) stay in registers
//     B. dh_{t-1} += da_t[:, gate block] . W_hh[gate block, :]   with the CTA's [512 x 64] slice of W_hh RESIDENT
//        in shared memory (MN-major B operand: no transposed weight copy); the four split-K partials go to a
//        double-buffered [4][B][H] scratch and are summed (fixed order, deterministic) by the next phase A.
template <int P, int CL>
__global__ void __launch_bounds__(LP_THREADS, 1)
lstm_bwd_persistent_kernel(const __grid_constant__ CUtensorMap mapDA, const __grid_constant__ CUtensorMap mapW,
                           const float* __restrict__ gates, const float* __restrict__ c, const float* __restrict__ dh0,
                           const float* __restrict__ dc0, int ld0, const float* __restrict__ dh_above, Drop drop,
                           float* __restrict__ dasum, __nv_bfloat16* __restrict__ dap, long long dap_plane,
                           float* __restrict__ dhbuf, float* __restrict__ dc_init, const int32_t* __restrict__ len, int T,
                           int B, int H, int KB, int S, unsigned int* counter, long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t W_TILE = 64 * 128;
  constexpr uint32_t A_PLANE = 128 * 128;
  const uint32_t w0 = base;
  const uint32_t a0 = w0 + (uint32_t)KB * P * W_TILE;
  const uint32_t bar0 = a0 + (uint32_t)S * P * A_PLANE;
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * LP_MAX_STAGES, wfull = bar0 + 16 * LP_MAX_STAGES,
                 tfull = wfull + 8, gobar = wfull + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar0 - raw) + 16 * LP_MAX_STAGES + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 64, ks = blockIdx.y, m0 = blockIdx.z * 128;
  const unsigned int G = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int k_base = ks * H;                 // this split's gate block inside the 4H contraction dimension
  const int tlast = dc_init ? 0 : 1;         // with dc_init: also produce d(initial state) (dh partials of step 0 + dc)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapDA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      // with a cluster of CL CTAs sharing the A tile, a stage is free only when all CL consumers have released it
      for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CL); }
      mbar_init(wfull, 1);
      mbar_init(tfull, 1);
      mbar_init(gobar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 64);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();           // peers' barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(wfull, (uint32_t)KB * P * W_TILE);
      for (int kb = 0; kb < KB; ++kb)
        for (int p = 0; p < P; ++p)
          tma_load_3d(w0 + (uint32_t)(kb * P + p) * W_TILE, &mapW, wfull, n0, k_base + kb * 64, p);
      int it = 0;
      const int gokb = (S < KB ? S : KB) - 1;
      for (int t = T - 1; t >= tlast; --t) {
        const unsigned int k = (unsigned int)(T - 1 - t);
        grid_wait(counter, (2 * k + 1) * G);             // da_t is complete everywhere
        fence_proxy_async();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          mbar_expect_tx(full0 + 8 * s, P * A_PLANE);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            if (CL == 1) {
              tma_load_3d(a0 + (uint32_t)(s * P + p) * A_PLANE, &mapDA, full0 + 8 * s, k_base + kb * 64, t * B + m0, p);
            } else {
              constexpr uint32_t SL = 128 / CL;
              tma_load_3d_mc(a0 + (uint32_t)(s * P + p) * A_PLANE + crank * SL * 128, &mapDA, full0 + 8 * s,
                             k_base + kb * 64, t * B + m0 + (int)(crank * SL), p, kMask);
            }
          }
          if (kb == gokb) mbar_arrive(gobar);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, true);      // B = W_hh slice, MN-major
      mbar_wait(wfull, 0);
      int it = 0;
      for (int t = T - 1; t >= tlast; --t) {
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t dA[P], dB[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
              dA[p] = make_kmajor_sw128_desc(a0 + (uint32_t)(s * P + p) * A_PLANE + k * 32);
              dB[p] = make_mnmajor_sw128_desc(w0 + (uint32_t)(kb * P + p) * W_TILE + k * 2048);
            }
            if (P >= 2) {
              umma_f16(tmem_base, dA[0], dB[1], idesc, acc); acc = 1;
              umma_f16(tmem_base, dA[1], dB[0], idesc, acc);
            }
            umma_f16(tmem_base, dA[0], dB[0], idesc, acc); acc = 1;
          }
          if (CL == 1) umma_commit(empty0 + 8 * s); else umma_commit_mc(empty0 + 8 * s, kMask);
        }
        umma_commit(tfull);
      }
    }
  } else {
    // ===== 8 element-wise / epilogue warps =====
    const int et = threadIdx.x - 64;                         // 0..255
    const int q = warp & 3, cc = (warp - 2) >> 2;
    const int brow = m0 + q * 32 + lane;
    const int H4 = H >> 2;
    const long long items = (long long)B * H4;
    const long long gthreads = (long long)G * LP_EPI_THREADS;
    // Every thread owns the same <= 2 items (batch row b, 4 hidden units) at every step, so the dc carry lives in
    // registers and the step-invariant operands (gates_t, c_{t-1}) of the NEXT step are prefetched from HBM while the
    // tensor cores run this step's phase B.
    constexpr int NI = 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    bool valid[NI];
    int bq[NI], ucol[NI], first_t[NI];
    float4 gi[NI], gf[NI], go[NI], gg[NI], cp[NI], cn[NI], dcr[NI];
    // sum over all steps of this thread's da elements: the bias gradients (both biases receive sum(da))
    // need only the column sums of da, so fp32 da never goes to HBM -- just these [B][4H] per-row partial sums at the end
    float4 bsi[NI], bsf[NI], bso[NI], bsg[NI];
#pragma unroll
    for (int n = 0; n < NI; ++n) {
      bsi[n] = bsf[n] = bso[n] = bsg[n] = z;
      const long long i = (long long)cta * LP_EPI_THREADS + et + (long long)n * gthreads;
      valid[n] = i < items;
      bq[n] = valid[n] ? (int)(i / H4) : 0;
      ucol[n] = valid[n] ? (int)(i % H4) * 4 : 0;
      first_t[n] = valid[n] ? (len ? T - len[bq[n]] : 0) : T;          // row active iff t >= first_t
      gi[n] = gf[n] = go[n] = gg[n] = cp[n] = cn[n] = dcr[n] = z;
      if (T - 1 >= first_t[n]) {
        const size_t row = (size_t)(T - 1) * B + bq[n];
        const float* g = gates + row * 4 * H + ucol[n];
        gi[n] = *reinterpret_cast<const float4*>(g);
        gf[n] = *reinterpret_cast<const float4*>(g + H);
        go[n] = *reinterpret_cast<const float4*>(g + 2 * H);
        gg[n] = *reinterpret_cast<const float4*>(g + 3 * H);
        cp[n] = *reinterpret_cast<const float4*>(c + row * H + ucol[n]);
        cn[n] = *reinterpret_cast<const float4*>(c + (row + B) * H + ucol[n]);
        dcr[n] = *reinterpret_cast<const float4*>(dc0 + (size_t)bq[n] * ld0 + ucol[n]);
      }
    }
    const size_t BHs = (size_t)B * H;
    for (int t = T - 1; t >= 0; --t) {
      const unsigned int k = (unsigned int)(T - 1 - t);
      if (et == 0) LP_STAMP(t, 0);
      if (t < T - 1) {                                       // dh_t (split-K sums of step t+1) complete everywhere
        if (et == 0) grid_wait(counter, (2 * k) * G);
        named_bar_sync(2, LP_EPI_THREADS);
      }
      if (et == 0) LP_STAMP(t, 1);
      // ---- phase A: cell backward, element-wise ----
      const float* dh_src = dhbuf + (size_t)((t + 1) & 1) * 4 * BHs;      // 4 split-K partials of dh_t
      float* dh_part = dhbuf + ((size_t)(t & 1) * 4 + ks) * BHs;          // this split's partial of dh_{t-1}
      float4 dai[NI], daf[NI], dao[NI], dag[NI];
      size_t row[NI];
#pragma unroll
      for (int n = 0; n < NI; ++n) {
        row[n] = (size_t)t * B + bq[n];
        dai[n] = daf[n] = dao[n] = dag[n] = z;
        if (!valid[n]) continue;
        if (t >= first_t[n]) {
          const size_t o = (size_t)bq[n] * H + ucol[n];
          float4 dh;
          if (t == T - 1) {
            dh = *reinterpret_cast<const float4*>(dh0 + (size_t)bq[n] * ld0 + ucol[n]);
          } else {
            const float* ds = dh_src + o;
            float4 d0 = *reinterpret_cast<const float4*>(ds), d1 = *reinterpret_cast<const float4*>(ds + BHs),
                   d2 = *reinterpret_cast<const float4*>(ds + 2 * BHs), d3 = *reinterpret_cast<const float4*>(ds + 3 * BHs);
            dh = make_float4((d0.x + d1.x) + (d2.x + d3.x), (d0.y + d1.y) + (d2.y + d3.y), (d0.z + d1.z) + (d2.z + d3.z),
                             (d0.w + d1.w) + (d2.w + d3.w));
          }
          if (dh_above) {
            float4 ua = *reinterpret_cast<const float4*>(dh_above + row[n] * H + ucol[n]);
            float4 mk = drop_at4(drop, (uint64_t)row[n] * H + ucol[n]);
            dh.x += ua.x * mk.x; dh.y += ua.y * mk.y; dh.z += ua.z * mk.z; dh.w += ua.w * mk.w;
          }
#define LB(kk)                                                                     \
          { float tc = fast_tanh(cn[n].kk);                                          \
            float dct = dcr[n].kk + dh.kk * go[n].kk * (1.0f - tc * tc);             \
            dao[n].kk = dh.kk * tc * go[n].kk * (1.0f - go[n].kk);                   \
            dai[n].kk = dct * gg[n].kk * gi[n].kk * (1.0f - gi[n].kk);               \
            daf[n].kk = dct * cp[n].kk * gf[n].kk * (1.0f - gf[n].kk);               \
            dag[n].kk = dct * gi[n].kk * (1.0f - gg[n].kk * gg[n].kk);               \
            dcr[n].kk = dct * gf[n].kk; }
          LB(x) LB(y) LB(z) LB(w)
#undef LB
#define ACC4(a, b) a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          ACC4(bsi[n], dai[n]) ACC4(bsf[n], daf[n]) ACC4(bso[n], dao[n]) ACC4(bsg[n], dag[n])
#undef ACC4
        }
        // (1) what phase B consumes through TMA: da_t as bf16 planes
        const float4 gsrc[4] = {dai[n], daf[n], dao[n], dag[n]};
#pragma unroll
        for (int gI = 0; gI < 4; ++gI) {
          __nv_bfloat16 pl[3][4];
          split3(gsrc[gI].x, pl[0][0], pl[1][0], pl[2][0]);
          split3(gsrc[gI].y, pl[0][1], pl[1][1], pl[2][1]);
          split3(gsrc[gI].z, pl[0][2], pl[1][2], pl[2][2]);
          split3(gsrc[gI].w, pl[0][3], pl[1][3], pl[2][3]);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            uint2 ov;
            ov.x = (uint32_t)__bfloat16_as_ushort(pl[p][0]) | ((uint32_t)__bfloat16_as_ushort(pl[p][1]) << 16);
            ov.y = (uint32_t)__bfloat16_as_ushort(pl[p][2]) | ((uint32_t)__bfloat16_as_ushort(pl[p][3]) << 16);
            *reinterpret_cast<uint2*>(dap + (size_t)p * dap_plane + row[n] * 4 * H + (size_t)gI * H + ucol[n]) = ov;
          }
        }
      }
      // (2) publish da_t (barrier 2k+1)
      if (et == 0) LP_STAMP(t, 2);
      fence_proxy_async();
      named_bar_sync(1, LP_EPI_THREADS);
      if (et == 0) { LP_STAMP(t, 3); grid_arrive(counter); LP_STAMP(t, 4); }
      named_bar_sync(3, LP_EPI_THREADS);          // keep the SM's memory pipeline clear until the release is out
      if (t >= tlast) mbar_wait(gobar, k & 1u);      // ... and until the producer has seen da_t complete and issued its loads
      // (3) off the critical path: the prefetch of step t-1's gates / cell states
#pragma unroll
      for (int n = 0; n < NI; ++n) {
        if (!valid[n]) continue;
        if (t == 0 && dc_init) *reinterpret_cast<float4*>(dc_init + (size_t)bq[n] * H + ucol[n]) = dcr[n];
        if (t > 0) {
          cn[n] = cp[n];                                               // c_{t-1} becomes the "new" cell of step t-1
          if (t - 1 >= first_t[n]) {
            const size_t rp = row[n] - B;
            const float* g = gates + rp * 4 * H + ucol[n];
            gi[n] = *reinterpret_cast<const float4*>(g);
            gf[n] = *reinterpret_cast<const float4*>(g + H);
            go[n] = *reinterpret_cast<const float4*>(g + 2 * H);
            gg[n] = *reinterpret_cast<const float4*>(g + 3 * H);
            cp[n] = *reinterpret_cast<const float4*>(c + rp * H + ucol[n]);
          }
        }
      }
      if (t < tlast) break;
      // ---- phase B epilogue: split-K partial of dh_{t-1} ----
      mbar_wait(tfull, k & 1u);
      tc_fence_after();
      if (et == 0) LP_STAMP(t, 5);
      float acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), acc);
      if (brow < B) {                                          // one full 128 B line per thread, deterministic
        float* dst = dh_part + (size_t)brow * H + n0 + cc * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
      tc_fence_before();
      named_bar_sync(1, LP_EPI_THREADS);
      if (et == 0) { LP_STAMP(t, 6); grid_arrive(counter); LP_STAMP(t, 7); }      // barrier 2k+2
    }
#pragma unroll
    for (int n = 0; n < NI; ++n) {
      if (!valid[n]) continue;
      float* dr = dasum + (size_t)bq[n] * 4 * H + ucol[n];
      *reinterpret_cast<float4*>(dr) = bsi[n];
      *reinterpret_cast<float4*>(dr + H) = bsf[n];
      *reinterpret_cast<float4*>(dr + 2 * H) = bso[n];
      *reinterpret_cast<float4*>(dr + 3 * H) = bsg[n];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();           // no CTA leaves while a peer can still multicast into it / signal it
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// cooperative launch (all CTAs co-resident: the kernels spin on grid barriers), optionally with thread-block clusters
static cudaError_t launch_coop(const void* fn, dim3 grid, int cluster_x, void** args, size_t smem, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(LP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int n = 0;
  attrs[n].id = cudaLaunchAttributeCooperative; attrs[n].val.cooperative = 1; ++n;
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster_x; attrs[n].val.clusterDim.y = 1; attrs[n].val.clusterDim.z = 1; ++n;
  }
  cfg.attrs = attrs; cfg.numAttrs = n;
  return cudaLaunchKernelExC(&cfg, fn, args);
}
static int g_cluster = -1;                    // A-tile multicast cluster size (NVQA_LSTM_CLUSTER=1 disables)
static int cluster_pref() {
  if (g_cluster < 0) {
    const char* e = getenv("NVQA_LSTM_CLUSTER");
    g_cluster = e ? atoi(e) : 1;   // measured on B200: multicast does not shorten the load phase (SM ingest-bound)
    if (g_cluster != 1 && g_cluster != 2 && g_cluster != 4 && g_cluster != 8) g_cluster = 1;
  }
  return g_cluster;
}

static int persistent_limits(int* num_sms, int* max_smem) {
  static int sms = 0, smem = 0;
  if (!sms) {
    int dev = 0;
    NVQA_CUDA(cudaGetDevice(&dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  *num_sms = sms; *max_smem = smem;
  return 0;
}

// dh_init[b][u] = sum of the four split-K partials of d h_{-1} left in dhbuf slot 0 by step 0 (fixed order)
__global__ void __launch_bounds__(256) sum4_kernel(const float* __restrict__ part, float* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4* p = reinterpret_cast<const float4*>(part);
  float4 a = p[i], b = p[i + n4], c = p[i + 2 * n4], d = p[i + 3 * n4];
  reinterpret_cast<float4*>(out)[i] = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z),
                                                  (a.w + b.w) + (c.w + d.w));
}

int lstm_bwd_persistent(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, const float* gates, const float* c,
                        const float* dh0, const float* dc0, int ld0, const float* dh_above, Drop d, float* dasum,
                        __nv_bfloat16* dap, long long dap_plane_rows, float* dhbuf, float* dh_init, float* dc_init,
                        const int32_t* len, int T, int B, int H, unsigned int* counter) {
  if ((dh_init == nullptr) != (dc_init == nullptr)) return -1;
  if (P < 1 || P > 2) return -1;
  if (H % 64 != 0 || H < 64) return -1;
  int num_sms = 0, max_smem = 0;
  NVQA_TRY(persistent_limits(&num_sms, &max_smem));
  const int KB = H / 64;
  dim3 grid(H / 64, 4, ceil_div(B, 128));
  if ((int)(grid.x * grid.y * grid.z) > num_sms) return -1;
  const size_t wbytes = (size_t)KB * P * 8192, misc = 1024 + 256;
  if ((size_t)max_smem < misc + wbytes + 2 * (size_t)P * 16384) return -1;
  int S = (int)(((size_t)max_smem - misc - wbytes) / ((size_t)P * 16384));
  if (S > LP_MAX_STAGES) S = LP_MAX_STAGES;
  const size_t smem = wbytes + (size_t)S * P * 16384 + misc;

  __nv_bfloat16* wp = nullptr;
  int pitch = 0;
  NVQA_TRY(prepare_planes(ws, s, P, Wh, 4 * H, H, H, true, &wp, &pitch));
  CUtensorMap mapW, mapDA;
  NVQA_TRY(get_map(ws, wp, 4 * H, pitch, P, 64, &mapW));             // MN-major B: 64 k-rows x 64 columns
  int CL = cluster_pref();
  if ((int)grid.x % CL != 0) CL = 1;
  if (dap_plane_rows <= 0) dap_plane_rows = (long long)T * B;
  NVQA_TRY(get_map(ws, dap, T * B, 4 * H, P, 128 / CL, &mapDA, dap_plane_rows * 4 * H));
  NVQA_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
  long long dap_plane = dap_plane_rows * 4 * H;
  int KBv = KB, Sv = S;
  long long* dbg = nullptr;
  static const bool want_dbg = getenv("NVQA_LSTM_DEBUG") != nullptr;
  if (want_dbg) {
    NVQA_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbg), (size_t)T * 8 * sizeof(long long)));
    NVQA_CUDA(cudaMemsetAsync(dbg, 0, (size_t)T * 8 * sizeof(long long), s));
  }
  void* args[] = {&mapDA, &mapW, &gates, &c, &dh0, &dc0, &ld0, &dh_above, &d, &dasum, &dap, &dap_plane, &dhbuf, &dc_init,
                  &len, &T, &B, &H, &KBv, &Sv, &counter, &dbg};
  const void* fn = nullptr;
#define LP_PICK(P_, CL_) if (P == P_ && CL == CL_) fn = (const void*)lstm_bwd_persistent_kernel<P_, CL_>
  LP_PICK(1, 1); LP_PICK(1, 2); LP_PICK(1, 4); LP_PICK(1, 8); LP_PICK(2, 1); LP_PICK(2, 2); LP_PICK(2, 4); LP_PICK(2, 8);
#undef LP_PICK
  NVQA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  if (CL > 4) NVQA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaError_t le = launch_coop(fn, grid, CL, args, smem, s);
  if (le != cudaSuccess && CL > 1) {          // clusters could not be made co-resident: fall back to unicast loads
    (void)cudaGetLastError();
    g_cluster = 1;
    return lstm_bwd_persistent(s, ws, P, Wh, gates, c, dh0, dc0, ld0, dh_above, d, dasum, dap, dap_plane_rows, dhbuf, dh_init,
                               dc_init, len, T, B, H, counter);
  }
  NVQA_CUDA(le);
  ++g_launches;
  if (dh_init) {
    const long long n4 = (long long)B * H / 4;
    sum4_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(dhbuf, dh_init, n4);
    NVQA_LAUNCHED();
  }
  if (want_dbg) {
    std::vector<long long> hbuf((size_t)T * 8);
    NVQA_CUDA(cudaStreamSynchronize(s));
    NVQA_CUDA(cudaMemcpy(hbuf.data(), dbg, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    fprintf(stderr, "lstm_bwd_persistent timeline (cycles): t | bar2_passed phaseA_done A_alldone A_arrived mma_done B_alldone B_arrived | step\n");
    for (int t = T - 2; t >= 1; --t) {
      const long long* e = &hbuf[(size_t)t * 8];
      fprintf(stderr, "%2d | %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld\n", t, e[1] - e[0], e[2] - e[0], e[3] - e[0],
              e[4] - e[0], e[5] - e[0], e[6] - e[0], e[7] - e[0], e[0] - hbuf[(size_t)(t + 1) * 8]);
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
int lstm_fwd_persistent(cudaStream_t s, UmmaWorkspace* ws, int P, const float* Wh, float* pre, float* c, float* h,
                        __nv_bfloat16* hp, long long hp_plane_rows, float* xdrop_next, const int32_t* len, Drop d, int T,
                        int B, int H, unsigned int* counter) {
  if (P < 1 || P > 2) return -1;                 // P = 3: W_hh does not fit in shared memory next to the A ring
  if (H % 64 != 0 || H < 64) return -1;
  static int num_sms = 0, max_smem = 0;
  if (!num_sms) {
    int dev = 0;
    NVQA_CUDA(cudaGetDevice(&dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  const int KB = H / 64;
  dim3 grid(H / 16, ceil_div(B, 128));
  if ((int)(grid.x * grid.y) > num_sms) return -1;
  const size_t wbytes = (size_t)KB * P * 8192, misc = 1024 + 256;
  int S = (int)(((size_t)max_smem - misc - wbytes) / ((size_t)P * 16384));
  if ((size_t)max_smem < misc + wbytes + 2 * (size_t)P * 16384) return -1;
  if (S > LP_MAX_STAGES) S = LP_MAX_STAGES;
  const size_t smem = wbytes + (size_t)S * P * 16384 + misc;

  __nv_bfloat16* wp = nullptr;
  int pitch = 0;
  NVQA_TRY(prepare_planes(ws, s, P, Wh, 4 * H, H, H, true, &wp, &pitch));
  CUtensorMap mapW, mapH;
  NVQA_TRY(get_map(ws, wp, 4 * H, pitch, P, 8, &mapW));
  int CL = cluster_pref();
  if ((int)grid.x % CL != 0) CL = 1;
  if (hp_plane_rows <= 0) hp_plane_rows = (long long)(T + 1) * B;
  NVQA_TRY(get_map(ws, hp, (T + 1) * B, H, P, 128 / CL, &mapH, hp_plane_rows * H));
  NVQA_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int) * 32 * grid.y, s));      // one counter line per batch tile
  long long hp_plane = hp_plane_rows * H;
  int KBv = KB, Sv = S;
  long long* dbg = nullptr;
  static const bool want_dbg = getenv("NVQA_LSTM_DEBUG") != nullptr;
  if (getenv("NVQA_POLL_NS")) {
    int ns = atoi(getenv("NVQA_POLL_NS"));
    NVQA_CUDA(cudaMemcpyToSymbol(g_poll_ns, &ns, sizeof(int)));
  }
  const size_t dbg_n = (size_t)T * 8 + 1024 + 1024;
  if (want_dbg) {
    NVQA_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbg), dbg_n * sizeof(long long)));
    NVQA_CUDA(cudaMemsetAsync(dbg, 0, dbg_n * sizeof(long long), s));
  }
  void* args[] = {&mapH, &mapW, &pre, &c, &h, &hp, &hp_plane, &xdrop_next, &len, &d, &T, &B, &H, &KBv, &Sv, &counter, &dbg};
  const void* fn = nullptr;
#define LP_PICK(P_, CL_) if (P == P_ && CL == CL_) fn = (const void*)lstm_fwd_persistent_kernel<P_, CL_>
  LP_PICK(1, 1); LP_PICK(1, 2); LP_PICK(1, 4); LP_PICK(1, 8); LP_PICK(2, 1); LP_PICK(2, 2); LP_PICK(2, 4); LP_PICK(2, 8);
#undef LP_PICK
  NVQA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  if (CL > 4) NVQA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaError_t le = launch_coop(fn, grid, CL, args, smem, s);
  if (le != cudaSuccess && CL > 1) {          // clusters could not be made co-resident: fall back to unicast loads
    (void)cudaGetLastError();
    g_cluster = 1;
    return lstm_fwd_persistent(s, ws, P, Wh, pre, c, h, hp, hp_plane_rows, xdrop_next, len, d, T, B, H, counter);
  }
  NVQA_CUDA(le);
  ++g_launches;
  if (want_dbg) {   // per-step timeline of CTA (0,0) in SM cycles, relative to the step's first stamp
    std::vector<long long> hbuf(dbg_n);
    NVQA_CUDA(cudaStreamSynchronize(s));
    NVQA_CUDA(cudaMemcpy(hbuf.data(), dbg, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    fprintf(stderr, "lstm_fwd_persistent timeline (cycles after the grid barrier opened): t | first_data last_data mma_done epi_start epi_stored epi_alldone arrived | step\n");
    for (int t = 1; t < T; ++t) {
      const long long* e = &hbuf[(size_t)t * 8];
      fprintf(stderr, "%2d | %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld\n", t, e[3] - e[1], e[0] - e[1], e[4] - e[1],
              e[5] - e[1], e[6] - e[1], e[2] - e[1], e[7] - e[1], e[1] - hbuf[(size_t)(t - 1) * 8 + 1]);
    }
    {   // wall-clock (globaltimer, ns) spread over all CTAs at step 10 -> 11
      const int G = grid.x * grid.y;
      const long long* g = &hbuf[(size_t)T * 8];
      long long open10_min = 1LL << 62, open10_max = 0, arr10_min = 1LL << 62, arr10_max = 0, open11_min = 1LL << 62, open11_max = 0;
      int slowest = 0;
      for (int i = 0; i < G; ++i) {
        open10_min = std::min(open10_min, g[i * 4]); open10_max = std::max(open10_max, g[i * 4]);
        arr10_min = std::min(arr10_min, g[i * 4 + 1]);
        if (g[i * 4 + 1] > arr10_max) { arr10_max = g[i * 4 + 1]; slowest = i; }
        open11_min = std::min(open11_min, g[i * 4 + 2]); open11_max = std::max(open11_max, g[i * 4 + 2]);
      }
      fprintf(stderr, "step 10 wall clock (ns, relative to first CTA seeing the barrier open): open %lld..%lld | arrive %lld..%lld "
              "(slowest CTA %d = n-slice %d, m-tile %d) | next open %lld..%lld\n", 0LL, open10_max - open10_min,
              arr10_min - open10_min, arr10_max - open10_min, slowest, slowest % (int)grid.x, slowest / (int)grid.x,
              open11_min - open10_min, open11_max - open10_min);
      for (int y = 0; y < (int)grid.y; y += 3) {
        std::vector<long long> o10, a10, o11;
        long long base = 1LL << 62;
        for (int x = 0; x < (int)grid.x; ++x) base = std::min(base, g[(y * grid.x + x) * 4]);
        for (int x = 0; x < (int)grid.x; ++x) {
          o10.push_back(g[(y * grid.x + x) * 4] - base); a10.push_back(g[(y * grid.x + x) * 4 + 1] - base);
          o11.push_back(g[(y * grid.x + x) * 4 + 2] - base);
        }
        fprintf(stderr, "m-tile %d per n-slice: smid | open10 first_data last_data tfull arrive10 | open11 (ns)\n", y);
        const long long* x3 = &hbuf[(size_t)T * 8 + 1024];
        for (int x = 0; x < (int)grid.x; ++x) {
          const long long* e = &x3[(y * grid.x + x) * 4];
          fprintf(stderr, "  n%02d sm%3lld | %5lld %5lld %5lld %5lld %5lld | %5lld\n", x, e[2], o10[x], e[0] - base, e[1] - base,
                  e[3] - base, a10[x], o11[x]);
        }
      }
      fprintf(stderr, "arrive(ns) by m-tile:");
      for (int y = 0; y < (int)grid.y; ++y) {
        long long mn = 1LL << 62, mx = 0;
        for (int x = 0; x < (int)grid.x; ++x) { long long v = g[(y * grid.x + x) * 4 + 1] - open10_min; mn = std::min(mn, v); mx = std::max(mx, v); }
        fprintf(stderr, "  m%d: %lld..%lld", y, mn, mx);
      }
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace nvqa
