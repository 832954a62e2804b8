// tcgen05 / TMEM / TMA GEMM engine for sm_100a (hand-written PTX, no CUTLASS, no cuBLAS).
//
//   C[M x N] (fp32) = (beta ? C : 0) + A . B^T + bias0 + bias1        A: [M x K], B: [N x K]
//
// Operands are bf16 "planes": an fp32 value x is split as x = p0 + p1 + p2 (+ O(2^-24 |x|)) with
// p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1).  With P planes per operand the kernel issues
//   P = 1:  (0,0)                                   -> plain bf16-operand GEMM (1e-2 mode)
//   P = 2:  (0,0) (0,1) (1,0)                       -> ~2^-16 relative per product
//   P = 3:  (0,0) (0,1) (1,0) (1,1) (0,2) (2,0)     -> fp32-equivalent (dropped terms <= 2^-25)
// tcgen05.mma kind::f16 instructions per K step, all accumulating in fp32 in the same TMEM tile.
// This is the "fp32 parity on tensor cores" mode of the arch1 step: the reference's nn.Linear GEMMs
// (misc/LSTM.lua:41-42, misc/netdef.lua:10-11, 002_train_baseline.lua:154) are fp32 SGEMMs.
//
// Kernel anatomy (PERSISTENT: one CTA per SM walks the 128 x BN output tiles tile = blockIdx.x, + gridDim.x, ...; 192 threads):
//   warp 0   : TMA producer   - cp.async.bulk.tensor.3d (SWIZZLE_128B boxes of 64 bf16 x rows x 1 plane), one ring across tiles
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2-5: epilogue       - tcgen05.ld (32 lanes x 32 columns per warp), bias / beta, fp32 stores
// smem: STAGES x { P A-planes [128 x 64] , P B-planes [BN x 64] } ring with full/empty mbarriers.
// TMEM: TWO accumulator stages of BN columns (tfull / tempty mbarriers): the epilogue of tile i (128 KB of stores at
// BN = 256) overlaps the main loop of tile i + 1 -- for the short-K products of the step (input projections K = 200 / 512,
// vocabulary projection K = 512) the epilogue was a third of a tile's time.
#include <unordered_map>
#include <vector>

#include <algorithm>

#include "umma_engine.cuh"
#include "umma_ptx.cuh"

namespace nvqa {

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
constexpr int UG_BM = 128;
constexpr int UG_BK = 64;                  // bf16 elements per 128-byte swizzle row
constexpr int UG_THREADS = 192;

template <int BN, int P>
struct UgCfg {
  static constexpr int A_PLANE = UG_BM * UG_BK * 2;           // 16 KB
  static constexpr int B_PLANE = BN * UG_BK * 2;
  static constexpr int STAGE = P * (A_PLANE + B_PLANE);
  static constexpr int EPI_PITCH = 36;                        // floats per row of an epilogue staging chunk (32 + 4: conflict-free)
  static constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;    // one [32 rows x 32 columns] chunk per epilogue warp
  static constexpr int MAX_STAGES = (227 * 1024 - 4096 - EPI_BYTES) / STAGE;
  static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * 1024 /*two bias tiles*/ + EPI_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;          // two accumulator stages (a power of two >= 32)
  static_assert(STAGES >= 2, "pipeline needs two stages");
};

template <int BN, int P, bool AMN, bool BMN>
__global__ void __launch_bounds__(UG_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int N,
                 int K, float* __restrict__ C, int ldc, int beta, const float* __restrict__ bias0,
                 const float* __restrict__ bias1, int a_row0, int b_row0, int kb_per_split, long long c_split_stride,
                 int tiles_n, int tiles_m, int splits) {
  using Cfg = UgCfg<BN, P>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + Cfg::STAGES * Cfg::STAGE);
  const uint32_t full0 = base + Cfg::STAGES * Cfg::STAGE;            // full[s]  at full0 + 8 s
  const uint32_t empty0 = full0 + 8 * Cfg::STAGES;                   // empty[s]
  const uint32_t tfull = empty0 + 8 * Cfg::STAGES;                   // accumulator stage a ready   (tfull + 8 a)
  const uint32_t tempty = tfull + 16;                                 // accumulator stage a drained (tempty + 8 a)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);
  float* bias_all = reinterpret_cast<float*>(gen + Cfg::STAGES * Cfg::STAGE + 256);    // bias0 + bias1 of the tile, per stage
  float* epi_all = bias_all + 512;                                                      // per-warp staging chunks of the epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_all = (K + UG_BK - 1) / UG_BK;
  const int total_tiles = tiles_n * tiles_m * splits;                 // tile = (z * tiles_m + y) * tiles_n + x

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull + 8 * a, 1); mbar_init(tempty + 8 * a, 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();                      // everything above touched only this CTA's shared / tensor memory (common.cuh)

  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop, one elected lane issues (see elect_one_sync in umma_ptx.cuh) =====
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tz = tile / (tiles_n * tiles_m), ty = (tile / tiles_n) % tiles_m, tx = tile % tiles_n;
    const int m0 = ty * UG_BM, n0 = tx * BN;
    const int kb0 = tz * kb_per_split;                                // split-K: this tile's k-block range
    const int nkb = min(nkb_all, kb0 + kb_per_split) - kb0;
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      if (elect_one_sync()) {
        mbar_expect_tx(full0 + 8 * s, Cfg::STAGE);
        const uint32_t sa = base + s * Cfg::STAGE;
        const uint32_t sb = sa + P * Cfg::A_PLANE;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (!AMN) {
            tma_load_3d(sa + p * Cfg::A_PLANE, &mapA, full0 + 8 * s, (kb0 + kb) * UG_BK, a_row0 + m0, p);
          } else {
#pragma unroll
            for (int c = 0; c < UG_BM / 64; ++c)
              tma_load_3d(sa + p * Cfg::A_PLANE + c * 8192, &mapA, full0 + 8 * s, m0 + 64 * c,
                          a_row0 + (kb0 + kb) * UG_BK, p);
          }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (!BMN) {
            tma_load_3d(sb + p * Cfg::B_PLANE, &mapB, full0 + 8 * s, (kb0 + kb) * UG_BK, b_row0 + n0, p);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_3d(sb + p * Cfg::B_PLANE + c * 8192, &mapB, full0 + 8 * s, n0 + 64 * c,
                          b_row0 + (kb0 + kb) * UG_BK, p);
          }
        }
      }
      __syncwarp();
    }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (uniform control flow and descriptors), one elected lane issues;
    // under `if (lane == 0)` every UTCHMMA was wrapped in an ELECT / R2UR / BRA.U.ANY loop (~80 cycles per MMA) =====
    constexpr uint32_t idesc = make_idesc_bf16(UG_BM, BN, AMN, BMN);
    int it = 0, li = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
    const int tz = tile / (tiles_n * tiles_m);
    const int kb0 = tz * kb_per_split;
    const int nkb = min(nkb_all, kb0 + kb_per_split) - kb0;
    const int as = li & 1;                                            // accumulator stage of this tile
    const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
    mbar_wait(tempty + 8 * as, (((uint32_t)li >> 1) & 1u) ^ 1u);      // the epilogue has drained the tile two tiles back
    tc_fence_after();
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        uint32_t acc = kb > 0 ? 1u : 0u;
        const uint32_t sa = base + s * Cfg::STAGE;
        const uint32_t sb = sa + P * Cfg::A_PLANE;
#pragma unroll
        for (int k = 0; k < UG_BK / 16; ++k) {
          // K advance per UMMA_K = 16 bf16: +32 bytes inside the 128 B swizzle row (K-major), or 16 k-rows of
          // 128 B = +2048 bytes (MN-major)
          uint64_t da[P], db[P];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            da[p] = AMN ? make_mnmajor_sw128_desc(sa + p * Cfg::A_PLANE + k * 2048)
                        : make_kmajor_sw128_desc(sa + p * Cfg::A_PLANE + k * 32);
            db[p] = BMN ? make_mnmajor_sw128_desc(sb + p * Cfg::B_PLANE + k * 2048)
                        : make_kmajor_sw128_desc(sb + p * Cfg::B_PLANE + k * 32);
          }
          // smallest terms first
          if (P == 3) {
            umma_f16(tacc, da[0], db[2], idesc, acc); acc = 1;
            umma_f16(tacc, da[2], db[0], idesc, acc);
            umma_f16(tacc, da[1], db[1], idesc, acc);
          }
          if (P >= 2) {
            umma_f16(tacc, da[0], db[1], idesc, acc); acc = 1;
            umma_f16(tacc, da[1], db[0], idesc, acc);
          }
          umma_f16(tacc, da[0], db[0], idesc, acc); acc = 1;
        }
        umma_commit(empty0 + 8 * s);          // smem slot free once these MMAs have read it
        if (kb == nkb - 1) umma_commit(tfull + 8 * as);  // accumulator complete
      }
      __syncwarp();
    }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    float* const C0 = C;
    int li = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
    const int tz = tile / (tiles_n * tiles_m), ty = (tile / tiles_n) % tiles_m, tx = tile % tiles_n;
    const int m0 = ty * UG_BM, n0 = tx * BN;
    const int as = li & 1;
    C = C0 + (size_t)tz * c_split_stride;
    float* bias_s = bias_all + as * 256;
    // stage the tile's bias row (read below as a shared-memory broadcast); the named barrier also orders the reuse of
    // the stage: nobody passes it before all four warps have finished the tile before
    for (int j = threadIdx.x - 64; j < BN; j += 128) {
      float bsum = 0.f;
      if (n0 + j < N) {
        if (bias0) bsum += __ldg(bias0 + n0 + j);
        if (bias1) bsum += __ldg(bias1 + n0 + j);
      }
      bias_s[j] = bsum;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tfull + 8 * as, ((uint32_t)li >> 1) & 1u);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
    // tcgen05.ld hands a lane one ROW x 32 columns; stored like that, a warp instruction would touch 32 different cache
    // lines with 16 bytes each (measured: the short-K GEMMs of the step were bound by exactly these stores, 5 B/clk/SM).
    // Each 32 x 32 chunk therefore goes through a per-warp shared-memory tile and leaves as 4 rows x 128 contiguous bytes
    // per instruction: full lines.
    float* stg = epi_all + (warp - 2) * 32 * Cfg::EPI_PITCH;
    const int sr = lane >> 3, sc = (lane & 7) * 4;              // store phase: lane -> (row within a group of 4, 4 columns)
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      if (n0 + cc * 32 >= N) break;
      float v[32];
      tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stg + lane * Cfg::EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int col = n0 + cc * 32 + sc;
      const float4 bb = *reinterpret_cast<const float4*>(bias_s + cc * 32 + sc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sr, row = m0 + q * 32 + r;
        if (row >= M || col >= N) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + r * Cfg::EPI_PITCH + sc);
        float o4[4] = {x.x + bb.x, x.y + bb.y, x.z + bb.z, x.w + bb.w};
        float* cp = C + (size_t)row * ldc + col;
        if (vec && col + 3 < N) {
          if (beta) {
            const float4 o = *reinterpret_cast<const float4*>(cp);
            o4[0] += o.x; o4[1] += o.y; o4[2] += o.z; o4[3] += o.w;
          }
          *reinterpret_cast<float4*>(cp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (col + t < N) cp[t] = (beta ? cp[t] : 0.f) + o4[t];
        }
      }
      __syncwarp();                              // the chunk is overwritten by the next one
    }
    tc_fence_before();                          // this warp's tcgen05.ld of the stage are complete (wait::ld inside tmem_ld32)
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty + 8 * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2): two CTAs on adjacent SMs compute one 256 x BN output tile.  Every CTA stages its own
// 128 A rows and only HALF of the B tile; one tcgen05.mma.cta_group::2 issued by the pair's leader multiplies both.
// Why: the single-CTA kernel streams P x (16 KB + BN x 128 B) per k-block through each SM -- 96 KB at BN = 256, P = 2 --
// and the step's big GEMMs were measured at 8.9-9.2 TB/s of L2 -> SM traffic, i.e. at the chip's L2 throughput ceiling
// (~6300 B/clk, B300_MICROARCH), with the tensor pipe 57 % active.  A pair fetches the B tile once for 256 rows: a third
// less traffic per FLOP, and the 64 KB stages fit three deep instead of two.
//   barriers   full[s]   leader only; the leader arms expect_tx for BOTH CTAs' bytes, the peer's TMA loads signal it remotely
//              empty[s]  every CTA its own, arrived by the leader's multicast tcgen05.commit
//              tfull[a]  every CTA its own (multicast commit); tempty[a] leader only, 8 arrivals = 4 epilogue warps x 2 CTAs
template <int BN, int P>
struct UgCfg2 {
  static constexpr int A_PLANE = UG_BM * UG_BK * 2;           // this CTA's 128 rows: 16 KB
  static constexpr int B_PLANE = (BN / 2) * UG_BK * 2;        // this CTA's half of the B tile
  static constexpr int STAGE = P * (A_PLANE + B_PLANE);
  static constexpr int EPI_PITCH = 36;                        // padded tile of the st.global fallback (fits the two boxes)
  static constexpr int EPI_WARP = 2 * 4096;                   // per epilogue warp: two 32 x 32 fp32 boxes (SWIZZLE_128B) in flight
  static constexpr int EPI_BYTES = 4 * EPI_WARP;
  static constexpr int TAIL = 256 /*barriers*/ + 2 * 1024 /*two bias tiles*/;
  static constexpr int MAX_STAGES = (227 * 1024 - TAIL - EPI_BYTES) / STAGE;
  static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
  // no alignment slack: the dynamic shared memory is declared __align__(1024) (and the kernel traps if it is not)
  static constexpr int SMEM = STAGES * STAGE + EPI_BYTES + TAIL;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(STAGES >= 2, "pipeline needs two stages");
  static_assert(STAGE % 1024 == 0 && SMEM <= 227 * 1024, "shared-memory layout");
  static_assert(32 * EPI_PITCH * 4 <= EPI_WARP, "fallback tile must fit the warp's staging area");
};

template <int BN, int P, bool AMN, bool BMN>
__global__ void __launch_bounds__(UG_THREADS, 1)
umma_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapC, int M, int N,
                      int K, float* __restrict__ C, int ldc, int beta, const float* __restrict__ bias0,
                      const float* __restrict__ bias1, int a_row0, int b_row0, int kb_per_split, long long c_split_stride,
                      int tiles_n, int tiles_m2, int splits) {
  using Cfg = UgCfg2<BN, P>;
  extern __shared__ uint8_t smem_pair_raw[];
  const uint32_t base = smem_u32(smem_pair_raw);
  if (base & 1023u) __trap();                                         // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen = smem_pair_raw;
  // [operand stages][epilogue staging 4 warps x 2 boxes][barriers 256 B][two bias tiles]
  constexpr int BAR_OFF = Cfg::STAGES * Cfg::STAGE + Cfg::EPI_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + BAR_OFF);
  const uint32_t full0 = base + BAR_OFF;
  const uint32_t empty0 = full0 + 8 * Cfg::STAGES;
  const uint32_t tfull = empty0 + 8 * Cfg::STAGES;
  const uint32_t tempty = tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);
  float* bias_all = reinterpret_cast<float*>(gen + BAR_OFF + 256);
  uint8_t* epi_all = gen + Cfg::STAGES * Cfg::STAGE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                            // 0 = the pair's leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb_all = (K + UG_BK - 1) / UG_BK;
  const int total_tiles = tiles_n * tiles_m2 * splits;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + 8 * a, 1); mbar_init(tempty + 8 * a, 8); }
    fence_barrier_init();
  }
  cluster_sync_all();                                                 // both CTAs' mbarriers exist; both are ready to allocate
  if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();                      // everything above touched only the pair's shared / tensor memory (common.cuh)

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of the B tile, bytes counted on the LEADER's full barrier =====
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
    const int tz = tile / (tiles_n * tiles_m2), ty = (tile / tiles_n) % tiles_m2, tx = tile % tiles_n;
    // the last column tile may be narrow: the pair multiplies N = neff columns only (next multiple of 16 of what is left),
    // of which each CTA supplies neff / 2 -- so the peer's half starts neff / 2 columns in, not BN / 2
    const int nrem = N - tx * BN;
    const int neff = nrem >= BN ? BN : ((nrem + 15) & ~15);
    const int m0 = ty * 2 * UG_BM + (int)rank * UG_BM, n0 = tx * BN + (int)rank * (neff / 2);
    const int kb0 = tz * kb_per_split;
    const int nkb = min(nkb_all, kb0 + kb_per_split) - kb0;
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      if (elect_one_sync()) {
        if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * Cfg::STAGE);
        const uint32_t fb = mapa_u32(full0 + 8 * s, 0);
        const uint32_t sa = base + s * Cfg::STAGE;
        const uint32_t sb = sa + P * Cfg::A_PLANE;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (!AMN) {
            tma_load_3d_pair(sa + p * Cfg::A_PLANE, &mapA, fb, (kb0 + kb) * UG_BK, a_row0 + m0, p);
          } else {
#pragma unroll
            for (int c = 0; c < UG_BM / 64; ++c)
              tma_load_3d_pair(sa + p * Cfg::A_PLANE + c * 8192, &mapA, fb, m0 + 64 * c, a_row0 + (kb0 + kb) * UG_BK, p);
          }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (!BMN) {
            tma_load_3d_pair(sb + p * Cfg::B_PLANE, &mapB, fb, (kb0 + kb) * UG_BK, b_row0 + n0, p);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c)
              tma_load_3d_pair(sb + p * Cfg::B_PLANE + c * 8192, &mapB, fb, n0 + 64 * c, b_row0 + (kb0 + kb) * UG_BK, p);
          }
        }
      }
      __syncwarp();
    }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader's warp issues for the pair =====
    int it = 0, li = 0;
    for (int tile = pair; tile < (rank == 0 ? total_tiles : 0); tile += npairs, ++li) {
    const int tz = tile / (tiles_n * tiles_m2);
    const int nrem = N - (tile % tiles_n) * BN;
    const uint32_t idesc = make_idesc_bf16(2 * UG_BM, nrem >= BN ? BN : ((nrem + 15) & ~15), AMN, BMN);   // see the producer
    const int kb0 = tz * kb_per_split;
    const int nkb = min(nkb_all, kb0 + kb_per_split) - kb0;
    const int as = li & 1;
    const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
    mbar_wait_cluster(tempty + 8 * as, (((uint32_t)li >> 1) & 1u) ^ 1u);   // BOTH CTAs have drained the tile two tiles back
    tc_fence_after();
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(it / Cfg::STAGES) & 1u;
      mbar_wait_cluster(full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        uint32_t acc = kb > 0 ? 1u : 0u;
        const uint32_t sa = base + s * Cfg::STAGE;
        const uint32_t sb = sa + P * Cfg::A_PLANE;
#pragma unroll
        for (int k = 0; k < UG_BK / 16; ++k) {
          uint64_t da[P], db[P];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            da[p] = AMN ? make_mnmajor_sw128_desc(sa + p * Cfg::A_PLANE + k * 2048)
                        : make_kmajor_sw128_desc(sa + p * Cfg::A_PLANE + k * 32);
            db[p] = BMN ? make_mnmajor_sw128_desc(sb + p * Cfg::B_PLANE + k * 2048)
                        : make_kmajor_sw128_desc(sb + p * Cfg::B_PLANE + k * 32);
          }
          if (P == 3) {
            umma_f16_pair(tacc, da[0], db[2], idesc, acc); acc = 1;
            umma_f16_pair(tacc, da[2], db[0], idesc, acc);
            umma_f16_pair(tacc, da[1], db[1], idesc, acc);
          }
          if (P >= 2) {
            umma_f16_pair(tacc, da[0], db[1], idesc, acc); acc = 1;
            umma_f16_pair(tacc, da[1], db[0], idesc, acc);
          }
          umma_f16_pair(tacc, da[0], db[0], idesc, acc); acc = 1;
        }
        umma_commit_pair(empty0 + 8 * s);
        if (kb == nkb - 1) umma_commit_pair(tfull + 8 * as);
      }
      __syncwarp();
    }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 accumulator rows -> 32 x 32 boxes in shared memory -> TMA stores =====
    // A warp reads its 32 TMEM lanes 32 columns at a time (lane = row), adds the bias, writes the row into a SWIZZLE_128B
    // box (16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free) and one lane hands the box to the TMA unit, which
    // clips it at the tensor's bounds; two boxes per warp alternate, the load of chunk cc + 1 is in flight meanwhile.
    // beta: the box is ADDED to C by the TMA unit (cp.reduce ... .add).  tma_out = 0 (C not 16-byte aligned / tiny
    // shapes): padded tile + st.global as in round 1.
    const int q = warp & 3;
    const int tma_out = (beta >> 1) & 1;
    beta &= 1;
    const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    float* const C0 = C;
    const uint32_t tempty_leader = mapa_u32(tempty, 0);
    uint8_t* const wst = epi_all + (warp - 2) * Cfg::EPI_WARP;
    const uint32_t wst_s = smem_u32(wst);
    const bool has_bias = bias0 != nullptr || bias1 != nullptr;
    int li = 0;
    uint32_t g = 0;                                   // running box counter of this warp (buffer = g & 1)
    for (int tile = pair; tile < total_tiles; tile += npairs, ++li) {
    const int tz = tile / (tiles_n * tiles_m2), ty = (tile / tiles_n) % tiles_m2, tx = tile % tiles_n;
    const int m0 = ty * 2 * UG_BM + (int)rank * UG_BM, n0 = tx * BN;
    const int as = li & 1;
    C = C0 + (size_t)tz * c_split_stride;
    float* bias_s = bias_all + as * 256;
    if (has_bias) {
      for (int j = threadIdx.x - 64; j < BN; j += 128) {
        float bsum = 0.f;
        if (n0 + j < N) {
          if (bias0) bsum += __ldg(bias0 + n0 + j);
          if (bias1) bsum += __ldg(bias1 + n0 + j);
        }
        bias_s[j] = bsum;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(tfull + 8 * as, ((uint32_t)li >> 1) & 1u);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(q * 32) << 16);
    const int nch = min(BN / 32, (N - n0 + 31) / 32);
    if (tma_out) {
      const int row0 = m0 + q * 32;
      float va[32], vb[32];
      tmem_ld32_nowait(tacc, va);
#pragma unroll 1
      for (int cc = 0; cc < nch; cc += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float* v = h ? vb : va;
          const int c = cc + h;
          if (c < nch) {
            tmem_wait_ld(v);
            if (c + 1 < nch) tmem_ld32_nowait(tacc + (uint32_t)((c + 1) * 32), h ? va : vb);
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
              }
            }
            const uint32_t boff = (g & 1u) * 4096u;
            if (lane == 0) bulk_wait_read<1>();         // the box stored two chunks back has been read out of this buffer
            __syncwarp();
            uint8_t* rowp = wst + boff + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < M) {
              if (beta) tma_reduce_add_3d(&mapC, wst_s + boff, n0 + c * 32, row0, tz);
              else tma_store_3d(&mapC, wst_s + boff, n0 + c * 32, row0, tz);
              bulk_commit();
            }
            ++g;
          }
        }
      }
    } else {
    float* stg = reinterpret_cast<float*>(wst);
    const int sr = lane >> 3, sc = (lane & 7) * 4;
#pragma unroll 1
    for (int cc = 0; cc < nch; ++cc) {
      float v[32];
      tmem_ld32(tacc + (uint32_t)(cc * 32), v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stg + lane * Cfg::EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int col = n0 + cc * 32 + sc;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bias) bb = *reinterpret_cast<const float4*>(bias_s + cc * 32 + sc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + sr, row = m0 + q * 32 + r;
        if (row >= M || col >= N) continue;
        const float4 x = *reinterpret_cast<const float4*>(stg + r * Cfg::EPI_PITCH + sc);
        float o4[4] = {x.x + bb.x, x.y + bb.y, x.z + bb.z, x.w + bb.w};
        float* cp = C + (size_t)row * ldc + col;
        if (vec && col + 3 < N) {
          if (beta) {
            const float4 o = *reinterpret_cast<const float4*>(cp);
            o4[0] += o.x; o4[1] += o.y; o4[2] += o.z; o4[3] += o.w;
          }
          *reinterpret_cast<float4*>(cp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (col + t < N) cp[t] = (beta ? cp[t] : 0.f) + o4[t];
        }
      }
      __syncwarp();
    }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (rank == 0) mbar_arrive(tempty + 8 * as);
      else mbar_arrive_remote(tempty_leader + 8 * as);
    }
    }
    if (tma_out && lane == 0) bulk_wait_all();        // the stores are complete (global writes performed) before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the leader's MMAs into the peer's tensor memory and both epilogues are done
  if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 -> P bf16 planes, K-major [P][rows][Kp]
// ------------------------------------------------------------------------------------------------
// row-major fp32 [rows x K] (ld) -> planes [P][rows][Kp]
template <int P>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ src, int rows, int K, int ld, int Kp, __nv_bfloat16* __restrict__ dst) {
  pdl_entry();
  const int K4 = Kp >> 2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * K4) return;
  const int k = (int)(i % K4) * 4;
  const int64_t row = i / K4;
  const float* s = src + row * ld + k;
  float x[4];
  if (k + 3 < K && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
    float4 t = *reinterpret_cast<const float4*>(s);
    x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (k + j < K) ? s[j] : 0.f;
  }
  __nv_bfloat16 p[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split3(x[j], p[0][j], p[1][j], p[2][j]);
  const int64_t plane = (int64_t)rows * Kp;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(p[q][0]) | ((uint32_t)__bfloat16_as_ushort(p[q][1]) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(p[q][2]) | ((uint32_t)__bfloat16_as_ushort(p[q][3]) << 16);
    *reinterpret_cast<uint2*>(dst + q * plane + row * Kp + k) = o;
  }
}

// several matrices in ONE launch (all weight matrices of a step): blockIdx.y selects the job
struct SplitJob { const float* src; __nv_bfloat16* dst; int rows, K, ld, Kp; };
struct SplitJobs { SplitJob j[16]; };
template <int P>
__global__ void __launch_bounds__(256) split_planes_batched_kernel(SplitJobs jobs) {
  pdl_entry();
  const SplitJob jb = jobs.j[blockIdx.y];
  const int K4 = jb.Kp >> 2;
  const int64_t total = (int64_t)jb.rows * K4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K4) * 4;
    const int64_t row = i / K4;
    const float* s = jb.src + row * jb.ld + k;
    float x[4];
    if (k + 3 < jb.K && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
      float4 t = *reinterpret_cast<const float4*>(s);
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = (k + j < jb.K) ? s[j] : 0.f;
    }
    __nv_bfloat16 p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(x[j], p[0][j], p[1][j], p[2][j]);
    const int64_t plane = (int64_t)jb.rows * jb.Kp;
#pragma unroll
    for (int q = 0; q < P; ++q) {
      uint2 o;
      o.x = (uint32_t)__bfloat16_as_ushort(p[q][0]) | ((uint32_t)__bfloat16_as_ushort(p[q][1]) << 16);
      o.y = (uint32_t)__bfloat16_as_ushort(p[q][2]) | ((uint32_t)__bfloat16_as_ushort(p[q][3]) << 16);
      *reinterpret_cast<uint2*>(jb.dst + q * plane + row * jb.Kp + k) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int umma_workspace_create(UmmaWorkspace** out, size_t transient_bytes, size_t static_bytes, size_t act_bytes) {
  UmmaWorkspace* ws = new UmmaWorkspace();
  ws->static_bytes = (static_bytes + 1023) & ~(size_t)1023;
  ws->act_bytes = (act_bytes + 1023) & ~(size_t)1023;
  ws->side_bytes = transient_bytes ? (size_t)48 << 20 : 0;           // split-K partials of the head GEMMs: <= 8 x 500 x 1024 x 4 B
  ws->bytes = ws->static_bytes + ws->act_bytes + transient_bytes + ws->side_bytes;
  ws->defer_bytes = transient_bytes ? (size_t)80 << 20 : 0;          // 2 x 40 MB: 9 splits of a [2048 x 512] fp32 tile set = 37.7 MB
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ws->base), ws->bytes + ws->defer_bytes);
  if (e == cudaSuccess && ws->defer_bytes) {
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&ws->reduce_gemm_done[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ws->reduce_done[i], cudaEventDisableTiming);
    }
  }
  if (e != cudaSuccess) {
    set_error(std::string("umma workspace cudaMalloc: ") + cudaGetErrorString(e));
    delete ws;
    return 1;
  }
  *out = ws;
  return 0;
}
void umma_workspace_destroy(UmmaWorkspace* ws) {
  if (!ws) return;
  for (int i = 0; i < 2; ++i) {
    if (ws->reduce_gemm_done[i]) cudaEventDestroy(ws->reduce_gemm_done[i]);
    if (ws->reduce_done[i]) cudaEventDestroy(ws->reduce_done[i]);
  }
  cudaFree(ws->base);
  delete ws;
}
// the static operands (weights) changed: drop their cached planes
void umma_workspace_invalidate(UmmaWorkspace* ws) {
  if (!ws) return;
  ws->cache.clear();
  ws->static_top = 0;
}
// a new forward pass overwrites the activations: drop their cached planes
void umma_workspace_new_forward(UmmaWorkspace* ws) {
  if (!ws) return;
  ws->act_cache.clear();
  ws->act_top = 0;
  ws->reduce_stream = nullptr;       // a step that failed half-way must not leave the deferral on
}

// fp32 row-major [rows x cols] (ld) -> bf16 planes [P][rows][colsp]; no transposition: an MN-major operand is
// consumed as such through MN-major shared-memory descriptors
static int prepare_planes_impl(UmmaWorkspace* ws, cudaStream_t s, int P, const float* src, int rows, int K,
                               int ld, int is_static, __nv_bfloat16** out, int* Kp_out, bool reserve_only) {
  const int Kp = (K + 7) & ~7;
  *Kp_out = Kp;
  PlaneKey key{src, rows, K, ld, 1, P};
  if (is_static == 1) {
    auto it = ws->cache.find(key);
    if (it != ws->cache.end()) { *out = it->second; return 0; }
  } else if (is_static == 2) {
    auto it = ws->act_cache.find(key);
    if (it != ws->act_cache.end()) { *out = it->second; return 0; }
  }
  const size_t need = ((size_t)P * rows * Kp * 2 + 1023) & ~(size_t)1023;
  const size_t trans0 = ws->tbase();
  __nv_bfloat16* dst;
  if (is_static == 1 && ws->static_top + need <= ws->static_bytes) {
    dst = reinterpret_cast<__nv_bfloat16*>(ws->base + ws->static_top);
    ws->static_top += need;
    ws->cache[key] = dst;
  } else if (is_static == 2 && ws->act_top + need <= ws->act_bytes) {
    dst = reinterpret_cast<__nv_bfloat16*>(ws->base + ws->static_bytes + ws->act_top);
    ws->act_top += need;
    ws->act_cache[key] = dst;
  } else {
    NVQA_CHECK(trans0 + ws->ttop() + need <= ws->tlimit(), "umma workspace too small");
    dst = reinterpret_cast<__nv_bfloat16*>(ws->base + trans0 + ws->ttop());
    ws->ttop() += need;
  }
  if (reserve_only) {             // reserve_planes: the producer kernel of `src` writes the planes itself
    *out = dst;
    return 0;
  }
  int64_t n = (int64_t)rows * (Kp / 4);
  const dim3 sgrid((unsigned)ceil_div(n, 256));
  if (P == 1) NVQA_CUDA(launch_pdl(split_planes_kernel<1>, sgrid, dim3(256), 0, s, src, rows, K, ld, Kp, dst));
  else if (P == 2) NVQA_CUDA(launch_pdl(split_planes_kernel<2>, sgrid, dim3(256), 0, s, src, rows, K, ld, Kp, dst));
  else NVQA_CUDA(launch_pdl(split_planes_kernel<3>, sgrid, dim3(256), 0, s, src, rows, K, ld, Kp, dst));
  NVQA_LAUNCHED();
  *out = dst;
  return 0;
}

// Reserves (and registers in the cache of class 1 / 2) the planes of the fp32 matrix `src` WITHOUT splitting it: the kernel
// that produces `src` writes the planes too (layout [P][rows][pitch], pitch = K rounded up to 8), so the GEMMs that later
// name `src` as an operand find them ready.  Returns the existing planes if `src` is already cached.
int reserve_planes(UmmaWorkspace* ws, int P, const float* src, int rows, int K, int ld, int cache_class, __nv_bfloat16** out,
                   int* pitch_out) {
  NVQA_CHECK(cache_class == 1 || cache_class == 2, "reserve_planes: only cached operand classes can be produced upstream");
  return prepare_planes_impl(ws, nullptr, P, src, rows, K, ld, cache_class, out, pitch_out, true);
}
int prepare_planes(UmmaWorkspace* ws, cudaStream_t s, int P, const float* src, int rows, int K, int ld, int is_static,
                   __nv_bfloat16** out, int* Kp_out) {
  return prepare_planes_impl(ws, s, P, src, rows, K, ld, is_static, out, Kp_out, false);
}

// Pre-split up to 16 weight matrices (row-major fp32 [rows x K], leading dimension K) into cached planes with ONE launch;
// matrices already cached are skipped.  Called at the start of a forward pass: the per-GEMM prepare_planes then hits.
int presplit_weights(UmmaWorkspace* ws, cudaStream_t s, int P, const float* const* src, const int* rows, const int* K, int n) {
  SplitJobs jobs;
  int nj = 0;
  int64_t most = 0;
  for (int i = 0; i < n && nj < 16; ++i) {
    if (!src[i]) continue;
    PlaneKey key{src[i], rows[i], K[i], K[i], 1, P};
    if (ws->cache.find(key) != ws->cache.end()) continue;
    const int Kp = (K[i] + 7) & ~7;
    const size_t need = ((size_t)P * rows[i] * Kp * 2 + 1023) & ~(size_t)1023;
    if (ws->static_top + need > ws->static_bytes) continue;           // no room: split per GEMM as before
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ws->base + ws->static_top);
    ws->static_top += need;
    ws->cache[key] = dst;
    jobs.j[nj++] = SplitJob{src[i], dst, rows[i], K[i], K[i], Kp};
    most = std::max<int64_t>(most, (int64_t)rows[i] * (Kp / 4));
  }
  if (!nj) return 0;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(most, 256), 2048), nj);
  if (P == 1) NVQA_CUDA(launch_pdl(split_planes_batched_kernel<1>, grid, dim3(256), 0, s, jobs));
  else if (P == 2) NVQA_CUDA(launch_pdl(split_planes_batched_kernel<2>, grid, dim3(256), 0, s, jobs));
  else NVQA_CUDA(launch_pdl(split_planes_batched_kernel<3>, grid, dim3(256), 0, s, jobs));
  NVQA_LAUNCHED();
  return 0;
}

// tensor map over planes [P][rows][colsp]: box = 64 columns (128 B, SWIZZLE_128B) x box_rows x 1 plane
int get_map(UmmaWorkspace* ws, const __nv_bfloat16* planes, int rows, int Kp, int P, int box_rows, CUtensorMap* out,
            long long plane_stride, int box_planes) {
  if (plane_stride <= 0) plane_stride = (long long)rows * Kp;
  MapKey key{planes, rows, Kp, P, box_rows, plane_stride, box_planes};
  auto it = ws->maps.find(key);
  if (it != ws->maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  NVQA_CHECK(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)P};
  cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)UG_BK, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(planes), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return 1;
  }
  if (ws->maps.size() > 4096) ws->maps.clear();
  ws->maps[key] = m;
  *out = m;
  return 0;
}

int get_map_kb(UmmaWorkspace* ws, const __nv_bfloat16* planes, int rows, int Kp, int P, int box_rows, int box_kb,
               CUtensorMap* out, long long plane_stride) {
  NVQA_CHECK(Kp % UG_BK == 0 && box_kb >= 1 && box_kb <= Kp / UG_BK, "get_map_kb: pitch must be a multiple of 64");
  if (plane_stride <= 0) plane_stride = (long long)rows * Kp;
  MapKey key{planes, rows, Kp, P, box_rows, plane_stride, P + 256 * box_kb};        // (3-D keys have box_planes <= 3)
  auto it = ws->maps.find(key);
  if (it != ws->maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  NVQA_CHECK(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)UG_BK, (cuuint64_t)rows, (cuuint64_t)P, (cuuint64_t)(Kp / UG_BK)};
  cuuint64_t strides[3] = {(cuuint64_t)Kp * 2, (cuuint64_t)plane_stride * 2, (cuuint64_t)UG_BK * 2};
  cuuint32_t box[4] = {(cuuint32_t)UG_BK, (cuuint32_t)box_rows, (cuuint32_t)P, (cuuint32_t)box_kb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(planes), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D) failed with code " + std::to_string((int)r));
    return 1;
  }
  if (ws->maps.size() > 4096) ws->maps.clear();
  ws->maps[key] = m;
  *out = m;
  return 0;
}

struct UgLaunch {
  int a_row0 = 0, b_row0 = 0, splits = 1, kb_per_split = 1 << 30, cta_cap = 0;
  long long c_split_stride = 0;
  const CUtensorMap* mapc = nullptr;       // pair kernel: output tensor map of the TMA-store epilogue (null: st.global)
};

// tensor map over the fp32 output [splits][M][N] (row pitch ld, split stride in elements): boxes of 32 x 32, SWIZZLE_128B
static int get_cmap(UmmaWorkspace* ws, float* C, int M, int N, int ld, int splits, long long split_stride, CUtensorMap* out) {
  if (split_stride <= 0) split_stride = (long long)M * ld;
  MapKey key{reinterpret_cast<const __nv_bfloat16*>(C), M, ld, splits, N, split_stride, 7777};
  auto it = ws->maps.find(key);
  if (it != ws->maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  NVQA_CHECK(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)split_stride * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed with code " + std::to_string((int)r));
    return 1;
  }
  if (ws->maps.size() > 4096) ws->maps.clear();
  ws->maps[key] = m;
  *out = m;
  return 0;
}

template <int BN, int P, bool AMN, bool BMN>
static int launch_umma(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, float* C,
                       int ldc, bool beta, const float* b0, const float* b1, const UgLaunch& L) {
  using Cfg = UgCfg<BN, P>;
  static bool attr_set = false;
  if (!attr_set) {
    NVQA_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<BN, P, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0, persist = -1;
  if (!num_sms) {
    int dev = 0;
    NVQA_CUDA(cudaGetDevice(&dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (persist < 0) { const char* e = getenv("NVQA_GEMM_PERSIST"); persist = e ? atoi(e) : 1; }   // 0: one tile per CTA
  const int tn = ceil_div(N, BN), tm = ceil_div(M, UG_BM);
  const long long total = (long long)tn * tm * L.splits;
  NVQA_CHECK(total < (1ll << 31), "umma_gemm: too many tiles");
  dim3 grid((unsigned)(persist ? std::min<long long>(total, L.cta_cap > 0 ? std::min(L.cta_cap, num_sms) : num_sms) : total));
  NVQA_CUDA(launch_pdl(umma_gemm_kernel<BN, P, AMN, BMN>, grid, dim3(UG_THREADS), Cfg::SMEM, s, ma, mb, M, N, K, C, ldc,
                       beta ? 1 : 0, b0, b1, L.a_row0, L.b_row0, L.kb_per_split, L.c_split_stride, tn, tm, L.splits));
  NVQA_LAUNCHED();
  return 0;
}

template <int BN, int P>
static int launch_umma_major(bool amn, bool bmn, cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, int M,
                             int N, int K, float* C, int ldc, bool beta, const float* b0, const float* b1,
                             const UgLaunch& L) {
  if (!amn && !bmn) return launch_umma<BN, P, false, false>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  if (!amn && bmn) return launch_umma<BN, P, false, true>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  if (amn && !bmn) return launch_umma<BN, P, true, false>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  return launch_umma<BN, P, true, true>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
}


template <int BN, int P, bool AMN, bool BMN>
static int launch_umma_pair(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, float* C,
                            int ldc, bool beta, const float* b0, const float* b1, const UgLaunch& L) {
  using Cfg = UgCfg2<BN, P>;
  static bool attr_set = false;
  const void* fn = (const void*)umma_gemm_pair_kernel<BN, P, AMN, BMN>;
  if (!attr_set) {
    NVQA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    NVQA_CUDA(cudaGetDevice(&dev));
    NVQA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int tn = ceil_div(N, BN), tm2 = ceil_div(M, 2 * UG_BM), splits = L.splits, kbs = L.kb_per_split;
  long long cstride = L.c_split_stride;
  int a0 = L.a_row0, b0r = L.b_row0, beta_i = (beta ? 1 : 0) | (L.mapc ? 2 : 0);      // bit 1: TMA-store epilogue
  const long long total = (long long)tn * tm2 * splits;
  NVQA_CHECK(total < (1ll << 30), "umma_gemm: too many tiles");
  const int sm_cap = L.cta_cap > 0 ? std::min(L.cta_cap, num_sms) : num_sms;
  const int pairs = (int)std::min<long long>(total, std::max(1, sm_cap / 2));
  const CUtensorMap& mc = L.mapc ? *L.mapc : ma;                                       // (unused by the kernel when bit 1 is clear)
  void* args[] = {(void*)&ma, (void*)&mb, (void*)&mc, &M, &N, &K, &C, &ldc, &beta_i, (void*)&b0, (void*)&b1, &a0, &b0r, &kbs, &cstride, &tn, &tm2, &splits};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(UG_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // common.cuh: pdl_entry() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  NVQA_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ++g_launches;
  return 0;
}
template <int BN, int P>
static int launch_umma_pair_major(bool amn, bool bmn, cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, int M,
                                  int N, int K, float* C, int ldc, bool beta, const float* b0, const float* b1,
                                  const UgLaunch& L) {
  if (!amn && !bmn) return launch_umma_pair<BN, P, false, false>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  if (!amn && bmn) return launch_umma_pair<BN, P, false, true>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  if (amn && !bmn) return launch_umma_pair<BN, P, true, false>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
  return launch_umma_pair<BN, P, true, true>(s, ma, mb, M, N, K, C, ldc, beta, b0, b1, L);
}

// C[m][n] = (beta ? C : 0) + sum_z part[z][m][n] + bias0[n] + bias1[n]   (fixed summation order: deterministic)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, int splits, long long stride, int M, int N, int ldp,
                     float* __restrict__ C, int ldc, int beta, const float* __restrict__ bias0,
                     const float* __restrict__ bias1) {
  pdl_entry();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int n = (int)(i % N);
  const long long m = i / N;
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += part[(size_t)z * stride + m * ldp + n];
  if (bias0) acc += bias0[n];
  if (bias1) acc += bias1[n];
  float* dst = C + m * ldc + n;
  *dst = (beta ? *dst : 0.f) + acc;
}

// operand = fp32 source to be split into planes here, or ready-made planes produced by an upstream kernel
static int resolve_operand(UmmaWorkspace* ws, cudaStream_t s, int P, const UmmaOperand& op, int rows, int cols,
                           const __nv_bfloat16** planes, int* pitch, long long* plane_stride, int* row0, int* bound) {
  if (op.planes) {
    *planes = op.planes; *pitch = op.pitch; *plane_stride = (long long)op.plane_rows * op.pitch;
    *row0 = op.row_offset; *bound = op.row_offset + rows;
    return 0;
  }
  __nv_bfloat16* p = nullptr;
  NVQA_TRY(prepare_planes(ws, s, P, op.src, rows, cols, op.ld, op.is_static, &p, pitch));
  *planes = p; *plane_stride = (long long)rows * *pitch; *row0 = 0; *bound = rows;
  return 0;
}

int umma_gemm_ops(cudaStream_t s, int planes, const UmmaOperand& A, const UmmaOperand& B, int M, int N, int K, float* C,
                  int ldc, bool beta, const float* bias0, const float* bias1, UmmaWorkspace* ws) {
  NVQA_CHECK(ws, "umma_gemm: no workspace");
  NVQA_CHECK(planes >= 1 && planes <= 3, "umma_gemm: planes must be 1..3");
  if (M <= 0 || N <= 0) return 0;
  ws->ttop() = 0;                         // stream order makes the previous GEMM's transient planes reusable
  // planes keep the source's row-major shape: [M x K] / [N x K] when K-major, [K x M] / [K x N] when MN-major
  const __nv_bfloat16 *pa = nullptr, *pb = nullptr;
  int pitch_a = 0, pitch_b = 0, bound_a = 0, bound_b = 0;
  long long ps_a = 0, ps_b = 0;
  UgLaunch L;
  NVQA_TRY(resolve_operand(ws, s, planes, A, A.kmajor ? M : K, A.kmajor ? K : M, &pa, &pitch_a, &ps_a, &L.a_row0, &bound_a));
  NVQA_TRY(resolve_operand(ws, s, planes, B, B.kmajor ? N : K, B.kmajor ? K : N, &pb, &pitch_b, &ps_b, &L.b_row0, &bound_b));
  // tile shape: the widest N tile that still gives every SM a CTA; split K when the grid would be too small
  // (a side-stream GEMM owns only cta_cap SMs: shaped as for a machine of that size)
  int num_sms = ws->side && ws->cta_cap > 0 ? ws->cta_cap : 148;
  L.cta_cap = ws->side ? ws->cta_cap : 0;
  const int mt = ceil_div(M, UG_BM);
  // CTA pairs (cta_group::2, 256 x BN tiles) for the wide tiles of two-plane products: NVQA_GEMM_PAIR=0 disables
  static int pair_on = -1, shape_v2 = -1;
  if (pair_on < 0) { const char* e = getenv("NVQA_GEMM_PAIR"); pair_on = e ? atoi(e) : 1; }
  if (shape_v2 < 0) { const char* e = getenv("NVQA_GEMM_SHAPE_V2"); shape_v2 = e ? atoi(e) : 1; }
  const bool pairable = pair_on && M > UG_BM;                      // (and BN >= 128, below)
  const int slots = pairable ? num_sms / 2 : num_sms;              // work units (pair tiles / tiles) the machine runs at once
  const int mrows = pairable ? ceil_div(M, 2 * UG_BM) : mt;
  int BN = 64;
  // the pair kernel multiplies only the next multiple of 16 of the columns a (last) tile really has, so a 256-wide tile
  // also serves 128 < N < 256 (N = 200, the embedding width: 208 instead of 2 x 128 columns)
  const int nmin256 = (shape_v2 && pairable) ? 129 : 256;
  const int n256 = ceil_div(N, 256) == 1 ? ((N + 15) & ~15) : 256;      // effective width of a 256-tile
  if (planes <= 2 && N >= nmin256 && (long)mt * ceil_div(N, 256) >= num_sms / 2) {
    BN = 256;
    // a persistent grid runs ceil(units / slots) waves: half-width tiles (8 % more operand traffic per flop) win when
    // they fill the last wave better -- layer-2 dgrad [13000 x 512]: 102 pair tiles on 74 pairs = 2 waves of 256 columns
    // against 204 = 3 waves of 128 (0.122 -> 0.107 ms for the two dgrad GEMMs of the step)
    const long w256 = (long)ceil_div(mrows * ceil_div(N, 256), slots) * n256 * 100;
    const long w128 = (long)ceil_div(mrows * ceil_div(N, 128), slots) * 128 * 108;
    if (shape_v2 && w128 < w256) BN = 128;
  } else if ((long)mt * ceil_div(N, 128) >= num_sms / 2) BN = 128;
  else if (planes <= 2 && N >= nmin256 && K >= 2048) BN = 256;    // few tiles but a long K: wide tiles + split-K
  else if (N >= 128 && K >= 2048) BN = 128;
  const int tiles = mt * ceil_div(N, BN);
  const int nkb = ceil_div(K, UG_BK);
  int splits = 1;
  if (tiles < num_sms / 2 && nkb >= 16) {
    if (shape_v2) {
      // waves x (k-blocks per split + ~3 k-blocks of prologue / epilogue): more than one wave is allowed when it fills the
      // machine better -- weight gradients [2048 x 512], K = 13000: 16 pair tiles x 4 splits = 64 of 74 pairs busy for 51
      // k-blocks against x 9 = 144 units = 2 waves of 23
      const int units = (BN >= 128 ? mrows : mt) * ceil_div(N, BN), sl = BN >= 128 ? slots : num_sms;
      // a split product also costs a reduction kernel on the stream (~4 us with its launch), expressed in k-blocks of this
      // tile width (~6 BN cycles each); free when the reduction is deferred to the side stream
      const int red_kb = ws->reduce_stream && !ws->side ? 0 : 8000 / (6 * BN);
      long best = -1;
      for (int sp = 1; sp <= 16 && sp <= (nkb / 8 > 1 ? nkb / 8 : 1); ++sp) {
        const int kbs = ceil_div(nkb, sp), eff = ceil_div(nkb, kbs);
        const long cost = (long)ceil_div(units * eff, sl) * (kbs + 3) + (eff > 1 ? red_kb : 0);
        if (best < 0 || cost < best) { best = cost; splits = eff; }
      }
    } else {
      splits = num_sms / tiles;
      if (splits > 8) splits = 8;
      if (splits > nkb / 8) splits = nkb / 8;
      if (splits < 1) splits = 1;
    }
  }
  // experiments: NVQA_GEMM_OVERRIDE="MxNxK:BN:splits,..." replaces the shape rule for the listed products
  {
    struct Ov { int M, N, K, BN, splits; };
    static std::vector<Ov> ov;
    static bool parsed = false;
    if (!parsed) {
      parsed = true;
      if (const char* e = getenv("NVQA_GEMM_OVERRIDE")) {
        Ov o;
        int used = 0;
        while (sscanf(e, "%dx%dx%d:%d:%d%n", &o.M, &o.N, &o.K, &o.BN, &o.splits, &used) == 5) {
          ov.push_back(o);
          e += used;
          if (*e == ',') ++e;
        }
      }
    }
    for (const Ov& o : ov)
      if (o.M == M && o.N == N && o.K == K && !ws->side && (o.BN == 64 || o.BN == 128 || (o.BN == 256 && planes <= 2))) {
        BN = o.BN;
        splits = o.splits < 1 ? 1 : (o.splits > nkb ? nkb : o.splits);
      }
  }
  float* Cout = C;
  int ldo = ldc, defer_slot = -1;
  bool beta_k = beta;
  const float *b0k = bias0, *b1k = bias1;
  if (splits > 1) {
    L.kb_per_split = ceil_div(nkb, splits);
    splits = ceil_div(nkb, L.kb_per_split);                       // no empty split
    L.splits = splits;
    const bool can_defer = ws->reduce_stream && !ws->side && ws->defer_bytes;
    const size_t room = can_defer ? ws->defer_bytes / 2 : ws->tlimit() - ws->tbase() - ws->ttop();
    while (splits > 2 && (size_t)splits * M * N * sizeof(float) > room) {        // fewer, longer splits if the partials do not fit
      L.kb_per_split = ceil_div(nkb, splits - 1);
      splits = ceil_div(nkb, L.kb_per_split);
    }
    L.splits = splits;
    const size_t need = (size_t)splits * M * N * sizeof(float);
    defer_slot = (ws->reduce_stream && !ws->side && ws->defer_bytes && need <= ws->defer_bytes / 2) ? ws->reduce_slot : -1;
    if (defer_slot >= 0) {
      ws->reduce_slot ^= 1;
      if (ws->reduce_pending[defer_slot]) NVQA_CUDA(cudaStreamWaitEvent(s, ws->reduce_done[defer_slot], 0));   // slot still being reduced
      Cout = reinterpret_cast<float*>(ws->base + ws->bytes + (size_t)defer_slot * (ws->defer_bytes / 2));
    } else {
      NVQA_CHECK(ws->tbase() + ws->ttop() + need <= ws->tlimit(), "umma workspace too small for split-K partials");
      Cout = reinterpret_cast<float*>(ws->base + ws->tbase() + ws->ttop());
      ws->ttop() += (need + 1023) & ~(size_t)1023;
    }
    ldo = N;
    L.c_split_stride = (long long)M * N;
    beta_k = false; b0k = b1k = nullptr;
  }
  const bool use_pair = pairable && BN >= 128;
  CUtensorMap ma, mb;
  NVQA_TRY(get_map(ws, pa, bound_a, pitch_a, planes, A.kmajor ? UG_BM : 64, &ma, ps_a));
  NVQA_TRY(get_map(ws, pb, bound_b, pitch_b, planes, B.kmajor ? (use_pair ? BN / 2 : BN) : 64, &mb, ps_b));
  CUtensorMap mc;
  static int tma_epi = -1;
  if (tma_epi < 0) { const char* e = getenv("NVQA_GEMM_TMA_STORE"); tma_epi = e ? atoi(e) : 1; }
  if (use_pair && tma_epi && M >= 32 && N >= 32 && (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(Cout) & 15) == 0 &&
      (L.c_split_stride & 3) == 0) {
    NVQA_TRY(get_cmap(ws, Cout, M, N, ldo, splits, L.c_split_stride, &mc));
    L.mapc = &mc;
  }
  int rc = 1;
#define NVQA_UGP(BN_, P_) \
  rc = launch_umma_pair_major<BN_, P_>(!A.kmajor, !B.kmajor, s, ma, mb, M, N, K, Cout, ldo, beta_k, b0k, b1k, L)
  if (use_pair) {
    if (BN == 256) { if (planes == 1) NVQA_UGP(256, 1); else NVQA_UGP(256, 2); }
    else { if (planes == 1) NVQA_UGP(128, 1); else if (planes == 2) NVQA_UGP(128, 2); else NVQA_UGP(128, 3); }
  } else
#undef NVQA_UGP
#define NVQA_UG(BN_, P_) \
  rc = launch_umma_major<BN_, P_>(!A.kmajor, !B.kmajor, s, ma, mb, M, N, K, Cout, ldo, beta_k, b0k, b1k, L)
  if (BN == 256) {
    if (planes == 1) NVQA_UG(256, 1); else NVQA_UG(256, 2);
  } else if (BN == 128) {
    if (planes == 1) NVQA_UG(128, 1); else if (planes == 2) NVQA_UG(128, 2); else NVQA_UG(128, 3);
  } else {
    if (planes == 1) NVQA_UG(64, 1); else if (planes == 2) NVQA_UG(64, 2); else NVQA_UG(64, 3);
  }
#undef NVQA_UG
  if (rc) return rc;
  if (splits > 1) {
    cudaStream_t rs = s;
    if (defer_slot >= 0) {
      rs = ws->reduce_stream;
      NVQA_CUDA(cudaEventRecord(ws->reduce_gemm_done[defer_slot], s));
      NVQA_CUDA(cudaStreamWaitEvent(rs, ws->reduce_gemm_done[defer_slot], 0));
    }
    NVQA_CUDA(launch_pdl(splitk_reduce_kernel, dim3((unsigned)ceil_div((long long)M * N, 256)), dim3(256), 0, rs, Cout, splits,
                         L.c_split_stride, M, N, N, C, ldc, beta ? 1 : 0, bias0, bias1));
    NVQA_LAUNCHED();
    if (defer_slot >= 0) {
      NVQA_CUDA(cudaEventRecord(ws->reduce_done[defer_slot], rs));
      ws->reduce_pending[defer_slot] = true;
    }
  }
  return 0;
}

int umma_gemm(cudaStream_t s, int planes, bool a_kmajor, bool b_kmajor, int M, int N, int K, const float* A, int lda,
              const float* B, int ldb, float* C, int ldc, bool beta, const float* bias0, const float* bias1,
              UmmaWorkspace* ws, int a_static, int b_static) {
  UmmaOperand a, b;
  a.src = A; a.ld = lda; a.kmajor = a_kmajor; a.is_static = a_static;
  b.src = B; b.ld = ldb; b.kmajor = b_kmajor; b.is_static = b_static;
  return umma_gemm_ops(s, planes, a, b, M, N, K, C, ldc, beta, bias0, bias1, ws);
}

}  // namespace nvqa
