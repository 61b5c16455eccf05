// K11: the ConvNeXt block's MLP as ONE tcgen05 kernel per direction (include/b200at_model.h: b200at_mlp_fused).
//
//   forward   out = x + GELU(t2 W1^T + b1) (gamma W2)^T + gamma b2          models/convnext.py:42-49
//   backward  dt2 = ((dout (gamma W2)) * GELU'(z + b1)) W1                   (input gradient of the same lines)
//
// Both are   OUT[M,C] = f( A[M,C] Wa[4C,C]^T ; Z ) Wb[C,4C]^T   with an elementwise f on the 4C-wide hidden, so the hidden
// activation (a / da / dz: 4C bf16 per pixel, the largest tensors of the block) never goes to HBM between the two
// GEMMs.  Unfused (three kernels) the forward moves 8 B and the input-gradient pass 10 B per hidden element; fused
// they move 2 B (z written once for the backward / z read once).
//
// Per 128-row tile, the hidden dimension is walked in chunks of 64 columns:
//   GEMM-a  acc_a[128 x 64]  = A_tile[128 x C] . Wa_chunk[64 x C]^T          (TMEM, double buffered by chunk parity)
//   f       16 epilogue warps in two groups of 8 (group = chunk parity): tcgen05.ld -> round to bf16 where the unfused
//           path stores (z / da) -> bias / GELU / GELU' -> bf16 -> st.shared in the K-major SWIZZLE_128B operand layout
//           (P[2]); the z chunk travels through its own swizzled buffer (Z[2]): TMA store forward, TMA load backward
//   GEMM-b  acc_b[128 x C] += P_chunk[128 x 64] . Wb_chunk[C x 64]^T          (TMEM, double buffered across tiles)
// then acc_b (+ bias2 + residual) -> bf16 -> global, deferred to after the warp's first chunk of the next tile.
// Persistent, one CTA per SM, 640 threads, mbarrier-only synchronisation:
//   warp 0      three independent TMA producer lanes: A tile + GEMM-a weight ring / GEMM-b weight ring / saved z (backward)
//   warp 1      MMA issuer: GEMM-a of chunk g+2 is issued as soon as acc_a of chunk g has been drained, before GEMM-b of g
//   warps 2, 3  TMA store issuers, one per chunk parity (z forward; the optional a / dz output straight from P);
//               warp 2 also allocates / frees TMEM
//   warps 4-19  the elementwise stage + final epilogue
// Two rules this kernel learnt the hard way: a buffer that an asynchronous writer (TMA) refills may only be
// released after the loads from it have RETURNED (mbarrier.arrive does not wait for them: `consume_loads`), and every
// shared-memory access goes through explicit ld/st.shared on 32-bit addresses (pointer arithmetic on the re-aligned
// dynamic shared base degrades to generic LD/ST).
// The hidden values are rounded to bf16 exactly where the unfused path stores them (z, da): forward z / a / out are
// bit-identical to the three-kernel path, backward dz within one bf16 step (profiles/debug/mlp_determinism.py).
// B200AT_MLP_DEBUG (race hunting only; bit 2 makes the optional output wrong on purpose): 1 = final epilogue at the tile
// end instead of deferred, 2 = skip the TMA store of P, 4 = plain instead of backed-off waits in the store warps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200at_gelu.cuh"
#include "b200at_launch.cuh"
#include "b200at_tcgen05.cuh"
#include "../../include/b200at_model.h"

namespace {

using namespace b200at_tc;
typedef __nv_bfloat16 bf16;

constexpr int kChunk = 64;                 // hidden columns per chunk = one SWIZZLE_128B k-block of GEMM-b
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);   // 640

template <int C>
struct MlpCfg {
  static constexpr int KB = (C + 63) / 64;                 // k-blocks of GEMM-a (K = C; TMA zero-fills the tail)
  static constexpr int KSteps = C / 16;                    // UMMA K = 16 steps of GEMM-a
  static constexpr int NC = 4 * C / kChunk;                // chunks per tile
  static constexpr int AStages = C <= 128 ? 2 : 1;
  static constexpr int WStages = 2;
  static constexpr int ABytes = KB * 128 * 128;            // KB x [128 rows x 128 B]
  static constexpr int WaBytes = KB * kChunk * 128;        // KB x [64 rows x 128 B]
  static constexpr int WbBytes = C * 128;                  // [C rows x 128 B]
  static constexpr int WStageBytes = WaBytes + WbBytes;
  static constexpr int PBytes = 128 * 128;                 // [128 rows x 128 B]
  static constexpr int OffW = AStages * ABytes;
  static constexpr int OffP = OffW + WStages * WStageBytes;
  static constexpr int OffZ = OffP + 2 * PBytes;           // z chunk staging [2][128 rows x 128 B]
  static constexpr int OffBias = OffZ + 2 * PBytes;        // fp32 bias1[4C], bias2[C]
  static constexpr int OffBars = OffBias + 5 * C * 4;
  static constexpr int Smem = 1024 + OffBars + 256;
  static constexpr int TmemCols = (2 * C + 2 * kChunk) <= 256 ? 256 : 512;   // acc_b[2] + acc_a[2]
  static_assert(C % 16 == 0 && C <= 256, "GEMM-b is one UMMA of N = C");
  static_assert(Smem <= 227 * 1024, "shared memory budget");
  static_assert(2 * C + 2 * kChunk <= 512, "TMEM budget");
};

struct MlpParams {
  const float* bias1;   // [4C]
  const float* bias2;   // [C] or null
  const bf16* residual; // [M][C] or null
  bf16* z;              // [M][4C]  forward: written (pre-activation without bias); backward: read
  bf16* p_out;          // [M][4C] or null: forward a = GELU(z + b1); backward dz
  bf16* out;            // [M][C]
  int M, tiles_m;
  int debug;            // B200AT_MLP_DEBUG bits (race hunting): 1 final epilogue at the tile end, 2 skip the P store, 4 plain waits in the store warps
};

template <int C, int MODE>   // MODE 0 forward, 1 backward
__global__ void __launch_bounds__(kThreads, 1) mlp_kernel(const __grid_constant__ CUtensorMap map_a,
                                                          const __grid_constant__ CUtensorMap map_wa,
                                                          const __grid_constant__ CUtensorMap map_wb,
                                                          const __grid_constant__ CUtensorMap map_z,
                                                          const __grid_constant__ CUtensorMap map_p,
                                                          const MlpParams p) {
  typedef MlpCfg<C> T;
  extern __shared__ __align__(1024) uint8_t mlp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mlp_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sW = smem + T::OffW;
  uint8_t* sP = smem + T::OffP;
  uint8_t* sZ = smem + T::OffZ;
  float* sBias1 = reinterpret_cast<float*>(smem + T::OffBias);
  float* sBias2 = sBias1 + 4 * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::OffBars);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* wa_full = bars + 4;       // [2]  GEMM-a weight chunk (released as soon as GEMM-a retires)
  uint64_t* wa_empty = bars + 6;      // [2]
  uint64_t* ta_full = bars + 8;       // [2]  GEMM-a accumulator ready
  uint64_t* ta_empty = bars + 10;     // [2]
  uint64_t* p_full = bars + 12;       // [2]  operand chunk written
  uint64_t* p_empty = bars + 14;      // [2]
  uint64_t* tb_full = bars + 26;      // [2]  GEMM-b accumulator of a tile complete
  uint64_t* tb_empty = bars + 28;     // [2]
  uint64_t* z_full = bars + 18;       // [2]  z chunk in shared memory (forward: written by the warps; backward: TMA)
  uint64_t* z_empty = bars + 20;      // [2]
  uint64_t* wb_full = bars + 22;      // [2]  GEMM-b weight chunk
  uint64_t* wb_empty = bars + 24;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wa) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wb) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_z) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_p) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
      mbar_init(&wa_full[s], 1); mbar_init(&wa_empty[s], 1);
      mbar_init(&wb_full[s], 1); mbar_init(&wb_empty[s], 1);
      mbar_init(&tb_full[s], 1); mbar_init(&tb_empty[s], kEpiWarps);
      mbar_init(&ta_full[s], 1); mbar_init(&ta_empty[s], kEpiWarps / 2);
      mbar_init(&p_full[s], kEpiWarps / 2); mbar_init(&p_empty[s], p.p_out ? 2 : 1);   // UMMA (+ the TMA store)
      mbar_init(&z_full[s], MODE == 0 ? kEpiWarps / 2 : 1); mbar_init(&z_empty[s], MODE == 0 ? 1 : kEpiWarps / 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)T::TmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4 * C; i += kThreads) sBias1[i] = p.bias1[i];
  for (int i = threadIdx.x; i < C; i += kThreads) sBias2[i] = p.bias2 ? p.bias2[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_b = tmem_base;                       // acc_b[b]: columns [b C, (b+1) C), b = tile parity
  const uint32_t tmem_a = tmem_base + (uint32_t)(2 * C);   // acc_a[s]: columns 2C + 64 s

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producers
    // Three independent streams, one lane each, so that a stream blocked on its ring never delays another one:
    // lane 0 the A tiles and the GEMM-a weight chunks (slot free as soon as GEMM-a retires: runs a full chunk ahead),
    // lane 1 the GEMM-b weight chunks (slot free when GEMM-b of two chunks ago retires), lane 2 the saved
    // pre-activation chunks of the backward pass.
    if (lane == 0) {
      uint32_t g = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++t) {
        const int as = t % T::AStages;
        const uint32_t aph = (uint32_t)(t / T::AStages) & 1u;
        mbar_wait_relaxed(&a_empty[as], aph ^ 1u);
        mbar_expect_tx(&a_full[as], T::ABytes);
#pragma unroll
        for (int kb = 0; kb < T::KB; ++kb)
          tma_load_2d(&map_a, &a_full[as], sA + as * T::ABytes + kb * (128 * 128), kb * 64, tile * 128);
        for (int j = 0; j < T::NC; ++j, ++g) {
          const int ws = (int)(g & 1u);
          mbar_wait_relaxed(&wa_empty[ws], ((g >> 1) & 1u) ^ 1u);
          uint8_t* wa = sW + ws * T::WaBytes;
          mbar_expect_tx(&wa_full[ws], T::WaBytes);
#pragma unroll
          for (int kb = 0; kb < T::KB; ++kb)
            tma_load_2d(&map_wa, &wa_full[ws], wa + kb * (kChunk * 128), kb * 64, j * kChunk);
        }
      }
    } else if (lane == 1) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
        for (int j = 0; j < T::NC; ++j, ++g) {
          const int ws = (int)(g & 1u);
          mbar_wait_relaxed(&wb_empty[ws], ((g >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&wb_full[ws], T::WbBytes);
          tma_load_2d(&map_wb, &wb_full[ws], sW + 2 * T::WaBytes + ws * T::WbBytes, j * kChunk, 0);
        }
      }
    } else if (lane == 2 && MODE == 1) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
        for (int j = 0; j < T::NC; ++j, ++g) {
          const int ws = (int)(g & 1u);
          mbar_wait_relaxed(&z_empty[ws], ((g >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&z_full[ws], T::PBytes);
          tma_load_2d(&map_z, &z_full[ws], sZ + ws * T::PBytes, j * kChunk, tile * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Issue order: GEMM-a of chunk g+2 goes out as soon as the elementwise stage has read acc_a of chunk g (ta_empty),
    // i.e. BEFORE waiting for that stage to finish chunk g -- so a group of epilogue warps always finds its next
    // accumulator complete and never waits for the tensor pipe; GEMM-b of chunk g follows when its operand is written.
    const uint32_t idesc_a = make_idesc(128, kChunk);
    const uint32_t idesc_b = make_idesc(128, C);
    const int my_tiles = (p.tiles_m - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * T::NC;     // chunks this CTA walks through
    uint32_t a_waited = 0;                                 // A tiles whose arrival this warp has observed
    auto gemm_a = [&](uint32_t gg) {
      const int s = (int)(gg & 1u);
      const uint32_t ph = (gg >> 1) & 1u;
      const uint32_t tt = gg / T::NC;                      // local tile index of chunk gg
      const int as = (int)(tt % T::AStages);
      if (tt >= a_waited) {                                // first chunk of a tile: its A tile must have landed
        mbar_wait(&a_full[as], (tt / T::AStages) & 1u);
        a_waited = tt + 1;
      }
      mbar_wait(&wa_full[s], ph);                          // weight chunk gg landed
      mbar_wait(&ta_empty[s], ph ^ 1u);                    // the elementwise stage has drained acc_a[s]
      tc_fence_after();
      if (elect_one()) {
        const uint8_t* a_tile = sA + as * T::ABytes;
        const uint8_t* wa = sW + s * T::WaBytes;
#pragma unroll
        for (int k = 0; k < T::KSteps; ++k) {
          const uint64_t da = make_desc(a_tile + (k >> 2) * (128 * 128)) + (uint64_t)(2 * (k & 3));
          const uint64_t db = make_desc(wa + (k >> 2) * (kChunk * 128)) + (uint64_t)(2 * (k & 3));
          umma(tmem_a + (uint32_t)(s * kChunk), da, db, idesc_a, k != 0);
        }
        umma_commit(&ta_full[s]);
        umma_commit(&wa_empty[s]);
        if (gg % T::NC == T::NC - 1) umma_commit(&a_empty[as]);   // last GEMM-a of the tile: A tile free when it retires
      }
      __syncwarp();
    };
    auto gemm_b = [&](uint32_t gg) {
      const int s = (int)(gg & 1u);
      const uint32_t ph = (gg >> 1) & 1u;
      const uint32_t tt = gg / T::NC;
      const int j = (int)(gg % T::NC);
      mbar_wait(&wb_full[s], ph);
      mbar_wait(&p_full[s], ph);                           // operand chunk written by its group of epilogue warps
      const int bb = (int)(tt & 1u);
      if (j == 0) mbar_wait(&tb_empty[bb], ((tt >> 1) & 1u) ^ 1u);   // the tile two back has been drained from acc_b[bb]
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = make_desc(sP + s * T::PBytes);
        const uint64_t db = make_desc(sW + 2 * T::WaBytes + s * T::WbBytes);
#pragma unroll
        for (int k = 0; k < kChunk / 16; ++k)
          umma(tmem_b + (uint32_t)(bb * C), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_b, (j | k) != 0);
        umma_commit(&p_empty[s]);
        umma_commit(&wb_empty[s]);
        if (j == T::NC - 1) umma_commit(&tb_full[bb]);
      }
      __syncwarp();
    };
    if (total > 0) {
      gemm_a(0);
      gemm_a(1);
      for (uint32_t g = 0; g < total; ++g) {
        const bool next_tile = (g % T::NC) + 2 >= (uint32_t)T::NC;   // chunk g+2 belongs to the next tile
        if (g + 2 < total && !next_tile) gemm_a(g + 2);
        gemm_b(g);
        if (g + 2 < total && next_tile) gemm_a(g + 2);     // may wait for the next A tile: keep GEMM-b ahead of it
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------------------------ TMA stores of the hidden chunks (z forward; a / dz)
    // warp 2 serves chunk parity 0, warp 3 parity 1: bulk-async groups are per thread, so each waits only for its own
    // buffer to be read before handing it back.
    if ((MODE == 0 || p.p_out != nullptr) && elect_one()) {
      const int s = warp - 2;
      for (int tile = blockIdx.x, t = 0; tile < p.tiles_m; tile += gridDim.x, ++t) {
        for (int j = s; j < T::NC; j += 2) {
          const uint32_t g = (uint32_t)(t * T::NC + j);
          const uint32_t ph = (g >> 1) & 1u;
          if (MODE == 0) {
            mbar_wait_relaxed(&z_full[s], ph);               // the warps' writes are fenced into the async proxy
            tma_store_2d(&map_z, sZ + s * T::PBytes, j * kChunk, tile * 128);   // rows past M are clipped
            tma_store_commit();
            tma_store_wait_read();                           // shared memory has been read: the buffer may be rewritten
            mbar_arrive(&z_empty[s]);
          }
          if (p.p_out != nullptr) {
            if (p.debug & 4) mbar_wait(&p_full[s], ph); else mbar_wait_relaxed(&p_full[s], ph);
            if (!(p.debug & 2)) tma_store_2d(&map_p, sP + s * T::PBytes, j * kChunk, tile * 128);
            tma_store_commit();
            tma_store_wait_read();
            mbar_arrive(&p_empty[s]);
          }
        }
      }
      tma_store_wait_all();                                  // writes complete before the CTA exits
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ elementwise stage + final epilogue
    // Two groups of 8 warps: group s owns chunk parity s, i.e. accumulator acc_a[s] and operand buffer P[s].  The
    // groups run half a chunk period out of phase, so the latency one group exposes (TMEM load, proxy fence, barrier
    // round trips) is covered by the other group's GELU arithmetic on the same schedulers.
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int grp = (warp - kEpiWarp0) >> 3;               // chunk parity handled by this warp
    const int sub = ((warp - kEpiWarp0) >> 2) & 1;         // 32-column half of the chunk
    const int sub4 = (warp - kEpiWarp0) >> 2;              // 0..3: slice of the final epilogue
    const int r = q * 32 + lane;                           // row inside the tile
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const int pc = sub * 4;                                // first 16-byte piece of this thread in the 128-byte operand row
    const int s = grp;
    const uint32_t zrow_a = smem_u32(sZ) + (uint32_t)(s * T::PBytes + r * 128);   // this thread's row of the z chunk buffer
    const uint32_t prow_a = smem_u32(sP) + (uint32_t)(s * T::PBytes + r * 128);   // ... of the operand chunk buffer
    const uint32_t swz = (uint32_t)(r & 7);                 // SWIZZLE_128B: 16-byte piece index ^ (row & 7)
    const uint32_t bias1_a = smem_u32(sBias1), bias2_a = smem_u32(sBias2);
    const uint32_t scratch_a = smem_u32(bars + 31);
    constexpr int kPieces = (C / 16 + 3) / 4;                // 16-column pieces of the result per warp (strided by 4)
    uint4 res[kPieces][2];
    // residual rows of a finished tile: requested one chunk of arithmetic before they are consumed
    auto final_prefetch = [&](int tile_id) {
      const int orow = tile_id * 128 + r;
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int piece = sub4 + 4 * i;
        res[i][0] = res[i][1] = make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 0 && p.residual != nullptr && orow < p.M && piece < C / 16) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (int64_t)orow * C + piece * 16);
          res[i][0] = __ldg(rp); res[i][1] = __ldg(rp + 1);
        }
      }
    };
    // ---- final epilogue of a tile: acc_b[tile parity] (+ bias2 + residual) -> bf16 -> global
    auto final_epilogue = [&](int tt, int tile_id) {
      const int bb = tt & 1;
      const int orow = tile_id * 128 + r;
      const bool ok = orow < p.M;
      mbar_wait(&tb_full[bb], (uint32_t)(tt >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int piece = sub4 + 4 * i;
        if (piece < C / 16) {
          uint32_t v[16];
          tmem_ld16(tmem_b + t_lane + (uint32_t)(bb * C + piece * 16), v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]);
          if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const uint4 b = lds128(bias2_a + (uint32_t)((piece * 16 + k) * 4));
              f[k] += __uint_as_float(b.x); f[k + 1] += __uint_as_float(b.y);
              f[k + 2] += __uint_as_float(b.z); f[k + 3] += __uint_as_float(b.w);
            }
            if (p.residual != nullptr) {
              float rr[16];
              unpack8(res[i][0], rr); unpack8(res[i][1], rr + 8);
#pragma unroll
              for (int k = 0; k < 16; ++k) f[k] += rr[k];
            }
          }
          if (ok) {
            uint4* op = reinterpret_cast<uint4*>(p.out + (int64_t)orow * C + piece * 16);
            op[0] = pack8(f); op[1] = pack8(f + 8);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tb_empty[bb]);
    };
    int t = 0;
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++t) {
      for (int j = grp; j < T::NC; j += 2) {
        const uint32_t g = (uint32_t)(t * T::NC + j);
        const uint32_t ph = (g >> 1) & 1u;
        if (j == grp && t > 0) final_prefetch(tile - (int)gridDim.x);
        uint4 zq[4];
        if (MODE == 1) {                                    // saved pre-activation chunk, landed by TMA
          mbar_wait(&z_full[s], ph);
#pragma unroll
          for (int h = 0; h < 4; ++h) zq[h] = lds128(zrow_a + ((((uint32_t)(pc + h)) ^ swz) << 4));
          // the loads must have RETURNED before the buffer goes back to the TMA producer, which refills it at once
          // (mbarrier.arrive does not wait for outstanding loads: seen as z of chunk g+2 in the registers of chunk g)
          consume_loads(zq[0].x ^ zq[1].x ^ zq[2].x ^ zq[3].x, scratch_a);
          __syncwarp();
          if (lane == 0) mbar_arrive(&z_empty[s]);
        } else {
          mbar_wait(&z_empty[s], ph ^ 1u);                  // the TMA store of two chunks ago has read the buffer
        }
        // bias of this warp's first 8 columns: requested before the accumulator wait (the st.shared between the later
        // requests are volatile asm the compiler will not move a load across, so the prefetch is spelled out: without it
        // the FADD2 that adds the bias held 15 % of the warp-stall samples at C = 192, profiles/r02_mlp192_ncu.txt)
        // Backward only: measured, it takes 8 / 3 us off the backward at C = 96 / 192 and adds 1 / 5 us to the forward
        // (profiles/r02_ops_bench_mlp_bias_prefetch.txt).
        const uint32_t b1_a = bias1_a + (uint32_t)((j * kChunk + sub * 32) * 4);
        uint4 ba, bb;
        if (MODE == 1) { ba = lds128(b1_a); bb = lds128(b1_a + 16); }
        mbar_wait(&ta_full[s], ph);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld16(tmem_a + t_lane + (uint32_t)(s * kChunk + sub * 32), v);
        tmem_ld16(tmem_a + t_lane + (uint32_t)(s * kChunk + sub * 32 + 16), v + 16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ta_empty[s]);
        uint4 stored[4];                                    // the values the unfused path stores in bf16 (z / da)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[h * 8 + i]);
          stored[h] = pack8(f);
          if (MODE == 0) sts128(zrow_a + ((((uint32_t)(pc + h)) ^ swz) << 4), stored[h]);
        }
        if (MODE == 0) {                                    // z chunk complete in shared memory -> TMA store (warp 2 / 3)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&z_full[s]);
        }
        mbar_wait(&p_empty[s], ph ^ 1u);                    // GEMM-b (and the TMA store) of two chunks ago have read it
#pragma unroll
        for (int h = 0; h < 4; ++h) {                       // 8 columns at a time
          float f[8];
          unpack8(stored[h], f);
          if (MODE == 0) { ba = lds128(b1_a + (uint32_t)(h * 32)); bb = lds128(b1_a + (uint32_t)(h * 32 + 16)); }
          const float bias[8] = {__uint_as_float(ba.x), __uint_as_float(ba.y), __uint_as_float(ba.z), __uint_as_float(ba.w),
                                 __uint_as_float(bb.x), __uint_as_float(bb.y), __uint_as_float(bb.z), __uint_as_float(bb.w)};
          if (MODE == 1 && h < 3) {                         // the next 8 columns' bias, in flight during this GELU'
            ba = lds128(b1_a + (uint32_t)((h + 1) * 32));
            bb = lds128(b1_a + (uint32_t)((h + 1) * 32 + 16));
          }
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float2 g = b200at_gelu2(b200at_fadd2(make_float2(f[i], f[i + 1]), make_float2(bias[i], bias[i + 1])));
              f[i] = g.x; f[i + 1] = g.y;
            }
          } else {
            float zf[8];
            unpack8(zq[h], zf);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float2 g = b200at_fmul2(make_float2(f[i], f[i + 1]),
                                            b200at_gelu_grad2(b200at_fadd2(make_float2(zf[i], zf[i + 1]),
                                                                           make_float2(bias[i], bias[i + 1]))));
              f[i] = g.x; f[i + 1] = g.y;
            }
          }
          sts128(prow_a + ((((uint32_t)(pc + h)) ^ swz) << 4), pack8(f));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        // the previous tile's result leaves after this warp's first chunk of the next tile: its GEMM-b has long
        // retired by then (no wait), and the two groups drain at different times
        if (!(p.debug & 1) && j == grp && t > 0) final_epilogue(t - 1, tile - (int)gridDim.x);
      }
      if (p.debug & 1) final_epilogue(t, tile);
    }
    if (t > 0 && !(p.debug & 1)) {
      final_prefetch((int)blockIdx.x + (t - 1) * (int)gridDim.x);
      final_epilogue(t - 1, (int)blockIdx.x + (t - 1) * (int)gridDim.x);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)T::TmemCols)
                 : "memory");
  }
}

template <int C, int MODE>
int launch(const CUtensorMap& ma, const CUtensorMap& mwa, const CUtensorMap& mwb, const CUtensorMap& mz,
           const CUtensorMap& mp, const MlpParams& p, int grid, cudaStream_t s) {
  static b200at::SmemConfig configured;
  cudaError_t e = b200at::ensure_dynamic_smem(mlp_kernel<C, MODE>, (int)MlpCfg<C>::Smem, configured);
  if (e != cudaSuccess) return (int)e;
  mlp_kernel<C, MODE><<<grid, kThreads, MlpCfg<C>::Smem, s>>>(ma, mwa, mwb, mz, mp, p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int b200at_mlp_fused_supported(int64_t C) { return C == 96 || C == 128 || C == 192; }

extern "C" int b200at_mlp_fused(const void* a, const void* wa, const void* wb, const float* bias1, const float* bias2,
                                const void* residual, void* z, void* p_out, void* out, int64_t M, int64_t C,
                                int backward, void* stream) {
  if (M <= 0) return 0;
  if (!b200at_mlp_fused_supported(C) || !bias1 || !z) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(wa) | reinterpret_cast<uintptr_t>(wb) |
       reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(p_out) |
       reinterpret_cast<uintptr_t>(residual)) & 15)
    return (int)cudaErrorInvalidValue;
  MlpParams p;
  p.bias1 = bias1; p.bias2 = bias2; p.residual = (const bf16*)residual;
  p.z = (bf16*)z; p.p_out = (bf16*)p_out; p.out = (bf16*)out;
  p.M = (int)M; p.tiles_m = (int)((M + 127) / 128);
  { const char* e = getenv("B200AT_MLP_DEBUG"); p.debug = e ? atoi(e) : 0; }
  CUtensorMap ma, mwa, mwb, mz, mp;
  if (!make_map_kmajor(&ma, a, M, C, 128) || !make_map_kmajor(&mwa, wa, 4 * C, C, kChunk) ||
      !make_map_kmajor(&mwb, wb, C, 4 * C, (int)C) || !make_map_kmajor(&mz, z, M, 4 * C, 128) ||
      !make_map_kmajor(&mp, p_out ? p_out : z, M, 4 * C, 128))
    return (int)cudaErrorUnknown;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.tiles_m < sms ? p.tiles_m : sms;
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 128) return backward ? launch<128, 1>(ma, mwa, mwb, mz, mp, p, grid, s) : launch<128, 0>(ma, mwa, mwb, mz, mp, p, grid, s);
  if (C == 96) return backward ? launch<96, 1>(ma, mwa, mwb, mz, mp, p, grid, s) : launch<96, 0>(ma, mwa, mwb, mz, mp, p, grid, s);
  return backward ? launch<192, 1>(ma, mwa, mwb, mz, mp, p, grid, s) : launch<192, 0>(ma, mwa, mwb, mz, mp, p, grid, s);
}
