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
//   GEMM-a  acc_a[128 x 64]  = A_tile[128 x C] . Wa_chunk[64 x C]^T          (TMEM, double buffered)
//   f       16 epilogue warps: tcgen05.ld -> bias / GELU / GELU' (z chunk stored / loaded in bf16) -> bf16 ->
//           shared memory in the K-major SWIZZLE_128B operand layout (double buffered)
//   GEMM-b  acc_b[128 x C] += P_chunk[128 x 64] . Wb_chunk[C x 64]^T         (TMEM, lives for the whole tile)
// then acc_b (+ bias2 + residual) -> bf16 -> global.  Persistent, one CTA per SM, warp-specialised:
//   warp 0 TMA producer (A tile ring, weight-chunk ring), warp 1 MMA issuer (GEMM-a of chunk j+1 is issued before
//   GEMM-b of chunk j, so the tensor pipe works under the GELU of the previous chunk), warp 2 TMEM allocator,
//   warps 4-19 the elementwise stage + final epilogue.  mbarrier-only synchronisation.
// The hidden values are rounded to bf16 exactly where the unfused path stores them (z, da), so both paths agree
// to the accumulation order of the second GEMM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "b200at_gelu.cuh"
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
  static constexpr int AStages = C <= 96 ? 2 : 1;
  static constexpr int WStages = 2;
  static constexpr int ABytes = KB * 128 * 128;            // KB x [128 rows x 128 B]
  static constexpr int WaBytes = KB * kChunk * 128;        // KB x [64 rows x 128 B]
  static constexpr int WbBytes = C * 128;                  // [C rows x 128 B]
  static constexpr int WStageBytes = WaBytes + WbBytes;
  static constexpr int PBytes = 128 * 128;                 // [128 rows x 128 B]
  static constexpr int OffW = AStages * ABytes;
  static constexpr int OffP = OffW + WStages * WStageBytes;
  static constexpr int OffBias = OffP + 2 * PBytes;        // fp32 bias1[4C], bias2[C]
  static constexpr int OffBars = OffBias + 5 * C * 4;
  static constexpr int Smem = 1024 + OffBars + 256;
  static constexpr int TmemCols = (C + 2 * kChunk) <= 256 ? 256 : 512;
  static_assert(C % 16 == 0 && C <= 256, "GEMM-b is one UMMA of N = C");
  static_assert(Smem <= 227 * 1024, "shared memory budget");
  static_assert(C + 2 * kChunk <= 512, "TMEM budget");
};

struct MlpParams {
  const float* bias1;   // [4C]
  const float* bias2;   // [C] or null
  const bf16* residual; // [M][C] or null
  bf16* z;              // [M][4C]  forward: written (pre-activation without bias); backward: read
  bf16* p_out;          // [M][4C] or null: forward a = GELU(z + b1); backward dz
  bf16* out;            // [M][C]
  int M, tiles_m;
};

template <int C, int MODE>   // MODE 0 forward, 1 backward
__global__ void __launch_bounds__(kThreads, 1) mlp_kernel(const __grid_constant__ CUtensorMap map_a,
                                                          const __grid_constant__ CUtensorMap map_wa,
                                                          const __grid_constant__ CUtensorMap map_wb,
                                                          const MlpParams p) {
  typedef MlpCfg<C> T;
  extern __shared__ __align__(1024) uint8_t mlp_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mlp_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sW = smem + T::OffW;
  uint8_t* sP = smem + T::OffP;
  float* sBias1 = reinterpret_cast<float*>(smem + T::OffBias);
  float* sBias2 = sBias1 + 4 * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::OffBars);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* w_full = bars + 4;        // [2]
  uint64_t* w_empty = bars + 6;       // [2]
  uint64_t* ta_full = bars + 8;       // [2]  GEMM-a accumulator ready
  uint64_t* ta_empty = bars + 10;     // [2]
  uint64_t* p_full = bars + 12;       // [2]  operand chunk written
  uint64_t* p_empty = bars + 14;      // [2]
  uint64_t* tb_full = bars + 16;      // GEMM-b accumulator complete
  uint64_t* tb_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wa) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wb) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
      mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1);
      mbar_init(&ta_full[s], 1); mbar_init(&ta_empty[s], kEpiWarps);
      mbar_init(&p_full[s], kEpiWarps); mbar_init(&p_empty[s], 1);
    }
    mbar_init(tb_full, 1); mbar_init(tb_empty, kEpiWarps);
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
  const uint32_t tmem_b = tmem_base;                       // acc_b: columns [0, C)
  const uint32_t tmem_a = tmem_base + (uint32_t)C;         // acc_a[s]: columns C + 64 s

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t g = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++t) {
        const int as = t % T::AStages;
        const uint32_t aph = (uint32_t)(t / T::AStages) & 1u;
        mbar_wait(&a_empty[as], aph ^ 1u);
        mbar_expect_tx(&a_full[as], T::ABytes);
#pragma unroll
        for (int kb = 0; kb < T::KB; ++kb)
          tma_load_2d(&map_a, &a_full[as], sA + as * T::ABytes + kb * (128 * 128), kb * 64, tile * 128);
        for (int j = 0; j < T::NC; ++j, ++g) {
          const int ws = (int)(g & 1u);
          const uint32_t wph = (g >> 1) & 1u;
          mbar_wait(&w_empty[ws], wph ^ 1u);
          uint8_t* w = sW + ws * T::WStageBytes;
          mbar_expect_tx(&w_full[ws], T::WStageBytes);
#pragma unroll
          for (int kb = 0; kb < T::KB; ++kb)
            tma_load_2d(&map_wa, &w_full[ws], w + kb * (kChunk * 128), kb * 64, j * kChunk);
          tma_load_2d(&map_wb, &w_full[ws], w + T::WaBytes, j * kChunk, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_a = make_idesc(128, kChunk);
    const uint32_t idesc_b = make_idesc(128, C);
    uint32_t g = 0;
    int t = 0;
    auto gemm_a = [&](uint32_t gg, const uint8_t* a_tile) {
      const int s = (int)(gg & 1u);
      const uint32_t ph = (gg >> 1) & 1u;
      mbar_wait(&w_full[s], ph);                           // weight chunk gg landed
      mbar_wait(&ta_empty[s], ph ^ 1u);                    // the elementwise stage has drained acc_a[s]
      tc_fence_after();
      if (elect_one()) {
        const uint8_t* wa = sW + s * T::WStageBytes;
#pragma unroll
        for (int k = 0; k < T::KSteps; ++k) {
          const uint64_t da = make_desc(a_tile + (k >> 2) * (128 * 128)) + (uint64_t)(2 * (k & 3));
          const uint64_t db = make_desc(wa + (k >> 2) * (kChunk * 128)) + (uint64_t)(2 * (k & 3));
          umma(tmem_a + (uint32_t)(s * kChunk), da, db, idesc_a, k != 0);
        }
        umma_commit(&ta_full[s]);
      }
      __syncwarp();
    };
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++t) {
      const int as = t % T::AStages;
      const uint32_t aph = (uint32_t)(t / T::AStages) & 1u;
      const uint8_t* a_tile = sA + as * T::ABytes;
      mbar_wait(&a_full[as], aph);
      tc_fence_after();
      gemm_a(g, a_tile);
      for (int j = 0; j < T::NC; ++j, ++g) {
        if (j + 1 < T::NC) {
          gemm_a(g + 1, a_tile);
        } else {
          if (elect_one()) umma_commit(&a_empty[as]);      // every GEMM-a of this tile issued: A tile free when they retire
          __syncwarp();
        }
        const int s = (int)(g & 1u);
        const uint32_t ph = (g >> 1) & 1u;
        mbar_wait(&p_full[s], ph);                         // operand chunk written by the 16 epilogue warps
        if (j == 0) mbar_wait(tb_empty, (uint32_t)(t & 1) ^ 1u);   // previous tile's acc_b drained
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = make_desc(sP + s * T::PBytes);
          const uint64_t db = make_desc(sW + s * T::WStageBytes + T::WaBytes);
#pragma unroll
          for (int k = 0; k < kChunk / 16; ++k)
            umma(tmem_b, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_b, (j | k) != 0);
          umma_commit(&p_empty[s]);
          umma_commit(&w_empty[s]);
          if (j == T::NC - 1) umma_commit(tb_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ elementwise stage + final epilogue
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int sub = (warp - kEpiWarp0) >> 2;               // 16-column slice of the chunk
    const int r = q * 32 + lane;                           // row inside the tile
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const int pc = sub * 2;                                // first 16-byte piece of this thread in the 128-byte operand row
    uint32_t g = 0;
    int t = 0;
    for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x, ++t) {
      const int row = tile * 128 + r;
      const bool row_ok = row < p.M;
      bf16* zrow = p.z + (int64_t)row * (4 * C) + sub * 16;
      bf16* prow = p.p_out ? p.p_out + (int64_t)row * (4 * C) + sub * 16 : nullptr;
      for (int j = 0; j < T::NC; ++j, ++g) {
        const int s = (int)(g & 1u);
        const uint32_t ph = (g >> 1) & 1u;
        uint4 zlo = make_uint4(0u, 0u, 0u, 0u), zhi = zlo;
        if (MODE == 1 && row_ok) {                          // saved pre-activation: in flight while GEMM-a finishes
          zlo = __ldcs(reinterpret_cast<const uint4*>(zrow + j * kChunk));
          zhi = __ldcs(reinterpret_cast<const uint4*>(zrow + j * kChunk) + 1);
        }
        mbar_wait(&ta_full[s], ph);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tmem_a + t_lane + (uint32_t)(s * kChunk + sub * 16), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ta_empty[s]);
        float f[16], zf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        const float* b1 = sBias1 + j * kChunk + sub * 16;
        uint4 lo = pack8(f), hi = pack8(f + 8);             // the value the unfused path stores in bf16 (z / da)
        if (MODE == 0) {
          if (row_ok) {
            reinterpret_cast<uint4*>(zrow + j * kChunk)[0] = lo;
            reinterpret_cast<uint4*>(zrow + j * kChunk)[1] = hi;
          }
          unpack8(lo, f); unpack8(hi, f + 8);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b1 + i);
            f[i] = b200at_gelu(f[i] + b.x); f[i + 1] = b200at_gelu(f[i + 1] + b.y);
            f[i + 2] = b200at_gelu(f[i + 2] + b.z); f[i + 3] = b200at_gelu(f[i + 3] + b.w);
          }
        } else {
          unpack8(lo, f); unpack8(hi, f + 8);
          unpack8(zlo, zf); unpack8(zhi, zf + 8);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(b1 + i);
            f[i] *= b200at_gelu_grad(zf[i] + b.x); f[i + 1] *= b200at_gelu_grad(zf[i + 1] + b.y);
            f[i + 2] *= b200at_gelu_grad(zf[i + 2] + b.z); f[i + 3] *= b200at_gelu_grad(zf[i + 3] + b.w);
          }
        }
        lo = pack8(f); hi = pack8(f + 8);
        if (prow && row_ok) {
          reinterpret_cast<uint4*>(prow + j * kChunk)[0] = lo;
          reinterpret_cast<uint4*>(prow + j * kChunk)[1] = hi;
        }
        mbar_wait(&p_empty[s], ph ^ 1u);                    // GEMM-b that read this buffer two chunks ago has retired
        uint8_t* prow_s = sP + s * T::PBytes + r * 128;     // K-major SWIZZLE_128B: 16-byte piece index ^ (row & 7)
        *reinterpret_cast<uint4*>(prow_s + (((pc) ^ (r & 7)) << 4)) = lo;
        *reinterpret_cast<uint4*>(prow_s + (((pc + 1) ^ (r & 7)) << 4)) = hi;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the UMMA reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
      }
      // ---- final epilogue of the tile: acc_b (+ bias2 + residual) -> bf16 -> global
      mbar_wait(tb_full, (uint32_t)(t & 1));
      tc_fence_after();
      for (int piece = sub; piece < C / 16; piece += 4) {
        uint32_t v[16];
        tmem_ld16(tmem_b + t_lane + (uint32_t)(piece * 16), v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += sBias2[piece * 16 + i];
          if (p.residual && row_ok) {
            float rr[16];
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (int64_t)row * C + piece * 16);
            unpack8(__ldg(rp), rr); unpack8(__ldg(rp + 1), rr + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += rr[i];
          }
        }
        if (row_ok) {
          uint4* op = reinterpret_cast<uint4*>(p.out + (int64_t)row * C + piece * 16);
          op[0] = pack8(f); op[1] = pack8(f + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tb_empty);
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
int launch(const CUtensorMap& ma, const CUtensorMap& mwa, const CUtensorMap& mwb, const MlpParams& p, int grid,
           cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpCfg<C>::Smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  mlp_kernel<C, MODE><<<grid, kThreads, MlpCfg<C>::Smem, s>>>(ma, mwa, mwb, p);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int b200at_mlp_fused_supported(int64_t C) { return C == 96 || C == 192; }

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
  CUtensorMap ma, mwa, mwb;
  if (!make_map_kmajor(&ma, a, M, C, 128) || !make_map_kmajor(&mwa, wa, 4 * C, C, kChunk) ||
      !make_map_kmajor(&mwb, wb, C, 4 * C, (int)C))
    return (int)cudaErrorUnknown;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.tiles_m < sms ? p.tiles_m : sms;
  cudaStream_t s = (cudaStream_t)stream;
  if (C == 96) return backward ? launch<96, 1>(ma, mwa, mwb, p, grid, s) : launch<96, 0>(ma, mwa, mwb, p, grid, s);
  return backward ? launch<192, 1>(ma, mwa, mwb, p, grid, s) : launch<192, 0>(ma, mwa, mwb, p, grid, s);
}
