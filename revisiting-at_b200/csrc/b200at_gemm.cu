// K10: tcgen05 / TMEM / TMA GEMM for the dense pwconv (MLP) layers of the ConvNeXt block, with the
// block's elementwise tail fused into the epilogue (include/b200at_model.h: b200at_gemm_bf16).
//
//   C[M,N] = epilogue( A[M,K] · B[N,K]^T )        A, B bf16 K-major (row-major, K contiguous), fp32 accumulate
//
// Persistent, warp-specialised, one CTA per SM (cta_group::1):
//   warp 0   TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 64 bf16 along K) -> kStages-deep smem ring
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma.kind::f16 (UMMA 128 x BLOCK_N x 16), accumulators in
//                           TMEM (2 x BLOCK_N fp32 columns, double buffered across output tiles)
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     tcgen05.ld 32x32b -> registers -> bias / GELU / GELU' / residual -> bf16 -> XOR-swizzled
//                           per-warp staging tile in shared memory -> 128-byte-line coalesced global stores
//                           (two warps per TMEM lane quarter, alternating 64-column strips)
// Synchronisation is mbarrier only (full/empty per smem stage, full/empty per TMEM accumulator).
// Out-of-range rows / K tails are zero-filled by TMA; the epilogue masks its loads/stores on M.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "b200at_gelu.cuh"
#include "b200at_launch.cuh"
#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kEpilogueWarp0 = 4;      // warps 0..3: TMA / MMA / TMEM-alloc / idle
// Epilogue warps EW (template parameter): 8 (warps 4..11, 4-stage operand ring) for the light epilogues; 16 (warps 4..19,
// 3-stage ring) for the GELU / GELU' epilogues, whose ~30 instructions per element on 8 warps took twice the tile's MMA
// time (profiles/r01_gemm_bench_v2.txt) -- with 16 warps (4 per TMEM lane quarter, one 64-column strip each at
// BLOCK_N = 256) the epilogue of tile i hides under the MMAs of tile i+1 (double-buffered accumulator).
constexpr int kStripCols = 64;         // columns per epilogue strip (128 B of bf16 per row)
constexpr int kStageTileBytes = 32 * kStripCols * 2;   // per-warp staging tile: 32 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major: 1), [32,46) SBO >> 4 (1024 B between
//   8-row groups), [46,48) version = 1 (Blackwell), [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(const void* smem) {
  const uint32_t addr = smem_u32(smem);
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 32 rows x 128 B staging tile, 16-byte pieces XOR-swizzled by (row & 7): conflict-free both for the
// row-per-lane writes and for the line-per-8-lanes reads
__device__ __forceinline__ uint4* stage_piece(uint8_t* tile, int row, int piece) {
  return reinterpret_cast<uint4*>(tile + row * (kStripCols * 2) + ((piece ^ (row & 7)) << 4));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct GemmParams {
  bf16* c;            // [M][N] output
  bf16* c2;           // optional second output (EPI_BIAS_GELU: the pre-activation z = acc), or null
  const bf16* aux;    // EPI_RESIDUAL: residual [M][N];  EPI_GELU_GRAD: saved pre-activation z [M][N]
  const float* bias;  // [N] or null
  float* colsum;      // EPI_GELU_GRAD only, or null: [N] fp32, += column sums of the (bf16-rounded) output -- the pwconv1 bias gradient
  int M, N, K;
  int block_n, tiles_m, tiles_n;
  int tile_rows;      // output rows an M tile really holds (128; the implicit-GEMM convolution: whole output image rows)
  // implicit GEMM of a 3x3 stride-2 pad-1 convolution on NHWC input (b200at_conv3x3s2_fwd): k-block kb = tap (kh, kw), the A
  // tile of tile tm is the TMA box {64 channels, conv_ow pixels at stride 2, conv_rows rows at stride 2} at (kw - 1, 2 oh0 + kh - 1)
  int conv;           // 0: plain GEMM
  int conv_rows, conv_tiles_per_img;
  uint32_t conv_a_bytes;
};

// ------------------------------------------------------------------------------------------------ GELU / GELU' epilogues
// The first version of these two epilogues ran the generic strip code above with 64 columns in registers: 34 instructions
// per element (half of them address arithmetic, re-materialised lane ids and one scalar LDG per bias value) and, for GELU',
// a row-per-lane read of the saved pre-activation (32 lines touched per instruction, 16 useful bytes each) -- 69 / 87 us at
// 25088 x 1536 x 384 against 32 + 32 / 33 + 44 us for the plain GEMM followed by the elementwise kernel
// (profiles/r02_gemm_epilogue_ncu.txt).  This version walks a warp's columns 32 at a time:
//   aux (GELU': the saved z) comes in as 64-byte row segments, 8 rows per instruction, through the warp's staging tile, and
//       the loads of the next pass are in flight during the arithmetic of the current one (the first pass of a tile is
//       requested before the accumulator barrier);
//   bias sits in shared memory (one broadcast LDS.128 per 4 columns);  the arithmetic is the packed fp32x2 GELU;
//   every shared-memory access is an explicit ld/st.shared on a 32-bit address computed once per warp.
constexpr int kPassCols = 32;
constexpr int kFusedTileBytes = 32 * kPassCols * 2;    // per-warp staging tile: 32 rows x 64 B, pieces XOR-swizzled by (row >> 1) & 3
constexpr int kMaxBiasCols = 8192;

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float2 v) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <int EPI, int EW>
__device__ __forceinline__ void fused_epilogue(const bf16* __restrict__ aux, bf16* __restrict__ c_out, bf16* __restrict__ c2_out,
                                               int M, int N, int block_n, int tiles_n, int num_tiles, uint32_t tmem_base,
                                               uint64_t* tfull, uint64_t* tempty, uint32_t tile_a, uint32_t bias_a, int warp,
                                               int lane, bool colsum) {
  const int q = warp & 3, sub = (warp - kEpilogueWarp0) >> 2;        // TMEM lane quarter; which 64-column group of each 256
  const uint32_t row_a = tile_a + (uint32_t)lane * 64u;               // row-per-lane view
  const uint32_t rsw = (uint32_t)((lane >> 1) & 3);
  const int lr = lane >> 2, lp = lane & 3;                            // line view: 4 lanes per 64-byte row segment, 8 rows
  const uint32_t line_a = tile_a + (uint32_t)lr * 64u + (((uint32_t)lp ^ (uint32_t)((lr >> 1) & 3)) << 4);   // + 512 k
  int it = 0;
  for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, ++it) {
    const int tm = tile_id / tiles_n, tn = tile_id - tm * tiles_n;
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    const int row0 = tm * kBlockM + q * 32;
    const int ncol0 = tn * block_n;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * block_n);
    const int64_t line_off = (int64_t)(row0 + lr) * N + ncol0 + lp * 8;
    const int rows_left = M - row0 - lr;                              // line k is inside the matrix when 8 k < rows_left
    auto pass_col = [&](int e) { return sub * 64 + (e >> 1) * (EW / 4) * 64 + (e & 1) * kPassCols; };
    auto pass_width = [&](int e) {
      const int c = pass_col(e);
      int w = block_n - c;
      const int wn = N - ncol0 - c;
      w = w < wn ? w : wn;
      return w < kPassCols ? w : kPassCols;
    };
    uint4 zpre[4];
    auto load_aux = [&](int e) {
      const int c = pass_col(e), w = pass_width(e);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        zpre[k] = make_uint4(0u, 0u, 0u, 0u);
        if (lp * 8 < w && 8 * k < rows_left)
          zpre[k] = __ldg(reinterpret_cast<const uint4*>(aux + line_off + (int64_t)(8 * k) * N + c));
      }
    };
    auto store_tile = [&](bf16* __restrict__ dst, int c, int w, const uint4* o) {      // rows per lane in, 64-byte segments out
#pragma unroll
      for (int j = 0; j < 4; ++j) sts128(row_a + (((uint32_t)j ^ rsw) << 4), o[j]);
      __syncwarp();
      float2 cs[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 t = lds128(line_a + 512u * k);
        if (lp * 8 < w && 8 * k < rows_left) {
          *reinterpret_cast<uint4*>(dst + line_off + (int64_t)(8 * k) * N + c) = t;
          if (EPI == B200AT_EPI_GELU_GRAD && colsum) {          // what the weight-gradient GEMM will see: the rounded values
            cs[0] = b200at_fadd2(cs[0], bf2_to_f2(t.x)); cs[1] = b200at_fadd2(cs[1], bf2_to_f2(t.y));
            cs[2] = b200at_fadd2(cs[2], bf2_to_f2(t.z)); cs[3] = b200at_fadd2(cs[3], bf2_to_f2(t.w));
          }
        }
      }
      if (EPI == B200AT_EPI_GELU_GRAD && colsum) {              // 8 rows-of-lanes -> one: lanes 0..3 hold 8 column sums each
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            cs[i].x += __shfl_xor_sync(0xffffffffu, cs[i].x, off);
            cs[i].y += __shfl_xor_sync(0xffffffffu, cs[i].y, off);
          }
        }
        if (lr == 0 && lp * 8 < w) {
          const uint32_t a0 = bias_a + (uint32_t)((ncol0 + c + lp * 8) * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a0 + 8u * i), "f"(cs[i].x) : "memory");
            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a0 + 8u * i + 4u), "f"(cs[i].y) : "memory");
          }
        }
      }
      __syncwarp();
    };
    if (EPI == B200AT_EPI_GELU_GRAD || EPI == B200AT_EPI_RESIDUAL) load_aux(0);
    mbar_wait(&tfull[acc], acc_phase);
    tc_fence_after();
    for (int e = 0; pass_col(e) < block_n; ++e) {
      const int c = pass_col(e), w = pass_width(e);
      if (w <= 0) break;
      uint32_t v[32];
      tmem_ld16(t_row + (uint32_t)c, v);
      if (w > 16) tmem_ld16(t_row + (uint32_t)(c + 16), v + 16);
      uint4 zrow[4];
      if (EPI == B200AT_EPI_GELU_GRAD || EPI == B200AT_EPI_RESIDUAL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sts128(line_a + 512u * k, zpre[k]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) zrow[j] = lds128(row_a + (((uint32_t)j ^ rsw) << 4));
        __syncwarp();
        if (pass_col(e + 1) < block_n) load_aux(e + 1);
      }
      tmem_ld_wait();
      uint4 o[4];
      if (EPI == B200AT_EPI_GELU_GRAD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t zw[4] = {zrow[j].x, zrow[j].y, zrow[j].z, zrow[j].w};
          uint32_t ow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 g = make_float2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1]));
            ow[i] = f2_to_bf2(b200at_fmul2(g, b200at_gelu_grad2(bf2_to_f2(zw[i]))));
          }
          o[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        store_tile(c_out, c, w, o);
      } else if (EPI == B200AT_EPI_RESIDUAL) {            // (acc + bias) + residual, the order of the generic epilogue
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t ba = bias_a + (uint32_t)((ncol0 + c + 8 * j) * 4);
          const uint4 b0 = lds128(ba), b1 = lds128(ba + 16);
          const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const uint32_t zw[4] = {zrow[j].x, zrow[j].y, zrow[j].z, zrow[j].w};
          uint32_t ow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = b200at_fadd2(make_float2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1])),
                                          make_float2(__uint_as_float(bw[2 * i]), __uint_as_float(bw[2 * i + 1])));
            ow[i] = f2_to_bf2(b200at_fadd2(t, bf2_to_f2(zw[i])));
          }
          o[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        store_tile(c_out, c, w, o);
      } else {
        float2 f[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t ba = bias_a + (uint32_t)((ncol0 + c + 8 * j) * 4);
          const uint4 b0 = lds128(ba), b1 = lds128(ba + 16);
          const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t ow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[4 * j + i] = b200at_fadd2(make_float2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1])),
                                        make_float2(__uint_as_float(bw[2 * i]), __uint_as_float(bw[2 * i + 1])));
            ow[i] = f2_to_bf2(f[4 * j + i]);
          }
          o[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        if (c2_out != nullptr) store_tile(c2_out, c, w, o);                 // the pre-activation (bias included)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t ow[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ow[i] = f2_to_bf2(b200at_gelu2(f[4 * j + i]));
          o[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        store_tile(c_out, c, w, o);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[acc]);
  }
}

template <int EPI, int EW, bool CONV = false>
__global__ void __launch_bounds__(32 * (kEpilogueWarp0 + EW), 1) gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B block_n*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = kBlockM * kBlockK * 2;
  const uint32_t b_bytes = (uint32_t)p.block_n * kBlockK * 2;
  // CONV: the whole weight matrix (9 k-blocks) stays in shared memory behind the A ring -- re-fetching 110 KB of weights for
  // every 112-pixel tile made the kernel L2-bandwidth-bound (839 MB of L2 reads, 91 us at 112 x 112 x 48 -> 96, batch 128)
  const uint32_t stage_bytes = CONV ? a_bytes : a_bytes + b_bytes;
  constexpr int kStages = CONV ? 5 : (EW > 8 ? 3 : 4);
  uint8_t* wres = smem + kStages * stage_bytes;                        // CONV: [num_k][b_bytes]
  const uint32_t wres_bytes = CONV ? (uint32_t)((p.K + kBlockK - 1) / kBlockK) * b_bytes : 0u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes + wres_bytes);
  uint64_t* full = bars;                   // [kStages]
  uint64_t* empty = bars + kStages;        // [kStages]
  uint64_t* tfull = bars + 2 * kStages;    // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* wfull = bars + 2 * kStages + 5;   // CONV: the resident weights have landed
  uint8_t* staging = smem + kStages * stage_bytes + wres_bytes + 256;   // [EW][kStageTileBytes]  (EW == 16: [EW][kFusedTileBytes], then bias)
  float* sbias = reinterpret_cast<float*>(staging + EW * kFusedTileBytes);   // fused epilogues only: [tiles_n * block_n]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (p.K + kBlockK - 1) / kBlockK;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const uint32_t tmem_cols = p.block_n * 2 <= 32 ? 32 : (p.block_n * 2 <= 64 ? 64 : (p.block_n * 2 <= 128 ? 128 : (p.block_n * 2 <= 256 ? 256 : 512)));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EW); }
    if (CONV) mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  constexpr bool kFused = EW == 16 || (EPI == B200AT_EPI_RESIDUAL && !CONV);   // epilogues that run fused_epilogue
  if (kFused && (EPI == B200AT_EPI_BIAS_GELU || EPI == B200AT_EPI_RESIDUAL)) {
    const int padded = p.tiles_n * p.block_n;
    for (int i = threadIdx.x; i < padded; i += blockDim.x) sbias[i] = (p.bias != nullptr && i < p.N) ? p.bias[i] : 0.0f;
  }
  if (EW == 16 && EPI == B200AT_EPI_GELU_GRAD && p.colsum != nullptr) {     // the same region: this CTA's column sums
    const int padded = p.tiles_n * p.block_n;
    for (int i = threadIdx.x; i < padded; i += blockDim.x) sbias[i] = 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (CONV) {                                  // tiles_n == 1: one weight matrix for every tile
        mbar_expect_tx(wfull, wres_bytes);
        for (int kb = 0; kb < num_k; ++kb) tma_load_2d(&map_b, wfull, wres + kb * b_bytes, kb * kBlockK, 0);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          if (CONV) {
            const int img = tm / p.conv_tiles_per_img, oh0 = (tm - img * p.conv_tiles_per_img) * p.conv_rows;
            const int kh = kb / 3, kw = kb - 3 * kh;
            mbar_expect_tx(&full[stage], p.conv_a_bytes);
            tma_load_4d(&map_a, &full[stage], sa, 0, kw - 1, 2 * oh0 + kh - 1, img);
          } else {
            mbar_expect_tx(&full[stage], stage_bytes);
            tma_load_2d(&map_a, &full[stage], sa, kb * kBlockK, tm * kBlockM);
            tma_load_2d(&map_b, &full[stage], sa + a_bytes, kb * kBlockK, tn * p.block_n);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(kBlockM, p.block_n);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (CONV) mbar_wait(wfull, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* sa = smem + stage * stage_bytes;
          const uint64_t da = make_desc(sa), db = make_desc(CONV ? wres + kb * b_bytes : sa + a_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 B along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            umma(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);                 // frees the smem stage when these MMAs retire
          if (kb == num_k - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kEpilogueWarp0 && kFused) {
    fused_epilogue<EPI, EW>(p.aux, p.c, p.c2, p.M, p.N, p.block_n, p.tiles_n, num_tiles, tmem_base, tfull, tempty,
                        smem_u32(staging) + (uint32_t)((warp - kEpilogueWarp0) * kFusedTileBytes), smem_u32(sbias), warp, lane,
                        p.colsum != nullptr);
  } else if (warp >= kEpilogueWarp0) {
    // ------------------------------------------------------------------ epilogue (TMEM -> regs -> smem -> global)
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int half = (warp - kEpilogueWarp0) >> 2;     // which of the EW / 4 warps of that quarter
    uint8_t* tile = staging + (warp - kEpilogueWarp0) * kStageTileBytes;
    int it = 0;
    for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, ++it) {
      const int tm = tile_id / p.tiles_n, tn = tile_id % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row0 = tm * p.tile_rows + q * 32;       // first row of this warp's quarter
      const int row = row0 + lane;
      const bool row_ok = q * 32 + lane < p.tile_rows && row < p.M;
      const int64_t row_off = (int64_t)row * p.N;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
      for (int s0 = half * kStripCols; s0 < p.block_n; s0 += (EW / 4) * kStripCols) {
        const int width = (p.block_n - s0) < kStripCols ? (p.block_n - s0) : kStripCols;   // multiple of 16
        const int col0 = tn * p.block_n + s0;
        float f[kStripCols];
#pragma unroll
        for (int c = 0; c < kStripCols; c += 16) {
          if (c < width) {
            uint32_t v[16];
            tmem_ld16(t_row + (uint32_t)(s0 + c), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[c + i] = __uint_as_float(v[i]);
          }
        }
        tmem_ld_wait();
        const bool live = row_ok;
#pragma unroll
        for (int c = 0; c < kStripCols; c += 16) {
          if (c < width && col0 + c < p.N) {
            float* g = f + c;
            if (EPI == B200AT_EPI_BIAS || EPI == B200AT_EPI_BIAS_GELU || EPI == B200AT_EPI_RESIDUAL) {
              if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i) g[i] += __ldg(p.bias + col0 + c + i);
              }
            }
            if (EPI == B200AT_EPI_RESIDUAL && live) {
              float r[16];
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.aux + row_off + col0 + c)), r);
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.aux + row_off + col0 + c + 8)), r + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) g[i] += r[i];
            }
            if (EPI == B200AT_EPI_GELU_GRAD && live) {
              float z[16];
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.aux + row_off + col0 + c)), z);
              unpack8(__ldg(reinterpret_cast<const uint4*>(p.aux + row_off + col0 + c + 8)), z + 8);
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const float2 r = b200at_fmul2(make_float2(g[i], g[i + 1]), b200at_gelu_grad2(make_float2(z[i], z[i + 1])));
                g[i] = r.x; g[i + 1] = r.y;
              }
            }
          }
        }
        // one or two outputs leave through the staging tile: rows per lane in, 128-byte lines out
        const int n_out = (EPI == B200AT_EPI_BIAS_GELU && p.c2 != nullptr) ? 2 : 1;
        for (int o = 0; o < n_out; ++o) {
          bf16* dst = (n_out == 2 && o == 0) ? p.c2 : p.c;
          if (EPI == B200AT_EPI_BIAS_GELU && o == n_out - 1) {
#pragma unroll
            for (int i = 0; i < kStripCols; i += 2) {
              const float2 r = b200at_gelu2(make_float2(f[i], f[i + 1]));
              f[i] = r.x; f[i + 1] = r.y;
            }
          }
#pragma unroll
          for (int c = 0; c < kStripCols; c += 8) {
            if (c < width) *stage_piece(tile, lane, c >> 3) = pack8(f + c);
          }
          __syncwarp();
          const int piece = lane & 7;                   // 16-byte piece of the 128-byte row segment
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = (lane >> 3) + 4 * j;          // 4 rows per store instruction
            const int col = col0 + piece * 8;
            if (piece * 8 < width && q * 32 + r < p.tile_rows && row0 + r < p.M && col < p.N)
              *reinterpret_cast<uint4*>(dst + (int64_t)(row0 + r) * p.N + col) = *stage_piece(tile, r, piece);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
  if (EW == 16 && EPI == B200AT_EPI_GELU_GRAD && p.colsum != nullptr) {     // one global atomic per column and CTA
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      const float v = sbias[i];
      if (v != 0.0f) atomicAdd(p.colsum + i, v);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [rows][K] bf16 row-major tensor, boxes of (64 along K) x box_rows, 128 B swizzle, zero fill outside
bool make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_block_n(int64_t N) {
  // largest multiple of 16 that divides N and is <= 256 (UMMA N limit at M = 128); else 128 with a masked tail
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 128;
}

int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int grid, cudaStream_t s) {
  constexpr int kStages = 5;
  const size_t smem = 1024 + (size_t)kStages * (kBlockM * kBlockK * 2) + (size_t)((p.K + kBlockK - 1) / kBlockK) * p.block_n * kBlockK * 2 +
                      256 + (size_t)8 * kStageTileBytes;
  if (smem > 227 * 1024) return -1;
  static b200at::SmemConfig configured;
  cudaError_t e = b200at::ensure_dynamic_smem(gemm_kernel<B200AT_EPI_NONE, 8, true>, 227 * 1024, configured);
  if (e != cudaSuccess) return (int)e;
  gemm_kernel<B200AT_EPI_NONE, 8, true><<<grid, 32 * (kEpilogueWarp0 + 8), smem, s>>>(ma, mb, p);
  return (int)cudaGetLastError();
}

template <int EPI>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int grid, cudaStream_t s) {
  constexpr int EW = (EPI == B200AT_EPI_BIAS_GELU || EPI == B200AT_EPI_GELU_GRAD) ? 16 : 8;
  constexpr int kStages = EW > 8 ? 3 : 4;
  constexpr bool kFused = EW == 16 || EPI == B200AT_EPI_RESIDUAL;
  const size_t smem = 1024 + (size_t)kStages * (kBlockM * kBlockK * 2 + (size_t)p.block_n * kBlockK * 2) + 256 +
                      (kFused ? (size_t)EW * kFusedTileBytes + sizeof(float) * (size_t)p.tiles_n * p.block_n
                              : (size_t)EW * kStageTileBytes);
  if (kFused && (int64_t)p.tiles_n * p.block_n > kMaxBiasCols) return (int)cudaErrorInvalidValue;
  static b200at::SmemConfig configured;
  cudaError_t e = b200at::ensure_dynamic_smem(gemm_kernel<EPI, EW>, 227 * 1024, configured);
  if (e != cudaSuccess) return (int)e;
  gemm_kernel<EPI, EW><<<grid, 32 * (kEpilogueWarp0 + EW), smem, s>>>(ma, mb, p);
  return (int)cudaGetLastError();
}

}  // namespace

static int gemm_bf16_impl(const void* a, const void* b, void* c, void* c2, const void* aux, const float* bias, float* colsum,
                         int64_t M, int64_t N, int64_t K, int epilogue, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (N % 16 || K % 8) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15)
    return (int)cudaErrorInvalidValue;
  GemmParams p;
  p.c = (bf16*)c; p.c2 = (bf16*)c2; p.aux = (const bf16*)aux; p.bias = bias; p.colsum = colsum;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.block_n = pick_block_n(N);
  p.tiles_m = (int)((M + kBlockM - 1) / kBlockM);
  p.tiles_n = (int)((N + p.block_n - 1) / p.block_n);
  p.tile_rows = kBlockM;
  p.conv = 0; p.conv_rows = 0; p.conv_tiles_per_img = 1; p.conv_a_bytes = 0;
  CUtensorMap ma, mb;
  if (!make_map(&ma, a, M, K, kBlockM) || !make_map(&mb, b, N, K, p.block_n)) return (int)cudaErrorUnknown;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sms ? tiles : sms;
  cudaStream_t s = (cudaStream_t)stream;
  switch (epilogue) {
    case B200AT_EPI_NONE: return launch<B200AT_EPI_NONE>(ma, mb, p, grid, s);
    case B200AT_EPI_BIAS: return launch<B200AT_EPI_BIAS>(ma, mb, p, grid, s);
    case B200AT_EPI_BIAS_GELU: return launch<B200AT_EPI_BIAS_GELU>(ma, mb, p, grid, s);
    case B200AT_EPI_RESIDUAL: return launch<B200AT_EPI_RESIDUAL>(ma, mb, p, grid, s);
    case B200AT_EPI_GELU_GRAD: return launch<B200AT_EPI_GELU_GRAD>(ma, mb, p, grid, s);
    default: return (int)cudaErrorInvalidValue;
  }
}

extern "C" int b200at_gemm_bf16(const void* a, const void* b, void* c, void* c2, const void* aux, const float* bias,
                                int64_t M, int64_t N, int64_t K, int epilogue, void* stream) {
  return gemm_bf16_impl(a, b, c, c2, aux, bias, nullptr, M, N, K, epilogue, stream);
}

// C = (A B^T) * GELU'(aux) as b200at_gemm_bf16(..., B200AT_EPI_GELU_GRAD), and colsum[n] += sum_m C[m][n] of the bf16-rounded
// result: the pwconv1 bias gradient without a separate pass over dz (b200at_colsum_bf16).  colsum: fp32 [N], accumulated into.
extern "C" int b200at_gemm_gelu_grad_colsum(const void* a, const void* b, void* c, const void* aux, float* colsum, int64_t M,
                                            int64_t N, int64_t K, void* stream) {
  if (colsum == nullptr) return (int)cudaErrorInvalidValue;
  return gemm_bf16_impl(a, b, c, nullptr, aux, nullptr, colsum, M, N, K, B200AT_EPI_GELU_GRAD, stream);
}

// out[b][oh][ow][co] = sum_{kh,kw,ci} x[b][2 oh + kh - 1][2 ow + kw - 1][ci] w[co][ci][kh][kw]   (Conv2d(k=3, s=2, p=1), no bias)
// as an implicit GEMM on the kernel above: M = output pixels (tiles of whole output rows), N = Cout, K = 9 taps x 64 (the
// channels of a tap padded to one SWIZZLE_128B k-block; TMA zero-fills channels >= Cin and the padding ring).
// wk: [Cout][9 * 64] bf16, wk[co][(kh * 3 + kw) * 64 + ci] = w[co][ci][kh][kw], zero for ci >= Cin.  Returns -1 for shapes it
// does not take (the caller keeps the library convolution for those).
extern "C" int b200at_conv3x3s2_fwd(const void* x, const void* wk, void* y, int64_t B, int64_t H, int64_t W, int64_t Cin,
                                    int64_t Cout, void* stream) {
  if (B <= 0) return 0;
  if (H % 2 || W % 2 || Cin % 8 || Cin > 64 || Cout % 16 || Cout > 256) return -1;
  const int64_t OH = H / 2, OW = W / 2;
  if (OW > 128 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (reinterpret_cast<uintptr_t>(wk) & 15))
    return -1;
  int rows = (int)(kBlockM / OW);
  while (rows > 1 && OH % rows) --rows;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorUnknown;
  GemmParams p;
  p.c = (bf16*)y; p.c2 = nullptr; p.aux = nullptr; p.bias = nullptr; p.colsum = nullptr;
  p.M = (int)(B * OH * OW); p.N = (int)Cout; p.K = 9 * kBlockK;
  p.block_n = pick_block_n(Cout);
  p.tile_rows = rows * (int)OW;
  p.tiles_m = (int)(B * (OH / rows));
  p.tiles_n = (int)((Cout + p.block_n - 1) / p.block_n);
  p.conv = 1; p.conv_rows = rows; p.conv_tiles_per_img = (int)(OH / rows);
  p.conv_a_bytes = (uint32_t)(kBlockK * 2 * OW * rows);
  if ((int64_t)B * OH * OW > 0x7fffffff) return -1;
  CUtensorMap ma, mb;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(2 * OW), (cuuint32_t)(2 * rows), 1};
    const cuuint32_t estr[4] = {1, 2, 2, 1};
    if (fn(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorUnknown;
  }
  if (!make_map(&mb, wk, Cout, 9 * kBlockK, p.block_n)) return (int)cudaErrorUnknown;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.tiles_m * p.tiles_n;
  if (p.tiles_n != 1) return -1;
  return launch_conv(ma, mb, p, tiles < sms ? tiles : sms, (cudaStream_t)stream);
}
