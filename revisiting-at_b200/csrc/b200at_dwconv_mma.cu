// K7 (second design): depthwise 7x7, stride 1, pad 3, NHWC bf16 -- forward and input gradient on the tensor cores.
// Reference op: /root/reference/models/convnext.py:28,39 (`self.dwconv`), its dgrad = same correlation with flipped taps.
//
// Why tensor cores for a depthwise conv.  The fp32-FMA kernel (b200at_convnext.cu, dwconv7_kernel) is bound by the FMA
// pipe and its shared-memory operand path: 12-15 TFMA/s of the chip's 37, i.e. 0.19 of the HBM roofline
// (profiles/r01_ncu_dwconv_stem0_v11_summary.txt).  A depthwise conv has no channel contraction, but along one image
// row it IS a matrix product with a banded Toeplitz matrix:
//
//     out[h][w] = sum_dh sum_k in[h + dh][k] * T_dh[k][w],        T_dh[k][w] = taps[dh][k - w]  (0 <= k - w < 7)
//
// per channel.  With M = 16 output rows (any 16 rows of the tile: ldmatrix takes one address per row), K = a 16-wide window
// of input columns and N = 8 output columns, one mma.sync.m16n8k16 does 56 useful of its 128 MACs per row (14 of the 16
// window columns, 7 taps of each 8 outputs) and the Toeplitz B fragment depends only on (channel, dh) -- shift
// invariance -- so it lives in 14 registers for the whole tile.  Measured HMMA issue rate on B200: 826-880 mma/us/SM
// (profiles/r02_hmma_rate.txt) = 110 useful TMAC/s for this mapping, against 15 TMAC/s of the FMA kernel.
//
// The channel is a batch dimension of these products, while memory is channel-innermost, so a tile is transposed on chip
// into [channel][row][column] planes of bf16 with the column contiguous (what ldmatrix wants; row stride an odd multiple of
// 16 bytes: conflict-free ldmatrix and fragment stores).  What bounds the kernel is the load/store unit's wavefront rate
// (one 128-byte shared-memory / L1 wavefront per cycle per SM; v1-v2 of this file, profiles/r02_dwconv_mma_ncu_v*.txt:
// per-lane LDG.128 / STG.128 of 16-channel pixels cost one wavefront per pixel and were 40 % of all wavefronts), so
// global memory is touched only by TMA and both transpositions are 8x8 matrix moves:
//   TMA in   one cp.async.bulk.tensor.4d per tile: box (16 channels, 8 NT + 8 columns from w = -4, TH + 6 rows from
//            h0 - 3, NB images), zero fill outside the image = the conv padding; SWIZZLE_32B so that 8 consecutive pixels
//            of one channel octet are 8 distinct 16-byte bank groups.  Issued for tile i+1 as soon as phase 1 of tile i
//            has drained the staging buffer: the load runs under phases 2-3.
//   phase 1  staging -> planes: ldmatrix.trans.x4 (16 pixels x 16 channels; lane (g, q) receives channel g / 8 + g of
//            pixel pairs 2q, 8 + 2q as packed words) + 4 STS.32 into the planes (plane stride = 4 mod 8 words: conflict-free).
//   phase 2  warp w owns channel w: Toeplitz fragments from the staged taps, then per 16-row M tile and tap row dh
//            ceil((NT+1)/2) ldmatrix.x4 + NT mma (window j uses 8-column chunks j, j+1: each chunk is loaded once and used
//            by two windows), fp32 accumulators (+ bias), results packed to bf16 IN PLACE over input rows that later
//            M tiles no longer need.
//   phase 3  planes -> output staging: 4 LDS.32 (+ the `add` tile, TMA-loaded into the same staging buffer during
//            phase 2: the residual-gradient join of the block's backward, models/convnext.py:49) + stmatrix.trans.x4.
//   TMA out  one cp.async.bulk.tensor.4d store per tile (columns / rows / images outside the tensor are clipped).
// One persistent CTA of 16 warps per SM (every phase is bound by the same load/store unit, so co-resident CTAs would
// not overlap anything); tiles are taken round-robin, channel group fastest (neighbouring CTAs read the same pixels).
// Weights enter the products rounded to bf16 (what autocast hands to the reference's conv); accumulation is fp32; the
// residual join rounds twice (conv result to bf16, then the sum), once more than the FMA kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200at_launch.cuh"
#include "b200at_tma.cuh"
#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kCG = 16;          // channels per tile
constexpr int kWarps = 16;       // one channel per warp in phase 2
constexpr int kThreads = 32 * kWarps;

struct DwmParams {
  const float* wt;     // [49][C] taps (already flipped by the caller for the input gradient)
  const float* bias;   // [C] or null
  int B, H, W, C;
  int TH, NB, IH;      // output rows per image in a tile, images per tile, TH + 6
  int S;               // plane row stride, bytes (odd multiple of 16)
  int plane_bytes;     // NB * IH * S rounded to 16 mod 32 (plane stride = 4 mod 8 words)
  int sa_bytes, so_bytes;   // input staging / add+output staging (multiples of 1024)
  int tiles_h, groups_b, cgroups, total_tiles;
};

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm4_trans(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// NT = number of 8-column output windows (W <= 8 NT).  Plane columns: p = w + 4, 8 NT + 8 of them; the staging rows hold
// BOXW = 16 * G16 >= 8 NT + 8 pixels of 32 bytes (16 channels), i.e. G16 matrix-move groups of 16 pixels.
// MULTI: tiles hold several images (NB > 1; the row -> (image, row) split costs an integer division per row).
template <int NT, bool ADD, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) dwconv7_mma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_add,
                                                                 const __grid_constant__ CUtensorMap map_y,
                                                                 const DwmParams p) {
  constexpr int kCols = 8 * NT + 8;
  constexpr int G16 = (kCols + 15) / 16;
  constexpr int kBoxW = 16 * G16;
  constexpr int kRowBytes = kBoxW * 32;                 // staging bytes per image row (multiple of 256)
  extern __shared__ __align__(1024) uint8_t dwm_smem_raw[];
  uint8_t* smem = dwm_smem_raw + ((1024u - (b200at::smem_u32(dwm_smem_raw) & 1023u)) & 1023u);
  const uint32_t sa = b200at::smem_u32(smem);            // input staging
  const uint32_t so = sa + p.sa_bytes;                  // add / output staging
  const uint32_t planes = so + p.so_bytes;
  uint8_t* tail = smem + p.sa_bytes + p.so_bytes + (size_t)kCG * p.plane_bytes + 128;
  float* ws = reinterpret_cast<float*>(tail);           // [kCG][49] taps, then [kCG] bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + sizeof(float) * (49 * kCG + kCG));   // [0] input, [1] add
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t pb = (uint32_t)p.plane_bytes;
  // lane's address offset inside a 16-pixel group of a staging row, for the 8x8 matrix moves: matrix mi = (pixel half
  // mi >> 1, channel octet mi & 1), row r8 = pixel; SWIZZLE_32B flips the octet where bit 7 of the offset is set
  const uint32_t mat_off = (uint32_t)(32 * (8 * (mi >> 1) + r8) + 16 * ((mi & 1) ^ ((r8 >> 2) & 1)));

  auto decode = [&](int tile, int& cg, int& n0, int& h0) {
    cg = tile % p.cgroups;
    int s = tile / p.cgroups;
    const int th = s % p.tiles_h;
    n0 = (s / p.tiles_h) * p.NB; h0 = th * p.TH;
  };
  const uint32_t in_bytes = (uint32_t)(p.NB * p.IH) * kRowBytes, add_bytes = (uint32_t)(p.NB * p.TH) * kRowBytes;
  if (tid == 0) {
    b200at::tma_prefetch_desc(&map_x);
    b200at::tma_prefetch_desc(&map_y);
    if (ADD) b200at::tma_prefetch_desc(&map_add);
    b200at::mbar_init(&bars[0], 1);
    b200at::mbar_init(&bars[1], 1);
    b200at::mbar_fence_init();
    if ((int)blockIdx.x < p.total_tiles) {
      int cg, n0, h0;
      decode(blockIdx.x, cg, n0, h0);
      b200at::mbar_expect_tx(&bars[0], in_bytes);
      b200at::tma_load_4d(&map_x, &bars[0], smem, cg * kCG, -4, h0 - 3, n0);
    }
  }
  __syncthreads();

  int it = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
    int cg, n0, h0;
    decode(tile, cg, n0, h0);
    const int c0 = cg * kCG;
    const int rows_out_img = min(p.TH, p.H - h0);        // output rows of this tile per image
    const int rows_in = p.NB * p.IH;
    const int R = p.NB * rows_out_img;                   // output rows of the tile (rows of all its images)
    // taps of this channel group: fetched into registers now (the L2 latency hides under phase 1), parked in shared memory
    // channel-major after phase 1 (the lanes of a warp then read consecutive taps of ONE channel); every warp is past
    // phase 2 of the previous tile here, and the barrier after phase 1 publishes them
    static_assert(49 * kCG <= 2 * kThreads, "two taps per thread");
    const int k1 = tid + kThreads;
    const float tap0 = p.wt[(int64_t)(tid / kCG) * p.C + c0 + (tid % kCG)];
    const float tap1 = k1 < 49 * kCG ? p.wt[(int64_t)(k1 / kCG) * p.C + c0 + (k1 % kCG)] : 0.0f;
    const float bias_v = (tid < kCG && p.bias) ? p.bias[c0 + tid] : 0.0f;

    // ---- phase 1: staging -> channel planes
    b200at::mbar_wait(&bars[0], (uint32_t)(it & 1));
    for (int t = warp; t < rows_in * G16; t += kWarps) {
      const int ri = t / G16, grp = t - ri * G16;
      uint32_t v[4];
      ldsm4_trans(v, sa + ri * kRowBytes + grp * 512 + mat_off);
      const uint32_t dst = planes + g * pb + ri * p.S + (16 * grp + 2 * q) * 2;
      sts32(dst, v[0]);
      sts32(dst + 8 * pb, v[1]);
      if (16 * grp + 8 < kCols) {
        sts32(dst + 16, v[2]);
        sts32(dst + 8 * pb + 16, v[3]);
      }
    }
    ws[(tid % kCG) * 49 + tid / kCG] = tap0;
    if (k1 < 49 * kCG) ws[(k1 % kCG) * 49 + k1 / kCG] = tap1;
    if (tid < kCG) ws[49 * kCG + tid] = bias_v;
    __syncthreads();
    if (tid == 0) {
      b200at::fence_proxy_async();
      b200at::tma_store_wait_read();                     // the previous tile's store no longer reads the output staging
      const int next = tile + gridDim.x;
      if (next < p.total_tiles) {
        int cg1, n1, h1;
        decode(next, cg1, n1, h1);
        b200at::mbar_expect_tx(&bars[0], in_bytes);
        b200at::tma_load_4d(&map_x, &bars[0], smem, cg1 * kCG, -4, h1 - 3, n1);
      }
      if (ADD) {
        b200at::mbar_expect_tx(&bars[1], add_bytes);
        b200at::tma_load_4d(&map_add, &bars[1], smem + p.sa_bytes, c0, -4, h0, n0);
      }
    }

    // ---- phase 2: Toeplitz products, channel = warp
    {
      const uint32_t plane = planes + warp * pb;
      // B[k][n] = tap[dh][k - n - 1]: window row k = 2q (+1) (+8 for the second register), output column n = g.  For a
      // lane exactly one of its two registers meets the 7-tap band (2q - g - 1 >= -1: the first, else the second).
      uint32_t bfr[7][2];
      {
        const int j0 = 2 * q - g - 1;
        const bool first = j0 >= -1;
        const int j = first ? j0 : j0 + 8;
        const float* wc = ws + warp * 49;
#pragma unroll
        for (int dh = 0; dh < 7; ++dh) {
          const float w0 = (j >= 0 && j < 7) ? wc[dh * 7 + j] : 0.0f;
          const float w1 = (j + 1 >= 0 && j + 1 < 7) ? wc[dh * 7 + j + 1] : 0.0f;
          const uint32_t v = pack_bf16x2(w0, w1);
          bfr[dh][0] = first ? v : 0u;
          bfr[dh][1] = first ? 0u : v;
        }
      }
      const float bias = ws[49 * kCG + warp];
      const int mtiles = (R + 15) >> 4;
#pragma unroll 1
      for (int t = 0; t < mtiles; ++t) {
        int r = 16 * t + (mi & 1) * 8 + r8;
        r = r < R ? r : 0;
        const int img = MULTI ? r / rows_out_img : 0, hr = r - img * rows_out_img;
        const uint32_t a_base = plane + (img * p.IH + hr) * p.S + (mi >> 1) * 16;
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[j][0] = bias; acc[j][1] = bias; acc[j][2] = bias; acc[j][3] = bias; }
#pragma unroll
        for (int dh = 0; dh < 7; ++dh) {
          uint32_t a[(NT + 2) / 2][4];
#pragma unroll
          for (int cp = 0; cp < (NT + 2) / 2; ++cp) ldsm4(a[cp], a_base + dh * p.S + cp * 32);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if ((j & 1) == 0) mma16816(acc[j], a[j / 2][0], a[j / 2][1], a[j / 2][2], a[j / 2][3], bfr[dh][0], bfr[dh][1]);
            else mma16816(acc[j], a[j / 2][2], a[j / 2][3], a[(j + 1) / 2][0], a[(j + 1) / 2][1], bfr[dh][0], bfr[dh][1]);
          }
        }
        // results over the plane rows that no later M tile reads (output row hr sits in plane row hr)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ro = 16 * t + g + 8 * half;
          if (ro < R) {
            const int im = MULTI ? ro / rows_out_img : 0, ho = ro - im * rows_out_img;
            const uint32_t dst = plane + (im * p.IH + ho) * p.S + (2 * q + 4) * 2;
#pragma unroll
            for (int j = 0; j < NT; ++j) sts32(dst + j * 16, pack_bf16x2(acc[j][2 * half], acc[j][2 * half + 1]));
          }
        }
      }
    }
    __syncthreads();

    // ---- phase 3: planes (+ add tile) -> output staging, pixel-major again
    if (ADD) b200at::mbar_wait(&bars[1], (uint32_t)(it & 1));
    for (int t = warp; t < R * G16; t += kWarps) {
      const int r = t / G16, grp = t - r * G16;
      const int img = MULTI ? r / rows_out_img : 0, hr = r - img * rows_out_img;
      const uint32_t src = planes + g * pb + (img * p.IH + hr) * p.S + (16 * grp + 2 * q) * 2;
      uint32_t v[4];
      v[0] = lds32(src);
      v[1] = lds32(src + 8 * pb);
      if (16 * grp + 8 < kCols) { v[2] = lds32(src + 16); v[3] = lds32(src + 8 * pb + 16); }
      else { v[2] = 0u; v[3] = 0u; }
      const uint32_t st = so + (img * p.TH + hr) * kRowBytes + grp * 512 + mat_off;
      if (ADD) {
        uint32_t u[4];
        ldsm4_trans(u, st);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = add_bf16x2(v[i], u[i]);
      }
      stsm4_trans(st, v);
    }
    b200at::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      // the store box starts at w = 0 (a store with a negative start coordinate faults: "illegal instruction" on B200),
      // i.e. 4 pixels = 128 bytes into the staging rows; the swizzle is a function of the absolute shared-memory address,
      // so the shifted source keeps the pattern.  The last 4 pixels of every box row come from the next staging row: they
      // are columns >= W, clipped.
      b200at::tma_store_4d(&map_y, smem + p.sa_bytes + 128, c0, 0, h0, n0);
      b200at::tma_store_commit();
    }
  }
  if (tid == 0) b200at::tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------- v4: two warp groups
// The three phases of a tile are all bound by the load/store unit, but inside ONE CTA they ran back to back with barriers
// in between, and the unit idled 45 % of the time (profiles/r02_dwconv_mma_ncu_v3.txt).  Here warps 0-7 ("T") only move
// data -- phase 1 of tile i+1 and phase 3 of tile i-1 -- while warps 8-15 ("M", two channels each) run phase 2 of tile i
// on the other of two plane buffers.  Hand-off through mbarriers (planes_full / planes_done, one arrival per warp);
// the T group synchronises with itself through named barrier 1.
__device__ __forceinline__ void bar_t_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int NT, bool ADD, bool MULTI>
__global__ void __launch_bounds__(kThreads, 1) dwconv7_mma_pp_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                    const __grid_constant__ CUtensorMap map_add,
                                                                    const __grid_constant__ CUtensorMap map_y,
                                                                    const DwmParams p) {
  constexpr int kCols = 8 * NT + 8;
  constexpr int G16 = (kCols + 15) / 16;
  constexpr int kBoxW = 16 * G16;
  constexpr int kRowBytes = kBoxW * 32;
  constexpr int kWs = 49 * kCG + kCG;                     // floats per tap buffer
  extern __shared__ __align__(1024) uint8_t dwm_smem_raw[];
  uint8_t* smem = dwm_smem_raw + ((1024u - (b200at::smem_u32(dwm_smem_raw) & 1023u)) & 1023u);
  const uint32_t sa = b200at::smem_u32(smem);
  const uint32_t so = sa + p.sa_bytes;
  const uint32_t planes0 = so + p.so_bytes;
  const uint32_t pbuf = (uint32_t)kCG * p.plane_bytes + 128;                 // bytes per plane buffer
  uint8_t* tail = smem + p.sa_bytes + p.so_bytes + 2 * (size_t)pbuf;
  float* ws0 = reinterpret_cast<float*>(tail);            // [2][kWs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * sizeof(float) * kWs);
  uint64_t* in_full = &bars[0];
  uint64_t* add_full = &bars[1];
  uint64_t* planes_full = &bars[2];                       // [2]
  uint64_t* planes_done = &bars[4];                       // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t pb = (uint32_t)p.plane_bytes;
  const uint32_t mat_off = (uint32_t)(32 * (8 * (mi >> 1) + r8) + 16 * ((mi & 1) ^ ((r8 >> 2) & 1)));
  auto decode = [&](int tile, int& cg, int& n0, int& h0) {
    cg = tile % p.cgroups;
    int s = tile / p.cgroups;
    const int th = s % p.tiles_h;
    n0 = (s / p.tiles_h) * p.NB; h0 = th * p.TH;
  };
  const uint32_t in_bytes = (uint32_t)(p.NB * p.IH) * kRowBytes, add_bytes = (uint32_t)(p.NB * p.TH) * kRowBytes;
  const int ntiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  if (tid == 0) {
    b200at::tma_prefetch_desc(&map_x);
    b200at::tma_prefetch_desc(&map_y);
    if (ADD) b200at::tma_prefetch_desc(&map_add);
    b200at::mbar_init(in_full, 1);
    b200at::mbar_init(add_full, 1);
    for (int b = 0; b < 2; ++b) { b200at::mbar_init(&planes_full[b], 8); b200at::mbar_init(&planes_done[b], 8); }
    b200at::mbar_fence_init();
    if (ntiles > 0) {
      int cg, n0, h0;
      decode(blockIdx.x, cg, n0, h0);
      b200at::mbar_expect_tx(in_full, in_bytes);
      b200at::tma_load_4d(&map_x, in_full, smem, cg * kCG, -4, h0 - 3, n0);
      if (ADD) {
        b200at::mbar_expect_tx(add_full, add_bytes);
        b200at::tma_load_4d(&map_add, add_full, smem + p.sa_bytes, cg * kCG, -4, h0, n0);
      }
    }
  }
  __syncthreads();

  if (warp < 8) {
    // ================================================================ T group: phase 1 of tile i, phase 3 of tile i - 1
    auto phase3 = [&](int j) {      // tile index j (of this CTA): planes[j & 1] -> output staging -> TMA store
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      int cg, n0, h0;
      decode(tile, cg, n0, h0);
      const int rows_out_img = min(p.TH, p.H - h0);
      const int R = p.NB * rows_out_img;
      const uint32_t planes = planes0 + (j & 1) * pbuf;
      b200at::mbar_wait(&planes_done[j & 1], (uint32_t)((j >> 1) & 1));
      if (ADD) b200at::mbar_wait(add_full, (uint32_t)(j & 1));
      for (int t = warp; t < R * G16; t += 8) {
        const int r = t / G16, grp = t - r * G16;
        const int img = MULTI ? r / rows_out_img : 0, hr = r - img * rows_out_img;
        const uint32_t src = planes + g * pb + (img * p.IH + hr) * p.S + (16 * grp + 2 * q) * 2;
        uint32_t v[4];
        v[0] = lds32(src);
        v[1] = lds32(src + 8 * pb);
        if (16 * grp + 8 < kCols) { v[2] = lds32(src + 16); v[3] = lds32(src + 8 * pb + 16); }
        else { v[2] = 0u; v[3] = 0u; }
        const uint32_t st = so + (img * p.TH + hr) * kRowBytes + grp * 512 + mat_off;
        if (ADD) {
          uint32_t u[4];
          ldsm4_trans(u, st);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = add_bf16x2(v[i], u[i]);
        }
        stsm4_trans(st, v);
      }
      b200at::fence_proxy_async();
      bar_t_sync();
      if (tid == 0) {
        b200at::tma_store_4d(&map_y, smem + p.sa_bytes + 128, cg * kCG, 0, h0, n0);
        b200at::tma_store_commit();
        if (ADD && j + 1 < ntiles) {                      // the next tile's residual-gradient tile, once the store has read SO
          int cg1, n1, h1;
          decode(tile + (int)gridDim.x, cg1, n1, h1);
          b200at::tma_store_wait_read();
          b200at::mbar_expect_tx(add_full, add_bytes);
          b200at::tma_load_4d(&map_add, add_full, smem + p.sa_bytes, cg1 * kCG, -4, h1, n1);
        }
      }
    };
    for (int i = 0; i < ntiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      int cg, n0, h0;
      decode(tile, cg, n0, h0);
      const int c0 = cg * kCG;
      const int rows_in = p.NB * p.IH;
      const uint32_t planes = planes0 + (i & 1) * pbuf;
      float* ws = ws0 + (i & 1) * kWs;
      // taps of this channel group into registers (L2 latency under the wait below)
      float tapv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kk = tid + 256 * k;
        tapv[k] = kk < 49 * kCG ? p.wt[(int64_t)(kk / kCG) * p.C + c0 + (kk % kCG)] : 0.0f;
      }
      const float bias_v = (tid < kCG && p.bias) ? p.bias[c0 + tid] : 0.0f;
      // planes[i & 1] was last read by phase 3 of tile i - 2 (this group, two iterations ago, behind a group barrier)
      b200at::mbar_wait(in_full, (uint32_t)(i & 1));
      for (int t = warp; t < rows_in * G16; t += 8) {
        const int ri = t / G16, grp = t - ri * G16;
        uint32_t v[4];
        ldsm4_trans(v, sa + ri * kRowBytes + grp * 512 + mat_off);
        const uint32_t dst = planes + g * pb + ri * p.S + (16 * grp + 2 * q) * 2;
        sts32(dst, v[0]);
        sts32(dst + 8 * pb, v[1]);
        if (16 * grp + 8 < kCols) {
          sts32(dst + 16, v[2]);
          sts32(dst + 8 * pb + 16, v[3]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kk = tid + 256 * k;
        if (kk < 49 * kCG) ws[(kk % kCG) * 49 + kk / kCG] = tapv[k];
      }
      if (tid < kCG) ws[49 * kCG + tid] = bias_v;
      __syncwarp();
      if (lane == 0) b200at::mbar_arrive(&planes_full[i & 1]);           // release: planes + taps of tile i are complete
      if (tid == 0) b200at::tma_store_wait_read();                        // the previous store has read the output staging
      bar_t_sync();                                                       // every T warp is done reading the input staging
      if (tid == 0 && i + 1 < ntiles) {
        int cg1, n1, h1;
        decode(tile + (int)gridDim.x, cg1, n1, h1);
        b200at::fence_proxy_async();
        b200at::mbar_expect_tx(in_full, in_bytes);
        b200at::tma_load_4d(&map_x, in_full, smem, cg1 * kCG, -4, h1 - 3, n1);
      }
      if (i > 0) phase3(i - 1);
    }
    if (ntiles > 0) phase3(ntiles - 1);
    if (tid == 0) b200at::tma_store_wait_all();
  } else {
    // ================================================================ M group: phase 2, channels 2 (warp - 8), 2 (warp - 8) + 1
    const int mw = warp - 8;
    for (int i = 0; i < ntiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      int cg, n0, h0;
      decode(tile, cg, n0, h0);
      const int rows_out_img = min(p.TH, p.H - h0);
      const int R = p.NB * rows_out_img;
      const uint32_t planes = planes0 + (i & 1) * pbuf;
      const float* ws = ws0 + (i & 1) * kWs;
      b200at::mbar_wait(&planes_full[i & 1], (uint32_t)((i >> 1) & 1));
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * mw + cc;
        const uint32_t plane = planes + c * pb;
        uint32_t bfr[7][2];
        {
          const int j0 = 2 * q - g - 1;
          const bool first = j0 >= -1;
          const int j = first ? j0 : j0 + 8;
          const float* wc = ws + c * 49;
#pragma unroll
          for (int dh = 0; dh < 7; ++dh) {
            const float w0 = (j >= 0 && j < 7) ? wc[dh * 7 + j] : 0.0f;
            const float w1 = (j + 1 >= 0 && j + 1 < 7) ? wc[dh * 7 + j + 1] : 0.0f;
            const uint32_t v = pack_bf16x2(w0, w1);
            bfr[dh][0] = first ? v : 0u;
            bfr[dh][1] = first ? 0u : v;
          }
        }
        const float bias = ws[49 * kCG + c];
        const int mtiles = (R + 15) >> 4;
#pragma unroll 1
        for (int t = 0; t < mtiles; ++t) {
          int r = 16 * t + (mi & 1) * 8 + r8;
          r = r < R ? r : 0;
          const int img = MULTI ? r / rows_out_img : 0, hr = r - img * rows_out_img;
          const uint32_t a_base = plane + (img * p.IH + hr) * p.S + (mi >> 1) * 16;
          float acc[NT][4];
#pragma unroll
          for (int j = 0; j < NT; ++j) { acc[j][0] = bias; acc[j][1] = bias; acc[j][2] = bias; acc[j][3] = bias; }
#pragma unroll
          for (int dh = 0; dh < 7; ++dh) {
            uint32_t a[(NT + 2) / 2][4];
#pragma unroll
            for (int cp = 0; cp < (NT + 2) / 2; ++cp) ldsm4(a[cp], a_base + dh * p.S + cp * 32);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              if ((j & 1) == 0) mma16816(acc[j], a[j / 2][0], a[j / 2][1], a[j / 2][2], a[j / 2][3], bfr[dh][0], bfr[dh][1]);
              else mma16816(acc[j], a[j / 2][2], a[j / 2][3], a[(j + 1) / 2][0], a[(j + 1) / 2][1], bfr[dh][0], bfr[dh][1]);
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int ro = 16 * t + g + 8 * half;
            if (ro < R) {
              const int im = MULTI ? ro / rows_out_img : 0, ho = ro - im * rows_out_img;
              const uint32_t dst = plane + (im * p.IH + ho) * p.S + (2 * q + 4) * 2;
#pragma unroll
              for (int j = 0; j < NT; ++j) sts32(dst + j * 16, pack_bf16x2(acc[j][2 * half], acc[j][2 * half + 1]));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) b200at::mbar_arrive(&planes_done[i & 1]);            // release: the outputs of this warp's channels
    }
  }
}

template <int NT, bool ADD, bool MULTI>
int launch_one(const CUtensorMap& mx, const CUtensorMap& ma, const CUtensorMap& my, const DwmParams& p, size_t smem, int grid,
               cudaStream_t s) {
  // two warp groups on two plane buffers when they fit (B200AT_DWM_PP=0: the single-group kernel)
  static const bool pp_on = [] { const char* e = getenv("B200AT_DWM_PP"); return e == nullptr || e[0] != '0'; }();
  const size_t smem_pp = 1024 + (size_t)p.sa_bytes + p.so_bytes + 2 * ((size_t)kCG * p.plane_bytes + 128) +
                         2 * sizeof(float) * (49 * kCG + kCG) + 64;
  if (pp_on && !MULTI && smem_pp <= 225 * 1024) {     // several images per tile: the single group measured faster (r02n)
    static b200at::SmemConfig conf_pp;
    cudaError_t e = b200at::ensure_dynamic_smem(dwconv7_mma_pp_kernel<NT, ADD, MULTI>, (int)smem_pp, conf_pp);
    if (e != cudaSuccess) return (int)e;
    dwconv7_mma_pp_kernel<NT, ADD, MULTI><<<grid, kThreads, smem_pp, s>>>(mx, ma, my, p);
    return (int)cudaGetLastError();
  }
  static b200at::SmemConfig conf;
  cudaError_t e = b200at::ensure_dynamic_smem(dwconv7_mma_kernel<NT, ADD, MULTI>, (int)smem, conf);
  if (e != cudaSuccess) return (int)e;
  dwconv7_mma_kernel<NT, ADD, MULTI><<<grid, kThreads, smem, s>>>(mx, ma, my, p);
  return (int)cudaGetLastError();
}

template <int NT>
int launch(const CUtensorMap& mx, const CUtensorMap& ma, const CUtensorMap& my, const DwmParams& p, bool add, size_t smem,
           int grid, cudaStream_t s) {
  if (p.NB > 1) return add ? launch_one<NT, true, true>(mx, ma, my, p, smem, grid, s) : launch_one<NT, false, true>(mx, ma, my, p, smem, grid, s);
  return add ? launch_one<NT, true, false>(mx, ma, my, p, smem, grid, s) : launch_one<NT, false, false>(mx, ma, my, p, smem, grid, s);
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

inline size_t smem_need(int NT, int TH, int NB, DwmParams* p) {
  const int cols = 8 * NT + 8, boxw = (cols + 15) / 16 * 16;
  int units = NT + 1;                                   // 16-byte units of the plane columns
  if ((units & 1) == 0) units += 1;                     // odd multiple of 16 bytes: conflict-free ldmatrix rows
  const int S = units * 16, IH = TH + 6;
  int plane = NB * IH * S;
  if ((plane & 31) != 16) plane += 16;                  // plane stride = 4 mod 8 words: conflict-free matrix-move stores
  const int sa = (NB * IH * boxw * 32 + 1023) / 1024 * 1024, so = (NB * TH * boxw * 32 + 1023) / 1024 * 1024;
  if (p) { p->S = S; p->IH = IH; p->TH = TH; p->NB = NB; p->plane_bytes = plane; p->sa_bytes = sa; p->so_bytes = so; }
  return 1024 + (size_t)sa + so + (size_t)kCG * plane + 128 + sizeof(float) * (49 * kCG + kCG) + 64;
}

}  // namespace

// Returns -1 when the shape is outside this kernel (the caller falls back to the FMA kernel), else the cudaError_t.
int b200at_dwconv7_mma_launch(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B,
                              int64_t H, int64_t W, int64_t C, void* stream) {
  // Maps narrower than 20 columns (stages 2-3 at 224 px) stay on the FMA kernel: with 16 or 8 output columns per row the
  // windows are mostly padding and the per-tile overheads dominate (profiles/r02_ops_bench_dwconv_mma_v3.txt: 49 vs 39 us at
  // 14 x 14 x 384, 33 vs 27 us at 7 x 7 x 768).  B200AT_DWM_MINW overrides the threshold (tests run every shape with 1).
  static const int min_w = env_int("B200AT_DWM_MINW", 20);
  if (W < min_w) return -1;
  if (C % kCG || W < 1 || W > 80 || H < 1 || H > 4096 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15) || (add && (reinterpret_cast<uintptr_t>(add) & 15)))
    return -1;
  int NT = (int)((W + 7) / 8);                          // instantiated window counts
  if (NT == 6) NT = 7;
  if (NT == 8 || NT == 9) NT = 10;
  DwmParams p;
  p.wt = wt; p.bias = bias;
  p.B = (int)B; p.H = (int)H; p.W = (int)W; p.C = (int)C;
  // tile: NB images x TH output rows x the full width x 16 channels.  The M tiles take 16 output rows of the tile (of any of
  // its images), so NB * TH close to a multiple of 16 keeps the tensor pipe full; planes + both staging buffers must fit
  // the SM's shared memory.
  static const int th_force = env_int("B200AT_DWM_TH", 0), nb_force = env_int("B200AT_DWM_NB", 0);
  const size_t cap = 200 * 1024;
  int TH = H > 28 ? 16 : (int)H, NB = 1;
  if (th_force > 0) TH = th_force;
  if (TH > H) TH = (int)H;
  if (smem_need(NT, TH, 1, nullptr) > cap) TH = 16 < TH ? 16 : TH;
  if (smem_need(NT, TH, 1, nullptr) > cap) return -1;
  if (H <= 28) {                                         // small maps: several images per tile
    while (NB < 8 && 2 * NB <= B && 2 * NB * TH <= 112 && smem_need(NT, TH, 2 * NB, nullptr) <= cap) NB *= 2;
  }
  if (nb_force > 0 && smem_need(NT, TH, nb_force, nullptr) <= cap) NB = nb_force;
  const size_t smem = smem_need(NT, TH, NB, &p);
  p.tiles_h = (int)((H + TH - 1) / TH);
  p.groups_b = (int)((B + NB - 1) / NB);
  p.cgroups = (int)(C / kCG);
  const int64_t total = (int64_t)p.cgroups * p.tiles_h * p.groups_b;
  if (total > 0x7fffffff) return -1;
  p.total_tiles = (int)total;
  const int boxw = ((8 * NT + 8) + 15) / 16 * 16;
  CUtensorMap mx, ma, my;
  if (!b200at::make_map_nhwc_bf16(&mx, x, B, H, W, C, kCG, boxw, p.IH, NB, true)) return (int)cudaErrorUnknown;
  if (!b200at::make_map_nhwc_bf16(&my, y, B, H, W, C, kCG, boxw, TH, NB, true)) return (int)cudaErrorUnknown;
  if (add) { if (!b200at::make_map_nhwc_bf16(&ma, add, B, H, W, C, kCG, boxw, TH, NB, true)) return (int)cudaErrorUnknown; }
  else ma = my;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(total < sms ? total : sms);
  cudaStream_t s = (cudaStream_t)stream;
  switch (NT) {
    case 1: return launch<1>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 2: return launch<2>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 3: return launch<3>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 4: return launch<4>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 5: return launch<5>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 7: return launch<7>(mx, ma, my, p, add != nullptr, smem, grid, s);
    case 10: return launch<10>(mx, ma, my, p, add != nullptr, smem, grid, s);
    default: return -1;
  }
}
