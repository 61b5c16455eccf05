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
// The channel is a batch dimension of these products, while memory is channel-innermost, so a tile is transposed on its
// way into shared memory: [channel][row][column] planes of bf16 with the column contiguous (what ldmatrix wants), row stride
// an odd multiple of 16 bytes (conflict-free ldmatrix, conflict-free fragment stores).
//   phase 1  global -> registers -> planes: a thread takes 2 adjacent pixels x 8 channels (2 x LDG.128; the lane pair
//            of a pixel covers its 32 contiguous bytes), PRMT pairs them per channel, 8 x STS.32.  Halo / padding is
//            zero-filled here.
//   phase 2  warp w owns channels 2w, 2w+1: Toeplitz fragments from the staged taps, then per 16-row M tile and tap row dh
//            ceil((NT+1)/2) ldmatrix.x4 + NT mma (window j uses 8-column chunks j, j+1: each chunk is loaded once and used
//            by two windows), fp32 accumulators (+ bias), results packed to bf16 IN PLACE over input rows that later
//            M tiles no longer need.
//   phase 3  planes -> registers -> global: the inverse transposition, (+ `add`: the residual-gradient join of the
//            block's backward, models/convnext.py:49), 32-byte contiguous stores per pixel.
// Phases of different CTAs overlap on an SM (3-4 CTAs resident); no barriers other than the two __syncthreads.
// Weights enter the products rounded to bf16 (what autocast hands to the reference's conv); accumulation is fp32.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200at_launch.cuh"
#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kCG = 16;          // channels per CTA
constexpr int kThreads = 256;    // 8 warps x 2 channels

struct DwmParams {
  const bf16* x;
  const float* wt;     // [49][C] taps (already flipped by the caller for the input gradient)
  const float* bias;   // [C] or null
  const bf16* add;     // [B][H][W][C] or null
  bf16* y;
  int B, H, W, C;
  int TH, NB, IH;      // output rows per image in a tile, images per tile, TH + 6
  int S;               // plane row stride, bytes (odd multiple of 16)
  int plane_bytes;     // NB * IH * S (multiple of 16)
  int tiles_h, groups_b, cgroups;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ uint4 ldg128(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// byte offset of channel plane c: planes 8..15 start 64 bytes (16 banks) later, so the two octets a warp of the
// transposing phases works on at once land in disjoint halves of the 32 banks
__device__ __forceinline__ uint32_t plane_off(int c, uint32_t plane_bytes) { return (uint32_t)c * plane_bytes + ((c & 8) ? 64u : 0u); }

// NT = number of 8-column output windows (W <= 8 NT).  Plane columns: p = w + 4, 8 NT + 8 of them.
template <int NT, bool ADD>
__global__ void __launch_bounds__(kThreads, (NT > 7 ? 2 : 3)) dwconv7_mma_kernel(const DwmParams p) {
  extern __shared__ __align__(16) uint8_t dwm_smem[];
  const uint32_t planes = smem_addr(dwm_smem);
  float* ws = reinterpret_cast<float*>(dwm_smem + (size_t)kCG * p.plane_bytes + 128);  // [kCG][49] taps, then [kCG] bias
  const int tid = threadIdx.x;

  int bid = blockIdx.x;
  const int cg = bid % p.cgroups; bid /= p.cgroups;
  const int th = bid % p.tiles_h; bid /= p.tiles_h;
  const int n0 = bid * p.NB, h0 = th * p.TH, c0 = cg * kCG;
  const int rows_out_img = min(p.TH, p.H - h0);          // output rows of this tile per image
  const int rows_in = p.NB * p.IH;
  const int PAIRS = (p.W + 1) >> 1;

  // channel-major in shared memory: the lanes of a warp then read consecutive taps of ONE channel (conflict-free; the
  // tap-major form cost as many shared-memory wavefronts as all the ldmatrix of the tile: profiles/r02_dwconv_mma_ncu_v1.txt)
  for (int k = tid; k < 49 * kCG; k += kThreads) ws[(k % kCG) * 49 + k / kCG] = p.wt[(int64_t)(k / kCG) * p.C + c0 + (k % kCG)];
  if (tid < kCG) ws[49 * kCG + tid] = p.bias ? p.bias[c0 + tid] : 0.0f;

  // ---- phase 1: transpose the halo tile into channel planes.  Task = (plane row, pixel pair, channel octet); the two
  // octets of a pixel pair sit in adjacent lanes, so one LDG.128 of a warp covers 16 pixels x 32 contiguous bytes (one
  // L1 wavefront per pixel instead of two: the load/store unit's wavefront rate is what bounds this kernel).
  {
    const int ntasks = rows_in * PAIRS * 2;
    const uint32_t pb = (uint32_t)p.plane_bytes;
    for (int t = tid; t < ntasks; t += kThreads) {
      const int o = t & 1, tp = t >> 1;
      const int ri = tp / PAIRS, pp = tp - ri * PAIRS;
      const int img = ri / p.IH, rr = ri - img * p.IH;
      const int n = n0 + img, h = h0 + rr - 3, w = 2 * pp;
      uint4 a0 = make_uint4(0u, 0u, 0u, 0u), a1 = a0;
      if (n < p.B && h >= 0 && h < p.H) {
        const bf16* src = p.x + (((int64_t)n * p.H + h) * p.W + w) * p.C + c0 + 8 * o;
        a0 = ldg128(src);
        if (w + 1 < p.W) a1 = ldg128(src + p.C);
      }
      const uint32_t dst = planes + plane_off(8 * o, pb) + ri * p.S + (w + 4) * 2;
      sts32(dst + 0 * pb, __byte_perm(a0.x, a1.x, 0x5410)); sts32(dst + 1 * pb, __byte_perm(a0.x, a1.x, 0x7632));
      sts32(dst + 2 * pb, __byte_perm(a0.y, a1.y, 0x5410)); sts32(dst + 3 * pb, __byte_perm(a0.y, a1.y, 0x7632));
      sts32(dst + 4 * pb, __byte_perm(a0.z, a1.z, 0x5410)); sts32(dst + 5 * pb, __byte_perm(a0.z, a1.z, 0x7632));
      sts32(dst + 6 * pb, __byte_perm(a0.w, a1.w, 0x5410)); sts32(dst + 7 * pb, __byte_perm(a0.w, a1.w, 0x7632));
    }
    // zero the padding columns: words [0, 2) and [(W + 5) / 2, 4 NT + 4) of every plane row
    const int first_pad = (p.W + 5) >> 1;
    const int npad = 2 + (4 * NT + 4 - first_pad);
    const int nz = kCG * rows_in * npad;
    for (int t = tid; t < nz; t += kThreads) {
      const int k = t % npad, row = t / npad;
      const int c = row / rows_in, ri = row - c * rows_in;
      const int word = k < 2 ? k : first_pad + (k - 2);
      sts32(planes + plane_off(c, pb) + ri * p.S + word * 4, 0u);
    }
  }
  __syncthreads();

  // ---- phase 2: Toeplitz products
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int mi = lane >> 3, r8 = lane & 7;
    const int R = p.NB * rows_out_img;                   // output rows of the tile (rows of all its images)
    const int mtiles = (R + 15) >> 4;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = warp * 2 + cc;
      const uint32_t plane = planes + plane_off(c, (uint32_t)p.plane_bytes);
      // B[k][n] = tap[dh][k - n - 1]: window row k = 2q (+1) (+8 for the second register), output column n = g.  For a
      // lane exactly one of its two registers meets the 7-tap band (2q - g - 1 >= -1: the first, else the second).
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
#pragma unroll 1
      for (int t = 0; t < mtiles; ++t) {
        int r = 16 * t + (mi & 1) * 8 + r8;
        r = r < R ? r : 0;
        const int img = r / rows_out_img, hr = r - img * rows_out_img;
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
            const int im = ro / rows_out_img, ho = ro - im * rows_out_img;
            const uint32_t dst = plane + (im * p.IH + ho) * p.S + (2 * q + 4) * 2;
#pragma unroll
            for (int j = 0; j < NT; ++j) sts32(dst + j * 16, pack_bf16x2(acc[j][2 * half], acc[j][2 * half + 1]));
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 3: transpose back and store (same task shape as phase 1: 32 contiguous bytes per lane pair)
  {
    const int R = p.NB * rows_out_img;
    const int ntasks = R * PAIRS * 2;
    const uint32_t pb = (uint32_t)p.plane_bytes;
    for (int t = tid; t < ntasks; t += kThreads) {
      const int o = t & 1, tp = t >> 1;
      const int r = tp / PAIRS, pp = tp - r * PAIRS;
      const int img = r / rows_out_img, hr = r - img * rows_out_img;
      const int n = n0 + img, h = h0 + hr, w = 2 * pp;
      if (n >= p.B) continue;
      const uint32_t src = planes + plane_off(8 * o, pb) + (img * p.IH + hr) * p.S + (w + 4) * 2;
      uint32_t wd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wd[i] = lds32(src + i * pb);
      uint4 o0, o1;                  // pixel w, pixel w + 1: channels 8 o .. 8 o + 7
      o0.x = __byte_perm(wd[0], wd[1], 0x5410); o1.x = __byte_perm(wd[0], wd[1], 0x7632);
      o0.y = __byte_perm(wd[2], wd[3], 0x5410); o1.y = __byte_perm(wd[2], wd[3], 0x7632);
      o0.z = __byte_perm(wd[4], wd[5], 0x5410); o1.z = __byte_perm(wd[4], wd[5], 0x7632);
      o0.w = __byte_perm(wd[6], wd[7], 0x5410); o1.w = __byte_perm(wd[6], wd[7], 0x7632);
      const int64_t off = (((int64_t)n * p.H + h) * p.W + w) * p.C + c0 + 8 * o;
      const bool second = w + 1 < p.W;
      if (ADD) {
        const uint4 r0 = ldg128(p.add + off);
        o0.x = add_bf16x2(o0.x, r0.x); o0.y = add_bf16x2(o0.y, r0.y); o0.z = add_bf16x2(o0.z, r0.z); o0.w = add_bf16x2(o0.w, r0.w);
        if (second) {
          const uint4 r1 = ldg128(p.add + off + p.C);
          o1.x = add_bf16x2(o1.x, r1.x); o1.y = add_bf16x2(o1.y, r1.y); o1.z = add_bf16x2(o1.z, r1.z); o1.w = add_bf16x2(o1.w, r1.w);
        }
      }
      *reinterpret_cast<uint4*>(p.y + off) = o0;
      if (second) *reinterpret_cast<uint4*>(p.y + off + p.C) = o1;
    }
  }
}

template <int NT>
int launch(const DwmParams& p, size_t smem, int grid, cudaStream_t s) {
  static std::atomic<uint64_t> conf_add{0}, conf_plain{0};
  cudaError_t e;
  if (p.add) {
    e = b200at::ensure_dynamic_smem(dwconv7_mma_kernel<NT, true>, (int)smem, conf_add);
    if (e != cudaSuccess) return (int)e;
    dwconv7_mma_kernel<NT, true><<<grid, kThreads, smem, s>>>(p);
  } else {
    e = b200at::ensure_dynamic_smem(dwconv7_mma_kernel<NT, false>, (int)smem, conf_plain);
    if (e != cudaSuccess) return (int)e;
    dwconv7_mma_kernel<NT, false><<<grid, kThreads, smem, s>>>(p);
  }
  return (int)cudaGetLastError();
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// Returns -1 when the shape is outside this kernel (the caller falls back to the FMA kernel), else the cudaError_t.
int b200at_dwconv7_mma_launch(const void* x, const float* wt, const float* bias, const void* add, void* y, int64_t B,
                              int64_t H, int64_t W, int64_t C, void* stream) {
  if (C % kCG || W < 1 || W > 80 || H < 1 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) ||
      (add && (reinterpret_cast<uintptr_t>(add) & 15)))
    return -1;
  const int NTw = (int)((W + 7) / 8);
  int NT = NTw;                                         // instantiated window counts
  if (NT == 6) NT = 7;
  if (NT == 8 || NT == 9) NT = 10;
  DwmParams p;
  p.x = (const bf16*)x; p.wt = wt; p.bias = bias; p.add = (const bf16*)add; p.y = (bf16*)y;
  p.B = (int)B; p.H = (int)H; p.W = (int)W; p.C = (int)C;
  int units = NT + 1;                                   // 16-byte units of the 8 NT + 8 plane columns
  if ((units & 1) == 0) units += 1;                     // odd multiple of 16 bytes: conflict-free ldmatrix rows
  p.S = units * 16;
  // tile: NB images x TH output rows x the full width x 16 channels.  The M tiles take 16 output rows of the tile (of any of
  // its images), so NB * TH close to a multiple of 16 keeps the tensor pipe full; shared memory (planes hold TH + 6 rows per
  // image) bounds it to 3-4 resident CTAs per SM.
  static const int th_force = env_int("B200AT_DWM_TH", 0), nb_force = env_int("B200AT_DWM_NB", 0);
  int TH, NB;
  if (H >= 32) { TH = 16; NB = 1; }
  else if (H > 16) { TH = (int)((H + 1) / 2); NB = 1; }         // 28 -> 14, 20 -> 10
  else if (H > 8) { TH = (int)H; NB = 4; }                      // 14 -> 4 x 14 = 56 rows, 10 -> 4 x 10
  else { TH = (int)H; NB = 8; }                                 // 7 -> 8 x 7 = 56 rows
  if (th_force > 0) TH = th_force;
  if (nb_force > 0) NB = nb_force;
  if (TH > H) TH = (int)H;
  p.TH = TH; p.NB = NB; p.IH = TH + 6;
  p.plane_bytes = NB * p.IH * p.S;
  p.tiles_h = (int)((H + TH - 1) / TH);
  p.groups_b = (int)((B + NB - 1) / NB);
  p.cgroups = (int)(C / kCG);
  const size_t smem = (size_t)kCG * p.plane_bytes + 128 + sizeof(float) * (49 * kCG + kCG);
  if (smem > 200 * 1024) return -1;
  const int64_t grid = (int64_t)p.cgroups * p.tiles_h * p.groups_b;
  if (grid > 0x7fffffff) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  switch (NT) {
    case 1: return launch<1>(p, smem, (int)grid, s);
    case 2: return launch<2>(p, smem, (int)grid, s);
    case 3: return launch<3>(p, smem, (int)grid, s);
    case 4: return launch<4>(p, smem, (int)grid, s);
    case 5: return launch<5>(p, smem, (int)grid, s);
    case 7: return launch<7>(p, smem, (int)grid, s);
    case 10: return launch<10>(p, smem, (int)grid, s);
    default: return -1;
  }
}
