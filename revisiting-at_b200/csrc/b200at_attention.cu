// sm_100a multi-head self-attention for the ViT-S-CvSt blocks (include/b200at_model.h, K12).
// Reference math: timm 0.8 `vision_transformer.Attention.forward` (un-vendored; call sites
// /root/reference/utils_architecture.py:271-301):
//     q, k, v = qkv(x).reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
//     o = softmax(q k^T * scale) v ; o.transpose(1, 2).reshape(B, N, H * 64)
// The sequence is 197 tokens (<= 208): the whole K / V of one (image, head) lives in shared memory, one CTA
// per (image, head), one warp per block of 16 query rows.  The products run on mma.sync m16n8k16 bf16 with
// fp32 accumulation: at 197 tokens x 64 channels the S = Q K^T tile is 1.5 UMMA tiles and the op is ~8 % of
// the block's FLOPs (the qkv / proj / MLP GEMMs, which are on tcgen05, carry the rest), so the
// register-resident online softmax of the mma.sync form is what matters here, not tensor-pipe peak.
// Forward: online softmax over 16-key blocks, P stays in registers (accumulator layout == A-fragment
// layout), saves the per-row log2-sum-exp.  Backward: pass A (warp = 16 query rows) recomputes P and
// produces dQ; pass B (warp = 16 key rows) recomputes P^T and produces dK, dV -- no atomics, deterministic.
// Operands whose contraction index is the token index (V in P V, K in dS K, dO in P^T dO, Q in dS^T Q) are read from
// the row-major shared tiles through ldmatrix.trans.  The backward runs as two kernels (dQ; dK+dV), 7 warps per CTA
// and 60 KB of shared memory each, so 21 warps share an SM (B200AT_ATTN_BWD=1 selects the older single-kernel form).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/b200at_model.h"

namespace {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

constexpr int kD = 64;            // head dimension
constexpr int kRS = kD + 8;       // row stride (elements) of the row-major tiles: 144 B => conflict-free fragment loads
constexpr int kMaxBlocks = 13;    // 16-row blocks per sequence: N <= 208

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragments (16 rows x 64 k) of a row-major shared tile: a[kk][0..3]
__device__ __forceinline__ void load_a_frags(const bf16* tile, int r0, int g, int t, uint32_t (*a)[4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = lds32(tile + (r0 + g) * kRS + kk * 16 + 2 * t);
    a[kk][1] = lds32(tile + (r0 + g + 8) * kRS + kk * 16 + 2 * t);
    a[kk][2] = lds32(tile + (r0 + g) * kRS + kk * 16 + 8 + 2 * t);
    a[kk][3] = lds32(tile + (r0 + g + 8) * kRS + kk * 16 + 8 + 2 * t);
  }
}
// four 8x8 b16 matrices, one 16-byte row address per lane (lanes 8j..8j+7 address the rows of matrix j); thread
// (g, t) receives elements (row g, columns 2t, 2t+1) of each: exactly the B-fragment words of mma.m16n8k16 when the
// matrix rows are the n index and its columns the (contiguous) k index.
__device__ __forceinline__ void ldmatrix_x4(uint32_t* r, const bf16* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// acc[2][4] (16 x 16) += A(16 x 64, fragments) . T[c0 .. c0+15][0..63]^T      (T row-major, contraction over its columns)
__device__ __forceinline__ void mma_rowmajor_b(float (*acc)[4], const uint32_t (*a)[4], const bf16* tile, int c0,
                                               int lane) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    // matrix j of the x4 load = T[c0+nt*8 .. +7][half*32 + 8j .. +7]: (b0, b1) of k-steps 2*half and 2*half+1
    const bf16* row = tile + (c0 + nt * 8 + (lane & 7)) * kRS + (lane >> 3) * 8;
    uint32_t b[8];
    ldmatrix_x4(b, row);
    ldmatrix_x4(b + 4, row + 32);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) mma16816(acc[nt], a[kk], b[2 * kk], b[2 * kk + 1]);
  }
}
// acc[8][4] (16 x 64) += A(16 x 16, one fragment) . Tt[0..63][c0 .. c0+15]^T   (Tt = transposed tile [64][ts])
__device__ __forceinline__ void mma_transposed_b(float (*acc)[4], const uint32_t* a, const bf16* tt, int ts, int c0,
                                                 int lane) {
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    // matrices: (rows dp*16 .. +7, cols c0 / c0+8), (rows dp*16+8 .. +15, cols c0 / c0+8): (b0, b1) of dt = 2dp, 2dp+1
    const bf16* row = tt + (dp * 16 + ((lane >> 4) << 3) + (lane & 7)) * ts + c0 + ((lane >> 3) & 1) * 8;
    uint32_t b[4];
    ldmatrix_x4(b, row);
    mma16816(acc[2 * dp], a, b[0], b[1]);
    mma16816(acc[2 * dp + 1], a, b[2], b[3]);
  }
}
__device__ __forceinline__ float ex2(float x) {       // 2^x, MUFU.EX2; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const bf16* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// acc[8][4] (16 x 64) += A(16 x 16 tokens, one fragment) . T[c0 .. c0+15][0..63]   (T row-major [token][channel])
__device__ __forceinline__ void mma_tokens_b(float (*acc)[4], const uint32_t* a, const bf16* tile, int c0, int lane) {
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    // stored 8x8 blocks (tokens x channels): (c0, 16dp), (c0+8, 16dp), (c0, 16dp+8), (c0+8, 16dp+8); transposed on load,
    // thread (g, t) gets (T[.. + 2t][.. + g], T[.. + 2t + 1][.. + g]) = the B-fragment word with k = token, n = channel
    const bf16* row = tile + (c0 + ((lane >> 3) & 1) * 8 + (lane & 7)) * kRS + dp * 16 + (lane >> 4) * 8;
    uint32_t b[4];
    ldmatrix_x4_trans(b, row);
    mma16816(acc[2 * dp], a, b[0], b[1]);
    mma16816(acc[2 * dp + 1], a, b[2], b[3]);
  }
}
// 16-byte chunk `ch` (8 bf16) of token row `tok` of one head: zero beyond the sequence
__device__ __forceinline__ uint4 load_chunk(const bf16* base, int64_t row_stride, int tok, int ch, int N) {
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (tok < N) v = *reinterpret_cast<const uint4*>(base + (int64_t)tok * row_stride + ch * 8);
  return v;
}
__device__ __forceinline__ void store_rowmajor(bf16* tile, int tok, int ch, const uint4& v) {
  *reinterpret_cast<uint4*>(tile + tok * kRS + ch * 8) = v;
}
__device__ __forceinline__ void store_transposed(bf16* tt, int ts, int tok, int ch, const uint4& v) {
  const bf16* e = reinterpret_cast<const bf16*>(&v);
#pragma unroll
  for (int i = 0; i < 8; ++i) tt[(ch * 8 + i) * ts + tok] = e[i];
}

// write a 16 x 64 accumulator tile (rows r0+g, r0+g+8) as bf16 to out[(tok) * row_stride + col]
__device__ __forceinline__ void store_tile(bf16* out, int64_t row_stride, int r0, int g, int t, int N, float (*acc)[4],
                                           float s0, float s1) {
  const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int col = dt * 8 + 2 * t;
    if (ra < N) *reinterpret_cast<uint32_t*>(out + (int64_t)ra * row_stride + col) = pack2(acc[dt][0] * s0, acc[dt][1] * s0);
    if (rb < N) *reinterpret_cast<uint32_t*>(out + (int64_t)rb * row_stride + col) = pack2(acc[dt][2] * s1, acc[dt][3] * s1);
  }
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(32 * kMaxBlocks, 2)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse, int N, int H, float c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nkb = (N + 15) >> 4, npad = nkb << 4;
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);          // [npad][kRS]
  bf16* Vs = Ks + npad * kRS;                            // [npad][kRS]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int64_t rs = 3 * (int64_t)H * kD;                // qkv row stride (elements)
  const bf16* qb = qkv + (int64_t)b * N * rs + h * kD;
  const bf16* kb = qb + H * kD;
  const bf16* vb = kb + H * kD;
  for (int idx = threadIdx.x; idx < npad * 8; idx += blockDim.x) {
    const int tok = idx >> 3, ch = idx & 7;                // 8 lanes = one 128-byte token row: coalesced
    store_rowmajor(Ks, tok, ch, load_chunk(kb, rs, tok, ch, N));
    store_rowmajor(Vs, tok, ch, load_chunk(vb, rs, tok, ch, N));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16;
  uint32_t qa[4][4];
  {
    const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = ra < N ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)ra * rs + kk * 16 + 2 * t) : 0u;
      qa[kk][1] = rb < N ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)rb * rs + kk * 16 + 2 * t) : 0u;
      qa[kk][2] = ra < N ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)ra * rs + kk * 16 + 8 + 2 * t) : 0u;
      qa[kk][3] = rb < N ? *reinterpret_cast<const uint32_t*>(qb + (int64_t)rb * rs + kk * 16 + 8 + 2 * t) : 0u;
    }
  }
  __syncthreads();
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float acc[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) acc[dt][0] = acc[dt][1] = acc[dt][2] = acc[dt][3] = 0.f;
  for (int jb = 0; jb < nkb; ++jb) {
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_rowmajor_b(s, qa, Ks, jb * 16, lane);
    if (jb == nkb - 1) {                                          // only the last block has columns beyond the sequence
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (jb * 16 + nt * 8 + 2 * t + (e & 1) >= N) s[nt][e] = -INFINITY;
    }
    // running maxima are kept on the raw scores (c > 0 preserves the order); the scale enters through one FFMA
    float x0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
    float x1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1));
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);          // finite from block 0 on (column 0 is always live)
    const float a0 = ex2((m0 - n0) * c), a1 = ex2((m1 - n1) * c);
    m0 = n0; m1 = n1;
    const float o0 = -n0 * c, o1 = -n1 * c;
    float p[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      p[nt][0] = ex2(fmaf(s[nt][0], c, o0)); p[nt][1] = ex2(fmaf(s[nt][1], c, o0));
      p[nt][2] = ex2(fmaf(s[nt][2], c, o1)); p[nt][3] = ex2(fmaf(s[nt][3], c, o1));
    }
    l0 = l0 * a0 + (p[0][0] + p[0][1] + p[1][0] + p[1][1]);
    l1 = l1 * a1 + (p[0][2] + p[0][3] + p[1][2] + p[1][3]);
    if (__any_sync(0xffffffffu, a0 != 1.f || a1 != 1.f)) {        // usually false after the first blocks
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { acc[dt][0] *= a0; acc[dt][1] *= a0; acc[dt][2] *= a1; acc[dt][3] *= a1; }
    }
    const uint32_t pa[4] = {pack2(p[0][0], p[0][1]), pack2(p[0][2], p[0][3]), pack2(p[1][0], p[1][1]), pack2(p[1][2], p[1][3])};
    mma_tokens_b(acc, pa, Vs, jb * 16, lane);                     // V row-major, transposed by ldmatrix
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  bf16* ob = o + (int64_t)b * N * H * kD + h * kD;
  store_tile(ob, (int64_t)H * kD, r0, g, t, N, acc, 1.f / l0, 1.f / l1);
  if (t == 0) {
    float* lb = lse + ((int64_t)b * H + h) * N;
    if (r0 + g < N) lb[r0 + g] = m0 * c + log2f(l0);
    if (r0 + g + 8 < N) lb[r0 + g + 8] = m1 * c + log2f(l1);
  }
}

// ----------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(32 * kMaxBlocks, 1)
attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                const float* __restrict__ lse, bf16* __restrict__ dqkv, int N, int H, float c, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nkb = (N + 15) >> 4, npad = nkb << 4, ts = npad + 8;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);          // row-major [npad][kRS]
  bf16* Ks = Qs + npad * kRS;
  bf16* Vs = Ks + npad * kRS;
  bf16* Gs = Vs + npad * kRS;                            // dO
  bf16* Qt = Gs + npad * kRS;                            // transposed [64][ts]
  bf16* Kt = Qt + kD * ts;
  bf16* Gt = Kt + kD * ts;
  float* Ls = reinterpret_cast<float*>(Gt + kD * ts);    // log2-sum-exp per query [npad]
  float* Ds = Ls + npad;                                 // rowsum(dO * O) per query [npad]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int64_t rs = 3 * (int64_t)H * kD, os = (int64_t)H * kD;
  const bf16* qb = qkv + (int64_t)b * N * rs + h * kD;
  const bf16* kb = qb + H * kD;
  const bf16* vb = kb + H * kD;
  const bf16* gb = d_o + (int64_t)b * N * os + h * kD;
  const bf16* ob = o + (int64_t)b * N * os + h * kD;
  for (int idx = threadIdx.x; idx < npad * 8; idx += blockDim.x) {
    const int tok = idx % npad, ch = idx / npad;
    const uint4 q = load_chunk(qb, rs, tok, ch, N), k = load_chunk(kb, rs, tok, ch, N);
    const uint4 v = load_chunk(vb, rs, tok, ch, N), gg = load_chunk(gb, os, tok, ch, N);
    store_rowmajor(Qs, tok, ch, q); store_transposed(Qt, ts, tok, ch, q);
    store_rowmajor(Ks, tok, ch, k); store_transposed(Kt, ts, tok, ch, k);
    store_rowmajor(Vs, tok, ch, v);
    store_rowmajor(Gs, tok, ch, gg); store_transposed(Gt, ts, tok, ch, gg);
  }
  // D[tok] = sum_d dO[tok][d] * O[tok][d]: two lanes per token, 32 channels each, fixed order
  for (int idx = threadIdx.x; idx < npad * 2; idx += blockDim.x) {
    const int tok = idx >> 1, half = idx & 1;
    float d = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const uint4 gg = load_chunk(gb, os, tok, half * 4 + ch, N), oo = load_chunk(ob, os, tok, half * 4 + ch, N);
      const bf16* ge = reinterpret_cast<const bf16*>(&gg);
      const bf16* oe = reinterpret_cast<const bf16*>(&oo);
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(__bfloat162float(ge[i]), __bfloat162float(oe[i]), d);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    if (half == 0) {
      Ds[tok] = d;
      Ls[tok] = tok < N ? lse[((int64_t)b * H + h) * N + tok] : 0.f;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16;
  bf16* dq_out = dqkv + (int64_t)b * N * rs + h * kD;
  bf16* dk_out = dq_out + H * kD;
  bf16* dv_out = dk_out + H * kD;

  {  // ---- pass A: this warp's 16 query rows -> dQ
    uint32_t qa[4][4], ga[4][4];
    load_a_frags(Qs, r0, g, t, qa);
    load_a_frags(Gs, r0, g, t, ga);
    const float L0 = Ls[r0 + g], L1 = Ls[r0 + g + 8], D0 = Ds[r0 + g], D1 = Ds[r0 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) dq[dt][0] = dq[dt][1] = dq[dt][2] = dq[dt][3] = 0.f;
    for (int jb = 0; jb < nkb; ++jb) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_rowmajor_b(s, qa, Ks, jb * 16, lane);
      mma_rowmajor_b(dp, ga, Vs, jb * 16, lane);
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float L = (e < 2) ? L0 : L1, D = (e < 2) ? D0 : D1;
          float p = ex2(fmaf(s[nt][e], c, -L));
          if (jb == nkb - 1 && jb * 16 + nt * 8 + 2 * t + (e & 1) >= N) p = 0.f;   // columns beyond the sequence
          ds[nt][e] = p * (dp[nt][e] - D) * scale;
        }
      const uint32_t da[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                              pack2(ds[1][2], ds[1][3])};
      mma_transposed_b(dq, da, Kt, ts, jb * 16, lane);
    }
    store_tile(dq_out, rs, r0, g, t, N, dq, 1.f, 1.f);
  }
  {  // ---- pass B: this warp's 16 key rows -> dK, dV   (everything transposed: rows = keys, columns = queries)
    uint32_t ka[4][4], va[4][4];
    load_a_frags(Ks, r0, g, t, ka);
    load_a_frags(Vs, r0, g, t, va);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
      dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
    }
    for (int ib = 0; ib < nkb; ++ib) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_rowmajor_b(st, ka, Qs, ib * 16, lane);
      mma_rowmajor_b(dpt, va, Gs, ib * 16, lane);
      float pt[2][4], dst[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = ib * 16 + nt * 8 + 2 * t + (e & 1);
          // queries beyond the sequence have dO = 0 and D = 0 (zero-filled tiles): their finite p contributes nothing
          const float p = ex2(fmaf(st[nt][e], c, -Ls[qi]));
          pt[nt][e] = p;
          dst[nt][e] = p * (dpt[nt][e] - Ds[qi]) * scale;
        }
      const uint32_t pa[4] = {pack2(pt[0][0], pt[0][1]), pack2(pt[0][2], pt[0][3]), pack2(pt[1][0], pt[1][1]),
                              pack2(pt[1][2], pt[1][3])};
      const uint32_t da[4] = {pack2(dst[0][0], dst[0][1]), pack2(dst[0][2], dst[0][3]), pack2(dst[1][0], dst[1][1]),
                              pack2(dst[1][2], dst[1][3])};
      mma_transposed_b(dv, pa, Gt, ts, ib * 16, lane);
      mma_transposed_b(dk, da, Qt, ts, ib * 16, lane);
    }
    store_tile(dk_out, rs, r0, g, t, N, dk, 1.f, 1.f);
    store_tile(dv_out, rs, r0, g, t, N, dv, 1.f, 1.f);
  }
}


// ------------------------------------------------------------------- backward, split form (default)
// Two kernels instead of one, each with only row-major tiles in shared memory (60 KB) and 7 warps per CTA (two
// CTAs per (image, head)), so three CTAs = 21 warps share an SM instead of one CTA of 13: the single-kernel form is
// latency-bound at 28 % issue utilisation (profiles/r01_attention_ncu_v2.txt).  Operands whose contraction index is
// the token index (K in dS K, dO in P^T dO, Q in dS^T Q) come straight from the row-major tiles through
// ldmatrix.trans -- no transposed staging copies, no 2-byte scatter stores.
// A fragments (16 rows x 64) straight from global memory (rows beyond the sequence: zero)
__device__ __forceinline__ void load_a_frags_global(const bf16* base, int64_t rs, int r0, int g, int t, int N,
                                                    uint32_t (*a)[4]) {
  const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = ra < N ? *reinterpret_cast<const uint32_t*>(base + (int64_t)ra * rs + kk * 16 + 2 * t) : 0u;
    a[kk][1] = rb < N ? *reinterpret_cast<const uint32_t*>(base + (int64_t)rb * rs + kk * 16 + 2 * t) : 0u;
    a[kk][2] = ra < N ? *reinterpret_cast<const uint32_t*>(base + (int64_t)ra * rs + kk * 16 + 8 + 2 * t) : 0u;
    a[kk][3] = rb < N ? *reinterpret_cast<const uint32_t*>(base + (int64_t)rb * rs + kk * 16 + 8 + 2 * t) : 0u;
  }
}
__device__ __forceinline__ float dot_bf16x2(uint32_t a, uint32_t b) {
  const float2 x = __bfloat1622float2(*reinterpret_cast<const bf162*>(&a));
  const float2 y = __bfloat1622float2(*reinterpret_cast<const bf162*>(&b));
  return fmaf(x.x, y.x, x.y * y.y);
}

constexpr int kSplitWarps = 7;       // warps per CTA of the split kernels; 2 CTAs cover the <= 13 row blocks

// dQ: warp = 16 query rows; K, V of the head in shared memory
__global__ void __maxnreg__(96)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                   const float* __restrict__ lse, bf16* __restrict__ dqkv, int N, int H, float c, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nkb = (N + 15) >> 4, npad = nkb << 4;
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + npad * kRS;
  const int bh = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int b = bh / H, h = bh % H;
  const int64_t rs = 3 * (int64_t)H * kD, os = (int64_t)H * kD;
  const bf16* qb = qkv + (int64_t)b * N * rs + h * kD;
  const bf16* kb = qb + H * kD;
  const bf16* vb = kb + H * kD;
  const bf16* gb = d_o + (int64_t)b * N * os + h * kD;
  const bf16* ob = o + (int64_t)b * N * os + h * kD;
  for (int idx = threadIdx.x; idx < npad * 8; idx += blockDim.x) {
    const int tok = idx >> 3, ch = idx & 7;                 // 8 lanes = one 128-byte token row: coalesced
    store_rowmajor(Ks, tok, ch, load_chunk(kb, rs, tok, ch, N));
    store_rowmajor(Vs, tok, ch, load_chunk(vb, rs, tok, ch, N));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rbk = half * kSplitWarps + warp;                // this warp's block of 16 query rows
  const int r0 = rbk * 16;
  uint32_t qa[4][4], ga[4][4];
  float D0 = 0.f, D1 = 0.f, L0 = 0.f, L1 = 0.f;
  if (rbk < nkb) {
    load_a_frags_global(qb, rs, r0, g, t, N, qa);
    load_a_frags_global(gb, os, r0, g, t, N, ga);
    uint32_t oa[4][4];
    load_a_frags_global(ob, os, r0, g, t, N, oa);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      D0 += dot_bf16x2(ga[kk][0], oa[kk][0]) + dot_bf16x2(ga[kk][2], oa[kk][2]);
      D1 += dot_bf16x2(ga[kk][1], oa[kk][1]) + dot_bf16x2(ga[kk][3], oa[kk][3]);
    }
    const float* lb = lse + ((int64_t)b * H + h) * N;
    if (r0 + g < N) L0 = lb[r0 + g];
    if (r0 + g + 8 < N) L1 = lb[r0 + g + 8];
  }
  D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
  D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  __syncthreads();
  if (rbk >= nkb) return;
  float dq[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) dq[dt][0] = dq[dt][1] = dq[dt][2] = dq[dt][3] = 0.f;
  for (int jb = 0; jb < nkb; ++jb) {
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_rowmajor_b(s, qa, Ks, jb * 16, lane);
    mma_rowmajor_b(dp, ga, Vs, jb * 16, lane);
    float ds[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float L = (e < 2) ? L0 : L1, D = (e < 2) ? D0 : D1;
        float p = ex2(fmaf(s[nt][e], c, -L));
        if (jb == nkb - 1 && jb * 16 + nt * 8 + 2 * t + (e & 1) >= N) p = 0.f;   // columns beyond the sequence
        ds[nt][e] = p * (dp[nt][e] - D) * scale;
      }
    const uint32_t da[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]), pack2(ds[1][0], ds[1][1]),
                            pack2(ds[1][2], ds[1][3])};
    mma_tokens_b(dq, da, Ks, jb * 16, lane);
  }
  store_tile(dqkv + (int64_t)b * N * rs + h * kD, rs, r0, g, t, N, dq, 1.f, 1.f);
}

// dK, dV: warp = 16 key rows; Q, dO of the head (+ log-sum-exp, D per query) in shared memory
__device__ __forceinline__ void attn_bwd_dkv_body(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                  const bf16* __restrict__ d_o, const float* __restrict__ lse,
                                                  bf16* __restrict__ dqkv, int N, int H, float c, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nkb = (N + 15) >> 4, npad = nkb << 4;
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Gs = Qs + npad * kRS;
  float* Ls = reinterpret_cast<float*>(Gs + npad * kRS);
  float* Ds = Ls + npad;
  const int bh = blockIdx.x >> 1, half = blockIdx.x & 1;
  const int b = bh / H, h = bh % H;
  const int64_t rs = 3 * (int64_t)H * kD, os = (int64_t)H * kD;
  const bf16* qb = qkv + (int64_t)b * N * rs + h * kD;
  const bf16* kb = qb + H * kD;
  const bf16* vb = kb + H * kD;
  const bf16* gb = d_o + (int64_t)b * N * os + h * kD;
  const bf16* ob = o + (int64_t)b * N * os + h * kD;
  for (int idx = threadIdx.x; idx < npad * 8; idx += blockDim.x) {
    const int tok = idx >> 3, ch = idx & 7;
    const uint4 gg = load_chunk(gb, os, tok, ch, N), oo = load_chunk(ob, os, tok, ch, N);
    store_rowmajor(Qs, tok, ch, load_chunk(qb, rs, tok, ch, N));
    store_rowmajor(Gs, tok, ch, gg);
    // D[tok] = sum_d dO * O: the 8 chunks of a token are 8 consecutive lanes (npad * 8 and blockDim are multiples of 32)
    const bf16* ge = reinterpret_cast<const bf16*>(&gg);
    const bf16* oe = reinterpret_cast<const bf16*>(&oo);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) d = fmaf(__bfloat162float(ge[i]), __bfloat162float(oe[i]), d);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (ch == 0) {
      Ds[tok] = d;
      Ls[tok] = tok < N ? lse[((int64_t)b * H + h) * N + tok] : 0.f;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rbk = half * kSplitWarps + warp;                // this warp's block of 16 key rows
  const int r0 = rbk * 16;
  uint32_t ka[4][4], va[4][4];
  if (rbk < nkb) {
    load_a_frags_global(kb, rs, r0, g, t, N, ka);
    load_a_frags_global(vb, rs, r0, g, t, N, va);
  }
  __syncthreads();
  if (rbk >= nkb) return;
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
    dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
  }
  for (int ib = 0; ib < nkb; ++ib) {
    float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_rowmajor_b(st, ka, Qs, ib * 16, lane);
    mma_rowmajor_b(dpt, va, Gs, ib * 16, lane);
    uint32_t pa[4], da[4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float pt[4], dst[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = ib * 16 + nt * 8 + 2 * t + (e & 1);
        // queries beyond the sequence have dO = 0 and D = 0 (zero-filled tiles): their finite p contributes nothing
        pt[e] = ex2(fmaf(st[nt][e], c, -Ls[qi]));
        dst[e] = pt[e] * (dpt[nt][e] - Ds[qi]) * scale;
      }
      pa[2 * nt] = pack2(pt[0], pt[1]); pa[2 * nt + 1] = pack2(pt[2], pt[3]);
      da[2 * nt] = pack2(dst[0], dst[1]); da[2 * nt + 1] = pack2(dst[2], dst[3]);
    }
    mma_tokens_b(dv, pa, Gs, ib * 16, lane);
    mma_tokens_b(dk, da, Qs, ib * 16, lane);
  }
  bf16* dk_out = dqkv + (int64_t)b * N * rs + H * kD + h * kD;
  store_tile(dk_out, rs, r0, g, t, N, dk, 1.f, 1.f);
  store_tile(dk_out + H * kD, rs, r0, g, t, N, dv, 1.f, 1.f);
}

// two register budgets of the same body: 96 (three CTAs = 21 warps per SM, spills) / 128 (two CTAs, none; default)
__global__ void __maxnreg__(96)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                    const float* __restrict__ lse, bf16* __restrict__ dqkv, int N, int H, float c, float scale) {
  attn_bwd_dkv_body(qkv, o, d_o, lse, dqkv, N, H, c, scale);
}
__global__ void __maxnreg__(128)
attn_bwd_dkv128_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                       const float* __restrict__ lse, bf16* __restrict__ dqkv, int N, int H, float c, float scale) {
  attn_bwd_dkv_body(qkv, o, d_o, lse, dqkv, N, H, c, scale);
}

size_t split_smem(int N, bool with_rows) {
  const int npad = (N + 15) / 16 * 16;
  return (size_t)(2 * npad * kRS) * sizeof(bf16) + (with_rows ? 2 * npad * sizeof(float) : 0);
}

size_t fwd_smem(int N) {
  const int npad = (N + 15) / 16 * 16;
  return (size_t)(2 * npad * kRS) * sizeof(bf16);
}
size_t bwd_smem(int N) {
  const int npad = (N + 15) / 16 * 16;
  return (size_t)(4 * npad * kRS + 3 * kD * (npad + 8)) * sizeof(bf16) + 2 * npad * sizeof(float);
}

}  // namespace

extern "C" {

int b200at_attn_fwd(const void* qkv, void* o, float* lse, int64_t B, int64_t N, int64_t H, float scale, void* stream) {
  if (B <= 0) return 0;
  if (N <= 0 || N > 16 * kMaxBlocks || H <= 0) return (int)cudaErrorInvalidValue;
  const size_t smem = fwd_smem((int)N);
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int threads = 32 * (int)((N + 15) / 16);
  attn_fwd_kernel<<<(unsigned)(B * H), threads, smem, (cudaStream_t)stream>>>(
      (const bf16*)qkv, (bf16*)o, lse, (int)N, (int)H, scale * 1.4426950408889634f);
  return (int)cudaGetLastError();
}

int b200at_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int64_t B, int64_t N,
                    int64_t H, float scale, void* stream) {
  if (B <= 0) return 0;
  if (N <= 0 || N > 16 * kMaxBlocks || H <= 0) return (int)cudaErrorInvalidValue;
  static const bool single = []() { const char* v = getenv("B200AT_ATTN_BWD"); return v && v[0] == '1'; }();
  if (!single) {
    const float c2 = scale * 1.4426950408889634f;
    const size_t sa = split_smem((int)N, false), sb = split_smem((int)N, true);
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa);
    if (e1 != cudaSuccess) return (int)e1;
    // measured (profiles/r01_vit_bench_v4.txt): 128 registers / 2 CTAs per SM 363 us, 96 registers / 3 CTAs 484 us (spills)
    static const bool dkv128 = []() { const char* v = getenv("B200AT_ATTN_DKV_REGS"); return !(v && atoi(v) == 96); }();
    auto dkv = dkv128 ? attn_bwd_dkv128_kernel : attn_bwd_dkv_kernel;
    e1 = cudaFuncSetAttribute(dkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
    if (e1 != cudaSuccess) return (int)e1;
    attn_bwd_dq_kernel<<<(unsigned)(2 * B * H), 32 * kSplitWarps, sa, (cudaStream_t)stream>>>(
        (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, (bf16*)dqkv, (int)N, (int)H, c2, scale);
    dkv<<<(unsigned)(2 * B * H), 32 * kSplitWarps, sb, (cudaStream_t)stream>>>(
        (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, (bf16*)dqkv, (int)N, (int)H, c2, scale);
    return (int)cudaGetLastError();
  }
  const size_t smem = bwd_smem((int)N);
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int threads = 32 * (int)((N + 15) / 16);
  attn_bwd_kernel<<<(unsigned)(B * H), threads, smem, (cudaStream_t)stream>>>(
      (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, (bf16*)dqkv, (int)N, (int)H,
      scale * 1.4426950408889634f, scale);
  return (int)cudaGetLastError();
}

}  // extern "C"
