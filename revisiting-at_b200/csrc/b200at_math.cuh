// Per-element / per-sample arithmetic of the APGD attack step, shared by the sm_100a kernels
// (b200at_attack.cu) and by the host-compiled checker in tests/hostcheck/ (test infrastructure:
// it lets the CPU test-suite run these exact bodies against the oracle without a GPU).
//
// Contract (reference autopgd_train_clean.py:213-226): every operation is rounded
// to fp32 on its own -- no FMA contraction.  Device code uses the __f*_rn intrinsics (never
// contracted) and the TU is additionally built with -fmad=false; host code is built with
// -ffp-contract=off.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define B200AT_HD __host__ __device__ __forceinline__
#define B200AT_ADD(a, b) __fadd_rn((a), (b))
#define B200AT_SUB(a, b) __fsub_rn((a), (b))
#define B200AT_MUL(a, b) __fmul_rn((a), (b))
#define B200AT_DIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define B200AT_HD __host__ __device__ __forceinline__
#define B200AT_ADD(a, b) ((a) + (b))
#define B200AT_SUB(a, b) ((a) - (b))
#define B200AT_MUL(a, b) ((a) * (b))
#define B200AT_DIV(a, b) ((a) / (b))
#else
#define B200AT_HD static inline
#define B200AT_ADD(a, b) ((a) + (b))
#define B200AT_SUB(a, b) ((a) - (b))
#define B200AT_MUL(a, b) ((a) * (b))
#define B200AT_DIV(a, b) ((a) / (b))
#endif

// ---- per-sample state block: float rows [B200AT_ST_ROWS][B]; integer rows are bit-cast int32 ----
enum {
  B200AT_ST_STEP = 0,            // step size (autopgd_train_clean.py:169)
  B200AT_ST_LOSS_BEST = 1,       // :199
  B200AT_ST_LOSS_BEST_LAST = 2,  // loss_best_last_check :200
  B200AT_ST_REDUCED_LAST = 3,    // reduced_last_check :201
  B200AT_ST_ACC = 4,             // int32 0/1, robust-so-far mask :197,296
  B200AT_ST_FLAGS = 5,           // int32 pending image ops for the next pass (B200AT_F_*)
  B200AT_ST_LOSS_CUR = 6,        // loss at the current iterate
  B200AT_ST_TOPK = 7,            // l1: fraction of coordinates moved :163,355
  B200AT_ST_SP_OLD = 8,          // l1: sparsity at the previous checkpoint :164,359
  B200AT_ST_SP_BEST = 9,         // int32 l1: nnz(x_best - x)
  B200AT_ST_SP_ADV = 10,         // int32 l1: nnz(x_adv - x), written by the step kernel
  B200AT_ST_PRED = 11,           // int32 0/1 prediction correct at the current iterate
  // iterate-log protocol (small n_iter): every iterate / gradient stays in its own slot and the
  // image ops of :304,:321-324,:345-346 become per-sample slot indices (int32) instead of copies
  B200AT_ST_IDX_CUR = 12,        // slot holding this sample's x_adv (differs from the newest slot after a restore)
  B200AT_ST_IDX_OLD = 13,        // slot holding x_adv_old
  B200AT_ST_IDX_BEST = 14,       // slot holding x_best
  B200AT_ST_IDX_BEST_ADV = 15,   // slot holding x_best_adv
  B200AT_ST_GIDX_CUR = 16,       // slot holding grad
  B200AT_ST_GIDX_BEST = 17,      // slot holding grad_best
  B200AT_ST_ROWS = 20
};
#define B200AT_LOG_MAX_SLOTS 8
enum {
  B200AT_F_IMPROVED = 1,   // x_best <- x_adv, grad_best <- grad   (:321-324)
  B200AT_F_WRITE_ADV = 2,  // x_best_adv <- x_adv                  (:304)
  B200AT_F_RESTORE = 4     // x_adv <- x_best, grad <- grad_best   (:345-346, :361-362)
};
enum { B200AT_NORM_LINF = 0, B200AT_NORM_L2 = 1, B200AT_NORM_L1 = 2 };

B200AT_HD int32_t b200at_f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
B200AT_HD float b200at_i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

// torch.max / torch.min / clamp semantics: NaN propagates.  On the device this is the single
// FMNMX.NAN instruction (PTX max.NaN / min.NaN); the update kernels are close to issue-bound otherwise.
#if defined(__CUDA_ARCH__)
B200AT_HD float b200at_max(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
B200AT_HD float b200at_min(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
#else
B200AT_HD float b200at_max(float a, float b) { return (a != a || b != b) ? (a + b) : (a < b ? b : a); }
B200AT_HD float b200at_min(float a, float b) { return (a != a || b != b) ? (a + b) : (b < a ? b : a); }
#endif
B200AT_HD float b200at_clamp01(float v) { return b200at_min(b200at_max(v, 0.0f), 1.0f); }
// torch.sign: 0 for +-0 and NaN
B200AT_HD float b200at_sign(float g) { return (float)((g > 0.0f) - (g < 0.0f)); }

// l-inf move with momentum (:214-226).  xc = current iterate, xo = previous iterate.
B200AT_HD float b200at_linf_elem(float x, float xc, float xo, float g, float step, float eps, float a,
                                 float one_minus_a) {
  const float lo = B200AT_SUB(x, eps), hi = B200AT_ADD(x, eps);
  const float g2 = B200AT_SUB(xc, xo);
  float z = B200AT_ADD(xc, B200AT_MUL(step, b200at_sign(g)));
  z = b200at_clamp01(b200at_min(b200at_max(z, lo), hi));
  float w = B200AT_ADD(xc, B200AT_MUL(B200AT_SUB(z, xc), a));
  w = B200AT_ADD(w, B200AT_MUL(g2, one_minus_a));
  return b200at_clamp01(b200at_min(b200at_max(w, lo), hi));
}

// Per-sample bookkeeping after a forward pass (:194-205 for iter < 0, :291-364 otherwise).
// `loss_steps` is [n_iter][B].  ckpt_k > 0 marks a checkpoint iteration with window k.
// step_full = float32(alpha*eps), step_min = float32(alpha*eps/10) (l1 only).
// has_grad: a new gradient is being computed at this iterate (false on the last iteration, :281-283).
B200AT_HD void b200at_bookkeep_sample(float* st, float* loss_steps, int B, int b, float loss, int pred,
                                      int iter, int n_iter, int ckpt_k, int norm_kind, float step_full,
                                      float step_min, float n_fts, int has_grad) {
  float* step = st + (int64_t)B200AT_ST_STEP * B + b;
  float* loss_best = st + (int64_t)B200AT_ST_LOSS_BEST * B + b;
  float* loss_best_last = st + (int64_t)B200AT_ST_LOSS_BEST_LAST * B + b;
  float* reduced_last = st + (int64_t)B200AT_ST_REDUCED_LAST * B + b;
  float* acc = st + (int64_t)B200AT_ST_ACC * B + b;
  float* flags = st + (int64_t)B200AT_ST_FLAGS * B + b;
  float* sp_best = st + (int64_t)B200AT_ST_SP_BEST * B + b;
  const float* sp_adv = st + (int64_t)B200AT_ST_SP_ADV * B + b;
  st[(int64_t)B200AT_ST_LOSS_CUR * B + b] = loss;
  st[(int64_t)B200AT_ST_PRED * B + b] = b200at_i2f(pred);
  float* idx_cur = st + (int64_t)B200AT_ST_IDX_CUR * B + b;
  float* idx_old = st + (int64_t)B200AT_ST_IDX_OLD * B + b;
  float* idx_best = st + (int64_t)B200AT_ST_IDX_BEST * B + b;
  float* idx_best_adv = st + (int64_t)B200AT_ST_IDX_BEST_ADV * B + b;
  float* gidx_cur = st + (int64_t)B200AT_ST_GIDX_CUR * B + b;
  float* gidx_best = st + (int64_t)B200AT_ST_GIDX_BEST * B + b;
  // the iterate just evaluated lives in slot iter+1; the move that produced it read x_adv from idx_cur
  const float slot = b200at_i2f(iter + 1);
  if (iter >= 0) *idx_old = *idx_cur;
  *idx_cur = slot;
  if (has_grad) *gidx_cur = slot;

  if (iter < 0) {
    *idx_old = slot; *idx_best = slot; *idx_best_adv = slot; *gidx_best = slot;  // first evaluation: everything "improves" so the next pass seeds x_best/grad_best/x_best_adv
    *acc = b200at_i2f(pred);
    *loss_best = loss;
    *loss_best_last = loss;
    *reduced_last = 1.0f;
    *sp_best = *sp_adv;
    *flags = b200at_i2f(B200AT_F_IMPROVED | B200AT_F_WRITE_ADV);
    return;
  }
  int32_t fl = 0;
  *acc = b200at_i2f(b200at_f2i(*acc) & pred);
  if (!pred) { fl |= B200AT_F_WRITE_ADV; *idx_best_adv = slot; }
  loss_steps[(int64_t)iter * B + b] = loss;
  if (loss > *loss_best) {  // strict; false for NaN
    fl |= B200AT_F_IMPROVED;
    *loss_best = loss;
    *sp_best = *sp_adv;
    *idx_best = slot;
    *gidx_best = *gidx_cur;  // the gradient just computed, or the stale one on the last iteration (:323)
  }
  if (ckpt_k > 0) {
    if (norm_kind != B200AT_NORM_L1) {
      float ups = 0.0f;
      for (int c = 0; c < ckpt_k; ++c) {  // Python-style negative row wrap (:119)
        int r1 = (iter - c) % n_iter, r0 = (iter - c - 1) % n_iter;
        if (r1 < 0) r1 += n_iter;
        if (r0 < 0) r0 += n_iter;
        ups += (loss_steps[(int64_t)r1 * B + b] > loss_steps[(int64_t)r0 * B + b]) ? 1.0f : 0.0f;
      }
      const float thr = (float)((double)ckpt_k * 0.75);
      const float osc = (ups <= thr) ? 1.0f : 0.0f;
      const float stalled = B200AT_MUL(B200AT_SUB(1.0f, *reduced_last), (*loss_best_last >= *loss_best) ? 1.0f : 0.0f);
      const float f = b200at_max(osc, stalled);
      *reduced_last = f;
      *loss_best_last = *loss_best;
      if (f > 0.0f) {
        *step = B200AT_DIV(*step, 2.0f);
        fl |= B200AT_F_RESTORE;
      }
    } else {  // l1 sparsity adaptation (:351-364)
      float* topk = st + (int64_t)B200AT_ST_TOPK * B + b;
      float* sp_old = st + (int64_t)B200AT_ST_SP_OLD * B + b;
      const float sp = (float)b200at_f2i(*sp_best);
      const int red = B200AT_DIV(sp, *sp_old) < 0.95f;
      *topk = B200AT_DIV(B200AT_DIV(sp, n_fts), 1.5f);
      float s = red ? step_full : B200AT_DIV(*step, 1.5f);
      s = b200at_min(b200at_max(s, step_min), step_full);
      *step = s;
      *sp_old = sp;
      if (red) fl |= B200AT_F_RESTORE;
    }
  }
  if (fl & B200AT_F_RESTORE) { *idx_cur = *idx_best; *gidx_cur = *gidx_best; }
  *flags = b200at_i2f(fl);
}

// ------------------------------------------------------------------------------------------------
// l2 move (autopgd_train_clean.py:228-237).  Three dependent per-sample norms:
// ||g||, ||z - x||, ||w - x||; each phase recomputes the elementwise chain up to its reduction.
// Division by a per-sample constant c = norm + 1e-12 whose correctly rounded reciprocal rc is computed once per sample:
// q0 = RN(d rc), e = d - c q0 (exact, one FMA), q = RN(q0 + e rc) -- Markstein's sequence, equal to the IEEE quotient
// except for a last-place difference in rare cases (the l2 path is tolerance-checked at 1e-6: the norms themselves depend
// on the summation order).  3 instructions instead of the ~20 of __fdiv_rn, 6 divisions per element and iteration.
#if defined(__CUDA_ARCH__)
B200AT_HD float b200at_div_by(float d, float c, float rc) {
  const float q0 = __fmul_rn(d, rc);
  const float e = __fmaf_rn(-c, q0, d);
  return __fmaf_rn(e, rc, q0);
}
B200AT_HD float b200at_rcp(float c) { return __frcp_rn(c); }
#else
B200AT_HD float b200at_div_by(float d, float c, float rc) { (void)rc; return d / c; }
B200AT_HD float b200at_rcp(float c) { return 1.0f / c; }
#endif
B200AT_HD float b200at_l2_z(float xc, float g, float step, float gden, float grcp) {
  return B200AT_ADD(xc, b200at_div_by(B200AT_MUL(step, g), gden, grcp));
}
// clamp(x + d / (||d|| + 1e-12) * min(eps, ||d||), 0, 1);  den = ||d|| + 1e-12, rcp = 1 / den
B200AT_HD float b200at_l2_ball(float x, float d, float nrm, float den, float rcp, float eps) {
  const float q = b200at_div_by(d, den, rcp);
  return b200at_clamp01(B200AT_ADD(x, B200AT_MUL(q, b200at_min(eps, nrm))));
}
B200AT_HD float b200at_momentum(float xc, float z, float xo, float a, float one_minus_a) {
  const float w = B200AT_ADD(xc, B200AT_MUL(B200AT_SUB(z, xc), a));
  return B200AT_ADD(w, B200AT_MUL(B200AT_SUB(xc, xo), one_minus_a));
}
// the three norms of a sample with their denominators (norm + 1e-12) and reciprocals
struct B200atL2Norms {
  float n[3], den[3], rcp[3];
};
template <int PHASE>
B200AT_HD B200atL2Norms b200at_l2_norms(const float* sums) {
  B200atL2Norms r;
  for (int i = 0; i < 3; ++i) {
    r.n[i] = i < PHASE ? sqrtf(sums[i]) : 0.0f;
    r.den[i] = B200AT_ADD(r.n[i], 1e-12f);
    r.rcp[i] = i < PHASE ? b200at_rcp(r.den[i]) : 0.0f;
  }
  return r;
}
// value whose square is accumulated in phase 0/1/2, or the new iterate in phase 3
template <int PHASE>
B200AT_HD float b200at_l2_elem(float x, float xc, float xo, float g, float step, float eps, float a,
                               float one_minus_a, const B200atL2Norms& nm) {
  if (PHASE == 0) return g;
  const float d1 = B200AT_SUB(b200at_l2_z(xc, g, step, nm.den[0], nm.rcp[0]), x);
  if (PHASE == 1) return d1;
  const float z1 = b200at_l2_ball(x, d1, nm.n[1], nm.den[1], nm.rcp[1], eps);
  const float d2 = B200AT_SUB(b200at_momentum(xc, z1, xo, a, one_minus_a), x);
  if (PHASE == 2) return d2;
  return b200at_l2_ball(x, d2, nm.n[2], nm.den[2], nm.rcp[2], eps);
}

// ------------------------------------------------------------------------------------------------
// l1 move (autopgd_train_clean.py:239-250) + projection onto {||.||_1 <= eps} ∩ box (:24-91).
//   thr  = k-th smallest |grad| (exact order statistic, radix select on the fp32 bit pattern)
//   y    = (xc + step * s / (nnz + 1e-10)) - x,  s = sign(grad) where |grad| >= thr else 0
//   a    = box excess of x+y (>= 0), b = |y|;  if sum(a) + eps - sum(b) < 0 the water level alpha with
//          sum_i clamp(alpha, a_i, b_i) = sum(b) - eps is found by sectioning alpha's bit pattern
//   out  = (x + y) + sign(y) * d,  d = -clamp(alpha, a, b)  (or -a when no l1 shrink is needed)
B200AT_HD uint32_t b200at_l1_key(float g) { return (uint32_t)b200at_f2i(g) & 0x7fffffffu; }

B200AT_HD int64_t b200at_l1_rank(float topk, int64_t n) {  // clamp((1-topk)*n, 0, n-1).long()  (:241)
  float r = B200AT_MUL(B200AT_SUB(1.0f, topk), (float)n);
  r = b200at_min(b200at_max(r, 0.0f), (float)(n - 1));
  return (int64_t)r;
}

B200AT_HD float b200at_l1_y(float x, float xc, float g, float step, float thr, float nnz) {
  const float ag = fabsf(g);
  const float s = (ag >= thr) ? b200at_sign(g) : 0.0f;
  const float moved = B200AT_ADD(xc, B200AT_DIV(B200AT_MUL(step, s), B200AT_ADD(nnz, 1e-10f)));
  return B200AT_SUB(moved, x);
}
// u = min(0, min(1 - x - y, x + y)) <= 0  (:37-39);  a = -u, b = |y|
B200AT_HD float b200at_l1_u(float x, float y) {
  const float t = b200at_min(B200AT_SUB(B200AT_SUB(1.0f, x), y), B200AT_ADD(x, y));
  return b200at_min(0.0f, t);
}
B200AT_HD float b200at_l1_level(float alpha, float a, float b) { return b200at_min(b200at_max(a, alpha), b); }
B200AT_HD float b200at_l1_out(float x, float y, float u, int need, float alpha) {
  const float d = need ? -b200at_l1_level(alpha, -u, fabsf(y)) : u;
  return B200AT_ADD(B200AT_ADD(x, y), B200AT_MUL(b200at_sign(y), d));
}
// sectioning of alpha's bit pattern: 31 value bits, 5 per pass (shifts 26,21,...,1) then the last bit
#define B200AT_L1_PASSES 7
B200AT_HD int b200at_l1_shift(int pass) { return pass < 6 ? 26 - 5 * pass : 0; }
B200AT_HD int b200at_l1_ncand(int pass) { return pass < 6 ? 31 : 1; }
B200AT_HD float b200at_l1_cand(uint32_t prefix, int k, int pass) {
  return b200at_i2f((int32_t)(prefix | ((uint32_t)k << b200at_l1_shift(pass))));
}
