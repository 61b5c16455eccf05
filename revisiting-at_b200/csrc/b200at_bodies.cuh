// Kernel bodies of the image-sized APGD passes, one call per VEC consecutive elements.
// The sm_100a kernels (b200at_attack.cu) call these from their grid loops; tests/hostcheck
// compiles the same bodies for the host to check indexing / flag handling against the oracle.
#pragma once
#include "b200at_math.cuh"

template <int VEC>
struct B200atVec {
  float v[VEC];
};

#if defined(__CUDA_ARCH__)
// streaming (evict-first) 16-byte accesses for data that is touched once per pass
template <int VEC>
__device__ __forceinline__ B200atVec<VEC> b200at_ld_stream(const float* p) {
  B200atVec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __ldcs(p + i);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void b200at_st_stream(float* p, const B200atVec<VEC>& r) {
  if constexpr (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) __stcs(p + i, r.v[i]);
  }
}
// L2-resident (ld.global.cg: no L1 allocation, normal L2 eviction priority) 16-byte loads for operands that a later
// phase of the SAME kernel reads again (the single-launch l2 move)
template <int VEC>
__device__ __forceinline__ B200atVec<VEC> b200at_ld_keep(const float* p) {
  B200atVec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __ldcg(p + i);
  }
  return r;
}
// default-policy store: the new iterate is read next by the model's first layer, keep it in L2
template <int VEC>
__device__ __forceinline__ void b200at_st_keep(float* p, const B200atVec<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
  }
}
#else
template <int VEC>
B200AT_HD B200atVec<VEC> b200at_ld_stream(const float* p) {
  B200atVec<VEC> r;
  for (int i = 0; i < VEC; ++i) r.v[i] = p[i];
  return r;
}
template <int VEC>
B200AT_HD B200atVec<VEC> b200at_ld_keep(const float* p) { return b200at_ld_stream<VEC>(p); }
template <int VEC>
B200AT_HD void b200at_st_stream(float* p, const B200atVec<VEC>& r) {
  for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
}
template <int VEC>
B200AT_HD void b200at_st_keep(float* p, const B200atVec<VEC>& r) {
  for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
}
#endif

// Image buffers of one attack call, all [B][n] fp32 in the same dense layout.
struct B200atImages {
  const float* x;      // clean input (never written)
  float* x_adv;        // current iterate (rewritten only where a restore is pending)
  const float* x_old;  // previous iterate   (may alias x_adv on the first move, or x_new)
  float* x_new;        // next iterate       (may alias x_old: each element is read before it is written)
  const float* grad;   // dL/dx at x_adv
  float* x_best;
  float* grad_best;
  float* x_best_adv;
  const float* st;     // per-sample state rows (b200at_math.cuh)
  int64_t B, n;
};

// ---- K1: fused l-inf update + pending best/adv/restore image ops (autopgd_train_clean.py:213-226,
// :304, :321-324, :345-346).  Algorithmic traffic 20 B/element (+4 per pending predicated write).
template <int VEC>
B200AT_HD void b200at_linf_body(const B200atImages& p, int64_t vi, float eps, float a, float one_minus_a) {
  const int64_t e = vi * VEC;
  const int b = (int)(e / p.n);
  const int32_t fl = b200at_f2i(p.st[(int64_t)B200AT_ST_FLAGS * p.B + b]);
  const float step = p.st[(int64_t)B200AT_ST_STEP * p.B + b];
  const bool improved = fl & B200AT_F_IMPROVED;
  const bool write_adv = fl & B200AT_F_WRITE_ADV;
  const bool restore = (fl & B200AT_F_RESTORE) && !improved;  // improved => x_best == x_adv already

  const B200atVec<VEC> x = b200at_ld_stream<VEC>(p.x + e);
  const B200atVec<VEC> xo = b200at_ld_stream<VEC>(p.x_old + e);
  B200atVec<VEC> xc, g;
  if (!restore) {
    xc = b200at_ld_stream<VEC>(p.x_adv + e);
    g = b200at_ld_stream<VEC>(p.grad + e);
    if (write_adv) b200at_st_stream<VEC>(p.x_best_adv + e, xc);
    if (improved) {
      b200at_st_stream<VEC>(p.x_best + e, xc);
      b200at_st_stream<VEC>(p.grad_best + e, g);
    }
  } else {
    if (write_adv) b200at_st_stream<VEC>(p.x_best_adv + e, b200at_ld_stream<VEC>(p.x_adv + e));
    xc = b200at_ld_stream<VEC>(p.x_best + e);
    g = b200at_ld_stream<VEC>(p.grad_best + e);
    b200at_st_stream<VEC>(p.x_adv + e, xc);  // becomes x_old of the next move
  }
  B200atVec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) o.v[i] = b200at_linf_elem(x.v[i], xc.v[i], xo.v[i], g.v[i], step, eps, a, one_minus_a);
  b200at_st_keep<VEC>(p.x_new + e, o);
}

// ---- final pass: apply the pending x_best / x_best_adv writes after the last forward ----
template <int VEC>
B200AT_HD void b200at_flush_body(const B200atImages& p, int64_t vi) {
  const int64_t e = vi * VEC;
  const int b = (int)(e / p.n);
  const int32_t fl = b200at_f2i(p.st[(int64_t)B200AT_ST_FLAGS * p.B + b]);
  if (!(fl & (B200AT_F_IMPROVED | B200AT_F_WRITE_ADV))) return;
  const B200atVec<VEC> xc = b200at_ld_stream<VEC>(p.x_adv + e);
  if (fl & B200AT_F_WRITE_ADV) b200at_st_stream<VEC>(p.x_best_adv + e, xc);
  if (fl & B200AT_F_IMPROVED) b200at_st_stream<VEC>(p.x_best + e, xc);
}

// ---- entry pass: x_adv = clamp(x, 0, 1) (:141); returns nnz(x_adv - x) of this vector (l1 state) ----
template <int VEC>
B200AT_HD int b200at_init_body(const float* x, float* x_adv, int64_t vi) {
  const int64_t e = vi * VEC;
  const B200atVec<VEC> xi = b200at_ld_stream<VEC>(x + e);
  B200atVec<VEC> o;
  int nnz = 0;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    o.v[i] = b200at_clamp01(xi.v[i]);
    nnz += (B200AT_SUB(o.v[i], xi.v[i]) != 0.0f);
  }
  b200at_st_keep<VEC>(x_adv + e, o);
  return nnz;
}

// ---- FGSM random start (fgsm_train.py:79-83): x_adv = x + (2t-1)*eps*noise_level [, clamp to [0,1]] ----
template <int VEC>
B200AT_HD void b200at_fgsm_start_body(const float* x, const float* noise, float* x_adv, int64_t vi, float eps,
                                      float noise_level, int skip_projection) {
  const int64_t e = vi * VEC;
  const B200atVec<VEC> xi = b200at_ld_stream<VEC>(x + e), t = b200at_ld_stream<VEC>(noise + e);
  B200atVec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float r = B200AT_MUL(B200AT_MUL(B200AT_SUB(B200AT_MUL(2.0f, t.v[i]), 1.0f), eps), noise_level);
    const float v = B200AT_ADD(xi.v[i], r);
    o.v[i] = skip_projection ? v : b200at_clamp01(v);
  }
  b200at_st_keep<VEC>(x_adv + e, o);
}

// ---- FGSM step (fgsm_train.py:93-96): x_adv + alpha*eps*sign(grad), then eps-box around x and [0,1] ----
template <int VEC>
B200AT_HD void b200at_fgsm_step_body(const float* x, const float* x_adv, const float* grad, float* out, int64_t vi,
                                     float eps, float step, int skip_projection) {
  const int64_t e = vi * VEC;
  const B200atVec<VEC> xi = b200at_ld_stream<VEC>(x + e), xa = b200at_ld_stream<VEC>(x_adv + e),
                       g = b200at_ld_stream<VEC>(grad + e);
  B200atVec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    float v = B200AT_ADD(xa.v[i], B200AT_MUL(step, b200at_sign(g.v[i])));
    if (!skip_projection) {
      const float d = b200at_min(b200at_max(B200AT_SUB(v, xi.v[i]), -eps), eps);
      v = b200at_clamp01(B200AT_ADD(xi.v[i], d));
    }
    o.v[i] = v;
  }
  b200at_st_keep<VEC>(out + e, o);
}

// ---- K2: l2 update, one phase per call (autopgd_train_clean.py:228-237).  Phases 0..2 return the
// partial sum of squares of this vector; phase 3 writes the new iterate and applies the pending ops
// exactly like the l-inf body.  `sums` = {||g||^2, ||z-x||^2, ||w-x||^2} of this sample so far.
// Per-sample constants of one l2 phase: read / derived once per sample by the single-launch kernel (the per-vector form
// below rebuilds them for every vector: a 64-bit division for the sample index, three square roots, two state loads).
struct B200atL2Ctx {
  float step;
  B200atL2Norms nm;
  bool improved, write_adv, restore;
};
template <int PHASE>
B200AT_HD B200atL2Ctx b200at_l2_ctx(const B200atImages& p, int b, const float* sums) {
  B200atL2Ctx c;
  const int32_t fl = b200at_f2i(p.st[(int64_t)B200AT_ST_FLAGS * p.B + b]);
  c.step = p.st[(int64_t)B200AT_ST_STEP * p.B + b];
  c.improved = fl & B200AT_F_IMPROVED;
  c.write_adv = fl & B200AT_F_WRITE_ADV;
  c.restore = (fl & B200AT_F_RESTORE) && !c.improved;
  c.nm = b200at_l2_norms<PHASE>(sums);
  return c;
}

// KEEP: load with the L2-resident policy (phases 0..2 of the single-launch form: the next phase re-reads the operands).
// `e` = first element of the vector (of sample b, whose constants are in c).  Split in a load half and an apply half so
// that a caller can issue the loads of several vectors before the first store (x_new may alias x_old: the compiler
// cannot move a later load above an earlier store itself).
// `gpre`: the gradient vector if the caller already holds it (shared-memory copy made in phase 0), else null.
template <int VEC>
struct B200atL2Ops { B200atVec<VEC> x, xc, xo, g; };

template <int PHASE, int VEC, bool KEEP>
B200AT_HD void b200at_l2_load(const B200atImages& p, int64_t e, const B200atL2Ctx& c, const B200atVec<VEC>* gpre,
                              B200atL2Ops<VEC>& o) {
  const bool restore = c.restore;
  if (gpre) o.g = *gpre;
  else o.g = KEEP ? b200at_ld_keep<VEC>((restore ? p.grad_best : p.grad) + e)
                  : b200at_ld_stream<VEC>((restore ? p.grad_best : p.grad) + e);
  if (PHASE > 0) {
    o.x = KEEP ? b200at_ld_keep<VEC>(p.x + e) : b200at_ld_stream<VEC>(p.x + e);
    o.xc = KEEP ? b200at_ld_keep<VEC>((restore ? p.x_best : p.x_adv) + e)
                : b200at_ld_stream<VEC>((restore ? p.x_best : p.x_adv) + e);
  }
  if (PHASE > 1) o.xo = KEEP ? b200at_ld_keep<VEC>(p.x_old + e) : b200at_ld_stream<VEC>(p.x_old + e);
}

template <int PHASE, int VEC>
B200AT_HD float b200at_l2_apply(const B200atImages& p, int64_t e, const B200atL2Ctx& c, float eps, float a,
                                float one_minus_a, const B200atL2Ops<VEC>& in) {
  const bool improved = c.improved, write_adv = c.write_adv, restore = c.restore;
  if (PHASE == 3) {
    if (!restore) {
      if (write_adv) b200at_st_stream<VEC>(p.x_best_adv + e, in.xc);
      if (improved) {
        b200at_st_stream<VEC>(p.x_best + e, in.xc);
        b200at_st_stream<VEC>(p.grad_best + e, in.g);
      }
    } else {
      if (write_adv) b200at_st_stream<VEC>(p.x_best_adv + e, b200at_ld_stream<VEC>(p.x_adv + e));
      b200at_st_stream<VEC>(p.x_adv + e, in.xc);
    }
  }
  B200atVec<VEC> o;
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    o.v[i] = b200at_l2_elem<PHASE>(PHASE > 0 ? in.x.v[i] : 0.0f, PHASE > 0 ? in.xc.v[i] : 0.0f,
                                   PHASE > 1 ? in.xo.v[i] : 0.0f, in.g.v[i], c.step, eps, a, one_minus_a, c.nm);
    acc = B200AT_ADD(acc, B200AT_MUL(o.v[i], o.v[i]));
  }
  if (PHASE == 3) b200at_st_keep<VEC>(p.x_new + e, o);
  return acc;
}

template <int PHASE, int VEC, bool KEEP>
B200AT_HD float b200at_l2_body_ctx(const B200atImages& p, int64_t e, const B200atL2Ctx& c, float eps, float a,
                                   float one_minus_a, const B200atVec<VEC>* gpre = nullptr) {
  B200atL2Ops<VEC> in;
  b200at_l2_load<PHASE, VEC, KEEP>(p, e, c, gpre, in);
  return b200at_l2_apply<PHASE, VEC>(p, e, c, eps, a, one_minus_a, in);
}

// per-vector form (four-launch kernels, tests/hostcheck): vi = vector index over the whole batch
template <int PHASE, int VEC>
B200AT_HD float b200at_l2_body(const B200atImages& p, int64_t vi, float eps, float a, float one_minus_a,
                               const float* sums) {
  const int64_t e = vi * VEC;
  const int b = (int)(e / p.n);
  return b200at_l2_body_ctx<PHASE, VEC, false>(p, e, b200at_l2_ctx<PHASE>(p, b, sums), eps, a, one_minus_a);
}

// ---- iterate-log variants (n_iter + 1 <= B200AT_LOG_MAX_SLOTS): no image copies at all -------------
struct B200atSlots {
  const float* x[B200AT_LOG_MAX_SLOTS];   // iterates x_adv^(0..)
  const float* g[B200AT_LOG_MAX_SLOTS];   // gradients at those iterates
};

// Slot pick without dynamic indexing of the kernel-parameter struct (that would spill the struct to local
// memory): 3-level select tree over the 8 constant-indexed pointers.
B200AT_HD const float* b200at_pick(const float* const* a, int i) {
  const float* a01 = (i & 1) ? a[1] : a[0];
  const float* a23 = (i & 1) ? a[3] : a[2];
  const float* a45 = (i & 1) ? a[5] : a[4];
  const float* a67 = (i & 1) ? a[7] : a[6];
  const float* lo = (i & 2) ? a23 : a01;
  const float* hi = (i & 2) ? a67 : a45;
  return (i & 4) ? hi : lo;
}

// K1 (log): x_new = move(x, x_adv, x_adv_old, grad) with the three operands picked per sample by slot
// index.  Exactly 20 B/element (16 on the first move, where x_old and x_adv are the same slot).
template <int VEC>
B200AT_HD void b200at_linf_log_body(const B200atSlots& sl, const float* x, float* x_new, const float* st, int64_t B,
                                    int64_t n, int64_t vi, float eps, float a, float one_minus_a) {
  const int64_t e = vi * VEC;
  const int b = (int)(e / n);
  const float step = st[(int64_t)B200AT_ST_STEP * B + b];
  const int ic = b200at_f2i(st[(int64_t)B200AT_ST_IDX_CUR * B + b]) & (B200AT_LOG_MAX_SLOTS - 1);
  const int io = b200at_f2i(st[(int64_t)B200AT_ST_IDX_OLD * B + b]) & (B200AT_LOG_MAX_SLOTS - 1);
  const int ig = b200at_f2i(st[(int64_t)B200AT_ST_GIDX_CUR * B + b]) & (B200AT_LOG_MAX_SLOTS - 1);
  const B200atVec<VEC> xv = b200at_ld_stream<VEC>(x + e);
  const B200atVec<VEC> xc = b200at_ld_stream<VEC>(b200at_pick(sl.x, ic) + e);
  const B200atVec<VEC> xo = (io == ic) ? xc : b200at_ld_stream<VEC>(b200at_pick(sl.x, io) + e);
  const B200atVec<VEC> g = b200at_ld_stream<VEC>(b200at_pick(sl.g, ig) + e);
  B200atVec<VEC> o;
#pragma unroll
  for (int i = 0; i < VEC; ++i) o.v[i] = b200at_linf_elem(xv.v[i], xc.v[i], xo.v[i], g.v[i], step, eps, a, one_minus_a);
  b200at_st_keep<VEC>(x_new + e, o);
}

// final assembly: x_best[b] = slot[idx_best[b]][b], x_best_adv[b] = slot[idx_best_adv[b]][b]
template <int VEC>
B200AT_HD void b200at_gather_body(const B200atSlots& sl, float* x_best, float* x_best_adv, const float* st,
                                  int64_t B, int64_t n, int64_t vi) {
  const int64_t e = vi * VEC;
  const int b = (int)(e / n);
  const int ib = b200at_f2i(st[(int64_t)B200AT_ST_IDX_BEST * B + b]) & (B200AT_LOG_MAX_SLOTS - 1);
  const int ia = b200at_f2i(st[(int64_t)B200AT_ST_IDX_BEST_ADV * B + b]) & (B200AT_LOG_MAX_SLOTS - 1);
  const B200atVec<VEC> vb = b200at_ld_stream<VEC>(b200at_pick(sl.x, ib) + e);
  b200at_st_keep<VEC>(x_best + e, vb);
  b200at_st_stream<VEC>(x_best_adv + e, (ia == ib) ? vb : b200at_ld_stream<VEC>(b200at_pick(sl.x, ia) + e));
}
