// TEST INFRASTRUCTURE ONLY: host build of the kernel bodies in revisiting-at_b200/csrc/b200at_bodies.cuh
// so the CPU test-suite can check indexing, flag protocol and arithmetic against the oracle without a
// GPU.  The product never loads this library (see revisiting-at_b200/_abi.py: CUDA library or error).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/hostcheck/build.py)
#include "../../revisiting-at_b200/csrc/b200at_bodies.cuh"
#include <algorithm>
#include <vector>

template <int VEC>
static void linf_all(const B200atImages& p, float eps, float a, float oma) {
  const int64_t nvec = p.B * p.n / VEC;
  for (int64_t v = 0; v < nvec; ++v) b200at_linf_body<VEC>(p, v, eps, a, oma);
}

template <int PHASE, int VEC>
static void l2_phase(const B200atImages& p, float eps, float a, float oma, float* sums /*[B][3]*/) {
  const int64_t nvec_row = p.n / VEC;
  for (int64_t b = 0; b < p.B; ++b) {
    double acc = 0.0;
    for (int64_t v = 0; v < nvec_row; ++v)
      acc += b200at_l2_body<PHASE, VEC>(p, b * nvec_row + v, eps, a, oma, sums + 3 * b);
    if (PHASE < 3) sums[3 * b + PHASE] = (float)acc;
  }
}

extern "C" {

void hc_init(const float* x, float* x_adv, float* st, int64_t B, int64_t n, float step0, float topk0, int vec) {
  for (int64_t i = 0; i < (int64_t)B200AT_ST_ROWS * B; ++i) st[i] = 0.f;
  for (int64_t b = 0; b < B; ++b) {
    int nnz = 0;
    if (vec == 4) for (int64_t v = 0; v < n / 4; ++v) nnz += b200at_init_body<4>(x, x_adv, b * (n / 4) + v);
    else for (int64_t v = 0; v < n; ++v) nnz += b200at_init_body<1>(x, x_adv, b * n + v);
    st[(int64_t)B200AT_ST_SP_ADV * B + b] = b200at_i2f(nnz);
    st[(int64_t)B200AT_ST_STEP * B + b] = step0;
    st[(int64_t)B200AT_ST_TOPK * B + b] = topk0;
    st[(int64_t)B200AT_ST_SP_OLD * B + b] = (float)n;
    st[(int64_t)B200AT_ST_REDUCED_LAST * B + b] = 1.0f;
  }
}

void hc_linf_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad, float* x_best,
                  float* grad_best, float* x_best_adv, const float* st, int64_t B, int64_t n, float eps, float a,
                  int vec) {
  B200atImages p{x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, st, B, n};
  const float oma = (float)(1.0 - (double)a);
  if (vec == 4) linf_all<4>(p, eps, a, oma); else linf_all<1>(p, eps, a, oma);
}

void hc_l2_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad, float* x_best,
                float* grad_best, float* x_best_adv, const float* st, float* sums, int64_t B, int64_t n, float eps,
                float a, int vec) {
  B200atImages p{x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, st, B, n};
  const float oma = (float)(1.0 - (double)a);
  if (vec == 4) { l2_phase<0, 4>(p, eps, a, oma, sums); l2_phase<1, 4>(p, eps, a, oma, sums);
                  l2_phase<2, 4>(p, eps, a, oma, sums); l2_phase<3, 4>(p, eps, a, oma, sums); }
  else { l2_phase<0, 1>(p, eps, a, oma, sums); l2_phase<1, 1>(p, eps, a, oma, sums);
         l2_phase<2, 1>(p, eps, a, oma, sums); l2_phase<3, 1>(p, eps, a, oma, sums); }
}

// l1 step: same helper arithmetic and the same algorithm as the device path (exact order statistic,
// bit-pattern sectioning of the water level); the reductions are plain host loops.
void hc_l1_step(const float* x, float* x_adv, float* x_new, const float* grad, float* x_best, float* grad_best,
                float* x_best_adv, float* st, int64_t B, int64_t n, float eps) {
  for (int64_t b = 0; b < B; ++b) {
    const int32_t fl = b200at_f2i(st[(int64_t)B200AT_ST_FLAGS * B + b]);
    const bool improved = fl & B200AT_F_IMPROVED, write_adv = fl & B200AT_F_WRITE_ADV;
    const bool restore = (fl & B200AT_F_RESTORE) && !improved;
    const float step = st[(int64_t)B200AT_ST_STEP * B + b];
    const float* X = x + b * n;
    const float* XC = (restore ? x_best : x_adv) + b * n;
    const float* G = (restore ? grad_best : grad) + b * n;
    std::vector<uint32_t> keys(n);
    for (int64_t i = 0; i < n; ++i) keys[i] = b200at_l1_key(G[i]);
    const int64_t rank = b200at_l1_rank(st[(int64_t)B200AT_ST_TOPK * B + b], n);
    std::nth_element(keys.begin(), keys.begin() + rank, keys.end());
    const float thr = b200at_i2f((int32_t)keys[rank]);
    float nnz = 0.f;
    for (int64_t i = 0; i < n; ++i) nnz += (fabsf(G[i]) >= thr && b200at_sign(G[i]) != 0.f) ? 1.f : 0.f;
    std::vector<float> y(n), u(n);
    double sb = 0, sa = 0;
    for (int64_t i = 0; i < n; ++i) {
      y[i] = b200at_l1_y(X[i], XC[i], G[i], step, thr, nnz);
      u[i] = b200at_l1_u(X[i], y[i]);
      sb += fabsf(y[i]); sa -= u[i];
    }
    const float c = eps - (float)sb;
    const int need = ((float)sa + c < 0.f);
    uint32_t prefix = 0;
    if (need) {
      for (int pass = 0; pass < B200AT_L1_PASSES; ++pass) {
        int kstar = 0;
        for (int k = 1; k <= b200at_l1_ncand(pass); ++k) {
          const float cand = b200at_l1_cand(prefix, k, pass);
          double gk = 0;
          for (int64_t i = 0; i < n; ++i) gk += b200at_l1_level(cand, -u[i], fabsf(y[i]));
          if ((float)gk + c < 0.f) kstar = k; else break;
        }
        prefix |= (uint32_t)kstar << b200at_l1_shift(pass);
      }
    }
    const float alpha = b200at_i2f((int32_t)prefix);
    int moved = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t e = b * n + i;
      if (!restore) {
        if (write_adv) x_best_adv[e] = XC[i];
        if (improved) { x_best[e] = XC[i]; grad_best[e] = G[i]; }
      } else if (write_adv) {
        x_best_adv[e] = x_adv[e];
      }
      const float o = b200at_l1_out(X[i], y[i], u[i], need, alpha);
      moved += (B200AT_SUB(o, X[i]) != 0.f);
      x_new[e] = o;
    }
    st[(int64_t)B200AT_ST_SP_ADV * B + b] = b200at_i2f(moved);
  }
}

void hc_flush(const float* x_adv, float* x_best, float* x_best_adv, const float* st, int64_t B, int64_t n, int vec) {
  B200atImages p{nullptr, const_cast<float*>(x_adv), nullptr, nullptr, nullptr, x_best, nullptr, x_best_adv, st, B, n};
  if (vec == 4) for (int64_t v = 0; v < B * n / 4; ++v) b200at_flush_body<4>(p, v);
  else for (int64_t v = 0; v < B * n; ++v) b200at_flush_body<1>(p, v);
}

void hc_bookkeep(float* st, float* loss_steps, int64_t B, const float* loss, const int* pred, int iter, int n_iter,
                 int ckpt_k, int norm_kind, float step_full, float step_min, int64_t n_fts, int has_grad) {
  for (int64_t b = 0; b < B; ++b)
    b200at_bookkeep_sample(st, loss_steps, (int)B, (int)b, loss[b], pred[b], iter, n_iter, ckpt_k, norm_kind,
                           step_full, step_min, (float)n_fts, has_grad);
}

static void fill(B200atSlots& sl, const float* const* xs, const float* const* gs, int n_slots) {
  for (int i = 0; i < B200AT_LOG_MAX_SLOTS; ++i) {
    sl.x[i] = xs[i < n_slots ? i : 0];
    sl.g[i] = gs ? gs[i < n_slots ? i : 0] : nullptr;
  }
}

void hc_linf_step_log(const float* x, const float* const* xs, const float* const* gs, int n_slots, float* x_new,
                      const float* st, int64_t B, int64_t n, float eps, float a, int vec) {
  B200atSlots sl; fill(sl, xs, gs, n_slots);
  const float oma = (float)(1.0 - (double)a);
  if (vec == 4) for (int64_t v = 0; v < B * n / 4; ++v) b200at_linf_log_body<4>(sl, x, x_new, st, B, n, v, eps, a, oma);
  else for (int64_t v = 0; v < B * n; ++v) b200at_linf_log_body<1>(sl, x, x_new, st, B, n, v, eps, a, oma);
}

void hc_gather_best(const float* const* xs, int n_slots, float* x_best, float* x_best_adv, const float* st, int64_t B,
                    int64_t n, int vec) {
  B200atSlots sl; fill(sl, xs, nullptr, n_slots);
  if (vec == 4) for (int64_t v = 0; v < B * n / 4; ++v) b200at_gather_body<4>(sl, x_best, x_best_adv, st, B, n, v);
  else for (int64_t v = 0; v < B * n; ++v) b200at_gather_body<1>(sl, x_best, x_best_adv, st, B, n, v);
}

void hc_fgsm_start(const float* x, const float* noise, float* x_adv, int64_t total, float eps, float nl, int skip,
                   int vec) {
  if (vec == 4) for (int64_t v = 0; v < total / 4; ++v) b200at_fgsm_start_body<4>(x, noise, x_adv, v, eps, nl, skip);
  else for (int64_t v = 0; v < total; ++v) b200at_fgsm_start_body<1>(x, noise, x_adv, v, eps, nl, skip);
}

void hc_fgsm_step(const float* x, const float* x_adv, const float* grad, float* out, int64_t total, float eps,
                  float step, int skip, int vec) {
  if (vec == 4) for (int64_t v = 0; v < total / 4; ++v) b200at_fgsm_step_body<4>(x, x_adv, grad, out, v, eps, step, skip);
  else for (int64_t v = 0; v < total; ++v) b200at_fgsm_step_body<1>(x, x_adv, grad, out, v, eps, step, skip);
}

}  // extern "C"
