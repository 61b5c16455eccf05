/* b200at -- C ABI of the B200-native APGD attack-step kernels (libb200at.so).
 *
 * The reference (nmndeep/revisiting-at) is pure Python and has no FFI; what these entry points
 * replace are the eager torch op sequences inside `apgd_train` / `fgsm_train`.  Each declaration
 * cites the reference lines it stands in for (paths relative to the reference root).
 *
 * Conventions:
 *   - plain device pointers + sizes; no torch types; caller owns every buffer;
 *   - every call launches on `stream` (a cudaStream_t passed as void*), never allocates, never
 *     synchronises, never throws; the return value is the cudaError_t of the launch (0 = ok);
 *   - image buffers are [B][n] fp32, dense, all in the same element order (n = C*H*W);
 *   - `state` is [B200AT_ST_ROWS][B] fp32 rows (integer rows bit-cast to int32), layout in
 *     revisiting-at_b200/csrc/b200at_math.cuh and mirrored in revisiting-at_b200/_abi.py;
 *   - re-entrant, no global state.
 */
#ifndef B200AT_H
#define B200AT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AT_ABI_VERSION 1

/* logits dtype codes */
#define B200AT_DT_F32 0
#define B200AT_DT_BF16 1
#define B200AT_DT_F16 2
/* loss codes: criterion_dict keys 'ce' / 'dlr' (autopgd_train_clean.py:113-114) */
#define B200AT_LOSS_CE 0
#define B200AT_LOSS_DLR 1
#define B200AT_LOSS_DLR_TARGETED 2 /* only through b200at_loss_bookkeep_targeted */

int b200at_abi_version(void);

/* autopgd_train_clean.py:135-146,163-171: x_adv = clamp(x,0,1); per-sample state zeroed and seeded
 * (step = step0 = float32(alpha*eps), reduced_last_check = 1, topk = topk0, sp_old = n).
 * x_best / x_best_adv / grad_best are NOT written here: the first bookkeeping call marks every
 * sample "improved" and the first step (or flush) pass seeds them. */
int b200at_apgd_init(const float* x, float* x_adv, float* state, int64_t B, int64_t n, float step0, float topk0,
                     void* stream);

/* autopgd_train_clean.py:113,179-205 (iter = -1) and :273-349 / :351-364 (iter >= 0):
 * per-sample loss ('ce' hard/soft targets, 'dlr'), dL/dlogits (what autograd would hand to the
 * model's backward for loss.sum()), prediction, acc &= pred, strict best-loss compare, loss_steps
 * history, check_oscillation (:116-121) + step halving at a checkpoint (ckpt_k > 0 = window k),
 * l1 sparsity adaptation.  Leaves the pending image ops in state[FLAGS].
 * Exactly one of y_hard ([B] int64) / y_soft ([B][C] fp32) is non-null. dlogits / loss_out may be null.
 * A hard label (or target class) outside [0, C) traps the kernel -- the launch fails at the next synchronisation,
 * like the device-side assert of torch's CUDA cross_entropy that the reference would hit; there is no host check
 * (it would be a synchronisation per forward). */
int b200at_loss_bookkeep(const void* logits, int logits_dtype, const int64_t* y_hard, const float* y_soft,
                         void* dlogits, float* loss_out, float* state, float* loss_steps, int64_t B, int64_t C,
                         int iter, int n_iter, int ckpt_k, int norm_kind, int loss_kind, float step_full,
                         float step_min, int64_t n_fts, void* stream);

/* Same kernel with the targeted DLR loss (autopgd_train_clean.py:106-111 `dlr_loss_targeted`; the loss of
 * AutoAttack's APGD-T, AA_eval.py:226-239):  -(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12), y_target [B] int64.
 * The prediction / robust mask still compares with y_hard.  C >= 4. */
int b200at_loss_bookkeep_targeted(const void* logits, int logits_dtype, const int64_t* y_hard, const int64_t* y_target,
                                  void* dlogits, float* loss_out, float* state, float* loss_steps, int64_t B,
                                  int64_t C, int iter, int n_iter, int ckpt_k, int norm_kind, float step_full,
                                  float step_min, int64_t n_fts, void* stream);

/* autopgd_train_clean.py:213-226,260 fused with the image side of :304, :321-324, :345-346:
 * one pass that (1) applies the pending x_best / grad_best / x_best_adv writes and the restore from
 * x_best/grad_best for flagged samples, (2) moves x_adv by step*sign(grad) with momentum `a`
 * (1.0 on the first move, 0.75 after), projected on the fp32 eps-ball around x and on [0,1].
 * x_new receives the new iterate; it may alias x_old (in place).  On the first move pass
 * x_old = x_adv.  Algorithmic traffic: 20 B / element. */
int b200at_linf_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad,
                     float* x_best, float* grad_best, float* x_best_adv, const float* state, int64_t B, int64_t n,
                     float eps, float a, void* stream);

/* Iterate-log form of b200at_linf_step for short attacks (n_iter + 1 <= B200AT_LOG_MAX_SLOTS = 8, the training
 * configuration): every iterate x_adv^(k) and gradient stays in its own buffer ("slot"), and the masked
 * copies of autopgd_train_clean.py:304,:321-324,:345-346 are replaced by per-sample slot indices kept by
 * b200at_loss_bookkeep in state[IDX_*].  One launch moves exactly 20 B/element (16 on the first move).
 * x_slots / g_slots: HOST arrays of n_slots device pointers (slot k = iterate k / gradient at iterate k;
 * unused gradient slots may repeat a valid pointer). */
#define B200AT_LOG_MAX_SLOTS 8
int b200at_linf_step_log(const float* x, const float* const* x_slots, const float* const* g_slots, int n_slots,
                         float* x_new, const float* state, int64_t B, int64_t n, float eps, float a, void* stream);
/* end of an iterate-log attack: x_best[b] = x_slots[IDX_BEST[b]][b], x_best_adv[b] = x_slots[IDX_BEST_ADV[b]][b] */
int b200at_gather_best(const float* const* x_slots, int n_slots, float* x_best, float* x_best_adv,
                       const float* state, int64_t B, int64_t n, void* stream);

/* autopgd_train_clean.py:228-237 (+ the same pending image ops as b200at_linf_step): l2 move with
 * momentum; the three dependent per-sample norms (||grad||, ||z-x||, ||w-x||) are deterministic
 * two-level sums.  scratch: >= B200AT_L2_SCRATCH_FLOATS(B) floats, contents irrelevant on entry. */
#define B200AT_L2_SCRATCH_FLOATS(B) (3 * 32 * (B))
int b200at_l2_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad,
                   float* x_best, float* grad_best, float* x_best_adv, const float* state, float* scratch,
                   int64_t B, int64_t n, float eps, float a, void* stream);

/* autopgd_train_clean.py:239-250 with L1_projection (:24-91) (+ the pending image ops): sparse sign step on
 * the top-k |grad| coordinates (exact order statistic by radix select), then projection of x+delta onto the
 * l1 ball of radius eps intersected with [0,1]^n.  No momentum, x_old is not read; x_new must NOT alias
 * x_adv.  Writes nnz(x_new - x) to state[SP_ADV].  scratch: >= B200AT_L1_SCRATCH_WORDS(B) 4-byte words.
 * 12 kernel launches + 1 memset: the per-sample decisions between the image passes are made in the prologue of the
 * next pass, not by helper launches. */
#define B200AT_L1_SCRATCH_WORDS(B) ((3 * 2048 + 32 + 32 * 2 + 2 * 32 * 32) * (B))
int b200at_l1_step(const float* x, float* x_adv, float* x_new, const float* grad, float* x_best, float* grad_best,
                   float* x_best_adv, float* state, void* scratch, int64_t B, int64_t n, float eps, void* stream);

/* image side of autopgd_train_clean.py:304,:322 after the LAST forward: pending x_best / x_best_adv writes. */
int b200at_flush_best(const float* x_adv, float* x_best, float* x_best_adv, const float* state, int64_t B,
                      int64_t n, void* stream);

/* fgsm_train.py:79-83: random start x_adv = x + (2*noise-1)*eps*noise_level (noise ~ U[0,1), drawn by the
 * caller so the RNG stream stays torch's), clamped to [0,1] unless skip_projection. `total` = B*n. */
int b200at_fgsm_start(const float* x, const float* noise, float* x_adv, int64_t total, float eps,
                      float noise_level, int skip_projection, void* stream);

/* fgsm_train.py:93-96: out = x_adv + step*sign(grad) with step = float32(alpha*eps); unless
 * skip_projection: out = clamp(x + clamp(out - x, -eps, eps), 0, 1).  out may alias x_adv. */
int b200at_fgsm_step(const float* x, const float* x_adv, const float* grad, float* out, int64_t total, float eps,
                     float step, int skip_projection, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200AT_H */
