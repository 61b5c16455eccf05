// sm_100a kernels for the APGD attack step + their C ABI (include/b200at.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (see csrc/build.py).
//
// All image-sized passes are pure HBM streams: 16-byte coalesced accesses, streaming cache policy on
// everything that is touched once, default policy on the new iterate (the model's first layer reads
// it next, the 126 MB L2 can hold it).  No shared memory, no tensor cores: there is no reuse.
// Entry points never allocate, never synchronise, never throw; they launch on the caller's stream
// and return the cudaError_t of the launch as int.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "b200at_bodies.cuh"
#include "../../include/b200at.h"

namespace {

constexpr int kThreads = 256;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int grid_for(int64_t items, int per_cta) {
  int64_t g = (items + per_cta - 1) / per_cta;
  return (int)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads) linf_step_kernel(B200atImages p, int64_t nvec, float eps, float a,
                                                              float one_minus_a) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_linf_body<VEC>(p, vi, eps, a, one_minus_a);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) flush_kernel(B200atImages p, int64_t nvec) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_flush_body<VEC>(p, vi);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) fgsm_start_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ noise,
                                                               float* __restrict__ x_adv, int64_t nvec, float eps,
                                                               float noise_level, int skip) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_fgsm_start_body<VEC>(x, noise, x_adv, vi, eps, noise_level, skip);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) fgsm_step_kernel(const float* __restrict__ x, const float* x_adv,
                                                              const float* __restrict__ grad, float* out,
                                                              int64_t nvec, float eps, float step, int skip) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_fgsm_step_body<VEC>(x, x_adv, grad, out, vi, eps, step, skip);
}

// One CTA never straddles two samples (grid.y = sample), so the nnz count reduces per CTA.
template <int VEC>
__global__ void __launch_bounds__(kThreads) init_kernel(const float* __restrict__ x, float* __restrict__ x_adv,
                                                         float* __restrict__ st, int64_t B, int64_t n, float step0,
                                                         float topk0) {
  const int b = blockIdx.y;
  const int64_t nvec_row = n / VEC;
  int nnz = 0;
  for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvec_row; v += (int64_t)gridDim.x * kThreads)
    nnz += b200at_init_body<VEC>(x, x_adv, (int64_t)b * nvec_row + v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
  int* sp_adv = reinterpret_cast<int*>(st + (int64_t)B200AT_ST_SP_ADV * B + b);
  if ((threadIdx.x & 31) == 0 && nnz) atomicAdd(sp_adv, nnz);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st[(int64_t)B200AT_ST_STEP * B + b] = step0;
    st[(int64_t)B200AT_ST_TOPK * B + b] = topk0;
    st[(int64_t)B200AT_ST_SP_OLD * B + b] = (float)n;
    st[(int64_t)B200AT_ST_REDUCED_LAST * B + b] = 1.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// K6 + K5: per-sample loss, dL/dlogits, prediction and the whole per-sample bookkeeping
// (autopgd_train_clean.py:113, :179-205, :273-349).  One warp per row of logits [B][C].
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// (value, index) max with torch semantics: NaN wins, first occurrence wins ties
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

struct LossArgs {
  const void* logits;
  const int64_t* y_hard;  // [B] or null
  const float* y_soft;    // [B][C] or null
  void* dlogits;          // [B][C] same dtype as logits, or null
  float* loss_out;        // [B] or null
  float* st;
  float* loss_steps;
  int B, C, iter, n_iter, ckpt_k, norm_kind, loss_kind;
  float step_full, step_min, n_fts;
};

template <typename T>
__global__ void __launch_bounds__(128) loss_bookkeep_kernel(LossArgs p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const T* z = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.C;
  T* dz = p.dlogits ? reinterpret_cast<T*>(p.dlogits) + (int64_t)b * p.C : nullptr;
  const float* ys = p.y_soft ? p.y_soft + (int64_t)b * p.C : nullptr;
  const float ninf = __int_as_float(0xff800000);

  // pass 1: top-1 of the logits (+ label = argmax of the soft target, + its mass)
  float m1 = ninf; int i1 = 0x7fffffff;
  float ym = ninf; int yi = 0x7fffffff; float ysum = 0.0f;
  for (int c = lane; c < p.C; c += 32) {
    const float v = to_f32<T>(z[c]);
    if (better(v, c, m1, i1)) { m1 = v; i1 = c; }
    if (ys) {
      const float t = ys[c];
      ysum += t;
      if (better(t, c, ym, yi)) { ym = t; yi = c; }
    }
  }
  warp_argmax(m1, i1);
  int label;
  if (ys) { warp_argmax(ym, yi); ysum = warp_sum(ysum); label = yi; }
  else label = (int)p.y_hard[b];
  const int pred = (i1 == label);

  float loss;
  if (p.loss_kind == B200AT_LOSS_CE) {
    // pass 2: log-sum-exp;  pass 3: loss (soft) and dL/dz = softmax * sum(y) - y
    float se = 0.0f;
    for (int c = lane; c < p.C; c += 32) se += expf(to_f32<T>(z[c]) - m1);
    se = warp_sum(se);
    const float lse = logf(se);
    float acc = 0.0f;
    for (int c = lane; c < p.C; c += 32) {
      const float lsm = (to_f32<T>(z[c]) - m1) - lse;
      const float t = ys ? ys[c] : (c == label ? 1.0f : 0.0f);
      if (ys) acc -= t * lsm;
      else if (c == label) acc = -lsm;
      if (dz) dz[c] = from_f32<T>(expf(lsm) * (ys ? ysum : 1.0f) - t);
    }
    loss = warp_sum(acc);
  } else {
    // DLR (:99-104): -(z_y - z_other) / (z_(1) - z_(3) + 1e-12); top-3 by two more masked argmax sweeps
    float m2 = ninf, m3 = ninf; int i2 = 0x7fffffff, i3 = 0x7fffffff;
    for (int c = lane; c < p.C; c += 32) {
      const float v = to_f32<T>(z[c]);
      if (c != i1 && better(v, c, m2, i2)) { m2 = v; i2 = c; }
    }
    warp_argmax(m2, i2);
    for (int c = lane; c < p.C; c += 32) {
      const float v = to_f32<T>(z[c]);
      if (c != i1 && c != i2 && better(v, c, m3, i3)) { m3 = v; i3 = c; }
    }
    warp_argmax(m3, i3);
    const float zy = to_f32<T>(z[label]);
    const int io = pred ? i2 : i1;
    const float zo = pred ? m2 : m1;
    const float den = (m1 - m3) + 1e-12f;
    const float num = zy - zo;
    loss = -num / den;
    if (dz) {
      // d(-num/den) = -(e_y - e_o)/den + num/den^2 * (e_1 - e_3)
      const float q = num / (den * den);
      for (int c = lane; c < p.C; c += 32) {
        float gz = 0.0f;
        if (c == label) gz -= 1.0f / den;
        if (c == io) gz += 1.0f / den;
        if (c == i1) gz += q;
        if (c == i3) gz -= q;
        dz[c] = from_f32<T>(gz);
      }
    }
  }
  if (lane == 0) {
    if (p.loss_out) p.loss_out[b] = loss;
    b200at_bookkeep_sample(p.st, p.loss_steps, p.B, b, loss, pred, p.iter, p.n_iter, p.ckpt_k, p.norm_kind,
                           p.step_full, p.step_min, p.n_fts);
  }
}

}  // namespace

extern "C" {

int b200at_abi_version(void) { return B200AT_ABI_VERSION; }

int b200at_apgd_init(const float* x, float* x_adv, float* state, int64_t B, int64_t n, float step0, float topk0,
                     void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  cudaError_t e = cudaMemsetAsync(state, 0, sizeof(float) * B200AT_ST_ROWS * B, s);
  if (e != cudaSuccess) return (int)e;
  const bool v4 = (n % 4 == 0) && aligned16(x) && aligned16(x_adv);
  const int64_t per_row = v4 ? n / 4 : n;
  int gx = grid_for(per_row, kThreads * 4);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, (unsigned)B);
  if (B > 65535) return (int)cudaErrorInvalidValue;
  if (v4) init_kernel<4><<<grid, kThreads, 0, s>>>(x, x_adv, state, B, n, step0, topk0);
  else init_kernel<1><<<grid, kThreads, 0, s>>>(x, x_adv, state, B, n, step0, topk0);
  return (int)cudaGetLastError();
}

int b200at_linf_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad,
                     float* x_best, float* grad_best, float* x_best_adv, const float* state, int64_t B, int64_t n,
                     float eps, float a, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  B200atImages p{x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, B, n};
  const float oma = (float)(1.0 - (double)a);
  const bool v4 = (n % 4 == 0) && aligned16(x) && aligned16(x_adv) && aligned16(x_old) && aligned16(x_new) &&
                  aligned16(grad) && aligned16(x_best) && aligned16(grad_best) && aligned16(x_best_adv);
  const int64_t nvec = v4 ? B * n / 4 : B * n;
  if (v4) linf_step_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(p, nvec, eps, a, oma);
  else linf_step_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(p, nvec, eps, a, oma);
  return (int)cudaGetLastError();
}

int b200at_flush_best(const float* x_adv, float* x_best, float* x_best_adv, const float* state, int64_t B,
                      int64_t n, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  B200atImages p{nullptr, const_cast<float*>(x_adv), nullptr, nullptr, nullptr, x_best, nullptr, x_best_adv, state, B, n};
  const bool v4 = (n % 4 == 0) && aligned16(x_adv) && aligned16(x_best) && aligned16(x_best_adv);
  const int64_t nvec = v4 ? B * n / 4 : B * n;
  if (v4) flush_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(p, nvec);
  else flush_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(p, nvec);
  return (int)cudaGetLastError();
}

int b200at_fgsm_start(const float* x, const float* noise, float* x_adv, int64_t total, float eps,
                       float noise_level, int skip_projection, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (total <= 0) return (int)cudaSuccess;
  const bool v4 = (total % 4 == 0) && aligned16(x) && aligned16(noise) && aligned16(x_adv);
  const int64_t nvec = v4 ? total / 4 : total;
  if (v4) fgsm_start_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(x, noise, x_adv, nvec, eps, noise_level, skip_projection);
  else fgsm_start_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(x, noise, x_adv, nvec, eps, noise_level, skip_projection);
  return (int)cudaGetLastError();
}

int b200at_fgsm_step(const float* x, const float* x_adv, const float* grad, float* out, int64_t total, float eps,
                      float step, int skip_projection, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (total <= 0) return (int)cudaSuccess;
  const bool v4 = (total % 4 == 0) && aligned16(x) && aligned16(x_adv) && aligned16(grad) && aligned16(out);
  const int64_t nvec = v4 ? total / 4 : total;
  if (v4) fgsm_step_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(x, x_adv, grad, out, nvec, eps, step, skip_projection);
  else fgsm_step_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(x, x_adv, grad, out, nvec, eps, step, skip_projection);
  return (int)cudaGetLastError();
}

int b200at_loss_bookkeep(const void* logits, int logits_dtype, const int64_t* y_hard, const float* y_soft,
                         void* dlogits, float* loss_out, float* state, float* loss_steps, int64_t B, int64_t C,
                         int iter, int n_iter, int ckpt_k, int norm_kind, int loss_kind, float step_full,
                         float step_min, int64_t n_fts, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0) return (int)cudaSuccess;
  if ((y_hard == nullptr) == (y_soft == nullptr)) return (int)cudaErrorInvalidValue;
  if (loss_kind == B200AT_LOSS_DLR && (y_soft != nullptr || C < 3)) return (int)cudaErrorInvalidValue;
  LossArgs p{logits, y_hard, y_soft, dlogits, loss_out, state, loss_steps, (int)B, (int)C, iter, n_iter, ckpt_k,
             norm_kind, loss_kind, step_full, step_min, (float)n_fts};
  const int grid = (int)((B + 3) / 4);
  switch (logits_dtype) {
    case B200AT_DT_F32: loss_bookkeep_kernel<float><<<grid, 128, 0, s>>>(p); break;
    case B200AT_DT_BF16: loss_bookkeep_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p); break;
    case B200AT_DT_F16: loss_bookkeep_kernel<__half><<<grid, 128, 0, s>>>(p); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
