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
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdlib.h>

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
__global__ void __launch_bounds__(kThreads, 8) linf_log_kernel(B200atSlots sl, const float* __restrict__ x,
                                                             float* __restrict__ x_new, const float* __restrict__ st,
                                                             int64_t B, int64_t n, int64_t nvec, float eps, float a,
                                                             float one_minus_a) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_linf_log_body<VEC>(sl, x, x_new, st, B, n, vi, eps, a, one_minus_a);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) gather_kernel(B200atSlots sl, float* __restrict__ x_best,
                                                           float* __restrict__ x_best_adv,
                                                           const float* __restrict__ st, int64_t B, int64_t n,
                                                           int64_t nvec) {
  const int64_t vi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (vi < nvec) b200at_gather_body<VEC>(sl, x_best, x_best_adv, st, B, n, vi);
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

// ------------------------------------------------------------------------------------------------
// K2: l2 update.  grid = (kChunks, B): a CTA reduces one contiguous chunk of one sample, so the three
// dependent per-sample norms are two-level, fixed-order (deterministic) sums: thread -> warp -> CTA
// partial in scratch[phase][b][chunk]; the next phase adds the kChunks partials with a warp butterfly.
constexpr int kChunks = 32;

template <int THREADS = kThreads>
__device__ __forceinline__ float cta_sum(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (THREADS / 32) ? smem[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// sum of the kChunks partials of `row` (fixed butterfly order => same value in every CTA of the sample)
__device__ __forceinline__ float chunk_total(const float* row) {
  float t = row[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

template <int PHASE, int VEC>
__global__ void __launch_bounds__(kThreads) l2_phase_kernel(B200atImages p, float* __restrict__ scratch, float eps,
                                                             float a, float one_minus_a) {
  __shared__ float red[kThreads / 32];
  const int b = blockIdx.y, c = blockIdx.x;
  const int64_t nvec_row = p.n / VEC;
  const int64_t per = (nvec_row + kChunks - 1) / kChunks;
  const int64_t v0 = c * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  float sums[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int ph = 0; ph < PHASE; ++ph) sums[ph] = chunk_total(scratch + ((int64_t)ph * p.B + b) * kChunks);
  float acc = 0.0f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads)
    acc += b200at_l2_body<PHASE, VEC>(p, (int64_t)b * nvec_row + v, eps, a, one_minus_a, sums);
  if (PHASE < 3) {
    const float t = cta_sum(acc, red);
    if (threadIdx.x == 0) scratch[((int64_t)PHASE * p.B + b) * kChunks + c] = t;
  }
}

// Single-launch form: a cluster of kL2Cluster CTAs owns one sample and runs the four phases back to back; the per-sample
// norms cross the cluster through distributed shared memory (one cluster barrier per norm) instead of through global
// scratch and a kernel boundary.  Each CTA sums a contiguous eighth of the sample (thread -> warp -> CTA, fixed order), the
// 8 partials are added in rank order by every thread: deterministic, but a different association than the four-launch
// form (the l2 path is tolerance-checked: 1e-6 on x).  The operands of a sample (4 x 602 KB at 224 px) are read from HBM
// once: phases 0..2 load with the L2-resident policy and the same CTA re-reads them microseconds later; occupancy is
// held at 2 CTAs per SM (37 samples x 2.4 MB in flight, inside the 126 MB L2) by the dynamic shared-memory request.
// HBM traffic 20 B/element instead of 52.
constexpr int kL2Cluster = 8;
// threads per CTA: 512 (2 CTAs x 16 warps per SM, four vectors in flight per thread) measured 140 us per step at
// 1024 x 3 x 224 x 224, 256 threads 161 us (profiles/r02_k1_driver_l2.txt); B200AT_L2_THREADS=256 selects the latter.
constexpr int kL2SmemCap = 100 * 1024;    // dynamic shared memory per CTA: the gradient slice; also bounds residency to 2 CTAs / SM

// `gbuf`: this CTA's slice of the gradient in shared memory (filled in phase 0, read by phases 1..3: 12 of the 52 B/element
// the four phases read come from there instead of from L2), or null when the slice does not fit (n / 8 floats > the buffer).
template <int PHASE, int VEC, int kL2Threads>
__device__ __forceinline__ float l2_cluster_phase(const B200atImages& p, int b, int rank, float eps, float a,
                                                  float one_minus_a, const float* sums, float* red, float* gbuf) {
  const int nvec_row = (int)(p.n / VEC);                 // n < 2^31 (checked by the entry point)
  const int per = (nvec_row + kL2Cluster - 1) / kL2Cluster;
  const int v0 = rank * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  const B200atL2Ctx ctx = b200at_l2_ctx<PHASE>(p, b, sums);
  const int64_t base = (int64_t)b * p.n;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  B200atVec<VEC>* gv = reinterpret_cast<B200atVec<VEC>*>(gbuf);
  // loads of a vector: the gradient from HBM (phase 0, streamed and parked in shared memory) or from shared memory
  auto load = [&](int v, B200atL2Ops<VEC>& in) {
    const int64_t e = base + (int64_t)v * VEC;
    if (gbuf == nullptr) { b200at_l2_load<PHASE, VEC, (PHASE < 3)>(p, e, ctx, nullptr, in); return; }
    if (PHASE == 0) {
      in.g = b200at_ld_stream<VEC>((ctx.restore ? p.grad_best : p.grad) + e);
      gv[v - v0] = in.g;
    } else {
      const B200atVec<VEC> g = gv[v - v0];
      b200at_l2_load<PHASE, VEC, (PHASE < 3)>(p, e, ctx, &g, in);
    }
  };
  int v = v0 + threadIdx.x;
  if (kL2Threads <= 256) {
    for (; v + 3 * kL2Threads < v1; v += 4 * kL2Threads) {    // four vectors: all loads first, then the (aliasing) stores
      B200atL2Ops<VEC> in[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) load(v + k * kL2Threads, in[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        acc[k] += b200at_l2_apply<PHASE, VEC>(p, base + (int64_t)(v + k * kL2Threads) * VEC, ctx, eps, a, one_minus_a, in[k]);
    }
  } else {
    // 512 threads leave 64 registers per thread: holding four vectors' operands spills (210 us); vector by vector, the
    // 32 warps of the SM supply the loads in flight (140 us)
    for (; v + 3 * kL2Threads < v1; v += 4 * kL2Threads) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        B200atL2Ops<VEC> in;
        load(v + k * kL2Threads, in);
        acc[k] += b200at_l2_apply<PHASE, VEC>(p, base + (int64_t)(v + k * kL2Threads) * VEC, ctx, eps, a, one_minus_a, in);
      }
    }
  }
  for (; v < v1; v += kL2Threads) {
    B200atL2Ops<VEC> in;
    load(v, in);
    acc[0] += b200at_l2_apply<PHASE, VEC>(p, base + (int64_t)v * VEC, ctx, eps, a, one_minus_a, in);
  }
  const float acc0 = acc[0], acc1 = acc[1], acc2 = acc[2], acc3 = acc[3];
  if (PHASE == 3) return 0.0f;
  return cta_sum<kL2Threads>((acc0 + acc1) + (acc2 + acc3), red);    // valid in warp 0
}

// the kL2Cluster partials of one phase, gathered from the cluster's CTAs and added in rank order
__device__ __forceinline__ float l2_cluster_total(cooperative_groups::cluster_group& cluster, float* part, int phase) {
  float t = 0.0f;
#pragma unroll
  for (int r = 0; r < kL2Cluster; ++r) t += cluster.map_shared_rank(part, r)[phase];
  return t;
}

template <int VEC, int kL2Threads>
__global__ void __launch_bounds__(kL2Threads) l2_cluster_kernel(B200atImages p, float eps, float a, float one_minus_a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float l2_gbuf[];         // kL2SmemCap bytes
  __shared__ float red[kL2Threads / 32];
  __shared__ float part[3];
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int64_t slice = ((p.n / VEC + kL2Cluster - 1) / kL2Cluster) * VEC;
  float* gbuf = (slice * 4 <= kL2SmemCap) ? l2_gbuf : nullptr;
  float sums[3] = {0.f, 0.f, 0.f};
  float t = l2_cluster_phase<0, VEC, kL2Threads>(p, b, rank, eps, a, one_minus_a, sums, red, gbuf);
  if (threadIdx.x == 0) part[0] = t;
  cluster.sync();
  sums[0] = l2_cluster_total(cluster, part, 0);
  t = l2_cluster_phase<1, VEC, kL2Threads>(p, b, rank, eps, a, one_minus_a, sums, red, gbuf);
  if (threadIdx.x == 0) part[1] = t;
  cluster.sync();
  sums[1] = l2_cluster_total(cluster, part, 1);
  t = l2_cluster_phase<2, VEC, kL2Threads>(p, b, rank, eps, a, one_minus_a, sums, red, gbuf);
  if (threadIdx.x == 0) part[2] = t;
  cluster.sync();
  sums[2] = l2_cluster_total(cluster, part, 2);
  l2_cluster_phase<3, VEC, kL2Threads>(p, b, rank, eps, a, one_minus_a, sums, red, gbuf);
  cluster.sync();   // nobody leaves while a peer may still be reading its partials
}

template <int VEC, int THREADS>
int launch_l2_cluster(const B200atImages& p, float eps, float a, float oma, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kL2Cluster, (unsigned)p.B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = kL2SmemCap;
  cfg.stream = s;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(l2_cluster_kernel<VEC, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL2SmemCap);
    if (e != cudaSuccess) return (int)e;
    configured[dev & 63] = true;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kL2Cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, l2_cluster_kernel<VEC, THREADS>, p, eps, a, oma);
}

template <int VEC>
int launch_l2(const B200atImages& p, float* scratch, float eps, float a, float oma, cudaStream_t s) {
  static const bool four_launches = [] { const char* e = getenv("B200AT_L2_PHASES"); return e != nullptr && e[0] == '4'; }();
  static const bool narrow = [] { const char* e = getenv("B200AT_L2_THREADS"); return e != nullptr && atoi(e) == 256; }();
  if (!four_launches) return narrow ? launch_l2_cluster<VEC, 256>(p, eps, a, oma, s) : launch_l2_cluster<VEC, 512>(p, eps, a, oma, s);
  dim3 grid(kChunks, (unsigned)p.B);
  l2_phase_kernel<0, VEC><<<grid, kThreads, 0, s>>>(p, scratch, eps, a, oma);
  l2_phase_kernel<1, VEC><<<grid, kThreads, 0, s>>>(p, scratch, eps, a, oma);
  l2_phase_kernel<2, VEC><<<grid, kThreads, 0, s>>>(p, scratch, eps, a, oma);
  l2_phase_kernel<3, VEC><<<grid, kThreads, 0, s>>>(p, scratch, eps, a, oma);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3 + K4: l1 update.  Per sample: exact k-th order statistic of |grad| by a 3-level radix select
// (11+11+9 bits, shared-memory histograms merged with integer atomics => deterministic), then the
// l1-ball ∩ box projection by sectioning the bit pattern of the water level (7 passes x 31 candidates),
// then the final write.  grid = (kChunks, B) for the image passes; tiny per-sample kernels in between.
// Scratch words per sample: see B200AT_L1_SCRATCH_WORDS in include/b200at.h.
constexpr int kHistBins = 2048;
constexpr int kMetaWords = 32;
// per-sample meta words.  The scan of radix level L and the decision of sectioning pass P are made in the PROLOGUE of the
// next image pass by every CTA of the sample (redundantly: 8 KB of histogram / 4 KB of partial sums from L2) instead of by
// 1-warp helper launches in between; CTA 0 of the sample publishes the result for the launches after that.  Every result
// has its own slot, so a CTA publishing slot L never races with a slower CTA of the same launch still reading slot L - 1.
enum { M_PREFIX = 0 /*..2*/, M_RANK = 3 /*..5*/, M_COUNT_LT = 6 /*..8*/, M_NAN = 9, M_ZERO = 10, M_THR = 11, M_NNZ = 12,
       M_C = 13, M_NEED = 14, M_ALPHA = 16 /*..22: water-level prefix after pass 0..6*/ };

struct L1Scratch {
  int* hist;      // [3][B][kHistBins]
  int* meta;      // [B][kMetaWords]
  float* sums;    // [B][kChunks][2]
  float* sect;    // [2][B][kChunks][32]   (double-buffered by pass parity: pass p+1's prologue reads what pass p wrote)
};
__host__ __device__ inline L1Scratch l1_scratch(void* base, int64_t B) {
  L1Scratch s;
  s.hist = reinterpret_cast<int*>(base);
  s.meta = s.hist + 3 * B * kHistBins;
  s.sums = reinterpret_cast<float*>(s.meta + B * kMetaWords);
  s.sect = s.sums + B * kChunks * 2;
  return s;
}

struct L1Sel {  // per-sample selection of the current iterate / gradient (pending restore)
  const float* xc; const float* g; bool improved, write_adv, restore; float step;
};
__device__ __forceinline__ L1Sel l1_select(const B200atImages& p, int b) {
  const int32_t fl = b200at_f2i(p.st[(int64_t)B200AT_ST_FLAGS * p.B + b]);
  L1Sel s;
  s.improved = fl & B200AT_F_IMPROVED;
  s.write_adv = fl & B200AT_F_WRITE_ADV;
  s.restore = (fl & B200AT_F_RESTORE) && !s.improved;
  s.xc = s.restore ? p.x_best : p.x_adv;
  s.g = s.restore ? p.grad_best : p.grad;
  s.step = p.st[(int64_t)B200AT_ST_STEP * p.B + b];
  return s;
}

struct L1Scan { uint32_t prefix; int rank, count_lt; };

// Locate the bin of radix level LEVEL that holds the wanted rank (whole CTA, kThreads threads; every thread returns the
// same result).  Level 0 starts from the rank the top-k fraction asks for, levels 1-2 from the published level below.
template <int LEVEL>
__device__ __forceinline__ L1Scan l1_scan_dev(const float* __restrict__ st, const L1Scratch& S, int64_t B, int64_t n, int b,
                                              int* sm /* >= kThreads/32 + 2 ints */) {
  const int* meta = S.meta + b * kMetaWords;
  const int* gh = S.hist + ((int64_t)LEVEL * B + b) * kHistBins;
  int rank, prev_clt = 0;
  uint32_t prev = 0u;
  if (LEVEL == 0) rank = (int)b200at_l1_rank(st[(int64_t)B200AT_ST_TOPK * B + b], n);
  else { rank = meta[M_RANK + LEVEL - 1]; prev = (uint32_t)meta[M_PREFIX + LEVEL - 1]; prev_clt = meta[M_COUNT_LT + LEVEL - 1]; }
  constexpr int per = kHistBins / kThreads;  // 8 consecutive bins per thread
  int loc[per], tot = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) { loc[i] = gh[threadIdx.x * per + i]; tot += loc[i]; }
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  int* warp_tot = sm;
  int* found = sm + kThreads / 32;           // [0] bin, [1] exclusive count
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  if (threadIdx.x == 0) { found[0] = -1; found[1] = 0; }
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) base += warp_tot[w];
  int excl = base + incl - tot;
  if (rank >= excl && rank < excl + tot) {
#pragma unroll
    for (int i = 0; i < per; ++i) {
      if (rank < excl + loc[i]) { found[0] = threadIdx.x * per + i; found[1] = excl; break; }
      excl += loc[i];
    }
  }
  __syncthreads();
  const int bin = found[0] < 0 ? 0 : found[0];
  const int ex = found[0] < 0 ? 0 : found[1];
  __syncthreads();                           // `sm` is reused by the caller
  L1Scan r;
  r.prefix = LEVEL == 0 ? (uint32_t)bin : (LEVEL == 1 ? ((prev << 11) | bin) : ((prev << 9) | bin));
  r.rank = rank - ex;
  r.count_lt = prev_clt + ex;
  return r;
}
__device__ __forceinline__ void l1_publish_scan(int* meta, int level, const L1Scan& r) {
  meta[M_PREFIX + level] = (int)r.prefix;
  meta[M_RANK + level] = r.rank;
  meta[M_COUNT_LT + level] = r.count_lt;
}

template <int LEVEL, int VEC>
__global__ void __launch_bounds__(kThreads) l1_hist_kernel(B200atImages p, void* scratch, float* st_rw) {
  __shared__ int h[kHistBins];
  __shared__ int scan_sm[kThreads / 32 + 2];
  const int b = blockIdx.y, c = blockIdx.x;
  const L1Scratch S = l1_scratch(scratch, p.B);
  const L1Sel sel = l1_select(p, b);
  uint32_t prefix = 0u;
  if (LEVEL == 0) {
    if (c == 0 && threadIdx.x == 0) *reinterpret_cast<int*>(st_rw + (int64_t)B200AT_ST_SP_ADV * p.B + b) = 0;
  } else {
    const L1Scan r = l1_scan_dev<LEVEL - 1>(p.st, S, p.B, p.n, b, scan_sm);
    prefix = r.prefix;
    if (c == 0 && threadIdx.x == 0) l1_publish_scan(S.meta + b * kMetaWords, LEVEL - 1, r);
  }
  for (int i = threadIdx.x; i < kHistBins; i += kThreads) h[i] = 0;
  __syncthreads();
  const int64_t nvec_row = p.n / VEC;
  const int64_t per = (nvec_row + kChunks - 1) / kChunks;
  const int64_t v0 = c * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  int nan_cnt = 0, zero_cnt = 0;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    const B200atVec<VEC> g = b200at_ld_stream<VEC>(sel.g + ((int64_t)b * nvec_row + v) * VEC);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const uint32_t key = b200at_l1_key(g.v[i]);
      if (LEVEL == 0) {
        atomicAdd(&h[key >> 20], 1);
        nan_cnt += key > 0x7f800000u;
        zero_cnt += key == 0u;
      } else if (LEVEL == 1) {
        if ((key >> 20) == prefix) atomicAdd(&h[(key >> 9) & 0x7ff], 1);
      } else {
        if ((key >> 9) == prefix) atomicAdd(&h[key & 0x1ff], 1);
      }
    }
  }
  __syncthreads();
  int* gh = S.hist + ((int64_t)LEVEL * p.B + b) * kHistBins;
  for (int i = threadIdx.x; i < kHistBins; i += kThreads)
    if (h[i]) atomicAdd(&gh[i], h[i]);
  if (LEVEL == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nan_cnt += __shfl_xor_sync(0xffffffffu, nan_cnt, o);
      zero_cnt += __shfl_xor_sync(0xffffffffu, zero_cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (nan_cnt) atomicAdd(&S.meta[b * kMetaWords + M_NAN], nan_cnt);
      if (zero_cnt) atomicAdd(&S.meta[b * kMetaWords + M_ZERO], zero_cnt);
    }
  }
}

// water-level prefix after sectioning pass `pass` (:71-88): keep the candidates still below the level.  Every warp of every
// CTA of the sample evaluates it from the 32 chunk partials of that pass (lane k = candidate k + 1, fixed chunk order).
__device__ __forceinline__ uint32_t l1_decide_dev(const L1Scratch& S, int64_t B, int b, int pass, float cc) {
  const int k = threadIdx.x & 31;
  const int* meta = S.meta + b * kMetaWords;
  const uint32_t prev = pass == 0 ? 0u : (uint32_t)meta[M_ALPHA + pass - 1];
  const int ncand = b200at_l1_ncand(pass);
  const float* sect = S.sect + ((int64_t)(pass & 1) * B + b) * kChunks * 32;
  float gk = 0.0f;
  if (k < ncand)
    for (int c = 0; c < kChunks; ++c) gk += sect[c * 32 + k];
  const bool below = (k < ncand) && (gk + cc < 0.0f);
  const unsigned m = __ballot_sync(0xffffffffu, below);
  const int kstar = __ffs(~m) - 1;  // number of leading candidates still below the level
  return prev | ((uint32_t)kstar << b200at_l1_shift(pass));
}

// sum|y| and sum(a) partials (prologue: the scan of the last radix level -> threshold and nnz, published by CTA 0)
template <int VEC>
__global__ void __launch_bounds__(kThreads) l1_sums_kernel(B200atImages p, void* scratch) {
  __shared__ float red[kThreads / 32];
  __shared__ int scan_sm[kThreads / 32 + 2];
  const int b = blockIdx.y, c = blockIdx.x;
  const L1Scratch S = l1_scratch(scratch, p.B);
  int* meta = S.meta + b * kMetaWords;
  const L1Scan r = l1_scan_dev<2>(p.st, S, p.B, p.n, b, scan_sm);
  const float thr = b200at_i2f((int32_t)r.prefix);
  float nnz = 0.0f;
  if (r.prefix <= 0x7f800000u) nnz = (float)((int)p.n - r.count_lt - meta[M_NAN] - (r.prefix == 0u ? meta[M_ZERO] : 0));
  if (c == 0 && threadIdx.x == 0) { l1_publish_scan(meta, 2, r); meta[M_THR] = b200at_f2i(thr); meta[M_NNZ] = b200at_f2i(nnz); }
  const L1Sel sel = l1_select(p, b);
  const int64_t nvec_row = p.n / VEC;
  const int64_t per = (nvec_row + kChunks - 1) / kChunks;
  const int64_t v0 = c * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  float sb = 0.0f, sa = 0.0f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    const int64_t e = ((int64_t)b * nvec_row + v) * VEC;
    const B200atVec<VEC> x = b200at_ld_stream<VEC>(p.x + e), xc = b200at_ld_stream<VEC>(sel.xc + e),
                         g = b200at_ld_stream<VEC>(sel.g + e);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float y = b200at_l1_y(x.v[i], xc.v[i], g.v[i], sel.step, thr, nnz);
      sb += fabsf(y);
      sa -= b200at_l1_u(x.v[i], y);
    }
  }
  const float tb = cta_sum(sb, red);
  const float ta = cta_sum(sa, red);
  if (threadIdx.x == 0) {
    S.sums[((int64_t)b * kChunks + c) * 2 + 0] = tb;
    S.sums[((int64_t)b * kChunks + c) * 2 + 1] = ta;
  }
}

// Sectioning pass `pass`: g(alpha_k) = sum_i clamp(alpha_k, a_i, b_i) for the 31 candidates of this pass; samples that need
// no l1 shrink are skipped.  Prologue: pass 0 derives c = eps - sum|y| and need = (sum(a) + c < 0) (:49-52) from the partial
// sums, later passes the water-level prefix from the previous pass's partial sums.
template <int VEC>
__global__ void __launch_bounds__(kThreads) l1_section_kernel(B200atImages p, void* scratch, int pass, float eps) {
  __shared__ float red[kThreads / 32];
  const int b = blockIdx.y, c = blockIdx.x;
  const L1Scratch S = l1_scratch(scratch, p.B);
  int* meta = S.meta + b * kMetaWords;
  const bool pub = c == 0 && threadIdx.x == 0;
  const float thr = b200at_i2f(meta[M_THR]), nnz = b200at_i2f(meta[M_NNZ]);
  uint32_t prefix = 0u;
  if (pass == 0) {
    float tb = S.sums[((int64_t)b * kChunks + (threadIdx.x & 31)) * 2 + 0];
    float ta = S.sums[((int64_t)b * kChunks + (threadIdx.x & 31)) * 2 + 1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tb += __shfl_xor_sync(0xffffffffu, tb, o);
      ta += __shfl_xor_sync(0xffffffffu, ta, o);
    }
    const float cc = eps - tb;
    const int need = (ta + cc < 0.0f) ? 1 : 0;
    if (pub) { meta[M_C] = b200at_f2i(cc); meta[M_NEED] = need; }
    if (!need) return;
  } else {
    if (!meta[M_NEED]) return;
    prefix = l1_decide_dev(S, p.B, b, pass - 1, b200at_i2f(meta[M_C]));
    if (pub) meta[M_ALPHA + pass - 1] = (int)prefix;
  }
  const L1Sel sel = l1_select(p, b);
  const int ncand = b200at_l1_ncand(pass);
  float cand[31], acc[31];
#pragma unroll
  for (int k = 0; k < 31; ++k) { cand[k] = b200at_l1_cand(prefix, k + 1, pass); acc[k] = 0.0f; }
  const int64_t nvec_row = p.n / VEC;
  const int64_t per = (nvec_row + kChunks - 1) / kChunks;
  const int64_t v0 = c * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    const int64_t e = ((int64_t)b * nvec_row + v) * VEC;
    const B200atVec<VEC> x = b200at_ld_stream<VEC>(p.x + e), xc = b200at_ld_stream<VEC>(sel.xc + e),
                         g = b200at_ld_stream<VEC>(sel.g + e);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float y = b200at_l1_y(x.v[i], xc.v[i], g.v[i], sel.step, thr, nnz);
      const float a = -b200at_l1_u(x.v[i], y), bb = fabsf(y);
#pragma unroll
      for (int k = 0; k < 31; ++k) acc[k] += b200at_l1_level(cand[k], a, bb);
    }
  }
  float* out = S.sect + (((int64_t)(pass & 1) * p.B + b) * kChunks + c) * 32;
#pragma unroll
  for (int k = 0; k < 31; ++k) {
    if (k < ncand) {
      const float t = cta_sum(acc[k], red);
      if (threadIdx.x == 0) out[k] = t;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) l1_final_kernel(B200atImages p, void* scratch, float* st_rw) {
  const int b = blockIdx.y, c = blockIdx.x;
  const L1Scratch S = l1_scratch(scratch, p.B);
  const L1Sel sel = l1_select(p, b);
  const int* meta = S.meta + b * kMetaWords;
  const float thr = b200at_i2f(meta[M_THR]), nnz = b200at_i2f(meta[M_NNZ]);
  const int need = meta[M_NEED];
  float alpha = 0.0f;
  if (need) alpha = b200at_i2f((int32_t)l1_decide_dev(S, p.B, b, B200AT_L1_PASSES - 1, b200at_i2f(meta[M_C])));
  const int64_t nvec_row = p.n / VEC;
  const int64_t per = (nvec_row + kChunks - 1) / kChunks;
  const int64_t v0 = c * per, v1 = (v0 + per < nvec_row) ? v0 + per : nvec_row;
  int moved = 0;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    const int64_t e = ((int64_t)b * nvec_row + v) * VEC;
    const B200atVec<VEC> x = b200at_ld_stream<VEC>(p.x + e), xc = b200at_ld_stream<VEC>(sel.xc + e),
                         g = b200at_ld_stream<VEC>(sel.g + e);
    if (!sel.restore) {
      if (sel.write_adv) b200at_st_stream<VEC>(p.x_best_adv + e, xc);
      if (sel.improved) {
        b200at_st_stream<VEC>(p.x_best + e, xc);
        b200at_st_stream<VEC>(p.grad_best + e, g);
      }
    } else if (sel.write_adv) {
      b200at_st_stream<VEC>(p.x_best_adv + e, b200at_ld_stream<VEC>(p.x_adv + e));
    }
    B200atVec<VEC> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float y = b200at_l1_y(x.v[i], xc.v[i], g.v[i], sel.step, thr, nnz);
      o.v[i] = b200at_l1_out(x.v[i], y, b200at_l1_u(x.v[i], y), need, alpha);
      moved += (B200AT_SUB(o.v[i], x.v[i]) != 0.0f);
    }
    b200at_st_keep<VEC>(p.x_new + e, o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) moved += __shfl_xor_sync(0xffffffffu, moved, o);
  if ((threadIdx.x & 31) == 0 && moved)
    atomicAdd(reinterpret_cast<int*>(st_rw + (int64_t)B200AT_ST_SP_ADV * p.B + b), moved);
}

// 12 kernel launches (+ one memset) per move: three histogram passes, the sums pass, seven sectioning passes, the final
// pass.  It was 23 launches + 2 memsets: three scan, one need and seven decide launches of one CTA / one warp per
// sample sat in between.
template <int VEC>
int launch_l1(const B200atImages& p, void* scratch, float* st_rw, float eps, cudaStream_t s) {
  const L1Scratch S = l1_scratch(scratch, p.B);
  cudaError_t e = cudaMemsetAsync(S.hist, 0, sizeof(int) * (3 * kHistBins + kMetaWords) * p.B, s);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(kChunks, (unsigned)p.B);
  l1_hist_kernel<0, VEC><<<grid, kThreads, 0, s>>>(p, scratch, st_rw);
  l1_hist_kernel<1, VEC><<<grid, kThreads, 0, s>>>(p, scratch, st_rw);
  l1_hist_kernel<2, VEC><<<grid, kThreads, 0, s>>>(p, scratch, st_rw);
  l1_sums_kernel<VEC><<<grid, kThreads, 0, s>>>(p, scratch);
  for (int pass = 0; pass < B200AT_L1_PASSES; ++pass) l1_section_kernel<VEC><<<grid, kThreads, 0, s>>>(p, scratch, pass, eps);
  l1_final_kernel<VEC><<<grid, kThreads, 0, s>>>(p, scratch, st_rw);
  return (int)cudaGetLastError();
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
  const int64_t* y_target;  // [B] or null (targeted DLR only)
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
  else {
    const int64_t yb = p.y_hard[b];
    // a class index outside [0, C) is a caller bug; torch's CUDA cross_entropy (the reference's :113) raises a
    // device-side assert for it, and DLR below would index z[label] out of bounds: fail as loudly, no host sync
    if (yb < 0 || yb >= (int64_t)p.C) __trap();
    label = (int)yb;
  }
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
    // targeted DLR (:106-111): -(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12); one more sweep for z_(4)
    const bool targeted = p.loss_kind == B200AT_LOSS_DLR_TARGETED;
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
    if (targeted) {
      float m4 = ninf; int i4 = 0x7fffffff;
      for (int c = lane; c < p.C; c += 32) {
        const float v = to_f32<T>(z[c]);
        if (c != i1 && c != i2 && c != i3 && better(v, c, m4, i4)) { m4 = v; i4 = c; }
      }
      warp_argmax(m4, i4);
      const int64_t tb = p.y_target[b];
      if (tb < 0 || tb >= (int64_t)p.C) __trap();
      const int it = (int)tb;
      const float num = zy - to_f32<T>(z[it]);
      const float den = (m1 - 0.5f * (m3 + m4)) + 1e-12f;
      loss = -num / den;
      if (dz) {
        // d(-num/den) = -(e_y - e_t)/den + num/den^2 * (e_1 - e_3/2 - e_4/2)
        const float q = num / (den * den);
        for (int c = lane; c < p.C; c += 32) {
          float gz = 0.0f;
          if (c == label) gz -= 1.0f / den;
          if (c == it) gz += 1.0f / den;
          if (c == i1) gz += q;
          if (c == i3) gz -= 0.5f * q;
          if (c == i4) gz -= 0.5f * q;
          dz[c] = from_f32<T>(gz);
        }
      }
    } else {
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
  }
  if (lane == 0) {
    if (p.loss_out) p.loss_out[b] = loss;
    b200at_bookkeep_sample(p.st, p.loss_steps, p.B, b, loss, pred, p.iter, p.n_iter, p.ckpt_k, p.norm_kind,
                           p.step_full, p.step_min, p.n_fts, p.dlogits != nullptr);
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

static bool fill_slots(B200atSlots& sl, const float* const* xs, const float* const* gs, int n_slots, bool& v4) {
  if (n_slots < 1 || n_slots > B200AT_LOG_MAX_SLOTS) return false;
  for (int i = 0; i < B200AT_LOG_MAX_SLOTS; ++i) {
    sl.x[i] = xs[i < n_slots ? i : 0];
    sl.g[i] = gs ? gs[i < n_slots ? i : 0] : nullptr;
    v4 = v4 && aligned16(sl.x[i]) && (!gs || aligned16(sl.g[i]));
  }
  return true;
}

int b200at_linf_step_log(const float* x, const float* const* x_slots, const float* const* g_slots, int n_slots,
                         float* x_new, const float* state, int64_t B, int64_t n, float eps, float a, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  B200atSlots sl;
  bool v4 = (n % 4 == 0) && aligned16(x) && aligned16(x_new);
  if (!fill_slots(sl, x_slots, g_slots, n_slots, v4)) return (int)cudaErrorInvalidValue;
  const float oma = (float)(1.0 - (double)a);
  const int64_t nvec = v4 ? B * n / 4 : B * n;
  if (v4) linf_log_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(sl, x, x_new, state, B, n, nvec, eps, a, oma);
  else linf_log_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(sl, x, x_new, state, B, n, nvec, eps, a, oma);
  return (int)cudaGetLastError();
}

int b200at_gather_best(const float* const* x_slots, int n_slots, float* x_best, float* x_best_adv,
                       const float* state, int64_t B, int64_t n, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  B200atSlots sl;
  bool v4 = (n % 4 == 0) && aligned16(x_best) && aligned16(x_best_adv);
  if (!fill_slots(sl, x_slots, nullptr, n_slots, v4)) return (int)cudaErrorInvalidValue;
  const int64_t nvec = v4 ? B * n / 4 : B * n;
  if (v4) gather_kernel<4><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(sl, x_best, x_best_adv, state, B, n, nvec);
  else gather_kernel<1><<<grid_for(nvec, kThreads), kThreads, 0, s>>>(sl, x_best, x_best_adv, state, B, n, nvec);
  return (int)cudaGetLastError();
}

int b200at_l2_step(const float* x, float* x_adv, const float* x_old, float* x_new, const float* grad,
                   float* x_best, float* grad_best, float* x_best_adv, const float* state, float* scratch,
                   int64_t B, int64_t n, float eps, float a, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  if (B > 65535 || n >= (int64_t)1 << 31) return (int)cudaErrorInvalidValue;
  B200atImages p{x, x_adv, x_old, x_new, grad, x_best, grad_best, x_best_adv, state, B, n};
  const float oma = (float)(1.0 - (double)a);
  const bool v4 = (n % 4 == 0) && aligned16(x) && aligned16(x_adv) && aligned16(x_old) && aligned16(x_new) &&
                  aligned16(grad) && aligned16(x_best) && aligned16(grad_best) && aligned16(x_best_adv);
  return v4 ? launch_l2<4>(p, scratch, eps, a, oma, s) : launch_l2<1>(p, scratch, eps, a, oma, s);
}

int b200at_l1_step(const float* x, float* x_adv, float* x_new, const float* grad, float* x_best, float* grad_best,
                   float* x_best_adv, float* state, void* scratch, int64_t B, int64_t n, float eps, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0 || n <= 0) return (int)cudaSuccess;
  if (B > 65535 || n >= (int64_t)1 << 31) return (int)cudaErrorInvalidValue;
  B200atImages p{x, x_adv, nullptr, x_new, grad, x_best, grad_best, x_best_adv, state, B, n};
  const bool v4 = (n % 4 == 0) && aligned16(x) && aligned16(x_adv) && aligned16(x_new) && aligned16(grad) &&
                  aligned16(x_best) && aligned16(grad_best) && aligned16(x_best_adv);
  return v4 ? launch_l1<4>(p, scratch, state, eps, s) : launch_l1<1>(p, scratch, state, eps, s);
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

static int launch_loss_bookkeep(const void* logits, int logits_dtype, const int64_t* y_hard, const float* y_soft,
                                const int64_t* y_target, void* dlogits, float* loss_out, float* state,
                                float* loss_steps, int64_t B, int64_t C, int iter, int n_iter, int ckpt_k,
                                int norm_kind, int loss_kind, float step_full, float step_min, int64_t n_fts,
                                void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (B <= 0) return (int)cudaSuccess;
  if ((y_hard == nullptr) == (y_soft == nullptr)) return (int)cudaErrorInvalidValue;
  if (loss_kind == B200AT_LOSS_DLR && (y_soft != nullptr || C < 3)) return (int)cudaErrorInvalidValue;
  if (loss_kind == B200AT_LOSS_DLR_TARGETED && (y_soft != nullptr || y_target == nullptr || C < 4))
    return (int)cudaErrorInvalidValue;
  if (loss_kind < B200AT_LOSS_CE || loss_kind > B200AT_LOSS_DLR_TARGETED) return (int)cudaErrorInvalidValue;
  LossArgs p{logits, y_hard, y_soft, y_target, dlogits, loss_out, state, loss_steps, (int)B, (int)C, iter, n_iter,
             ckpt_k, norm_kind, loss_kind, step_full, step_min, (float)n_fts};
  const int grid = (int)((B + 3) / 4);
  switch (logits_dtype) {
    case B200AT_DT_F32: loss_bookkeep_kernel<float><<<grid, 128, 0, s>>>(p); break;
    case B200AT_DT_BF16: loss_bookkeep_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p); break;
    case B200AT_DT_F16: loss_bookkeep_kernel<__half><<<grid, 128, 0, s>>>(p); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

int b200at_loss_bookkeep(const void* logits, int logits_dtype, const int64_t* y_hard, const float* y_soft,
                         void* dlogits, float* loss_out, float* state, float* loss_steps, int64_t B, int64_t C,
                         int iter, int n_iter, int ckpt_k, int norm_kind, int loss_kind, float step_full,
                         float step_min, int64_t n_fts, void* stream) {
  if (loss_kind == B200AT_LOSS_DLR_TARGETED) return (int)cudaErrorInvalidValue;   // needs the targeted entry point
  return launch_loss_bookkeep(logits, logits_dtype, y_hard, y_soft, nullptr, dlogits, loss_out, state, loss_steps, B, C,
                              iter, n_iter, ckpt_k, norm_kind, loss_kind, step_full, step_min, n_fts, stream);
}

int b200at_loss_bookkeep_targeted(const void* logits, int logits_dtype, const int64_t* y_hard, const int64_t* y_target,
                                  void* dlogits, float* loss_out, float* state, float* loss_steps, int64_t B,
                                  int64_t C, int iter, int n_iter, int ckpt_k, int norm_kind, float step_full,
                                  float step_min, int64_t n_fts, void* stream) {
  return launch_loss_bookkeep(logits, logits_dtype, y_hard, nullptr, y_target, dlogits, loss_out, state, loss_steps, B,
                              C, iter, n_iter, ckpt_k, norm_kind, B200AT_LOSS_DLR_TARGETED, step_full, step_min, n_fts,
                              stream);
}

}  // extern "C"
