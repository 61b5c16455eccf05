// Host-side launch helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace b200at {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: a process that
// touches a second GPU must set it there too, and a launch that needs more than an earlier one must raise it.
// `state` (one per kernel instantiation, zero-initialised) remembers the largest size set on each device ordinal;
// racing threads at worst both make the (idempotent) call.
struct SmemConfig {
  std::atomic<int> bytes[64];
};
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, SmemConfig& state) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::atomic<int>& cur = state.bytes[dev & 63];
  if (cur.load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) {
    int seen = cur.load(std::memory_order_relaxed);
    while (seen < bytes && !cur.compare_exchange_weak(seen, bytes, std::memory_order_release)) {}
  }
  return e;
}

}  // namespace b200at
