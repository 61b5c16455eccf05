// Host-side launch helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace b200at {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: a process that
// touches a second GPU must set it there too.  One bit per device ordinal, set after the first successful call on
// that device; racing threads at worst both make the (idempotent) call.
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

}  // namespace b200at
