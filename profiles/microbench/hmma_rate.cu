// Microbenchmark: issue rate of the legacy tensor path (mma.sync.m16n8k16 bf16, SASS HMMA.16816.F32.BF16) on sm_100a,
// alone and interleaved with ldmatrix.x4 in the ratio of the Toeplitz depthwise-conv kernel (4 ldmatrix per 7 mma).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  __shared__ __align__(16) uint16_t tile[32 * 72 * 4];
  for (int i = threadIdx.x; i < 32 * 72 * 4; i += 256) tile[i] = (uint16_t)(0x3c00 + (i & 7));
  __syncthreads();
  float acc[7][4];
#pragma unroll
  for (int j = 0; j < 7; ++j) for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;
  uint32_t a[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) for (int q = 0; q < 4; ++q) a[j][q] = 0x3c003c00u + threadIdx.x + j;
  const uint32_t b0 = 0x3c003c00u, b1 = 0x3c003c01u;
  const int lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile) + ((lane & 15) * 72 + (lane >> 4) * 8) * 2;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ldsm4(a[j], base + j * 32 + (it & 7) * 144);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (MODE == 1) {
        const uint32_t af[4] = {a[j >> 1][(j & 1) * 2], a[j >> 1][(j & 1) * 2 + 1], a[(j + 1) >> 1][((j + 1) & 1) * 2], a[(j + 1) >> 1][((j + 1) & 1) * 2 + 1]};
        mma16816(acc[j], af, b0, b1);
      } else {
        mma16816(acc[j], a[j], b0, b1);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) for (int q = 0; q < 4; ++q) s += acc[j][q];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * 256);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * ctas_per_sm, 256>>>(out, 100);
  cudaEventRecord(e0);
  k<MODE><<<sms * ctas_per_sm, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * ctas_per_sm * 8 * iters * 7;
  printf("%-34s ctas/sm=%d  %.3f ms  %.1f G mma/s  = %.1f dense TFLOP/s (m16n8k16 = 4096 flop)  %.2f mma/us/SM\n", name, ctas_per_sm, ms,
         mmas / ms * 1e-6, mmas * 4096 / ms * 1e-9, mmas / ms * 1e-3 / sms);
  cudaFree(out);
}

int main() {
  run<0>("mma.sync only (7 chains/warp)", 1);
  run<0>("mma.sync only (7 chains/warp)", 2);
  run<1>("4 ldmatrix.x4 + 7 mma.sync", 1);
  run<1>("4 ldmatrix.x4 + 7 mma.sync", 2);
  return 0;
}
