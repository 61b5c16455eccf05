// Which fp32 FMA forms does an sm_100a SM sustain?  (design input for the depthwise 7x7 kernel, which is
// FMA-issue bound.)  Every variant runs ILP independent accumulator chains per thread, 8 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_forms fma_forms.cu && ./fma_forms
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float cw[64];

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

constexpr int ILP = 16;
constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int uidx) {
  float a[ILP], b[ILP];
  float2 a2[ILP / 2], b2[ILP / 2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; }
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) { a2[i] = make_float2(a[2 * i], a[2 * i + 1]); b2[i] = make_float2(b[2 * i], b[2 * i + 1]); }
  const float x = in[threadIdx.x + 64];
  const float2 x2 = make_float2(x, in[threadIdx.x + 65]);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) a[i] = fmaf(x, b[i], a[i]);                       // 3 registers
      if (MODE == 1 && i < ILP / 2) a2[i] = ffma2(x2, b2[i], a2[i]);     // packed, 3 register pairs
      if (MODE == 2) a[i] = fmaf(x, cw[i], a[i]);                      // constant-bank operand, immediate offset
      if (MODE == 3) a[i] = fmaf(x, cw[uidx + i], a[i]);               // constant, warp-uniform runtime offset
      if (MODE == 4) a[i] = fmaf(a[i], 1.0001f, 0.5f);                 // immediate form
      if (MODE == 5 && i < ILP / 2) {                                  // packed with a broadcast constant pair
        a2[i] = ffma2(x2, make_float2(cw[2 * i], cw[2 * i + 1]), a2[i]);
      }
      if (MODE == 6) {                                                 // 3-reg FFMA interleaved with an ALU-pipe op
        a[i] = fmaf(x, b[i], a[i]);
        b[i] = __int_as_float(__float_as_int(b[i]) ^ (it & 1));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + b[i];
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) s += a2[i].x + a2[i].y + b2[i].x;
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread, float* out, float* in) {
  const int grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, in, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, in, 16);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double fmas = fma_per_thread * grid * 256.0;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-44s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at %d MHz max clock)\n", name, ms, fmas / ms / 1e9,
         fmas / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
}

int main() {
  float *out, *in;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpyToSymbol(cw, h, sizeof(h));
  const double n = (double)ILP * ITERS;
  run<0>("FFMA  reg,reg,reg", n, out, in);
  run<1>("FFMA2 regpair x3 (fma.rn.f32x2)", n, out, in);
  run<2>("FFMA  reg,c[imm],reg", n, out, in);
  run<3>("FFMA  reg,c[uniform runtime idx],reg", n, out, in);
  run<4>("FFMA  reg,imm,imm", n, out, in);
  run<5>("FFMA2 regpair,const pair,regpair", n, out, in);
  run<6>("FFMA reg x3 + LOP3 (alu pipe) interleaved", n, out, in);
  return 0;
}
