// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, 8 independent chains per thread.
// Prints FMA/clk/SM for both so kernels know whether pairing fp32 lanes buys FMA-pipe throughput.
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  if (MODE == 0) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fma1(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    u64 x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float v = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma2(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.5f); else k<1><<<sms, warps * 32>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)sms * warps * 32 * iters * 16;   // fp32 FMAs (both modes do 16 per thread per iteration)
      printf("warps/SM %2d  %s : %.3f ms  %.1f fp32-FMA/clk/SM (at %d MHz nominal)  %.1f TFLOP/s\n", warps, mode ? "FFMA2" : "FFMA ", ms,
             fma / (ms * 1e-3) / ((double)clk * 1e3) / sms, clk / 1000, 2 * fma / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
