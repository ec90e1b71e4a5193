// HBM write-only vs copy bandwidth on this GPU: conv0_apply writes 12.6 GB per launch and reads ~0.1 GB, so its
// roofline is the write-only figure, not the copy figure.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void fill2(uint4* a, uint4* b, size_t n16, int rows) {   // same pattern as conv0_apply: 2 streams, 16 B/thread
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) {
    a[i] = make_uint4(i, 1, 2, 3);
    b[i] = make_uint4(i, 4, 5, 6);
  }
}
__global__ void copy1(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) b[i] = a[i];
}
int main() {
  const size_t bytes = 6291456000ull;   // one hi (or lo) array of the 64 x 30 s batch: 64*95999*512*2 ~ 6.29 GB
  uint4 *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int grid = 148 * 2; grid <= 148 * 16; grid *= 2) {
    for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); fill2<<<grid, 256>>>(a, b, bytes / 16, 0); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("fill 2 x 6.29 GB grid %5d: %.3f ms  %.1f GB/s (write only)\n", grid, ms, 2.0 * bytes / ms / 1e6);
  }
  for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); cudaMemsetAsync(a, 1, bytes); cudaMemsetAsync(b, 2, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemset 2 x 6.29 GB: %.3f ms  %.1f GB/s (write only)\n", ms, 2.0 * bytes / ms / 1e6);
  for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); copy1<<<148 * 8, 256>>>(a, b, bytes / 16); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  cudaEventElapsedTime(&ms, e0, e1);
  printf("copy 6.29 GB -> 6.29 GB: %.3f ms  %.1f GB/s (read + write)\n", ms, 2.0 * bytes / ms / 1e6);
  for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpy D2D 6.29 GB: %.3f ms  %.1f GB/s (read + write)\n", ms, 2.0 * bytes / ms / 1e6);
  return 0;
}
