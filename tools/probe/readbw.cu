// Read-only HBM bandwidth ceiling: every CTA streams its own slice of a large buffer with 16-byte loads
// (8 independent loads per thread in flight), sum reduced to defeat dead-code elimination.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 2) rd(const uint4* __restrict__ p, size_t n, unsigned* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (; i + 7 * st < n; i += 8 * st) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcs(p + i + j * st);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) *out = acc;
}
int main() {
  const size_t bytes = (size_t)16 << 30;
  uint4* p; unsigned* o;
  cudaMalloc(&p, bytes); cudaMalloc(&o, 4); cudaMemset(p, 1, bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int grid : {148 * 2, 148 * 4, 148 * 8}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      rd<<<grid, 1024>>>(p, bytes / 16, o);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 2) printf("grid %d: %.3f ms, %.1f GB/s read-only\n", grid, ms, bytes / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
