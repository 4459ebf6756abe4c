// Do CTAs of TWO kernels that both allocate tensor memory share an SM?  (cudaOccupancyMaxActiveBlocksPerMultiprocessor says 1 CTA/SM
// for a kernel with tcgen05.alloc; this asks the hardware.)  Two launches of 148 x 192 threads on two streams, each CTA allocates 64
// TMEM columns, notes its SM and spins for 2 ms.  Output: how many SMs held a CTA of each launch at the same time.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <bool TMEM>
__global__ void __launch_bounds__(192, 2) k(unsigned long long* out, int cols, unsigned long long spin_ns) {
  extern __shared__ uint8_t dyn[];
  __shared__ uint32_t slot;
  const unsigned long long t0 = gtime();
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  const unsigned long long t1 = gtime();
  dyn[threadIdx.x] = (uint8_t)threadIdx.x;
  while (gtime() - t1 < spin_ns) { __nanosleep(200); }
  const unsigned long long t2 = gtime();
  if (threadIdx.x == 0) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[blockIdx.x * 4 + 0] = smid; out[blockIdx.x * 4 + 1] = t0; out[blockIdx.x * 4 + 2] = t1; out[blockIdx.x * 4 + 3] = t2;
  }
  if (TMEM) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols) : "memory");
  }
}
template <bool TMEM>
void run(const char* name, int nk, size_t smem, int cols) {
  auto kern = k<TMEM>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 192, smem);
  std::vector<cudaStream_t> st(nk);
  std::vector<unsigned long long*> d(nk);
  for (int i = 0; i < nk; ++i) { cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking); cudaMalloc(&d[i], 148 * 4 * 8); cudaMemset(d[i], 0, 148 * 4 * 8); }
  cudaDeviceSynchronize();
  for (int i = 0; i < nk; ++i) kern<<<148, 192, smem, st[i]>>>(d[i], cols, 2000000ull);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<std::vector<unsigned long long>> h(nk, std::vector<unsigned long long>(148 * 4));
  for (int i = 0; i < nk; ++i) cudaMemcpy(h[i].data(), d[i], 148 * 4 * 8, cudaMemcpyDeviceToHost);
  // overlap: per SM, CTAs of launch 0 and launch j alive at the same time
  unsigned long long tmin = ~0ull, tmax = 0;
  for (int i = 0; i < nk; ++i) for (int c = 0; c < 148; ++c) { tmin = std::min(tmin, h[i][c * 4 + 1]); tmax = std::max(tmax, h[i][c * 4 + 3]); }
  int shared_sm = 0, same_kernel_double = 0;
  for (int sm = 0; sm < 160; ++sm) {
    bool ov = false;
    for (int c0 = 0; c0 < 148; ++c0) {
      if (h[0][c0 * 4] != (unsigned long long)sm) continue;
      for (int j = 1; j < nk; ++j) for (int c1 = 0; c1 < 148; ++c1)
        if (h[j][c1 * 4] == (unsigned long long)sm && h[j][c1 * 4 + 2] < h[0][c0 * 4 + 3] && h[0][c0 * 4 + 2] < h[j][c1 * 4 + 3]) ov = true;
      for (int c1 = c0 + 1; c1 < 148; ++c1)
        if (h[0][c1 * 4] == (unsigned long long)sm && h[0][c1 * 4 + 2] < h[0][c0 * 4 + 3] && h[0][c0 * 4 + 2] < h[0][c1 * 4 + 3]) ++same_kernel_double;
    }
    shared_sm += ov;
  }
  printf("%s: %d launches, smem %zu, tmem cols %d, occupancy API %d CTA/SM, err %s: total span %.3f ms (2.0 = concurrent, %d.0 = serial), SMs shared by launch 0 and another %d, SMs with two CTAs of launch 0: %d\n",
         name, nk, smem, cols, occ, cudaGetErrorString(e), (tmax - tmin) * 1e-6, 2 * nk, shared_sm, same_kernel_double);
}
int main() {
  run<false>("plain", 2, 100 * 1024, 0);
  run<true>("tmem ", 2, 100 * 1024, 64);
  run<true>("tmem ", 2, 100 * 1024, 256);
  run<true>("tmem ", 3, 64 * 1024, 64);
  run<true>("tmem ", 2, 100 * 1024, 512);
  return 0;
}
