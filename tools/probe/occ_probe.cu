// occupancy probe: which resource stops two 256-thread CTAs from sharing an SM?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int REGS, bool TMEM>
__global__ void __launch_bounds__(256, 2) k(float* out, int n) {
  extern __shared__ uint8_t dyn[];
  __shared__ uint32_t slot;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = out[(threadIdx.x * 64 + i) % n];
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64) : "memory");
  }
  for (int it = 0; it < n; ++it)
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = acc[i] * acc[(i + 1) % 64] + dyn[(i + it) % 1024];
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
  out[threadIdx.x] = s;
}
template <class K>
void probe(const char* name, K kern) {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  for (size_t smem : {0ul, 48000ul, 96384ul}) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    printf("%s regs %d static %zu dyn %zu -> %d CTAs/SM\n", name, fa.numRegs, fa.sharedSizeBytes, smem, occ);
  }
}
int main() {
  probe("plain", k<0, false>);
  probe("tmem ", k<0, true>);
  int v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, 0); printf("regs/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxBlocksPerMultiprocessor, 0); printf("blocks/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerMultiProcessor, 0); printf("threads/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrReservedSharedMemoryPerBlock, 0); printf("reserved smem/block %d\n", v);
  return 0;
}
