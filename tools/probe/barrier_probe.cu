// grid-barrier microbenchmark: 148 CTAs x 256 threads (optionally 2 groups x 148 CTAs, 2 per SM), N barriers per launch
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
template <int V>
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target, int nc) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += (unsigned)nc;
    unsigned v;
    if (V == 0) {  // current: red.release + relaxed poll + acq_rel fence + all-proxy fence
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    } else if (V == 1) {  // no proxy fence
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else if (V == 2) {  // acquire poll, no fences
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
    } else if (V == 3) {  // relaxed red + relaxed poll (no ordering at all: lower bound)
      asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
    } else if (V == 4) {  // acquire poll + proxy fence
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
      asm volatile("fence.proxy.async;" ::: "memory");
    } else if (V == 5) {  // last arriver publishes a flag in another line; everyone polls the flag
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
      if (old + 1 == target) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 64), "r"(target) : "memory");
      else do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar + 64) : "memory"); } while ((int)(v - target) < 0);
    }
  }
  __syncthreads();
}
template <int V>
__global__ void __launch_bounds__(256, 2) k(unsigned* bars, unsigned* seats, float* data, int n, int nc, int work) {
  __shared__ int seat[2];
  if (threadIdx.x == 0) {
    int grp = 0, cta = blockIdx.x;
    if (gridDim.x > (unsigned)nc) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      grp = atomicAdd(seats + smid, 1u) & 1;
      for (;;) { cta = atomicAdd(seats + 256 + grp, 1u); if (cta < nc) break; grp ^= 1; }
    }
    seat[0] = grp; seat[1] = cta;
  }
  __syncthreads();
  unsigned* bar = bars + 256 * seat[0];
  unsigned target = 0;
  float* mine = data + ((size_t)seat[0] * nc + seat[1]) * 256;
  for (int i = 0; i < n; ++i) {
    if (work) mine[threadIdx.x] = (float)i;  // a store per thread before the barrier, as a real phase has
    grid_sync<V>(bar, target, nc);
  }
}
template <int V>
void run(const char* name, int groups, int work) {
  unsigned *bars, *seats; float* data;
  cudaMalloc(&bars, 4096); cudaMalloc(&seats, 2048); cudaMalloc(&data, 2 * 148 * 256 * 4);
  const int n = 2000, nc = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(bars, 0, 4096); cudaMemset(seats, 0, 2048);
    void* args[] = {&bars, &seats, &data, (void*)&n, (void*)&nc, &work};
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k<V>, dim3(nc * groups), dim3(256), args, 0, 0);
    cudaEventRecord(e1);
    if (e != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("%s: run failed %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-44s groups %d work %d: %.3f us per barrier\n", name, groups, work, best * 1e3 / n);
  cudaFree(bars); cudaFree(seats); cudaFree(data);
}
int main() {
  for (int groups = 1; groups <= 2; ++groups)
    for (int work = 0; work <= 1; ++work) {
      run<0>("V0 red.release, relaxed poll, acq_rel+proxy", groups, work);
      run<1>("V1 same, no proxy fence", groups, work);
      run<2>("V2 red.release, acquire poll", groups, work);
      run<3>("V3 relaxed everything (lower bound)", groups, work);
      run<4>("V4 red.release, acquire poll, proxy fence", groups, work);
      run<5>("V5 atom + flag line", groups, work);
    }
  return 0;
}
