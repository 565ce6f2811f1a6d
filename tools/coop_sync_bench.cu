// Micro-benchmark: cost of grid-wide barriers on B200 (cooperative launch), cg::grid.sync() vs a hand-rolled
// monotonic-counter barrier.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_sync_bench coop_sync_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_cg(int nsync, float* out) {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < nsync; ++i) g.sync();
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f;
}

// monotonic counter barrier: `bar` must be 0 at kernel start (reset by the last CTA at the end)
__device__ __forceinline__ void grid_bar(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar));
    } while (v < target);
  }
  __syncthreads();
}
__global__ void k_hand(int nsync, unsigned* bar, float* out) {
  for (int i = 0; i < nsync; ++i) grid_bar(bar, (unsigned)(i + 1) * gridDim.x);
  // self-clean: last CTA out resets
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(bar + 1, 1u);
    if (t == gridDim.x - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
  }
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f;
}
__global__ void k_empty(float* out) { if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f; }

int main() {
  float* out; unsigned* bar;
  cudaMalloc(&out, 4); cudaMalloc(&bar, 8); cudaMemset(bar, 0, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 200;
  for (int grid : {148, 296, 592}) {
    for (int nsync : {0, 1, 2, 4}) {
      void* args[] = {&nsync, &out};
      for (int w = 0; w < 10; ++w) cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(256), args, 0, 0);
      cudaEventRecord(a);
      for (int i = 0; i < iters; ++i) cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(256), args, 0, 0);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      void* args2[] = {&nsync, &bar, &out};
      for (int w = 0; w < 10; ++w) cudaLaunchCooperativeKernel((void*)k_hand, dim3(grid), dim3(256), args2, 0, 0);
      cudaEventRecord(a);
      for (int i = 0; i < iters; ++i) cudaLaunchCooperativeKernel((void*)k_hand, dim3(grid), dim3(256), args2, 0, 0);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms2; cudaEventElapsedTime(&ms2, a, b);
      printf("grid %d nsync %d: cg %.2f us/launch, hand %.2f us/launch  (%s)\n", grid, nsync, ms * 1e3 / iters, ms2 * 1e3 / iters,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) k_empty<<<592, 256>>>(out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("plain empty launch: %.2f us\n", ms * 1e3 / iters);
  return 0;
}
