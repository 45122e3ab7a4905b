// Microbenchmark: cost of one software grid barrier on 113 / 148 CTAs of 512 threads (the PCG kernel's shape), variants.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o grid_barrier grid_barrier.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

template <int V>
__global__ void __launch_bounds__(512, 1) k_bar(unsigned* bar, double* data, int iters) {
  unsigned target = 0;
  cg::grid_group grid = cg::this_grid();
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    // a little "work": every thread publishes a value others read after the barrier
    if (threadIdx.x < 6) data[blockIdx.x * 6 + threadIdx.x] = acc + it;
    if (V == 0) {                      // fence + atomicAdd (returning) + volatile poll + fence
      __syncthreads();
      if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*(volatile unsigned*)bar < target) {}
        __threadfence();
      }
      __syncthreads();
    } else if (V == 1) {               // fence + red + relaxed poll + fence
      __syncthreads();
      if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (ld_relaxed(bar) < target) {}
        __threadfence();
      }
      __syncthreads();
    } else if (V == 2) {               // red.release + acquire poll (no explicit fences)
      __syncthreads();
      if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (ld_acquire(bar) < target) {}
      }
      __syncthreads();
    } else if (V == 3) {               // red.release + relaxed poll, no trailing fence (readers use ld.cg)
      __syncthreads();
      if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (ld_relaxed(bar) < target) {}
      }
      __syncthreads();
    } else if (V == 4) {               // cooperative groups
      grid.sync();
    } else if (V == 5) {               // no fences at all (lower bound: atomic + poll)
      __syncthreads();
      if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (ld_relaxed(bar) < target) {}
      }
      __syncthreads();
    }
    const int nb = (blockIdx.x + 1) % gridDim.x;
    acc += __ldcg(data + nb * 6 + (threadIdx.x % 6));
  }
  if (acc == -1.0) data[0] = acc;
}

template <int V>
float run(int grid, int iters, unsigned* bar, double* data) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  void* args[] = {&bar, &data, &iters};
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(bar, 0, 4);
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k_bar<V>, dim3(grid), dim3(512), args, 0, 0);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("variant %d failed: %s\n", V, cudaGetErrorString(cudaGetLastError())); return -1.f; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best * 1e3f / iters;
}

int main() {
  unsigned* bar; double* data;
  cudaMalloc(&bar, 64); cudaMalloc(&data, 148 * 6 * 8); cudaMemset(data, 0, 148 * 6 * 8);
  const char* names[] = {"fence+atomicAdd+volatile poll+fence", "fence+red+relaxed poll+fence", "red.release+ld.acquire poll", "red.release+relaxed poll (no trailing fence)",
                         "cooperative_groups grid.sync()", "red.relaxed+relaxed poll (no fences: lower bound)"};
  for (int grid : {113, 148, 32}) {
    printf("grid %d x 512 threads, us per barrier (incl. one 48-B store and one ld.cg per thread):\n", grid);
    printf("  [0] %-52s %.2f\n", names[0], run<0>(grid, 2000, bar, data));
    printf("  [1] %-52s %.2f\n", names[1], run<1>(grid, 2000, bar, data));
    printf("  [2] %-52s %.2f\n", names[2], run<2>(grid, 2000, bar, data));
    printf("  [3] %-52s %.2f\n", names[3], run<3>(grid, 2000, bar, data));
    printf("  [4] %-52s %.2f\n", names[4], run<4>(grid, 2000, bar, data));
    printf("  [5] %-52s %.2f\n", names[5], run<5>(grid, 2000, bar, data));
  }
  return 0;
}
