// Micro-probe: per-launch cost of back-to-back kernels as a function of dynamic shared memory,
// carveout preference, TMEM allocation and alternation with a no-smem kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_plain(float* p) { if (threadIdx.x == 0 && blockIdx.x == 0 && p) p[0] += 1.f; }
__global__ void k_smem(float* p) {
  extern __shared__ float s[];
  s[threadIdx.x] = 1.f; __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0 && p) p[0] += s[1];
}
__global__ void k_tmem(float* p, int cols) {
  extern __shared__ float s[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  s[threadIdx.x] = 1.f;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(cols) : "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0 && p) p[0] += s[1];
}
template <typename F> float timeit(F f, int n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return 1e3f * ms / n;
}
int main() {
  float* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  const int N = 2000;
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148, 296}) {
    printf("grid %d\n", grid);
    printf("  plain                       %.2f us\n", timeit([&] { k_plain<<<grid, 256>>>(d); }, N));
    for (int kb : {16, 48, 100, 200}) {
      printf("  smem %3d KB                 %.2f us\n", kb, timeit([&] { k_smem<<<grid, 256, kb * 1024>>>(d); }, N));
      printf("  smem %3d KB alt plain       %.2f us per pair\n", kb, timeit([&] { k_smem<<<grid, 256, kb * 1024>>>(d); k_plain<<<grid, 256>>>(d); }, N));
    }
    printf("  tmem 64 cols, smem 100 KB   %.2f us\n", timeit([&] { k_tmem<<<grid, 256, 100 * 1024>>>(d, 64); }, N));
    printf("  tmem 512 cols, smem 100 KB  %.2f us\n", timeit([&] { k_tmem<<<grid, 256, 100 * 1024>>>(d, 512); }, N));
    printf("  tmem 64 alt plain           %.2f us per pair\n", timeit([&] { k_tmem<<<grid, 256, 100 * 1024>>>(d, 64); k_plain<<<grid, 256>>>(d); }, N));
  }
  cudaFuncSetAttribute(k_plain, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_smem, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  printf("with carveout=100 on both:\n");
  for (int kb : {16, 100, 200})
    printf("  smem %3d KB alt plain       %.2f us per pair\n", kb, timeit([&] { k_smem<<<148, 256, kb * 1024>>>(d); k_plain<<<148, 256>>>(d); }, N));
  // the same chains captured into a CUDA graph: GPU-side spacing of dependent kernels without the host's launch cost
  {
    cudaStream_t cs; cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    auto graph_time = [&](int variant, int len) {
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
      for (int i = 0; i < len; ++i) {
        if (variant == 0) k_plain<<<148, 256, 0, cs>>>(d);
        else if (variant == 1) { k_smem<<<148, 256, 100 * 1024, cs>>>(d); k_plain<<<148, 256, 0, cs>>>(d); }
        else k_tmem<<<148, 256, 100 * 1024, cs>>>(d, 64);
      }
      cudaStreamEndCapture(cs, &g);
      cudaGraphInstantiate(&ge, g, 0);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaGraphLaunch(ge, cs); cudaStreamSynchronize(cs);
      cudaEventRecord(a, cs);
      for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, cs);
      cudaEventRecord(b, cs); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      return 1e3f * ms / (5 * len);
    };
    printf("CUDA graph, chain of 400 dependent launches (grid 148):\n");
    printf("  plain                       %.2f us per launch\n", graph_time(0, 400));
    printf("  smem 100 KB alt plain       %.2f us per pair\n", graph_time(1, 200));
    printf("  tmem 64 cols, smem 100 KB   %.2f us per launch\n", graph_time(2, 400));
    // stream launches with the host kept far ahead of the GPU by a long first kernel are not needed: the graph numbers
    // bound what a stream can do
  }
  // events around every launch (what the library's profiler does)
  cudaEvent_t e[2]; cudaEventCreate(&e[0]); cudaEventCreate(&e[1]);
  float acc = 0;
  for (int i = 0; i < 200; ++i) {
    cudaEventRecord(e[0]); k_plain<<<148, 256>>>(d); cudaEventRecord(e[1]); cudaEventSynchronize(e[1]);
    float ms; cudaEventElapsedTime(&ms, e[0], e[1]); if (i >= 100) acc += ms;
  }
  printf("event-bracketed plain kernel: %.2f us\n", 1e3f * acc / 100);
  acc = 0;
  for (int i = 0; i < 200; ++i) {
    k_plain<<<148, 256>>>(d);
    cudaEventRecord(e[0]); k_smem<<<148, 256, 100 * 1024>>>(d); cudaEventRecord(e[1]); cudaEventSynchronize(e[1]);
    float ms; cudaEventElapsedTime(&ms, e[0], e[1]); if (i >= 100) acc += ms;
  }
  printf("event-bracketed smem-100KB kernel after a plain one: %.2f us\n", 1e3f * acc / 100);
  return 0;
}
