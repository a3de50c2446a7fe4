// Data-parallel optimizer step: sum-all-reduce of the flat gradient bucket + Adam in ONE kernel over NVLink / NVSwitch peer
// memory (one process per GPU, every rank launches the same kernel in the same step).
//
// Replaces   dist.all_reduce(flat_grads, SUM)  +  dg_adam_step   (two launches + NCCL's own protocol, ~100 us exposed per
// iteration at 8 GPUs for a 4.4 MB bucket, SCALE_r01.json) by a two-shot exchange that the kernel does itself:
//   barrier A   every rank's bucket is complete (kernel order on its stream) -> flag to every peer, wait for every peer's flag;
//   phase 1     rank r owns shard r of the bucket: it reads that shard from every rank's bucket (peer loads in rank order, or
//               one multimem.ld_reduce through the NVSwitch multicast mapping), and writes the sum into shard r of EVERY
//               rank's bucket (peer stores / multimem.st) - each rank moves 2 (W-1)/W of the bucket instead of W-1 buckets;
//   barrier B   all my stores are out (system fence, last block signals) -> flag to every peer; wait for every peer's flag:
//               my bucket now holds the global sum, and nobody reads or writes it any more;
//   phase 2     Adam over the whole flat buffer from the reduced bucket (same arithmetic as adam_kernel) - every rank applies
//               identical sums, so the replicas stay bit-identical without a parameter broadcast.
// The buckets and the flag blocks live in symmetric memory (torch.distributed._symmetric_memory on the host side,
// downgan_b200/dp.py); the kernel only sees raw peer-mapped pointers.  All waits are bounded (trap instead of hanging the box).
#include "dg_common.cuh"

namespace dg {
namespace {

constexpr int DP_MAX_WORLD = 8;
constexpr int DP_THREADS = 256;
constexpr int DP_FLAG_B = 8;         // flags[0..7]: barrier A by source rank, flags[8..15]: barrier B, flags[16]: local block counter
constexpr int DP_FLAG_COUNTER = 16;

struct DpArgs {
  float* grad_peer[DP_MAX_WORLD];      // every rank's gradient bucket, peer-mapped (index = rank; [rank] is the local one)
  uint32_t* flag_peer[DP_MAX_WORLD];   // every rank's flag block
  float* grad_mc;                      // multicast mapping of the bucket (NVSwitch reduces / broadcasts), or null
  float *param, *m, *v;
  long long n;
  int rank, world;
  uint32_t epoch;
  float b1, b2, eps, step_size, inv_sqrt_bc2, gscale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {  // bypasses L1: the peer rewrites this address every step
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_f32(float* p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ float4 mc_ld_reduce_v4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_v4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// bounded spin on one of MY flags until it reaches `epoch` (monotonic counters, wrap-safe compare)
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    if (clock64() - t0 > 120000000000LL) __trap();  // ~60 s: a peer died or the ranks are out of step
  }
}

__global__ void __launch_bounds__(DP_THREADS) dp_allreduce_adam_kernel(const DpArgs a) {
  __shared__ bool last;
  const int tid = threadIdx.x;
  uint32_t* my_flags = a.flag_peer[a.rank];
  // ---- barrier A: all buckets complete
  if (blockIdx.x == 0 && tid < a.world) st_release_sys(a.flag_peer[tid] + a.rank, a.epoch);
  if (tid < a.world) wait_flag(my_flags + tid, a.epoch);
  __syncthreads();
  // ---- phase 1: reduce my shard, publish it to every bucket
  const long long n4 = a.n >> 2;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank, hi = min(n4, lo + per);
  for (long long i = lo + (long long)blockIdx.x * DP_THREADS + tid; i < hi; i += (long long)gridDim.x * DP_THREADS) {
    if (a.grad_mc) {
      mc_st_v4(a.grad_mc + 4 * i, mc_ld_reduce_v4(a.grad_mc + 4 * i));
    } else {
      float4 s = ld_sys_v4(a.grad_peer[0] + 4 * i);
      for (int p = 1; p < a.world; ++p) {
        const float4 t = ld_sys_v4(a.grad_peer[p] + 4 * i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      for (int p = 0; p < a.world; ++p) st_sys_v4(a.grad_peer[p] + 4 * i, s);
    }
  }
  if (a.rank == a.world - 1 && blockIdx.x == 0) {  // scalar tail (n % 4 elements) belongs to the last rank
    for (long long i = (n4 << 2) + tid; i < a.n; i += DP_THREADS) {
      float s = 0.f;
      for (int p = 0; p < a.world; ++p) s += ld_sys_f32(a.grad_peer[p] + i);
      for (int p = 0; p < a.world; ++p) st_sys_f32(a.grad_peer[p] + i, s);
    }
  }
  // ---- barrier B: my stores are visible everywhere; every peer's stores into my bucket are visible here
  __threadfence_system();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(my_flags + DP_FLAG_COUNTER, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (tid == 0) my_flags[DP_FLAG_COUNTER] = 0;  // every block has arrived: ready for the next launch
    if (tid < a.world) st_release_sys(a.flag_peer[tid] + DP_FLAG_B + a.rank, a.epoch);
  }
  if (tid < a.world) wait_flag(my_flags + DP_FLAG_B + tid, a.epoch);
  __syncthreads();
  // ---- phase 2: Adam over the whole flat buffer from the reduced bucket (L1 holds no line of it: __ldcg)
  const float* g = a.grad_peer[a.rank];
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= a.gscale;
    mi = a.b1 * mi + (1.f - a.b1) * gi;
    vi = a.b2 * vi + (1.f - a.b2) * gi * gi;
    pi -= a.step_size * mi / (sqrtf(vi) * a.inv_sqrt_bc2 + a.eps);
  };
  float4* p4 = reinterpret_cast<float4*>(a.param);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  for (long long i = (long long)blockIdx.x * DP_THREADS + tid; i < n4; i += (long long)gridDim.x * DP_THREADS) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldcg(reinterpret_cast<const float4*>(g) + i);
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * DP_THREADS + tid; i < a.n; i += (long long)gridDim.x * DP_THREADS)
    upd(a.param[i], __ldcg(g + i), a.m[i], a.v[i]);
}

}  // namespace
}  // namespace dg

extern "C" int dg_dp_allreduce_adam(const dg_dp_peers* peers, float* params, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                    float beta1, float beta2, float eps, int step, float grad_scale, unsigned int epoch, void* stream) {
  using namespace dg;
  DG_CHECK(peers && params && exp_avg && exp_avg_sq && n > 0, "dg_dp_allreduce_adam: null argument");
  DG_CHECK(peers->world >= 1 && peers->world <= DP_MAX_WORLD && peers->rank >= 0 && peers->rank < peers->world,
           "dg_dp_allreduce_adam: rank %d / world %d (at most %d ranks on one node)", peers->rank, peers->world, DP_MAX_WORLD);
  DpArgs a;
  memset(&a, 0, sizeof(a));
  uintptr_t align = (uintptr_t)params | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)peers->grad_multicast;
  for (int p = 0; p < peers->world; ++p) {
    DG_CHECK(peers->grad_ptrs[p] && peers->flag_ptrs[p], "dg_dp_allreduce_adam: missing peer pointer %d", p);
    a.grad_peer[p] = (float*)peers->grad_ptrs[p];
    a.flag_peer[p] = (uint32_t*)peers->flag_ptrs[p];
    align |= (uintptr_t)peers->grad_ptrs[p];
  }
  DG_CHECK((align & 15) == 0, "dg_dp_allreduce_adam: buffers must be 16-byte aligned");
  a.grad_mc = (float*)peers->grad_multicast;
  a.param = params; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.rank = peers->rank; a.world = peers->world; a.epoch = epoch;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.step_size = (float)(lr / bc1); a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  a.gscale = grad_scale;
  cudaStream_t st = (cudaStream_t)stream;
  // every block waits inside the kernel (barriers A and B): the grid must be co-resident -> at most one block per SM
  int dev = 0, sms = 0;
  DG_CUDA(cudaGetDevice(&dev));
  DG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long want = ((n >> 2) + DP_THREADS - 1) / DP_THREADS;
  const int grid = (int)std::max(1LL, std::min<long long>(want, std::min(sms, 128)));
  Prof prof(PC_ADAM, 0.0, (double)n * 28.0, st);
  dp_allreduce_adam_kernel<<<grid, DP_THREADS, 0, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}
