// Probe for the position-contracted (weight-gradient) tcgen05 formulation on TMA-written swizzled NHWC tiles:
//   D[(g, ci)][co] = sum_k x[k + g + shift][ci] * dy[k][co]
// A = x as an MN-major operand whose M-atoms (CX channels each) are ONE POSITION apart (LBO = row bytes), i.e.
// overlapping atoms = the kx taps of a 3x3 kernel row folded into M; B = dy MN-major.  Also reports where the
// rows of an M = 64 accumulator live in TMEM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
constexpr int R = 128;  // rows (positions) staged per operand
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
// CX = channels of x per position (16/32/64), CD = channels of dy (16/32/64), M = 64 or 128
template <int CX, int CD>
__global__ void probe(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap md, float* out, int M, int shift, int ksteps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t sx = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sd = sx + R * CX * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(R * (CX + CD) * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sx), "l"(&mx), "r"(0), "r"(0), "r"(smem_u32(&bars[0])) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sd), "l"(&md), "r"(0), "r"(0), "r"(smem_u32(&bars[0])) : "memory");
    mbar_wait(smem_u32(&bars[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t lx = (CX == 16) ? 6ull : (CX == 32 ? 4ull : 2ull), ld = (CD == 16) ? 6ull : (CD == 32 ? 4ull : 2ull);
    // instruction descriptor: bf16 x bf16 -> f32, A and B both MN-major
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CD >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t ad = mk_desc(sx + (shift + 16 * ks) * CX * 2, CX * 2 /*LBO: next atom = next position*/, 8 * CX * 2, lx);
      const uint64_t bd = mk_desc(sd + (16 * ks) * CD * 2, 0, 8 * CD * 2, ld);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(ks > 0 ? 1 : 0) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  __syncthreads();
  mbar_wait(smem_u32(&bars[1]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int nc = 0; nc < CD; nc += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + nc;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[tid * CD + nc + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}
static CUtensorMap make2d(void* p, int C, int rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  const CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}
template <int CX, int CD> void run() {
  std::vector<bf16> hx((size_t)R * CX), hd((size_t)R * CD);
  for (int r = 0; r < R; ++r) for (int c = 0; c < CX; ++c) hx[(size_t)r * CX + c] = __float2bfloat16((float)((int)(((unsigned)(r * 131 + c * 71 + 7) * 2654435761u) >> 28) - 8));
  for (int r = 0; r < R; ++r) for (int c = 0; c < CD; ++c) hd[(size_t)r * CD + c] = __float2bfloat16((float)((int)(((unsigned)(r * 977 + c * 37 + 3) * 2246822519u) >> 28) - 8));
  bf16 *dx, *dd; cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dd, hd.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dd, hd.data(), hd.size() * 2, cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout, 128 * CD * 4);
  CUtensorMap mx = make2d(dx, CX, R), md = make2d(dd, CD, R);
  cudaFuncSetAttribute(probe<CX, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho((size_t)128 * CD);
  for (int M : {128, 64})
    for (int shift : {0, 19})
      for (int ksteps : {3}) {
        cudaMemset(dout, 0xFF, 128 * CD * 4);
        probe<CX, CD><<<1, 128, 48 * 1024>>>(mx, md, dout, M, shift, ksteps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CX %d CD %d M %d: %s\n", CX, CD, M, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        // expected rows m = g*CX + ci
        int found_at[128]; int nfound = 0;
        for (int m = 0; m < M; ++m) {
          const int g = m / CX, ci = m % CX;
          std::vector<float> want(CD);
          for (int co = 0; co < CD; ++co) {
            float s = 0;
            for (int k = 0; k < 16 * ksteps; ++k) s += __bfloat162float(hx[(size_t)(k + g + shift) * CX + ci]) * __bfloat162float(hd[(size_t)k * CD + co]);
            want[co] = s;
          }
          found_at[m] = -1;
          for (int lane = 0; lane < 128; ++lane) {
            bool eq = true;
            for (int co = 0; co < CD; ++co) if (ho[(size_t)lane * CD + co] != want[co]) { eq = false; break; }
            if (eq) { found_at[m] = lane; ++nfound; break; }
          }
        }
        printf("CX %2d CD %2d M %3d shift %2d ksteps %d: rows found %3d/%3d; row->lane:", CX, CD, M, shift, ksteps, nfound, M);
        for (int m : {0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 127}) if (m < M) printf(" %d->%d", m, found_at[m]);
        printf("\n");
      }
}
int main() {
  cudaFree(0);
  run<16, 16>();
  run<16, 32>();
  run<32, 64>();
  run<64, 64>();
  return 0;
}
