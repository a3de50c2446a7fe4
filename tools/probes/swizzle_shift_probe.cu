// Probe: does a tcgen05.mma A descriptor over a TMA-written SWIZZLE_{32,64,128}B K-major tile tolerate a
// start address shifted by an arbitrary number of rows (the 3x3-tap shift of the implicit-GEMM conv)?
// X is [R][C] bf16 (C = 16, 32 or 64 channels = one swizzle row), B = 16x16 identity, so
// D[r][n] must equal X[r + shift][kk*16 + n].  Tested with descriptor base_offset = 0 and = (addr>>7)&7.
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
constexpr int R = 192;  // rows staged
template <int ROWB>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int shift, int kk, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t sa = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sb = sa + R * ROWB;  // B operand after A (1024-aligned since R*ROWB is)
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B: identity 16x16 in canonical no-swizzle K-major: [K/8 = 2 planes][16 rows n][8 k]
  for (int i = tid; i < 2 * 16 * 8; i += blockDim.x) {
    const int pl = i / 128, n = (i / 8) % 16, k8 = i % 8;
    const int k = pl * 8 + k8;
    reinterpret_cast<bf16*>(smem_raw + (sb - smem_u32(smem_raw)))[i] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(R * ROWB) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(sa), "l"(&tmap), "r"(0), "r"(0), "r"(smem_u32(&bars[0])) : "memory");
    mbar_wait(smem_u32(&bars[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_addr = sa + shift * ROWB + kk * 32;
    const uint64_t layout = (ROWB == 32) ? 6ull : (ROWB == 64 ? 4ull : 2ull);
    uint64_t ad = 0;
    ad |= (uint64_t)((a_addr >> 4) & 0x3FFFu);
    ad |= (uint64_t)1 << 16;                          // LBO (unused for swizzled K-major)
    ad |= (uint64_t)(((8 * ROWB) >> 4) & 0x3FFFu) << 32;  // SBO = 8 rows
    ad |= (uint64_t)1 << 46;
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    ad |= layout << 61;
    uint64_t bd = 0;
    bd |= (uint64_t)((sb >> 4) & 0x3FFFu);
    bd |= (uint64_t)((256 >> 4) & 0x3FFFu) << 16;  // LBO = plane stride
    bd |= (uint64_t)((128 >> 4) & 0x3FFFu) << 32;  // SBO
    bd |= (uint64_t)1 << 46;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  }
  __syncthreads();
  mbar_wait(smem_u32(&bars[1]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
template <int ROWB> int run() {
  const int C = ROWB / 2;
  std::vector<bf16> hx((size_t)R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) hx[(size_t)r * C + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251) - 125.f);
  bf16* dx; cudaMalloc(&dx, hx.size() * 2); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  float* dout; cudaMalloc(&dout, 128 * 16 * 4);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)ROWB};
  cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)R};
  cuuint32_t es[2] = {1, 1};
  const CUtensorMapSwizzle sw = ROWB == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(probe<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(128 * 16);
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int kk = 0; kk < ROWB / 32; ++kk) {
      printf("rowbytes %d base_offset=%s kk %d: shifts ok =", ROWB, ubo ? "(addr>>7)&7" : "0", kk);
      for (int shift : {0, 1, 2, 3, 5, 8, 9, 18, 19, 20, 37, 63}) {
        cudaMemset(dout, 0, 128 * 16 * 4);
        probe<ROWB><<<1, 128, 48 * 1024>>>(m, dout, shift, kk, ubo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" [shift %d: %s]", shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int row = 0; row < 128; ++row) for (int n = 0; n < 16; ++n) {
          const float want = __bfloat162float(hx[(size_t)(row + shift) * C + kk * 16 + n]);
          if (ho[row * 16 + n] != want) ++bad;
        }
        printf(" %d:%s", shift, bad ? "NO" : "ok");
        if (bad && shift <= 1) printf("(%d bad)", bad);
      }
      printf("\n");
    }
  return 0;
}
int main() {
  cudaFree(0);
  if (run<32>()) return 1;
  if (run<64>()) return 1;
  if (run<128>()) return 1;
  return 0;
}
