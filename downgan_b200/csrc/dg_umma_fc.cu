// tcgen05 kernels for the critic classifier's hidden layer (classifier.0: 8F*(fine/16)^2 -> 100, critic.py:94-96), sm_100a.
//
// STATUS: validated on a B200 in round 2 (tests/test_gpu_fc_umma.py: scalars and every gradient against the CUDA-core
// kernels at B = 5 / 16 / 64; profiles/README.md has the A/B) and ON by default (dg_set_tuning(14, 0) selects the CUDA-core
// kernels of dg_kernels.cu, which remain for fp32 mode and for shapes outside fc_umma_supported).
//
// profiles/per_layer_roofline_r01g.md: the five classifier launches of a critic iteration take 138 us more than their
// roofline time on the CUDA cores (~10 TFLOP/s).  All three products are small GEMMs with one long dimension (K = 8192):
//   forward        y[b][j]   = sum_k x[b][k] w[j][k]          M = samples, N = units,  contraction over k   (split-K over CTAs)
//   input gradient dx[b][k]  = sum_j dz[b][j] w[j][k] * m     M = samples, N = k-tile, contraction over j   (one k-tile per CTA)
//   weight grad    dW[j][k] += sum_b dz[b][j] x[b][k]         M = units,   N = k-tile, contraction over b   (one k-tile per CTA)
// Each CTA stages ONE operand pair in shared memory in the no-swizzle core-matrix layouts documented in dg_umma.cuh
// (K-major planar [k/8][row][8] as in dg_umma_conv.cu, MN-major planar [mn/8][k][8] as in dg_umma_wgrad_im2col.cu),
// converting the fp32 operands (weights, dz) to bf16 on the way, issues 7..12 tcgen05.mma per M-tile from one thread and
// drains the TMEM accumulator with eight warps (two per lane quarter, alternating 16-column pieces).
#include <stdlib.h>

#include <algorithm>

#include <cuda.h>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int FCU_THREADS = 256;
constexpr int FCU_KT = 128;   // k columns per CTA (forward: split-K chunk; gradients: output tile)
constexpr int FCU_NP = 112;   // hidden units padded to a multiple of 16 (MMA N of the forward, contraction of the input gradient)

__device__ __forceinline__ uint32_t fcu_pk2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void fcu_sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// eight consecutive fp32 values p[0..7] (elements >= n_valid read as zero) rounded to one 16-byte bf16 chunk
__device__ __forceinline__ uint4 fcu_load8(const float* p, int n_valid, bool vec_ok) {
  float v[8];
  if (n_valid >= 8 && vec_ok) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (i < n_valid) ? p[i] : 0.f;
  }
  return make_uint4(fcu_pk2(v[0], v[1]), fcu_pk2(v[2], v[3]), fcu_pk2(v[4], v[5]), fcu_pk2(v[6], v[7]));
}

struct FcuArgs {
  const void* x;    // bf16 [NB][K]   (forward / weight gradient: activations; input gradient: the mask = same tensor)
  const float* w;   // fp32 [N][K]
  const float* dz;  // fp32 [NB][N]
  float* y;         // forward: fp32 [NB][N], accumulated with atomics (zeroed by the host)
  void* dx;         // input gradient: bf16 [NB][K]
  float* dw;        // weight gradient: fp32 [N][K], accumulated with atomics
  int NB, K, N;
  int rows;         // staged sample rows: NB rounded up to 128 (forward, input gradient) or to 16 (weight gradient)
  float slope;
};

// ------------------------------------------------------------------------------------------------------------------
// forward: CTA c contracts k in [c*128, c*128+128).  A = x (K-major planar, 16 planes x rows x 16 B), B = w (K-major planar,
// 16 planes x 112 x 16 B, rounded to bf16 here).  D[sample][unit], one 112-column accumulator per 128-sample M-tile.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCU_THREADS) fc_fwd_umma_kernel(const FcuArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * FCU_KT;
  const int mtiles = a.rows >> 7;
  const uint32_t PBA = (uint32_t)a.rows * 16u, PBB = (uint32_t)FCU_NP * 16u;
  const uint32_t sA = smem_u32(smem), sB = sA + 16u * PBA;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  if (tid == 32) mbar_init(smem_u32(&mbar), 1);
  const bf16* xb = (const bf16*)a.x;
  for (int i = tid; i < a.rows * 16; i += FCU_THREADS) {  // plane fastest: 256 contiguous global bytes per sample row
    const int r = i >> 4, p = i & 15;
    const bool ok = r < a.NB;
    cp_async16(sA + p * PBA + r * 16, ok ? xb + (size_t)r * a.K + k0 + p * 8 : xb, ok ? 16 : 0);
  }
  for (int i = tid; i < FCU_NP * 16; i += FCU_THREADS) {
    const int j = i >> 4, p = i & 15;
    const uint4 v = fcu_load8(a.w + (size_t)(j < a.N ? j : 0) * a.K + k0 + p * 8, j < a.N ? 8 : 0, true);
    fcu_sts16(sB + p * PBB + j * 16, v);
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc(128, FCU_NP);
    const uint64_t kA = (uint64_t)((2 * PBA) >> 4), kB = (uint64_t)((2 * PBB) >> 4);
    for (int mt = 0; mt < mtiles; ++mt) {
      uint64_t ad = smem_desc(sA + mt * 128 * 16, PBA, 128), bd = smem_desc(sB, PBB, 128);
      for (int ks = 0; ks < FCU_KT / 16; ++ks, ad += kA, bd += kB) umma_f16(tmem + mt * FCU_NP, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after();
  const int q = warp & 3, half = warp >> 2;
  for (int mt = 0; mt < mtiles; ++mt) {
    const int b = mt * 128 + q * 32 + lane;
    int piece = 0;
    for (int nc = 0; nc < FCU_NP; nc += 16, ++piece) {
      if ((piece & 1) != half) continue;  // warp-uniform
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + mt * FCU_NP + nc, v);
      if (b < a.NB) {
        float* yr = a.y + (size_t)b * a.N + nc;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          if (nc + c + 3 < a.N && (a.N & 3) == 0) red_add_v4(yr + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
          else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (nc + c + e < a.N) atomicAdd(yr + c + e, v[c + e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------------------
// input gradient: CTA c produces dx[:, c*128 .. c*128+128).  A = dz (K-major planar over j: 14 planes x rows x 16 B, rounded
// to bf16 here), B = w (MN-major planar: [k/8 = 16 planes][j = 112][8 k], rounded here).  Epilogue: x lrelu'(mask), bf16 store.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCU_THREADS) fc_dgrad_umma_kernel(const FcuArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * FCU_KT;
  const int mtiles = a.rows >> 7;
  constexpr int JP = FCU_NP / 8;  // 14 planes of 8 hidden units
  const uint32_t PBA = (uint32_t)a.rows * 16u, PBB = (uint32_t)FCU_NP * 16u;
  const uint32_t sA = smem_u32(smem), sB = sA + (uint32_t)JP * PBA;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  if (tid == 32) mbar_init(smem_u32(&mbar), 1);
  const bool dz_vec = (a.N & 3) == 0;
  for (int i = tid; i < a.rows * JP; i += FCU_THREADS) {
    const int r = i / JP, p = i - r * JP;
    const int nv = (r < a.NB) ? max(0, min(8, a.N - p * 8)) : 0;
    const uint4 v = fcu_load8(a.dz + (size_t)(r < a.NB ? r : 0) * a.N + p * 8, nv, dz_vec);
    fcu_sts16(sA + p * PBA + r * 16, v);
  }
  for (int i = tid; i < FCU_NP * 16; i += FCU_THREADS) {  // k-group fastest: 512 contiguous global bytes per hidden unit
    const int j = i >> 4, g = i & 15;
    const uint4 v = fcu_load8(a.w + (size_t)(j < a.N ? j : 0) * a.K + k0 + g * 8, j < a.N ? 8 : 0, true);
    fcu_sts16(sB + g * PBB + j * 16, v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc(128, FCU_KT, 0, 1);
    const uint64_t kA = (uint64_t)((2 * PBA) >> 4);
    for (int mt = 0; mt < mtiles; ++mt) {
      uint64_t ad = smem_desc(sA + mt * 128 * 16, PBA, 128), bd = smem_desc(sB, 128, PBB);
      for (int ks = 0; ks < FCU_NP / 16; ++ks, ad += kA, bd += 16) umma_f16(tmem + mt * FCU_KT, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after();
  const int q = warp & 3, half = warp >> 2;
  const bf16* mb = (const bf16*)a.x;
  bf16* dxb = (bf16*)a.dx;
  for (int mt = 0; mt < mtiles; ++mt) {
    const int b = mt * 128 + q * 32 + lane;
    int piece = 0;
    for (int nc = 0; nc < FCU_KT; nc += 16, ++piece) {
      if ((piece & 1) != half) continue;  // warp-uniform
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + mt * FCU_KT + nc, v);
      if (b < a.NB) {
        const size_t o = (size_t)b * a.K + k0 + nc;
        const uint4 m0 = *reinterpret_cast<const uint4*>(mb + o), m1 = *reinterpret_cast<const uint4*>(mb + o + 8);
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        uint32_t ow[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float ma = __uint_as_float(mw[e] << 16), mbv = __uint_as_float(mw[e] & 0xFFFF0000u);
          ow[e] = fcu_pk2(v[2 * e] * (ma > 0.f ? 1.f : a.slope), v[2 * e + 1] * (mbv > 0.f ? 1.f : a.slope));
        }
        *reinterpret_cast<uint4*>(dxb + o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(dxb + o + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient: CTA c accumulates dW[:, c*128 .. c*128+128).  A = dz^T (MN-major planar: [j/8 = 16 planes][b = rows][8 j],
// rounded to bf16 here; planes past the 100 real units are zero), B = x (MN-major planar: [k/8 = 16 planes][b][8 k], cp.async).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCU_THREADS) fc_wgrad_umma_kernel(const FcuArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * FCU_KT;
  const uint32_t PB = (uint32_t)a.rows * 16u;
  const uint32_t sA = smem_u32(smem), sB = sA + 16u * PB;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 128);
  if (tid == 32) mbar_init(smem_u32(&mbar), 1);
  const bf16* xb = (const bf16*)a.x;
  for (int i = tid; i < a.rows * 16; i += FCU_THREADS) {  // k-group fastest: 256 contiguous global bytes per sample row
    const int b = i >> 4, g = i & 15;
    const bool ok = b < a.NB;
    cp_async16(sB + g * PB + b * 16, ok ? xb + (size_t)b * a.K + k0 + g * 8 : xb, ok ? 16 : 0);
  }
  const bool dz_vec = (a.N & 3) == 0;
  for (int i = tid; i < a.rows * 16; i += FCU_THREADS) {
    const int b = i >> 4, g = i & 15;
    const int nv = (b < a.NB) ? max(0, min(8, a.N - g * 8)) : 0;
    const uint4 v = fcu_load8(a.dz + (size_t)(b < a.NB ? b : 0) * a.N + (nv > 0 ? g * 8 : 0), nv, dz_vec);
    fcu_sts16(sA + g * PB + b * 16, v);
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc(128, FCU_KT, 1, 1);
    // Descriptors are rebuilt from 32-bit parts per K-step and the loop is not unrolled: with 64-bit descriptor increments
    // ptxas 12.9 unrolled this loop by four and dropped the instruction that writes the HIGH word of the B descriptor pair
    // (both descriptors share it), so the MMA took a stale kernel-parameter word as its stride field - an illegal shared
    // memory access whose appearance depended on the address of `x` (tools/sass_desc_check.py finds the pattern; it runs in
    // tests/test_abi.py on every build).
    const uint32_t hi = ((PB >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t lo_a = ((sA >> 4) & 0x3FFFu) | (8u << 16), lo_b = ((sB >> 4) & 0x3FFFu) | (8u << 16);
#pragma unroll 1
    for (int ks = 0; ks < a.rows / 16; ++ks) {
      const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(lo_a + 16u * ks);
      const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(lo_b + 16u * ks);
      umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after();
  const int q = warp & 3, half = warp >> 2;
  const int j = q * 32 + lane;  // accumulator row = hidden unit
  int piece = 0;
  for (int nc = 0; nc < FCU_KT; nc += 16, ++piece) {
    if ((piece & 1) != half) continue;  // warp-uniform
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + nc, v);
    if (j < a.N) {
      float* d = a.dw + (size_t)j * a.K + k0 + nc;
#pragma unroll
      for (int c = 0; c < 16; c += 4) red_add_v4(d + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

inline int fcu_round(int v, int m) { return (v + m - 1) / m * m; }

template <typename Kern>
int fcu_attr(Kern k, size_t smem) {
  DG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

}  // namespace

// shapes the tcgen05 classifier kernels take: bf16 activations, K a multiple of 128, up to 112 hidden units, up to 256 rows
bool fc_umma_supported(int NB, int K, int N, int x_bf) {
  return g_tune[14] && x_bf && K % FCU_KT == 0 && K >= FCU_KT && N >= 8 && N <= FCU_NP && NB >= 1 && NB <= 256;
}

// y += x w^T (y zeroed by the caller)
int fc_fwd_umma(const void* x, const float* w, float* y, int NB, int K, int N, cudaStream_t st) {
  if (ablate(6)) return 0;
  FcuArgs a{};
  a.x = x; a.w = w; a.y = y; a.NB = NB; a.K = K; a.N = N; a.rows = fcu_round(NB, 128);
  const size_t smem = (size_t)16 * a.rows * 16 + (size_t)16 * FCU_NP * 16 + 256;
  static bool attr = false;
  if (!attr) { DG_TRY(fcu_attr(fc_fwd_umma_kernel, 16 * 256 * 16 + 16 * FCU_NP * 16 + 256)); attr = true; }
  fc_fwd_umma_kernel<<<K / FCU_KT, FCU_THREADS, smem, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

// dx = (dz w) * lrelu'(mask), bf16 dx and mask
int fc_dgrad_umma(const float* dz, const float* w, void* dx, int NB, int K, int N, const void* mask, float slope, cudaStream_t st) {
  if (ablate(6)) return 0;
  FcuArgs a{};
  a.x = mask; a.w = w; a.dz = dz; a.dx = dx; a.NB = NB; a.K = K; a.N = N; a.rows = fcu_round(NB, 128); a.slope = slope;
  const size_t smem = (size_t)(FCU_NP / 8) * a.rows * 16 + (size_t)16 * FCU_NP * 16 + 256;
  static bool attr = false;
  if (!attr) { DG_TRY(fcu_attr(fc_dgrad_umma_kernel, (FCU_NP / 8) * 256 * 16 + 16 * FCU_NP * 16 + 256)); attr = true; }
  fc_dgrad_umma_kernel<<<K / FCU_KT, FCU_THREADS, smem, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

// DG_FC_CHECK=1 (diagnostic): print every operand's extent against the allocation that holds it
static void fcu_check(const char* what, const void* p, size_t bytes) {
  typedef CUresult (*Fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static Fn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &q, cudaEnableDefault, &r) != cudaSuccess || !q) return;
    fn = (Fn)q;
  }
  CUdeviceptr base = 0; size_t size = 0;
  const CUresult r = fn(&base, &size, (CUdeviceptr)p);
  const long long off = (long long)((CUdeviceptr)p - base);
  fprintf(stderr, "[fc check] %s: ptr %p needs %zu bytes; allocation base %p size %zu (offset %lld, room %lld)%s rc=%d\n", what, p, bytes,
          (void*)base, size, off, (long long)size - off, ((long long)size - off < (long long)bytes) ? "  <-- OUT OF RANGE" : "", (int)r);
}

// dw += dz^T x
int fc_wgrad_umma(const float* dz, const void* x, float* dw, int NB, int K, int N, cudaStream_t st) {
  if (ablate(6)) return 0;
  static const bool chk = getenv("DG_FC_CHECK") != nullptr;
  if (chk) {
    fprintf(stderr, "[fc check] fc_wgrad_umma NB %d K %d N %d\n", NB, K, N);
    fcu_check("dz", dz, (size_t)NB * N * 4);
    fcu_check("x", x, (size_t)NB * K * 2);
    fcu_check("dw", dw, (size_t)N * K * 4);
  }
  FcuArgs a{};
  a.x = x; a.dz = dz; a.dw = dw; a.NB = NB; a.K = K; a.N = N; a.rows = fcu_round(NB, 16);
  const size_t smem = (size_t)2 * 16 * a.rows * 16 + 256;
  static bool attr = false;
  if (!attr) { DG_TRY(fcu_attr(fc_wgrad_umma_kernel, 2 * 16 * 256 * 16 + 256)); attr = true; }
  fc_wgrad_umma_kernel<<<K / FCU_KT, FCU_THREADS, smem, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
