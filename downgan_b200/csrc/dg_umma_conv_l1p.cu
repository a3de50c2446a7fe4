// tcgen05 forward conv for the 1- / 2-channel fp32 boundary layers WITHOUT an im2col build (round 2), sm_100a.
// (critic features.0: 2 -> 16k on the fine fields, its JVP twin, and the data gradient of the generator's last conv.)
//
// dg_umma_conv_l1.cu builds the im2col row of every position (18 values -> 32 bf16 columns: 9 shared loads, 16 packs and
// four 16-byte stores per position), and that builder loop bounds the layer at ~2.5x its HBM time.  Here the kx taps are
// folded into the operand itself:
//   * the builders write ONE 16-byte chunk per input position p = (y, x):  R[p] = bf16{ in[y][x-1][:], in[y][x][:],
//     in[y][x+1][:], 0.. }  (3 taps x Ci <= 6 of the 8 elements; neighbours come from warp shuffles, image borders are
//     zero) - one coalesced global load, two shuffles and one shared store per position, rows at pitch W (no pad columns);
//   * a K = 8 chunk of the A operand for kernel row ky is then the chunk of the position one image row further down, so
//     with the K-major no-swizzle layout (rows 16 B apart, the two K halves LBO apart) an M-tile of 128 consecutive output
//     positions needs two MMAs:  K halves (ky = 0, ky = 1) with LBO = one image row, and (ky = 2, zero weights) - the
//     operand is the staged tile itself at a start-address offset, as in the trunk kernel;
//   * the weight image is [4 chunks: ky = 0, 1, 2, zero][16 output channels][8] bf16, built once per CTA.
// Roles: 2 x 4 builder warps (the groups alternate tiles), 1 MMA warp (one elected lane), 4 epilogue warps (one per TMEM lane quarter);
// 3-slot tile ring, 2 accumulator stages, mbarrier pipelines.  A tile = TR whole image rows = n_mt * 128 positions.
//
// MEASURED (round 2, 192 x (2 -> 16) at 128x128): 42.8 us alone against 50.5 us for the im2col kernel, but the cfg-2 step is
// 1.4 % SLOWER with it (43.0 k vs 43.6 k samples/s, two A/B pairs): its twelve busy warps per CTA take issue slots from the
// kernels that run beside it on the side streams.  It is therefore OFF by default (dg_set_tuning(20, 1) selects it) and kept
// as the starting point for a leaner builder (profiles/README.md).
#include <algorithm>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int P1_GROUP_WARPS = 4, P1_GROUPS = 2, P1_BUILD_WARPS = P1_GROUP_WARPS * P1_GROUPS;  // two builder groups alternate tiles
constexpr int P1_BUILDERS = P1_GROUP_WARPS * 32;                                               // arrivals per tile
constexpr int P1_MMA_WARP = P1_BUILD_WARPS, P1_EPI_WARP0 = P1_MMA_WARP + 1, P1_EPI_WARPS = 4;
constexpr int P1_THREADS = (P1_EPI_WARP0 + P1_EPI_WARPS) * 32;  // 416
constexpr int P1_NSTAGE = 3;
constexpr int P1_MAX_MT = 4;  // M-tiles per tile: 64 TMEM columns per accumulator stage
constexpr int P1_NP = 11;     // staged positions per builder lane: (TR + 3) * W / 128 <= 10.x for 32 <= W <= 256

struct P1Args {
  ConvOp op;
  int tiles_total, tiles_per_img, TR, n_mt, stage_bytes;
};

__device__ __forceinline__ void mbar_arrive_p1(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// BACKOFF: the MMA and epilogue warps of this kernel mostly wait for the builders; polling in a tight loop they took half of
// the SM's issue slots from the eight builder warps (ncu: 21 M of 29 M warp instructions were the wait loop), so they sleep
// between polls
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait_p1(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
    if (BACKOFF && spin > 0) __nanosleep(128);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ uint32_t elect_one_sync_p1() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ uint32_t pk2p(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void sts16p(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int CI>
__global__ void __launch_bounds__(P1_THREADS, 2) conv_l1p_kernel(const P1Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * P1_NSTAGE + 4];  // full[3], empty[3], tfull[2], tempty[2]
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[16];
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * 16;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (P1_NSTAGE + s); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (2 * P1_NSTAGE + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (2 * P1_NSTAGE + 2 + q); };
  const int my_tiles = ((int)blockIdx.x < a.tiles_total) ? (a.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == P1_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), 128);
  if (tid == 0) {
    for (int s = 0; s < P1_NSTAGE; ++s) { mbar_init(full_bar(s), P1_BUILDERS); mbar_init(empty_bar(s), 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), P1_EPI_WARPS); }
  }
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sW = s0 + P1_NSTAGE * a.stage_bytes;  // weight image [4 chunks][16 rows][8] bf16
  {
    const int CoP = (op.Co + 15) & ~15;
    bf16* wimg = reinterpret_cast<bf16*>(smem + (size_t)P1_NSTAGE * a.stage_bytes);
    for (int i = tid; i < 4 * 16 * 8; i += P1_THREADS) {
      const int ky = i >> 7, n = (i >> 3) & 15, j = i & 7;  // element j of chunk ky = (kx = j / CI, ci = j % CI)
      float w = 0.f;
      if (ky < 3 && j < 3 * CI) w = op.w[(size_t)((ky * 3 + j / CI) * CI + j % CI) * CoP + co0 + n];
      wimg[i] = __float2bfloat16_rn(w);
    }
    if (tid < 16) sbias[tid] = op.bias ? op.bias[co0 + tid] : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int W = op.Win, H = op.Hin, TR = a.TR;
  const int Wl = 31 - __clz(W);
  const int tile_pos = TR * W;           // output positions per tile (n_mt * 128)
  const int staged = (TR + 3) * W;       // staged input positions: rows y0 - 1 .. y0 + TR + 1

  if (warp < P1_BUILD_WARPS) {
    // ================= builders =================
    const float* xg = (const float*)op.x.p;
    const int grp = warp / P1_GROUP_WARPS, wg = warp % P1_GROUP_WARPS;
    for (int it = grp; it < my_tiles; it += P1_GROUPS) {
      const int s = it % P1_NSTAGE;
      const int t = blockIdx.x + it * gridDim.x;
      const int n = t / a.tiles_per_img, y0 = (t - n * a.tiles_per_img) * TR;
      mbar_wait_p1(empty_bar(s), (((uint32_t)(it / P1_NSTAGE)) & 1u) ^ 1u);
      const uint32_t sR = s0 + s * a.stage_bytes;
      // A warp covers 32 consecutive x of one staged row (W is a multiple of 32).  Phase 1 issues every global load of the
      // tile (own pixel; lanes 0 / 31 also the neighbour outside the warp's span), phase 2 exchanges neighbours by shuffle
      // and writes the chunks - one global-load latency per tile instead of one per row.
      float c0[P1_NP], c1[P1_NP], e0[P1_NP], e1[P1_NP];
#pragma unroll
      for (int k = 0; k < P1_NP; ++k) {
        c0[k] = c1[k] = e0[k] = e1[k] = 0.f;
        const int i = (k * P1_GROUP_WARPS + wg) * 32 + lane;
        if (i < staged) {
          const int rr = i >> Wl, x = i & (W - 1);
          const int gy = y0 - 1 + rr;
          if (gy >= 0 && gy < H) {
            const float* src = xg + ((size_t)(n * H + gy) * W + x) * CI;
            if (CI == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(src)); c0[k] = v.x; c1[k] = v.y; }
            else c0[k] = __ldg(src);
            const int xe = (lane == 0) ? x - 1 : x + 1;  // the neighbour outside the warp's 32 positions
            if ((lane == 0 || lane == 31) && xe >= 0 && xe < W) {
              const float* se = xg + ((size_t)(n * H + gy) * W + xe) * CI;
              e0[k] = __ldg(se);
              if (CI == 2) e1[k] = __ldg(se + 1);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < P1_NP; ++k) {
        const int i = (k * P1_GROUP_WARPS + wg) * 32 + lane;
        if ((k * P1_GROUP_WARPS + wg) * 32 < staged) {  // warp-uniform
          float l0 = __shfl_up_sync(0xFFFFFFFFu, c0[k], 1), r0 = __shfl_down_sync(0xFFFFFFFFu, c0[k], 1), l1 = 0.f, r1 = 0.f;
          if (CI == 2) { l1 = __shfl_up_sync(0xFFFFFFFFu, c1[k], 1); r1 = __shfl_down_sync(0xFFFFFFFFu, c1[k], 1); }
          if (lane == 0) { l0 = e0[k]; l1 = e1[k]; }
          if (lane == 31) { r0 = e0[k]; r1 = e1[k]; }
          uint4 v;
          if (CI == 2) v = make_uint4(pk2p(l0, l1), pk2p(c0[k], c1[k]), pk2p(r0, r1), 0u);
          else v = make_uint4(pk2p(l0, c0[k]), pk2p(r0, 0.f), 0u, 0u);
          sts16p(sR + (uint32_t)i * 16u, v);
        }
      }
      fence_proxy_async();
      mbar_arrive_p1(full_bar(s));
    }
  } else if (warp == P1_MMA_WARP) {
    // ================= MMA issue =================
    const uint32_t idesc = instr_desc(128, 16);
    const uint32_t rowB = (uint32_t)W * 16u;                       // one image row of chunks
    const uint64_t bd01 = smem_desc(sW, 256, 128), bd23 = smem_desc(sW + 512, 256, 128);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % P1_NSTAGE, q = it & 1;
      mbar_wait_p1<true>(tempty_bar(q), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      mbar_wait_p1<true>(full_bar(s), ((uint32_t)(it / P1_NSTAGE)) & 1u);
      tc_fence_after();
      if (elect_one_sync_p1()) {
        const uint32_t sR = s0 + s * a.stage_bytes;
        for (int mt = 0; mt < a.n_mt; ++mt) {
          // output position p of the tile reads staged positions p (ky = 0), p + W (ky = 1), p + 2W (ky = 2), p + 3W (x 0)
          const uint64_t ad01 = smem_desc(sR + (uint32_t)(mt * 128) * 16u, rowB, 128);
          const uint64_t ad23 = smem_desc(sR + (uint32_t)(mt * 128) * 16u + 2u * rowB, rowB, 128);
          umma_f16(tmem + q * 64 + mt * 16, ad01, bd01, idesc, 0u);
          umma_f16(tmem + q * 64 + mt * 16, ad23, bd23, idesc, 1u);
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(q));
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: warp w reads TMEM lane quarter w % 4 =================
    const int lq = warp & 3;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const int q = it & 1;
      const long long p0 = (long long)(blockIdx.x + it * gridDim.x) * tile_pos;
      mbar_wait_p1<true>(tfull_bar(q), ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      for (int mt = 0; mt < a.n_mt; ++mt) {
        const size_t p = (size_t)(p0 + mt * 128 + lq * 32 + lane);
        uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
        if (op.act == ACT_MASK) {
          const uint4* mp = reinterpret_cast<const uint4*>((const bf16*)op.mask.p + p * op.mask.pitch + op.mask.coff + co0);
          m0 = mp[0]; m1 = mp[1];
        }
        float v[16];
        tmem_ld16(tmem + lane_base + q * 64 + mt * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += sbias[j];
        if (op.act == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
        } else if (op.act == ACT_MASK) {
          const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[2 * k] *= (__uint_as_float(w[k] << 16) > 0.f ? 1.f : op.slope);
            v[2 * k + 1] *= (__uint_as_float(w[k] & 0xFFFF0000u) > 0.f ? 1.f : op.slope);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)op.y.p + p * op.y.pitch + op.y.coff + co0);
        dst[0] = make_uint4(pk2p(v[0], v[1]), pk2p(v[2], v[3]), pk2p(v[4], v[5]), pk2p(v[6], v[7]));
        dst[1] = make_uint4(pk2p(v[8], v[9]), pk2p(v[10], v[11]), pk2p(v[12], v[13]), pk2p(v[14], v[15]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_p1(tempty_bar(q));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P1_MMA_WARP) tmem_dealloc(tmem, 128);
}

bool plan_p1(const ConvOp& op, P1Args& a) {
  if (op.transposed || op.stride != 1 || op.shuffle != SHUF_NONE || op.r1.p || op.r2.p || op.s_acc != 1.f) return false;
  if (op.Hin != op.Hout || op.Win != op.Wout) return false;
  if (op.Ci < 1 || op.Ci > 2 || op.Co % 16 || !op.w) return false;
  if (op.x.bf || op.x.pitch != op.Ci || op.x.coff != 0) return false;
  if (((uintptr_t)op.x.p & 7) != 0) return false;
  if (!op.y.bf || op.y.pitch % 8 || op.y.coff % 8) return false;
  if (op.act == ACT_MASK && (!op.mask.bf || op.mask.pitch % 8 || op.mask.coff % 8)) return false;
  const int W = op.Win;
  if ((W & (W - 1)) || W < 32 || W > 256) return false;
  if ((long long)op.B * op.Hin * op.Win >= (1LL << 31)) return false;
  int TR = 0;
  for (int mt = P1_MAX_MT; mt >= 1; --mt) {  // the largest tile of whole rows that divides the image
    if ((mt * 128) % W) continue;
    const int tr = mt * 128 / W;
    if (tr >= 1 && op.Hin % tr == 0) { TR = tr; break; }
  }
  if (TR == 0) return false;
  a.op = op;
  a.TR = TR;
  a.n_mt = TR * W / 128;
  a.tiles_per_img = op.Hin / TR;
  a.tiles_total = op.B * a.tiles_per_img;
  a.stage_bytes = (TR + 3) * W * 16;
  return true;
}

}  // namespace

bool conv_l1p_supported(const ConvOp& op) {
  P1Args a;
  return plan_p1(op, a);
}

int conv_l1p(const ConvOp& op, cudaStream_t st) {
  if (ablate(5)) return 0;
  P1Args a;
  if (!plan_p1(op, a)) { set_error("conv_l1p: unsupported shape"); return DG_ERR_INVALID; }
  const size_t smem = (size_t)P1_NSTAGE * a.stage_bytes + 1024;
  const long long px = (long long)op.B * op.Hout * op.Wout;
  Prof prof(PC_CONV_UMMA, 2.0 * px * op.Co * op.Ci * 9.0, (double)px * op.Co * 2.0 + (double)px * op.Ci * 4.0, st);
  const int n_chunks = op.Co / 16;
  int gx = std::max(1, std::min(a.tiles_total, (148 * 2) / n_chunks));
  const int per = (a.tiles_total + gx - 1) / gx;
  gx = (a.tiles_total + per - 1) / per;
#define P1_LAUNCH(CI)                                                                                                   \
  do {                                                                                                                  \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      DG_CUDA(cudaFuncSetAttribute(conv_l1p_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));      \
      DG_CUDA(cudaFuncSetAttribute(conv_l1p_kernel<CI>, cudaFuncAttributePreferredSharedMemoryCarveout,                 \
                                   cudaSharedmemCarveoutMaxShared));                                                    \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    conv_l1p_kernel<CI><<<dim3(gx, n_chunks), P1_THREADS, smem, st>>>(a);                                               \
  } while (0)
  if (op.Ci == 1) P1_LAUNCH(1);
  else P1_LAUNCH(2);
#undef P1_LAUNCH
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
