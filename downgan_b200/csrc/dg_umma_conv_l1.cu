// tcgen05 forward conv for the few-channel fp32 boundary layer (critic features.0: Ci = 2 -> Co = 16k, also its
// JVP and the data-gradient of the generator's last conv), sm_100a.
//
// K = 9*Ci <= 27 cannot come from a tap-shifted tile, so the im2col rows are BUILT: warp-specialised builders stage
// the TR + 2 fp32 input rows of a 256-position tile in shared memory (cp.async, zero halo) and write
// R[pos][tap*Ci + ci] (bf16, 32 columns, K-major planar [4 planes][256][8]); one elected thread issues
// 2 M-tiles x 2 K-steps of tcgen05.mma (M = 128, N = 16) into a double-buffered TMEM accumulator; four epilogue warps
// apply bias + LeakyReLU (or the mask) and store 16 bf16 channels (32 bytes) per position.  The three roles
// overlap through mbarrier pipelines (3-slot tile ring, 2 accumulator stages).
#include <algorithm>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int C1_BUILDERS = 256, C1_THREADS = 416;  // 8 builder warps, 1 MMA warp, 4 epilogue warps
constexpr int C1_MMA_WARP = 8, C1_EPI_WARP0 = 9;
constexpr int C1_TPOS = 256, C1_PB = C1_TPOS * 16, C1_NSTAGE = 3;

struct C1Args {
  ConvOp op;
  int tiles_total, xs_bytes, stage_bytes;
};

__device__ __forceinline__ void mbar_arrive_c1(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_c1(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
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
__device__ __forceinline__ uint32_t elect_one_sync_c1() {
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
__device__ __forceinline__ uint32_t pk2c(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void sts16c(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int CI>
__global__ void __launch_bounds__(C1_THREADS, 2) conv_l1_kernel(const C1Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * C1_NSTAGE + 4];  // full[3], empty[3], tfull[2], tempty[2]
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[16];
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * 16;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C1_NSTAGE + s); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (2 * C1_NSTAGE + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (2 * C1_NSTAGE + 2 + q); };
  const int my_tiles = ((int)blockIdx.x < a.tiles_total) ? (a.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == C1_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), 64);
  if (tid == 0) {
    for (int s = 0; s < C1_NSTAGE; ++s) { mbar_init(full_bar(s), C1_BUILDERS); mbar_init(empty_bar(s), 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), 4); }
  }
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sW = s0 + C1_NSTAGE * a.stage_bytes;  // weight image [4 planes][16 rows][8] bf16
  // zero the ring (halo columns of the staged rows, im2col columns >= 9*CI) and build the weight image
  for (uint32_t i = tid * 16; i < (uint32_t)(C1_NSTAGE * a.stage_bytes); i += C1_THREADS * 16) sts16c(s0 + i, make_uint4(0, 0, 0, 0));
  {
    const int CoP = (op.Co + 15) & ~15;
    bf16* wimg = reinterpret_cast<bf16*>(smem + (size_t)C1_NSTAGE * a.stage_bytes);
    for (int i = tid; i < 4 * 16 * 8; i += C1_THREADS) {
      const int pl = i >> 7, n = (i >> 3) & 15, k8 = i & 7, k = pl * 8 + k8;
      wimg[i] = __float2bfloat16_rn(k < 9 * CI ? op.w[(size_t)k * CoP + co0 + n] : 0.f);
    }
    if (tid < 16) sbias[tid] = op.bias ? op.bias[co0 + tid] : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int W = op.Win, H = op.Hin;
  const int Wl = 31 - __clz(W), TR = C1_TPOS >> Wl, PWx = W + 2;
  const int tiles_per_img = H / TR;

  if (warp < C1_MMA_WARP) {
    // ================= builders =================
    const float* xg = (const float*)op.x.p;
    auto issue_x = [&](int it) {
      const int s = it % C1_NSTAGE;
      const int t = blockIdx.x + it * gridDim.x;
      const int n = t / tiles_per_img, y0 = (t - n * tiles_per_img) * TR;
      const uint32_t xs = s0 + s * a.stage_bytes + 4 * C1_PB;
      const int chunks = (TR + 2) * W;  // one pixel (CI floats) per copy
      for (int i = tid; i < chunks; i += C1_BUILDERS) {
        const int rr = i >> Wl, cc = i & (W - 1);
        const int gy = y0 - 1 + rr;
        const bool ok = gy >= 0 && gy < H;
        const float* src = xg + ((size_t)(n * H + (ok ? gy : 0)) * W + cc) * CI;
        const uint32_t dst = xs + (uint32_t)((rr * PWx + cc + 1) * CI) * 4u;
        if (CI == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(ok ? 8 : 0) : "memory");
        else {
#pragma unroll
          for (int c = 0; c < CI; ++c)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4u * c), "l"(src + c), "r"(ok ? 4 : 0) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (my_tiles > 0) issue_x(0);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % C1_NSTAGE;
      const bool more = it + 1 < my_tiles;
      if (more) {
        mbar_wait_c1(empty_bar((it + 1) % C1_NSTAGE), (((uint32_t)((it + 1) / C1_NSTAGE)) & 1u) ^ 1u);
        issue_x(it + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (it == 0) mbar_wait_c1(empty_bar(0), 1u);  // fresh barrier: passes
      asm volatile("bar.sync 1, %0;" ::"n"(C1_BUILDERS) : "memory");  // every builder's row copies have landed
      const uint32_t sR = s0 + s * a.stage_bytes;
      const float* xs = reinterpret_cast<const float*>(smem + (size_t)s * a.stage_bytes + 4 * C1_PB);
      const int r = tid >> Wl, c = tid & (W - 1);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float* px = xs + ((r + tap / 3) * PWx + c + tap % 3) * CI;
        if (CI == 2) {
          const float2 t = *reinterpret_cast<const float2*>(px);
          v[tap * CI] = t.x; v[tap * CI + 1] = t.y;
        } else {
#pragma unroll
          for (int ch = 0; ch < CI; ++ch) v[tap * CI + ch] = px[ch];
        }
      }
#pragma unroll
      for (int pl = 0; pl < 4; ++pl)
        sts16c(sR + pl * C1_PB + tid * 16, make_uint4(pk2c(v[8 * pl], v[8 * pl + 1]), pk2c(v[8 * pl + 2], v[8 * pl + 3]),
                                                      pk2c(v[8 * pl + 4], v[8 * pl + 5]), pk2c(v[8 * pl + 6], v[8 * pl + 7])));
      fence_proxy_async();
      mbar_arrive_c1(full_bar(s));
    }
  } else if (warp == C1_MMA_WARP) {
    // ================= MMA issue =================
    const uint32_t idesc = instr_desc(128, 16);
    const uint64_t bd0 = smem_desc(sW, 256, 128);  // K-major planar: LBO = plane stride (16 rows x 16 B), SBO = 8 rows
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % C1_NSTAGE, q = it & 1;
      mbar_wait_c1(tempty_bar(q), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      mbar_wait_c1(full_bar(s), ((uint32_t)(it / C1_NSTAGE)) & 1u);
      tc_fence_after();
      if (elect_one_sync_c1()) {
        const uint64_t ad0 = smem_desc(s0 + s * a.stage_bytes, C1_PB, 128);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16(tmem + q * 32 + mt * 16, ad0 + (uint64_t)(mt * 128 + ks * ((2 * C1_PB) >> 4)), bd0 + (uint64_t)(ks * ((2 * 256) >> 4)), idesc,
                     ks > 0 ? 1u : 0u);
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
      const long long p0 = (long long)(blockIdx.x + it * gridDim.x) * C1_TPOS;
      mbar_wait_c1(tfull_bar(q), ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const size_t p = (size_t)(p0 + mt * 128 + lq * 32 + lane);
        uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
        uint32_t mbits = 0;
        if (op.act == ACT_MASK) {
          if (op.bits_in) {
            mbits = op.bits_in[p * (size_t)(op.Co >> 4) + (co0 >> 4)];
          } else {
            const uint4* mp = reinterpret_cast<const uint4*>((const bf16*)op.mask.p + p * op.mask.pitch + op.mask.coff + co0);
            m0 = mp[0]; m1 = mp[1];
          }
        }
        float v[16];
        tmem_ld16(tmem + lane_base + q * 32 + mt * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += sbias[j];
        if (op.act == ACT_LRELU) {
          if (op.bits_out) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            op.bits_out[p * (size_t)(op.Co >> 4) + (co0 >> 4)] = (unsigned short)w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
        } else if (op.act == ACT_MASK && op.bits_in) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= ((mbits >> j) & 1u) ? 1.f : op.slope;
        } else if (op.act == ACT_MASK) {
          const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[2 * k] *= (__uint_as_float(w[k] << 16) > 0.f ? 1.f : op.slope);
            v[2 * k + 1] *= (__uint_as_float(w[k] & 0xFFFF0000u) > 0.f ? 1.f : op.slope);
          }
        }
        uint4* dst = reinterpret_cast<uint4*>((bf16*)op.y.p + p * op.y.pitch + op.y.coff + co0);
        dst[0] = make_uint4(pk2c(v[0], v[1]), pk2c(v[2], v[3]), pk2c(v[4], v[5]), pk2c(v[6], v[7]));
        dst[1] = make_uint4(pk2c(v[8], v[9]), pk2c(v[10], v[11]), pk2c(v[12], v[13]), pk2c(v[14], v[15]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_c1(tempty_bar(q));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C1_MMA_WARP) tmem_dealloc(tmem, 64);
}

bool plan_c1(const ConvOp& op, C1Args& a) {
  if (op.transposed || op.stride != 1 || op.shuffle != SHUF_NONE || op.r1.p || op.r2.p || op.s_acc != 1.f) return false;
  if (op.Hin != op.Hout || op.Win != op.Wout) return false;
  if (op.Ci < 1 || op.Ci > 3 || op.Co % 16 || !op.w) return false;
  if (op.x.bf || op.x.pitch != op.Ci || op.x.coff != 0) return false;
  if (!op.y.bf || op.y.pitch % 8 || op.y.coff % 8) return false;
  if (op.act == ACT_MASK && (!op.mask.bf || op.mask.pitch % 8 || op.mask.coff % 8)) return false;
  const int W = op.Win;
  if ((W & (W - 1)) || W > C1_TPOS || W < 2) return false;
  const int TR = C1_TPOS / W;
  if (op.Hin % TR) return false;
  if ((long long)op.B * op.Hin * op.Win >= (1LL << 31)) return false;
  a.op = op;
  a.tiles_total = (int)((long long)op.B * op.Hin * op.Win / C1_TPOS);
  a.xs_bytes = ((TR + 2) * (W + 2) * op.Ci * 4 + 127) & ~127;
  a.stage_bytes = 4 * C1_PB + a.xs_bytes;
  return true;
}

}  // namespace

bool conv_l1_supported(const ConvOp& op) {
  C1Args a;
  return plan_c1(op, a);
}

int conv_l1(const ConvOp& op, cudaStream_t st) {
  if (ablate(5)) return 0;
  C1Args a;
  if (!plan_c1(op, a)) { set_error("conv_l1: unsupported shape"); return DG_ERR_INVALID; }
  const size_t smem = (size_t)C1_NSTAGE * a.stage_bytes + 1024;
  const long long px = (long long)op.B * op.Hout * op.Wout;
  Prof prof(PC_CONV_UMMA, 2.0 * px * op.Co * op.Ci * 9.0, (double)px * op.Co * 2.0 + (double)px * op.Ci * 4.0, st);
  const int n_chunks = op.Co / 16;
  int gx = std::max(1, std::min(a.tiles_total, (148 * 2) / n_chunks));
  const int per = (a.tiles_total + gx - 1) / gx;
  gx = (a.tiles_total + per - 1) / per;
#define C1_LAUNCH(CI)                                                                                                   \
  do {                                                                                                                  \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      DG_CUDA(cudaFuncSetAttribute(conv_l1_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));       \
      DG_CUDA(cudaFuncSetAttribute(conv_l1_kernel<CI>, cudaFuncAttributePreferredSharedMemoryCarveout,                  \
                                   cudaSharedmemCarveoutMaxShared));                                                    \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    conv_l1_kernel<CI><<<dim3(gx, n_chunks), C1_THREADS, smem, st>>>(a);                                                \
  } while (0)
  if (op.Ci == 1) C1_LAUNCH(1);
  else if (op.Ci == 2) C1_LAUNCH(2);
  else C1_LAUNCH(3);
#undef C1_LAUNCH
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
