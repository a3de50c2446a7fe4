// Streaming implicit-GEMM 3x3 convolution on tcgen05 / TMEM for sm_100a: BOTH operands arrive by TMA per K-step.
//
// Why a second conv kernel (profiles/ws_trace_r02.md): the weights-stationary kernel of dg_umma_conv_ws.cu keeps a tile's
// input resident and the CTA's weight chunk resident, which is right for the wide, shallow layers (Ci, Co = 16..32 on 64x64 /
// 128x128 maps).  On the late critic layers (Ci, Co = 64..128 on 32x32 .. 8x8 maps, critic.py:53-92) its per-tile timeline shows
// the MMA chain itself as the limit: the weights of a layer (147 - 295 KB) do not fit next to the input ring, so the output
// channels are split into N = 32 / 64 chunks whose tcgen05.mma are bound by the shared-memory operand feed (A 4 KB + B 1 - 2 KB
// per instruction: ~50 clk for 16 - 32 clk of tensor work), and a tile is rows of ONE image in padded-row order, so an 8x8 map
// fills 36 - 72 of the 128 MMA rows.  Here
//   * an M-tile is 128 OUTPUT positions = one TMA box {Cb channels, bw, bh, bn images} with bw*bh*bn = 128 (8x8 maps: two images
//     per tile, 16x16: half an image, 32x32: four rows), so every MMA row is a real output,
//   * N = all output channels (<= 256) in one accumulator, so an MMA is tensor-bound (M128 N128 K16 = 64 clk for 8 KB of operands),
//   * the K loop runs over (tap, 64-channel block): per step one A box - the input shifted by the tap, conv padding = TMA
//     out-of-bounds zero fill, stride 2 = TMA element strides - and one B box [N x Cb] of a K-major bf16 weight image
//     [tap][Co][Ci], both SWIZZLE_128B (64B / 32B for Cb = 32 / 16), into a ring of 4 - 7 stages,
//   * stride-2 data gradients run as four output-parity classes with 1 / 2 / 2 / 4 taps each (no zero insertion).
// The price is L2 -> SMEM traffic: the input is re-read once per tap (9x) and the weights once per tile; both stay in the 126 MB
// L2, whose ~12 TB/s is the bound this kernel runs against (DESIGN.md §3.1b has the per-layer arithmetic).
// Roles as in the weights-stationary kernel: warp 0 TMA producer, warp 1 MMA issue (one elected lane), warps 2..9 epilogue
// (TMEM -> registers -> bias / residuals / LeakyReLU or mask -> bf16 -> NHWC store), two accumulator stages.
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include <cuda.h>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int IG_THREADS = 320;
constexpr int IG_PROD_WARP = 0;
constexpr int IG_MMA_WARP = 1;
constexpr int IG_EPI_WARP0 = 2;
constexpr int IG_EPI_WARPS = 8;
constexpr int IG_MAX_STAGE = 8;
constexpr int IG_MAX_SMEM = 227 * 1024 - 2048;

struct IgTap { int dy, dx, wt; };

struct IgArgs {
  ConvOp op;
  int Cb, nblk, kpb;          // channel block (16 / 32 / 64), blocks per tap, K = 16 steps per block
  int N;                      // MMA columns = output channels rounded up to 16
  int bw, bh, bn, lbw, lbh;   // tile box in positions (powers of two, bw*bh*bn = 128)
  int tiles_x, tiles_y, tiles_n;
  int Ht, Wt;                 // tile-space grid (output grid; the INPUT grid for stride-2 data gradients)
  int es;                     // TMA element stride of the input map (2: stride-2 forward)
  int out_s;                  // output pixel = tile-space pixel * out_s + class parity (2: stride-2 data gradient)
  int ncls;
  int ntaps[4];
  IgTap taps[4][9];
  int nstage;
  unsigned a_bytes, stage_bytes, tx_bytes;
  int items_total;            // tiles * classes
  int tmem_cols;
};

__device__ __forceinline__ void ig_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ig_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ig_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void ig_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// bounded wait (try_wait suspends in hardware between polls); a pipeline bug must trap, not hang the GPU box
__device__ __forceinline__ void ig_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void ig_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ig_tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t ig_elect() {
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
__device__ __forceinline__ void ig_ld16(const TV& t, size_t i, float* v) {
  if (t.bf) {
    const uint4* p = reinterpret_cast<const uint4*>((const bf16*)t.p + i);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 q = p[h];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[h * 8 + 2 * k] = __uint_as_float(w[k] << 16);
        v[h * 8 + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>((const float*)t.p + i);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 q = p[h];
      v[4 * h] = q.x; v[4 * h + 1] = q.y; v[4 * h + 2] = q.z; v[4 * h + 3] = q.w;
    }
  }
}
__device__ __forceinline__ void ig_st16(const TV& t, size_t i, const float* v) {
  if (t.bf) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    uint4* p = reinterpret_cast<uint4*>((bf16*)t.p + i);
    p[0] = make_uint4(w[0], w[1], w[2], w[3]);
    p[1] = make_uint4(w[4], w[5], w[6], w[7]);
  } else {
    float4* p = reinterpret_cast<float4*>((float*)t.p + i);
#pragma unroll
    for (int h = 0; h < 4; ++h) p[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
  }
}

__global__ void __launch_bounds__(IG_THREADS, 1) conv_ig_kernel(const __grid_constant__ CUtensorMap amap,
                                                                const __grid_constant__ CUtensorMap bmap, const IgArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * IG_MAX_STAGE + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[256];
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.nstage;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (IG_MAX_STAGE + s); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (2 * IG_MAX_STAGE + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (2 * IG_MAX_STAGE + 2 + q); };

  const uint32_t sa0 = (smem_u32(smem) + 1023u) & ~1023u;  // swizzle atoms repeat every 1024 bytes
  const int my_items = ((int)blockIdx.x < a.items_total) ? (a.items_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int tiles_per_n = a.tiles_x * a.tiles_y;
  // item -> (class, tile origin in tile space)
  auto decode_item = [&](int item, int& cls, int& n0, int& y0, int& x0) {
    cls = item % a.ncls;
    int t = item / a.ncls;
    const int tn = t / tiles_per_n;
    t -= tn * tiles_per_n;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    n0 = tn * a.bn; y0 = ty * a.bh; x0 = tx * a.bw;
  };
  // producer state: the K-stage stream of this CTA = for item, for tap, for channel block
  int p_it = 0, p_t = 0, p_blk = 0, p_k = 0;
  auto produce = [&](int upto_k) {  // whole producer warp; issues K-stages [p_k, upto_k) (or until the items run out)
    while (p_it < my_items && p_k < upto_k) {
      int cls, n0, y0, x0;
      decode_item(blockIdx.x + p_it * gridDim.x, cls, n0, y0, x0);
      const IgTap tp = a.taps[cls][p_t];
      const int s = p_k % S;
      ig_wait(empty_bar(s), (((uint32_t)(p_k / S)) & 1u) ^ 1u);
      if (ig_elect()) {
        const uint32_t sa = sa0 + s * a.stage_bytes;
        ig_expect_tx(full_bar(s), a.tx_bytes);
        ig_tma_4d(sa, &amap, p_blk * a.Cb, x0 * a.es + tp.dx, y0 * a.es + tp.dy, n0, full_bar(s));
        ig_tma_3d(sa + a.a_bytes, &bmap, p_blk * a.Cb, 0, tp.wt, full_bar(s));
      }
      __syncwarp();
      ++p_k;
      if (++p_blk == a.nblk) { p_blk = 0; if (++p_t == a.ntaps[cls]) { p_t = 0; ++p_it; } }
    }
  };

  if (warp == IG_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  if (warp == IG_PROD_WARP) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap) : "memory");
      for (int s = 0; s < IG_MAX_STAGE; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int q = 0; q < 2; ++q) {
        mbar_init(tfull_bar(q), 1);
        mbar_init(tempty_bar(q), IG_EPI_WARPS);
      }
    }
    __syncwarp();
    // the ring is filled BEFORE the block-wide barrier: the first loads' latency (tensor-map fetch + L2 / HBM round trip)
    // overlaps the TMEM allocation, the bias staging and the barrier itself
    produce(S);
  }
  for (int i = tid; i < a.N; i += IG_THREADS) sbias[i] = (op.bias && i < op.Co) ? op.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == IG_PROD_WARP) {
    // ================= TMA producer: one A box + one B box per (tap, channel block) =================
    produce(0x7fffffff);
  } else if (warp == IG_MMA_WARP) {
    // ================= MMA issue (whole warp runs the uniform loops, one elected lane issues) =================
    const uint32_t idesc = instr_desc(128, a.N);
    // K-major swizzled operands, rows of Cb*2 bytes, 8-row atoms: SWIZZLE_32B / 64B / 128B layout codes 6 / 4 / 2
    const uint64_t layout = (a.kpb == 1) ? 6ull : (a.kpb == 2 ? 4ull : 2ull);
    const uint64_t desc0 = smem_desc(0, 16, 8u * 32u * (uint32_t)a.kpb) | (layout << 61);
    int k = 0;
    for (int it = 0; it < my_items; ++it) {
      const int q = it & 1;
      const int cls = (blockIdx.x + it * gridDim.x) % a.ncls;
      const int nk = a.ntaps[cls] * a.nblk;
      ig_wait(tempty_bar(q), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem + q * a.N;
      for (int kk = 0; kk < nk; ++kk, ++k) {
        const int s = k % S;
        ig_wait(full_bar(s), ((uint32_t)(k / S)) & 1u);
        tc_fence_after();
        if (ig_elect()) {
          const uint32_t a16 = (sa0 + s * a.stage_bytes) >> 4, b16 = a16 + (a.a_bytes >> 4);
          for (int kc = 0; kc < a.kpb; ++kc)
            umma_f16(d_tmem, desc0 + (uint64_t)(a16 + 2 * kc), desc0 + (uint64_t)(b16 + 2 * kc), idesc, (kk > 0 || kc > 0) ? 1u : 0u);
          umma_commit(empty_bar(s));  // ring slot reusable once these MMAs have read it
          if (kk == nk - 1) umma_commit(tfull_bar(q));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue =================
    const int lq = warp & 3;                      // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - IG_EPI_WARP0) >> 2;  // the two warps of a quarter alternate 16-column pieces
    const int r = lq * 32 + lane;                 // accumulator row = position inside the tile box
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const int ix = r & (a.bw - 1), iy = (r >> a.lbw) & (a.bh - 1), in = r >> (a.lbw + a.lbh);
    const int npieces = a.N >> 4;
    const bool pre_mask = (op.act == ACT_MASK) && op.mask.bf;
    for (int it = 0; it < my_items; ++it) {
      const int q = it & 1;
      int cls, n0, y0, x0;
      decode_item(blockIdx.x + it * gridDim.x, cls, n0, y0, x0);
      const int n = n0 + in, yt = y0 + iy, xt = x0 + ix;
      const bool valid = (n < op.B) && (yt < a.Ht) && (xt < a.Wt);
      const int yo = yt * a.out_s + (cls >> 1), xo = xt * a.out_s + (cls & 1);
      const size_t pix = ((size_t)n * op.Hout + yo) * op.Wout + xo;
      auto mask_ptr = [&](int nc) {
        return reinterpret_cast<const uint4*>((const bf16*)op.mask.p + pix * op.mask.pitch + op.mask.coff + nc);
      };
      uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
      if (pre_mask && valid && half < npieces && half * 16 < op.Co) { const uint4* mp = mask_ptr(half * 16); m0 = mp[0]; m1 = mp[1]; }
      ig_wait(tfull_bar(q), ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      const uint32_t acc0 = tmem + lane_base + q * a.N;
      for (int p = half; p < npieces; p += 2) {
        const int nc = p * 16;
        uint32_t raw[16];
        ig_tmem_ld16(acc0 + nc, raw);
        uint4 nm0 = m0, nm1 = m1;
        if (pre_mask && valid && p + 2 < npieces && (p + 2) * 16 < op.Co) { const uint4* mp = mask_ptr((p + 2) * 16); nm0 = mp[0]; nm1 = mp[1]; }
        ig_tmem_wait();
        if (valid && nc < op.Co) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          if (op.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += sbias[nc + j];
          }
          if (op.s_acc != 1.f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= op.s_acc;
          }
          if (op.r1.p) {
            float t[16];
            ig_ld16(op.r1, pix * op.r1.pitch + op.r1.coff + nc, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s1, t[j], v[j]);
          }
          if (op.r2.p) {
            float t[16];
            ig_ld16(op.r2, pix * op.r2.pitch + op.r2.coff + nc, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s2, t[j], v[j]);
          }
          if (op.act == ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
          } else if (op.act == ACT_MASK) {
            if (pre_mask) {
              const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xFFFF0000u);
                v[2 * k] *= (lo > 0.f ? 1.f : op.slope);
                v[2 * k + 1] *= (hi > 0.f ? 1.f : op.slope);
              }
            } else {
              float t[16];
              ig_ld16(op.mask, pix * op.mask.pitch + op.mask.coff + nc, t);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= (t[j] > 0.f ? 1.f : op.slope);
            }
          }
          if (op.Co < 16) {  // narrow layer: only the first Co columns exist
            const size_t o = pix * op.y.pitch + op.y.coff;
            if (op.y.bf) {
#pragma unroll
              for (int j = 0; j < 15; ++j)  // static indices: v[] stays in registers
                if (j < op.Co) ((bf16*)op.y.p)[o + j] = __float2bfloat16_rn(v[j]);
            } else if (op.Co == 2 && ((o & 1) == 0)) {
              *reinterpret_cast<float2*>((float*)op.y.p + o) = make_float2(v[0], v[1]);
            } else {
#pragma unroll
              for (int j = 0; j < 15; ++j)
                if (j < op.Co) ((float*)op.y.p)[o + j] = v[j];
            }
          } else {
            ig_st16(op.y, pix * op.y.pitch + op.y.coff + nc, v);
          }
        }
        m0 = nm0; m1 = nm1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ig_mbar_arrive(tempty_bar(q));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == IG_MMA_WARP) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

bool plan_ig(const ConvOp& op, IgArgs& a) {
  if (!op.w_ig || !op.x.bf) return false;
  const bool narrow = op.Co < 16;  // e.g. conv3.2 (F -> 2): N = 16 MMA columns of a zero-padded weight image, only Co of them stored
  if (op.Ci % 16 || (!narrow && op.Co % 16) || op.Co > 256 || op.Co < 1) return false;
  if (narrow && (!op.narrow_ok || op.r1.p || op.r2.p || op.act == ACT_MASK)) return false;
  if (op.shuffle != SHUF_NONE) return false;
  if (op.x.pitch % 8 || op.x.coff % 8) return false;
  auto aligned = [](const TV& t) {
    if (!t.p) return true;
    return t.bf ? (t.pitch % 8 == 0 && t.coff % 8 == 0) : (t.pitch % 4 == 0 && t.coff % 4 == 0);
  };
  if (!aligned(op.r1) || !aligned(op.r2) || !aligned(op.mask) || (!narrow && !aligned(op.y))) return false;
  int mode;  // 0: stride 1, 1: stride-2 forward, 2: stride-2 data gradient
  if (op.transposed) {
    if (op.Hout != 2 * op.Hin || op.Wout != 2 * op.Win) return false;
    mode = 2;
  } else if (op.stride == 2) {
    if (op.Hin != 2 * op.Hout || op.Win != 2 * op.Wout) return false;
    mode = 1;
  } else {
    if (op.stride != 1 || op.Hin != op.Hout || op.Win != op.Wout) return false;
    mode = 0;
  }
  a = IgArgs{};
  a.op = op;
  a.Ht = (mode == 2) ? op.Hin : op.Hout;
  a.Wt = (mode == 2) ? op.Win : op.Wout;
  if (!is_pow2(a.Wt) || !is_pow2(a.Ht)) return false;  // tile boxes are powers of two; other grids stay on the halo-tile kernels
  a.Cb = (op.Ci % 64 == 0) ? 64 : (op.Ci % 32 == 0 ? 32 : 16);
  a.nblk = op.Ci / a.Cb;
  a.kpb = a.Cb / 16;
  a.N = round_up(op.Co, 16);
  a.bw = std::min(a.Wt, 128);
  a.bh = std::min(a.Ht, 128 / a.bw);
  a.bn = 128 / (a.bw * a.bh);
  a.lbw = ilog2(a.bw); a.lbh = ilog2(a.bh);
  a.tiles_x = a.Wt / a.bw; a.tiles_y = a.Ht / a.bh; a.tiles_n = (op.B + a.bn - 1) / a.bn;
  a.es = (mode == 1) ? 2 : 1;
  a.out_s = (mode == 2) ? 2 : 1;
  if (a.bw * a.es > 256 || a.bh * a.es > 256) return false;  // TMA box extent (traversal) per dimension
  a.ncls = (mode == 2) ? 4 : 1;
  for (int cls = 0; cls < a.ncls; ++cls) {
    int n = 0;
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      IgTap t;
      t.wt = tap;
      if (mode == 0 || mode == 1) { t.dy = ky - 1; t.dx = kx - 1; }
      else {
        // output pixel (2y + py, 2x + px) gathers dz[y + (ky == 0), x + (kx == 0)] over the taps of matching parity
        const int py = cls >> 1, px = cls & 1;
        if (((py == 0) != (ky == 1)) || ((px == 0) != (kx == 1))) continue;
        t.dy = (ky == 0) ? 1 : 0; t.dx = (kx == 0) ? 1 : 0;
      }
      a.taps[cls][n++] = t;
    }
    a.ntaps[cls] = n;
  }
  a.a_bytes = (unsigned)(128 * a.Cb * 2);
  const unsigned b_bytes = (unsigned)(a.N * a.Cb * 2);
  a.tx_bytes = a.a_bytes + b_bytes;
  a.stage_bytes = (a.a_bytes + b_bytes + 1023u) & ~1023u;
  a.nstage = (int)std::min<size_t>(IG_MAX_STAGE, (size_t)(IG_MAX_SMEM - 1024) / a.stage_bytes);
  if (a.nstage < 2) return false;
  a.items_total = a.tiles_x * a.tiles_y * a.tiles_n * a.ncls;
  int pc = 32;
  while (pc < 2 * a.N) pc <<= 1;
  a.tmem_cols = pc;
  return true;
}

struct IgMapKey {
  const void* base; int a, b, c, d, e, f, g, h;
  bool operator==(const IgMapKey& o) const {
    return base == o.base && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && f == o.f && g == o.g && h == o.h;
  }
};
std::vector<std::pair<IgMapKey, CUtensorMap>> g_ig_maps;
std::mutex g_ig_mu;

CUtensorMapSwizzle ig_swizzle(int Cb) {
  return Cb == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (Cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// input: NHWC view {C, W, H, B} (bf16), box {Cb, bw*es, bh*es, bn} with element strides {1, es, es, 1}
int ig_map_a(const ConvOp& op, const IgArgs& a, CUtensorMap* out) {
  IgMapKey k{(const bf16*)op.x.p + op.x.coff, op.Ci, op.Win, op.Hin, op.B, op.x.pitch, a.bw * a.es, a.bh * a.es, a.bn * 4 + a.es};
  std::lock_guard<std::mutex> lock(g_ig_mu);
  for (auto& e : g_ig_maps)
    if (e.first == k) { *out = e.second; return 0; }
  cuuint64_t dims[4] = {(cuuint64_t)op.Ci, (cuuint64_t)op.Win, (cuuint64_t)op.Hin, (cuuint64_t)op.B};
  cuuint64_t strides[3] = {(cuuint64_t)op.x.pitch * 2, (cuuint64_t)op.Win * op.x.pitch * 2, (cuuint64_t)op.Hin * op.Win * op.x.pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)a.Cb, (cuuint32_t)(a.bw * a.es), (cuuint32_t)(a.bh * a.es), (cuuint32_t)a.bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)a.es, (cuuint32_t)a.es, 1};
  CUtensorMap m;
  const CUresult r = encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(k.base), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, ig_swizzle(a.Cb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_ig: cuTensorMapEncodeTiled(input) failed (%d): C %d W %d H %d B %d pitch %d box %d x %d x %d es %d", (int)r, op.Ci,
              op.Win, op.Hin, op.B, op.x.pitch, a.bw * a.es, a.bh * a.es, a.bn, a.es);
    return DG_ERR_CUDA;
  }
  if (g_ig_maps.size() > 4096) g_ig_maps.clear();
  g_ig_maps.emplace_back(k, m);
  *out = m;
  return 0;
}

// weights: K-major image [9][N][Ci] (bf16) as {Ci, N, 9}, box {Cb, N, 1}
int ig_map_b(const ConvOp& op, const IgArgs& a, CUtensorMap* out) {
  IgMapKey k{op.w_ig, op.Ci, a.N, 9, a.Cb, -1, -1, -1, -1};
  std::lock_guard<std::mutex> lock(g_ig_mu);
  for (auto& e : g_ig_maps)
    if (e.first == k) { *out = e.second; return 0; }
  cuuint64_t dims[3] = {(cuuint64_t)op.Ci, (cuuint64_t)a.N, 9};
  cuuint64_t strides[2] = {(cuuint64_t)op.Ci * 2, (cuuint64_t)a.N * op.Ci * 2};
  cuuint32_t box[3] = {(cuuint32_t)a.Cb, (cuuint32_t)a.N, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(op.w_ig), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, ig_swizzle(a.Cb), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_ig: cuTensorMapEncodeTiled(weights) failed (%d): Ci %d N %d Cb %d", (int)r, op.Ci, a.N, a.Cb);
    return DG_ERR_CUDA;
  }
  if (g_ig_maps.size() > 4096) g_ig_maps.clear();
  g_ig_maps.emplace_back(k, m);
  *out = m;
  return 0;
}

// fp32 packed [tap][Ci][CoP] -> bf16 K-major [tap][CoP][Ci]; element offsets are shared with the fp32 packed buffer.
// Two (source, destination, table) sets per launch: the forward and the data-gradient images of a network in one go.
struct PackIgSet { const float* packed; bf16* dst; const UmmaPackDesc* table; int n; };
__global__ void pack_ig_kernel(const PackIgSet s0, const PackIgSet s1) {
  const bool second = (int)blockIdx.y >= s0.n;
  const PackIgSet& s = second ? s1 : s0;
  const int e = second ? (int)blockIdx.y - s0.n : (int)blockIdx.y;
  if (e >= s.n) return;
  const UmmaPackDesc d = s.table[e];
  const int total = 9 * d.Ci * d.CoP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % d.Ci;
    const int t = i / d.Ci;
    const int co = t % d.CoP, tap = t / d.CoP;
    s.dst[d.off + i] = __float2bfloat16_rn(s.packed[d.off + ((size_t)tap * d.Ci + ci) * d.CoP + co]);
  }
}

}  // namespace

int pack_ig2(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, const float* packed2,
             void* dst2_bf16, const UmmaPackDesc* table2_dev, int n2, int max_elems2, cudaStream_t st) {
  if (n + n2 <= 0) return 0;
  const int bx = std::max(1, std::min(64, (std::max(max_elems, max_elems2) + 255) / 256));
  const PackIgSet s0{packed, (bf16*)dst_bf16, table_dev, n}, s1{packed2, (bf16*)dst2_bf16, table2_dev, n2};
  pack_ig_kernel<<<dim3(bx, n + n2), 256, 0, st>>>(s0, s1);
  DG_LAUNCH_CHECK();
  return 0;
}
int pack_ig(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, cudaStream_t st) {
  return pack_ig2(packed, dst_bf16, table_dev, n, max_elems, nullptr, nullptr, nullptr, 0, 1, st);
}

bool conv_ig_supported(const ConvOp& op) {
  IgArgs a;
  return plan_ig(op, a);
}

bool umma_ws_supported(const ConvOp& op);
bool umma_supported(const ConvOp& op);

// Shapes on which the streaming kernel is the faster of the two tcgen05 conv kernels (measured on the B200,
// profiles/conv_ig_ab_r02.md: kernel time, ws / ig): 128 -> 128 stride 2 on 16x16 (critic L7) 2.1x forward, 1.8x data gradient;
// 64 -> 128 on 16x16 (L6) 1.16x forward, 0.93x data gradient; 64 -> 64 stride 2 (L5) 0.9x at 192 images but 1.07 - 1.14x at 64
// (the JVP / generator-iteration passes); everything with fewer than 64 channels on either side or maps above 32x32 is 1.5 - 6x
// slower (the 9x re-read of a shallow, wide input costs more than the small-N MMAs it avoids).  Shapes the weights-stationary
// kernel does not take at all (Ci not in {16,32,64,128}: the dense-block layers of F = 32 / 64 generators, 256-channel critic
// layers) always come here: 3 - 9x faster than the per-tile cp.async kernel, 10x+ over the CUDA cores.
// DG_IG=0 never, DG_IG=2 wherever it is supported (tests, tools/ig_ab.py).
bool conv_ig_preferred(const ConvOp& op) {
  static const int force = getenv("DG_IG") ? atoi(getenv("DG_IG")) : 1;
  if (force == 0 || !g_tune[16]) return false;
  IgArgs a;
  if (!plan_ig(op, a)) return false;
  if (force == 2) return true;
  if (!(op.w_umma && umma_ws_supported(op))) return op.Co >= 64 || op.Co < 16 || !(op.w_umma && umma_supported(op));  // (Co = 32: 0.8 - 0.9x of the cp.async kernel)
  const int npos = a.Ht * a.Wt;
  if (op.Ci >= 128 && op.Co >= 128 && npos <= 256) return true;                               // L7 forward / data gradient / JVP
  if (op.Ci >= 64 && op.Co >= 128 && npos <= 256 && !op.transposed) return true;                // L6 forward / JVP
  if (op.Ci >= 64 && op.Co >= 64 && npos <= 256 && op.B <= 96) return true;                      // L5 .. L7, both directions, small batches
  return false;
}

int conv_ig(const ConvOp& op, cudaStream_t st) {
  if (ablate(7)) return 0;
  IgArgs a;
  if (!plan_ig(op, a)) { set_error("conv_ig: unsupported shape"); return DG_ERR_INVALID; }
  CUtensorMap ma, mb;
  DG_TRY(ig_map_a(op, a, &ma));
  DG_TRY(ig_map_b(op, a, &mb));
  const size_t smem = (size_t)a.nstage * a.stage_bytes + 1024;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const double taps = op.transposed ? 2.25 : 9.0;
  Prof prof(PC_CONV_UMMA, 2.0 * total * op.Co * op.Ci * taps,
            (double)total * op.Co * (op.y.bf ? 2 : 4) + (double)op.B * op.Hin * op.Win * op.Ci * 2.0, st);
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(conv_ig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_MAX_SMEM));
    attr_set = true;
  }
  int gx = std::min(a.items_total, 148);
  const int per = (a.items_total + gx - 1) / gx;
  gx = (a.items_total + per - 1) / per;  // even out the tail
  static const bool show_plan = getenv("DG_IG_PLAN") != nullptr;
  if (show_plan)
    fprintf(stderr, "[ig plan] Ci %d Co %d Hout %d B %d stride %d transposed %d: Cb %d nblk %d N %d box %dx%dx%d stages %d items %d grid %d smem %zu\n",
            op.Ci, op.Co, op.Hout, op.B, op.stride, op.transposed, a.Cb, a.nblk, a.N, a.bw, a.bh, a.bn, a.nstage, a.items_total, gx, smem);
  conv_ig_kernel<<<gx, IG_THREADS, smem, st>>>(ma, mb, a);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
