// Warp-specialised, persistent, TMA-fed tcgen05 / TMEM implicit-GEMM 3x3 convolution for sm_100a.
//
// Same operand layouts as dg_umma_conv.cu (planar no-swizzle K-major input tile, tap = descriptor
// start-address shift, weights staged once per CTA), but the three phases of a tile run in different
// warps and overlap across tiles through mbarrier pipelines:
//
//   warp 0      producer  one elected thread issues cp.async.bulk.tensor (TMA) loads of the halo-padded
//                         tile of tile i+1, i+2 into a ring of NSTAGE shared-memory buffers: one box
//                         {8 ch, PW, rows} per 8-channel plane of an NHWC tensor map, conv padding =
//                         TMA out-of-bounds zero fill, stride-2 parity sub-images = TMA element strides
//                         (full[s]: expect_tx + complete_tx)
//   warp 1      MMA       one elected thread issues the tcgen05.mma chain of tile i into TMEM
//                         accumulator stage i&1, commits to empty[s] (ring slot free) and tfull[a]
//   warps 2..9  epilogue  tcgen05.ld of accumulator stage a (two warps per TMEM lane quarter, alternating
//                         16-column pieces), fused epilogue, global store, then tempty[a] hands the
//                         accumulator back to the MMA warp
//
// so the tensor pipe, the L2->SMEM stream and the epilogue's HBM stores of three different tiles are in
// flight at once inside one CTA.  LeakyReLU masks of the next piece are prefetched while the current
// piece is processed.  nn.PixelShuffle(2) is folded into the store: the CTA stages its weight rows
// (i,j)-major so that one 16-column accumulator piece is 16 consecutive channels of ONE output pixel.
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include <cuda.h>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int WS_THREADS = 320;
constexpr int WS_PROD_WARP = 0;
constexpr int WS_MMA_WARP = 1;
constexpr int WS_EPI_WARP0 = 2;
constexpr int WS_EPI_WARPS = 8;
constexpr int WS_MAX_STAGE = 4;
constexpr int WS_MAX_SMEM = 227 * 1024 - 3072;
enum Mode { S1 = 0, S2_FWD = 1, S2_DGRAD = 2 };

struct WsArgs {
  ConvOp op;
  int mode;
  int TH, PW, n_mt, RB, tiles_per_img, NT, nplanes, nsub, CoP;  // RB: bytes of one TMA box region (1024-aligned)
  int nblk, Cb;  // channel blocks of Cb = 16/32/64 channels = one SWIZZLE_32/64/128B row per position
  int Ht, Wt;
  unsigned magic_np, magic_pw;
  unsigned w_off;    // byte offset of the weight image in dynamic smem (after the ring)
  unsigned a_bytes;  // bytes of one ring slot
  int tiles_total, nstage;
  int acc_cols;   // TMEM columns of one accumulator stage
  int tmem_cols;  // allocation (power of two >= 2*acc_cols)
  int perm;       // 1: PixelShuffle store, weight rows staged (i,j)-major
  int Fsh;        // channels after the shuffle (Co/4) when perm
  unsigned tile_tx;  // bytes one tile's TMA loads deliver
  unsigned over;     // tail pad: bytes the last M-tile's rows may read past the last box region
  unsigned long long* trace;  // debug timeline (DG_WS_TRACE=1), null otherwise
  int early;                  // 1: the producer fills the ring before the weight staging / block barrier (dg_set_tuning(17, .))
  int bits_tile;              // 1: sign-bit masks of a whole tile are fetched before the wait for its MMAs (dg_set_tuning(21, 2))
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

// bounded wait without the clock reads of um::mbar_wait (try_wait suspends in hardware between polls)
__device__ __forceinline__ void mbar_wait_ws(uint32_t bar, uint32_t parity) {
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
    if (spin > (1u << 26)) __trap();  // a pipeline bug must trap, not hang the GPU box
  }
}

__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ld16f(const TV& t, size_t i, float* v) {
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
__device__ __forceinline__ void st16f(const TV& t, size_t i, const float* v) {
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

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace slot: [block][role 0..2][it < 16][event 0..3]
#define WS_TRACE(role, it, ev)                                                                          \
  do {                                                                                                  \
    if (a.trace && (it) < 16) a.trace[(((size_t)blockIdx.x * 3 + (role)) * 16 + (it)) * 4 + (ev)] = gtimer(); \
  } while (0)

// one 16-column accumulator piece of one tile row
struct Piece {
  bool valid;
  int n, yo, xo, nc;  // sample, tile-space output pixel, first column inside the CTA's chunk
  int col;            // TMEM column offset inside the accumulator stage
};

__device__ __forceinline__ uint32_t elect_one_sync() {
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

// SIMPLE (compile time): no bias / accumulator scale / residuals, plain NHWC bf16 store of full 16-column pieces, activation =
// none, LeakyReLU or a mask that is prefetched (sign bits or bf16) - every critic layer and most generator-tail layers.  The
// epilogue warps are bound by the length of their per-piece instruction chain; this drops the run-time checks of the
// general epilogue (about a fifth of the chain).
template <int MODE, int KCS, bool SIMPLE>
__global__ void __launch_bounds__(WS_THREADS, 2) conv_ws_kernel(const __grid_constant__ CUtensorMap tmap, const WsArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * WS_MAX_STAGE + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[256];
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * a.NT;
  constexpr int mode = MODE;
  const int PW = a.PW, S = a.nstage;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WS_MAX_STAGE + s); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (2 * WS_MAX_STAGE + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (2 * WS_MAX_STAGE + 2 + q); };

  const uint32_t sa0 = (smem_u32(smem) + 1023u) & ~1023u;  // swizzle atoms repeat every 1024 bytes
  const int my_tiles = ((int)blockIdx.x < a.tiles_total) ? (a.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // TMA loads of tile `it` into ring slot it % S (whole producer warp; one elected lane issues)
  auto produce_tile = [&](int it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const int s = it % S;
    if (lane == 0) WS_TRACE(0, it, 0);
    mbar_wait_ws(empty_bar(s), (((uint32_t)(it / S)) & 1u) ^ 1u);
    if (lane == 0) WS_TRACE(0, it, 1);
    const uint32_t sa = sa0 + s * a.a_bytes;
    const int n = tile / a.tiles_per_img;
    const int y0 = (tile - n * a.tiles_per_img) * a.TH;
    if (elect_one_sync()) {
      mbar_expect_tx(full_bar(s), a.tile_tx);
#pragma unroll
      for (int sub = 0; sub < ((MODE == S2_FWD) ? 4 : 1); ++sub) {
        int cx, cy;
        if (MODE == S1) { cx = -1; cy = y0 - 1; }
        else if (MODE == S2_FWD) { cx = -2 + (sub & 1); cy = 2 * (y0 - 1) + (sub >> 1); }
        else { cx = 0; cy = y0; }
        for (int blk = 0; blk < a.nblk; ++blk)
          tma_load_4d(sa + (sub * a.nblk + blk) * a.RB, &tmap, blk * a.Cb, cx, cy, n, full_bar(s));
      }
    }
    __syncwarp();
    if (lane == 0) WS_TRACE(0, it, 2);
  };
  const int pre_tiles = a.early ? min(S, my_tiles) : 0;

  if (warp == WS_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  if (warp == WS_PROD_WARP) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      for (int s = 0; s < WS_MAX_STAGE; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int q = 0; q < 2; ++q) {
        mbar_init(tfull_bar(q), 1);
        mbar_init(tempty_bar(q), WS_EPI_WARPS);
      }
    }
    __syncwarp();
    // The first ring slots are filled BEFORE the weights are staged and the block synchronises: the first tiles' load latency
    // (tensor-map fetch, L2 / HBM round trip) overlaps the weight staging, the TMEM allocation and the barrier.  The ring
    // [sa0, sa0 + S * a_bytes) and the weight image behind it are disjoint.  (PDL: the activations are the previous kernel's.)
    if (pre_tiles > 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int it = 0; it < pre_tiles; ++it) produce_tile(it);
  }
  const uint32_t sw = sa0 + a.w_off;
  {
    // weights: planes [tap*nplanes + pl], NT rows of 16 B each; row j of the chunk is output channel
    // co0 + j, or (PixelShuffle store) the channel whose shuffled position is (i,j)-major column co0 + j
    const int wrows = 9 * a.nplanes * a.NT;
    const uint4* wsrc = reinterpret_cast<const uint4*>(op.w_umma);
    for (int i = tid; i < wrows; i += WS_THREADS) {
      const int tp = i / a.NT, row = i - tp * a.NT;
      const int colp = co0 + row;
      const int src = a.perm ? 4 * (colp % a.Fsh) + colp / a.Fsh : colp;
      cp_async16(sw + i * 16, wsrc + (size_t)tp * a.CoP + src, 16);
    }
    for (int row = tid; row < a.NT; row += WS_THREADS) {
      const int colp = co0 + row;
      const int src = a.perm ? 4 * (colp % a.Fsh) + colp / a.Fsh : colp;
      sbias[row] = (op.bias && src < op.Co) ? op.bias[src] : 0.f;
    }
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Programmatic dependent launch: everything above (TMEM allocation, barrier init, weight staging - weights are
  // never written by the kernel launched just before a conv) may overlap the tail of the previous kernel; the
  // activations it produced are only touched after this wait.  Our own dependents may start their prologue now.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  constexpr int ncls = (MODE == S2_DGRAD) ? 4 : 1;

  if (warp == WS_PROD_WARP) {
    // ================= TMA producer =================
    for (int it = pre_tiles; it < my_tiles; ++it) produce_tile(it);
  } else if (warp == WS_MMA_WARP) {
    // ================= MMA issue =================
    // The whole warp runs the (warp-uniform) loops so that descriptors live in uniform registers; one
    // elected lane issues.  Descriptor low words are in 16-byte units: a tap is a constant add.
    const uint32_t idesc = instr_desc(128, a.NT);
    const uint32_t wplane = a.NT * 16;
    constexpr int KPB = (KCS < 4) ? KCS : 4;       // K-steps (16 channels) per channel block
    constexpr uint32_t R16 = 2 * KPB;               // bytes of one position row / 16
    constexpr uint64_t A_LAYOUT = (KPB == 1) ? 6ull : (KPB == 2 ? 4ull : 2ull);  // SWIZZLE_32B / 64B / 128B
    const uint32_t kB = (2 * wplane) >> 4;
    // A: swizzled K-major rows of 32*KPB bytes, 8-row atoms (SBO); the swizzle XOR acts on absolute smem
    // address bits, so a tap is a plain start-address shift by whole rows (tools/probes/swizzle_shift_probe.cu)
    const uint64_t adesc0 = (smem_desc(0, 16, 8 * 32 * KPB) | (A_LAYOUT << 61)), bdesc0 = smem_desc(0, wplane, 128);
    const uint32_t regA = (uint32_t)(a.RB >> 4), tapB = (uint32_t)((a.nplanes * wplane) >> 4);
    const uint32_t subA = regA * a.nblk;
    const uint32_t sw16 = sw >> 4;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % S, q = it & 1;
      if (lane == 0) WS_TRACE(1, it, 0);
      mbar_wait_ws(tempty_bar(q), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
      if (lane == 0) WS_TRACE(1, it, 1);
      mbar_wait_ws(full_bar(s), ((uint32_t)(it / S)) & 1u);
      if (lane == 0) WS_TRACE(1, it, 2);
      tc_fence_after();
      const uint32_t sa16 = (sa0 + s * a.a_bytes) >> 4;
      const uint32_t acc0 = tmem + q * a.acc_cols;
      if (elect_one_sync()) {
        for (int mt = 0; mt < a.n_mt; ++mt) {
#pragma unroll
          for (int cls = 0; cls < ncls; ++cls) {
            const uint32_t d_tmem = acc0 + (mt * ncls + cls) * a.NT;
            uint32_t acc = 0;
            constexpr int TAP_UNROLL = (KCS <= 2) ? 9 : 1;  // wide-K layers: keep the descriptor set in registers small
#pragma unroll TAP_UNROLL
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap % 3;
              uint32_t aoff;
              if (MODE == S1) {
                aoff = (ky * PW + kx) * R16;
              } else if (MODE == S2_FWD) {
                const int sub = ((ky == 1) ? 0 : 2) + ((kx == 1) ? 0 : 1);
                aoff = sub * subA + (((ky == 0) ? 0 : 1) * PW + ((kx == 0) ? 0 : 1)) * R16;
              } else {
                const int py = cls >> 1, px = cls & 1;
                if (((py == 0) != (ky == 1)) || ((px == 0) != (kx == 1))) continue;
                aoff = (((ky == 0) ? 1 : 0) * PW + ((kx == 0) ? 1 : 0)) * R16;
              }
              const uint32_t alo = sa16 + mt * 128 * R16 + aoff, blo = sw16 + tap * tapB;
#pragma unroll
              for (int kc = 0; kc < KCS; ++kc) {
                umma_f16(d_tmem, adesc0 + (uint64_t)(alo + (kc / KPB) * regA + (kc % KPB) * 2), bdesc0 + (uint64_t)(blo + kc * kB), idesc,
                         acc);
                acc = 1;
              }
            }
          }
        }
        umma_commit(empty_bar(s));  // ring slot reusable once these MMAs have read it
        umma_commit(tfull_bar(q));  // accumulator stage complete
      }
      __syncwarp();
      if (lane == 0) WS_TRACE(1, it, 3);
    }
  } else {
    // ================= epilogue =================
    const int lq = warp & 3;                     // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - WS_EPI_WARP0) >> 2;  // the two warps of a quarter alternate pieces
    const int row_in_tile = lq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const int ppr_log = 31 - __clz(a.NT >> 4);  // 16-column pieces per (m-tile, class) = 1 << ppr_log
    const int npieces = (a.n_mt * ncls) << ppr_log;
    const bool mask_bits = (op.act == ACT_MASK) && op.bits_in != nullptr;
    const bool pre_mask = (op.act == ACT_MASK) && (op.mask.bf || mask_bits);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int q = it & 1;
      const int n = tile / a.tiles_per_img;
      const int y0 = (tile - n * a.tiles_per_img) * a.TH;
      const int rows = min(a.TH, a.Ht - y0);
      auto decode = [&](int p) {
        Piece pc;
        const int u = p >> ppr_log, j = p - (u << ppr_log);
        const int mt = u / ncls, cls = u - mt * ncls;
        const int qq = mt * 128 + row_in_tile;
        const int r = (int)__umulhi((unsigned)qq, a.magic_pw), c = qq - r * PW;
        pc.valid = (r < rows) && (c < a.Wt);
        pc.n = n; pc.yo = y0 + r; pc.xo = c;
        if (mode == S2_DGRAD) { pc.yo = 2 * pc.yo + (cls >> 1); pc.xo = 2 * pc.xo + (cls & 1); }
        pc.nc = j * 16;
        pc.col = u * a.NT + j * 16;
        return pc;
      };
      auto mask_ptr = [&](const Piece& pc) {
        const size_t p = ((size_t)pc.n * op.Hout + pc.yo) * op.Wout + pc.xo;
        return reinterpret_cast<const uint4*>((const bf16*)op.mask.p + p * op.mask.pitch + op.mask.coff + co0 + pc.nc);
      };
      // sign-bit masks (ConvOp::bits_in): one 2-byte word per piece, carried in m0.x
      auto load_mask = [&](const Piece& pc, uint4& a0, uint4& a1) {
        if (mask_bits) {
          const size_t p = ((size_t)pc.n * op.Hout + pc.yo) * op.Wout + pc.xo;
          a0.x = op.bits_in[p * (size_t)(op.Co >> 4) + ((co0 + pc.nc) >> 4)];
        } else {
          const uint4* mp = mask_ptr(pc);
          a0 = mp[0]; a1 = mp[1];
        }
      };
      Piece cur = decode(half < npieces ? half : 0);
      if (half >= npieces) cur.valid = false;
      uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
      // bits_tile: the sign words of ALL this warp's pieces of the tile (2 bytes each, up to eight) are requested here, before
      // the wait for the tile's MMAs, and carried packed in two 64-bit registers - no load sits between an accumulator read
      // and its store any more
      const bool tile_bits = mask_bits && a.bits_tile && npieces <= 16;
      unsigned long long pkA = 0ull, pkB = 0ull;
      if (tile_bits) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int p = half + 2 * k;
          if (p < npieces) {
            const Piece pc = decode(p);
            if (pc.valid) {
              const size_t px = ((size_t)pc.n * op.Hout + pc.yo) * op.Wout + pc.xo;
              const unsigned long long w = op.bits_in[px * (size_t)(op.Co >> 4) + ((co0 + pc.nc) >> 4)];
              if (k < 4) pkA |= w << (16 * k); else pkB |= w << (16 * (k - 4));
            }
          }
        }
      } else if (pre_mask && cur.valid) load_mask(cur, m0, m1);
      if (tid == WS_EPI_WARP0 * 32) WS_TRACE(2, it, 0);
      mbar_wait_ws(tfull_bar(q), ((uint32_t)(it >> 1)) & 1u);
      if (tid == WS_EPI_WARP0 * 32) WS_TRACE(2, it, 1);
      tc_fence_after();
      const uint32_t acc0 = tmem + lane_base + q * a.acc_cols;
      for (int p = half; p < npieces; p += 2) {
        uint32_t raw[16];
        tmem_ld16_raw(acc0 + cur.col, raw);
        Piece nxt = cur;
        uint4 nm0 = m0, nm1 = m1;
        if (p + 2 < npieces) {
          nxt = decode(p + 2);
          if (pre_mask && !tile_bits && nxt.valid) load_mask(nxt, nm0, nm1);
        }
        if (tile_bits) {
          const int k = (p - half) >> 1;
          m0.x = (uint32_t)((k < 4 ? pkA >> (16 * k) : pkB >> (16 * (k - 4))) & 0xFFFFull);
          nm0 = m0;
        }
        tmem_ld_wait();
        if (cur.valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          const int nc = cur.nc;
          const size_t pix = ((size_t)cur.n * op.Hout + cur.yo) * op.Wout + cur.xo;
          if (!SIMPLE && op.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += sbias[nc + j];
          }
          if (!SIMPLE && op.s_acc != 1.f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= op.s_acc;
          }
          if (!SIMPLE && op.r1.p) {
            float t[16];
            ld16f(op.r1, pix * op.r1.pitch + op.r1.coff + co0 + nc, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s1, t[j], v[j]);
          }
          if (!SIMPLE && op.r2.p) {
            float t[16];
            ld16f(op.r2, pix * op.r2.pitch + op.r2.coff + co0 + nc, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s2, t[j], v[j]);
          }
          if (op.act == ACT_LRELU) {
            if (op.bits_out) {  // sign bits of this piece for the data-gradient / JVP epilogues (plain NHWC stores only)
              uint32_t w = 0;
#pragma unroll
              for (int j = 0; j < 16; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
              op.bits_out[pix * (size_t)(op.Co >> 4) + ((co0 + nc) >> 4)] = (unsigned short)w;
            }
            if (op.slope >= 0.f && op.slope <= 1.f) {  // max(v, slope*v) == LeakyReLU for 0 <= slope <= 1
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], v[j] * op.slope);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
            }
          } else if (op.act == ACT_MASK) {
            if (mask_bits) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= ((m0.x >> j) & 1u) ? 1.f : op.slope;
            } else if (SIMPLE || pre_mask) {
              const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xFFFF0000u);
                v[2 * k] *= (lo > 0.f ? 1.f : op.slope);
                v[2 * k + 1] *= (hi > 0.f ? 1.f : op.slope);
              }
            } else {
              float t[16];
              ld16f(op.mask, pix * op.mask.pitch + op.mask.coff + co0 + nc, t);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= (t[j] > 0.f ? 1.f : op.slope);
            }
          }
          if (SIMPLE) {
            uint32_t w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
              w[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
            uint4* dst = reinterpret_cast<uint4*>((bf16*)op.y.p + pix * op.y.pitch + op.y.coff + co0 + nc);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
          } else if (op.Co < 16) {  // narrow layer: only the first Co columns exist
            const size_t o = pix * op.y.pitch + op.y.coff;
            if (op.y.bf) {
#pragma unroll
              for (int j = 0; j < 15; ++j)  // static indices: v[] must stay in registers
                if (j < op.Co) ((bf16*)op.y.p)[o + j] = __float2bfloat16_rn(v[j]);
            } else if (op.Co == 2 && ((o & 1) == 0)) {
              *reinterpret_cast<float2*>((float*)op.y.p + o) = make_float2(v[0], v[1]);
            } else {
#pragma unroll
              for (int j = 0; j < 15; ++j)
                if (j < op.Co) ((float*)op.y.p)[o + j] = v[j];
            }
          } else if (op.shuffle == SHUF_NONE) {
            st16f(op.y, pix * op.y.pitch + op.y.coff + co0 + nc, v);
          } else if (op.shuffle == SHUF_PIXEL) {  // (i,j)-major columns: one piece = 16 channels of one shuffled pixel
            const int colp = co0 + nc;
            const int sq = colp / a.Fsh, cc0 = colp - sq * a.Fsh;
            const size_t qq = ((size_t)cur.n * (2 * op.Hout) + 2 * cur.yo + (sq >> 1)) * (2 * op.Wout) + 2 * cur.xo + (sq & 1);
            st16f(op.y, qq * op.y.pitch + op.y.coff + cc0, v);
          } else {  // inverse shuffle: data-gradient w.r.t. the pre-shuffle activation, kept (i,j)-major (PackDesc::ps):
            // the 16 channels of this piece are contiguous in the (2i + j) block of the coarse pixel - one 32-byte store
            const size_t qq = ((size_t)cur.n * (op.Hout >> 1) + (cur.yo >> 1)) * (op.Wout >> 1) + (cur.xo >> 1);
            st16f(op.y, qq * op.y.pitch + op.y.coff + (2 * (cur.yo & 1) + (cur.xo & 1)) * op.Co + co0 + nc, v);
          }
        }
        cur = nxt; m0 = nm0; m1 = nm1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(q));
      if (tid == WS_EPI_WARP0 * 32) WS_TRACE(2, it, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == WS_MMA_WARP) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

bool plan_ws(const ConvOp& op, WsArgs& a, int& ctas_per_sm) {
  if (!op.w_umma || !op.x.bf) return false;
  const bool narrow = op.Co < 16;  // e.g. conv3.2 (16 -> 2): N = 16 MMA columns, only Co of them stored
  if (op.Ci % 16 || (!narrow && op.Co % 16) || op.Co > 256 || op.Co < 1) return false;
  if (narrow && (!op.narrow_ok || op.shuffle != SHUF_NONE || op.r1.p || op.r2.p || op.act == ACT_MASK)) return false;
  if (op.Ci != 16 && op.Ci != 32 && op.Ci != 64 && op.Ci != 128) return false;  // KCS template instances
  if (op.x.pitch % 8 || op.x.coff % 8) return false;
  int mode;
  if (op.transposed) {
    if (op.Hout != 2 * op.Hin || op.Wout != 2 * op.Win) return false;
    mode = S2_DGRAD;
  } else if (op.stride == 2) {
    if (op.Hin != 2 * op.Hout || op.Win != 2 * op.Wout) return false;
    mode = S2_FWD;
  } else {
    if (op.stride != 1 || op.Hin != op.Hout || op.Win != op.Wout) return false;
    mode = S1;
  }
  auto aligned = [](const TV& t) {
    if (!t.p) return true;
    return t.bf ? (t.pitch % 8 == 0 && t.coff % 8 == 0) : (t.pitch % 4 == 0 && t.coff % 4 == 0);
  };
  if (!aligned(op.r1) || !aligned(op.r2) || !aligned(op.mask)) return false;
  if (!narrow && !aligned(op.y)) return false;
  const int perm = op.shuffle == SHUF_PIXEL;
  if (perm && (op.Co % 64)) return false;  // Co/4 shuffled channels in whole 16-column pieces
  const int Ht = (mode == S2_DGRAD) ? op.Hin : op.Hout, Wt = (mode == S2_DGRAD) ? op.Win : op.Wout;
  const int PW = (mode == S1) ? Wt + 2 : Wt + 1;
  const int nplanes = op.Ci / 8, nsub = (mode == S2_FWD) ? 4 : 1, ncls = (mode == S2_DGRAD) ? 4 : 1;
  const int halo = (mode == S1) ? 2 * PW + 2 : PW + 1;
  const int hrows = (mode == S1) ? 2 : 1;
  const int Cb = std::min(op.Ci, 64), nblk = op.Ci / Cb;
  if (PW * ((mode == S2_FWD) ? 2 : 1) > 256) return false;  // TMA box extent (traversal) per dimension
  // joint choice of the output-channel chunk NT (weights staged per CTA), CTAs per SM and tile height:
  // minimise (waves of tiles per CTA) x (bytes staged + accumulator rows drained + fixed cost).  A TMA box
  // region holds exactly its (rows x PW) positions; the MMA rows of the last M-tile may read past it (into
  // the next region / the weights / a tail pad) - those rows are discarded in the epilogue.
  double best = 1e300;
  int bestTH = 0, best_mt = 0, best_stage = 0, best_cps = 1, NT = 0;
  size_t best_over = 0;
  const int CoE = narrow ? 16 : op.Co;  // MMA columns
  // DG_WS_FORCE="NT:cps:TH" (0 = free) restricts the search (tools/ws_sweep.py); read on every call
  int f_nt = 0, f_cps = 0, f_th = 0;
  if (const char* f = getenv("DG_WS_FORCE")) sscanf(f, "%d:%d:%d", &f_nt, &f_cps, &f_th);
  for (int cand : {256, 128, 64, 32, 16}) {
    if (cand > CoE || CoE % cand) continue;
    if (ncls * cand > 256) continue;
    if (f_nt && cand != f_nt) continue;
    const size_t wbytes = (size_t)9 * op.Ci * cand * 2;
    if (wbytes > 112 * 1024) continue;
    const int n_chunks = CoE / cand;
    for (int cps = 2; cps >= 1; --cps) {
      if (f_cps && cps != f_cps) continue;
      const size_t budget = ((cps == 2) ? (size_t)(113 * 1024 - 2048) : (size_t)WS_MAX_SMEM) - 1024;
      const int acc_max = (cps == 2) ? 128 : 256;
      for (int TH = 1; TH <= Ht; ++TH) {
        if (f_th && TH != f_th) { if (TH > f_th) break; continue; }
        const int span = TH * PW - (PW - Wt);
        const int n_mt = (span + 127) / 128;
        if (n_mt > 8 || n_mt * ncls * cand > acc_max) break;
        if ((TH + hrows) * ((mode == S2_FWD) ? 2 : 1) > 256) break;  // TMA box extent (traversal) per dimension
        const size_t box_pos = (size_t)(TH + hrows) * PW;
        const size_t reg_b = ((box_pos * Cb * 2) + 1023) & ~(size_t)1023;
        const size_t tile_b = (size_t)nsub * nblk * reg_b;
        const size_t need_pos = (size_t)n_mt * 128 + halo;
        const size_t over = need_pos * Cb * 2 > reg_b ? need_pos * Cb * 2 - reg_b : 0;
        if (wbytes + over + 2 * tile_b > budget) break;
        const int stages = (int)std::min<size_t>(WS_MAX_STAGE, (budget - wbytes - over) / tile_b);
        const int ntiles = (Ht + TH - 1) / TH;
        const long long tiles_total = (long long)ntiles * op.B;
        const int G = std::max(1, (148 * cps) / n_chunks);
        const double waves = (double)((tiles_total + G - 1) / G);
        const double in_b = (double)(TH + hrows) * PW * nsub * op.Ci * 2.0;
        const double out_b = (double)n_mt * 128 * ncls * cand * 2.0;
        static const double cps2 = getenv("DG_WS_CPS2") ? atof(getenv("DG_WS_CPS2")) : 0.7;  // measured (bench sweep)
        static const double fixed = getenv("DG_WS_FIXED") ? atof(getenv("DG_WS_FIXED")) : 6144.0;
        const double cost = waves * cps * (in_b + out_b + fixed) * (cps == 2 ? cps2 : 1.0) * (stages >= 3 ? 1.0 : 1.15);
        if (cost < best - 1e-9) {
          best = cost; bestTH = TH; best_mt = n_mt; best_stage = stages; best_cps = cps; NT = cand; best_over = over;
        }
      }
    }
  }
  if (bestTH == 0) return false;
  a.op = op;
  a.mode = mode; a.Ht = Ht; a.Wt = Wt;
  a.TH = bestTH; a.PW = PW; a.n_mt = best_mt; a.NT = NT; a.nplanes = nplanes; a.nsub = nsub;
  a.CoP = round_up(op.Co, 16);
  a.RB = (int)((((size_t)(bestTH + hrows) * PW * Cb * 2) + 1023) & ~(size_t)1023);
  a.over = (unsigned)best_over;
  a.nblk = nblk; a.Cb = Cb;
  a.tiles_per_img = (Ht + bestTH - 1) / bestTH;
  a.acc_cols = best_mt * ncls * NT;
  int pc = 32;
  while (pc < 2 * a.acc_cols) pc <<= 1;
  a.tmem_cols = pc;
  a.a_bytes = (unsigned)(nsub * nblk * a.RB);
  a.nstage = best_stage;
  a.w_off = a.nstage * a.a_bytes;
  a.tiles_total = a.tiles_per_img * op.B;
  a.magic_np = (unsigned)((0x100000000ULL + nplanes - 1) / nplanes);
  a.magic_pw = (unsigned)((0x100000000ULL + PW - 1) / PW);
  a.perm = perm;
  a.trace = nullptr;
  a.early = g_tune[17] ? 1 : 0;
  a.tile_tx = (unsigned)(nsub * PW * (bestTH + hrows) * op.Ci * 2);
  a.Fsh = perm ? op.Co / 4 : 1;
  ctas_per_sm = best_cps;
  return true;
}

}  // namespace

// NHWC activation view as a rank-4 bf16 tensor {C, W, H, B}; one box = Cb channels x bw x bh pixels of
// one sample.  Encoded maps are cached (the activation buffers of a model are fixed allocations).
struct MapKey {
  const void* base; int Ci, W, H, B, pitch, bw, bh, es;
  bool operator==(const MapKey& o) const {
    return base == o.base && Ci == o.Ci && W == o.W && H == o.H && B == o.B && pitch == o.pitch && bw == o.bw && bh == o.bh && es == o.es;
  }
};
static std::vector<std::pair<MapKey, CUtensorMap>> g_maps;
static std::mutex g_maps_mu;  // two handles may be driven from two host threads (one handle per thread)

static int get_map(const ConvOp& op, const WsArgs& a, CUtensorMap* out) {
  const int es = (a.mode == S2_FWD) ? 2 : 1;
  const int trows = a.TH + ((a.mode == S1) ? 2 : 1);
  MapKey k{(const bf16*)op.x.p + op.x.coff, op.Ci, op.Win, op.Hin, op.B, op.x.pitch, a.PW * es, trows * es, es};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  for (auto& e : g_maps)
    if (e.first == k) { *out = e.second; return 0; }
  cuuint64_t dims[4] = {(cuuint64_t)op.Ci, (cuuint64_t)op.Win, (cuuint64_t)op.Hin, (cuuint64_t)op.B};
  cuuint64_t strides[3] = {(cuuint64_t)op.x.pitch * 2, (cuuint64_t)op.Win * op.x.pitch * 2, (cuuint64_t)op.Hin * op.Win * op.x.pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)a.Cb, (cuuint32_t)k.bw, (cuuint32_t)k.bh, 1};
  const CUtensorMapSwizzle swz = a.Cb == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (a.Cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  CUtensorMap m;
  const CUresult r = encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(k.base), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): C %d W %d H %d B %d pitch %d box %d x %d es %d", (int)r, op.Ci, op.Win, op.Hin, op.B,
              op.x.pitch, k.bw, k.bh, es);
    return DG_ERR_CUDA;
  }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace_back(k, m);
  *out = m;
  return 0;
}

bool umma_ws_supported(const ConvOp& op) {
  WsArgs a;
  int cps;
  return plan_ws(op, a, cps);
}

int conv_umma_ws(const ConvOp& op, cudaStream_t st) {
  if (ablate(8)) return 0;
  WsArgs a;
  int cps = 1;
  if (!plan_ws(op, a, cps)) { set_error("conv_umma_ws: unsupported shape"); return DG_ERR_INVALID; }
  a.bits_tile = g_tune[21] >= 2;
  const size_t smem = (size_t)a.w_off + (size_t)9 * op.Ci * a.NT * 2 + a.over + 1024;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const double taps = op.transposed ? 2.25 : 9.0;
  Prof prof(PC_CONV_UMMA, 2.0 * total * op.Co * op.Ci * taps,
            (double)total * op.Co * (op.y.bf ? 2 : 4) + (double)op.B * op.Hin * op.Win * op.Ci * 2.0, st);
  const int n_chunks = std::max(1, op.Co / a.NT);
  int gx = std::max(1, (148 * cps) / n_chunks);
  gx = std::min(gx, a.tiles_total);
  // even out the tail: every CTA walks the same number of tiles (or one fewer)
  const int per = (a.tiles_total + gx - 1) / gx;
  gx = (a.tiles_total + per - 1) / per;
  const dim3 grid(gx, n_chunks);
  static const bool show_plan = getenv("DG_WS_PLAN") != nullptr;
  if (show_plan)
    fprintf(stderr, "[ws plan] mode %d Ci %d Co %d H %d B %d shuffle %d: TH %d n_mt %d NT %d stages %d cps %d grid %d x %d tiles %d smem %zu\n",
            a.mode, op.Ci, op.Co, op.Hout, op.B, op.shuffle, a.TH, a.n_mt, a.NT, a.nstage, cps, gx, n_chunks, a.tiles_total, smem);
  CUtensorMap tmap;
  DG_TRY(get_map(op, a, &tmap));
  static const bool tracing = getenv("DG_WS_TRACE") != nullptr;
  const size_t trace_n = (size_t)gx * 3 * 16 * 4;
  if (tracing) {
    cudaMalloc(&a.trace, trace_n * 8);
    cudaMemset(a.trace, 0, trace_n * 8);
  }
  const int kcs = a.nplanes >> 1;
#define WS_LAUNCH(M, K, S)                                                                                        \
  do {                                                                                                            \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      DG_CUDA(cudaFuncSetAttribute(conv_ws_kernel<M, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_MAX_SMEM)); \
      DG_CUDA(cudaFuncSetAttribute(conv_ws_kernel<M, K, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = grid; cfg.blockDim = dim3(WS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;            \
    cudaLaunchAttribute attr[1];                                                                                  \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                              \
    attr[0].val.programmaticStreamSerializationAllowed = g_tune[5] ? 1 : 0;                                       \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                           \
    DG_CUDA(cudaLaunchKernelEx(&cfg, conv_ws_kernel<M, K, S>, tmap, a));                                                      \
  } while (0)
#define WS_LAUNCH_KS(M, S)                    \
  do {                                        \
    if (kcs == 1) WS_LAUNCH(M, 1, S);         \
    else if (kcs == 2) WS_LAUNCH(M, 2, S);    \
    else if (kcs == 4) WS_LAUNCH(M, 4, S);    \
    else WS_LAUNCH(M, 8, S);                  \
  } while (0)
#define WS_LAUNCH_K(M)                        \
  do {                                        \
    if (simple) WS_LAUNCH_KS(M, true);        \
    else WS_LAUNCH_KS(M, false);              \
  } while (0)
  // SIMPLE epilogue (see the kernel): dg_set_tuning(22, 0) forces the general one
  const bool simple = g_tune[22] && !op.bias && op.s_acc == 1.f && !op.r1.p && !op.r2.p && op.Co >= 16 && op.Co % 16 == 0 &&
                      op.shuffle == SHUF_NONE && op.y.bf && !tracing &&
                      (op.act == ACT_NONE || op.act == ACT_LRELU || (op.act == ACT_MASK && (op.bits_in != nullptr || op.mask.bf)));
  if (a.mode == S1) WS_LAUNCH_K(S1);
  else if (a.mode == S2_FWD) WS_LAUNCH_K(S2_FWD);
  else WS_LAUNCH_K(S2_DGRAD);
#undef WS_LAUNCH_K
#undef WS_LAUNCH_KS
#undef WS_LAUNCH
  if (tracing) {
    std::vector<unsigned long long> h(trace_n);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), a.trace, trace_n * 8, cudaMemcpyDeviceToHost);
    cudaFree(a.trace);
    unsigned long long t0 = ~0ull;
    for (auto v : h) if (v && v < t0) t0 = v;
    fprintf(stderr, "[ws trace] mode %d Ci %d Co %d H %d B %d: TH %d n_mt %d NT %d stages %d cps %d grid %d x %d tiles %d smem %zu\n", a.mode,
            op.Ci, op.Co, op.Hout, op.B, a.TH, a.n_mt, a.NT, a.nstage, cps, gx, n_chunks, a.tiles_total, smem);
    for (int b : {0, gx / 2, gx - 1}) {
      static const char* names[3] = {"load [start, slot free, issued]", "mma  [start, acc free, tile full, issued]", "epi  [start, acc full, done]"};
      for (int r = 0; r < 3; ++r) {
        fprintf(stderr, "  cta %d %s\n   ", b, names[r]);
        for (int it = 0; it < 16; ++it) {
          const unsigned long long* e = &h[(((size_t)b * 3 + r) * 16 + it) * 4];
          if (!e[0]) break;
          fprintf(stderr, " |%d:", it);
          for (int k = 0; k < 4; ++k) if (e[k]) fprintf(stderr, " %.2f", (double)(e[k] - t0) * 1e-3);
        }
        fprintf(stderr, "\n");
      }
    }
  }
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
