// TMA-fed, warp-specialised tcgen05 weight gradient of a 3x3 convolution (stride 1 and 2), sm_100a.
//
//   dW[tap][ci][co] = sum over output positions p of  x[p (+) tap][ci] * dy[p][co]        (K = positions)
//
// Both operands are the NHWC tiles themselves, written by TMA with SWIZZLE_32/64/128B (row = the 16/32/64
// channels of one position) and read as MN-MAJOR tcgen05 operands (channels contiguous, K = consecutive
// positions = consecutive rows).  The swizzle XOR acts on absolute shared-memory address bits
// (tools/probes/swizzle_shift_probe.cu), so
//   * a tap (ky,kx) is a start-address shift of ky*PW + kx rows of the halo-padded x tile, and
//   * the M-atoms of the A descriptor may OVERLAP: with LBO = one row, atom g is the tile shifted by g
//     positions, i.e. the three kx taps of a kernel row are folded into M  (M = 3*Ci; used when 3*Ci <= 128)
//     (tools/probes/wgrad_desc_probe.cu).  Ci = 16 uses M = 64 MMAs (accumulator row i lives in TMEM lane
//     (i % 16) + 32 * (i / 16)), which halves the shared-memory operand traffic that bounds these small-N MMAs.
// Stride 2 stages x as four parity sub-images (TMA element strides) so that every tap is again a constant shift.
// Conv padding, the pad column of the dy tile and rows outside the image are TMA out-of-bounds zero fill; the
// ring is zeroed once so that K-steps running past a tile only ever see zeros (dy) or finite values (x).
//
//   warp 0  producer: TMA loads of (x tile, dy tile) of tile i+1.. into an NSTAGE ring   (full[s])
//   warp 1  MMA: one elected thread issues  ksteps x units  tcgen05.mma per tile into per-unit TMEM
//           accumulators that persist over all tiles of the CTA; commit -> empty[s]
//   all     epilogue once per CTA: tcgen05.ld, red.global.add.v4.f32 into dW (fp32, [tap][ci][CoP])
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int WW_THREADS = 128;
constexpr int WW_MAX_STAGE = 4;
constexpr int WW_MAX_SMEM = 227 * 1024 - 3072;
enum WMode { W_S1_FOLD = 0, W_S1_TAPS = 1, W_S2_TAPS = 2, W_S2_FOLD = 3 };
// stride 2, Ci <= 32: the taps that are consecutive positions of the SAME parity sub-image are folded into M
// (two atoms): 6 MMAs per K-step instead of 9.  unit -> {sub-image, row shift (x PW), col shift, taps of atom 0 / 1}
__constant__ int c_s2f_sub[6] = {3, 3, 1, 2, 2, 0};
__constant__ int c_s2f_dy[6] = {0, 1, 1, 0, 1, 1};
__constant__ int c_s2f_dx[6] = {0, 0, 0, 1, 1, 1};
__constant__ int c_s2f_tap0[6] = {0, 6, 3, 1, 7, 4};
__constant__ int c_s2f_tap1[6] = {2, 8, 5, -1, -1, -1};

struct WwArgs {
  WgradOp op;
  int mode, units;  // MMAs per K-step: 3 (kernel rows, kx folded into M) or 9 (taps)
  int M;            // 64 or 128
  int TH, PW, nks;  // dy rows per tile, padded width, K-steps (16 positions) per tile
  int NT, CoP;      // output-channel chunk per CTA (blockIdx.y)
  int xrow, drow;   // bytes of one position row of the x / dy tile (32, 64 or 128)
  int nblkx;        // 64-channel blocks of x (2 when Ci = 128)
  int nsub;         // 4 parity sub-images when stride 2
  unsigned xreg, dreg;      // bytes of one x box region / of the dy region (1024-aligned)
  unsigned stage_bytes, d_off;  // ring slot size, offset of the dy region inside a slot
  unsigned tx_bytes;
  int tiles_per_img, tiles_total, tiles_per_cta, nstage, tmem_cols;
  int ci_total, ci_off;  // dw row = tap * ci_total + ci_off + ci
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_ww(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ uint32_t elect_one_sync_w() {
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
// MN-major swizzled operand: LBO = byte distance between M/N atoms, SBO = 8 rows, layout = swizzle code
__device__ __forceinline__ uint64_t mn_desc(uint32_t lbo, uint32_t rowbytes) {
  const uint64_t layout = rowbytes == 32 ? 6ull : (rowbytes == 64 ? 4ull : 2ull);
  return smem_desc(0, lbo, 8 * rowbytes) | (layout << 61);
}

template <int MODE>
__device__ __forceinline__ void wgrad_ws_body(const CUtensorMap* mapx_p, const CUtensorMap* mapd_p, const WwArgs& a, uint8_t* smem,
                                              uint64_t* bars, uint32_t* tmem_slot_p) {
  uint32_t& tmem_slot = *tmem_slot_p;
  constexpr int UNITS = (MODE == W_S1_FOLD) ? 3 : (MODE == W_S2_FOLD ? 6 : 9);
  constexpr bool STRIDE2 = (MODE == W_S2_TAPS || MODE == W_S2_FOLD);
  const WgradOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * a.NT;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = min(a.tiles_total, t_begin + a.tiles_per_cta);
  const int my_tiles = max(0, t_end - t_begin);
  if (my_tiles == 0 || co0 >= op.Co) return;  // batched launches are sized for the largest op (uniform exit, before any barrier)
  const int S = a.nstage;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WW_MAX_STAGE + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * WW_MAX_STAGE);

  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < WW_MAX_STAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
  }
  const uint32_t sa0 = (smem_u32(smem) + 1023u) & ~1023u;
  {  // zero the ring once: K-steps that run past a tile must contract zeros (dy) with finite values (x)
    const uint32_t zbytes = (uint32_t)S * a.stage_bytes;
    for (uint32_t i = tid * 16; i < zbytes; i += WW_THREADS * 16)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sa0 + i), "r"(0u) : "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: the prologue above touched no global memory
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = t_begin + it;
      const int s = it % S;
      mbar_wait_ww(empty_bar(s), (((uint32_t)(it / S)) & 1u) ^ 1u);
      const uint32_t sx = sa0 + s * a.stage_bytes, sd = sx + a.d_off;
      const int n = tile / a.tiles_per_img;
      const int y0 = (tile - n * a.tiles_per_img) * a.TH;
      if (elect_one_sync_w()) {
        mbar_expect_tx(full_bar(s), a.tx_bytes);
#pragma unroll
        for (int sub = 0; sub < (STRIDE2 ? 4 : 1); ++sub) {
          int cx, cy;
          if (STRIDE2) { cx = 2 * op.col0 - 2 + (sub & 1); cy = 2 * (y0 - 1) + (sub >> 1); }
          else { cx = op.col0 - 1; cy = y0 - 1; }
          for (int blk = 0; blk < a.nblkx; ++blk)
            tma_load_4d(sx + (sub * a.nblkx + blk) * a.xreg, mapx_p, blk * 64, cx, cy, n, full_bar(s));
        }
        tma_load_4d(sd, mapd_p, co0, 0, y0, n, full_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issue =================
    const uint32_t idesc = instr_desc(a.M, a.NT, 1, 1);
    const uint32_t xr16 = (uint32_t)a.xrow >> 4, dr16 = (uint32_t)a.drow >> 4;
    // A atoms: folded kx taps are one position apart; Ci = 128 is two 64-channel blocks one region apart
    const uint32_t lboA = (a.nblkx > 1) ? a.xreg : (uint32_t)a.xrow;
    const uint64_t adesc0 = mn_desc(lboA, (uint32_t)a.xrow), bdesc0 = mn_desc(16, (uint32_t)a.drow);
    const uint32_t subA = (a.xreg >> 4) * a.nblkx;
    uint32_t uoff[UNITS];  // per-unit start offset inside the x slot, in 16-byte units
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
      if (MODE == W_S1_FOLD) {
        uoff[u] = (uint32_t)(u * a.PW) * xr16;
      } else if (MODE == W_S2_FOLD) {
        uoff[u] = c_s2f_sub[u] * subA + (uint32_t)(c_s2f_dy[u] * a.PW + c_s2f_dx[u]) * xr16;
      } else {
        const int ky = u / 3, kx = u % 3;
        if (MODE == W_S1_TAPS) {
          uoff[u] = (uint32_t)(ky * a.PW + kx) * xr16;
        } else {
          const int sub = ((ky == 1) ? 0 : 2) + ((kx == 1) ? 0 : 1);
          uoff[u] = sub * subA + (uint32_t)(((ky == 0) ? 0 : 1) * a.PW + ((kx == 0) ? 0 : 1)) * xr16;
        }
      }
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % S;
      mbar_wait_ww(full_bar(s), ((uint32_t)(it / S)) & 1u);
      tc_fence_after();
      const uint32_t sx16 = (sa0 + s * a.stage_bytes) >> 4, sd16 = sx16 + (a.d_off >> 4);
      if (elect_one_sync_w()) {
        for (int ks = 0; ks < a.nks; ++ks) {
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          const uint64_t bd = bdesc0 + (uint64_t)(sd16 + ks * 16 * dr16);
          const uint32_t xa = sx16 + ks * 16 * xr16;
#pragma unroll
          for (int u = 0; u < UNITS; ++u) umma_f16(tmem + u * a.NT, adesc0 + (uint64_t)(xa + uoff[u]), bd, idesc, acc);
        }
        umma_commit(empty_bar(s));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  }
  // ================= epilogue (all warps; warp w reads TMEM lane quarter w) =================
  if (my_tiles > 0) {
    mbar_wait_ww(done_bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int row;  // accumulator row held by this thread's TMEM lane
    if (a.M == 128) row = warp * 32 + lane;
    else row = (lane < 16) ? warp * 16 + lane : -1;
    int kx = 0, ci = row;
    bool valid = row >= 0;
    if (MODE == W_S1_FOLD || MODE == W_S2_FOLD) {
      kx = row / op.Ci; ci = row - kx * op.Ci;
      valid = valid && kx < ((MODE == W_S1_FOLD) ? 3 : 2);
    } else {
      valid = valid && row < op.Ci;
    }
    const bool valid_row = valid;
    for (int u = 0; u < UNITS; ++u) {
      int tap;
      if (MODE == W_S1_FOLD) tap = u * 3 + kx;
      else if (MODE == W_S2_FOLD) { tap = (kx == 0) ? c_s2f_tap0[u] : c_s2f_tap1[u]; valid = valid_row && tap >= 0; if (tap < 0) tap = 0; }
      else tap = u;
      for (int nc = 0; nc < a.NT; nc += 16) {
        float v[16];
        tmem_ld16(tmem + lane_base + u * a.NT + nc, v);
        if (valid) {
          float* dst = op.dw + ((size_t)tap * a.ci_total + a.ci_off + ci) * a.CoP + co0 + nc;
#pragma unroll
          for (int q = 0; q < 4; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

template <int MODE>
__global__ void __launch_bounds__(WW_THREADS) wgrad_ws_kernel(const __grid_constant__ CUtensorMap mapx,
                                                               const __grid_constant__ CUtensorMap mapd, const WwArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * WW_MAX_STAGE + 1];
  __shared__ uint32_t tmem_slot;
  wgrad_ws_body<MODE>(&mapx, &mapd, a, smem, bars, &tmem_slot);
}

// one launch for many ops of the same MODE: blockIdx.z selects the op; plans and tensor maps live in a device table
template <int MODE>
__global__ void __launch_bounds__(WW_THREADS) wgrad_ws_batched_kernel(const WwArgs* __restrict__ args, const CUtensorMap* __restrict__ maps,
                                                                       const int* __restrict__ op_index) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * WW_MAX_STAGE + 1];
  __shared__ uint32_t tmem_slot;
  __shared__ WwArgs sa;
  const int z = op_index[blockIdx.z];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(args + z);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (int i = threadIdx.x; i < (int)(sizeof(WwArgs) / 4); i += WW_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  wgrad_ws_body<MODE>(maps + 2 * z, maps + 2 * z + 1, sa, smem, bars, &tmem_slot);
}

bool plan_ww(const WgradOp& op, WwArgs& a) {
  if (!op.x.bf || !op.dy.bf || !op.dw) return false;
  if (op.Ci != 16 && op.Ci != 32 && op.Ci != 64 && op.Ci != 128) return false;
  if (op.Co % 16 || op.Co < 16 || op.Co > 256) return false;
  if (op.x.pitch % 8 || op.x.coff % 8 || op.dy.pitch % 8 || op.dy.coff % 8) return false;
  const int s = op.stride;
  if (s == 1) { if (op.Hin != op.Hout || op.Win != op.Wout) return false; }
  else if (s == 2) { if (op.Hin != 2 * op.Hout || op.Win != 2 * op.Wout) return false; }
  else return false;
  const int mode = (s == 2) ? (op.Ci <= 32 ? W_S2_FOLD : W_S2_TAPS) : (3 * op.Ci <= 128 ? W_S1_FOLD : W_S1_TAPS);
  const int units = (mode == W_S1_FOLD) ? 3 : (mode == W_S2_FOLD ? 6 : 9);
  int NT = 0;
  // TMEM columns of a CTA (DG_WW_TMEM, default 512): <= 256 lets a second tcgen05 CTA share the SM while this one runs
  static const int tmem_limit = getenv("DG_WW_TMEM") ? atoi(getenv("DG_WW_TMEM")) : 512;
  for (int cand : {64, 32, 16})
    if (op.Co % cand == 0 && units * cand <= tmem_limit) { NT = cand; break; }
  if (NT == 0) return false;
  const int PW = (s == 1) ? op.Wout + 2 : op.Wout + 1;
  if (PW * s > 256) return false;  // TMA box extent (traversal) per dimension
  const int xrow = std::min(op.Ci, 64) * 2, drow = NT * 2;
  const int nblkx = op.Ci > 64 ? 2 : 1, nsub = (s == 2) ? 4 : 1;
  const int hrows = (s == 1) ? 2 : 1;
  const int max_shift = (s == 1) ? 2 * PW + 2 : PW + 1;
  // tile height: the largest whose ring (>= 3 slots) fits, capped so that one slot stays a modest unit of work
  int bestTH = 0;
  WwArgs b{};
  for (int TH = 1; TH <= op.Hout; ++TH) {
    if ((TH + hrows) * s > 256) break;
    const int nks = (TH * PW + 15) / 16;
    const size_t xpos = std::max((size_t)(TH + hrows) * PW, (size_t)nks * 16 + max_shift + 8);
    const size_t xreg = (xpos * xrow + 1023) & ~(size_t)1023;
    const size_t dreg = ((size_t)nks * 16 * drow + 1023) & ~(size_t)1023;
    const size_t stage = xreg * nblkx * nsub + dreg;
    if (TH > 1 && stage > 48 * 1024) break;
    if (3 * stage + 1024 > (size_t)WW_MAX_SMEM) break;
    bestTH = TH;
    b.TH = TH; b.nks = nks; b.xreg = (unsigned)xreg; b.dreg = (unsigned)dreg; b.stage_bytes = (unsigned)stage;
    b.d_off = (unsigned)(xreg * nblkx * nsub);
    b.tx_bytes = (unsigned)((size_t)nsub * (TH + hrows) * PW * op.Ci * 2 + (size_t)TH * PW * drow);
  }
  if (bestTH == 0) return false;
  if (bestTH < op.Hout) {  // balance the tiles of an image (e.g. 16 rows: 8 + 8 instead of 14 + 2)
    const int nt = (op.Hout + bestTH - 1) / bestTH;
    const int TH = (op.Hout + nt - 1) / nt;
    if (TH < bestTH) {
      const int nks = (TH * PW + 15) / 16;
      const size_t xpos = std::max((size_t)(TH + hrows) * PW, (size_t)nks * 16 + max_shift + 8);
      const size_t xreg = (xpos * xrow + 1023) & ~(size_t)1023;
      const size_t dreg = ((size_t)nks * 16 * drow + 1023) & ~(size_t)1023;
      bestTH = TH;
      b.TH = TH; b.nks = nks; b.xreg = (unsigned)xreg; b.dreg = (unsigned)dreg;
      b.stage_bytes = (unsigned)(xreg * nblkx * nsub + dreg);
      b.d_off = (unsigned)(xreg * nblkx * nsub);
      b.tx_bytes = (unsigned)((size_t)nsub * (TH + hrows) * PW * op.Ci * 2 + (size_t)TH * PW * drow);
    }
  }
  a = b;
  a.op = op;
  a.mode = mode; a.units = units;
  a.M = (mode == W_S1_FOLD) ? (3 * op.Ci <= 64 ? 64 : 128) : (mode == W_S2_FOLD ? 64 : (op.Ci <= 64 ? 64 : 128));
  a.PW = PW; a.NT = NT; a.CoP = round_up(op.Co, 16);
  a.xrow = xrow; a.drow = drow; a.nblkx = nblkx; a.nsub = nsub;
  a.tiles_per_img = (op.Hout + a.TH - 1) / a.TH;
  a.tiles_total = a.tiles_per_img * op.B;
  a.nstage = (int)std::min<size_t>(WW_MAX_STAGE, ((size_t)WW_MAX_SMEM - 1024) / a.stage_bytes);
  // two CTAs per SM when their rings fit side by side (more loads in flight, two MMA issue streams)
  if ((size_t)a.nstage * a.stage_bytes + 1024 > 112 * 1024) {
    if (3 * (size_t)a.stage_bytes + 1024 <= 112 * 1024) a.nstage = 3;  // (two slots per CTA measured slower)
  }
  int pc = 32;
  while (pc < units * NT) pc <<= 1;
  a.tmem_cols = pc;
  a.ci_total = op.dw_ci_total > 0 ? op.dw_ci_total : op.Ci;
  a.ci_off = op.dw_ci_total > 0 ? op.dw_ci_off : 0;
  return true;
}

struct MapKeyW {
  const void* base; int C, W, H, B, pitch, bc, bw, bh, es, Wrow;
  bool operator==(const MapKeyW& o) const {
    return base == o.base && C == o.C && W == o.W && H == o.H && B == o.B && pitch == o.pitch && bc == o.bc && bw == o.bw && bh == o.bh &&
           es == o.es && Wrow == o.Wrow;
  }
};
std::vector<std::pair<MapKeyW, CUtensorMap>> g_maps_w;
std::mutex g_maps_w_mu;

// NHWC view {C, W, H, B} (bf16); box = bc channels x bw x bh (traversal extents) with element stride es on W and H
// Wrow > W: the view is a column strip [col0, col0 + W) of rows that are Wrow positions long (t already points at col0);
// columns past the strip read as zero (out-of-bounds fill), rows keep the full pitch
int get_map_w(const TV& t, int C, int W, int H, int B, int bc, int bw, int bh, int es, CUtensorMap* out, int Wrow = 0) {
  if (Wrow <= 0) Wrow = W;
  MapKeyW k{(const bf16*)t.p + t.coff, C, W, H, B, t.pitch, bc, bw, bh, es, Wrow};
  std::lock_guard<std::mutex> lock(g_maps_w_mu);
  for (auto& e : g_maps_w)
    if (e.first == k) { *out = e.second; return 0; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)t.pitch * 2, (cuuint64_t)Wrow * t.pitch * 2, (cuuint64_t)H * Wrow * t.pitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  const CUtensorMapSwizzle swz = bc == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap m;
  const CUresult r = encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(k.base), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): C %d W %d H %d B %d pitch %d box %d x %d x %d es %d", (int)r, C, W, H, B, t.pitch, bc, bw, bh, es);
    return DG_ERR_CUDA;
  }
  if (g_maps_w.size() > 4096) g_maps_w.clear();
  g_maps_w.emplace_back(k, m);
  *out = m;
  return 0;
}

}  // namespace

bool wgrad_ws_supported(const WgradOp& op) {
  WwArgs a;
  return plan_ww(op, a);
}

int wgrad_ws(const WgradOp& op, cudaStream_t st) {
  if (ablate(0)) return 0;
  WwArgs a;
  if (!plan_ww(op, a)) { set_error("wgrad_ws: unsupported shape"); return DG_ERR_INVALID; }
  const int s = op.stride;
  const int hrows = (s == 1) ? 2 : 1;
  CUtensorMap mx, md;
  if (op.Wout_full > 0) {
    // column strip: x keeps the whole rows (the strip's halo columns are real neighbours, only the image border is zero
    // fill; the kernel starts its boxes at column col0 * stride - 1), dy is the strip alone
    TV dys = op.dy;
    dys.coff += op.col0 * op.dy.pitch;
    DG_TRY(get_map_w(op.x, op.Ci, op.Wout_full * s, op.Hin, op.B, std::min(op.Ci, 64), a.PW * s, (a.TH + hrows) * s, s, &mx));
    DG_TRY(get_map_w(dys, op.Co, op.Wout, op.Hout, op.B, a.NT, a.PW, a.TH, 1, &md, op.Wout_full));
  } else {
    DG_TRY(get_map_w(op.x, op.Ci, op.Win, op.Hin, op.B, std::min(op.Ci, 64), a.PW * s, (a.TH + hrows) * s, s, &mx));
    DG_TRY(get_map_w(op.dy, op.Co, op.Wout, op.Hout, op.B, a.NT, a.PW, a.TH, 1, &md));
  }
  const int n_chunks = op.Co / a.NT;
  const size_t smem = (size_t)a.nstage * a.stage_bytes + 1024;
  const int cps = (2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
  // position split: enough CTAs to fill the chip, bounded so that the cross-CTA fp32 reductions stay small
  long long S = (148 * cps + n_chunks - 1) / n_chunks;
  static const long long cap_elems = getenv("DG_WW_CAP") ? atoll(getenv("DG_WW_CAP")) : 6000000LL;  // (bench sweep: 3 M -> 6 M = -5 % on the class, flat beyond)
  const long long cap = cap_elems / (9LL * op.Ci * op.Co) + 1;
  if (S > cap) S = cap;
  if (S > a.tiles_total) S = a.tiles_total;
  if (S < 1) S = 1;
  a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
  S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const dim3 grid((unsigned)S, (unsigned)n_chunks);
  {
  Prof prof(PC_WGRAD_UMMA, 2.0 * total * op.Co * op.Ci * 9.0,
            (double)total * op.Co * 2.0 + (double)op.B * op.Hin * op.Win * op.Ci * 2.0, st);
#define WW_LAUNCH(MODE)                                                                                                \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      DG_CUDA(cudaFuncSetAttribute(wgrad_ws_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_MAX_SMEM));  \
      DG_CUDA(cudaFuncSetAttribute(wgrad_ws_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    cudaLaunchConfig_t cfg = {};                                                                                       \
    cfg.gridDim = grid; cfg.blockDim = dim3(WW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;                 \
    cudaLaunchAttribute attr[1];                                                                                       \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
    attr[0].val.programmaticStreamSerializationAllowed = g_tune[5] ? 1 : 0;                                            \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                                \
    DG_CUDA(cudaLaunchKernelEx(&cfg, wgrad_ws_kernel<MODE>, mx, md, a));                                               \
  } while (0)
  if (a.mode == W_S1_FOLD) WW_LAUNCH(W_S1_FOLD);
  else if (a.mode == W_S1_TAPS) WW_LAUNCH(W_S1_TAPS);
  else if (a.mode == W_S2_FOLD) WW_LAUNCH(W_S2_FOLD);
  else WW_LAUNCH(W_S2_TAPS);
#undef WW_LAUNCH
  }
  DG_LAUNCH_CHECK();
  if (op.dbias) DG_TRY(colsum(op.dy, (size_t)total, op.Co, op.dbias, st));
  return 0;
}


size_t wgrad_ws_batch_bytes(int n_ops) {
  // a multiple of 256 so that consecutive tables in one allocation keep their tensor maps 64-byte aligned
  return (((size_t)n_ops * (sizeof(WwArgs) + 2 * sizeof(CUtensorMap) + sizeof(int)) + 256) + 255) & ~(size_t)255;
}

// Table layout: [CUtensorMap x 2n (64-byte aligned)] [WwArgs x n] [int x n: op indices grouped by mode]
int wgrad_ws_batched(const WgradOp* ops, int n, void* table_dev, std::vector<unsigned char>& shadow, int S_per_op, cudaStream_t st) {
  if (ablate(4)) return 0;
  if (n <= 0) return 0;
  const size_t maps_bytes = (size_t)n * 2 * sizeof(CUtensorMap), args_bytes = (size_t)n * sizeof(WwArgs);
  std::vector<unsigned char> tab(maps_bytes + args_bytes + (size_t)n * sizeof(int));
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(tab.data());
  WwArgs* args = reinterpret_cast<WwArgs*>(tab.data() + maps_bytes);
  int* order = reinterpret_cast<int*>(tab.data() + maps_bytes + args_bytes);
  unsigned smem[4] = {0, 0, 0, 0}, gx[4] = {1, 1, 1, 1};
  int count[4] = {0, 0, 0, 0};
  double flops = 0, bytes = 0;
  for (int i = 0; i < n; ++i) {
    WwArgs a;
    if (!plan_ww(ops[i], a)) { set_error("wgrad_ws_batched: op %d unsupported", i); return DG_ERR_INVALID; }
    if (ops[i].Co != a.NT) { set_error("wgrad_ws_batched: op %d needs output-channel chunks", i); return DG_ERR_INVALID; }
    if (ops[i].Wout_full > 0) { set_error("wgrad_ws_batched: op %d is a column strip", i); return DG_ERR_INVALID; }
    long long S = std::max(1, std::min(S_per_op, a.tiles_total));
    a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
    S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
    const int s = ops[i].stride, hrows = (s == 1) ? 2 : 1;
    CUtensorMap mx, md;
    DG_TRY(get_map_w(ops[i].x, ops[i].Ci, ops[i].Win, ops[i].Hin, ops[i].B, std::min(ops[i].Ci, 64), a.PW * s, (a.TH + hrows) * s, s, &mx));
    DG_TRY(get_map_w(ops[i].dy, ops[i].Co, ops[i].Wout, ops[i].Hout, ops[i].B, a.NT, a.PW, a.TH, 1, &md));
    memcpy(&maps[2 * i], &mx, sizeof(CUtensorMap));
    memcpy(&maps[2 * i + 1], &md, sizeof(CUtensorMap));
    memcpy(&args[i], &a, sizeof(WwArgs));
    gx[a.mode] = std::max(gx[a.mode], (unsigned)S);
    smem[a.mode] = std::max(smem[a.mode], (unsigned)((size_t)a.nstage * a.stage_bytes + 1024));
    ++count[a.mode];
    const double total = (double)ops[i].B * ops[i].Hout * ops[i].Wout;
    flops += 2.0 * total * ops[i].Co * ops[i].Ci * 9.0;
    bytes += total * ops[i].Co * 2.0 + (double)ops[i].B * ops[i].Hin * ops[i].Win * ops[i].Ci * 2.0;
  }
  int start[4], pos = 0;
  for (int m = 0; m < 4; ++m) {
    start[m] = pos;
    for (int i = 0; i < n; ++i)
      if (args[i].mode == m) order[pos++] = i;
  }
  if (shadow.size() != tab.size() || memcmp(shadow.data(), tab.data(), tab.size()) != 0) {
    shadow = tab;
    DG_CUDA(cudaMemcpyAsync(table_dev, shadow.data(), shadow.size(), cudaMemcpyHostToDevice, st));
  }
  const CUtensorMap* dmaps = reinterpret_cast<const CUtensorMap*>(table_dev);
  const WwArgs* dargs = reinterpret_cast<const WwArgs*>((const unsigned char*)table_dev + maps_bytes);
  const int* dorder = reinterpret_cast<const int*>((const unsigned char*)table_dev + maps_bytes + args_bytes);
  Prof prof(PC_WGRAD_UMMA, flops, bytes, st);
#define WWB_LAUNCH(MODE)                                                                                                      \
  do {                                                                                                                        \
    if (count[MODE] > 0) {                                                                                                    \
      static bool attr_set = false;                                                                                           \
      if (!attr_set) {                                                                                                        \
        DG_CUDA(cudaFuncSetAttribute(wgrad_ws_batched_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_MAX_SMEM)); \
        DG_CUDA(cudaFuncSetAttribute(wgrad_ws_batched_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout,           \
                                     cudaSharedmemCarveoutMaxShared));                                                        \
        attr_set = true;                                                                                                      \
      }                                                                                                                       \
      wgrad_ws_batched_kernel<MODE><<<dim3(gx[MODE], 1, (unsigned)count[MODE]), WW_THREADS, smem[MODE], st>>>(dargs, dmaps,   \
                                                                                                              dorder + start[MODE]); \
      DG_LAUNCH_CHECK();                                                                                                      \
    }                                                                                                                         \
  } while (0)
  WWB_LAUNCH(W_S1_FOLD);
  WWB_LAUNCH(W_S1_TAPS);
  WWB_LAUNCH(W_S2_TAPS);
  WWB_LAUNCH(W_S2_FOLD);
#undef WWB_LAUNCH
  return 0;
}

}  // namespace dg
