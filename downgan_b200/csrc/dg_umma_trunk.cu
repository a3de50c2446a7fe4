// Persistent fused RRDB trunk (forward) on tcgen05, sm_100a.
//
// One CTA owns one 16x16 coarse image for the WHOLE trunk (R residual-in-residual blocks = 3R dense
// blocks = 15R convolutions, networks/generator.py:36-41,52-53 of the reference).  The dense-block
// concat buffer (x, o1..o4 = 80 channels) never leaves shared memory: it is the planar tile
// [10 planes][18x18 padded positions][8 ch] that the tcgen05 A descriptors read directly (tap =
// start-address offset), and every epilogue writes its 16 output channels (bias + LeakyReLU, or
// 0.2*o5 + x [+ RRDB skip]) straight into the next two planes / back into planes 0-1.  Only the
// per-layer weight images stream in from L2 (double-buffered cp.async, prefetched during the MMAs
// of the previous layer).  Three warps issue the MMAs of the three M-tiles in parallel; the
// accumulators (3 x 16 TMEM columns) are read back by all four warps.  With `db_bufs` set, every
// slice is also stored to the per-block global concat buffers for the backward pass.
#include <stdlib.h>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int TF = 16;            // filters
constexpr int TW = 16;            // coarse grid edge
constexpr int TPW = TW + 2;       // padded width
constexpr int TNMT = 3;           // M-tiles: 128 + 128 + 64 rows >= the 16*18 = 288 padded positions (the third tile is an M = 64 MMA)
constexpr int T_THREADS = 320;    // 10 warps: warps 0-3 drain M-tile 0, warps 4-7 M-tile 1, warps 8-9 the 32 live rows of M-tile 2
constexpr int T_COLS_MT = 5 * TF; // TMEM columns per M-tile: one 16-column accumulator per layer of the block
constexpr int T_TMEM_COLS = 256;  // 3 * 80 = 240 -> power of two
constexpr int TPBPOS = 424;       // positions per plane incl. over-read slack (2*128 + 64 + 2*18 + 2 = 358 <= 424)
constexpr int TPB = TPBPOS * 16;  // plane stride in bytes
constexpr int T_X_BYTES = 10 * TPB;
constexpr int T_W_BYTES = 9 * 80 * TF * 2;
constexpr int T_SMEM = T_X_BYTES + 2 * T_W_BYTES;
constexpr float G_SLOPE = 0.01f, RES = 0.2f;
constexpr int T_DB_ELEMS_ = 9 * TF * TF * 15;  // weight elements of one dense block

struct TrunkArgs {
  const bf16* x_in; int in_pitch, in_coff;  // (B,16,16,pitch): conv1 output
  bf16* y_out; int out_pitch;               // (B,16,16,pitch): trunk output
  bf16* const* db_bufs;                     // [3R] concat buffers (pitch 80) or nullptr
  const bf16* w;                            // slice-major B-operand images, 5 per dense block (pack_trunk_slices)
  const float* bias;                        // 16 floats per dense conv, consecutive
  int R, B;
  int order;                                // MMA issue order, see trunk_issue_round
  int save_count;                           // samples [0, save_count) store their slices to db_bufs
  unsigned long long* trace;                // debug timeline (DG_TRUNK_TRACE), null otherwise
};

__device__ __forceinline__ unsigned long long gtimer_t() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TR_TRACE(ev)                                                                                   \
  do {                                                                                                 \
    if (a.trace && blockIdx.x == 0 && tid == 0 && L >= 30 && L < 62) a.trace[(L - 30) * 8 + (ev)] = gtimer_t(); \
  } while (0)
__device__ __forceinline__ uint32_t elect_one_sync_t() {
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
__device__ __forceinline__ void st_shared16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(uint4 q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(w[k] << 16);
    v[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}


// MMA round of slice j for the three M-tiles.  order 0: warps 0..2 each issue their own M-tile (three
// independent accumulate chains interleave in the tensor pipe); 1: one thread, M-tile after M-tile (M-tile 0
// completes first); 2: one thread, tap-major (chains interleaved).  Every M-tile commits to its own mbarrier.
__device__ __forceinline__ void trunk_issue_round(int order, int warp, uint32_t tmem, uint32_t sX, uint32_t wb, int j, int Nj,
                                                  uint64_t* mbar) {
  const uint32_t id128 = instr_desc(128, Nj), id64 = instr_desc(64, Nj);
  const uint32_t wplane = Nj * 16;
  const uint64_t bd0 = smem_desc(wb, wplane, 128);
  const uint32_t accf = (j > 0) ? 1u : 0u;
  if (order == 0) {
    if (warp < TNMT) {
      if (elect_one_sync_t()) {
        const uint64_t ad0 = smem_desc(sX + 2 * j * TPB + (warp * 128) * 16, TPB, 128);
        const uint32_t idj = (warp == 2) ? id64 : id128;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
          umma_f16(tmem + warp * T_COLS_MT + j * TF, ad0 + (uint64_t)((tap / 3) * TPW + tap % 3), bd0 + (uint64_t)((tap * 2 * wplane) >> 4), idj,
                   (tap > 0) ? 1u : accf);
        umma_commit(smem_u32(&mbar[warp]));
      }
      __syncwarp();
    }
    return;
  }
  if (warp == 0) {
    if (elect_one_sync_t()) {
      if (order == 1) {
#pragma unroll
        for (int mt = 0; mt < TNMT; ++mt) {
          const uint64_t ad0 = smem_desc(sX + 2 * j * TPB + (mt * 128) * 16, TPB, 128);
          const uint32_t idj = (mt == 2) ? id64 : id128;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            umma_f16(tmem + mt * T_COLS_MT + j * TF, ad0 + (uint64_t)((tap / 3) * TPW + tap % 3), bd0 + (uint64_t)((tap * 2 * wplane) >> 4), idj,
                     (tap > 0) ? 1u : accf);
          umma_commit(smem_u32(&mbar[mt]));
        }
      } else {
        const uint64_t ad0 = smem_desc(sX + 2 * j * TPB, TPB, 128);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int mt = 0; mt < TNMT; ++mt)
            umma_f16(tmem + mt * T_COLS_MT + j * TF, ad0 + (uint64_t)(mt * 128 + (tap / 3) * TPW + tap % 3),
                     bd0 + (uint64_t)((tap * 2 * wplane) >> 4), (mt == 2) ? id64 : id128, (tap > 0) ? 1u : accf);
#pragma unroll
        for (int mt = 0; mt < TNMT; ++mt) umma_commit(smem_u32(&mbar[mt]));
      }
    }
    __syncwarp();
  }
}

// Which accumulator row a thread drains: warp group h = warp / 4 owns M-tile h.  M-tiles 0 and 1 are M = 128 MMAs (row i =
// TMEM lane i: warp w reads rows 32*(w%4) + lane).  M-tile 2 is an M = 64 MMA whose row i lives in TMEM lane 32*(i/16) + i%16
// (tools/probes/wgrad_desc_probe.cu), so lanes 0-15 of warps 8 and 9 (TMEM lane quarters 0 and 1) hold rows 0..31 = padded
// positions 256..287; its rows 32..63 lie past the image and are never read.
struct TrunkSlot {
  int pos, pix, tile;
  bool valid;
  uint32_t tcol;   // TMEM address (lane quarter | column of layer 0) of the thread's accumulator row
};
__device__ __forceinline__ TrunkSlot trunk_slot(int warp, int lane) {
  TrunkSlot s;
  const int qd = warp & 3, h = warp >> 2;
  s.tile = h;
  const int q = (h < 2) ? h * 128 + qd * 32 + lane : 256 + 16 * qd + lane;
  const int y = q / TPW, x = q - y * TPW;
  s.valid = (x < TW) && (y < TW) && (h < 2 || lane < 16);
  s.pos = q + TPW + 1;
  s.pix = s.valid ? y * TW + x : 0;
  s.tcol = ((uint32_t)(qd * 32) << 16) + h * T_COLS_MT;
  return s;
}

__global__ void __launch_bounds__(T_THREADS, 2) trunk_fwd_kernel(const TrunkArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[TNMT];  // one per M-tile: its epilogue starts while the other tiles' MMAs run
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x;
  const uint32_t sX = smem_u32(smem);
  const uint32_t sW = sX + T_X_BYTES;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), T_TMEM_COLS);
  if (tid == 32) { for (int i = 0; i < TNMT; ++i) mbar_init(smem_u32(&mbar[i]), 1); }
  // zero the whole concat tile: halo ring and pad columns are the convolution's zero padding
  for (int i = tid; i < T_X_BYTES / 16; i += T_THREADS) st_shared16(sX + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  // image -> planes 0,1 (interior positions), weights of layer 0 -> buffer 0
  for (int i = tid; i < 256 * 2; i += T_THREADS) {
    const int pix = i >> 1, pl = i & 1;
    const int y = pix >> 4, x = pix & 15;
    cp_async16(sX + pl * TPB + ((y + 1) * TPW + x + 1) * 16,
               a.x_in + ((size_t)n * 256 + pix) * a.in_pitch + a.in_coff + pl * 8, 16);
  }
  for (int i = tid; i < 18 * 5 * TF; i += T_THREADS) cp_async16(sW + i * 16, reinterpret_cast<const uint4*>(a.w) + i, 16);
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const TrunkSlot S = trunk_slot(warp, lane);

  uint4 xr[2];  // RRDB input at this thread's position (bf16 x 16)
  const int total = a.R * 15;
  size_t w_elem = 0;  // element offset of the current slice's weight image
  // Input-stationary schedule: as soon as slice j of the concat buffer (x, o1..o4) exists, ONE round
  // of MMAs adds its contribution to ALL later layers j+1..5 (N = 16*(5-j) output channels, weight
  // image = slice j of W_{j+1..5} side by side).  Every A block is then read from shared memory once
  // per dense block instead of once per consuming layer (3x less operand traffic, the actual bound
  // of an N=16 MMA), and layer j+1's accumulator is complete after round j.
  for (int L = 0; L < total; ++L) {
    const int db = L / 5, k = L - db * 5 + 1, d = db % 3, j = k - 1;
    const int Nj = TF * (5 - j);
    const uint32_t wb = sW + (L & 1) * T_W_BYTES;
    if (k == 1 && d == 0) {
      xr[0] = ld_shared16(sX + S.pos * 16);
      xr[1] = ld_shared16(sX + TPB + S.pos * 16);
    }
    // ---- MMA round of slice j (9 taps, K = 16 channels, N = Nj per M-tile)
    TR_TRACE(0);
    trunk_issue_round(a.order, warp, tmem, sX, wb, j, Nj, mbar);
    TR_TRACE(1);
    // ---- prefetch the next slice's weights into the other buffer while the tensor pipe runs
    const size_t w_next = w_elem + (size_t)9 * TF * Nj;
    if (L + 1 < total) {
      const int Nn = (k == 5) ? 5 * TF : Nj - TF;
      const uint4* src = reinterpret_cast<const uint4*>(a.w + w_next);
      const uint32_t dst = sW + ((L + 1) & 1) * T_W_BYTES;
      for (int i = tid; i < 18 * Nn; i += T_THREADS) cp_async16(dst + i * 16, src + i, 16);
    }
    float bias[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) bias[q] = __ldg(a.bias + (size_t)L * 16 + q);
    // ---- epilogue of this thread's rows, M-tile by M-tile as their MMAs complete
    const bool saving = a.db_bufs && n < a.save_count;  // activations are kept for the first save_count samples only
    bf16* save_cur = saving ? a.db_bufs[db] : nullptr;
    bf16* save_next = (saving && db + 1 < a.R * 3) ? a.db_bufs[db + 1] : nullptr;
    {
      mbar_wait(smem_u32(&mbar[S.tile]), L & 1);
      TR_TRACE(2);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem + S.tcol + j * TF, v);
      if (S.valid) {
        const int pos = S.pos, pix = S.pix;
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] += bias[q];
        if (k < 5) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * G_SLOPE;
          const uint4 lo = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          const uint4 hi = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
          st_shared16(sX + (2 * k) * TPB + pos * 16, lo);
          st_shared16(sX + (2 * k + 1) * TPB + pos * 16, hi);
          if (save_cur) {
            uint4* g = reinterpret_cast<uint4*>(save_cur + ((size_t)n * 256 + pix) * 80 + 16 * k);
            g[0] = lo; g[1] = hi;
          }
        } else {
          float xo[16];
          unpack8(ld_shared16(sX + pos * 16), xo);
          unpack8(ld_shared16(sX + TPB + pos * 16), xo + 8);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = fmaf(RES, v[q], xo[q]);
          if (d == 2) {
            float xq[16];
            unpack8(xr[0], xq);
            unpack8(xr[1], xq + 8);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = fmaf(RES, v[q], xq[q]);
          }
          const uint4 lo = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          const uint4 hi = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
          st_shared16(sX + pos * 16, lo);
          st_shared16(sX + TPB + pos * 16, hi);
          if (L + 1 == total) {
            uint4* g = reinterpret_cast<uint4*>(a.y_out + ((size_t)n * 256 + pix) * a.out_pitch);
            g[0] = lo; g[1] = hi;
          } else if (save_next) {
            uint4* g = reinterpret_cast<uint4*>(save_next + ((size_t)n * 256 + pix) * 80);
            g[0] = lo; g[1] = hi;
          }
        }
      }
      TR_TRACE(3); TR_TRACE(4); TR_TRACE(5);
    }
    w_elem = w_next;
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TR_TRACE(6);
  }
  if (warp == 0) tmem_dealloc(tmem, T_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Backward (data-gradient chain) of the whole trunk, same structure as the forward kernel.
// With D = [dz5, dz4, dz3, dz2, dz1] the gradient w.r.t. slice k of a dense block's concat buffer
// is one conv over D[0:(5-k)], so the backward of a dense block is again a dense
// block: slice i of D feeds every later output in one N = 16*(5-i) MMA round.  The epilogue
// multiplies by the LeakyReLU mask taken from the saved forward buffer, the last round adds the
// block / RRDB skips.  Every D slice is also stored to global memory for the batched
// weight-gradient launch that follows.
// ---------------------------------------------------------------------------------------------
struct TrunkBwdArgs {
  const bf16* g_in;            // (B,16,16,16): dL/d(trunk output)
  bf16* g_out;                 // (B,16,16,16): dL/d(trunk input)
  const bf16* const* fwd_bufs; // [3R] saved concat buffers (pitch 80): masks
  bf16* const* d_bufs;         // [3R] dz buffers (pitch 80), written here
  const bf16* w;               // slice-major images of the dense data-gradient matrices, 5 per block
  int R, B;
  int order;
};

__global__ void __launch_bounds__(T_THREADS, 2) trunk_bwd_kernel(const TrunkBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[TNMT];  // one per M-tile: its epilogue starts while the other tiles' MMAs run
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x;
  const uint32_t sX = smem_u32(smem);
  const uint32_t sW = sX + T_X_BYTES;
  const int n_db = a.R * 3;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), T_TMEM_COLS);
  if (tid == 32) { for (int i = 0; i < TNMT; ++i) mbar_init(smem_u32(&mbar[i]), 1); }
  for (int i = tid; i < T_X_BYTES / 16; i += T_THREADS) st_shared16(sX + i * 16, make_uint4(0, 0, 0, 0));
  // weights of the LAST block's slice 0 -> buffer 0
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.w + (size_t)(n_db - 1) * T_DB_ELEMS_);
    for (int i = tid; i < 18 * 5 * TF; i += T_THREADS) cp_async16(sW + i * 16, src + i, 16);
  }
  const TrunkSlot S = trunk_slot(warp, lane);
  // incoming gradient at this thread's position (bf16 x 16)
  uint4 gin[2], gr[2];
  {
    const uint4* g = reinterpret_cast<const uint4*>(a.g_in + ((size_t)n * 256 + S.pix) * TF);
    gin[0] = S.valid ? g[0] : make_uint4(0, 0, 0, 0);
    gin[1] = S.valid ? g[1] : make_uint4(0, 0, 0, 0);
    gr[0] = gr[1] = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  int L = 0;
  for (int db = n_db - 1; db >= 0; --db) {
    const int d = db % 3;
    const float s_in = (d == 2) ? RES : 1.f;
    const bf16* fbuf = a.fwd_bufs[db];
    bf16* dbuf = a.d_bufs[db];
    if (d == 2) { gr[0] = gin[0]; gr[1] = gin[1]; }
    // slice 0 of D: dz5 = 0.2 * s_in * g
    if (S.valid) {
      float g[16];
      unpack8(gin[0], g);
      unpack8(gin[1], g + 8);
      const float sc = RES * s_in;
#pragma unroll
      for (int q = 0; q < 16; ++q) g[q] *= sc;
      const uint4 lo = make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
      const uint4 hi = make_uint4(pack2(g[8], g[9]), pack2(g[10], g[11]), pack2(g[12], g[13]), pack2(g[14], g[15]));
      st_shared16(sX + S.pos * 16, lo);
      st_shared16(sX + TPB + S.pos * 16, hi);
      uint4* gd = reinterpret_cast<uint4*>(dbuf + ((size_t)n * 256 + S.pix) * 80);
      gd[0] = lo; gd[1] = hi;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    size_t w_elem = (size_t)db * T_DB_ELEMS_;
    for (int j = 0; j < 5; ++j, ++L) {
      const int t = j + 1;  // output of this round: dz_{5-t} (t < 5) or the block-input gradient (t == 5)
      const int Nj = TF * (5 - j);
      const uint32_t wb = sW + (L & 1) * T_W_BYTES;
      trunk_issue_round(a.order, warp, tmem, sX, wb, j, Nj, mbar);
      // prefetch the next slice's weights (next slice of this block, or slice 0 of the previous block)
      const size_t w_next = (j < 4) ? w_elem + (size_t)9 * TF * Nj : (size_t)(db - 1) * T_DB_ELEMS_;
      if (j < 4 || db > 0) {
        const int Nn = (j == 4) ? 5 * TF : Nj - TF;
        const uint4* src = reinterpret_cast<const uint4*>(a.w + w_next);
        const uint32_t dst = sW + ((L + 1) & 1) * T_W_BYTES;
        for (int i = tid; i < 18 * Nn; i += T_THREADS) cp_async16(dst + i * 16, src + i, 16);
      }
      // mask source for this round (issued before the wait so the latency hides behind the MMAs)
      uint4 mk[2];
      if (t < 5) {
        const uint4* m = reinterpret_cast<const uint4*>(fbuf + ((size_t)n * 256 + S.pix) * 80 + TF * (5 - t));
        mk[0] = m[0]; mk[1] = m[1];
      }
      {
        mbar_wait(smem_u32(&mbar[S.tile]), L & 1);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem + S.tcol + j * TF, v);
        if (!S.valid) {
        } else if (t < 5) {
          float m[16];
          unpack8(mk[0], m);
          unpack8(mk[1], m + 8);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] *= (m[q] > 0.f ? 1.f : G_SLOPE);
          const uint4 lo = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          const uint4 hi = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
          st_shared16(sX + (2 * t) * TPB + S.pos * 16, lo);
          st_shared16(sX + (2 * t + 1) * TPB + S.pos * 16, hi);
          uint4* gd = reinterpret_cast<uint4*>(dbuf + ((size_t)n * 256 + S.pix) * 80 + TF * t);
          gd[0] = lo; gd[1] = hi;
        } else {
          float g[16];
          unpack8(gin[0], g);
          unpack8(gin[1], g + 8);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = fmaf(s_in, g[q], v[q]);
          if (d == 0) {
            unpack8(gr[0], g);
            unpack8(gr[1], g + 8);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] += g[q];
          }
          gin[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          gin[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        }
      }
      w_elem = w_next;
      cp_async_wait_all();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (S.valid) {
    uint4* g = reinterpret_cast<uint4*>(a.g_out + ((size_t)n * 256 + S.pix) * TF);
    g[0] = gin[0]; g[1] = gin[1];
  }
  if (warp == 0) tmem_dealloc(tmem, T_TMEM_COLS);
}

// fp32 packed per-layer weights [tap][16k][16] of one dense block  ->  slice-major bf16 images:
// slice j: [tap][2 planes][N_j = 16*(5-j) rows = layers j+1..5 side by side][8 channels of slice j]
constexpr int T_DB_ELEMS = 9 * TF * TF * 15;
__global__ void pack_trunk_kernel(const float* __restrict__ pk_db0, bf16* __restrict__ dst, int n_db, int bwd) {
  const int db = blockIdx.y;
  if (db >= n_db) return;
  const float* src = pk_db0 + (size_t)db * T_DB_ELEMS;
  bf16* out = dst + (size_t)db * T_DB_ELEMS;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < T_DB_ELEMS; e += gridDim.x * blockDim.x) {
    int j = 0, base = 0;
    while (j < 4 && e >= base + 9 * TF * TF * (5 - j)) { base += 9 * TF * TF * (5 - j); ++j; }
    const int Nj = TF * (5 - j);
    int r = e - base;
    const int c8 = r & 7; r >>= 3;
    const int row = r % Nj; r /= Nj;
    const int pl = r & 1, tap = r >> 1;
    const int k = j + 1 + row / TF, co = row % TF;        // consuming layer (1-based) and its output channel
    const int ci = TF * j + pl * 8 + c8;                   // input channel inside the concat buffer
    // forward: layers 1..k-1 of this block precede layer k.  backward: "layer" t = k consumes the 16t
    // rows [dz5..dz_{6-t}] of the dense data-gradient matrix Wt_{5-t}, stored after Wt_0..Wt_{4-t}.
    const size_t layer_off = bwd ? (size_t)9 * TF * TF * (15 - k * (k + 1) / 2) : (size_t)9 * TF * TF * (k - 1) * k / 2;
    out[e] = __float2bfloat16_rn(src[layer_off + ((size_t)tap * (TF * k) + ci) * TF + co]);
  }
}

}  // namespace

int pack_trunk_slices(const float* pk_first_dense, void* dst_bf16, int n_db, int bwd, cudaStream_t st) {
  if (n_db <= 0) return 0;
  pack_trunk_kernel<<<dim3(32, n_db), 256, 0, st>>>(pk_first_dense, (bf16*)dst_bf16, n_db, bwd);
  DG_LAUNCH_CHECK();
  return 0;
}

// RRDBs [r0, r0 + R) of the trunk (all pointers are those of the WHOLE trunk): g_in = dL/d(output of RRDB r0 + R - 1),
// g_out = dL/d(input of RRDB r0); chaining calls from the last RRDB range to the first reproduces the single launch.
int trunk_bwd_fused(const void* g_in, void* g_out, void* const* fwd_bufs_dev, void* const* d_bufs_dev, const void* w_slices,
                    int R, int B, cudaStream_t st, int r0) {
  if (ablate(3)) return 0;
  fwd_bufs_dev += 3 * r0;
  d_bufs_dev += 3 * r0;
  w_slices = (const bf16*)w_slices + (size_t)3 * r0 * T_DB_ELEMS_;
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
    // two CTAs (images) per SM need the full 228 KB carve-out: one CTA's epilogue then overlaps the other's MMAs
    DG_CUDA(cudaFuncSetAttribute(trunk_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  TrunkBwdArgs a;
  a.g_in = (const bf16*)g_in; a.g_out = (bf16*)g_out;
  a.fwd_bufs = (const bf16* const*)fwd_bufs_dev; a.d_bufs = (bf16* const*)d_bufs_dev;
  a.w = (const bf16*)w_slices; a.R = R; a.B = B; a.order = g_tune[3];
  const double px = (double)B * 256;
  Prof prof(PC_DENSE_UMMA, 2.0 * px * 16.0 * 9.0 * 16.0 * 15.0 * 3.0 * R, px * 80.0 * 2.0 * 2.0 * 3.0 * R, st);
  trunk_bwd_kernel<<<B, T_THREADS, T_SMEM, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

bool trunk_fused_supported(int F, int Hc, int R, int bf) { return bf && F == TF && Hc == TW && R >= 1; }

// x_in: conv1 output view; y_out: trunk output (pitch out_pitch); db_bufs_dev: device array of 3R
// concat-buffer pointers (pitch 80) or nullptr when the activations need not be kept.
int trunk_fwd_fused(const void* x_in, int in_pitch, int in_coff, void* y_out, int out_pitch, void* const* db_bufs_dev,
                    const void* w_umma, const float* bias, int R, int B, int save_count, cudaStream_t st) {
  if (ablate(2)) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(trunk_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
    DG_CUDA(cudaFuncSetAttribute(trunk_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  TrunkArgs a;
  a.x_in = (const bf16*)x_in; a.in_pitch = in_pitch; a.in_coff = in_coff;
  a.y_out = (bf16*)y_out; a.out_pitch = out_pitch;
  a.db_bufs = (bf16* const*)db_bufs_dev;
  a.w = (const bf16*)w_umma; a.bias = bias; a.R = R; a.B = B; a.order = g_tune[3];
  a.save_count = db_bufs_dev ? save_count : 0;
  a.trace = nullptr;
  static const bool tracing = getenv("DG_TRUNK_TRACE") != nullptr;
  if (tracing) { cudaMalloc(&a.trace, 32 * 8 * 8); cudaMemset(a.trace, 0, 32 * 8 * 8); }
  const double px = (double)B * 256;
  Prof prof(PC_DENSE_UMMA, 2.0 * px * 16.0 * 9.0 * 16.0 * 15.0 * 3.0 * R, px * 16.0 * 2.0 * 2.0, st);
  trunk_fwd_kernel<<<B, T_THREADS, T_SMEM, st>>>(a);
  DG_LAUNCH_CHECK();
  if (tracing) {
    unsigned long long h[32 * 8];
    cudaDeviceSynchronize();
    cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.trace);
    double sum[7] = {0};
    for (int l = 0; l < 32; ++l) {
      const unsigned long long* e = h + l * 8;
      // [issue, ->M-tile 0 ready, its epilogue, -, -, sync] seen by thread 0 (warp 0 drains M-tile 0)
      for (int k = 0; k < 6; ++k) sum[k] += (double)(e[k + 1] - e[k]);
    }
    fprintf(stderr, "[trunk trace] order %d, layers 30..61 of CTA 0, mean ns: issue %.0f | ->mt0 ready %.0f | epi0 %.0f | ->mt2 ready %.0f | epi2 %.0f | sync %.0f | layer %.0f\n",
            a.order, sum[0] / 32, sum[1] / 32, sum[2] / 32, sum[3] / 32, sum[4] / 32, sum[5] / 32, (double)(h[31 * 8 + 6] - h[0]) / 32);
  }
  return 0;
}

}  // namespace dg
