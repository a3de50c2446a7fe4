// tcgen05 weight gradient for convolutions with FEW input channels (9*Ci <= 256), stride 1, sm_100a.
//
// The generic wgrad kernel issues one MMA per tap and K-step with N = Co; with few channels every
// MMA re-reads a 4 KB (M = 128 padded) A block for a handful of useful rows.  Here the nine taps are
// folded into N instead: the CTA builds an im2col tile  R[pos][tap*Ci + ci]  (bf16, planar
// [NR/8 planes][positions][8]) in shared memory and issues ONE MMA per 16 positions,
//     D[co][tap*Ci+ci] += sum_pos dy[pos][co] * R[pos][tap*Ci+ci]        (A = dy^T, B = R, both MN-major)
// so the A block is read once per K-step instead of nine times.  Used for the critic's first layer
// (Ci = 2: the fp32 fine fields are gathered and rounded to bf16 while building R; 18 -> NR = 32
// columns) and for 16-channel inputs (NR = 144, vector copies).  Positions are a flat range over the
// whole batch (no halo structure is needed once im2col is explicit).
#include <algorithm>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int IC_THREADS = 256;
constexpr int IC_TPOS = 256;              // positions per tile (16 K-steps)
constexpr int IC_PB = IC_TPOS * 16;       // plane stride in bytes

struct IcArgs {
  WgradOp op;
  int CoP, NR, nplB, nplA, tmem_cols;
  long long total_pos;
  int tiles_total, tiles_per_cta;
  unsigned a_off;  // byte offset of the dy tile (after the im2col planes)
  unsigned xs_bytes;  // staged fp32 input rows per ring slot (pipelined first-layer kernel)
  int bias_n;         // staged first-layer kernel: > 0 = im2col column 9*Ci is 1 for samples < bias_n (fused bias gradient)
};

__device__ __forceinline__ float ldx(const void* p, int bf, size_t i) {
  return bf ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// few channels (CI <= 3): a thread gathers the 9*CI values of a position and packs them to bf16
template <int CI>
__device__ __forceinline__ void gather_tile(const WgradOp& op, uint32_t sB, long long p0, long long total_pos, int tid) {
  constexpr int NV = (9 * CI + 15) / 16 * 16;
  const int H = op.Hout, W = op.Wout, Hi = op.Hin, Wi = op.Win, st = op.stride;
  for (int pos = tid; pos < IC_TPOS; pos += IC_THREADS) {
    const long long p = p0 + pos;
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = 0.f;
    if (p < total_pos) {
      const int x = (int)(p % W);
      const long long q = p / W;
      const int y = (int)(q % H);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int gy = st * y + tap / 3 - 1, gx = st * x + tap % 3 - 1;
        if (gy >= 0 && gy < Hi && gx >= 0 && gx < Wi) {
          const size_t base = (((size_t)(q / H) * Hi + gy) * Wi + gx) * op.x.pitch + op.x.coff;
#pragma unroll
          for (int c = 0; c < CI; ++c) v[tap * CI + c] = ldx(op.x.p, op.x.bf, base + c);
        }
      }
    }
#pragma unroll
    for (int pl = 0; pl < NV / 8; ++pl)
      sts16(sB + pl * IC_PB + pos * 16, make_uint4(pk2(v[8 * pl], v[8 * pl + 1]), pk2(v[8 * pl + 2], v[8 * pl + 3]),
                                                   pk2(v[8 * pl + 4], v[8 * pl + 5]), pk2(v[8 * pl + 6], v[8 * pl + 7])));
  }
}

__global__ void __launch_bounds__(IC_THREADS) wgrad_im2col_kernel(const IcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const WgradOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = min(a.tiles_total, t_begin + a.tiles_per_cta);
  if (t_begin >= t_end) return;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  if (tid == 32) mbar_init(smem_u32(&mbar), 1);
  const uint32_t sB = smem_u32(smem), sA = sB + a.a_off;
  // everything the MMAs contract over must be finite: zero the B planes and the real dy planes once
  {
    const uint32_t zb = a.a_off + a.nplA * IC_PB;
    for (uint32_t i = tid * 16; i < zb; i += IC_THREADS * 16) sts16(sB + i, make_uint4(0, 0, 0, 0));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = instr_desc(128, a.NR, 1, 1);
  const int H = op.Hout, W = op.Wout, Hi = op.Hin, Wi = op.Win, Ci = op.Ci, cst = op.stride;
  const bf16* db = (const bf16*)op.dy.p;

  int it = 0;
  for (int t = t_begin; t < t_end; ++t, ++it) {
    const long long p0 = (long long)t * IC_TPOS;
    // ---- dy tile: [nplA planes][pos][8]
    for (int i = tid; i < IC_TPOS * a.nplA; i += IC_THREADS) {
      const int pl = i / IC_TPOS, pos = i - pl * IC_TPOS;
      const long long p = p0 + pos;
      const bool ok = p < a.total_pos;
      cp_async16(sA + pl * IC_PB + pos * 16, ok ? db + (size_t)p * op.dy.pitch + op.dy.coff + pl * 8 : db, ok ? 16 : 0);
    }
    // ---- im2col tile
    if (Ci % 8 == 0 && op.x.bf) {
      // 8-channel groups are contiguous in NHWC: one 16-byte copy per (position, tap, group)
      const int gpt = Ci >> 3;  // groups per tap
      const bf16* xb = (const bf16*)op.x.p;
      for (int i = tid; i < IC_TPOS * 9 * gpt; i += IC_THREADS) {
        const int r = i / IC_TPOS, pos = i - r * IC_TPOS, tap = r / gpt, g = r - tap * gpt;
        const long long p = p0 + pos;
        bool ok = p < a.total_pos;
        const bf16* src = xb;
        if (ok) {
          const int x = (int)(p % W);
          const long long q = p / W;
          const int y = (int)(q % H);
          const int gy = cst * y + tap / 3 - 1, gx = cst * x + tap % 3 - 1;
          ok = gy >= 0 && gy < Hi && gx >= 0 && gx < Wi;
          src = xb + (((size_t)(q / H) * Hi + gy) * Wi + gx) * op.x.pitch + op.x.coff + g * 8;
        }
        cp_async16(sB + (tap * gpt + g) * IC_PB + pos * 16, ok ? src : xb, ok ? 16 : 0);
      }
    } else if (Ci == 1) {
      gather_tile<1>(op, sB, p0, a.total_pos, tid);
    } else if (Ci == 2) {
      gather_tile<2>(op, sB, p0, a.total_pos, tid);
    } else {
      gather_tile<3>(op, sB, p0, a.total_pos, tid);
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      uint64_t ad = smem_desc(sA, 128, IC_PB), bd = smem_desc(sB, 128, IC_PB);
      for (int ks = 0; ks < IC_TPOS / 16; ++ks, ad += 16, bd += 16)
        umma_f16(tmem, ad, bd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), it & 1);  // single buffer: the tile is rebuilt in place
    tc_fence_after();
  }
  // ---- epilogue: lane = co, column = tap*Ci + ci
  const int co = (warp & 3) * 32 + (tid & 31);
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  int piece = 0;
  for (int nc = 0; nc < a.NR; nc += 16, ++piece) {
    if ((piece & 1) != (warp >> 2)) continue;
    float v[16];
    tmem_ld16(tmem + lane_base + nc, v);
    if (co < op.Co) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = nc + j;
        if (col < 9 * Ci) atomicAdd(op.dw + (size_t)col * a.CoP + co, v[j]);  // dw is [tap][ci][CoP]: row = tap*Ci+ci
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// Pipelined variant for the critic's first layer (Ci <= 3: fp32 fine fields, 9*Ci <= 32 im2col columns).
// Warp-specialised: warps 0..3 BUILD tile i+1 (gather + round the im2col rows, cp.async the dy rows) into
// the other ring slot while warp 4's elected thread issues the 16 MMAs of tile i (M = 64: Co <= 64 rows,
// accumulator row r in TMEM lane (r % 16) + 32 * (r / 16)); one accumulator for the whole CTA, read back
// once at the end.
// ---------------------------------------------------------------------------------------------------
constexpr int L1_THREADS = 288;       // 8 builder warps + 1 MMA warp
constexpr int L1_BUILDERS = 256;
constexpr int L1_NPLB = 4;            // im2col planes (NR = 32 columns)
constexpr int L1_NPLA = 8;            // planes an M = 64 A descriptor spans (Co / 8 of them are real)
constexpr int L1_NSTAGE = 3;

__device__ __forceinline__ void mbar_arrive_l1(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_l1(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ uint32_t elect_one_sync_l1() {
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

// ring slot = [4 im2col planes][nplA real dy planes]; an M = 64 A descriptor spans 8 planes from the dy base, the
// ones past the real planes alias the next slot / the tail pad (finite values feeding ignored accumulator rows)
template <int CI, bool STAGED>
__global__ void __launch_bounds__(L1_THREADS) wgrad_l1_kernel(const IcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * L1_NSTAGE + 1];  // full[], empty[], done
  __shared__ uint32_t tmem_slot;
  const WgradOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = min(a.tiles_total, t_begin + a.tiles_per_cta);
  const int my_tiles = max(0, t_end - t_begin);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (L1_NSTAGE + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * L1_NSTAGE);
  constexpr int MMA_WARP = L1_BUILDERS / 32;
  const uint32_t xs_bytes = STAGED ? (uint32_t)a.xs_bytes : 0u;  // fp32 input rows of the tile (zero halo)
  const uint32_t stage_bytes = (uint32_t)(L1_NPLB + a.nplA) * IC_PB + xs_bytes;
  if (warp == MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), 32);
  if (tid == 0) {
    for (int s = 0; s < L1_NSTAGE; ++s) { mbar_init(full_bar(s), L1_BUILDERS); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
  }
  const uint32_t s0 = smem_u32(smem);
  // everything the MMAs contract over must be finite: zero the ring and the tail pad once
  for (uint32_t i = tid * 16; i < L1_NSTAGE * stage_bytes + L1_NPLA * IC_PB; i += L1_THREADS * 16) sts16(s0 + i, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < MMA_WARP) {
    // ================= builders: one position per thread per tile =================
    const bf16* db = (const bf16*)op.dy.p;
    const unsigned H = (unsigned)op.Hout, W = (unsigned)op.Wout;
    const int Hi = op.Hin, Wi = op.Win, st = op.stride, pitch = op.x.pitch;
    const bool vec2 = (CI == 2) && !op.x.bf && ((pitch & 1) == 0) && ((op.x.coff & 1) == 0);
    // the 9*CI im2col values of this thread's position of a tile (fp32, rounded when stored): 32-bit index
    // arithmetic, the nine taps are constant offsets from the centre pixel, two fp32 channels = one 8-byte load
    auto gather = [&](int it, float* v) {
#pragma unroll
      for (int j = 0; j < 9 * CI; ++j) v[j] = 0.f;
      const long long p = (long long)(t_begin + it) * IC_TPOS + tid;
      if (p >= a.total_pos) return;
      const unsigned pu = (unsigned)p;
      const unsigned q = pu / W, x = pu - q * W;
      const unsigned n = q / H, y = q - n * H;
      const int cy = st * (int)y, cx = st * (int)x;
      const size_t centre = ((size_t)(n * (unsigned)Hi + (unsigned)cy) * (unsigned)Wi + (unsigned)cx) * pitch + op.x.coff;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int gy = cy + dy, gx = cx + dx;
        if (gy >= 0 && gy < Hi && gx >= 0 && gx < Wi) {
          const size_t base = centre + (long long)(dy * Wi + dx) * pitch;
          if (vec2) {
            const float2 t = *reinterpret_cast<const float2*>((const float*)op.x.p + base);
            v[tap * CI] = t.x; v[tap * CI + (CI > 1 ? 1 : 0)] = t.y;
          } else {
#pragma unroll
            for (int c = 0; c < CI; ++c) v[tap * CI + c] = ldx(op.x.p, op.x.bf, base + c);
          }
        }
      }
    };
    auto issue_dy = [&](int it) {
      const int s = it % L1_NSTAGE;
      if (!STAGED || it == 0) mbar_wait_l1(empty_bar(s), (((uint32_t)(it / L1_NSTAGE)) & 1u) ^ 1u);
      const uint32_t sA = s0 + s * stage_bytes + L1_NPLB * IC_PB;
      const long long p0 = (long long)(t_begin + it) * IC_TPOS;
      for (int i = tid; i < IC_TPOS * a.nplA; i += L1_BUILDERS) {
        const int pl = i / IC_TPOS, pos = i - pl * IC_TPOS;
        const long long p = p0 + pos;
        const bool ok = p < a.total_pos;
        cp_async16(sA + pl * IC_PB + pos * 16, ok ? db + (size_t)p * op.dy.pitch + op.dy.coff + pl * 8 : db, ok ? 16 : 0);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (STAGED) {
      // fp32 NHWC input with pitch == CI: the tile is TR whole image rows (TR * W = 256 positions); its TR + 2 input
      // rows are staged once in shared memory (zero halo), so a position's nine taps are constant smem offsets
      const int Wl = 31 - __clz((int)W), TR = IC_TPOS >> Wl, PWx = (int)W + 2;
      const int tiles_per_img = (int)H / TR;
      const float* xg = (const float*)op.x.p;
      auto issue_x = [&](int it) {
        const int s = it % L1_NSTAGE;
        const int t = t_begin + it;
        const int n = t / tiles_per_img, y0 = (t - n * tiles_per_img) * TR;
        const uint32_t xs = s0 + s * stage_bytes + (L1_NPLB + a.nplA) * IC_PB;
        const int chunks = (TR + 2) * (int)W;  // one pixel (CI floats) per copy
        for (int i = tid; i < chunks; i += L1_BUILDERS) {
          const int rr = i >> Wl, cc = i & ((int)W - 1);
          const int gy = y0 - 1 + rr;
          const bool ok = gy >= 0 && gy < Hi;
          const float* src = xg + ((size_t)(n * Hi + (ok ? gy : 0)) * Wi + cc) * CI;
          const uint32_t dst = xs + (uint32_t)((rr * PWx + cc + 1) * CI) * 4u;
          if (CI == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(ok ? 8 : 0) : "memory");
          else {
#pragma unroll
            for (int c = 0; c < CI; ++c)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4u * c), "l"(src + c), "r"(ok ? 4 : 0) : "memory");
          }
        }
      };
      if (my_tiles > 0) { issue_x(0); issue_dy(0); }
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % L1_NSTAGE;
        const bool more = it + 1 < my_tiles;
        if (more) {
          mbar_wait_l1(empty_bar((it + 1) % L1_NSTAGE), (((uint32_t)((it + 1) / L1_NSTAGE)) & 1u) ^ 1u);
          issue_x(it + 1);
          issue_dy(it + 1);
        }
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(L1_BUILDERS) : "memory");  // every builder's row copies have landed
        const uint32_t sB = s0 + s * stage_bytes;
        const float* xs = reinterpret_cast<const float*>(smem + (size_t)s * stage_bytes + (size_t)(L1_NPLB + a.nplA) * IC_PB);
        const int r = tid >> Wl, c = tid & ((int)W - 1);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        // fused bias gradient: a column of ones makes D[co][9*CI] = sum_pos dy[pos][co] (tiles never straddle samples)
        if (9 * CI < 32) v[9 * CI < 32 ? 9 * CI : 0] = ((t_begin + it) / tiles_per_img < a.bias_n) ? 1.f : 0.f;
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
          sts16(sB + pl * IC_PB + tid * 16, make_uint4(pk2(v[8 * pl], v[8 * pl + 1]), pk2(v[8 * pl + 2], v[8 * pl + 3]),
                                                       pk2(v[8 * pl + 4], v[8 * pl + 5]), pk2(v[8 * pl + 6], v[8 * pl + 7])));
        fence_proxy_async();
        mbar_arrive_l1(full_bar(s));
      }
    } else {
    float vn[9 * CI];
    if (my_tiles > 0) { issue_dy(0); gather(0, vn); }
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % L1_NSTAGE;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (j < 9 * CI) ? vn[j < 9 * CI ? j : 0] : 0.f;
      const bool more = it + 1 < my_tiles;
      if (more) { issue_dy(it + 1); gather(it + 1, vn); }  // next tile's loads fly while this one is finished
      const uint32_t sB = s0 + s * stage_bytes;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl)
        sts16(sB + pl * IC_PB + tid * 16, make_uint4(pk2(v[8 * pl], v[8 * pl + 1]), pk2(v[8 * pl + 2], v[8 * pl + 3]),
                                                     pk2(v[8 * pl + 4], v[8 * pl + 5]), pk2(v[8 * pl + 6], v[8 * pl + 7])));
      if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async();
      mbar_arrive_l1(full_bar(s));
    }
    }
  } else {
    // ================= MMA issue =================
    const uint32_t idesc = instr_desc(64, 32, 1, 1);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % L1_NSTAGE;
      mbar_wait_l1(full_bar(s), ((uint32_t)(it / L1_NSTAGE)) & 1u);
      tc_fence_after();
      const uint32_t sB = s0 + s * stage_bytes, sA = sB + L1_NPLB * IC_PB;
      if (elect_one_sync_l1()) {
        const uint64_t ad0 = smem_desc(sA, 128, IC_PB), bd0 = smem_desc(sB, 128, IC_PB);
#pragma unroll
        for (int ks = 0; ks < IC_TPOS / 16; ++ks)
          umma_f16(tmem, ad0 + (uint64_t)(16 * ks), bd0 + (uint64_t)(16 * ks), idesc, (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(empty_bar(s));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  }
  // ---- epilogue: M = 64 accumulator, row co lives in lane (co % 16) + 32 * (co / 16); column = tap*Ci + ci
  if (my_tiles > 0 && warp < 4) {  // warps 0..3 cover the four TMEM lane quarters
    mbar_wait_l1(done_bar, 0);
    tc_fence_after();
    const int lane = tid & 31;
    const int co = (lane < 16) ? warp * 16 + lane : -1;
#pragma unroll
    for (int nc = 0; nc < 32; nc += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + nc, v);
      if (co >= 0 && co < op.Co) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (nc + j < 9 * CI) atomicAdd(op.dw + (size_t)(nc + j) * a.CoP + co, v[j]);
          else if (STAGED && nc + j == 9 * CI && a.bias_n > 0) atomicAdd(op.dbias + co, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem, 32);
}

bool plan_ic(const WgradOp& op, IcArgs& a) {
  if (op.stride == 1) { if (op.Hin != op.Hout || op.Win != op.Wout) return false; }
  else if (op.stride == 2) { if (op.Hin != 2 * op.Hout || op.Win != 2 * op.Wout) return false; }
  else return false;
  if (!op.dy.bf || op.Co % 16 || op.Co > 128 || op.dy.pitch % 8 || op.dy.coff % 8) return false;
  const bool vec = (op.Ci % 8 == 0) && op.x.bf && op.x.pitch % 8 == 0 && op.x.coff % 8 == 0;
  if (!vec && op.Ci > 3) return false;
  if (9 * op.Ci > 256) return false;
  a.op = op;
  a.CoP = round_up(op.Co, 16);
  a.NR = round_up(9 * op.Ci, 16);
  a.nplB = a.NR / 8;
  a.nplA = op.Co / 8;
  int pc = 32;
  while (pc < a.NR) pc <<= 1;
  a.tmem_cols = pc;
  a.total_pos = (long long)op.B * op.Hout * op.Wout;
  a.tiles_total = (int)((a.total_pos + IC_TPOS - 1) / IC_TPOS);
  a.a_off = (unsigned)(a.nplB * IC_PB);
  return true;
}
// the A descriptor (M = 128) spans 16 planes from the dy tile base; the tail may be garbage but must be inside smem
unsigned ic_smem(const IcArgs& a) { return a.a_off + 16 * IC_PB; }

}  // namespace

bool wgrad_im2col_supported(const WgradOp& op) {
  IcArgs a;
  // policy: stride-2 layers run faster on the parity-sub-image kernel (measured), the kernel itself handles both
  return op.stride == 1 && plan_ic(op, a) && ic_smem(a) <= 227 * 1024 - 2048;
}

int wgrad_im2col(const WgradOp& op, cudaStream_t st) {
  if (ablate(1)) return 0;
  IcArgs a;
  if (!plan_ic(op, a)) { set_error("wgrad_im2col: unsupported shape"); return DG_ERR_INVALID; }
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(wgrad_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  const double total = (double)a.total_pos;
  const int bias_B = (op.dbias_B > 0 && op.dbias_B < op.B) ? op.dbias_B : op.B;
  bool fused_bias = false;
  a.bias_n = 0;
  {
    Prof prof(PC_WGRAD_UMMA, 2.0 * total * op.Co * op.Ci * 9.0, total * op.Co * 2.0 + total * op.Ci * (op.x.bf ? 2.0 : 4.0), st);
    if (op.Ci <= 3 && op.Co <= 64 && a.total_pos < (1LL << 31)) {
      // staged variant: fp32 NHWC input with pitch == Ci, stride 1, tiles = whole image rows
      const bool pow2 = (op.Wout & (op.Wout - 1)) == 0;
      const bool staged = !op.x.bf && op.x.pitch == op.Ci && op.x.coff == 0 && op.stride == 1 && pow2 && op.Wout <= IC_TPOS &&
                          (op.Hout % (IC_TPOS / op.Wout)) == 0 && (op.Ci == 2 || op.Ci == 1 || op.Ci == 3);
      a.xs_bytes = staged ? (unsigned)(((IC_TPOS / op.Wout + 2) * (op.Wout + 2) * op.Ci * 4 + 127) & ~127) : 0u;
      fused_bias = staged && op.dbias && g_tune[8];
      a.bias_n = fused_bias ? bias_B : 0;
      const unsigned l1_smem = (unsigned)(L1_NSTAGE * ((L1_NPLB + a.nplA) * IC_PB + a.xs_bytes) + L1_NPLA * IC_PB);
      long long S = std::min<long long>(a.tiles_total, 148LL * 2);
      a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
      S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
#define L1_LAUNCH(CI, ST)                                                                                                     \
  do {                                                                                                                        \
    static bool attr = false;                                                                                                 \
    if (!attr) {                                                                                                              \
      DG_CUDA(cudaFuncSetAttribute(wgrad_l1_kernel<CI, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));        \
      DG_CUDA(cudaFuncSetAttribute(wgrad_l1_kernel<CI, ST>, cudaFuncAttributePreferredSharedMemoryCarveout,                   \
                                   cudaSharedmemCarveoutMaxShared));                                                          \
      attr = true;                                                                                                            \
    }                                                                                                                         \
    wgrad_l1_kernel<CI, ST><<<(unsigned)S, L1_THREADS, l1_smem, st>>>(a);                                                     \
  } while (0)
      if (op.Ci == 1) { if (staged) L1_LAUNCH(1, true); else L1_LAUNCH(1, false); }
      else if (op.Ci == 2) { if (staged) L1_LAUNCH(2, true); else L1_LAUNCH(2, false); }
      else { if (staged) L1_LAUNCH(3, true); else L1_LAUNCH(3, false); }
#undef L1_LAUNCH
    } else {
      const unsigned smem = ic_smem(a);
      int per_sm = std::max(1, std::min(3, (int)((227u * 1024u) / (smem + 1024u))));
      long long S = std::min<long long>(a.tiles_total, 148LL * per_sm);
      a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
      S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
      wgrad_im2col_kernel<<<(unsigned)S, IC_THREADS, smem, st>>>(a);
    }
  }
  DG_LAUNCH_CHECK();
  if (op.dbias && !fused_bias) DG_TRY(colsum(op.dy, (size_t)bias_B * op.Hout * op.Wout, op.Co, op.dbias, st));
  return 0;
}

}  // namespace dg
