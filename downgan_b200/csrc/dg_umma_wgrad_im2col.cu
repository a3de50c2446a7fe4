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
  IcArgs a;
  if (!plan_ic(op, a)) { set_error("wgrad_im2col: unsupported shape"); return DG_ERR_INVALID; }
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(wgrad_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    attr_set = true;
  }
  const unsigned smem = ic_smem(a);
  int per_sm = std::max(1, std::min(3, (int)((227u * 1024u) / (smem + 1024u))));
  long long S = std::min<long long>(a.tiles_total, 148LL * per_sm);
  a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
  S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
  const double total = (double)a.total_pos;
  Prof prof(PC_WGRAD_UMMA, 2.0 * total * op.Co * op.Ci * 9.0,
            total * op.Co * 2.0 + total * op.Ci * (op.x.bf ? 2.0 : 4.0), st);
  wgrad_im2col_kernel<<<(unsigned)S, IC_THREADS, smem, st>>>(a);
  DG_LAUNCH_CHECK();
  if (op.dbias) DG_TRY(colsum(op.dy, (size_t)a.total_pos, op.Co, op.dbias, st));
  return 0;
}

}  // namespace dg
