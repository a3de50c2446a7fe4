// tcgen05 weight-gradient kernel for 3x3 convolutions (stride 1 and 2), sm_100a.
//
//   dW[tap][ci][co] = sum over output positions p of  x[p (+) tap][ci] * dy[p][co]
//
// is a GEMM whose contraction index is the pixel position.  Both operands are staged in shared
// memory in the same "planar" layout the forward kernel uses,  [C/8 planes][positions][8 channels],
// which is ALSO the tcgen05 canonical no-swizzle MN-major layout (8 channels contiguous, consecutive
// K-steps = consecutive positions 16 B apart; LBO = 128 B between groups of 8 positions, SBO = plane
// stride between groups of 8 channels).  So no transposition is needed: A = x^T (M = Ci, padded to
// 128 lanes), B = dy (N = a chunk of Co), one descriptor start-address offset per tap, and the nine
// per-tap accumulators (9 x NT columns) live in TMEM across all the tiles a CTA walks.  Stride-2
// convolutions stage x as four parity sub-images so that every tap is again a constant offset.
// Partial sums of different CTAs are combined with vectorised fp32 reductions (red.global.add.v4).
#include <algorithm>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int WG_THREADS_U = 256;  // 8 warps stage tiles; warps w and w+4 share TMEM lane quarter w
constexpr int WG_MAX_SMEM = 227 * 1024 - 2048;

struct WgArgs {
  WgradOp op;
  int CoP;
  int TH, PWt, npos16, PBx, PBd, nplx, npld, nsub;
  int tiles_per_img, tiles_total, tiles_per_cta;
  int NT, tmem_cols;
  unsigned x_bytes, d_off, d_bytes;   // per-buffer x region size, offset of the dy buffers, per-buffer dy size
  unsigned magic_nx, magic_nd, magic_pw;  // ceil(2^32 / nplx), ceil(2^32 / npld), ceil(2^32 / PWt)
};

__device__ __forceinline__ void wgrad_body(const WgArgs& a, uint8_t* smem, uint64_t* mbar, uint32_t* tmem_slot_p) {
  uint32_t& tmem_slot = *tmem_slot_p;
  const WgradOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5;
  if ((int)blockIdx.y * a.NT >= op.Co) return;
  const int co0 = blockIdx.y * a.NT;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = min(a.tiles_total, t_begin + a.tiles_per_cta);
  if (t_begin >= t_end) return;
  // bias gradient: thread t sums channel (t % NT) over positions t / NT, t / NT + 128 / NT, ...
  const int bc = tid % a.NT, bslice = tid / a.NT, bstep = WG_THREADS_U / a.NT;
  const bool bias_on = op.dbias != nullptr && bslice < bstep;
  float bsum = 0.f;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  if (tid == 32) { mbar_init(smem_u32(&mbar[0]), WG_THREADS_U / 32); mbar_init(smem_u32(&mbar[1]), WG_THREADS_U / 32); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int s = op.stride;
  const bf16* xb = (const bf16*)op.x.p;
  const bf16* db = (const bf16*)op.dy.p;
  const uint32_t idesc = instr_desc(128, a.NT, 1, 1);
  const int ksteps = a.npos16 >> 4;

  // The MMAs contract over every staged position, so everything they can read must be finite:
  // zero both x regions and both dy regions once; afterwards a tile only rewrites its own rows
  // (stale x rows of earlier tiles only ever meet dy == 0).
  {
    const uint32_t zbytes = a.d_off + 2 * a.d_bytes;
    for (uint32_t i = tid * 16; i < zbytes; i += WG_THREADS_U * 16)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sbase + i), "r"(0u) : "memory");
    __syncthreads();
  }
  const int xrows = a.TH + ((s == 1) ? 2 : 1);
  const int xrow_elems = a.PWt * a.nplx, drow_elems = a.PWt * a.npld;
  const int ystep = (s == 1) ? 1 : 2;

  int it = 0;
  for (int t = t_begin; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    if (it >= 2) mbar_wait(smem_u32(&mbar[buf]), ((it - 2) >> 1) & 1);  // MMAs that read this buffer are done
    const int n = t / a.tiles_per_img;
    const int y0 = (t - n * a.tiles_per_img) * a.TH;
    const uint32_t sx = sbase + buf * a.x_bytes;
    const uint32_t sd = sbase + a.d_off + buf * a.d_bytes;
    // ---- x tile(s): a thread owns (column, plane) pairs and walks the rows (zero fill = padding)
    for (int e = tid; e < xrow_elems; e += WG_THREADS_U) {
      const int c = (int)__umulhi((unsigned)e, a.magic_nx), pl = e - c * a.nplx;
      for (int sub = 0; sub < a.nsub; ++sub) {
        int gy, gx;
        if (s == 1) { gy = y0 - 1; gx = c - 1; }
        else { gy = 2 * (y0 - 1) + (sub >> 1); gx = 2 * (c - 1) + (sub & 1); }
        const bool okx = gx >= 0 && gx < op.Win;
        const bf16* src = xb + (((size_t)n * op.Hin + gy) * op.Win + (okx ? gx : 0)) * op.x.pitch + op.x.coff + pl * 8;
        const size_t sstep = (size_t)ystep * op.Win * op.x.pitch;
        uint32_t dst = sx + (sub * a.nplx + pl) * a.PBx + c * 16;
        for (int r = 0; r < xrows; ++r, gy += ystep, src += sstep, dst += a.PWt * 16) {
          const bool ok = okx && gy >= 0 && gy < op.Hin;
          cp_async16(dst, ok ? src : xb, ok ? 16 : 0);
        }
      }
    }
    // ---- dy tile rows (pad column and rows outside the image are zero-filled)
    for (int e = tid; e < drow_elems; e += WG_THREADS_U) {
      const int c = (int)__umulhi((unsigned)e, a.magic_nd), pl = e - c * a.npld;
      const bool okx = c < op.Wout;
      const bf16* src = db + (((size_t)n * op.Hout + y0) * op.Wout + (okx ? c : 0)) * op.dy.pitch + op.dy.coff + co0 + pl * 8;
      const size_t sstep = (size_t)op.Wout * op.dy.pitch;
      uint32_t dst = sd + pl * a.PBd + c * 16;
      int gy = y0;
      for (int r = 0; r < a.TH; ++r, ++gy, src += sstep, dst += a.PWt * 16) {
        const bool ok = okx && gy < op.Hout;
        cp_async16(dst, ok ? src : db, ok ? 16 : 0);
      }
    }
    cp_async_wait_all();
    fence_proxy_async();
    __syncthreads();
    if (bias_on) {  // column sums of the staged dy tile (zero outside the image)
      const uint8_t* dz = smem + a.d_off + buf * a.d_bytes + (bc >> 3) * a.PBd + (bc & 7) * 2;
      for (int pos = bslice; pos < a.npos16; pos += bstep)
        bsum += __bfloat162float(*reinterpret_cast<const bf16*>(dz + pos * 16));
    }
    // MMA issue spread over the warps: lane 0 of warp w issues taps w, w + 8 (each tap has its own
    // TMEM accumulator, so the issue streams are independent); every issuing warp commits.
    if ((tid & 31) == 0) {
      tc_fence_after();
      const uint64_t bd0 = smem_desc(sd, 128, a.PBd);
      for (int tap = warp; tap < 9; tap += WG_THREADS_U / 32) {
        const int ky = tap / 3, kx = tap - 3 * ky;
        int sub = 0, shift;
        if (s == 1) {
          shift = ky * a.PWt + kx;
        } else {
          sub = ((ky == 1) ? 0 : 2) + ((kx == 1) ? 0 : 1);
          shift = ((ky == 0) ? 0 : 1) * a.PWt + ((kx == 0) ? 0 : 1);
        }
        uint64_t ad = smem_desc(sx + sub * a.nplx * a.PBx + shift * 16, 128, a.PBx);
        uint64_t bd = bd0;
        for (int ks = 0; ks < ksteps; ++ks, ad += 16, bd += 16)  // +16 positions x 16 B, in descriptor units
          umma_f16(tmem + tap * a.NT, ad, bd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar[buf]));
    }
    __syncwarp();
  }
  // ---- wait for the last MMAs of both buffers
  {
    const int last = it - 1;
    mbar_wait(smem_u32(&mbar[last & 1]), (last >> 1) & 1);
    if (it >= 2) mbar_wait(smem_u32(&mbar[(last - 1) & 1]), ((last - 1) >> 1) & 1);
  }
  tc_fence_after();
  // ---- epilogue: lane = input channel ci, columns = [tap][co chunk]; fp32 reductions into dW
  const int ci = (warp & 3) * 32 + (tid & 31);
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  for (int tap = (warp >> 2); tap < 9; tap += 2) {   // the two warps of a lane quarter alternate taps
    for (int nc = 0; nc < a.NT; nc += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + tap * a.NT + nc, v);
      if (ci < op.Ci) {
        float* dst = op.dw + ((size_t)tap * op.Ci + ci) * a.CoP + co0 + nc;
#pragma unroll
        for (int q = 0; q < 4; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  if (bias_on) atomicAdd(op.dbias + co0 + bc, bsum);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

__global__ void __launch_bounds__(WG_THREADS_U) wgrad_umma_kernel(const WgArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;
  wgrad_body(a, smem, mbar, &tmem_slot);
}

// one launch for many layers: blockIdx.z selects the op (all ops share the launch geometry limits)
__global__ void __launch_bounds__(WG_THREADS_U) wgrad_umma_batched_kernel(const WgArgs* __restrict__ table) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ WgArgs sa;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(table + blockIdx.z);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sa);
    for (int i = threadIdx.x; i < (int)(sizeof(WgArgs) / 4); i += WG_THREADS_U) dst[i] = src[i];
  }
  __syncthreads();
  wgrad_body(sa, smem, mbar, &tmem_slot);
}

bool plan_wgrad(const WgradOp& op, WgArgs& a) {
  if (!op.x.bf || !op.dy.bf) return false;
  if (op.Ci % 16 || op.Co % 16 || op.Ci > 128 || op.Co > 256) return false;
  if (op.x.pitch % 8 || op.x.coff % 8 || op.dy.pitch % 8 || op.dy.coff % 8) return false;
  if (op.stride != 1 && op.stride != 2) return false;
  if (op.stride == 2 && ((op.Hin & 1) || (op.Win & 1))) return false;
  const int s = op.stride;
  const int NT = (op.Co % 48 == 0) ? 48 : ((op.Co % 32 == 0) ? 32 : 16);
  const int PWt = (s == 1) ? op.Wout + 2 : op.Wout + 1;
  const int nplx = op.Ci / 8, npld = NT / 8, nsub = (s == 1) ? 1 : 4;
  int bestTH = 0;
  WgArgs b{};
  for (int TH = 1; TH <= op.Hout; ++TH) {
    const int npos16 = (TH * PWt + 15) & ~15;
    if (TH > 1 && npos16 > 384) break;
    const int xpos = (npos16 + 2 * PWt + 2 + 7) & ~7;
    const unsigned PBx = xpos * 16, PBd = npos16 * 16;
    const unsigned x_bytes = nsub * nplx * PBx, d_bytes = npld * PBd;
    // the A descriptor always spans 16 planes (M = 128): the tail may alias later buffers but must stay inside smem
    const unsigned span_end = x_bytes /*buffer 1 base*/ + (unsigned)((nsub - 1) * nplx + 16) * PBx;
    unsigned total = 2 * x_bytes + 2 * d_bytes;
    if (span_end > total) total = span_end;
    if (total > (unsigned)WG_MAX_SMEM) break;
    if ((long long)xpos * nplx >= 65536 || (long long)npos16 * npld >= 65536) break;  // exact index arithmetic
    bestTH = TH;
    b.TH = TH; b.PWt = PWt; b.npos16 = npos16; b.PBx = (int)PBx; b.PBd = (int)PBd;
    b.x_bytes = x_bytes; b.d_off = 2 * x_bytes; b.d_bytes = d_bytes;
  }
  if (bestTH == 0) return false;
  a = b;
  a.op = op;
  a.CoP = round_up(op.Co, 16);
  a.nplx = nplx; a.npld = npld; a.nsub = nsub; a.NT = NT;
  a.magic_nx = (unsigned)((0x100000000ULL + nplx - 1) / nplx);
  a.magic_nd = (unsigned)((0x100000000ULL + npld - 1) / npld);
  a.magic_pw = (unsigned)((0x100000000ULL + PWt - 1) / PWt);
  a.tiles_per_img = (op.Hout + a.TH - 1) / a.TH;
  a.tiles_total = a.tiles_per_img * op.B;
  int cols = 9 * NT, pc = 32;
  while (pc < cols) pc <<= 1;
  a.tmem_cols = pc;
  return true;
}

unsigned smem_total(const WgArgs& a) {
  const unsigned span_end = a.x_bytes + (unsigned)((a.nsub - 1) * a.nplx + 16) * a.PBx;
  unsigned total = 2 * a.x_bytes + 2 * a.d_bytes;
  return span_end > total ? span_end : total;
}

}  // namespace

bool wgrad_umma_supported(const WgradOp& op) {
  WgArgs a;
  return plan_wgrad(op, a);
}

int wgrad_umma(const WgradOp& op, cudaStream_t st) {
  WgArgs a;
  if (!plan_wgrad(op, a)) { set_error("wgrad_umma: unsupported shape"); return DG_ERR_INVALID; }
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_MAX_SMEM));
    attr_set = true;
  }
  const int n_chunks = op.Co / a.NT;
  // position split: enough CTAs to fill the chip, bounded so that the cross-CTA reductions stay small
  long long S = (2 * 148 + n_chunks - 1) / n_chunks;
  const long long cap = 3000000LL / (9LL * op.Ci * op.Co) + 1;
  if (S > cap) S = cap;
  if (S > a.tiles_total) S = a.tiles_total;
  if (S < 1) S = 1;
  a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
  S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  Prof prof(PC_WGRAD_UMMA, 2.0 * total * op.Co * op.Ci * 9.0,
            (double)total * op.Co * 2.0 + (double)op.B * op.Hin * op.Win * op.Ci * 2.0, st);
  wgrad_umma_kernel<<<dim3((unsigned)S, (unsigned)n_chunks), WG_THREADS_U, smem_total(a), st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;  // the bias gradient (op.dbias) is accumulated inside the kernel
}

size_t wgrad_umma_args_size() { return sizeof(WgArgs); }

// Many weight gradients in ONE launch (blockIdx.z = op).  `table_dev` holds n * wgrad_umma_args_size()
// bytes; `shadow` is the host copy of what was last uploaded (re-uploaded only when the plan changes,
// e.g. another batch size).  S_per_op = position split per op.
int wgrad_umma_batched(const WgradOp* ops, int n, void* table_dev, std::vector<unsigned char>& shadow, int S_per_op,
                       cudaStream_t st) {
  if (n <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(wgrad_umma_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_MAX_SMEM));
    attr_set = true;
  }
  std::vector<WgArgs> tab((size_t)n);
  unsigned smem = 0, gx = 1, gy = 1;
  double flops = 0, bytes = 0;
  for (int i = 0; i < n; ++i) {
    WgArgs& a = tab[(size_t)i];
    if (!plan_wgrad(ops[i], a)) { set_error("wgrad_umma_batched: op %d unsupported", i); return DG_ERR_INVALID; }
    long long S = S_per_op;
    if (S > a.tiles_total) S = a.tiles_total;
    if (S < 1) S = 1;
    a.tiles_per_cta = (int)((a.tiles_total + S - 1) / S);
    S = (a.tiles_total + a.tiles_per_cta - 1) / a.tiles_per_cta;
    gx = std::max(gx, (unsigned)S);
    gy = std::max(gy, (unsigned)(ops[i].Co / a.NT));
    smem = std::max(smem, smem_total(a));
    const double total = (double)ops[i].B * ops[i].Hout * ops[i].Wout;
    flops += 2.0 * total * ops[i].Co * ops[i].Ci * 9.0;
    bytes += total * ops[i].Co * 2.0 + (double)ops[i].B * ops[i].Hin * ops[i].Win * ops[i].Ci * 2.0;
  }
  const size_t nbytes = sizeof(WgArgs) * (size_t)n;
  if (shadow.size() != nbytes || memcmp(shadow.data(), tab.data(), nbytes) != 0) {
    shadow.assign((const unsigned char*)tab.data(), (const unsigned char*)tab.data() + nbytes);
    DG_CUDA(cudaMemcpyAsync(table_dev, shadow.data(), nbytes, cudaMemcpyHostToDevice, st));
  }
  Prof prof(PC_WGRAD_UMMA, flops, bytes, st);
  wgrad_umma_batched_kernel<<<dim3(gx, gy, (unsigned)n), WG_THREADS_U, smem, st>>>((const WgArgs*)table_dev);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg
