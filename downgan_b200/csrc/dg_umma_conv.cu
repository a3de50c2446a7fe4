// tcgen05 / TMEM implicit-GEMM 3x3 convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// Shared-memory layout ("planar"): the halo-padded input tile is stored as
//     [Ci/8 planes][positions][8 channels]      (16 bytes per position per plane)
// with positions = tile rows x padded width, row-major.  This is the tcgen05 canonical NO-SWIZZLE
// K-major layout (core matrix = 8 consecutive positions x 16 B; SBO = 128 B; LBO = plane stride), so
// the A operand of a tap is the SAME tile at a start-address offset of (dy*PW+dx)*16 bytes: the
// tile is staged once and never re-gathered per tap.  MMA rows are consecutive padded positions;
// the pad column(s) of each image row are computed and discarded in the epilogue.
//
// Three modes share the kernel:
//   S1      stride-1 conv (also the data-gradient of a stride-1 conv, with flipped weights)
//   S2_FWD  stride-2 conv: the input is staged as four parity sub-images so that every tap is again
//           a constant position offset inside one sub-image
//   S2_DGRAD data-gradient of a stride-2 conv: the tile is in dy resolution, the four output parity
//           classes are four TMEM accumulators, each fed by its 1/2/2/4 contributing taps
// Weights are pre-packed in global memory in the matching B layout [tap*Ci/8 planes][Co][8]; a CTA
// stages the NT output channels of its blockIdx.y chunk.  The accumulator lives in TMEM (lane =
// position, column = output channel) and is read back with tcgen05.ld for the fused epilogue (bias,
// residual scale + skips, LeakyReLU / mask, bf16 pack, pixel-(un)shuffle addressing, channel-offset
// store into the dense-block concat buffer).
#include <algorithm>

#include "dg_umma.cuh"

namespace dg {
namespace {

using namespace um;

constexpr int UMMA_THREADS = 256;  // 8 warps: all stage tiles; warps w and w+4 share TMEM lane quarter w and split the epilogue
constexpr int MAX_SMEM = 227 * 1024 - 2048;  // dynamic limit (static smem of the kernel comes on top)
enum Mode { S1 = 0, S2_FWD = 1, S2_DGRAD = 2 };

struct UmmaArgs {
  ConvOp op;
  int mode;
  int TH, PW, n_mt, PB, tiles_per_img, NT, tmem_cols, nplanes, nsub, CoP;
  int Ht, Wt;      // tile-space extent (S1/S2_FWD: output grid; S2_DGRAD: dy grid)
  unsigned magic_np, magic_pw;  // ceil(2^32 / nplanes), ceil(2^32 / PW)
  unsigned w_off;  // byte offset of the weight image in dynamic smem (after the two input-tile buffers)
  unsigned a_bytes;  // bytes of one input-tile buffer
  int tiles_total, nbuf;
};

// 16 consecutive channels of a view at element index i (16-byte aligned, checked on the host)
__device__ __forceinline__ void load16(const TV& t, size_t i, float* v) {
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
__device__ __forceinline__ void store16(const TV& t, size_t i, const float* v) {
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
__device__ __forceinline__ void st1(const TV& t, size_t i, float v) {
  if (t.bf) ((bf16*)t.p)[i] = __float2bfloat16_rn(v);
  else ((float*)t.p)[i] = v;
}

// fused epilogue for 16 channels [nc, nc+16) of output pixel (n, yo, xo)
__device__ __forceinline__ void epilogue16(const ConvOp& op, float* v, int n, int yo, int xo, int nc) {
  const size_t p = ((size_t)n * op.Hout + yo) * op.Wout + xo;
  if (op.bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += op.bias[nc + j];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] *= op.s_acc;
  if (op.r1.p) {
    float t[16];
    load16(op.r1, p * op.r1.pitch + op.r1.coff + nc, t);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s1, t[j], v[j]);
  }
  if (op.r2.p) {
    float t[16];
    load16(op.r2, p * op.r2.pitch + op.r2.coff + nc, t);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaf(op.s2, t[j], v[j]);
  }
  if (op.act == ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
  } else if (op.act == ACT_MASK) {
    float t[16];
    load16(op.mask, p * op.mask.pitch + op.mask.coff + nc, t);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= (t[j] > 0.f ? 1.f : op.slope);
  }
  if (op.shuffle == SHUF_NONE) {
    store16(op.y, p * op.y.pitch + op.y.coff + nc, v);
  } else if (op.shuffle == SHUF_PIXEL) {  // nn.PixelShuffle(2) folded into the store addressing
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int co = nc + j, cc = co >> 2, ii = (co >> 1) & 1, jj = co & 1;
      const size_t qq = ((size_t)n * (2 * op.Hout) + 2 * yo + ii) * (2 * op.Wout) + 2 * xo + jj;
      st1(op.y, qq * op.y.pitch + op.y.coff + cc, v[j]);
    }
  } else {  // inverse shuffle: data-gradient w.r.t. the pre-shuffle activation
    const size_t qq = ((size_t)n * (op.Hout >> 1) + (yo >> 1)) * (op.Wout >> 1) + (xo >> 1);
    store16(op.y, qq * op.y.pitch + op.y.coff + (2 * (yo & 1) + (xo & 1)) * op.Co + nc, v);  // (i,j)-major, see PackDesc::ps
  }
}

// Persistent over tiles: a CTA stages its weight slice ONCE, then walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...; the cp.async staging of tile i+1 (other buffer) is issued right
// after the MMAs of tile i, so it overlaps the tensor pipe and the epilogue of tile i.
__global__ void __launch_bounds__(UMMA_THREADS) conv_umma_kernel(const UmmaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int co0 = blockIdx.y * a.NT;
  const int PW = a.PW, mode = a.mode;
  if ((int)blockIdx.x >= a.tiles_total) return;

  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
  const int nissue = min(UMMA_THREADS / 32, a.n_mt * ((mode == S2_DGRAD) ? 4 : 1));
  if (tid == 32) mbar_init(smem_u32(&mbar), (uint32_t)nissue);

  const uint32_t sa0 = smem_u32(smem);
  const uint32_t sw = sa0 + a.w_off;
  const bf16* xb = (const bf16*)op.x.p;
  const int trows = a.TH + ((mode == S1) ? 2 : 1);

  // stage one input tile (zero fill = conv padding); only the (TH + halo) rows are written, positions
  // beyond feed discarded rows only.  A thread owns (column c, plane pl) pairs and walks the rows, so
  // the per-chunk work is an add and a compare (no index decomposition in the inner loop).
  const int row_elems = PW * a.nplanes;
  auto load_tile = [&](int tile, uint32_t sa) {
    const int n = tile / a.tiles_per_img;
    const int y0 = (tile - n * a.tiles_per_img) * a.TH;
    const int ystep = (mode == S2_FWD) ? 2 : 1;
    for (int e = tid; e < row_elems; e += UMMA_THREADS) {
      const int c = (int)__umulhi((unsigned)e, a.magic_np), pl = e - c * a.nplanes;
      for (int sub = 0; sub < a.nsub; ++sub) {
        const int py = sub >> 1, px = sub & 1;
        int gy, gx;
        if (mode == S1) { gy = y0 - 1; gx = c - 1; }
        else if (mode == S2_FWD) { gy = 2 * (y0 - 1) + py; gx = 2 * (c - 1) + px; }
        else { gy = y0; gx = c; }
        const bool okx = gx >= 0 && gx < op.Win;
        const bf16* src = xb + (((size_t)n * op.Hin + gy) * op.Win + (okx ? gx : 0)) * op.x.pitch + op.x.coff + pl * 8;
        const size_t sstep = (size_t)ystep * op.Win * op.x.pitch;
        uint32_t dst = sa + (sub * a.nplanes + pl) * a.PB + c * 16;
        for (int r = 0; r < trows; ++r, gy += ystep, src += sstep, dst += PW * 16) {
          const bool ok = okx && gy >= 0 && gy < op.Hin;
          cp_async16(dst, ok ? src : xb, ok ? 16 : 0);
        }
      }
    }
  };
  {
    // weights: planes [tap*nplanes + pl], rows co0..co0+NT of CoP, 16 B per row
    const int wrows = 9 * a.nplanes * a.NT;
    const uint4* wsrc = reinterpret_cast<const uint4*>(op.w_umma);
    for (int i = tid; i < wrows; i += UMMA_THREADS) {
      const int tp = i / a.NT, row = i - tp * a.NT;
      cp_async16(sw + i * 16, wsrc + (size_t)tp * a.CoP + co0 + row, 16);
    }
  }
  load_tile(blockIdx.x, sa0);
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int row_in_tile = (warp & 3) * 32 + (tid & 31), ehalf = warp >> 2;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int ncls = (mode == S2_DGRAD) ? 4 : 1;
  const uint32_t idesc = instr_desc(128, a.NT);
  const int kcs = a.nplanes >> 1;
  const uint32_t wplane = a.NT * 16;

  int it = 0;
  for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++it) {
    const uint32_t sa = sa0 + ((a.nbuf == 2) ? (it & 1) * a.a_bytes : 0u);
    // ---- MMA issue, spread over warps: unit u = (M-tile, parity class) is issued by lane 0 of warp
    //      u % nissue; descriptors are built once per tap and advanced by plain adds per K-step
    if ((tid & 31) == 0 && warp < nissue) {
      const uint64_t kstepA = (uint64_t)((2 * a.PB) >> 4), kstepB = (uint64_t)((2 * wplane) >> 4);
      for (int u = warp; u < a.n_mt * ncls; u += nissue) {
        const int mt = u / ncls, cls = u - mt * ncls;
        const int py = cls >> 1, px = cls & 1;
        uint32_t acc = 0;
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          int sub = 0, shift;
          if (mode == S1) {
            shift = ky * PW + kx;
          } else if (mode == S2_FWD) {
            sub = ((ky == 1) ? 0 : 2) + ((kx == 1) ? 0 : 1);
            shift = ((ky == 0) ? 0 : 1) * PW + ((kx == 0) ? 0 : 1);
          } else {
            // output parity class (py,px): taps ky in {1} (py=0) or {0,2} (py=1); ky=0 reads dy one row below
            if (((py == 0) != (ky == 1)) || ((px == 0) != (kx == 1))) continue;
            shift = ((ky == 0) ? 1 : 0) * PW + ((kx == 0) ? 1 : 0);
          }
          uint64_t ad = smem_desc(sa + sub * a.nplanes * a.PB + (mt * 128 + shift) * 16, a.PB, 128);
          uint64_t bd = smem_desc(sw + tap * a.nplanes * wplane, wplane, 128);
          for (int kc = 0; kc < kcs; ++kc, ad += kstepA, bd += kstepB) {
            umma_f16(tmem + u * a.NT, ad, bd, idesc, acc);
            acc = 1;
          }
        }
      }
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    // ---- stage the next tile: into the other buffer while the MMAs run (nbuf == 2), or into the
    //      same buffer once they are done (nbuf == 1); either way it overlaps the epilogue
    const int next = tile + gridDim.x;
    if (a.nbuf == 2 && next < a.tiles_total) load_tile(next, sa0 + ((it + 1) & 1) * a.a_bytes);
    mbar_wait(smem_u32(&mbar), it & 1);
    tc_fence_after();
    if (a.nbuf == 1 && next < a.tiles_total) load_tile(next, sa0);

    // ---- epilogue: thread t owns accumulator row t of every M-tile
    const int n = tile / a.tiles_per_img;
    const int y0 = (tile - n * a.tiles_per_img) * a.TH;
    const int rows = min(a.TH, a.Ht - y0);
    int piece = 0;  // 16-column pieces alternate between the two warps of a lane quarter
    for (int mt = 0; mt < a.n_mt; ++mt) {
      const int q = mt * 128 + row_in_tile;
      const int r = (int)__umulhi((unsigned)q, a.magic_pw), c = q - r * PW;
      const bool valid = (r < rows) && (c < a.Wt);
      for (int cls = 0; cls < ncls; ++cls) {
        int yo = y0 + r, xo = c;
        if (mode == S2_DGRAD) { yo = 2 * yo + (cls >> 1); xo = 2 * xo + (cls & 1); }
        for (int nc = 0; nc < a.NT; nc += 16, ++piece) {
          if ((piece & 1) != ehalf) continue;
          float v[16];
          tmem_ld16(tmem + lane_base + (mt * ncls + cls) * a.NT + nc, v);
          if (valid) epilogue16(op, v, n, yo, xo, co0 + nc);
        }
      }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

bool plan(const ConvOp& op, UmmaArgs& a) {
  if (!op.w_umma || !op.x.bf) return false;
  if (op.Ci % 16 || op.Co % 16 || op.Co > 256) return false;
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
  if (op.shuffle == SHUF_NONE && !aligned(op.y)) return false;
  const int Ht = (mode == S2_DGRAD) ? op.Hin : op.Hout, Wt = (mode == S2_DGRAD) ? op.Win : op.Wout;
  const int PW = (mode == S1) ? Wt + 2 : Wt + 1;
  const int nplanes = op.Ci / 8, nsub = (mode == S2_FWD) ? 4 : 1, ncls = (mode == S2_DGRAD) ? 4 : 1;
  const int halo = (mode == S1) ? 2 * PW + 2 : PW + 1;
  // output-channel chunk per CTA: the largest divisor of Co whose weight image fits comfortably
  int NT = 0;
  for (int cand : {256, 128, 64, 32, 16}) {
    if (cand > op.Co || op.Co % cand) continue;
    if ((size_t)9 * op.Ci * cand * 2 > 112 * 1024) continue;
    if (ncls * cand > 512) continue;
    NT = cand;
    break;
  }
  if (NT == 0) return false;
  const size_t wbytes = (size_t)9 * op.Ci * NT * 2;
  double best = -1.0;
  int bestTH = 0, best_mt = 0;
  for (int TH = 1; TH <= Ht; ++TH) {
    const int span = TH * PW - (PW - Wt);
    const int n_mt = (span + 127) / 128;
    if (n_mt > 4 || n_mt * ncls * NT > 512) break;
    const int pb_pos = ((n_mt * 128 + halo) + 7) & ~7;
    const size_t smem = (size_t)nsub * nplanes * pb_pos * 16 + wbytes;
    if (smem > (size_t)MAX_SMEM) break;
    if ((TH + 2) * PW * nplanes >= 65536) break;  // loader index arithmetic is exact below 2^16
    const int ntiles = (Ht + TH - 1) / TH;
    // score = MMA row efficiency x halo re-read efficiency x chip fill (tiny tiles pay ~us of fixed
    // per-tile latency and re-stage their halo rows; too few tiles leave SMs idle)
    const double row_eff = (double)(Ht * Wt) / ((double)ntiles * n_mt * 128);
    const double halo_eff = (double)TH / (TH + ((mode == S1) ? 2 : 1));
    const double fill = std::min(1.0, (double)ntiles * op.B / 296.0);
    const double eff = row_eff * halo_eff * fill;
    if (eff > best + 1e-9) { best = eff; bestTH = TH; best_mt = n_mt; }
  }
  if (bestTH == 0) return false;
  a.op = op;
  a.mode = mode; a.Ht = Ht; a.Wt = Wt;
  a.TH = bestTH; a.PW = PW; a.n_mt = best_mt; a.NT = NT; a.nplanes = nplanes; a.nsub = nsub;
  a.CoP = round_up(op.Co, 16);
  a.PB = (((best_mt * 128 + halo) + 7) & ~7) * 16;
  a.tiles_per_img = (Ht + bestTH - 1) / bestTH;
  int cols = best_mt * ncls * NT, pc = 32;
  while (pc < cols) pc <<= 1;
  a.tmem_cols = pc;
  a.a_bytes = (unsigned)(nsub * nplanes * a.PB);
  // a second input buffer only when it is cheap: occupancy (co-resident CTAs) hides more latency
  a.nbuf = (2 * (size_t)a.a_bytes + wbytes <= 56 * 1024) ? 2 : 1;
  a.w_off = a.nbuf * a.a_bytes;
  a.tiles_total = a.tiles_per_img * op.B;
  a.magic_np = (unsigned)((0x100000000ULL + nplanes - 1) / nplanes);
  a.magic_pw = (unsigned)((0x100000000ULL + PW - 1) / PW);
  return true;
}

}  // namespace

bool umma_supported(const ConvOp& op) {
  UmmaArgs a;
  return plan(op, a);
}

int conv_umma(const ConvOp& op, cudaStream_t st) {
  UmmaArgs a;
  if (!plan(op, a)) { set_error("conv_umma: unsupported shape"); return DG_ERR_INVALID; }
  static bool attr_set = false;
  if (!attr_set) {
    DG_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    attr_set = true;
  }
  const size_t smem = (size_t)a.w_off + (size_t)9 * op.Ci * a.NT * 2;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const double taps = op.transposed ? 2.25 : 9.0;
  Prof prof(PC_CONV_UMMA, 2.0 * total * op.Co * op.Ci * taps,
            (double)total * op.Co * (op.y.bf ? 2 : 4) + (double)op.B * op.Hin * op.Win * op.Ci * 2.0, st);
  const int n_chunks = op.Co / a.NT;
  int per_sm = (int)((size_t)(MAX_SMEM + 2048) / (smem + 1024));  // co-resident CTAs by shared memory
  per_sm = std::max(1, std::min(per_sm, std::min(4, 512 / a.tmem_cols)));
  int gx = (148 * per_sm + n_chunks - 1) / n_chunks;
  gx = std::max(1, std::min(gx, a.tiles_total));
  conv_umma_kernel<<<dim3(gx, n_chunks), UMMA_THREADS, smem, st>>>(a);
  DG_LAUNCH_CHECK();
  return 0;
}

// fp32 packed [tap][ci][CoP]  ->  bf16 planes [(tap*Ci/8 + ci/8)][CoP][8]   (same element offsets)
__global__ void pack_umma_kernel(const float* __restrict__ src, bf16* __restrict__ dst, const UmmaPackDesc* __restrict__ tab) {
  const UmmaPackDesc d = tab[blockIdx.y];
  const long long n = 9LL * d.Ci * d.CoP;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(e % d.CoP);
    const long long r = e / d.CoP;
    const int ci = (int)(r % d.Ci), tap = (int)(r / d.Ci);
    const long long o = (((long long)tap * (d.Ci >> 3) + (ci >> 3)) * d.CoP + co) * 8 + (ci & 7);
    dst[d.off + o] = __float2bfloat16_rn(src[d.off + e]);
  }
}

int pack_umma(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, cudaStream_t st) {
  if (n == 0) return 0;
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  pack_umma_kernel<<<dim3(bx, n), 256, 0, st>>>(packed, (bf16*)dst_bf16, table_dev);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg

extern "C" int dg_has_tcgen05(void) { return 1; }
