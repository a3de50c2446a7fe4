// tcgen05 / TMEM implicit-GEMM 3x3 convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// Data layout in shared memory ("planar"): the halo-padded input tile is stored as
//     [Ci/8 planes][positions][8 channels]      (16 bytes per position per plane)
// with positions = (TH+2) x (W+2) row-major.  This is the tcgen05 canonical NO-SWIZZLE K-major
// layout (core matrix = 8 consecutive positions x 16 B; SBO = 128 B; LBO = plane stride), so the A
// operand of tap (dy,dx) is the SAME tile at a start-address offset of (dy*(W+2)+dx)*16 bytes:
// the tile is loaded once and never re-gathered per tap.  Rows of the MMA are consecutive padded
// positions; the two pad columns per image row are computed and discarded in the epilogue.
// Weights are pre-packed in global memory in the matching B layout [tap*Ci/8 planes][Co][8].
// The accumulator lives in TMEM (lane = output position, column = output channel) and is read back
// with tcgen05.ld for the fused epilogue (bias, residual scale + skips, LeakyReLU / mask, bf16
// pack, pixel-(un)shuffle addressing, channel-offset store into the dense-block concat buffer).
#include "dg_common.cuh"

namespace dg {

namespace {

constexpr int UMMA_THREADS = 128;
constexpr int MAX_SMEM = 227 * 1024 - 2048;  // dynamic limit (static smem of the kernel comes on top)

struct UmmaArgs {
  ConvOp op;
  int TH, PW, n_mt, PB, tiles_per_img, NT, tmem_cols, nplanes;
  unsigned w_off;  // byte offset of the weight image in dynamic smem
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a descriptor / pipeline bug must trap, not hang the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
// instruction descriptor for kind::f16, bf16 x bf16 -> fp32, K-major A and B
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive channels of a view at element index i (i % 8 == 0 guaranteed by the host checks)
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

__global__ void __launch_bounds__(UMMA_THREADS) conv_umma_kernel(const UmmaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const ConvOp& op = a.op;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.x / a.tiles_per_img;
  const int y0 = (blockIdx.x % a.tiles_per_img) * a.TH;
  const int rows = min(a.TH, op.Hout - y0);
  const int H = op.Hin, W = op.Win, PW = a.PW;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) mbar_init(smem_u32(&mbar), 1);

  // ---- stage the halo-padded input tile (zero fill = conv padding) and the weight image
  const uint32_t sa = smem_u32(smem);
  {
    const int npos = (a.TH + 2) * PW;
    const int total = npos * a.nplanes;
    const bf16* xb = (const bf16*)op.x.p;
    for (int i = tid; i < total; i += UMMA_THREADS) {
      const int pos = i / a.nplanes, pl = i - pos * a.nplanes;
      const int r = pos / PW, c = pos - r * PW;
      const int gy = y0 - 1 + r, gx = c - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const bf16* src = ok ? xb + (((size_t)n * H + gy) * W + gx) * op.x.pitch + op.x.coff + pl * 8 : xb;
      cp_async16(sa + pl * a.PB + pos * 16, src, ok ? 16 : 0);
    }
    const int wchunks = 9 * a.nplanes * a.NT;
    const uint4* wsrc = reinterpret_cast<const uint4*>(op.w_umma);
    const uint32_t sw = sa + a.w_off;
    for (int i = tid; i < wchunks; i += UMMA_THREADS) cp_async16(sw + i * 16, wsrc + i, 16);
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // ---- one thread issues every MMA of the tile: M-tiles x 9 taps x Ci/16 k-steps
  if (tid == 0) {
    const uint32_t idesc = instr_desc(128, a.NT);
    const uint32_t sw = sa + a.w_off;
    const int kcs = a.nplanes >> 1;
    for (int mt = 0; mt < a.n_mt; ++mt) {
      uint32_t first = 0;
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap - 3 * dy;
        const uint32_t a0 = sa + (mt * 128 + dy * PW + dx) * 16;
        const uint32_t b0 = sw + (tap * a.nplanes) * a.NT * 16;
        for (int kc = 0; kc < kcs; ++kc) {
          const uint64_t ad = smem_desc(a0 + 2 * kc * a.PB, a.PB, 128);
          const uint64_t bd = smem_desc(b0 + 2 * kc * a.NT * 16, a.NT * 16, 128);
          umma_f16(tmem + mt * a.NT, ad, bd, idesc, first);
          first = 1;
        }
      }
    }
    umma_commit(smem_u32(&mbar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after();

  // ---- epilogue: thread t owns accumulator row t of every M-tile
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int mt = 0; mt < a.n_mt; ++mt) {
    const int q = mt * 128 + tid;
    const int r = q / PW, c = q - r * PW;
    const bool valid = (r < rows) && (c < W);
    const int yo = y0 + r, xo = c;
    const size_t p = ((size_t)n * op.Hout + yo) * op.Wout + xo;
    for (int nc = 0; nc < a.NT; nc += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + mt * a.NT + nc, v);
      if (!valid || nc >= op.Co) continue;
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
      } else if (op.shuffle == SHUF_PIXEL) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = nc + j, cc = co >> 2, ii = (co >> 1) & 1, jj = co & 1;
          const size_t qq = ((size_t)n * (2 * op.Hout) + 2 * yo + ii) * (2 * op.Wout) + 2 * xo + jj;
          st1(op.y, qq * op.y.pitch + op.y.coff + cc, v[j]);
        }
      } else {
        const size_t qq = ((size_t)n * (op.Hout >> 1) + (yo >> 1)) * (op.Wout >> 1) + (xo >> 1);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          st1(op.y, qq * op.y.pitch + op.y.coff + 4 * (nc + j) + 2 * (yo & 1) + (xo & 1), v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)a.tmem_cols) : "memory");
}

bool plan(const ConvOp& op, UmmaArgs& a) {
  if (!op.w_umma || !op.x.bf || op.stride != 1 || op.transposed) return false;
  if (op.Ci % 16 || op.Co % 16 || op.Co > 256 || op.Hin != op.Hout || op.Win != op.Wout) return false;
  if (op.x.pitch % 8 || op.x.coff % 8) return false;
  auto aligned = [](const TV& t) {
    if (!t.p) return true;
    return t.bf ? (t.pitch % 8 == 0 && t.coff % 8 == 0) : (t.pitch % 4 == 0 && t.coff % 4 == 0);
  };
  if (!aligned(op.r1) || !aligned(op.r2) || !aligned(op.mask)) return false;
  if (op.shuffle == SHUF_NONE && !aligned(op.y)) return false;
  const int W = op.Win, H = op.Hin, PW = W + 2, NT = op.Co, nplanes = op.Ci / 8;
  const size_t wbytes = (size_t)9 * op.Ci * NT * 2;
  double best = -1.0;
  int bestTH = 0, best_mt = 0;
  for (int TH = 1; TH <= H; ++TH) {
    const int span = TH * PW - 2;
    const int n_mt = (span + 127) / 128;
    if (n_mt > 4 || n_mt * NT > 512) break;
    const int pb_pos = ((n_mt * 128 + 2 * PW + 2) + 7) & ~7;
    const size_t smem = (size_t)nplanes * pb_pos * 16 + wbytes;
    if (smem > (size_t)MAX_SMEM) break;
    const int ntiles = (H + TH - 1) / TH;
    const double eff = (double)(H * W) / ((double)ntiles * n_mt * 128);
    if (eff > best + 1e-9) { best = eff; bestTH = TH; best_mt = n_mt; }
  }
  if (bestTH == 0) return false;
  a.op = op;
  a.TH = bestTH; a.PW = PW; a.n_mt = best_mt; a.NT = NT; a.nplanes = nplanes;
  a.PB = (((best_mt * 128 + 2 * PW + 2) + 7) & ~7) * 16;
  a.tiles_per_img = (H + bestTH - 1) / bestTH;
  int cols = best_mt * NT, pc = 32;
  while (pc < cols) pc <<= 1;
  a.tmem_cols = pc;
  a.w_off = (unsigned)(nplanes * a.PB);
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
  Prof prof(PC_CONV_UMMA, 2.0 * total * op.Co * op.Ci * 9.0,
            (double)total * op.Co * (op.y.bf ? 2 : 4) + (double)total * op.Ci * 2.0, st);
  conv_umma_kernel<<<op.B * a.tiles_per_img, UMMA_THREADS, smem, st>>>(a);
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
