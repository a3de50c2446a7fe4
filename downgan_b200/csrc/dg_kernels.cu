// CUDA-core kernels of libdowngan_b200.so: direct 3x3 convolutions (the fp32
// parity mode and the HBM-bound skinny layers of the bf16 mode), the critic's
// linear layers, weight packing, gradient-penalty / loss reductions and Adam.
// Everything accumulates in fp32.  sm_100a only.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "dg_common.cuh"

namespace dg {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------
// profiling
// ---------------------------------------------------------------------------
bool g_prof_on = false;
namespace {
struct ProfRec { int cls; double flops, bytes; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;
}  // namespace
Prof::Prof(int cls, double flops, double bytes, cudaStream_t s) : st(s) {
  if (!g_prof_on) return;
  ProfRec r; r.cls = cls; r.flops = flops; r.bytes = bytes;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  idx = (int)g_prof.size();
  g_prof.push_back(r);
}
Prof::~Prof() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].b, st);
}

// ---------------------------------------------------------------------------
// typed access through a TV
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ldv(const void* p, int bf, size_t i) {
  return bf ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void stv(void* p, int bf, size_t i, float v) {
  if (bf) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}
__device__ __forceinline__ float lrelu_d(float a, float slope) { return a > 0.f ? 1.f : slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {  // sh: >= 32 floats
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  return v;  // valid in warp 0
}

// ---------------------------------------------------------------------------
// direct 3x3 convolution, one output pixel x 16 output channels per thread
// ---------------------------------------------------------------------------
constexpr int CONV_THREADS = 128;

__global__ void __launch_bounds__(CONV_THREADS) conv_direct_kernel(ConvOp op) {
  __shared__ __align__(16) float sw[9][16][16];
  const int CoP = (op.Co + 15) & ~15;
  const int co0 = blockIdx.y * 16;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const long long p = (long long)blockIdx.x * CONV_THREADS + threadIdx.x;
  const bool valid = p < total;
  int n = 0, yo = 0, xo = 0;
  if (valid) {
    xo = (int)(p % op.Wout);
    long long t = p / op.Wout;
    yo = (int)(t % op.Hout);
    n = (int)(t / op.Hout);
  }
  // per-tap input pixel (or -1)
  long long off[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    int yi, xi;
    bool ok = valid;
    if (op.transposed) {
      const int ty = yo + 1 - ky, tx = xo + 1 - kx;
      ok = ok && ty >= 0 && tx >= 0 && !(ty & 1) && !(tx & 1);
      yi = ty >> 1; xi = tx >> 1;
      ok = ok && yi < op.Hin && xi < op.Win;
    } else {
      yi = yo * op.stride + ky - 1; xi = xo * op.stride + kx - 1;
      ok = ok && yi >= 0 && xi >= 0 && yi < op.Hin && xi < op.Win;
    }
    off[tap] = ok ? (((long long)n * op.Hin + yi) * op.Win + xi) * op.x.pitch + op.x.coff : -1;
  }
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;

  for (int ci0 = 0; ci0 < op.Ci; ci0 += 16) {
    const int cn = min(16, op.Ci - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * cn * 16; i += CONV_THREADS) {
      const int tap = i / (cn * 16), r = i % (cn * 16), ci = r >> 4, j = r & 15;
      sw[tap][ci][j] = op.w[((size_t)tap * op.Ci + ci0 + ci) * CoP + co0 + j];
    }
    __syncthreads();
    if (!valid) continue;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (off[tap] < 0) continue;
      const size_t base = (size_t)off[tap] + ci0;
      if (cn == 16) {
        float xs[16];
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) xs[ci] = ldv(op.x.p, op.x.bf, base + ci);
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) {
          const float4* wr = reinterpret_cast<const float4*>(&sw[tap][ci][0]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = wr[q];
            acc[4 * q + 0] = fmaf(xs[ci], wv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xs[ci], wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xs[ci], wv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xs[ci], wv.w, acc[4 * q + 3]);
          }
        }
      } else {
        for (int ci = 0; ci < cn; ++ci) {
          const float xv = ldv(op.x.p, op.x.bf, base + ci);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(xv, sw[tap][ci][j], acc[j]);
        }
      }
    }
  }
  if (!valid) return;
  // epilogue
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int co = co0 + j;
    if (co >= op.Co) break;
    float v = acc[j];
    if (op.bias) v += op.bias[co];
    v *= op.s_acc;
    if (op.r1.p) v = fmaf(op.s1, ldv(op.r1.p, op.r1.bf, (size_t)p * op.r1.pitch + op.r1.coff + co), v);
    if (op.r2.p) v = fmaf(op.s2, ldv(op.r2.p, op.r2.bf, (size_t)p * op.r2.pitch + op.r2.coff + co), v);
    if (op.act == ACT_LRELU) v = v > 0.f ? v : v * op.slope;
    else if (op.act == ACT_MASK)
      v *= lrelu_d(ldv(op.mask.p, op.mask.bf, (size_t)p * op.mask.pitch + op.mask.coff + co), op.slope);
    size_t idx;
    if (op.shuffle == SHUF_NONE) {
      idx = (size_t)p * op.y.pitch + op.y.coff + co;
    } else if (op.shuffle == SHUF_PIXEL) {
      // out[n, 2y+i, 2x+j, c] = v[n, y, x, 4c+2i+j]   (nn.PixelShuffle(2), generator.py:73)
      const int c = co >> 2, i = (co >> 1) & 1, jj = co & 1;
      const size_t q = ((size_t)n * (2 * op.Hout) + 2 * yo + i) * (2 * op.Wout) + 2 * xo + jj;
      idx = q * op.y.pitch + op.y.coff + c;
    } else {
      // inverse addressing for the data-gradient of a pixel-shuffled activation
      const size_t q = ((size_t)n * (op.Hout >> 1) + (yo >> 1)) * (op.Wout >> 1) + (xo >> 1);
      idx = q * op.y.pitch + op.y.coff + (2 * (yo & 1) + (xo & 1)) * op.Co + co;  // (i,j)-major, see PackDesc::ps
    }
    stv(op.y.p, op.y.bf, idx, v);
  }
}

int conv_direct(const ConvOp& op, cudaStream_t st) {
  DG_CHECK(op.x.p && op.y.p && op.w, "conv_direct: null tensor");
  DG_CHECK(op.stride == 1 || op.stride == 2, "conv_direct: stride %d", op.stride);
  const long long total = (long long)op.B * op.Hout * op.Wout;
  dim3 grid((unsigned)((total + CONV_THREADS - 1) / CONV_THREADS), (unsigned)((op.Co + 15) / 16));
  const double taps = op.transposed ? 2.25 : 9.0;
  Prof prof(PC_CONV_DIRECT, 2.0 * total * op.Co * op.Ci * taps,
            (double)total * op.Co * (op.y.bf ? 2 : 4) + (double)op.B * op.Hin * op.Win * op.Ci * (op.x.bf ? 2 : 4), st);
  conv_direct_kernel<<<grid, CONV_THREADS, 0, st>>>(op);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// direct weight gradient: thread = (tap, ci) x 16 output channels
// ---------------------------------------------------------------------------
constexpr int WG_THREADS = 144;
constexpr int WG_SUB = 16;

__global__ void __launch_bounds__(WG_THREADS) wgrad_direct_kernel(WgradOp op, int pix_per_block) {
  __shared__ __align__(16) float sdy[WG_SUB][16];
  const int CoP = (op.Co + 15) & ~15;
  const int tap = threadIdx.x >> 4, cil = threadIdx.x & 15;
  const int ci = blockIdx.y * 16 + cil;
  const int co0 = blockIdx.z * 16;
  const int ky = tap / 3, kx = tap % 3;
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = min(total, p0 + pix_per_block);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (long long ps = p0; ps < p1; ps += WG_SUB) {
    __syncthreads();
    for (int i = threadIdx.x; i < WG_SUB * 16; i += WG_THREADS) {
      const int q = i >> 4, j = i & 15;
      const long long p = ps + q;
      float v = 0.f;
      if (p < p1 && co0 + j < op.Co) v = ldv(op.dy.p, op.dy.bf, (size_t)p * op.dy.pitch + op.dy.coff + co0 + j);
      sdy[q][j] = v;
    }
    __syncthreads();
    if (ci >= op.Ci) continue;
    const int qn = (int)min((long long)WG_SUB, p1 - ps);
    for (int q = 0; q < qn; ++q) {
      const long long p = ps + q;
      const int xo = (int)(p % op.Wout);
      const long long t = p / op.Wout;
      const int yo = (int)(t % op.Hout);
      const int n = (int)(t / op.Hout);
      const int yi = yo * op.stride + ky - 1, xi = xo * op.stride + kx - 1;
      if (yi < 0 || xi < 0 || yi >= op.Hin || xi >= op.Win) continue;
      const float xv = ldv(op.x.p, op.x.bf, (((size_t)n * op.Hin + yi) * op.Win + xi) * op.x.pitch + op.x.coff + ci);
      const float4* dr = reinterpret_cast<const float4*>(&sdy[q][0]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 d = dr[k];
        acc[4 * k + 0] = fmaf(xv, d.x, acc[4 * k + 0]);
        acc[4 * k + 1] = fmaf(xv, d.y, acc[4 * k + 1]);
        acc[4 * k + 2] = fmaf(xv, d.z, acc[4 * k + 2]);
        acc[4 * k + 3] = fmaf(xv, d.w, acc[4 * k + 3]);
      }
    }
  }
  if (ci >= op.Ci) return;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (co0 + j < op.Co) atomicAdd(&op.dw[((size_t)tap * op.Ci + ci) * CoP + co0 + j], acc[j]);
}

int wgrad_direct(const WgradOp& op, cudaStream_t st) {
  DG_CHECK(op.x.p && op.dy.p && op.dw, "wgrad_direct: null tensor");
  const long long total = (long long)op.B * op.Hout * op.Wout;
  const int cic = (op.Ci + 15) / 16, coc = (op.Co + 15) / 16;
  // aim for ~8 blocks per SM overall
  long long want = (148LL * 8 + cic * coc - 1) / (cic * coc);
  long long ppb = (total + want - 1) / want;
  ppb = ((ppb + WG_SUB - 1) / WG_SUB) * WG_SUB;
  if (ppb < 64) ppb = 64;
  if (ppb > 4096) ppb = 4096;
  dim3 grid((unsigned)((total + ppb - 1) / ppb), (unsigned)cic, (unsigned)coc);
  Prof prof(PC_WGRAD_DIRECT, 2.0 * total * op.Co * op.Ci * 9.0,
            (double)total * op.Co * (op.dy.bf ? 2 : 4) + (double)op.B * op.Hin * op.Win * op.Ci * (op.x.bf ? 2 : 4), st);
  wgrad_direct_kernel<<<grid, WG_THREADS, 0, st>>>(op, (int)ppb);
  DG_LAUNCH_CHECK();
  if (op.dbias) DG_TRY(colsum(op.dy, (size_t)total, op.Co, op.dbias, st));
  return 0;
}

// column sums over pixels (bias gradients).  Block partials are combined in fp64 by the last
// block to finish: critic bias gradients are sums of cancelling real/fake halves, and an fp32
// atomic chain loses ~1e-3 of the result there.  out[c] += sum.
// The partial buffer and the completion counter exist once per SLOT: launches on the library's side streams
// (register_side_stream) use their own slot, so a column sum there may overlap one on the caller's stream.
constexpr int COLSUM_MAX_BLOCKS = 1024, COLSUM_MAX_C = 256, COLSUM_SLOTS = 3;
__device__ float g_colsum_parts[COLSUM_SLOTS][COLSUM_MAX_BLOCKS * COLSUM_MAX_C];
__device__ unsigned int g_colsum_dones[COLSUM_SLOTS] = {0, 0, 0};
// every handle's side stream is registered (several handles may share a slot: their side streams never carry column
// sums at the same time - a call joins its side stream before it returns, and the deferred look-ahead chain has none)
static std::vector<std::pair<cudaStream_t, int>> g_side_streams;
void register_side_stream(cudaStream_t st, int slot) {
  unregister_side_stream(st);
  if (st && slot >= 1 && slot < COLSUM_SLOTS) g_side_streams.emplace_back(st, slot);
}
void unregister_side_stream(cudaStream_t st) {
  for (size_t i = 0; i < g_side_streams.size();)
    if (g_side_streams[i].first == st) g_side_streams.erase(g_side_streams.begin() + i);
    else ++i;
}
static int colsum_slot(cudaStream_t st) {
  for (const auto& e : g_side_streams)
    if (e.first == st) return e.second;
  return 0;
}

// Last-block combine of the per-block partial column sums, by all 256 threads of the block: thread t owns channel t % C' and every
// (256 / C')-th block (C' = C rounded up to a power of two), fp64 partials are then added in a fixed order - deterministic, and
// nblocks / (256 / C') dependent-free loads per thread instead of nblocks serial ones on C threads (the 888-block sums of the
// generator's tail layers spent 30 - 50 us in that loop).
__device__ __forceinline__ void colsum_combine(const float* __restrict__ part, unsigned nblocks, int C, float* __restrict__ out, int t) {
  __shared__ double shd[256];
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int Cc = min(C - c0, 256);
    int Cp = 1;
    while (Cp < Cc) Cp <<= 1;
    const int ngrp = 256 / Cp, c = t % Cp, grp = t / Cp;
    double acc = 0.0;
    if (c < Cc)
      for (unsigned b = grp; b < nblocks; b += ngrp) acc += (double)__ldcg(part + (size_t)b * C + c0 + c);
    shd[t] = acc;
    __syncthreads();
    if (grp == 0 && c < Cc) {
      double tot = 0.0;
      for (int g2 = 0; g2 < ngrp; ++g2) tot += shd[g2 * Cp + c];
      out[c0 + c] += (float)tot;
    }
    __syncthreads();
  }
}

__global__ void colsum_kernel(TV dy, size_t pixels, int C, float* out, size_t pix_per_block, int slot) {
  __shared__ float sh[8][33];
  __shared__ bool last;
  float* const g_colsum_part = g_colsum_parts[slot];
  unsigned int& g_colsum_done = g_colsum_dones[slot];
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  const size_t p1 = min(pixels, p0 + pix_per_block);
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + threadIdx.x;
    float s = 0.f;
    if (c < C)
      for (size_t p = p0 + threadIdx.y; p < p1; p += 8) s += ldv(dy.p, dy.bf, p * dy.pitch + dy.coff + c);
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
      g_colsum_part[(size_t)blockIdx.x * C + c] = t;
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0 && threadIdx.y == 0) last = (atomicAdd(&g_colsum_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  colsum_combine(g_colsum_part, gridDim.x, C, out, threadIdx.y * 32 + threadIdx.x);
  if (threadIdx.x == 0 && threadIdx.y == 0) g_colsum_done = 0;
}
// bf16 rows with C in {16,32,64,128,256}: 16-byte loads (8 channels per thread), four pixels in flight per
// thread; same fp64 last-block combine as colsum_kernel.
__global__ void __launch_bounds__(256) colsum_vec_kernel(TV dy, size_t pixels, int C, float* out, size_t pix_per_block, int slot) {
  __shared__ float sh[256 * 8];
  __shared__ bool last;
  float* const g_colsum_part = g_colsum_parts[slot];
  unsigned int& g_colsum_done = g_colsum_dones[slot];
  const int cpp = C >> 3;              // 16-byte chunks per pixel (a power of two <= 32)
  const int chunk = threadIdx.x & (cpp - 1), pl = threadIdx.x / cpp, npl = 256 / cpp;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  const size_t p1 = min(pixels, p0 + pix_per_block);
  const bf16* base = (const bf16*)dy.p + dy.coff + chunk * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto add8 = [&](const uint4& q) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[2 * k] += __uint_as_float(w[k] << 16);
      acc[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
    }
  };
  size_t p = p0 + pl;
  for (; p + 7 * (size_t)npl < p1; p += 8 * (size_t)npl) {  // eight 16-byte loads in flight per thread
    uint4 q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) q[u] = *reinterpret_cast<const uint4*>(base + (p + (size_t)u * npl) * dy.pitch);
#pragma unroll
    for (int u = 0; u < 8; ++u) add8(q[u]);
  }
  for (; p < p1; p += npl) add8(*reinterpret_cast<const uint4*>(base + p * dy.pitch));
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if ((int)threadIdx.x < C) {
    const int c = threadIdx.x, ch = c >> 3, j = c & 7;
    float t = 0.f;
    for (int q = 0; q < npl; ++q) t += sh[(q * cpp + ch) * 8 + j];
    g_colsum_part[(size_t)blockIdx.x * C + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&g_colsum_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  colsum_combine(g_colsum_part, gridDim.x, C, out, threadIdx.x);
  if (threadIdx.x == 0) g_colsum_done = 0;
}
// fp32 rows with C in {1, 2, 4} and no channel offset (the 2-channel fine fields: conv3.2's bias gradient sums dL/dfake): the
// generic kernel would use C of its 32 channel lanes; here a thread reads float4 = 4 / C whole pixels per load.
__global__ void __launch_bounds__(256) colsum_small_kernel(const float* __restrict__ x, size_t n_elems, int C, float* out,
                                                           size_t elems_per_block, int slot) {
  __shared__ float sh[256 * 4];
  __shared__ bool last;
  float* const g_colsum_part = g_colsum_parts[slot];
  unsigned int& g_colsum_done = g_colsum_dones[slot];
  const size_t e0 = (size_t)blockIdx.x * elems_per_block, e1 = min(n_elems, e0 + elems_per_block);  // multiples of 4
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t i = (e0 >> 2) + threadIdx.x; i < (e1 >> 2); i += 256) {
    const float4 v = __ldg(x4 + i);
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sh[threadIdx.x * 4 + j] = acc[j];
  __syncthreads();
  if ((int)threadIdx.x < C) {  // element j of a float4 belongs to channel j % C
    float t = 0.f;
    for (int q = 0; q < 256; ++q)
      for (int j = threadIdx.x; j < 4; j += C) t += sh[q * 4 + j];
    g_colsum_part[(size_t)blockIdx.x * C + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&g_colsum_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  colsum_combine(g_colsum_part, gridDim.x, C, out, threadIdx.x);
  if (threadIdx.x == 0) g_colsum_done = 0;
}
int colsum(TV dy, size_t pixels, int C, float* out, cudaStream_t st) {
  if (ablate(10)) return 0;
  DG_CHECK(C <= COLSUM_MAX_C, "colsum: %d channels > %d", C, COLSUM_MAX_C);
  if (!dy.bf && (C == 1 || C == 2 || C == 4) && dy.pitch == C && dy.coff == 0 && ((uintptr_t)dy.p & 15) == 0 && ((pixels * C) & 3) == 0 &&
      pixels >= 4096) {
    const size_t n = pixels * C;
    size_t epb = ((n + 148 * 4 - 1) / (148 * 4) + 1023) & ~(size_t)1023;
    const unsigned grid = (unsigned)((n + epb - 1) / epb);
    colsum_small_kernel<<<grid, 256, 0, st>>>((const float*)dy.p, n, C, out, epb, colsum_slot(st));
    DG_LAUNCH_CHECK();
    return 0;
  }
  const bool vec = dy.bf && (C == 16 || C == 32 || C == 64 || C == 128 || C == 256) && dy.pitch % 8 == 0 && dy.coff % 8 == 0 &&
                   pixels >= 4096;
  if (vec) {
    size_t ppb = (pixels + 148 * 6 - 1) / (148 * 6);
    if (ppb < 256) ppb = 256;
    const unsigned grid = (unsigned)((pixels + ppb - 1) / ppb);
    colsum_vec_kernel<<<grid, 256, 0, st>>>(dy, pixels, C, out, ppb, colsum_slot(st));
    DG_LAUNCH_CHECK();
    return 0;
  }
  size_t ppb = (pixels + 148 * 4 - 1) / (148 * 4);
  if (ppb < 64) ppb = 64;
  unsigned grid = (unsigned)((pixels + ppb - 1) / ppb);
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(dy, pixels, C, out, ppb, colsum_slot(st));
  DG_LAUNCH_CHECK();
  return 0;
}

// Bias gradients of every dense conv of the generator trunk in one launch (F = 16): the dz buffer of block b is
// (rows, 80) bf16 = [dz5 | dz4 | dz3 | dz2 | dz1]; out[(b*5 + k-1)*16 + c] += sum_rows D_b[row][(5-k)*16 + c].
__global__ void __launch_bounds__(320) colsum_dense_kernel(void* const* __restrict__ d_bufs, size_t rows, size_t rows_per_cta,
                                                            float* __restrict__ out) {
  __shared__ float sh[32][80];
  const int b = blockIdx.y;
  const bf16* D = (const bf16*)d_bufs[b];
  const int chunk = threadIdx.x % 10, rl = threadIdx.x / 10;  // 10 16-byte chunks per row, 32 rows in flight
  const size_t r0 = (size_t)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (size_t r = r0 + rl; r < r1; r += 32) {
    const uint4 q = *reinterpret_cast<const uint4*>(D + r * 80 + chunk * 8);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[2 * k] += __uint_as_float(w[k] << 16);
      acc[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[rl][chunk * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 80) {
    float t = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) t += sh[q][threadIdx.x];
    const int slice = threadIdx.x >> 4, c = threadIdx.x & 15;  // slice t holds dz_{5-t}
    atomicAdd(out + ((size_t)b * 5 + (4 - slice)) * 16 + c, t);
  }
}
int colsum_dense_blocks(void* const* d_bufs_dev, int n_blocks, size_t rows, float* out, cudaStream_t st) {
  if (ablate(10)) return 0;
  if (n_blocks <= 0) return 0;
  const unsigned gx = 8;
  const size_t rpc = (rows + gx - 1) / gx;
  colsum_dense_kernel<<<dim3(gx, (unsigned)n_blocks), 320, 0, st>>>(d_bufs_dev, rows, rpc, out);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// table-driven packing (one launch per network)
// ---------------------------------------------------------------------------
// One pass per table entry: OIHW fp32 -> packed fp32 [tap][row][CoP] and (d.umma) the bf16 tcgen05 B-operand
// image [(tap*R/8 + row/8)][CoP][8] at the same element offset of `udst`.  32-bit index arithmetic.
// A launch may carry TWO tables (forward and data-gradient images of one network): rows [0, n1) of the grid's y dimension
// use (dst, udst, tab), rows [n1, ..) use (dst2, udst2, tab2).
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, float* __restrict__ dst, bf16* __restrict__ udst,
                                                   const PackDesc* __restrict__ tab, int unpack, int n1,
                                                   float* __restrict__ dst2, bf16* __restrict__ udst2,
                                                   const PackDesc* __restrict__ tab2, bf16* __restrict__ igdst,
                                                   bf16* __restrict__ igdst2) {
  if ((int)blockIdx.y >= n1) { dst = dst2; udst = udst2; tab = tab2 - n1; igdst = igdst2; }
  const PackDesc d = tab[blockIdx.y];
  if (d.mode == 4) {
    // linear weight (Co = N rows, Ci = K cols) with the NCHW -> NHWC column permutation (K = C*HW, C = slice_off): per row a
    // (C x HW) -> (HW x C) transpose through a 32 x 33 shared tile, so both the reads and the writes are coalesced
    __shared__ float tile[32][33];
    const unsigned C = (unsigned)d.slice_off, K = (unsigned)d.Ci, HW = K / C;
    const unsigned tc = (C + 31) / 32, th = (HW + 31) / 32, per_row = tc * th, total = (unsigned)d.Co * per_row;
    const float* s_base = src + (unpack ? d.dst_off : d.src_off);
    float* d_base = dst + (unpack ? d.src_off : d.dst_off);
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (unsigned t = blockIdx.x; t < total; t += gridDim.x) {
      const unsigned j = t / per_row, r = t - j * per_row, c0 = (r / th) * 32, h0 = (r - (r / th) * th) * 32;
      // flat (reference) layout: [j][c][hw]; packed layout: [j][hw][c]
      if (!unpack) {
#pragma unroll
        for (unsigned i = ty; i < 32; i += 8)
          if (c0 + i < C && h0 + tx < HW) tile[i][tx] = s_base[(size_t)j * K + (c0 + i) * HW + h0 + tx];
        __syncthreads();
#pragma unroll
        for (unsigned i = ty; i < 32; i += 8)
          if (h0 + i < HW && c0 + tx < C) d_base[(size_t)j * K + (h0 + i) * C + c0 + tx] = tile[tx][i];
      } else {
#pragma unroll
        for (unsigned i = ty; i < 32; i += 8)
          if (h0 + i < HW && c0 + tx < C) tile[i][tx] = s_base[(size_t)j * K + (h0 + i) * C + c0 + tx];
        __syncthreads();
#pragma unroll
        for (unsigned i = ty; i < 32; i += 8)
          if (c0 + i < C && h0 + tx < HW) d_base[(size_t)j * K + (c0 + i) * HW + h0 + tx] = tile[tx][i];
      }
      __syncthreads();
    }
    return;
  }
  unsigned n;
  if (d.mode == 5) n = (unsigned)d.Co;
  else if (d.mode == 4) n = (unsigned)d.Co * (unsigned)d.Ci;
  else n = (unsigned)d.Co * (unsigned)d.Ci * 9u;
  const float* s_base = src + (unpack ? d.dst_off : d.src_off);
  float* d_base = dst + (unpack ? d.src_off : d.dst_off);
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    if (d.mode == 5) {  // bias copy (unpacking a pixel-shuffle layer's gradient: (i,j)-major columns back to 4c + 2i + j)
      d_base[e] = s_base[(unpack && d.ps) ? ps_perm(e, (unsigned)d.Co) : e];
      continue;
    }
    const unsigned r = e / 9u, tap = e - 9u * r;
    const unsigned co = r / (unsigned)d.Ci, ci = r - co * (unsigned)d.Ci;
    unsigned s_idx, tp, row, col, R, cols;
    if (d.mode == 3) {
      // slice of a dense-block weight W_j[co][slice_off+ci][tap] -> row dst_row_off+co, col ci, flipped tap
      s_idx = (co * (unsigned)d.src_ci_total + (unsigned)d.slice_off + ci) * 9u + tap;
      tp = 8u - tap; row = (unsigned)d.dst_row_off + co; col = ci; R = (unsigned)d.CoP /*rows_total*/; cols = (unsigned)d.dst_CoP;
    } else {
      s_idx = e;  // (co*Ci + ci)*9 + tap
      if (d.mode == 0) { tp = tap; row = ci; col = (unpack && d.ps) ? ps_perm(co, (unsigned)d.Co) : co; R = (unsigned)d.Ci; }
      else if (d.mode == 1) { tp = 8u - tap; row = d.ps ? ps_perm(co, (unsigned)d.Co) : co; col = ci; R = (unsigned)d.Co; }  // CoP = round16(Ci)
      else { tp = tap; row = co; col = ci; R = (unsigned)d.Co; }                          // mode 2
      cols = (unsigned)d.CoP;
    }
    const unsigned p_idx = (tp * R + row) * cols + col;
    if (unpack) {
      d_base[s_idx] = s_base[p_idx];
    } else {
      const float v = s_base[s_idx];
      d_base[p_idx] = v;
      if (d.umma) {
        const bf16 vb = __float2bfloat16_rn(v);
        udst[d.dst_off + (((tp * (R >> 3) + (row >> 3)) * cols + col) << 3) + (row & 7u)] = vb;
        // K-major image [tap][CoP][rows] of the streaming implicit-GEMM kernel (dg_umma_conv_ig.cu), same element offset
        if (igdst) igdst[d.dst_off + (tp * cols + col) * R + row] = vb;
      }
    }
  }
}
static inline int pack_blocks(int max_elems) {
  int bx = (max_elems + 1023) / 1024;  // ~4 elements per thread
  return bx < 1 ? 1 : (bx > 512 ? 512 : bx);
}
int pack_weights(const float* params, float* packed, void* packed_umma, const PackDesc* tab, int n, int max_elems, cudaStream_t st,
                 void* packed_ig) {
  if (ablate(9)) return 0;
  if (n == 0) return 0;
  pack_kernel<<<dim3(pack_blocks(max_elems), n), 256, 0, st>>>(params, packed, (bf16*)packed_umma, tab, 0, n, nullptr, nullptr, nullptr,
                                                               (bf16*)packed_ig, nullptr);
  DG_LAUNCH_CHECK();
  return 0;
}
// forward and data-gradient tables of one network in ONE launch
int pack_weights2(const float* params, float* packed, void* packed_umma, const PackDesc* tab, int n, int max_elems, float* packed2,
                  void* packed_umma2, const PackDesc* tab2, int n2, int max_elems2, cudaStream_t st, void* packed_ig, void* packed_ig2) {
  if (ablate(9)) return 0;
  if (n == 0 || n2 == 0) {
    DG_TRY(pack_weights(params, packed, packed_umma, tab, n, max_elems, st, packed_ig));
    return pack_weights(params, packed2, packed_umma2, tab2, n2, max_elems2, st, packed_ig2);
  }
  pack_kernel<<<dim3(pack_blocks(std::max(max_elems, max_elems2)), n + n2), 256, 0, st>>>(params, packed, (bf16*)packed_umma, tab, 0, n,
                                                                                       packed2, (bf16*)packed_umma2, tab2,
                                                                                       (bf16*)packed_ig, (bf16*)packed_ig2);
  DG_LAUNCH_CHECK();
  return 0;
}
int unpack_wgrads(const float* packed, float* grads, const PackDesc* tab, int n, int max_elems, cudaStream_t st) {
  if (ablate(9)) return 0;
  if (n == 0) return 0;
  pack_kernel<<<dim3(pack_blocks(max_elems), n), 256, 0, st>>>(packed, grads, nullptr, tab, 1, n, nullptr, nullptr, nullptr, nullptr, nullptr);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// layout + elementwise
// ---------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, TV dst, int C, size_t HW, size_t total_pix) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_pix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, hw = i % HW;
    for (int c = 0; c < C; ++c) stv(dst.p, dst.bf, i * dst.pitch + dst.coff + c, src[(n * C + c) * HW + hw]);
  }
}
__global__ void nhwc_to_nchw_kernel(TV src, float* __restrict__ dst, int C, size_t HW, size_t total_pix) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_pix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, hw = i % HW;
    for (int c = 0; c < C; ++c) dst[(n * C + c) * HW + hw] = ldv(src.p, src.bf, i * src.pitch + src.coff + c);
  }
}
static inline unsigned ew_grid(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}
int nchw_to_nhwc(const float* src, TV dst, int B, int C, int H, int W, cudaStream_t st) {
  const size_t HW = (size_t)H * W, tot = HW * B;
  nchw_to_nhwc_kernel<<<ew_grid(tot), 256, 0, st>>>(src, dst, C, HW, tot);
  DG_LAUNCH_CHECK();
  return 0;
}
int nhwc_to_nchw(TV src, float* dst, int B, int C, int H, int W, cudaStream_t st) {
  const size_t HW = (size_t)H * W, tot = HW * B;
  nhwc_to_nchw_kernel<<<ew_grid(tot), 256, 0, st>>>(src, dst, C, HW, tot);
  DG_LAUNCH_CHECK();
  return 0;
}

__global__ void scale_add_kernel(TV dst, TV a, float sa, TV b, float sb, size_t pixels, int C) {
  const size_t n = pixels * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / C;
    const int c = (int)(i % C);
    float v = sa * ldv(a.p, a.bf, p * a.pitch + a.coff + c);
    if (b.p) v = fmaf(sb, ldv(b.p, b.bf, p * b.pitch + b.coff + c), v);
    stv(dst.p, dst.bf, p * dst.pitch + dst.coff + c, v);
  }
}
int scale_add(TV dst, TV a, float sa, TV b, float sb, size_t pixels, int C, cudaStream_t st) {
  scale_add_kernel<<<ew_grid(pixels * C), 256, 0, st>>>(dst, a, sa, b, sb, pixels, C);
  DG_LAUNCH_CHECK();
  return 0;
}

// [real ; fake ; alpha*real + (1-alpha)*fake]  (wasserstein.py:91-94), NHWC fp32 out
__global__ void build_critic_input_kernel(const float* __restrict__ real, const float* __restrict__ fake, int fake_nchw,
                                          const float* __restrict__ alpha, float* __restrict__ dst, int B, int C,
                                          size_t HW, int mode) {
  const size_t tot = (size_t)B * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, hw = i % HW;
    const float a = alpha ? alpha[n] : 0.f;
    for (int c = 0; c < C; ++c) {
      const float r = real[(n * C + c) * HW + hw];
      const float f = fake_nchw ? fake[(n * C + c) * HW + hw] : fake[i * C + c];
      const float x = a * r + (1.f - a) * f;
      if (mode == 0 || mode == 2) {  // mode 2: [real ; fake] only (metric pass)
        dst[i * C + c] = r;
        dst[(tot + i) * C + c] = f;
        if (mode == 0) dst[(2 * tot + i) * C + c] = x;
      } else {
        dst[i * C + c] = x;
      }
    }
  }
}
int build_critic_input(const float* real, const float* fake, int fake_is_nchw, const float* alpha, float* dst, int B,
                       int C, int H, int W, int mode, cudaStream_t st) {
  const size_t HW = (size_t)H * W;
  Prof prof(PC_INTERP, 0.0, (double)HW * B * C * 4.0 * (mode == 0 ? 5.0 : (mode == 2 ? 4.0 : 3.0)), st);
  build_critic_input_kernel<<<ew_grid(HW * B), 256, 0, st>>>(real, fake, fake_is_nchw, alpha, dst, B, C, HW, mode);
  DG_LAUNCH_CHECK();
  return 0;
}

__global__ void fill_kernel(float* p, float v, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
int fill(float* p, float v, long long n, cudaStream_t st) {
  fill_kernel<<<ew_grid((size_t)n), 256, 0, st>>>(p, v, n);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// critic classifier (critic.py:94-99)
// ---------------------------------------------------------------------------
// y[b][j] = act(sum_k x[b][k] w[j][k] + bias[j]).  Split-K tiled GEMM on CUDA cores (0.3 GFLOP at cfg-2: the
// bound is re-reading x and w, so a block keeps a 32-sample x 104-unit tile over its K chunk in registers,
// operands staged through shared memory) + a tiny finish kernel for bias / activation.
constexpr int FCF_BM = 32, FCF_BN = 104, FCF_BK = 32, FCF_WP = FCF_BK + 4;  // w tile row pitch (floats): 16-byte rows, conflict-free
__device__ __forceinline__ void fc_cp16(void* dst, const void* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
// K % 8 == 0 and kchunk % 32 == 0 (checked by the host): 16-byte cp.async chunks, double-buffered k-tiles
template <bool XBF>
__global__ void __launch_bounds__(256) fc_fwd_tiled_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                            float* __restrict__ y, int NB, int K, int N, int kchunk) {
  __shared__ __align__(16) float ws[2][FCF_BN][FCF_WP];
  __shared__ __align__(16) unsigned char xs_raw[2][FCF_BM * FCF_BK * (XBF ? 2 : 4)];
  const int b0 = blockIdx.x * FCF_BM;
  const int k_begin = blockIdx.y * kchunk, k_end = min(K, k_begin + kchunk);
  const int tb = threadIdx.x >> 3, tj = threadIdx.x & 7;
  const int ntiles = (k_end - k_begin + FCF_BK - 1) / FCF_BK;
  constexpr int XE = XBF ? 8 : 4;              // x elements per 16-byte chunk
  constexpr int XCH = FCF_BM * FCF_BK / XE;    // x chunks per tile
  auto stage = [&](int t, int buf) {
    const int k0 = k_begin + t * FCF_BK;
    for (int i = threadIdx.x; i < XCH; i += 256) {
      const int r = i / (FCF_BK / XE), c = (i - r * (FCF_BK / XE)) * XE;
      const bool ok = (b0 + r < NB) && (k0 + c < k_end);
      const size_t off = ok ? (size_t)(b0 + r) * K + k0 + c : 0;
      fc_cp16(xs_raw[buf] + (size_t)(r * FCF_BK + c) * (XBF ? 2 : 4), (const char*)x + off * (XBF ? 2 : 4), ok);
    }
    for (int i = threadIdx.x; i < FCF_BN * (FCF_BK / 4); i += 256) {
      const int r = i >> 3, c = (i & 7) * 4;
      const bool ok = (r < N) && (k0 + c < k_end);
      fc_cp16(&ws[buf][r][c], w + (ok ? (size_t)r * K + k0 + c : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[FCF_BN / 8];
#pragma unroll
  for (int i = 0; i < FCF_BN / 8; ++i) acc[i] = 0.f;
  if (ntiles > 0) stage(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) { stage(t + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < FCF_BK; ++c) {
      float xv;
      if (XBF) xv = __bfloat162float(reinterpret_cast<const bf16*>(xs_raw[buf])[tb * FCF_BK + c]);
      else xv = reinterpret_cast<const float*>(xs_raw[buf])[tb * FCF_BK + c];
#pragma unroll
      for (int i = 0; i < FCF_BN / 8; ++i) acc[i] = fmaf(xv, ws[buf][tj + 8 * i][c], acc[i]);
    }
    __syncthreads();
  }
  if (b0 + tb < NB) {
#pragma unroll
    for (int i = 0; i < FCF_BN / 8; ++i) {
      const int j = tj + 8 * i;
      if (j < N) atomicAdd(&y[(size_t)(b0 + tb) * N + j], acc[i]);
    }
  }
}
__global__ void fc_finish_kernel(float* __restrict__ y, const float* __restrict__ bias, int NB, int N, int act, float slope,
                                 const float* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB * N) return;
  float s = y[i];
  if (bias) s += bias[i % N];
  if (act == ACT_LRELU) s = s > 0.f ? s : s * slope;
  else if (act == ACT_MASK) s *= lrelu_d(mask[i], slope);
  y[i] = s;
}
// fallback for wide layers (N > 104): one warp per (b, j)
__global__ void fc_fwd_kernel(const void* __restrict__ x, int x_bf, const float* __restrict__ w,
                              const float* __restrict__ bias, float* __restrict__ y, int NB, int K, int N, int act,
                              float slope, const float* __restrict__ mask) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= NB * N) return;
  const int b = warp / N, j = warp % N;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(ldv(x, x_bf, (size_t)b * K + k), w[(size_t)j * K + k], s);
  s = warp_sum(s);
  if (lane == 0) {
    if (bias) s += bias[j];
    if (act == ACT_LRELU) s = s > 0.f ? s : s * slope;
    else if (act == ACT_MASK) s *= lrelu_d(mask[(size_t)b * N + j], slope);
    y[(size_t)b * N + j] = s;
  }
}
// y = x w^T without bias / activation (the caller's next kernel applies them); false: shape not covered by the tiled kernel
bool fc_fwd_raw_supported(int K, int N) { return N <= FCF_BN && K % 8 == 0; }
int fc_fwd_raw(const void* x, int x_bf, const float* w, float* y, int NB, int K, int N, cudaStream_t st) {
  DG_CHECK(fc_fwd_raw_supported(K, N), "fc_fwd_raw: unsupported shape K=%d N=%d", K, N);
  Prof prof(PC_FC, 2.0 * NB * N * K, (double)N * K * 4.0 + (double)NB * K * (x_bf ? 2 : 4), st);
  const int nbt = (NB + FCF_BM - 1) / FCF_BM;
  int splits = (296 + nbt - 1) / nbt;
  splits = std::max(1, std::min(splits, (K + 63) / 64));
  const int kchunk = ((K + splits - 1) / splits + FCF_BK - 1) / FCF_BK * FCF_BK;
  splits = (K + kchunk - 1) / kchunk;
  DG_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)NB * N, st));
  if (fc_umma_supported(NB, K, N, x_bf)) return fc_fwd_umma(x, w, y, NB, K, N, st);
  if (x_bf) fc_fwd_tiled_kernel<true><<<dim3(nbt, splits), 256, 0, st>>>(x, w, y, NB, K, N, kchunk);
  else fc_fwd_tiled_kernel<false><<<dim3(nbt, splits), 256, 0, st>>>(x, w, y, NB, K, N, kchunk);
  DG_LAUNCH_CHECK();
  return 0;
}

// Classifier head of the fused critic iteration in ONE launch (single CTA, deterministic), for the 3B batch
// [real ; fake ; interpolates]: a9 = lrelu(y + b1) in place (classifier.1), scores = a9 w2 + b2 (classifier.2), the score
// means of the real / fake rows (wasserstein.py:46-47), the per-row loss seeds (-1/B, +1/B, 1 for the penalty's ones
// seed) and dz9 = seed * w2 * lrelu'(a9).  Replaces fc_finish + fc2_fwd + critic_means + critic_seed + fc2_seed.
__device__ unsigned int g_head_done = 0;  // completion counter (the fused iteration runs on one stream at a time)
__global__ void __launch_bounds__(256) critic_head_kernel(float* __restrict__ a9, const float* __restrict__ b1,
                                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                                          float* __restrict__ scores, float* __restrict__ seed,
                                                          float* __restrict__ dz9, float* __restrict__ scalars, int B, int K,
                                                          float slope) {
  __shared__ float part[2][8];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;  // one sample per warp
  if (b < 3 * B) {
    const float sd = b < B ? -1.f / B : (b < 2 * B ? 1.f / B : 1.f);
    float v = 0.f;
    for (int k = lane; k < K; k += 32) {
      const size_t i = (size_t)b * K + k;
      float s = a9[i] + b1[k];
      s = s > 0.f ? s : s * slope;
      a9[i] = s;
      const float w = w2[k];
      v = fmaf(s, w, v);
      dz9[i] = sd * w * lrelu_d(s, slope);
    }
    v = warp_sum(v);
    if (lane == 0) {
      scores[b] = v + b2[0];
      seed[b] = sd;
    }
  }
  // the last block to finish sums the real / fake scores in a fixed order (deterministic means)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&g_head_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  float r = 0.f, f = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) { r += __ldcg(scores + i); f += __ldcg(scores + B + i); }
  r = warp_sum(r);
  f = warp_sum(f);
  if (lane == 0) { part[0][warp] = r; part[1][warp] = f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    r = 0.f; f = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { r += part[0][k]; f += part[1][k]; }
    scalars[1] = r / B;
    scalars[2] = f / B;
    g_head_done = 0;
  }
}
bool critic_head_supported(int B) { return B >= 1; }
int critic_head(float* a9, const float* b1, const float* w2, const float* b2, float* scores, float* seed, float* dz9,
                float* scalars, int B, int K, float slope, cudaStream_t st) {
  critic_head_kernel<<<(3 * B + 7) / 8, 256, 0, st>>>(a9, b1, w2, b2, scores, seed, dz9, scalars, B, K, slope);
  DG_LAUNCH_CHECK();
  return 0;
}

int fc_fwd(const void* x, int x_bf, const float* w, const float* bias, float* y, int NB, int K, int N, int act,
           float slope, const float* mask, cudaStream_t st) {
  Prof prof(PC_FC, 2.0 * NB * N * K, (double)N * K * 4.0 + (double)NB * K * (x_bf ? 2 : 4), st);
  if (N > FCF_BN || K % 8) {
    const long long threads = (long long)NB * N * 32;
    fc_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, x_bf, w, bias, y, NB, K, N, act, slope, mask);
    DG_LAUNCH_CHECK();
    return 0;
  }
  const int nbt = (NB + FCF_BM - 1) / FCF_BM;
  int splits = (296 + nbt - 1) / nbt;
  splits = std::max(1, std::min(splits, (K + 63) / 64));
  const int kchunk = ((K + splits - 1) / splits + FCF_BK - 1) / FCF_BK * FCF_BK;
  splits = (K + kchunk - 1) / kchunk;
  DG_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)NB * N, st));
  if (fc_umma_supported(NB, K, N, x_bf)) DG_TRY(fc_fwd_umma(x, w, y, NB, K, N, st));
  else {
    if (x_bf) fc_fwd_tiled_kernel<true><<<dim3(nbt, splits), 256, 0, st>>>(x, w, y, NB, K, N, kchunk);
    else fc_fwd_tiled_kernel<false><<<dim3(nbt, splits), 256, 0, st>>>(x, w, y, NB, K, N, kchunk);
    DG_LAUNCH_CHECK();
  }
  fc_finish_kernel<<<(NB * N + 255) / 256, 256, 0, st>>>(y, bias, NB, N, act, slope, mask);
  DG_LAUNCH_CHECK();
  return 0;
}

// dx[b][k] = (sum_j dz[b][j] w[j][k]) * lrelu'(mask[b][k]); a thread owns one k for 16 samples (w read once per 16)
constexpr int FCD_BB = 16;
__global__ void __launch_bounds__(256) fc_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w, void* __restrict__ dx,
                                                        int dx_bf, int NB, int K, int N, const void* __restrict__ mask, int mask_bf,
                                                        float slope) {
  extern __shared__ __align__(16) float sdz[];  // [N][FCD_BB]
  const int b0 = blockIdx.y * FCD_BB;
  for (int i = threadIdx.x; i < N * FCD_BB; i += blockDim.x) {
    const int j = i / FCD_BB, bb = i - j * FCD_BB;
    sdz[i] = (b0 + bb < NB) ? dz[(size_t)(b0 + bb) * N + j] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[FCD_BB];
#pragma unroll
  for (int i = 0; i < FCD_BB; ++i) acc[i] = 0.f;
#pragma unroll 5
  for (int j = 0; j < N; ++j) {
    const float wv = w[(size_t)j * K + k];
    const float4* p = reinterpret_cast<const float4*>(sdz + j * FCD_BB);
#pragma unroll
    for (int q = 0; q < FCD_BB / 4; ++q) {
      const float4 d = p[q];
      acc[4 * q] = fmaf(d.x, wv, acc[4 * q]);
      acc[4 * q + 1] = fmaf(d.y, wv, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(d.z, wv, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(d.w, wv, acc[4 * q + 3]);
    }
  }
#pragma unroll
  for (int bb = 0; bb < FCD_BB; ++bb) {
    if (b0 + bb >= NB) break;
    float s = acc[bb];
    const size_t o = (size_t)(b0 + bb) * K + k;
    if (mask) s *= lrelu_d(ldv(mask, mask_bf, o), slope);
    stv(dx, dx_bf, o, s);
  }
}
int fc_dgrad(const float* dz, const float* w, void* dx, int dx_bf, int NB, int K, int N, const void* mask, int mask_bf,
             float slope, cudaStream_t st) {
  Prof prof(PC_FC, 2.0 * NB * N * K, (double)N * K * 4.0 + (double)NB * K * (dx_bf ? 2 : 4), st);
  if (mask && mask_bf && fc_umma_supported(NB, K, N, dx_bf)) return fc_dgrad_umma(dz, w, dx, NB, K, N, mask, slope, st);
  fc_dgrad_kernel<<<dim3((K + 255) / 256, (NB + FCD_BB - 1) / FCD_BB), 256, N * FCD_BB * sizeof(float), st>>>(
      dz, w, dx, dx_bf, NB, K, N, mask, mask_bf, slope);
  DG_LAUNCH_CHECK();
  return 0;
}

// dw[j][k] += sum_b dz[b][j] x[b][k]; a thread owns one k for 10 output units (x read once per 10)
constexpr int FCW_BJ = 10, FCW_MAXB = 1024;
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const float* __restrict__ dz, const void* __restrict__ x, int x_bf,
                                                        float* __restrict__ dw, int NB, int K, int N) {
  extern __shared__ __align__(16) float sdz[];  // [NB][FCW_BJ]
  const int j0 = blockIdx.y * FCW_BJ;
  for (int i = threadIdx.x; i < NB * FCW_BJ; i += blockDim.x) {
    const int b = i / FCW_BJ, jj = i - b * FCW_BJ;
    sdz[i] = (j0 + jj < N) ? dz[(size_t)b * N + j0 + jj] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[FCW_BJ];
#pragma unroll
  for (int i = 0; i < FCW_BJ; ++i) acc[i] = 0.f;
  int b = 0;
  for (; b + 3 < NB; b += 4) {
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = ldv(x, x_bf, (size_t)(b + u) * K + k);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int jj = 0; jj < FCW_BJ; ++jj) acc[jj] = fmaf(sdz[(b + u) * FCW_BJ + jj], xv[u], acc[jj]);
  }
  for (; b < NB; ++b) {
    const float xv = ldv(x, x_bf, (size_t)b * K + k);
#pragma unroll
    for (int jj = 0; jj < FCW_BJ; ++jj) acc[jj] = fmaf(sdz[b * FCW_BJ + jj], xv, acc[jj]);
  }
#pragma unroll
  for (int jj = 0; jj < FCW_BJ; ++jj)
    if (j0 + jj < N) atomicAdd(&dw[(size_t)(j0 + jj) * K + k], acc[jj]);  // (two sample ranges may accumulate concurrently)
}
int fc_wgrad(const float* dz, const void* x, int x_bf, float* dw, int NB, int K, int N, cudaStream_t st) {
  DG_CHECK(NB <= FCW_MAXB, "fc_wgrad: batch %d > %d", NB, FCW_MAXB);
  Prof prof(PC_FC, 2.0 * NB * N * K, (double)N * K * 8.0 + (double)NB * K * (x_bf ? 2 : 4), st);
  if (fc_umma_supported(NB, K, N, x_bf)) return fc_wgrad_umma(dz, x, dw, NB, K, N, st);
  fc_wgrad_kernel<<<dim3((K + 255) / 256, (N + FCW_BJ - 1) / FCW_BJ), 256, (size_t)NB * FCW_BJ * sizeof(float), st>>>(dz, x, x_bf, dw,
                                                                                                                   NB, K, N);
  DG_LAUNCH_CHECK();
  return 0;
}

// The small classifier gradients of the fused critic step in ONE launch (fp64 partial sums: the bias gradients
// are sums of cancelling real/fake halves):
//   d_fc1b[j] += sum_{b<n0} dz9[b][j];   d_fc2w[j] += sum_{b<n0} seed[b]*a9[b][j] + sum_{b<nv} vfc[b][j];
//   d_fc2b    += sum_{b<n0} seed[b]
__global__ void __launch_bounds__(256) critic_small_grads_kernel(const float* __restrict__ dz9, const float* __restrict__ seed,
                                                                  const float* __restrict__ a9, const float* __restrict__ vfc, int n0,
                                                                  int nv, int N, float* __restrict__ d_fc1b, float* __restrict__ d_fc2w,
                                                                  float* __restrict__ d_fc2b) {
  __shared__ double sh[3][8][32];
  const int jl = threadIdx.x & 31, g = threadIdx.x >> 5, j = blockIdx.x * 32 + jl;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (j < N) {
    for (int b = g; b < n0; b += 8) {
      s1 += (double)dz9[(size_t)b * N + j];
      s2 += (double)seed[b] * (double)a9[(size_t)b * N + j];
    }
    for (int b = g; b < nv; b += 8) s2 += (double)vfc[(size_t)b * N + j];
  }
  if (blockIdx.x == 0)
    for (int b = threadIdx.x; b < n0; b += 256) s3 += (double)seed[b];
  sh[0][g][jl] = s1; sh[1][g][jl] = s2; sh[2][g][jl] = s3;
  __syncthreads();
  if (g == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { t1 += sh[0][q][jl]; t2 += sh[1][q][jl]; }
    if (j < N) { d_fc1b[j] += (float)t1; d_fc2w[j] += (float)t2; }
    if (blockIdx.x == 0) {
      double t3 = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t3 += sh[2][q][jl];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t3 += __shfl_xor_sync(0xffffffffu, t3, o);
      if (jl == 0) d_fc2b[0] += (float)t3;
    }
  }
}
int critic_small_grads(const float* dz9, const float* seed, const float* a9, const float* vfc, int n0, int nv, int N, float* d_fc1b,
                       float* d_fc2w, float* d_fc2b, cudaStream_t st) {
  critic_small_grads_kernel<<<(N + 31) / 32, 256, 0, st>>>(dz9, seed, a9, vfc, n0, nv, N, d_fc1b, d_fc2w, d_fc2b);
  DG_LAUNCH_CHECK();
  return 0;
}

__global__ void fc2_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ s, int NB, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= NB) return;
  float v = 0.f;
  for (int k = lane; k < K; k += 32) v = fmaf(a[(size_t)warp * K + k], w[k], v);
  v = warp_sum(v);
  if (lane == 0) s[warp] = v + (bias ? bias[0] : 0.f);
}
int fc2_fwd(const float* a, const float* w, const float* bias, float* s, int NB, int K, cudaStream_t st) {
  fc2_fwd_kernel<<<(NB * 32 + 255) / 256, 256, 0, st>>>(a, w, bias, s, NB, K);
  DG_LAUNCH_CHECK();
  return 0;
}
// dz9[b][j] = seed[b] * w2[j] * lrelu'(a9[b][j])
__global__ void fc2_seed_kernel(const float* __restrict__ a9, const float* __restrict__ w2, const float* __restrict__ seed,
                                float* __restrict__ dz9, int NB, int K, float slope) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB * K) return;
  const int b = i / K, j = i % K;
  dz9[i] = seed[b] * w2[j] * lrelu_d(a9[i], slope);
}
int fc2_seed(const float* a9, const float* w2, const float* seed, float* dz9, int NB, int K, float slope, cudaStream_t st) {
  fc2_seed_kernel<<<(NB * K + 255) / 256, 256, 0, st>>>(a9, w2, seed, dz9, NB, K, slope);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// losses and gradient-penalty reductions
// ---------------------------------------------------------------------------
// scalars[1] = mean(scores[0:B]) (real), scalars[2] = mean(scores[B:2B]) (fake)
__global__ void critic_means_kernel(const float* __restrict__ scores, int B, float* __restrict__ scalars) {
  __shared__ float sh[32];
  float r = 0.f, f = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { r += scores[i]; f += scores[B + i]; }
  r = block_sum(r, sh);
  f = block_sum(f, sh);
  if (threadIdx.x == 0) { scalars[1] = r / B; scalars[2] = f / B; }
}
int critic_means(const float* scores, int B, float* scalars, cudaStream_t st) {
  critic_means_kernel<<<1, 256, 0, st>>>(scores, B, scalars);
  DG_LAUNCH_CHECK();
  return 0;
}

// sumsq[b] += sum over the sample of g^2 (vectorised, one atomic per block)
// per-sample sum of squares of the input gradient (wasserstein.py:110-114), then - in the LAST block to finish - the norms, the
// penalty value and dGP/dg coefficients  lam^2 * (2/B) * (n-1)/n  (wasserstein.py:114-117,40).  Block partials go to
// part[b][blockIdx.x] and are combined in a fixed order (deterministic, no float atomics, no memset launch); one launch instead
// of memset + reduction + finish.
__device__ unsigned int g_gp_done = 0;  // completion counter (one gradient-penalty pass at a time per process, see header)
constexpr int GPN_MAXBX = 32;
__global__ void __launch_bounds__(256) gp_norms_finish_kernel(const float* __restrict__ g, size_t per_sample, float* __restrict__ part,
                                                              int B, float lam, float* __restrict__ sumsq, float* __restrict__ norms,
                                                              float* __restrict__ coef, float* __restrict__ scalars, int write_loss) {
  __shared__ float sh[32];
  __shared__ bool last;
  const int b = blockIdx.y;
  const float4* gp = reinterpret_cast<const float4*>(g + (size_t)b * per_sample);
  const size_t n4 = per_sample >> 2;
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(gp + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n4 << 2; i < per_sample; ++i) { const float v = g[(size_t)b * per_sample + i]; s += v * v; }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    part[(size_t)b * GPN_MAXBX + blockIdx.x] = s;
    __threadfence();
    last = (atomicAdd(&g_gp_done, 1u) == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float t = 0.f;
    for (unsigned k = 0; k < gridDim.x; ++k) t += __ldcg(part + (size_t)i * GPN_MAXBX + k);
    sumsq[i] = t;
    const float n = sqrtf(t + 1e-12f);
    if (norms) norms[i] = n;
    if (coef) coef[i] = lam * lam * (2.f / B) * (n - 1.f) / n;
    acc += (n - 1.f) * (n - 1.f);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    const float gpv = lam * acc / B;
    scalars[3] = gpv;
    scalars[4] = lam * gpv;
    if (write_loss) scalars[0] = scalars[2] - scalars[1] + lam * gpv;
    g_gp_done = 0;
  }
}
static float* g_gp_part = nullptr;
static int g_gp_part_B = 0;
int gp_norms_finish(const float* g, int B, size_t per_sample, float gp_lambda, float* sumsq, float* norms, float* coef, float* scalars,
                    int write_loss, cudaStream_t st) {
  if (B > g_gp_part_B) {  // scratch for the block partials, grown on first use / larger batches (not on the steady-state path)
    if (g_gp_part) cudaFree(g_gp_part);
    DG_CUDA(cudaMalloc(&g_gp_part, sizeof(float) * (size_t)B * GPN_MAXBX));
    g_gp_part_B = B;
  }
  Prof prof(PC_GP_NORMS, 0.0, (double)per_sample * B * 4.0, st);
  unsigned bx = (unsigned)((per_sample / 4 + 1023) / 1024);
  if (bx < 1) bx = 1;
  if (bx > GPN_MAXBX) bx = GPN_MAXBX;
  gp_norms_finish_kernel<<<dim3(bx, B), 256, 0, st>>>(g, per_sample, g_gp_part, B, gp_lambda, sumsq, norms, coef, scalars, write_loss);
  DG_LAUNCH_CHECK();
  return 0;
}
int gp_norms(const float* g, int B, size_t per_sample, float* sumsq, cudaStream_t st) {
  (void)g; (void)B; (void)per_sample; (void)sumsq; (void)st;
  set_error("gp_norms: superseded by gp_norms_finish");
  return DG_ERR_STATE;
}
int gp_finish(const float* sumsq, int B, float lam, float* norms, float* coef, float* scalars, int write_loss, cudaStream_t st) {
  (void)sumsq; (void)B; (void)lam; (void)norms; (void)coef; (void)scalars; (void)write_loss; (void)st;
  set_error("gp_finish: superseded by gp_norms_finish");
  return DG_ERR_STATE;
}
__global__ void gp_scale_kernel(const float* __restrict__ g, const float* __restrict__ coef, float* __restrict__ u,
                                size_t per_sample, size_t total) {
  if ((per_sample & 3) == 0) {  // float4 path: a vector never straddles two samples
    const size_t n4 = total >> 2, ps4 = per_sample >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* u4 = reinterpret_cast<float4*>(u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const float c = __ldg(coef + i / ps4);
      const float4 v = __ldg(g4 + i);
      u4[i] = make_float4(c * v.x, c * v.y, c * v.z, c * v.w);
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    u[i] = coef[i / per_sample] * g[i];
}
int gp_scale(const float* g, const float* coef, float* u, int B, size_t per_sample, cudaStream_t st) {
  gp_scale_kernel<<<ew_grid(per_sample * B), 256, 0, st>>>(g, coef, u, per_sample, per_sample * B);
  DG_LAUNCH_CHECK();
  return 0;
}

// mean |a-b| (+ seed scale*sign(a-b)/n [+ d_add])   (losses.py:51-53)
__global__ void __launch_bounds__(256) l1_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float scale,
                                                 float* __restrict__ loss, float* __restrict__ d_a, const float* __restrict__ d_add,
                                                 int vec) {
  __shared__ float sh[32];
  float s = 0.f;
  const float k = scale / (float)n;
  auto one = [&](float x, float y, float add) {
    const float d = x - y;
    s += fabsf(d);
    return (d > 0.f ? k : (d < 0.f ? -k : 0.f)) + add;
  };
  const long long n4 = vec ? (n >> 2) : 0;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    float4 ad = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d_a && d_add) ad = __ldg(reinterpret_cast<const float4*>(d_add) + i);
    const float4 o = make_float4(one(x.x, y.x, ad.x), one(x.y, y.y, ad.y), one(x.z, y.z, ad.z), one(x.w, y.w, ad.w));
    if (d_a) reinterpret_cast<float4*>(d_a)[i] = o;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float o = one(a[i], b[i], (d_a && d_add) ? d_add[i] : 0.f);
    if (d_a) d_a[i] = o;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(loss, s / (float)n);
}
int l1_loss(const float* a, const float* b, long long n, float scale, float* loss_out, float* d_a, const float* d_add,
            cudaStream_t st) {
  DG_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  Prof prof(PC_L1, 0.0, (double)n * 4.0 * (2.0 + (d_a ? 1.0 : 0.0) + (d_add ? 1.0 : 0.0)), st);
  const int vec = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)d_a | (uintptr_t)d_add) & 15) == 0;
  l1_kernel<<<ew_grid((size_t)(vec ? (n + 3) / 4 : n)), 256, 0, st>>>(a, b, n, scale, loss_out, d_a, d_add, vec);
  DG_LAUNCH_CHECK();
  return 0;
}

// g_loss = -gamma*mean(c_fake) + content_lambda*l1   (wasserstein.py:74,78)
__global__ void gen_scalars_kernel(const float* __restrict__ scores, int B, const float* __restrict__ l1, float gamma,
                                   float clam, float* __restrict__ scalars) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += scores[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    const float m = s / B;
    scalars[1] = m;
    scalars[2] = l1[0];
    scalars[0] = -gamma * m + clam * l1[0];
  }
}
int gen_scalars(const float* scores, int B, const float* l1, float gamma, float clam, float* scalars, cudaStream_t st) {
  gen_scalars_kernel<<<1, 256, 0, st>>>(scores, B, l1, gamma, clam, scalars);
  DG_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// fused Adam over a flat buffer (torch.optim.Adam semantics, no amsgrad/decay)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float b1, float b2, float eps, float step_size,
                                                   float inv_sqrt_bc2, float gscale, int vec) {
  // 16-byte loads / stores over the flat buffer (vec: all four pointers 16-byte aligned), scalar grid-stride tail
  const long long n4 = vec ? (n >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    upd(p[i], g[i], m[i], v[i]);
}
int adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int step,
         float gscale, cudaStream_t st) {
  if (ablate(9)) return 0;
  Prof prof(PC_ADAM, 0.0, (double)n * 28.0, st);
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const int vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  adam_kernel<<<ew_grid((size_t)(vec ? (n + 3) / 4 : n)), 256, 0, st>>>(p, g, m, v, n, b1, b2, eps, (float)(lr / bc1),
                                                                       (float)(1.0 / sqrt(bc2)), gscale, vec);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg

extern "C" const char* dg_last_error(void) { return dg::last_error(); }

// Enable (1) / disable (0) per-launch event timing; enabling clears earlier records.
extern "C" int dg_profile(int enable) {
  if (enable) {
    for (auto& r : dg::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    dg::g_prof.clear();
  }
  dg::g_prof_on = enable != 0;
  return 0;
}
// out[cls*4 + {0,1,2,3}] = launches, total ms, algorithmic flops, algorithmic bytes. Synchronises the device.
extern "C" int dg_profile_report(double* out, int n_classes) {
  if (!out || n_classes < dg::PC_COUNT) { dg::set_error("dg_profile_report: need %d classes", (int)dg::PC_COUNT); return DG_ERR_INVALID; }
  DG_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < n_classes * 4; ++i) out[i] = 0.0;
  const char* dump = getenv("DG_PROFILE_DUMP");  // optional per-launch CSV: class,flops,bytes,ms
  FILE* f = dump ? fopen(dump, "a") : nullptr;
  struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{f};
  for (auto& r : dg::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    if (f) fprintf(f, "%d,%.0f,%.0f,%.6f\n", r.cls, r.flops, r.bytes, ms);
    out[r.cls * 4 + 0] += 1.0; out[r.cls * 4 + 1] += ms; out[r.cls * 4 + 2] += r.flops; out[r.cls * 4 + 3] += r.bytes;
  }
  return 0;
}
namespace dg {
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* addr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return CUDA_ERROR_NOT_FOUND;
    fn = (Fn)p;
  }
  return fn(map, dt, rank, addr, dims, strides, box, estr, il, sw, l2, oob);
}
}  // namespace dg
namespace dg { int g_tune[DG_TUNE_KEYS] = {1, 1, 1, 0, 1, 0, 1, 1, 1, 1, 4, 1, 1, 3, 1, 0, 1, 0, 2, 0, 0, 1, 1, 1, 0, 1}; }  // see include/downgan_b200.h: dg_set_tuning
extern "C" int dg_set_tuning(int key, int value) {
  if (key < 0 || key >= DG_TUNE_KEYS) { dg::set_error("dg_set_tuning: unknown key %d", key); return DG_ERR_INVALID; }
  const int prev = dg::g_tune[key];
  dg::g_tune[key] = value;
  if (key == 1)  // device-wide L1/shared split preference for kernels without their own (avoids carveout flips between launches)
    cudaDeviceSetCacheConfig(value ? cudaFuncCachePreferShared : cudaFuncCachePreferNone);
  return prev;
}
extern "C" int64_t dg_launch_count(void) { return (int64_t)dg::g_launches.load(); }
