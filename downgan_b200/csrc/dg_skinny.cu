// CUDA-core kernels for the "skinny" boundary layers of the two networks: the 2-channel fine
// field entering the critic (features.0: 2 -> F) and leaving the generator (conv3.2: F -> 2),
// with their data- and weight-gradients.  K = 18 or N = 2 cannot feed a tensor-core tile, and the
// fp32 input/output fields stay fp32 here; the kernels are bandwidth/FFMA bound streaming passes:
// each thread owns two horizontally adjacent pixels, weights are broadcast from shared memory.
#include "dg_common.cuh"

namespace dg {
namespace {

__device__ __forceinline__ float ldv(const void* p, int bf, size_t i) {
  return bf ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}

// ---------------------------------------------------------------------------------------------
// Ci = 2  ->  Co = 16 per blockIdx.y chunk, stride 1.  v = act(acc + bias)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_ci2_kernel(ConvOp op) {
  __shared__ __align__(16) float sw[9][2][16];
  const int CoP = (op.Co + 15) & ~15;
  const int co0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 288; i += 128) {
    const int tap = i / 32, r = i % 32, ci = r >> 4, j = r & 15;
    sw[tap][ci][j] = op.w[((size_t)tap * 2 + ci) * CoP + co0 + j];
  }
  __syncthreads();
  const int W = op.Win, H = op.Hin, W2 = W >> 1;
  const long long total = (long long)op.B * H * W2;
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= total) return;
  const int xp = (int)(t % W2) * 2;
  const long long t2 = t / W2;
  const int y = (int)(t2 % H), n = (int)(t2 / H);
  // 3 x 4 x 2 input window
  float in[3][4][2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int gy = y + r - 1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gx = xp + c - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const size_t idx = (((size_t)n * H + gy) * W + gx) * op.x.pitch + op.x.coff;
      in[r][c][0] = ok ? ldv(op.x.p, op.x.bf, idx) : 0.f;
      in[r][c][1] = ok ? ldv(op.x.p, op.x.bf, idx + 1) : 0.f;
    }
  }
  float acc[2][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float b = op.bias ? op.bias[co0 + j] : 0.f;
    acc[0][j] = b; acc[1][j] = b;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const float4* wr = reinterpret_cast<const float4*>(&sw[ky * 3 + kx][ci][0]);
        const float a0 = in[ky][kx][ci], a1 = in[ky][kx + 1][ci];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = wr[q];
          acc[0][4 * q] = fmaf(a0, w.x, acc[0][4 * q]); acc[0][4 * q + 1] = fmaf(a0, w.y, acc[0][4 * q + 1]);
          acc[0][4 * q + 2] = fmaf(a0, w.z, acc[0][4 * q + 2]); acc[0][4 * q + 3] = fmaf(a0, w.w, acc[0][4 * q + 3]);
          acc[1][4 * q] = fmaf(a1, w.x, acc[1][4 * q]); acc[1][4 * q + 1] = fmaf(a1, w.y, acc[1][4 * q + 1]);
          acc[1][4 * q + 2] = fmaf(a1, w.z, acc[1][4 * q + 2]); acc[1][4 * q + 3] = fmaf(a1, w.w, acc[1][4 * q + 3]);
        }
      }
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    const size_t p = ((size_t)n * H + y) * W + xp + px;
    float* v = acc[px];
    if (op.act == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * op.slope;
    } else if (op.act == ACT_MASK) {
      const size_t mo = p * op.mask.pitch + op.mask.coff + co0;
      if (op.mask.bf && ((mo & 7) == 0)) {  // 16 bf16 mask values as two 16-byte loads
        const uint4* mp = reinterpret_cast<const uint4*>((const bf16*)op.mask.p + mo);
        const uint4 m0 = mp[0], m1 = mp[1];
        const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[2 * k] *= (__uint_as_float(w[k] << 16) > 0.f ? 1.f : op.slope);
          v[2 * k + 1] *= (__uint_as_float(w[k] & 0xFFFF0000u) > 0.f ? 1.f : op.slope);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= (ldv(op.mask.p, op.mask.bf, mo + j) > 0.f ? 1.f : op.slope);
      }
    }
    const size_t o = p * op.y.pitch + op.y.coff + co0;
    if (op.y.bf) {
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t*>(&h);
      }
      uint4* dst = reinterpret_cast<uint4*>((bf16*)op.y.p + o);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      float4* dst = reinterpret_cast<float4*>((float*)op.y.p + o);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Ci = 16  ->  Co = 2, stride 1.  v = acc + bias   (fp32 or bf16 out, pitch >= 2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_co2_kernel(ConvOp op) {
  __shared__ __align__(16) float sw[9][16][2];
  const int CoP = (op.Co + 15) & ~15;
  for (int i = threadIdx.x; i < 288; i += 128) {
    const int tap = i / 32, r = i % 32, ci = r >> 1, j = r & 1;
    sw[tap][ci][j] = op.w[((size_t)tap * 16 + ci) * CoP + j];
  }
  __syncthreads();
  const int W = op.Win, H = op.Hin, W2 = W >> 1;
  const long long total = (long long)op.B * H * W2;
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= total) return;
  const int xp = (int)(t % W2) * 2;
  const long long t2 = t / W2;
  const int y = (int)(t2 % H), n = (int)(t2 / H);
  float acc[2][2];
  acc[0][0] = acc[1][0] = op.bias ? op.bias[0] : 0.f;
  acc[0][1] = acc[1][1] = op.bias ? op.bias[1] : 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int gy = y + r - 1;
    if (gy < 0 || gy >= H) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gx = xp + c - 1;
      if (gx < 0 || gx >= W) continue;
      float xv[16];
      const size_t idx = (((size_t)n * H + gy) * W + gx) * op.x.pitch + op.x.coff;
      if (op.x.bf) {
        const uint4* src = reinterpret_cast<const uint4*>((const bf16*)op.x.p + idx);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 q = src[h];
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            xv[h * 8 + 2 * k] = __uint_as_float(w[k] << 16);
            xv[h * 8 + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
          }
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>((const float*)op.x.p + idx);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 q = src[h];
          xv[4 * h] = q.x; xv[4 * h + 1] = q.y; xv[4 * h + 2] = q.z; xv[4 * h + 3] = q.w;
        }
      }
      // window column c feeds pixel 0 with kx = c and pixel 1 with kx = c - 1
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        if (c < 3) {
          const float2 w = *reinterpret_cast<const float2*>(&sw[r * 3 + c][ci][0]);
          acc[0][0] = fmaf(xv[ci], w.x, acc[0][0]); acc[0][1] = fmaf(xv[ci], w.y, acc[0][1]);
        }
        if (c > 0) {
          const float2 w = *reinterpret_cast<const float2*>(&sw[r * 3 + c - 1][ci][0]);
          acc[1][0] = fmaf(xv[ci], w.x, acc[1][0]); acc[1][1] = fmaf(xv[ci], w.y, acc[1][1]);
        }
      }
    }
  }
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    const size_t o = (((size_t)n * H + y) * W + xp + px) * op.y.pitch + op.y.coff;
    if (op.y.bf) {
      ((bf16*)op.y.p)[o] = __float2bfloat16_rn(acc[px][0]);
      ((bf16*)op.y.p)[o + 1] = __float2bfloat16_rn(acc[px][1]);
    } else {
      ((float*)op.y.p)[o] = acc[px][0];
      ((float*)op.y.p)[o + 1] = acc[px][1];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient with one 2-channel side and one 16-channel side (288 outputs), stride 1.
// lane = (tap = lane % 9, sub = lane / 9): a warp walks three pixels per iteration, each lane owns
// the 2 x 16 outputs of its tap, so one 8-byte and one 32/64-byte load feed 32 FMAs.
//   x2 == 1 : x has 2 channels (critic features.0), dy has 16
//   x2 == 0 : x has 16 channels, dy has 2 (generator conv3.2)
// ---------------------------------------------------------------------------------------------
constexpr int WS_WARPS = 8;

__device__ __forceinline__ void ld16(const TV& t, size_t i, bool ok, float* v) {
  if (!ok) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    return;
  }
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
__device__ __forceinline__ void ld2(const TV& t, size_t i, bool ok, float* v) {
  v[0] = ok ? ldv(t.p, t.bf, i) : 0.f;
  v[1] = ok ? ldv(t.p, t.bf, i + 1) : 0.f;
}

__global__ void __launch_bounds__(WS_WARPS * 32) wgrad_skinny_kernel(WgradOp op, int x2, long long pix_per_warp) {
  __shared__ float red[WS_WARPS][27][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tap = lane % 9, sub = lane / 9;
  const bool active = lane < 27;
  const int ky = tap / 3, kx = tap - 3 * ky;
  const int H = op.Hin, W = op.Win;
  const long long total = (long long)op.B * H * W;
  const long long gw = (long long)blockIdx.x * WS_WARPS + warp;
  const long long p0 = gw * pix_per_warp, p1 = min(total, p0 + pix_per_warp);
  float acc[2][16];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  if (active) {
    for (long long p = p0 + sub; p < p1; p += 3) {
      const int x = (int)(p % W);
      const long long t = p / W;
      const int y = (int)(t % H);
      const long long nrow = t - y;  // n*H
      const int gy = y + ky - 1, gx = x + kx - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const size_t xi = (((size_t)nrow + gy) * W + gx) * op.x.pitch + op.x.coff;
      const size_t di = (size_t)p * op.dy.pitch + op.dy.coff;
      float a2[2], a16[16];
      if (x2) { ld2(op.x, xi, ok, a2); ld16(op.dy, di, true, a16); }
      else { ld16(op.x, xi, ok, a16); ld2(op.dy, di, true, a2); }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc[0][j] = fmaf(a2[0], a16[j], acc[0][j]);
        acc[1][j] = fmaf(a2[1], a16[j], acc[1][j]);
      }
    }
  }
  // block reduction: rows = lane (tap, sub), 32 values each
  if (active) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) red[warp][lane][i * 16 + j] = acc[i][j];
  }
  __syncthreads();
  const int CoP = (op.Co + 15) & ~15;
  for (int i = threadIdx.x; i < 288; i += WS_WARPS * 32) {
    const int k = i / 32, e = i % 32;   // k = tap, e = c2*16 + c16
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WS_WARPS; ++w) s += red[w][k][e] + red[w][k + 9][e] + red[w][k + 18][e];
    const int c2 = e >> 4, c16 = e & 15;
    const int ci = x2 ? c2 : c16, co = x2 ? c16 : c2;
    atomicAdd(&op.dw[((size_t)k * op.Ci + ci) * CoP + co], s);
  }
}

bool aligned16(const TV& t) {
  return t.bf ? (t.pitch % 8 == 0 && t.coff % 8 == 0) : (t.pitch % 4 == 0 && t.coff % 4 == 0);
}

}  // namespace

bool conv_skinny_supported(const ConvOp& op) {
  if (op.transposed || op.stride != 1 || op.shuffle != SHUF_NONE || op.r1.p || op.r2.p) return false;
  if (op.Hin != op.Hout || op.Win != op.Wout || (op.Win & 1)) return false;
  if (op.Ci == 2 && op.Co % 16 == 0) return aligned16(op.y);
  if (op.Co == 2 && op.Ci == 16 && op.act == ACT_NONE) return aligned16(op.x);
  return false;
}

int conv_skinny(const ConvOp& op, cudaStream_t st) {
  const long long total = (long long)op.B * op.Hin * (op.Win >> 1);
  const unsigned gx = (unsigned)((total + 127) / 128);
  const long long px = (long long)op.B * op.Hout * op.Wout;
  Prof prof(PC_CONV_DIRECT, 2.0 * px * op.Co * op.Ci * 9.0,
            (double)px * op.Co * (op.y.bf ? 2 : 4) + (double)px * op.Ci * (op.x.bf ? 2 : 4), st);
  if (op.Ci == 2) conv_ci2_kernel<<<dim3(gx, op.Co / 16), 128, 0, st>>>(op);
  else conv_co2_kernel<<<gx, 128, 0, st>>>(op);
  DG_LAUNCH_CHECK();
  return 0;
}

bool wgrad_skinny_supported(const WgradOp& op) {
  if (op.stride != 1 || op.Hin != op.Hout || op.Win != op.Wout) return false;
  if (op.Ci == 2 && op.Co == 16) return aligned16(op.dy);
  if (op.Ci == 16 && op.Co == 2) return aligned16(op.x);
  return false;
}

int wgrad_skinny(const WgradOp& op, cudaStream_t st) {
  const long long total = (long long)op.B * op.Hin * op.Win;
  long long warps = 148LL * 4 * WS_WARPS;
  long long ppw = (total + warps - 1) / warps;
  if (ppw < 32) ppw = 32;
  const unsigned blocks = (unsigned)((total + ppw * WS_WARPS - 1) / (ppw * WS_WARPS));
  Prof prof(PC_WGRAD_DIRECT, 2.0 * total * op.Co * op.Ci * 9.0,
            (double)total * op.Co * (op.dy.bf ? 2 : 4) + (double)total * op.Ci * (op.x.bf ? 2 : 4), st);
  wgrad_skinny_kernel<<<blocks, WS_WARPS * 32, 0, st>>>(op, op.Ci == 2 ? 1 : 0, ppw);
  DG_LAUNCH_CHECK();
  if (op.dbias) DG_TRY(colsum(op.dy, (size_t)total, op.Co, op.dbias, st));
  return 0;
}

}  // namespace dg
