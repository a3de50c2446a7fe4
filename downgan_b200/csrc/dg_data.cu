// Data-path kernels either side of the training iteration, sm_100a:
//   * gather_rows      - the batch assembly of the reference's DataLoader (GAN/dataloader.py:25-33 `__getitem__` per sample +
//                        default collate = torch.stack) as ONE coalesced row gather from the HBM-resident dataset;
//   * metric_sums      - MAE / MSE between the real and the generated fine fields (GAN/losses.py:40-68 `content_loss`,
//                        `content_MSELoss`, called per batch by mlflow_tools/mlflow_epoch.py:53-63), one pass over both
//                        tensors, vectorised loads, warp-shuffle + fixed-order block combine (deterministic);
//   * lowpass_replicate - the frequency-separation filter of GAN/wasserstein_fs.py (config/hyperparams.py:31-35:
//                        ReplicationPad2d(2) followed by AvgPool2d(5, stride 1)) as one stencil pass.
// All three are HBM-bound; algorithmic bytes are stated at the launch sites (Prof records).
#include <algorithm>

#include "dg_common.cuh"

namespace dg {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dst[r][:] = src[idx[r]][:]; one row = row_elems floats (a multiple of 4, 16-byte aligned rows), grid.y = row
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const long long* __restrict__ idx,
                                                          float4* __restrict__ dst, long long row_vec, long long n_src) {
  const long long r = blockIdx.y;
  long long s = idx[r];
  if (s < 0 || s >= n_src) return;  // host validates; never read out of bounds
  const float4* in = src + s * row_vec;
  float4* out = dst + r * row_vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += (long long)gridDim.x * blockDim.x)
    out[i] = __ldg(in + i);
}
__global__ void gather_rows_scalar_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst,
                                          long long row_elems, long long n_src) {
  const long long r = blockIdx.y;
  const long long s = idx[r];
  if (s < 0 || s >= n_src) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_elems; i += (long long)gridDim.x * blockDim.x)
    dst[r * row_elems + i] = src[s * row_elems + i];
}

// partial[block] = {sum |a-b|, sum (a-b)^2} over this block's grid-stride share; n4 float4 elements + tail
constexpr int MS_THREADS = 256;
__global__ void __launch_bounds__(MS_THREADS) metric_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                                 double* __restrict__ partial) {
  __shared__ float sh[2][MS_THREADS / 32];
  float s1 = 0.f, s2 = 0.f;
  const long long n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s1 += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    s2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float d = a[i] - b[i];
      s1 += fabsf(d);
      s2 += d * d;
    }
  s1 = warp_sum_f(s1);
  s2 = warp_sum_f(s2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s1; sh[1][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < MS_THREADS / 32; ++k) { t1 += sh[0][k]; t2 += sh[1][k]; }
    partial[2 * blockIdx.x] = t1;
    partial[2 * blockIdx.x + 1] = t2;
  }
}
// out[0] = MAE, out[1] = MSE, out[2] = Wass = mean C(real) - mean C(fake) (losses.py:8-9), out[3] = mean C(real), out[4] = mean C(fake)
__global__ void metric_finish_kernel(const double* __restrict__ partial, int nblocks, long long n, const float* __restrict__ scores,
                                     int B, float* __restrict__ out) {
  __shared__ double sh[4][32];
  double t1 = 0.0, t2 = 0.0, r = 0.0, f = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { t1 += partial[2 * i]; t2 += partial[2 * i + 1]; }
  for (int i = threadIdx.x; i < B; i += blockDim.x) { r += scores[i]; f += scores[B + i]; }
  // fixed-order combine: lane partials -> shared -> thread 0
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    r += __shfl_xor_sync(0xffffffffu, r, o); f += __shfl_xor_sync(0xffffffffu, f, o);
  }
  if (l == 0) { sh[0][w] = t1; sh[1][w] = t2; sh[2][w] = r; sh[3][w] = f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    t1 = t2 = r = f = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t1 += sh[0][k]; t2 += sh[1][k]; r += sh[2][k]; f += sh[3][k]; }
    out[0] = (float)(t1 / (double)n);
    out[1] = (float)(t2 / (double)n);
    out[3] = (float)(r / B);
    out[4] = (float)(f / B);
    out[2] = out[3] - out[4];
    out[5] = out[6] = out[7] = 0.f;
  }
}

// Low-pass filter of GAN/wasserstein_fs.py:42-43 (hp.low(hp.rf(x)): ReplicationPad2d(R) then AvgPool2d(2R+1, stride 1)) on
// images whose element (n, h, w, c) lives at ((n*H + h)*W + w)*cs + c  (NHWC: cs = C, images = N; NCHW: cs = 1, images = N*C).
//   mode 0: y = low(x)        mode 1: y = x - low(x) (the high-pass part the critic sees)
//   mode 2: y = low^T(x), the adjoint the generator's backward needs: the clamped (replicated) taps of border outputs fold back
//           onto the border pixels, so along each axis the border input p = 0 collects weight (R - q + 1) from output q <= R
//           (mirrored at p = H-1) and interior inputs weight 1 from |p - q| <= R.
// One thread per output element; rows are read coalesced and the window re-reads hit L1/L2 (8-33 MB per call: the pass is
// launch-latency sized).
__global__ void __launch_bounds__(256) lowpass_replicate_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                                int cs, long long images, int R, int mode) {
  const long long total = images * H * W * cs;
  const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cs);
    long long t = i / cs;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const float* p = x + (t / H) * (long long)H * W * cs + c;
    float s = 0.f;
    if (mode != 2) {
      for (int dy = -R; dy <= R; ++dy) {
        const int yy = min(max(h + dy, 0), H - 1);
        const float* row = p + (long long)yy * W * cs;
        for (int dx = -R; dx <= R; ++dx) s += __ldg(row + (long long)min(max(w + dx, 0), W - 1) * cs);
      }
      s *= inv;
      y[i] = mode == 1 ? __ldg(x + i) - s : s;
    } else {
      for (int q = max(h - R, 0); q <= min(h + R, H - 1); ++q) {
        float wy = 1.f;
        if (h == 0) wy = (float)(R - q + 1);
        if (h == H - 1) wy = (float)(R - (H - 1 - q) + 1);
        if (H == 1) wy = (float)(2 * R + 1);
        const float* row = p + (long long)q * W * cs;
        float rs = 0.f;
        for (int r = max(w - R, 0); r <= min(w + R, W - 1); ++r) {
          float wx = 1.f;
          if (w == 0) wx = (float)(R - r + 1);
          if (w == W - 1) wx = (float)(R - (W - 1 - r) + 1);
          if (W == 1) wx = (float)(2 * R + 1);
          rs += wx * __ldg(row + (long long)r * cs);
        }
        s += wy * rs;
      }
      y[i] = s * inv;
    }
  }
}

}  // namespace

int gather_rows(const float* src, const long long* idx_dev, int n_rows, long long row_elems, long long n_src, float* dst, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  Prof prof(PC_LAYOUT, 0.0, 8.0 * n_rows * (double)row_elems, st);
  const bool vec = row_elems % 4 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0;
  const long long per = vec ? row_elems / 4 : row_elems;
  // grid.x blocks per row sized so that the whole grid is a few waves of 148 SMs
  int bx = (int)std::min<long long>((per + 255) / 256, std::max<long long>(1, (148LL * 8 + n_rows - 1) / n_rows));
  bx = std::max(bx, 1);
  if (vec) gather_rows_kernel<<<dim3(bx, n_rows), 256, 0, st>>>((const float4*)src, idx_dev, (float4*)dst, per, n_src);
  else gather_rows_scalar_kernel<<<dim3(bx, n_rows), 256, 0, st>>>(src, idx_dev, dst, row_elems, n_src);
  DG_LAUNCH_CHECK();
  return 0;
}

constexpr int MS_BLOCKS = 148 * 4;
size_t metric_scratch_bytes() { return sizeof(double) * 2 * MS_BLOCKS; }
int metric_sums(const float* a, const float* b, long long n, const float* scores, int B, double* scratch, float* out, cudaStream_t st) {
  Prof prof(PC_L1, 0.0, 8.0 * (double)n, st);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(MS_BLOCKS, ((n >> 2) + MS_THREADS - 1) / MS_THREADS));
  metric_sums_kernel<<<blocks, MS_THREADS, 0, st>>>(a, b, n, scratch);
  DG_LAUNCH_CHECK();
  metric_finish_kernel<<<1, 256, 0, st>>>(scratch, blocks, n, scores, B, out);
  DG_LAUNCH_CHECK();
  return 0;
}

int lowpass_replicate(const float* x, float* y, long long images, int H, int W, int cs, int radius, int mode, cudaStream_t st) {
  const long long total = images * H * W * cs;
  Prof prof(PC_LAYOUT, 0.0, 8.0 * (double)total, st);
  const long long b = std::min<long long>((total + 255) / 256, 148LL * 16);
  lowpass_replicate_kernel<<<(unsigned)std::max<long long>(b, 1), 256, 0, st>>>(x, y, H, W, cs, images, radius, mode);
  DG_LAUNCH_CHECK();
  return 0;
}

}  // namespace dg

// ---- extern "C" (include/downgan_b200.h) -------------------------------------------------------------------------------
extern "C" int dg_gather_rows(const float* src, const int64_t* idx, int n_rows, int64_t row_elems, int64_t n_src, float* dst,
                              void* stream) {
  DG_CHECK(src && idx && dst && n_rows >= 0 && row_elems > 0 && n_src > 0, "dg_gather_rows: bad argument");
  return dg::gather_rows(src, (const long long*)idx, n_rows, row_elems, n_src, dst, (cudaStream_t)stream);
}

extern "C" int dg_lowpass(const float* x, float* y, int64_t planes, int h, int w, int filter_size, int mode, void* stream) {
  DG_CHECK(x && y && planes > 0 && h > 0 && w > 0 && filter_size >= 1 && (filter_size & 1), "dg_lowpass: bad argument (odd filter_size)");
  DG_CHECK(mode >= 0 && mode <= 2, "dg_lowpass: mode %d", mode);
  DG_CHECK(x != y, "dg_lowpass: in-place filtering is not supported");
  return dg::lowpass_replicate(x, y, planes, h, w, 1, filter_size / 2, mode, (cudaStream_t)stream);
}
