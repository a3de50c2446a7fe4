// Shared declarations for libdowngan_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/downgan_b200.h"

namespace dg {

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// DG_SYNC_CHECK=1 (diagnostic): every launch check synchronises the device, so an asynchronous fault is reported at the
// launch site of the kernel that caused it, whatever stream or launch API it used.
// DG_ABLATE=<bit mask> (diagnostic, tools/ablate.sh): the launchers of the named kernel families return without launching, so
// a bench run shows how much of the LIVE step (streams overlapped) each family really costs - results are garbage.
// bit 0 wgrad_ws, 1 wgrad_im2col (first layer), 2 trunk forward, 3 trunk backward, 4 batched dense wgrads, 5 conv_l1,
// 6 classifier tcgen05 kernels, 7 conv_ig, 8 conv_ws, 9 pack / unpack / adam, 10 column sums
inline bool ablate(int bit) {
  static const int mask = getenv("DG_ABLATE") ? atoi(getenv("DG_ABLATE")) : 0;
  return (mask >> bit) & 1;
}
inline bool sync_check() {
  static const bool on = getenv("DG_SYNC_CHECK") != nullptr;
  return on;
}

#define DG_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      dg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return DG_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define DG_CHECK(cond, ...)                 \
  do {                                      \
    if (!(cond)) {                          \
      dg::set_error(__VA_ARGS__);           \
      return DG_ERR_INVALID;                \
    }                                       \
  } while (0)

#define DG_TRY(expr)              \
  do {                            \
    int s__ = (expr);             \
    if (s__ != 0) return s__;     \
  } while (0)

#define DG_LAUNCH_CHECK()                                                                \
  do {                                                                                   \
    dg::count_launch();                                                                  \
    cudaError_t e__ = dg::sync_check() ? cudaDeviceSynchronize() : cudaPeekAtLastError(); \
    if (e__ != cudaSuccess) {                                                            \
      dg::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return DG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

// ---- optional per-launch profiling (bench.py's roofline leg) ----------------
// When enabled (dg_profile(1)) every wrapped launch is bracketed by CUDA events
// on its own stream; dg_profile_report sums the durations per kernel class.
enum ProfClass {
  PC_CONV_DIRECT = 0, PC_WGRAD_DIRECT, PC_CONV_UMMA, PC_WGRAD_UMMA, PC_DENSE_UMMA, PC_FC,
  PC_L1, PC_GP_NORMS, PC_ADAM, PC_INTERP, PC_LAYOUT, PC_COUNT
};
extern bool g_prof_on;
extern int g_tune[];  // dg_set_tuning switches
struct Prof {
  int idx = -1;
  cudaStream_t st;
  Prof(int cls, double flops, double bytes, cudaStream_t s);
  ~Prof();
};

// cuTensorMapEncodeTiled resolved through the runtime (cudaGetDriverEntryPoint): the library does not link
// libcuda, so it still loads (symbol check, host logic) on a machine without a driver.
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* addr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob);

// ---- tensor views ---------------------------------------------------------
// NHWC activation view: element (n,y,x,c) lives at p[((n*H + y)*W + x)*pitch + coff + c].
// A dense-block concat buffer is ONE allocation with pitch = 5F; each conv
// writes its output slice through a view with a different coff.
struct TV {
  void* p = nullptr;
  int bf = 0;     // 1: bf16 storage, 0: fp32
  int pitch = 0;  // channels per pixel in the allocation
  int coff = 0;   // first channel of this view
};
inline TV tv(void* p, int bf, int pitch, int coff = 0) {
  TV t; t.p = p; t.bf = bf; t.pitch = pitch; t.coff = coff; return t;
}
inline TV tv_batch(const TV& t, size_t pixels_per_sample, int n0) {  // view starting at sample n0
  TV r = t;
  size_t off = (size_t)n0 * pixels_per_sample * t.pitch;
  r.p = t.bf ? (void*)((bf16*)t.p + off) : (void*)((float*)t.p + off);
  return r;
}

enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_MASK = 2 };
enum Shuf { SHUF_NONE = 0, SHUF_PIXEL = 1, SHUF_UNPIXEL = 2 };

// One 3x3 / pad 1 convolution (or its data-gradient) with a fused epilogue:
//   v = s_acc*(acc + bias) + s1*r1 + s2*r2 ; act ; store (optionally pixel-(un)shuffled)
struct ConvOp {
  TV x; int Hin = 0, Win = 0, Ci = 0;
  TV y; int Hout = 0, Wout = 0, Co = 0;
  int B = 0;
  const float* w = nullptr;   // packed [9][Ci][CoP] fp32, CoP = round_up(Co,16)
  const float* bias = nullptr;
  int stride = 1;
  int transposed = 0;         // 1: data-gradient of a stride-2 conv (scatter expressed as gather)
  float s_acc = 1.f;
  TV r1; float s1 = 0.f;
  TV r2; float s2 = 0.f;
  int act = ACT_NONE; float slope = 0.f;
  TV mask;                    // ACT_MASK: v *= (mask > 0 ? 1 : slope); indexed like the STORE location
  int shuffle = SHUF_NONE;
  // tcgen05 path extras (bf16 mode)
  const void* w_umma = nullptr;  // packed bf16 weights for the tcgen05 kernel (null -> direct kernel)
  int narrow_ok = 0;             // Co < 16: w_umma holds a zero-padded 16-column image (set only where one is packed)
  const void* w_ig = nullptr;    // K-major bf16 weight image [tap][CoP][Ci] for the streaming implicit-GEMM kernel (null: not packed)
  // LeakyReLU sign bits (critic, bf16 mode): the data-gradient / JVP epilogues only need the SIGN of the saved activation, so
  // the forward epilogues that support it also store one uint16 per (pixel, 16-channel group) - bit j = channel 16k + j is
  // positive - and the mask epilogues read those 2 bytes instead of 32 bytes of bf16 activations.  Both pointers are already
  // offset to the op's first sample; words are indexed [pixel][Co / 16] like the store location.  Kernels that do not
  // implement them ignore the fields (bits_in: they use `mask`; bits_out: conv_writes_bits() tells the caller).
  unsigned short* bits_out = nullptr;      // ACT_LRELU ops
  const unsigned short* bits_in = nullptr; // ACT_MASK ops
};

struct WgradOp {
  TV x; int Hin = 0, Win = 0, Ci = 0;
  TV dy; int Hout = 0, Wout = 0, Co = 0;
  int B = 0;
  int stride = 1;
  float* dw = nullptr;   // packed [9][Ci][CoP] fp32, accumulated with atomics
  float* dbias = nullptr;  // [Co] accumulated, may be null
  int dbias_B = 0;         // > 0: the bias gradient sums the first dbias_B samples only (0: all B)
  // channel-block ops (dg_umma_wgrad_ws.cu): this op covers input channels [dw_ci_off, dw_ci_off + Ci) of a layer with
  // dw_ci_total input channels; dw rows are tap * dw_ci_total + dw_ci_off + ci   (0 / 0: the op is the whole layer)
  int dw_ci_total = 0, dw_ci_off = 0;
  // column-strip ops (dg_umma_wgrad_ws.cu): Wout / Win are the widths of ONE strip of a layer whose maps are Wout_full
  // (Wout_full * stride) columns wide; the strip starts at output column col0.  x and dy still point at the full tensors.
  // dW is accumulated, so the strips of a layer simply add up.  (0 / 0: the op covers whole rows.)
  int Wout_full = 0, col0 = 0;
};

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline size_t packed_w_elems(int ci, int co) { return (size_t)9 * ci * round_up(co, 16); }

// ---- kernels (dg_kernels.cu) ------------------------------------------------
int conv_direct(const ConvOp& op, cudaStream_t st);
int wgrad_direct(const WgradOp& op, cudaStream_t st);

// packing: table-driven, one launch per network
struct PackDesc {
  long long src_off;   // offset of OIHW weight in flat params (or flat grads for unpack)
  long long dst_off;   // offset in packed buffer
  int Ci, Co, CoP;
  int mode;            // 0: fwd [tap][ci][co] = W[co][ci][tap]
                       // 1: dgrad stride-1 (flipped, transposed): [tap][co][ci'] = W[co][ci][8-tap], packed as Ci_op=Co, Co_op=Ci
                       // 2: dgrad stride-2 gather form (not flipped): [tap][co][ci]
  // dense-block dgrad (mode 3): source slice of W_j feeding input slice k
  int slice_off;       // first input channel of the slice inside W (k*F)
  int src_ci_total;    // Ci of the source weight (j*F)
  int dst_row_off;     // row (input channel of the dgrad op) where this slice starts
  int dst_CoP;         // CoP of the destination op
  int umma;            // 1: also emit the bf16 tcgen05 B-operand image (same element offset) while packing
  int ps;              // 1: the layer feeds nn.PixelShuffle(2).  Its data-gradient dz is kept (i,j)-major - channel
                       // (2i + j) * Co/4 + c instead of 4c + 2i + j - so that the inverse-shuffle store of the layer above is one
                       // contiguous 32-byte store per pixel.  Consequences handled here: the weight / bias GRADIENTS come out in
                       // that column order (un-permuted while unpacking, modes 0 and 5) and the data-gradient weight image has its
                       // rows in that order (mode 1, packing).
};
__host__ __device__ inline unsigned ps_perm(unsigned co, unsigned Co) { return (co & 3u) * (Co >> 2) + (co >> 2); }
// packed_ig (optional): also emit the K-major bf16 image [tap][CoP][rows] of the streaming implicit-GEMM kernel for d.umma entries
int pack_weights(const float* params, float* packed, void* packed_umma, const PackDesc* table_dev, int n, int max_elems, cudaStream_t st,
                 void* packed_ig = nullptr);
int pack_weights2(const float* params, float* packed, void* packed_umma, const PackDesc* table_dev, int n, int max_elems, float* packed2,
                  void* packed_umma2, const PackDesc* table2_dev, int n2, int max_elems2, cudaStream_t st, void* packed_ig = nullptr,
                  void* packed_ig2 = nullptr);
int unpack_wgrads(const float* packed, float* grads, const PackDesc* table_dev, int n, int max_elems, cudaStream_t st);

int nchw_to_nhwc(const float* src, TV dst, int B, int C, int H, int W, cudaStream_t st);
int nhwc_to_nchw(TV src, float* dst, int B, int C, int H, int W, cudaStream_t st);
// dst = s * src (+ s2 * src2)   over B*H*W pixels x C channels of NHWC views
int scale_add(TV dst, TV a, float sa, TV b, float sb, size_t pixels, int C, cudaStream_t st);
// critic input batch [real ; fake ; alpha*real + (1-alpha)*fake] from NCHW real and NHWC/NCHW fake
int build_critic_input(const float* real_nchw, const float* fake, int fake_is_nchw, const float* alpha,
                       float* dst_nhwc, int B, int C, int H, int W, int with_interp, cudaStream_t st);
int colsum(TV dy, size_t pixels, int C, float* out, cudaStream_t st);
// Streams the library creates for overlapped work get their own column-sum scratch slot (1 or 2); every other
// stream shares slot 0 (one caller stream at a time per process, include/downgan_b200.h threading note).
void register_side_stream(cudaStream_t st, int slot);
void unregister_side_stream(cudaStream_t st);

// linear layers (critic classifier)
int fc_fwd(const void* x, int x_bf, const float* w, const float* bias, float* y, int NB, int K, int N,
           int act, float slope, const float* mask, cudaStream_t st);   // y = act(x W^T + b) or mask*(x W^T)
int fc_dgrad(const float* dz, const float* w, void* dx, int dx_bf, int NB, int K, int N,
             const void* mask, int mask_bf, float slope, cudaStream_t st);  // dx = (dz W) * lrelu'(mask)
int fc_wgrad(const float* dz, const void* x, int x_bf, float* dw, int NB, int K, int N, cudaStream_t st);  // dw += dz^T x
int critic_small_grads(const float* dz9, const float* seed, const float* a9, const float* vfc, int n0, int nv, int N, float* d_fc1b,
                       float* d_fc2w, float* d_fc2b, cudaStream_t st);
int fc2_fwd(const float* a, const float* w, const float* bias, float* s, int NB, int K, cudaStream_t st);
int fc2_seed(const float* a9, const float* w2, const float* seed, float* dz9, int NB, int K, float slope, cudaStream_t st);

int critic_means(const float* scores, int B, float* scalars, cudaStream_t st);
// tcgen05 classifier kernels (dg_umma_fc.cu): behind dg_set_tuning(14, 1), not yet validated on hardware
bool fc_umma_supported(int NB, int K, int N, int x_bf);
int fc_fwd_umma(const void* x, const float* w, float* y, int NB, int K, int N, cudaStream_t st);  // y (zeroed) += x w^T
int fc_dgrad_umma(const float* dz, const float* w, void* dx, int NB, int K, int N, const void* mask, float slope, cudaStream_t st);
int fc_wgrad_umma(const float* dz, const void* x, float* dw, int NB, int K, int N, cudaStream_t st);  // dw += dz^T x
bool fc_fwd_raw_supported(int K, int N);
int fc_fwd_raw(const void* x, int x_bf, const float* w, float* y, int NB, int K, int N, cudaStream_t st);  // y = x w^T
bool critic_head_supported(int B);
int critic_head(float* a9, const float* b1, const float* w2, const float* b2, float* scores, float* seed, float* dz9,
                float* scalars, int B, int K, float slope, cudaStream_t st);
int gp_norms(const float* g, int B, size_t per_sample, float* sumsq, cudaStream_t st);
// sum of squares per sample + norms / penalty value / dGP/dg coefficients in ONE launch (last block finishes, fixed-order combine)
int gp_norms_finish(const float* g, int B, size_t per_sample, float gp_lambda, float* sumsq, float* norms, float* coef, float* scalars,
                    int write_loss, cudaStream_t st);
int gp_finish(const float* sumsq, int B, float gp_lambda, float* norms, float* coef, float* scalars, int write_loss, cudaStream_t st);
int gp_scale(const float* g, const float* coef, float* u, int B, size_t per_sample, cudaStream_t st);
int l1_loss(const float* a, const float* b, long long n, float scale, float* loss_out, float* d_a,
            const float* d_add, cudaStream_t st);
int adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
         int step, float gscale, cudaStream_t st);
int fill(float* p, float v, long long n, cudaStream_t st);
int gen_scalars(const float* scores, int B, const float* l1, float gamma, float content_lambda, float* scalars, cudaStream_t st);

// tcgen05 path (dg_umma.cu)
bool umma_supported(const ConvOp& op);
int conv_umma(const ConvOp& op, cudaStream_t st);
// warp-specialised persistent variant (dg_umma_conv_ws.cu); DG_CONV_WS=0 in the environment falls back to conv_umma
bool umma_ws_supported(const ConvOp& op);
// tcgen05 forward conv of the few-channel fp32 boundary layer (dg_umma_conv_l1.cu)
bool conv_l1_supported(const ConvOp& op);
int conv_l1(const ConvOp& op, cudaStream_t st);
bool conv_l1p_supported(const ConvOp& op);  // planar variant without the im2col build (dg_umma_conv_l1p.cu)
int conv_l1p(const ConvOp& op, cudaStream_t st);
// TMA-fed weight gradient on swizzled NHWC tiles (dg_umma_wgrad_ws.cu)
bool wgrad_ws_supported(const WgradOp& op);
int wgrad_ws(const WgradOp& op, cudaStream_t st);
// many weight gradients in two launches (one per kernel mode, blockIdx.z = op; device table of plans + tensor maps)
size_t wgrad_ws_batch_bytes(int n_ops);
int wgrad_ws_batched(const WgradOp* ops, int n, void* table_dev, std::vector<unsigned char>& shadow, int S_per_op, cudaStream_t st);
// bias gradients of all dense blocks: out[(b*5 + k-1)*16 + c] += sum over rows of D_b[:, (5-k)*16 + c]   (F = 16)
int colsum_dense_blocks(void* const* d_bufs_dev, int n_blocks, size_t rows, float* out, cudaStream_t st);
int conv_umma_ws(const ConvOp& op, cudaStream_t st);
// bf16 re-pack of fp32 packed conv weights [tap][Ci][CoP] into the tcgen05 B-operand image
// [(tap*Ci/8 + ci/8)][CoP][8]; element offsets are shared with the fp32 packed buffer.
struct UmmaPackDesc { long long off; int Ci, CoP; };
bool wgrad_umma_supported(const WgradOp& op);
int wgrad_umma(const WgradOp& op, cudaStream_t st);
int pack_umma(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, cudaStream_t st);
// streaming implicit-GEMM conv (dg_umma_conv_ig.cu): both operands by TMA per (tap, channel block); weight image [tap][CoP][Ci]
int pack_ig(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, cudaStream_t st);
int pack_ig2(const float* packed, void* dst_bf16, const UmmaPackDesc* table_dev, int n, int max_elems, const float* packed2,
             void* dst2_bf16, const UmmaPackDesc* table2_dev, int n2, int max_elems2, cudaStream_t st);
bool conv_ig_supported(const ConvOp& op);
bool conv_ig_preferred(const ConvOp& op);
int conv_ig(const ConvOp& op, cudaStream_t st);

}  // namespace dg

namespace dg {
// boundary-layer kernels (dg_skinny.cu)
bool conv_skinny_supported(const ConvOp& op);
int conv_skinny(const ConvOp& op, cudaStream_t st);
bool wgrad_skinny_supported(const WgradOp& op);
int wgrad_skinny(const WgradOp& op, cudaStream_t st);
}  // namespace dg

namespace dg {
// persistent fused RRDB trunk, forward (dg_umma_trunk.cu)
bool trunk_fused_supported(int F, int Hc, int R, int bf);
int trunk_fwd_fused(const void* x_in, int in_pitch, int in_coff, void* y_out, int out_pitch, void* const* db_bufs_dev,
                    const void* w_umma, const float* bias, int R, int B, int save_count, cudaStream_t st);
}  // namespace dg

namespace dg {
// batched tcgen05 weight gradients: one launch for many layers (dg_umma_wgrad.cu)
size_t wgrad_umma_args_size();
int wgrad_umma_batched(const WgradOp* ops, int n, void* table_dev, std::vector<unsigned char>& shadow, int S_per_op,
                       cudaStream_t st);
}  // namespace dg

namespace dg {
int pack_trunk_slices(const float* pk_first_dense, void* dst_bf16, int n_db, int bwd, cudaStream_t st);
int trunk_bwd_fused(const void* g_in, void* g_out, void* const* fwd_bufs_dev, void* const* d_bufs_dev, const void* w_slices,
                    int R, int B, cudaStream_t st, int r0 = 0);
}  // namespace dg

namespace dg {
// tcgen05 weight gradient with the taps folded into N via an im2col tile (dg_umma_wgrad_im2col.cu)
bool wgrad_im2col_supported(const WgradOp& op);
int wgrad_im2col(const WgradOp& op, cudaStream_t st);
}  // namespace dg

namespace dg {
// data-path kernels (dg_data.cu)
int gather_rows(const float* src, const long long* idx_dev, int n_rows, long long row_elems, long long n_src, float* dst, cudaStream_t st);
size_t metric_scratch_bytes();
// out[0..4] = MAE(a,b), MSE(a,b), mean(scores[0:B]) - mean(scores[B:2B]), and the two means; deterministic two-stage reduction
int metric_sums(const float* a, const float* b, long long n, const float* scores, int B, double* scratch, float* out, cudaStream_t st);
// replicate-pad box filter on images with pixel stride cs (NHWC: cs = C; NCHW planes: cs = 1); mode 0 low, 1 high-pass, 2 adjoint
int lowpass_replicate(const float* x, float* y, long long images, int H, int W, int cs, int radius, int mode, cudaStream_t st);
}  // namespace dg
