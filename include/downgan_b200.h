/*
 * downgan_b200.h — C ABI of libdowngan_b200.so
 *
 * B200 (sm_100a) implementation of the DoWnGAN WGAN-GP training iteration.
 * The reference (nannau/DoWnGAN) has no FFI layer: its boundary is the Python
 * class surface.  Each entry point below names the reference statement(s) it
 * replaces (paths under /root/reference/DoWnGAN).  The Python host in
 * downgan_b200/ binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative dg_status otherwise;
 *     dg_last_error() returns a thread-local message for the last failure.
 *   - no exceptions cross the ABI, no hidden device synchronisation (the
 *     unit-test primitives dg_conv3x3_* are the exception: they allocate
 *     scratch and synchronise the stream before returning), every
 *     call enqueues on the cudaStream_t passed as `void* stream`.  A handle
 *     also owns one non-blocking side stream: the fused iterations fork
 *     work onto it (weight gradients, the second chain of a look-ahead
 *     forward) and join it back with events before returning, so towards
 *     the caller a call is still plain work in order on `stream`.
 *   - the caller owns all tensors handed in (device pointers unless stated);
 *     handles own only packed weights and activation workspaces.
 *   - user-facing tensors are contiguous NCHW fp32 (the reference layout);
 *     parameters / gradients are ONE flat fp32 buffer per network holding the
 *     tensors in the reference's state_dict() order, each OIHW / (out,in).
 *   - one handle per device and per Python module; not thread-safe per handle.
 */
#ifndef DOWNGAN_B200_H
#define DOWNGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DG_ABI_VERSION 2

typedef enum dg_status {
  DG_OK = 0,
  DG_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  DG_ERR_CUDA = -2,      /* CUDA runtime / driver error      */
  DG_ERR_STATE = -3,     /* call order (e.g. bwd without saved fwd) */
  DG_ERR_NOMEM = -4
} dg_status;

typedef enum dg_precision {
  DG_FP32 = 0,  /* fp32 storage, CUDA-core FFMA convs: the <=1e-3 parity mode */
  DG_BF16 = 1   /* bf16 storage, tcgen05/TMEM convs, fp32 accumulate          */
} dg_precision;

/* Constructor arguments of Generator (networks/generator.py:58). */
typedef struct dg_generator_config {
  int filters;         /* F; the reference passes the coarse grid edge (GAN/stage.py:60) */
  int channels;        /* input covariates */
  int n_predictands;   /* output channels (2: u10, v10) */
  int num_res_blocks;  /* RRDB count, default 16 */
  int num_upsample;    /* x2 pixel-shuffle stages, default 3 */
  int coarse_dim;      /* H = W of the coarse grid */
  int max_batch;
  int precision;       /* dg_precision */
} dg_generator_config;

/* Constructor arguments of Critic (networks/critic.py:12). */
typedef struct dg_critic_config {
  int coarse_dim;      /* base width (GAN/stage.py:59) */
  int fine_dim;        /* H = W of the fine grid, multiple of 16 */
  int nc;              /* input channels */
  int max_batch;       /* per-call batch; the fused critic step runs 3x this internally */
  int precision;       /* dg_precision */
} dg_critic_config;

/* Hyper-parameters (config/hyperparams.py:16-22). */
typedef struct dg_hyper {
  float gp_lambda;       /* 10; applied twice as in wasserstein.py:40,117 */
  float gamma;           /* 0.01 */
  float content_lambda;  /* 5 */
  int freq_sep;          /* hyperparams.py:31 (False): != 0 runs the iterations of GAN/wasserstein_fs.py:28-91 - the critic and
                            the penalty see x - low(x), the content loss compares low(fake) with low(fine) */
  int filter_size;       /* hyperparams.py:32 (5): low = AvgPool2d(filter_size, stride 1) o ReplicationPad2d(filter_size // 2) */
} dg_hyper;

typedef struct dg_generator dg_generator;
typedef struct dg_critic dg_critic;

const char* dg_last_error(void);
int dg_abi_version(void);
/* 1 if the library was compiled with the tcgen05 conv path for sm_100a. */
int dg_has_tcgen05(void);

/* ---- handles ---------------------------------------------------------- */
int dg_generator_create(const dg_generator_config* cfg, dg_generator** out);
int dg_generator_destroy(dg_generator* g);
int dg_critic_create(const dg_critic_config* cfg, dg_critic** out);
int dg_critic_destroy(dg_critic* c);

/* Number of fp32 elements of the flat parameter buffer, and the offset of
 * the i-th state_dict tensor in it (i in reference order; returns -1 past the end). */
int64_t dg_generator_param_count(const dg_generator_config* cfg);
int64_t dg_generator_param_offset(const dg_generator_config* cfg, int index);
int64_t dg_critic_param_count(const dg_critic_config* cfg);
int64_t dg_critic_param_offset(const dg_critic_config* cfg, int index);

/* Re-pack flat OIHW fp32 parameters into the kernels' layouts.  Must be
 * called after every parameter change (optimizer step / load_state_dict). */
int dg_generator_pack(dg_generator* g, const float* params_flat, void* stream);
int dg_critic_pack(dg_critic* c, const float* params_flat, void* stream);
/* Like dg_critic_pack, but only records `params`: the pack launch is issued by the next entry point that needs the weights;
 * the fused critic iterations run it beside their batch assembly.  `params` must stay valid and unchanged until then. */
int dg_critic_pack_lazy(dg_critic* c, const float* params);

/* ---- Generator.forward (networks/generator.py:83-90) -------------------
 * coarse: (B, channels, H, H) NCHW fp32 -> fake: (B, n_predictands, 8H, 8H).
 * save_for_backward != 0 keeps every dense-block buffer for dg_generator_bwd. */
int dg_generator_fwd(dg_generator* g, const float* coarse, int batch, float* fake,
                     int save_for_backward, void* stream);
/* autograd of the above: d_fake NCHW fp32 -> grads_flat (overwritten, same
 * layout as params_flat); d_coarse may be NULL. */
int dg_generator_bwd(dg_generator* g, const float* d_fake, float* grads_flat,
                     float* d_coarse, void* stream);

/* ---- Critic.forward (networks/critic.py:101-106) -----------------------
 * x: (B, nc, fine, fine) NCHW fp32 -> scores: (B) fp32. */
int dg_critic_fwd(dg_critic* c, const float* x, int batch, float* scores, void* stream);
/* autograd of the above: d_scores (B) -> grads_flat (overwritten), d_x NCHW (may be NULL). */
int dg_critic_bwd(dg_critic* c, const float* d_scores, float* grads_flat, float* d_x, void* stream);

/* ---- WassersteinGAN._gp (GAN/wasserstein.py:87-117) --------------------
 * real, fake NCHW fp32, alpha (B) fp32 (injectable; the reference draws it
 * with torch.rand).  Writes gp_out[0] = gp_lambda * mean((||g||-1)^2) (ONE
 * lambda, as _gp returns), norms (B, may be NULL), and — when grads_flat is
 * not NULL — d(gp_lambda * gp_out)/dparams (both lambdas, i.e. the term that
 * enters critic_loss at wasserstein.py:40,49), via the closed-form
 * double-backward (no autograd graph). */
int dg_gp(dg_critic* c, const dg_hyper* hp, const float* real, const float* fake, const float* alpha,
          int batch, float* gp_out, float* norms, float* grads_flat, void* stream);

/* ---- content_loss (GAN/losses.py:40-55) --------------------------------
 * loss_out[0] = mean |a-b| over n elements; if d_a != NULL writes
 * scale * sign(a-b) / n  (the seed of hp.content_lambda * content_loss). */
int dg_l1_loss(const float* a, const float* b, int64_t n, float scale, float* loss_out, float* d_a, void* stream);

/* ---- torch.optim.Adam.step (GAN/stage.py:63-64; wasserstein.py:55,83) --
 * one fused launch over a flat buffer; step is the 1-based step count.
 * grad_scale multiplies the gradient first (1/world_size after a sum-allreduce). */
int dg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* ---- data-parallel optimizer step (one process per GPU; replaces the host's dist.all_reduce(flat_grads) followed by
 * C_optimizer.step() / G_optimizer.step(), wasserstein.py:55,83 under data parallelism): ONE kernel sums the flat gradient
 * bucket over all ranks through NVLink peer memory (two-shot: every rank reduces one shard from all buckets and publishes
 * it to all buckets; NVSwitch multimem.ld_reduce / multimem.st when grad_multicast is set) and applies Adam with
 * grad_scale (1/world).  Every rank must call it in the same step with the same epoch (a counter the caller increments
 * per call and per flag block, starting at 1).  grad_ptrs[r] / flag_ptrs[r]: rank r's bucket / flag block as mapped into
 * THIS process (symmetric memory; [rank] are the local ones); a flag block is 32 zero-initialised uint32.  The local bucket
 * holds the global sum afterwards.  world <= 8 (one node).  Waits inside the kernel are bounded (trap after ~60 s). */
typedef struct dg_dp_peers {
  int rank, world;
  void* grad_ptrs[8];
  void* flag_ptrs[8];
  void* grad_multicast; /* multicast mapping of the bucket, or NULL: plain peer loads / stores */
} dg_dp_peers;
int dg_dp_allreduce_adam(const dg_dp_peers* peers, float* params, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, int step, float grad_scale, unsigned int epoch, void* stream);

/* ---- fused iterations ---------------------------------------------------
 * _critic_train_iteration (wasserstein.py:27-52, up to but excluding
 * C_optimizer.step): generator forward (no graph), critic on real / fake /
 * interpolates as one 3B batch, GP closed form, all critic gradients.
 * scalars_out (device, 8 floats): [0] critic_loss [1] c_real_mean
 * [2] c_fake_mean [3] gp (one lambda) [4] penalty (both lambdas) [5..7] 0.
 * c_grads_flat is overwritten. */
int dg_critic_step(dg_generator* g, dg_critic* c, const dg_hyper* hp,
                   const float* coarse, const float* fine, const float* alpha, int batch,
                   float* c_grads_flat, float* scalars_out, void* stream);
/* _generator_train_iteration (wasserstein.py:65-80, excluding G_optimizer.step).
 * scalars_out: [0] g_loss [1] c_fake_mean [2] l1 (unweighted) [3..7] 0. */
int dg_generator_step(dg_generator* g, dg_critic* c, const dg_hyper* hp,
                      const float* coarse, const float* fine, int batch,
                      float* g_grads_flat, float* scalars_out, void* stream);

/* Look-ahead for the n_critic schedule (wasserstein.py:131-147): the critic iterations between two generator
 * updates see the same generator weights, so fake = G(coarse) for the next `total` samples (several batches,
 * concatenated along the batch axis, total <= max_batch) is computed in ONE pass and kept in the generator's
 * output buffer until the next generator forward.  dg_critic_step_fake is dg_critic_step with the fake taken
 * from samples [fake_offset, fake_offset + batch) of that buffer instead of being recomputed.  The first
 * `save_first` samples also keep their activations, so that the generator iteration on that batch
 * (dg_generator_step_saved = dg_generator_step without the forward) needs no second forward. */
int dg_generator_lookahead(dg_generator* g, const float* coarse, int total, int save_first, void* stream);
/* Same pass, but only samples [first_offset, first_offset + first_count) - the batch of the NEXT critic iteration - are
 * computed in stream order; the rest of the pass is enqueued on the handle's side stream and overlaps that critic
 * iteration.  Every later call that touches the generator handle (dg_critic_step_fake for samples outside the first range,
 * dg_generator_step_saved, any other generator entry point) first joins the side stream, so callers need no extra
 * synchronisation.  Requires save_first <= first_offset; dg_set_tuning(12, 0) makes it identical to dg_generator_lookahead. */
int dg_generator_lookahead_first(dg_generator* g, const float* coarse, int total, int save_first, int first_offset,
                                 int first_count, void* stream);
int dg_critic_step_fake(dg_generator* g, dg_critic* c, const dg_hyper* hp, int fake_offset,
                        const float* fine, const float* alpha, int batch,
                        float* c_grads_flat, float* scalars, void* stream);
int dg_generator_step_saved(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* fine, int batch,
                            float* g_grads_flat, float* scalars_out, void* stream);

/* Data-parallel overlap of the gradient all-reduce with the backward (the one exchange step of the path).
 * After dg_critic_defer_conv_grads(c, 1), dg_critic_step / dg_critic_step_fake return as soon as the classifier gradients
 * c_grads_flat[dg_critic_param_offset(cfg, 9) ..) are final in stream order, while the conv weight gradients still run on
 * the handle's side stream; the caller starts the all-reduce of that slice and calls dg_critic_step_finish, which joins the
 * side stream and writes the conv gradients c_grads_flat[0 .. offset).  Until then every other critic entry point
 * returns DG_ERR_STATE. */
int dg_critic_defer_conv_grads(dg_critic* c, int on);
int dg_critic_step_finish(dg_critic* c, float* c_grads_flat, void* stream);

/* ---- unit-testable conv primitives (NCHW fp32 in/out, OIHW fp32 weights)
 * 3x3, padding 1, stride 1 or 2.  `precision` selects the kernel family.
 * Replaces the cudnn_convolution / convolution_backward calls behind
 * networks/generator.py:24 and networks/critic.py:21-86. */
int dg_conv3x3_fwd(const float* x, const float* w, const float* bias, float* y,
                   int batch, int ci, int co, int hin, int win, int stride, float lrelu_slope,
                   int precision, void* stream);
int dg_conv3x3_dgrad(const float* dy, const float* w, float* dx,
                     int batch, int ci, int co, int hin, int win, int stride, int precision, void* stream);
int dg_conv3x3_wgrad(const float* x, const float* dy, float* dw, float* dbias,
                     int batch, int ci, int co, int hin, int win, int stride, int precision, void* stream);

/* ---- parity instrumentation (tests/ and tools/parity_report.py; no reference counterpart) ----------------
 * Copies an activation the handle still holds out as NCHW fp32 (`out` is a device pointer), so that a test can
 * take the LeakyReLU sign masks the CUDA path actually used (sign of the stored post-activation) and replay
 * them in the CPU oracle: with the masks pinned, every parameter gradient must meet north_star's tolerance;
 * without, bf16 rounding flips masks of near-zero pre-activations (profiles/parity_r02.md).
 * dg_generator_activation, after a forward that saved (dg_generator_fwd save=1 / dg_generator_step):
 *   which 0 .. 3R-1   dense-block concat buffer (batch, 5F, Hc, Hc): slices [x, o1, o2, o3, o4]
 *   which 1000        trunk output (batch, F, Hc, Hc)        1001  out1 + conv2(trunk) (batch, F, Hc, Hc)
 *   which 1100 + u    upsample stage u AFTER PixelShuffle (batch, F, Hc << (u+1), Hc << (u+1))
 *   which 1200        conv3.0 output (batch, F, Hf, Hf)
 * dg_critic_activation, samples [s0, s0 + batch) of the critic's last batch:
 *   which 1 .. 8      output of features.{2(which-1)} after LeakyReLU (batch, Co, H, H)
 *   which 9           classifier.0 output after LeakyReLU (batch, 100)
 *   which 101 .. 108  the interpolates' activations of the last fused critic iteration, samples [s0, s0+batch)
 *                     of its B interpolates (the iteration overwrites them in place with the JVP chain; they are
 *                     kept aside only while dg_set_tuning(15, 1) is on). */
int dg_generator_activation(dg_generator* g, int which, int batch, float* out, void* stream);
int dg_critic_activation(dg_critic* c, int which, int s0, int batch, float* out, void* stream);

/* The RRDB trunk alone (networks/generator.py:36-53, the res_blocks Sequential at :85): x (batch, F, Hc, Hc) NCHW
 * fp32 -> y same shape, on the handle's packed weights; activations are saved.  dg_generator_trunk_bwd is its
 * autograd: d_y -> d_x (may be NULL) and grads_flat (whole generator layout, overwritten; only the res_blocks
 * tensors are non-zero).  Unit-test entry points for the fused persistent trunk kernels. */
int dg_generator_trunk_fwd(dg_generator* g, const float* x, int batch, float* y, void* stream);
int dg_generator_trunk_bwd(dg_generator* g, const float* d_y, float* d_x, float* grads_flat, void* stream);

/* ---- per-batch metric pass (mlflow_tools/mlflow_epoch.py:53-63 gen_batch_and_log_metrics) -----------------------
 * out8 (device): [0] MAE = content_loss(real, fake) (GAN/losses.py:40-55)  [1] MSE = content_MSELoss (:58-68)
 * [2] Wass = wass_loss(mean C(real), mean C(fake)) = their difference (:8-9)  [3] mean C(real)  [4] mean C(fake)  [5..7] 0,
 * with fake = G(coarse) and both networks' CURRENT weights.  coarse == NULL: fake is taken from samples
 * [fake_offset, fake_offset + batch) of the last look-ahead pass (valid while the generator has not been updated since).
 * MS-SSIM (losses.py:12-38) is computed by the third-party pytorch_msssim package in the reference and is not provided. */
int dg_metrics(dg_generator* g, dg_critic* c, const float* coarse, int fake_offset, const float* fine, int batch,
               float* out8, void* stream);

/* ---- batch assembly (GAN/dataloader.py:25-33 __getitem__ per sample + torch's default collate) -------------------
 * dst[r] = src[idx[r]] for r < n_rows; a row is row_elems contiguous floats, src holds n_src rows, idx is a DEVICE array of
 * int64 row numbers (the shuffled order of the epoch).  One coalesced gather from the HBM-resident dataset. */
int dg_gather_rows(const float* src, const int64_t* idx, int n_rows, int64_t row_elems, int64_t n_src, float* dst, void* stream);

/* ---- frequency-separation filter (config/hyperparams.py:31-35; GAN/wasserstein_fs.py:41-47) ----------------------
 * x, y: `planes` contiguous (h, w) fp32 images (NCHW tensors: planes = N*C).  mode 0: y = low(x) = AvgPool2d(filter_size, 1)
 * (ReplicationPad2d(filter_size // 2)(x));  mode 1: y = x - low(x);  mode 2: y = low^T(x) (adjoint, used by the backward). */
int dg_lowpass(const float* x, float* y, int64_t planes, int h, int w, int filter_size, int mode, void* stream);

/* Number of kernel launches this library has enqueued since load (for bench.py's gpu_launches). */
int64_t dg_launch_count(void);

/* Per-launch CUDA-event timing for bench.py's roofline leg (no reference counterpart).
 * dg_profile(1) clears and starts recording, dg_profile(0) stops.  dg_profile_report
 * synchronises the device and writes, per kernel class c < DG_PROFILE_CLASSES,
 * out[4c..4c+3] = {launches, total milliseconds, algorithmic FLOPs, algorithmic bytes}.
 * Classes: 0 conv_direct 1 wgrad_direct 2 conv_tcgen05 3 wgrad_tcgen05 4 dense_block_tcgen05
 * 5 linear 6 l1_loss 7 gp_norm 8 adam 9 interpolate 10 layout. */
#define DG_PROFILE_CLASSES 11
int dg_profile(int enable);
int dg_profile_report(double* out, int n_classes);

/* Kernel-selection switches for A/B measurements (no reference counterpart); every switch defaults to
 * the fastest validated path.  key 0: warp-specialised tcgen05 conv (1) or the per-tile kernel (0);
 * key 1: device-wide cudaFuncCachePreferShared (1) or no preference (0); key 2: TMA-fed weight-gradient
 * kernel (1) or the cp.async kernels (0); key 3: trunk MMA issue order (0, 1, 2); key 4: dense-block weight
 * gradients on the TMA-fed kernel as channel blocks (1) or the cp.async batched kernel (0); key 5: programmatic
 * dependent launch of the TMA-fed conv / weight-gradient kernels (1) or plain stream order (0); key 6: tcgen05 kernel
 * for the few-channel fp32 boundary conv (1) or the CUDA-core kernel (0); key 7: narrow-output convs (Co < 16) on the
 * TMA-fed kernel with zero pad columns (1) or the CUDA-core kernel (0); key 8: features.0 bias gradient folded into the
 * first-layer weight-gradient kernel as an im2col column of ones (1) or a separate column-sum launch (0); key 9: weight
 * gradients (and their bias column sums) of the fused iterations on a library-owned side stream, forked and joined with
 * events inside the call, so they overlap the data-gradient / JVP chain (1) or everything in stream order (0).
 * key 9 == 2: the critic iteration as two chains over disjoint sample ranges (measured slower than 1).  key 10 = k > 0:
 * generator forward of more than 296 samples (look-ahead pass) split so that its last 32k samples run as a second
 * chain on the side stream (0: one chain).  key 11: classifier head of the fused critic iteration (bias + LeakyReLU,
 * classifier.2, score means, loss seeds, dz of the hidden layer) in one launch (1) or five (0).
 * key 12: dg_generator_lookahead_first defers everything but the first range to the side stream (1) or computes the whole
 * pass in stream order (0).
 * key 13 = k > 0 (with key 9 == 1): the real + fake rows of the weight gradients of critic layers 0 .. k-1 are enqueued on
 * the side stream during the input-gradient chain, their interpolates' rows during the JVP chain (0: one 3B launch per layer).
 * key 14: classifier.0 products (forward, input gradient, weight gradient) on the tcgen05 kernels of csrc/dg_umma_fc.cu (1)
 * (1, default: validated in round 2, tests/test_gpu_fc_umma.py; +1.8 % on the cfg-2 step) or on the CUDA-core kernels (0).
 * key 15: parity instrumentation - the fused critic iteration keeps a copy of the interpolates' activations for
 * dg_critic_activation(101..108) (0, default: off, no copy).
 * key 16: streaming implicit-GEMM conv kernel (csrc/dg_umma_conv_ig.cu) on the shapes where it is the faster tcgen05 kernel
 * (1, default) or never (0: every tcgen05 conv on the weights-stationary kernel, the round-1 behaviour).
 * key 17: the TMA producer of the weights-stationary conv kernel fills its ring before the CTA stages its weights and
 * synchronises (1) or after (0, default: measured -0.5 % on the cfg-2 step, profiles/README.md).
 * key 18 = k (1..4): the fused trunk backward runs as k launches over RRDB ranges; the dense convs' weight gradients of a range
 * go to the side stream and overlap the next range's data gradients (default 2; 1 = one launch, weight gradients afterwards).
 * key 19: the critic's late-layer weight gradients alternate between two side streams (1) or share one (0, default).
 * key 20: the 1- / 2-channel first-layer convolutions run on the planar kernel without an im2col build (1;
 * csrc/dg_umma_conv_l1p.cu: faster alone, measured 1.4 % slower in the overlapped cfg-2 step) or on the im2col kernel (0, default).
 * key 21: the critic's forward epilogues store LeakyReLU sign bits (2 bytes per pixel and 16 channels) and its data-gradient / JVP
 * epilogues read them instead of the saved bf16 activations (1, default) or not (0); 2 = as 1, and the weights-stationary kernel
 * fetches the sign words of a whole tile before it waits for the tile's MMAs (measured slower: the extra address arithmetic
 * lengthens the epilogue warps' instruction chains, 43.1 k vs 43.9 k samples/s).
 * key 22: the weights-stationary conv kernel uses its specialised epilogue (no bias / residual / shuffle checks) where the
 * layer allows it (1, default) or always the general one (0).
 * key 23: the two batched dense-block weight-gradient launches of the LAST trunk-backward range run beside each other on two
 * streams (1, default) or one after the other (0).
 * key 24: the fused critic iteration unpacks the classifier gradients while its conv weight gradients still run on the side
 * stream and only the nine conv entries after the join (1) or everything after the join (0, default: the extra launch costs
 * more than it hides, 44 424 vs 44 577 samples/s in a same-box A/B).
 * key 25 (with dg_critic_pack_lazy): the pending weight pack runs beside the critic iteration's batch assembly (1, default) or
 * in front of it (0).
 * Returns the previous value, or DG_ERR_INVALID for an unknown key. */
#define DG_TUNE_KEYS 26
int dg_set_tuning(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* DOWNGAN_B200_H */
