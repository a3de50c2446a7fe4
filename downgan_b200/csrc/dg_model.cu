// Host-side orchestration of the generator, the critic and the fused WGAN-GP
// iterations, plus the extern "C" surface declared in include/downgan_b200.h.
// Reference statements replaced are cited per function (paths under
// /root/reference/DoWnGAN).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "dg_common.cuh"

using namespace dg;

namespace {

constexpr float G_SLOPE = 0.01f;   // nn.LeakyReLU() default, networks/generator.py:26,72,79
constexpr float C_SLOPE = 0.2f;    // networks/critic.py:24..87,97
constexpr float RES_SCALE = 0.2f;  // networks/generator.py:19,45
constexpr int FC_HIDDEN = 100;     // networks/critic.py:95

struct Layer {
  int Ci = 0, Co = 0, stride = 1;
  long long w_off = 0, b_off = -1;    // offsets in the flat parameter buffer
  long long pk_off = 0, pkb_off = -1; // offsets in the packed buffers (weights / bias area)
  long long pkd_off = -1;             // packed data-gradient weights
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// A library-owned stream for work that only has to be finished by the end of an API call (weight gradients and
// their bias column sums): forked from / joined to the caller's stream with events, so the call keeps plain
// stream-order semantics for the caller.  Non-blocking, so it also overlaps a caller on the legacy default stream.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr, aux_ev = nullptr;
  bool dirty = false;  // work has been enqueued on `s` since the last join
  int create(int colsum_slot) {
    DG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    DG_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&aux_ev, cudaEventDisableTiming));
    register_side_stream(s, colsum_slot);
    return 0;
  }
  void destroy() {
    if (s) cudaStreamSynchronize(s);  // a deferred chain may still be using the handle's buffers
    if (s) unregister_side_stream(s);
    if (fork_ev) cudaEventDestroy(fork_ev);
    if (join_ev) cudaEventDestroy(join_ev);
    if (aux_ev) cudaEventDestroy(aux_ev);
    if (s) cudaStreamDestroy(s);
    s = nullptr; fork_ev = join_ev = aux_ev = nullptr;
  }
  // everything enqueued on `st` so far happens before what is enqueued on the side stream next
  int fork(cudaStream_t st) {
    DG_CUDA(cudaEventRecord(fork_ev, st));
    DG_CUDA(cudaStreamWaitEvent(s, fork_ev, 0));
    dirty = true;
    return 0;
  }
  // everything enqueued on the side stream so far happens before what is enqueued on `st` next
  int join(cudaStream_t st) {
    if (!dirty) return 0;
    DG_CUDA(cudaEventRecord(join_ev, s));
    DG_CUDA(cudaStreamWaitEvent(st, join_ev, 0));
    dirty = false;
    return 0;
  }
};

int dev_alloc(std::vector<void*>& pool, void** out, size_t bytes) {
  if (bytes == 0) bytes = 16;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
    return DG_ERR_NOMEM;
  }
  e = cudaMemset(p, 0, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMemset -> %s", cudaGetErrorString(e));
    return DG_ERR_CUDA;
  }
  pool.push_back(p);
  *out = p;
  return 0;
}

// DG_LOG_FALLBACK=1: one stderr line per distinct shape that leaves the tcgen05 kernels for the CUDA-core ones (diagnostic)
void log_fallback(const char* what, int ci, int co, int h, int w, int b, int stride, int bf) {
  static const bool on = getenv("DG_LOG_FALLBACK") != nullptr;
  if (!on) return;
  static std::vector<long long> seen;
  const long long key = ((((long long)ci * 1024 + co) * 4096 + h) * 4 + stride) * 2 + (what[0] == 'w');
  if (std::find(seen.begin(), seen.end(), key) != seen.end()) return;
  seen.push_back(key);
  fprintf(stderr, "[dg fallback] %s on CUDA cores: Ci %d Co %d %dx%d B %d stride %d %s\n", what, ci, co, h, w, b, stride, bf ? "bf16" : "fp32");
}

int run_wgrad(const WgradOp& w0, cudaStream_t st) {
  WgradOp w = w0;
  // Wide or non-power-of-two channel counts (the dense-block layers of F = 32 / 64 generators: Ci = 96, 160, 192, 256, 320):
  // the TMA-fed kernel takes power-of-two channel blocks <= 128, each writing its own rows of dW (as the batched launch of
  // the F = 16 trunk does); the bias gradient rides on the first block.
  if (g_tune[2] && w.x.bf && w.dy.bf && w.Ci % 16 == 0 && w.dw_ci_total == 0 && (w.Ci > 128 || (w.Ci & (w.Ci - 1)))) {
    std::vector<WgradOp> blocks;
    bool ok = true;
    for (int c0 = 0; c0 < w.Ci && ok;) {
      int cb = 128;
      while (cb > w.Ci - c0) cb >>= 1;
      WgradOp o = w;
      o.x.coff = w.x.coff + c0; o.Ci = cb; o.dw_ci_total = w.Ci; o.dw_ci_off = c0;
      if (c0 > 0) o.dbias = nullptr;
      ok = wgrad_ws_supported(o);
      blocks.push_back(o);
      c0 += cb;
    }
    if (ok) {
      for (const WgradOp& o : blocks) DG_TRY(wgrad_ws(o, st));
      return 0;
    }
  }
  // Maps too wide for one TMA box / one shared-memory stage (the 256-wide layers of cfg-4: conv3.0, the critic's stride-2
  // layers on 256x256 .. 64x64 maps at 32 .. 128 channels): 2 or 4 column strips, each a launch of the TMA-fed kernel that
  // adds into the same dW.
  if (g_tune[2] && w.x.bf && w.dy.bf && w.dw_ci_total == 0 && w.Wout_full == 0 && !wgrad_ws_supported(w) && w.Ci >= 16 && w.Ci <= 128 &&
      (w.Ci & (w.Ci - 1)) == 0 && w.Co % 16 == 0) {
    for (int ns = 2; ns <= 8; ns *= 2) {
      if (w.Wout % ns) break;
      WgradOp o = w;
      o.Wout_full = w.Wout; o.Wout = w.Wout / ns; o.Win = w.Win / ns; o.col0 = 0; o.dbias = nullptr;
      if (!wgrad_ws_supported(o)) continue;
      for (int k = 0; k < ns; ++k) {
        o.col0 = k * o.Wout;
        DG_TRY(wgrad_ws(o, st));
      }
      if (w.dbias) {
        const int nb = (w.dbias_B > 0 && w.dbias_B < w.B) ? w.dbias_B : w.B;
        DG_TRY(colsum(w.dy, (size_t)nb * w.Hout * w.Wout, w.Co, w.dbias, st));
      }
      return 0;
    }
  }
  // bias gradient over a leading part of the batch: only the first-layer tcgen05 kernel folds it in; elsewhere it is
  // a column sum over those samples
  const bool ws_path = g_tune[2] && wgrad_ws_supported(w);
  if (w.dbias && w.dbias_B > 0 && w.dbias_B < w.B && (ws_path || !wgrad_im2col_supported(w))) {
    DG_TRY(colsum(w.dy, (size_t)w.dbias_B * w.Hout * w.Wout, w.Co, w.dbias, st));
    w.dbias = nullptr;
  }
  if (ws_path) return wgrad_ws(w, st);
  if (wgrad_im2col_supported(w)) return wgrad_im2col(w, st);
  if (wgrad_skinny_supported(w)) return wgrad_skinny(w, st);
  if (wgrad_umma_supported(w)) return wgrad_umma(w, st);
  log_fallback("wgrad", w.Ci, w.Co, w.Hin, w.Win, w.B, w.stride, w.x.bf);
  return wgrad_direct(w, st);
}

// wrote_bits (optional): set to whether the kernel that took the op stores ConvOp::bits_out (the im2col first-layer kernel and the
// weights-stationary kernel do; the others ignore the field)
int run_conv(const ConvOp& op, cudaStream_t st, bool* wrote_bits = nullptr) {
  static const bool ws = !(getenv("DG_CONV_WS") && atoi(getenv("DG_CONV_WS")) == 0);
  if (wrote_bits) *wrote_bits = false;
  if (g_tune[6] && g_tune[20] && conv_l1p_supported(op)) return conv_l1p(op, st);
  if (g_tune[6] && conv_l1_supported(op)) { if (wrote_bits) *wrote_bits = op.bits_out != nullptr; return conv_l1(op, st); }
  if (op.Co < 16 && op.narrow_ok && ws && g_tune[0] && g_tune[7] && op.w_umma && umma_ws_supported(op))
    return conv_umma_ws(op, st);  // narrow output on tensor cores
  if (conv_skinny_supported(op)) return conv_skinny(op, st);
  if (op.w_ig && conv_ig_preferred(op)) return conv_ig(op, st);  // late critic layers / wide dense layers: streaming implicit GEMM
  if (ws && g_tune[0] && op.w_umma && umma_ws_supported(op)) {
    if (wrote_bits) *wrote_bits = op.bits_out != nullptr && op.act == ACT_LRELU && op.shuffle == SHUF_NONE && op.Co % 16 == 0;
    return conv_umma_ws(op, st);
  }
  if (op.w_umma && umma_supported(op)) return conv_umma(op, st);
  log_fallback("conv", op.Ci, op.Co, op.Hout, op.Wout, op.B, op.stride, op.x.bf);
  return conv_direct(op, st);
}

}  // namespace

// ===========================================================================
// Generator
// ===========================================================================
struct dg_generator {
  dg_generator_config cfg{};
  int bf = 0, F = 0, Hc = 0, R = 0, U = 0, Cin = 0, Cout = 0, Hf = 0, maxB = 0;
  size_t esz = 4;
  std::vector<Layer> layers;  // state_dict order: conv1, R*3*5 dense convs, conv2, U upsample convs, conv3.0, conv3.2
  long long n_params = 0;
  std::vector<void*> pool;
  // packed parameters
  float* pk = nullptr;       // forward weights + bias area
  float* pkd = nullptr;      // data-gradient weights
  float* gpk = nullptr;      // packed gradients (same offsets as pk)
  long long pk_elems = 0, pkd_elems = 0;
  PackDesc *tab_fwd = nullptr, *tab_dgrad = nullptr;
  int n_fwd = 0, n_dgrad = 0, max_fwd = 0, max_dgrad = 0;
  bf16 *pk_u = nullptr, *pkd_u = nullptr;  // tcgen05 B-operand images (bf16 mode)
  bf16 *pk_ig = nullptr, *pkd_ig = nullptr;  // K-major images [tap][CoP][Ci] for the streaming implicit-GEMM kernel
  bf16* pk_trunk = nullptr;                // slice-major images of the dense convs for the fused trunk kernel
  bf16* pkd_trunk = nullptr;               // same for the dense data-gradient matrices (fused trunk backward)
  void** d_ptrs_dev = nullptr;             // device copy of Dall[]
  UmmaPackDesc *utab_fwd = nullptr, *utab_dgrad = nullptr, *utab_igfwd = nullptr;
  int n_ufwd = 0, n_udgrad = 0, max_ufwd = 1, max_udgrad = 1, n_igfwd = 0;
  std::vector<long long> db_dgrad_off;  // [(r*3+d)*5 + k] packed offset of Wt_k
  bool packed = false;
  bool use_ig = false;  // some layer may run on the streaming implicit-GEMM kernel (F >= 32 or maps wider than a TMA halo box):
                        // the F = 16 / 128x128 generator never does, so its K-major weight images are not packed at all
  // activations
  void* x0 = nullptr;
  std::vector<void*> db;  // R*3 concat buffers (B,Hc,Hc,5F)
  void** db_ptrs_dev = nullptr;  // device copy of db[] for the fused trunk kernel
  void *trunk_out = nullptr, *t1 = nullptr, *c30 = nullptr;
  std::vector<void*> up;  // U post-shuffle activations
  float* fake = nullptr;  // (B,Hf,Hf,Cout) NHWC fp32
  // backward workspaces
  std::vector<void*> Dall;        // per-dense-block dz buffers (bf16 mode: weight gradients are batched after the dgrad chain)
  void* wg_table_dev = nullptr;   // device table of the batched weight-gradient launch
  std::vector<unsigned char> wg_shadow;
  void* wgws_table_dev = nullptr; // same for the TMA-fed kernel (plans + tensor maps)
  std::vector<unsigned char> wgws_shadow;
  std::vector<std::vector<unsigned char>> wgws_parts;  // host shadows of the per-range tables (trunk backward in RRDB ranges)
  void *D = nullptr, *gR = nullptr, *gx0 = nullptr, *gx1 = nullptr, *gT1 = nullptr, *gA = nullptr, *gB = nullptr;
  std::vector<void*> gU;   // gU[u]: dz of upsample stage u, (B, Hc<<u, Hc<<u, 4F); one buffer per stage (no ping-pong), so
                           // weight gradients on the side stream never race a later data-gradient store; gU[U-1] == gB
  SideStream side;
  float* dfake = nullptr;  // NHWC fp32
  float* fine_nhwc = nullptr;
  float* l1 = nullptr;
  int saved_batch = 0;
  int lookahead = 0;  // samples of g->fake produced by the last dg_generator_lookahead (0: none / overwritten)
  int ready_lo = 0, ready_hi = 0;  // dg_generator_lookahead_first: samples already final in stream order while side.dirty

  int idx_conv1() const { return 0; }
  int idx_db(int r, int d, int k) const { return 1 + (r * 3 + d) * 5 + (k - 1); }
  int idx_conv2() const { return 1 + 15 * R; }
  int idx_up(int u) const { return 2 + 15 * R + u; }
  int idx_c30() const { return 2 + 15 * R + U; }
  int idx_c32() const { return 3 + 15 * R + U; }
  TV act(void* p, int pitch, int coff = 0) const { return tv(p, bf, pitch, coff); }
};

static void gen_enumerate(const dg_generator_config& c, std::vector<Layer>& L, long long& n_params) {
  L.clear();
  long long off = 0;
  auto add = [&](int ci, int co) {
    Layer l;
    l.Ci = ci; l.Co = co; l.stride = 1;
    l.w_off = off; off += (long long)co * ci * 9;
    l.b_off = off; off += co;
    L.push_back(l);
  };
  const int F = c.filters;
  add(c.channels, F);
  for (int r = 0; r < c.num_res_blocks; ++r)
    for (int d = 0; d < 3; ++d)
      for (int k = 1; k <= 5; ++k) add(k * F, F);
  add(F, F);
  for (int u = 0; u < c.num_upsample; ++u) add(F, 4 * F);
  add(F, F);
  add(F, c.n_predictands);
  n_params = off;
}

extern "C" int64_t dg_generator_param_count(const dg_generator_config* cfg) {
  if (!cfg) return -1;
  std::vector<Layer> L; long long n;
  gen_enumerate(*cfg, L, n);
  return n;
}
extern "C" int64_t dg_generator_param_offset(const dg_generator_config* cfg, int index) {
  if (!cfg || index < 0) return -1;
  std::vector<Layer> L; long long n;
  gen_enumerate(*cfg, L, n);
  const int li = index / 2;
  if (li >= (int)L.size()) return -1;
  return (index & 1) ? L[li].b_off : L[li].w_off;
}

static int upload_table(std::vector<void*>& pool, const std::vector<PackDesc>& t, PackDesc** dev) {
  DG_TRY(dev_alloc(pool, (void**)dev, sizeof(PackDesc) * std::max<size_t>(t.size(), 1)));
  if (!t.empty()) DG_CUDA(cudaMemcpy(*dev, t.data(), sizeof(PackDesc) * t.size(), cudaMemcpyHostToDevice));
  return 0;
}

static int upload_utable(std::vector<void*>& pool, const std::vector<UmmaPackDesc>& t, UmmaPackDesc** dev) {
  DG_TRY(dev_alloc(pool, (void**)dev, sizeof(UmmaPackDesc) * std::max<size_t>(t.size(), 1)));
  if (!t.empty()) DG_CUDA(cudaMemcpy(*dev, t.data(), sizeof(UmmaPackDesc) * t.size(), cudaMemcpyHostToDevice));
  return 0;
}
static inline bool umma_ok(int ci, int co) { return ci % 16 == 0 && co % 16 == 0 && co <= 256; }
// operand image usable by the TMA conv kernel: narrow outputs (Co < 16) run as N = 16 with zero pad columns
static inline bool umma_img_ok(int ci, int co) { return ci % 16 == 0 && co <= 256 && (co % 16 == 0 || co < 16); }

extern "C" int dg_generator_create(const dg_generator_config* cfg, dg_generator** out) {
  DG_CHECK(cfg && out, "dg_generator_create: null argument");
  DG_CHECK(cfg->filters >= 1 && cfg->channels >= 1 && cfg->n_predictands >= 1 && cfg->num_res_blocks >= 0 &&
               cfg->num_upsample >= 0 && cfg->num_upsample <= 5 && cfg->coarse_dim >= 1 && cfg->max_batch >= 1,
           "dg_generator_create: bad config");
  DG_CHECK(cfg->precision == DG_FP32 || cfg->precision == DG_BF16, "dg_generator_create: precision %d", cfg->precision);
  dg_generator* g = new dg_generator();
  g->cfg = *cfg;
  g->bf = cfg->precision == DG_BF16;
  g->esz = g->bf ? 2 : 4;
  g->F = cfg->filters; g->Hc = cfg->coarse_dim; g->R = cfg->num_res_blocks; g->U = cfg->num_upsample;
  g->Cin = cfg->channels; g->Cout = cfg->n_predictands; g->Hf = g->Hc << g->U; g->maxB = cfg->max_batch;
  gen_enumerate(*cfg, g->layers, g->n_params);
  const int F = g->F;
  g->use_ig = g->bf && (F >= 32 || g->Hf > 254);
  // ---- packed layouts + tables
  std::vector<PackDesc> tf, td;
  auto is_up = [&](long li) { return (li >= g->idx_up(0) && li < g->idx_up(0) + g->U) ? 1 : 0; };  // layers feeding nn.PixelShuffle
  long long pk = 0, pkd = 0;
  int maxf = 1, maxd = 1;
  for (auto& l : g->layers) {
    l.pk_off = pk; pk += (long long)packed_w_elems(l.Ci, l.Co);
    PackDesc d{}; d.src_off = l.w_off; d.dst_off = l.pk_off; d.Ci = l.Ci; d.Co = l.Co; d.CoP = round_up(l.Co, 16); d.mode = 0;
    d.umma = g->bf && umma_img_ok(l.Ci, l.Co);
    d.ps = is_up(&l - g->layers.data());
    tf.push_back(d);
    maxf = std::max(maxf, l.Ci * l.Co * 9);
  }
  for (auto& l : g->layers) {
    l.pkb_off = pk; pk += round_up(l.Co, 16);
    PackDesc d{}; d.src_off = l.b_off; d.dst_off = l.pkb_off; d.Co = l.Co; d.Ci = 1; d.mode = 5;
    d.ps = is_up(&l - g->layers.data());
    tf.push_back(d);
  }
  g->pk_elems = pk;
  // data-gradient weights: plain layers (flipped + transposed)
  auto add_dgrad = [&](int li) {
    Layer& l = g->layers[li];
    l.pkd_off = pkd; pkd += (long long)packed_w_elems(l.Co, l.Ci);
    PackDesc d{}; d.src_off = l.w_off; d.dst_off = l.pkd_off; d.Ci = l.Ci; d.Co = l.Co; d.CoP = round_up(l.Ci, 16); d.mode = 1;
    d.umma = g->bf && umma_ok(l.Co, l.Ci);
    d.ps = is_up(li);
    td.push_back(d);
    maxd = std::max(maxd, l.Ci * l.Co * 9);
  };
  add_dgrad(g->idx_conv1());
  add_dgrad(g->idx_conv2());
  for (int u = 0; u < g->U; ++u) add_dgrad(g->idx_up(u));
  add_dgrad(g->idx_c30());
  add_dgrad(g->idx_c32());
  // dense blocks: Wt_k (k = 0..4) gathers slice k of W_j for j = k+1..5; rows ordered [dz5, dz4, ..., dz_{k+1}]
  g->db_dgrad_off.assign((size_t)g->R * 3 * 5, 0);
  for (int r = 0; r < g->R; ++r)
    for (int dd = 0; dd < 3; ++dd)
      for (int k = 0; k < 5; ++k) {
        const int rows = (5 - k) * F;
        const long long base = pkd;
        g->db_dgrad_off[(size_t)(r * 3 + dd) * 5 + k] = base;
        pkd += (long long)packed_w_elems(rows, F);
        for (int j = k + 1; j <= 5; ++j) {
          const Layer& l = g->layers[g->idx_db(r, dd, j)];
          PackDesc d{};
          d.src_off = l.w_off; d.dst_off = base; d.Ci = F; d.Co = F; d.CoP = rows; d.mode = 3;
          d.slice_off = k * F; d.src_ci_total = j * F; d.dst_row_off = (5 - j) * F; d.dst_CoP = round_up(F, 16);
          d.umma = g->bf && F % 16 == 0;
          td.push_back(d);
          maxd = std::max(maxd, F * F * 9);
        }
      }
  g->pkd_elems = pkd;
  g->n_fwd = (int)tf.size(); g->n_dgrad = (int)td.size(); g->max_fwd = maxf; g->max_dgrad = maxd;
  std::vector<UmmaPackDesc> uf, ud;
  if (g->bf) {
    for (auto& l : g->layers)
      if (umma_ok(l.Ci, l.Co)) { uf.push_back({l.pk_off, l.Ci, round_up(l.Co, 16)}); g->max_ufwd = std::max(g->max_ufwd, 9 * l.Ci * l.Co); }
    for (auto& l : g->layers)
      if (l.pkd_off >= 0 && umma_ok(l.Co, l.Ci)) { ud.push_back({l.pkd_off, l.Co, round_up(l.Ci, 16)}); g->max_udgrad = std::max(g->max_udgrad, 9 * l.Ci * l.Co); }
    if (F % 16 == 0)
      for (int i = 0; i < g->R * 3; ++i)
        for (int k = 0; k < 5; ++k) {
          ud.push_back({g->db_dgrad_off[(size_t)i * 5 + k], (5 - k) * F, round_up(F, 16)});
          g->max_udgrad = std::max(g->max_udgrad, 9 * (5 - k) * F * F);
        }
  }
  g->n_ufwd = (int)uf.size(); g->n_udgrad = (int)ud.size();
  // the streaming kernel also takes narrow outputs (conv3.2: F -> n_predictands as N = 16 with zero pad columns)
  std::vector<UmmaPackDesc> igf = uf;
  if (g->bf)
    for (auto& l : g->layers)
      if (!umma_ok(l.Ci, l.Co) && umma_img_ok(l.Ci, l.Co)) igf.push_back({l.pk_off, l.Ci, round_up(l.Co, 16)});
  g->n_igfwd = (int)igf.size();
  int s = 0;
#define GA(ptr, bytes) if ((s = dev_alloc(g->pool, (void**)&(ptr), (bytes))) != 0) { dg_generator_destroy(g); return s; }
  GA(g->pk, sizeof(float) * pk);
  GA(g->gpk, sizeof(float) * pk);
  GA(g->pkd, sizeof(float) * std::max<long long>(pkd, 1));
  GA(g->pk_u, sizeof(bf16) * (pk + 64));
  GA(g->pkd_u, sizeof(bf16) * (std::max<long long>(pkd, 1) + 64));
  GA(g->pk_ig, sizeof(bf16) * (pk + 64));
  GA(g->pkd_ig, sizeof(bf16) * (std::max<long long>(pkd, 1) + 64));
  GA(g->pk_trunk, sizeof(bf16) * ((size_t)std::max(1, g->R * 3) * 9 * F * F * 15 + 64));
  GA(g->pkd_trunk, sizeof(bf16) * ((size_t)std::max(1, g->R * 3) * 9 * F * F * 15 + 64));
  if ((s = upload_table(g->pool, tf, &g->tab_fwd)) != 0) { dg_generator_destroy(g); return s; }
  if ((s = upload_table(g->pool, td, &g->tab_dgrad)) != 0) { dg_generator_destroy(g); return s; }
  if ((s = upload_utable(g->pool, uf, &g->utab_fwd)) != 0) { dg_generator_destroy(g); return s; }
  if ((s = upload_utable(g->pool, ud, &g->utab_dgrad)) != 0) { dg_generator_destroy(g); return s; }
  if ((s = upload_utable(g->pool, igf, &g->utab_igfwd)) != 0) { dg_generator_destroy(g); return s; }
  // ---- activations
  const size_t B = g->maxB, pc = (size_t)g->Hc * g->Hc, pf = (size_t)g->Hf * g->Hf;
  GA(g->x0, B * pc * g->Cin * g->esz);
  const int ndb = std::max(1, g->R * 3);
  g->db.assign(ndb, nullptr);
  for (int i = 0; i < ndb; ++i) GA(g->db[i], B * pc * 5 * F * g->esz);
  GA(g->db_ptrs_dev, sizeof(void*) * ndb);
  if (cudaMemcpy(g->db_ptrs_dev, g->db.data(), sizeof(void*) * ndb, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("cudaMemcpy(db_ptrs) failed"); dg_generator_destroy(g); return DG_ERR_CUDA;
  }
  GA(g->trunk_out, B * pc * F * g->esz);
  GA(g->t1, B * pc * F * g->esz);
  g->up.assign(g->U, nullptr);
  for (int u = 0; u < g->U; ++u) GA(g->up[u], B * (pc << (2 * (u + 1))) * F * g->esz);
  GA(g->c30, B * pf * F * g->esz);
  GA(g->fake, B * pf * g->Cout * sizeof(float));
  GA(g->dfake, B * pf * g->Cout * sizeof(float));
  GA(g->fine_nhwc, B * pf * g->Cout * sizeof(float));
  GA(g->l1, 64);
  // backward
  GA(g->D, B * pc * 5 * F * g->esz);
  if (g->bf && F % 16 == 0 && 5 * F <= 128 && g->R > 0) {  // the tcgen05 wgrad kernel takes Ci <= 128
    g->Dall.assign((size_t)g->R * 3, nullptr);
    g->Dall[0] = g->D;
    for (int i = 1; i < g->R * 3; ++i) GA(g->Dall[i], B * pc * 5 * F * g->esz);
    GA(g->wg_table_dev, wgrad_umma_args_size() * (size_t)g->R * 15);
    GA(g->wgws_table_dev, 4 * wgrad_ws_batch_bytes(g->R * 15 * 2));  // up to four RRDB ranges, one table each
    GA(g->d_ptrs_dev, sizeof(void*) * g->Dall.size());
    if (cudaMemcpy(g->d_ptrs_dev, g->Dall.data(), sizeof(void*) * g->Dall.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("cudaMemcpy(d_ptrs) failed"); dg_generator_destroy(g); return DG_ERR_CUDA;
    }
  }
  GA(g->gR, B * pc * F * g->esz);
  GA(g->gx0, B * pc * F * g->esz);
  GA(g->gx1, B * pc * F * g->esz);
  GA(g->gT1, B * pc * F * g->esz);
  GA(g->gA, B * pf * F * g->esz);
  GA(g->gB, B * pf * F * g->esz);   // 4F channels at quarter resolution
  g->gU.assign(std::max(g->U, 1), nullptr);
  for (int u = 0; u < g->U; ++u) {
    if (u == g->U - 1) g->gU[u] = g->gB;
    else GA(g->gU[u], B * (pc << (2 * u)) * 4 * F * g->esz);
  }
  if ((s = g->side.create(2)) != 0) { dg_generator_destroy(g); return s; }
#undef GA
  *out = g;
  return 0;
}

extern "C" int dg_generator_destroy(dg_generator* g) {
  if (!g) return 0;
  g->side.destroy();
  for (void* p : g->pool) cudaFree(p);
  delete g;
  return 0;
}

extern "C" int dg_generator_pack(dg_generator* g, const float* params, void* stream) {
  DG_CHECK(g && params, "dg_generator_pack: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  const bool g_ig = g->use_ig && g_tune[16];  // K-major images of the streaming kernel ride on the same launch
  DG_TRY(pack_weights2(params, g->pk, g->pk_u, g->tab_fwd, g->n_fwd, g->max_fwd, g->pkd, g->pkd_u, g->tab_dgrad, g->n_dgrad,
                       g->max_dgrad, st, g_ig ? g->pk_ig : nullptr, g_ig ? g->pkd_ig : nullptr));
  if (trunk_fused_supported(g->F, g->Hc, g->R, g->bf))
  {
    DG_TRY(pack_trunk_slices(g->pk + g->layers[g->idx_db(0, 0, 1)].pk_off, g->pk_trunk, g->R * 3, 0, st));
    DG_TRY(pack_trunk_slices(g->pkd + g->db_dgrad_off[0], g->pkd_trunk, g->R * 3, 1, st));
  }
  g->packed = true;
  return 0;
}

// The RRDB trunk (generator.py:36-53; `self.res_blocks(out1)` at :85) over samples [s0, s0 + B): reads slice 0 of g->db[0],
// writes g->trunk_out.  Fused persistent tcgen05 kernel where the shape allows, per-layer convs otherwise.
static int gen_trunk_forward(dg_generator* g, int s0, int B, int save_count, cudaStream_t st) {
  const bool save = save_count > 0;
  const int F = g->F, Hc = g->Hc;
  const size_t pc = (size_t)Hc * Hc;
  auto V = [&](void* p, int pitch, int coff, size_t pixels) { return tv_batch(g->act(p, pitch, coff), pixels, s0); };
  auto conv = [&](int li, TV x, int H, TV y) {
    const Layer& l = g->layers[li];
    ConvOp op;
    op.x = x; op.Hin = H; op.Win = H; op.Ci = l.Ci;
    op.y = y; op.Hout = H; op.Wout = H; op.Co = l.Co;
    op.B = B; op.w = g->pk + l.pk_off; op.bias = g->pk + l.pkb_off;
    if (g->bf) { op.w_umma = g->pk_u + l.pk_off; op.w_ig = g->use_ig ? g->pk_ig + l.pk_off : nullptr; }
    return op;
  };
  const bool fused_trunk = trunk_fused_supported(F, Hc, g->R, g->bf);
  if (fused_trunk) {
    // persistent tcgen05 kernel: the whole RRDB trunk with the concat buffer resident in shared memory
    const Layer& l0 = g->layers[g->idx_db(0, 0, 1)];
    DG_TRY(trunk_fwd_fused(V(g->db[0], 5 * F, 0, pc).p, 5 * F, 0, V(g->trunk_out, F, 0, pc).p, F,
                           (save && s0 == 0) ? (void* const*)g->db_ptrs_dev : nullptr, g->pk_trunk, g->pk + l0.pkb_off, g->R, B,
                           s0 == 0 ? save_count : 0, st));
  }
  for (int r = 0; r < (fused_trunk ? 0 : g->R); ++r)
    for (int d = 0; d < 3; ++d) {
      void* buf = g->db[r * 3 + d];
      for (int k = 1; k <= 4; ++k) {
        ConvOp op = conv(g->idx_db(r, d, k), V(buf, 5 * F, 0, pc), Hc, V(buf, 5 * F, k * F, pc));
        op.act = ACT_LRELU; op.slope = G_SLOPE;
        DG_TRY(run_conv(op, st));
      }
      const bool last_db = (r == g->R - 1 && d == 2);
      void* nxt = last_db ? g->trunk_out : g->db[r * 3 + d + 1];
      const int npitch = last_db ? F : 5 * F;
      ConvOp op = conv(g->idx_db(r, d, 5), V(buf, 5 * F, 0, pc), Hc, V(nxt, npitch, 0, pc));
      if (d < 2) {  // 0.2*o5 + x
        op.s_acc = RES_SCALE; op.r1 = V(buf, 5 * F, 0, pc); op.s1 = 1.f;
      } else {      // 0.2*(0.2*o5 + x_db) + x_rrdb
        op.s_acc = RES_SCALE * RES_SCALE; op.r1 = V(buf, 5 * F, 0, pc); op.s1 = RES_SCALE;
        op.r2 = V(g->db[r * 3], 5 * F, 0, pc); op.s2 = 1.f;
      }
      DG_TRY(run_conv(op, st));
    }
  return 0;
}

// Generator.forward, networks/generator.py:83-90 (dense block :36-41, RRDB :52-53).
// Input already in g->x0 (NHWC); output in g->fake (NHWC fp32).
// `save_count` leading samples keep their activations for a later backward (0: inference / critic iterations).
// Generator.forward (generator.py:83-90) over samples [s0, s0 + B) of g->x0; the first save_count samples keep their
// dense-block activations (only meaningful for s0 == 0).
static int gen_forward_range(dg_generator* g, int s0, int B, int save_count, cudaStream_t st) {
  const int F = g->F, Hc = g->Hc;
  const size_t pc = (size_t)Hc * Hc;
  auto V = [&](void* p, int pitch, int coff, size_t pixels) { return tv_batch(g->act(p, pitch, coff), pixels, s0); };
  auto conv = [&](int li, TV x, int H, TV y) {
    const Layer& l = g->layers[li];
    ConvOp op;
    op.x = x; op.Hin = H; op.Win = H; op.Ci = l.Ci;
    op.y = y; op.Hout = H; op.Wout = H; op.Co = l.Co;
    op.B = B; op.w = g->pk + l.pk_off; op.bias = g->pk + l.pkb_off;
    if (g->bf) { op.w_umma = g->pk_u + l.pk_off; op.w_ig = g->use_ig ? g->pk_ig + l.pk_off : nullptr; }
    return op;
  };
  void* first = g->R > 0 ? g->db[0] : g->trunk_out;
  const int first_pitch = g->R > 0 ? 5 * F : F;
  {
    ConvOp op = conv(g->idx_conv1(), V(g->x0, g->Cin, 0, pc), Hc, V(first, first_pitch, 0, pc));
    DG_TRY(run_conv(op, st));
  }
  DG_TRY(gen_trunk_forward(g, s0, B, save_count, st));
  {  // out1 + conv2(trunk)
    ConvOp op = conv(g->idx_conv2(), V(g->trunk_out, F, 0, pc), Hc, V(g->t1, F, 0, pc));
    if (g->R > 0) { op.r1 = V(g->db[0], 5 * F, 0, pc); op.s1 = 1.f; }
    else { op.s_acc = 1.f; op.r1 = V(g->trunk_out, F, 0, pc); op.s1 = 1.f; }
    DG_TRY(run_conv(op, st));
  }
  void* cur = g->t1;
  int H = Hc;
  for (int u = 0; u < g->U; ++u) {  // conv -> LeakyReLU -> PixelShuffle(2)
    ConvOp op = conv(g->idx_up(u), V(cur, F, 0, (size_t)H * H), H, V(g->up[u], F, 0, (size_t)4 * H * H));
    op.act = ACT_LRELU; op.slope = G_SLOPE; op.shuffle = SHUF_PIXEL;
    DG_TRY(run_conv(op, st));
    cur = g->up[u];
    H *= 2;
  }
  {
    ConvOp op = conv(g->idx_c30(), V(cur, F, 0, (size_t)H * H), H, V(g->c30, F, 0, (size_t)H * H));
    op.act = ACT_LRELU; op.slope = G_SLOPE;
    DG_TRY(run_conv(op, st));
  }
  {
    ConvOp op = conv(g->idx_c32(), V(g->c30, F, 0, (size_t)H * H), H,
                     tv_batch(tv(g->fake, 0, g->Cout), (size_t)H * H, s0));
    op.narrow_ok = g->bf && umma_img_ok(F, g->Cout);
    DG_TRY(run_conv(op, st));
  }
  return 0;
}

// g_tune[10] = k > 0: when the batch is larger than one wave of the persistent trunk kernel (2 CTAs x 148 SMs), the last
// 32k samples run as a second chain on the side stream, so the tail convolutions of the first chain overlap the trunk
// kernel's partial second wave instead of waiting for it.
static int gen_forward_internal(dg_generator* g, int B, int save_count, cudaStream_t st) {
  const int nb = 32 * g_tune[10];
  const bool split = nb > 0 && g->side.s != nullptr && B > 2 * 148 && B - nb >= save_count && B - nb >= nb &&
                     trunk_fused_supported(g->F, g->Hc, g->R, g->bf);
  if (split) {
    DG_TRY(g->side.fork(st));
    DG_TRY(gen_forward_range(g, 0, B - nb, save_count, st));
    DG_TRY(gen_forward_range(g, B - nb, nb, 0, g->side.s));
    DG_TRY(g->side.join(st));
  } else {
    DG_TRY(gen_forward_range(g, 0, B, save_count, st));
  }
  g->saved_batch = save_count;
  return 0;
}

extern "C" int dg_generator_fwd(dg_generator* g, const float* coarse, int batch, float* fake, int save, void* stream) {
  DG_CHECK(g && coarse, "dg_generator_fwd: null argument");
  DG_CHECK(batch >= 1 && batch <= g->maxB, "dg_generator_fwd: batch %d outside [1,%d]", batch, g->maxB);
  if (!g->packed) { set_error("dg_generator_fwd: dg_generator_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), batch, g->Cin, g->Hc, g->Hc, st));
  g->lookahead = 0;
  DG_TRY(gen_forward_internal(g, batch, save ? batch : 0, st));
  if (fake) DG_TRY(nhwc_to_nchw(tv(g->fake, 0, g->Cout), fake, batch, g->Cout, g->Hf, g->Hf, st));
  return 0;
}

// Backward of the RRDB trunk: g->gR holds dL/d(trunk output) on entry and dL/d(trunk input) (through the trunk only) on exit;
// the dense convs' weight / bias gradients are accumulated into g->gpk.  `side_on` is cleared when the per-layer path had to
// join the side stream (it reuses its dz buffers block after block).
static int gen_trunk_backward(dg_generator* g, int B, bool& side_on, cudaStream_t st) {
  const int F = g->F, Hc = g->Hc;
  auto wgrad = [&](int li, TV x, int H, TV dy) {
    const Layer& l = g->layers[li];
    WgradOp w;
    w.x = x; w.Hin = H; w.Win = H; w.Ci = l.Ci; w.dy = dy; w.Hout = H; w.Wout = H; w.Co = l.Co; w.B = B; w.stride = 1;
    w.dw = g->gpk + l.pk_off; w.dbias = g->gpk + l.pkb_off;
    if (!side_on) return run_wgrad(w, st);
    DG_TRY(g->side.fork(st));
    return run_wgrad(w, g->side.s);
  };
  const size_t pix = (size_t)B * Hc * Hc;
  // trunk, RRDB by RRDB
  void* gxs[2] = {g->gx0, g->gx1};
  const bool batched_wgrad = !g->Dall.empty();
  std::vector<WgradOp> wops;
  if (batched_wgrad) wops.reserve((size_t)g->R * 15);
  void* const D0 = g->D;
  const bool fused_bwd = batched_wgrad && trunk_fused_supported(F, Hc, g->R, g->bf);
  if (!fused_bwd && side_on) {  // the per-layer trunk path reuses its dz / gradient buffers block after block
    DG_TRY(g->side.join(st));
    side_on = false;
  }
  if (fused_bwd) {
    // Persistent tcgen05 kernel: the data-gradient chain of the trunk, dz slices saved per block.  The chain is launched in
    // g_tune[18] RRDB ranges (last range first): the dense convs' weight gradients of a range only need that range's dz, so
    // they are enqueued on the side stream as soon as the range's kernel is (one image per CTA: 64 of 148 SMs at cfg-2) and run
    // beside the NEXT range's data gradients instead of after the whole chain (measured: profiles/README.md).
    // which: 0 = every channel block, 1 = only the 64-channel blocks (nine-tap units), 2 = only the 16- / 32-channel blocks
    auto dense_wgrads = [&](int r_lo, int r_hi, int part, cudaStream_t s2, int which = 0, bool with_colsum = true) -> int {
      std::vector<WgradOp> blk;
      blk.reserve((size_t)(r_hi - r_lo) * 15 * 2);
      for (int r = r_hi - 1; r >= r_lo; --r)
        for (int d = 2; d >= 0; --d)
          for (int k = 1; k <= 5; ++k) {
            const Layer& l = g->layers[g->idx_db(r, d, k)];
            WgradOp w;
            memset(&w, 0, sizeof(w));
            w.x = g->act(g->db[r * 3 + d], 5 * F, 0); w.Hin = Hc; w.Win = Hc; w.Ci = l.Ci;
            w.dy = g->act(g->Dall[(size_t)r * 3 + d], 5 * F, (5 - k) * F); w.Hout = Hc; w.Wout = Hc; w.Co = l.Co;
            w.B = B; w.stride = 1; w.dw = g->gpk + l.pk_off; w.dbias = g->gpk + l.pkb_off;
            if (!(g_tune[4] && F == 16)) { wops.push_back(w); continue; }
            for (int c0 = 0; c0 < w.Ci;) {  // power-of-two channel blocks for the TMA-fed kernel
              int cb = 64;
              while (cb > w.Ci - c0) cb >>= 1;
              WgradOp o = w;
              o.x.coff = w.x.coff + c0; o.Ci = cb; o.dbias = nullptr;
              o.dw_ci_total = w.Ci; o.dw_ci_off = c0;
              if (which == 0 || (which == 1) == (cb == 64)) blk.push_back(o);
              c0 += cb;
            }
          }
      if (!blk.empty()) {
        if ((int)g->wgws_parts.size() <= part) g->wgws_parts.resize(part + 1);
        char* table = (char*)g->wgws_table_dev + (size_t)part * wgrad_ws_batch_bytes(g->R * 15 * 2);
        DG_TRY(wgrad_ws_batched(blk.data(), (int)blk.size(), table, g->wgws_parts[part], 2, s2));
      }
      if (!with_colsum) return 0;
      // bias gradients of the range = column sums of its dz buffers
      const Layer& l0 = g->layers[g->idx_db(r_lo, 0, 1)];
      return colsum_dense_blocks((void* const*)g->d_ptrs_dev + 3 * r_lo, (r_hi - r_lo) * 3, (size_t)B * Hc * Hc, g->gpk + l0.pkb_off, s2);
    };
    int nsplit = std::max(1, std::min(std::min(g_tune[18], 4), g->R));
    if (!(side_on && g_tune[4] && F == 16)) nsplit = 1;
    for (int p = nsplit - 1; p >= 0; --p) {
      const int r_lo = (int)((long long)g->R * p / nsplit), r_hi = (int)((long long)g->R * (p + 1) / nsplit);
      DG_TRY(trunk_bwd_fused(g->gR, g->gR, (void* const*)g->db_ptrs_dev, (void* const*)g->d_ptrs_dev, g->pkd_trunk, r_hi - r_lo, B, st, r_lo));
      if (p > 0) {  // overlaps the next range's kernel
        DG_TRY(g->side.fork(st));
        DG_TRY(dense_wgrads(r_lo, r_hi, p, g->side.s));
      } else if (g_tune[23] && nsplit < 4) {
        // last range: nothing is left to overlap it with, so its two batched launches (64-channel blocks / narrower blocks) run
        // beside each other on the side stream and the caller's stream instead of one after the other
        DG_TRY(g->side.fork(st));
        DG_TRY(dense_wgrads(r_lo, r_hi, nsplit, g->side.s, 1, false));
        DG_TRY(dense_wgrads(r_lo, r_hi, p, st, 2, true));
      } else {
        DG_TRY(dense_wgrads(r_lo, r_hi, p, st));
      }
    }
  }
  for (int r = fused_bwd ? -1 : g->R - 1; r >= 0; --r) {
    // g->gR holds dL/d(RRDB_r output)
    void* gin = g->gR;
    float s_in = RES_SCALE;
    for (int d = 2; d >= 0; --d) {
      void* buf = g->db[r * 3 + d];
      if (batched_wgrad) g->D = g->Dall[(size_t)r * 3 + d];
      // dz5 = 0.2 * s_in * gin
      DG_TRY(scale_add(g->act(g->D, 5 * F, 0), g->act(gin, F), RES_SCALE * s_in, TV(), 0.f, pix, F, st));
      for (int k = 4; k >= 1; --k) {
        ConvOp op;
        op.x = g->act(g->D, 5 * F, 0); op.Hin = Hc; op.Win = Hc; op.Ci = (5 - k) * F;
        op.y = g->act(g->D, 5 * F, (5 - k) * F); op.Hout = Hc; op.Wout = Hc; op.Co = F;
        op.B = B; op.w = g->pkd + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + k];
        if (g->bf) { op.w_umma = g->pkd_u + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + k]; op.w_ig = g->use_ig ? g->pkd_ig + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + k] : nullptr; }
        op.act = ACT_MASK; op.slope = G_SLOPE; op.mask = g->act(buf, 5 * F, k * F);
        DG_TRY(run_conv(op, st));
      }
      // weight gradients of b1..b5: x = buf[0:kF], dy = dz_k
      for (int k = 1; k <= 5; ++k) {
        if (!batched_wgrad) {
          DG_TRY(wgrad(g->idx_db(r, d, k), g->act(buf, 5 * F, 0), Hc, g->act(g->D, 5 * F, (5 - k) * F)));
          continue;
        }
        const Layer& l = g->layers[g->idx_db(r, d, k)];
        WgradOp w;
        memset(&w, 0, sizeof(w));
        w.x = g->act(buf, 5 * F, 0); w.Hin = Hc; w.Win = Hc; w.Ci = l.Ci;
        w.dy = g->act(g->D, 5 * F, (5 - k) * F); w.Hout = Hc; w.Wout = Hc; w.Co = l.Co; w.B = B; w.stride = 1;
        w.dw = g->gpk + l.pk_off; w.dbias = g->gpk + l.pkb_off;
        wops.push_back(w);
      }
      // gx = conv(D, Wt_0) + s_in*gin (+ gR when this is the RRDB's first block)
      void* gout = (d == 0) ? g->gR : gxs[d & 1];
      ConvOp op;
      op.x = g->act(g->D, 5 * F, 0); op.Hin = Hc; op.Win = Hc; op.Ci = 5 * F;
      op.y = g->act(gout, F); op.Hout = Hc; op.Wout = Hc; op.Co = F;
      op.B = B; op.w = g->pkd + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + 0];
      if (g->bf) { op.w_umma = g->pkd_u + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + 0]; op.w_ig = g->use_ig ? g->pkd_ig + g->db_dgrad_off[(size_t)(r * 3 + d) * 5 + 0] : nullptr; }
      op.r1 = g->act(gin, F); op.s1 = s_in;
      if (d == 0) { op.r2 = g->act(g->gR, F); op.s2 = 1.f; }
      DG_TRY(run_conv(op, st));
      gin = gout;
      s_in = 1.f;
    }
  }
  g->D = D0;
  if (batched_wgrad && !wops.empty()) {
    // (TMA-fed path: already launched range by range above.)  All remaining dense-conv weight + bias gradients in one
    // tcgen05 launch of the cp.async kernel.
    DG_TRY(wgrad_umma_batched(wops.data(), (int)wops.size(), g->wg_table_dev, g->wg_shadow, 2, st));
  }
  return 0;
}

// Backward of Generator.forward given g->dfake (NHWC fp32); fills g->gpk and unpacks into grads_flat.
static int gen_backward_internal(dg_generator* g, float* grads_flat, float* d_coarse, cudaStream_t st) {
  const int B = g->saved_batch, F = g->F, Hc = g->Hc, Hf = g->Hf;
  if (B <= 0) { set_error("generator backward without a saved forward"); return DG_ERR_STATE; }
  DG_CUDA(cudaMemsetAsync(g->gpk, 0, sizeof(float) * g->pk_elems, st));
  // Weight gradients only feed the final unpack: with the side stream on they are enqueued there (after everything
  // enqueued on `st` so far, i.e. after the data-gradient that produced their dy) and joined before the unpack, so the
  // tail layers' weight gradients overlap the data-gradient chain and the 64-CTA trunk kernel.
  bool side_on = g_tune[9] && g->side.s != nullptr;
  auto wgrad = [&](int li, TV x, int H, TV dy) {
    const Layer& l = g->layers[li];
    WgradOp w;
    w.x = x; w.Hin = H; w.Win = H; w.Ci = l.Ci; w.dy = dy; w.Hout = H; w.Wout = H; w.Co = l.Co; w.B = B; w.stride = 1;
    w.dw = g->gpk + l.pk_off; w.dbias = g->gpk + l.pkb_off;
    if (!side_on) return run_wgrad(w, st);
    DG_TRY(g->side.fork(st));
    return run_wgrad(w, g->side.s);
  };
  auto dconv = [&](int li, TV dy, int H, TV dx) {  // data-gradient op of plain layer li: Ci_op = Co, Co_op = Ci
    const Layer& l = g->layers[li];
    ConvOp op;
    op.x = dy; op.Hin = H; op.Win = H; op.Ci = l.Co;
    op.y = dx; op.Hout = H; op.Wout = H; op.Co = l.Ci;
    op.B = B; op.w = g->pkd + l.pkd_off;
    if (g->bf) { op.w_umma = g->pkd_u + l.pkd_off; op.w_ig = g->use_ig ? g->pkd_ig + l.pkd_off : nullptr; }
    return op;
  };
  void* last_up = g->U > 0 ? g->up[g->U - 1] : g->t1;
  // conv3.2
  DG_TRY(wgrad(g->idx_c32(), g->act(g->c30, F), Hf, tv(g->dfake, 0, g->Cout)));
  {
    ConvOp op = dconv(g->idx_c32(), tv(g->dfake, 0, g->Cout), Hf, g->act(g->gA, F));
    op.act = ACT_MASK; op.slope = G_SLOPE; op.mask = g->act(g->c30, F);
    DG_TRY(run_conv(op, st));
  }
  // conv3.0
  DG_TRY(wgrad(g->idx_c30(), g->act(last_up, F), Hf, g->act(g->gA, F)));
  void* gcur = nullptr;  // gradient w.r.t. the input of the layer just processed
  {
    ConvOp op = dconv(g->idx_c30(), g->act(g->gA, F), Hf, g->U > 0 ? g->act(g->gU[g->U - 1], 4 * F) : g->act(g->gT1, F));
    if (g->U > 0) {  // store un-shuffled and masked by the pre-shuffle LeakyReLU (sign taken from the shuffled copy)
      op.shuffle = SHUF_UNPIXEL; op.act = ACT_MASK; op.slope = G_SLOPE; op.mask = g->act(last_up, F);
    }
    DG_TRY(run_conv(op, st));
    gcur = g->U > 0 ? g->gU[g->U - 1] : g->gT1;
  }
  // upsample stages, last to first; gcur = gU[u] holds dz of stage u as (B, H, H, 4F)
  for (int u = g->U - 1; u >= 0; --u) {
    const int H = Hc << u;
    void* xin = u > 0 ? g->up[u - 1] : g->t1;
    DG_TRY(wgrad(g->idx_up(u), g->act(xin, F), H, g->act(gcur, 4 * F)));
    void* dst = u > 0 ? g->gU[u - 1] : g->gT1;
    ConvOp op = dconv(g->idx_up(u), g->act(gcur, 4 * F), H, u > 0 ? g->act(dst, 4 * F) : g->act(dst, F));
    if (u > 0) { op.shuffle = SHUF_UNPIXEL; op.act = ACT_MASK; op.slope = G_SLOPE; op.mask = g->act(xin, F); }
    DG_TRY(run_conv(op, st));
    gcur = dst;
  }
  // gT1 = dL/d(out1 + conv2(trunk))
  DG_TRY(wgrad(g->idx_conv2(), g->act(g->trunk_out, F), Hc, g->act(g->gT1, F)));
  {
    ConvOp op = dconv(g->idx_conv2(), g->act(g->gT1, F), Hc, g->act(g->gR, F));
    DG_TRY(run_conv(op, st));
  }
  const size_t pix = (size_t)B * Hc * Hc;
  DG_TRY(gen_trunk_backward(g, B, side_on, st));
  // dL/d(out1) = gR (through the trunk / conv2) + gT1 (long skip)
  DG_TRY(scale_add(g->act(g->gx0, F), g->act(g->gR, F), 1.f, g->act(g->gT1, F), 1.f, pix, F, st));
  DG_TRY(wgrad(g->idx_conv1(), g->act(g->x0, g->Cin), Hc, g->act(g->gx0, F)));
  if (d_coarse) {
    ConvOp op = dconv(g->idx_conv1(), g->act(g->gx0, F), Hc, g->act(g->gx1, g->Cin));
    DG_TRY(run_conv(op, st));
    DG_TRY(nhwc_to_nchw(g->act(g->gx1, g->Cin), d_coarse, B, g->Cin, Hc, Hc, st));
  }
  DG_TRY(g->side.join(st));
  DG_TRY(unpack_wgrads(g->gpk, grads_flat, g->tab_fwd, g->n_fwd, g->max_fwd, st));
  return 0;
}

extern "C" int dg_generator_bwd(dg_generator* g, const float* d_fake, float* grads_flat, float* d_coarse, void* stream) {
  DG_CHECK(g && d_fake && grads_flat, "dg_generator_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  if (g->saved_batch <= 0) { set_error("dg_generator_bwd: no saved forward"); return DG_ERR_STATE; }
  DG_TRY(nchw_to_nhwc(d_fake, tv(g->dfake, 0, g->Cout), g->saved_batch, g->Cout, g->Hf, g->Hf, st));
  return gen_backward_internal(g, grads_flat, d_coarse, st);
}

// ---- parity instrumentation / unit-test entry points (include/downgan_b200.h) ---------------------------------------
extern "C" int dg_generator_activation(dg_generator* g, int which, int batch, float* out, void* stream) {
  DG_CHECK(g && out, "dg_generator_activation: null argument");
  DG_CHECK(batch >= 1 && batch <= g->saved_batch, "dg_generator_activation: batch %d outside the saved forward (%d samples)", batch,
           g->saved_batch);
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));
  const int F = g->F, Hc = g->Hc;
  if (which >= 0 && which < g->R * 3) return nhwc_to_nchw(g->act(g->db[which], 5 * F), out, batch, 5 * F, Hc, Hc, st);
  if (which == 1000) return nhwc_to_nchw(g->act(g->trunk_out, F), out, batch, F, Hc, Hc, st);
  if (which == 1001) return nhwc_to_nchw(g->act(g->t1, F), out, batch, F, Hc, Hc, st);
  if (which >= 1100 && which < 1100 + g->U) {
    const int H = Hc << (which - 1100 + 1);
    return nhwc_to_nchw(g->act(g->up[which - 1100], F), out, batch, F, H, H, st);
  }
  if (which == 1200) return nhwc_to_nchw(g->act(g->c30, F), out, batch, F, g->Hf, g->Hf, st);
  set_error("dg_generator_activation: unknown activation %d", which);
  return DG_ERR_INVALID;
}

extern "C" int dg_generator_trunk_fwd(dg_generator* g, const float* x, int batch, float* y, void* stream) {
  DG_CHECK(g && x && y, "dg_generator_trunk_fwd: null argument");
  DG_CHECK(batch >= 1 && batch <= g->maxB, "dg_generator_trunk_fwd: batch %d outside [1,%d]", batch, g->maxB);
  DG_CHECK(g->R > 0, "dg_generator_trunk_fwd: the generator has no residual blocks");
  if (!g->packed) { set_error("dg_generator_trunk_fwd: dg_generator_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));
  DG_TRY(nchw_to_nhwc(x, g->act(g->db[0], 5 * g->F, 0), batch, g->F, g->Hc, g->Hc, st));
  g->lookahead = 0;
  DG_TRY(gen_trunk_forward(g, 0, batch, batch, st));
  g->saved_batch = batch;
  return nhwc_to_nchw(g->act(g->trunk_out, g->F), y, batch, g->F, g->Hc, g->Hc, st);
}

extern "C" int dg_generator_trunk_bwd(dg_generator* g, const float* d_y, float* d_x, float* grads_flat, void* stream) {
  DG_CHECK(g && d_y && grads_flat, "dg_generator_trunk_bwd: null argument");
  DG_CHECK(g->R > 0, "dg_generator_trunk_bwd: the generator has no residual blocks");
  if (g->saved_batch <= 0) { set_error("dg_generator_trunk_bwd: no saved forward"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));
  const int B = g->saved_batch;
  DG_TRY(nchw_to_nhwc(d_y, g->act(g->gR, g->F), B, g->F, g->Hc, g->Hc, st));
  DG_CUDA(cudaMemsetAsync(g->gpk, 0, sizeof(float) * g->pk_elems, st));
  bool side_on = g_tune[9] && g->side.s != nullptr;
  DG_TRY(gen_trunk_backward(g, B, side_on, st));
  DG_TRY(g->side.join(st));
  DG_TRY(unpack_wgrads(g->gpk, grads_flat, g->tab_fwd, g->n_fwd, g->max_fwd, st));
  if (d_x) DG_TRY(nhwc_to_nchw(g->act(g->gR, g->F), d_x, B, g->F, g->Hc, g->Hc, st));
  return 0;
}

// ===========================================================================
// Critic
// ===========================================================================
struct dg_critic {
  dg_critic_config cfg{};
  int bf = 0, W = 0, Hf = 0, nc = 0, maxB = 0, NBmax = 0, fc_in = 0, Hlast = 0, Clast = 0;
  size_t esz = 4;
  Layer L[8];
  int Hin[8], Hout[8];
  long long n_params = 0;
  long long fc1w_off = 0, fc1b_off = 0, fc2w_off = 0, fc2b_off = 0;          // flat
  long long pk_fc1w = 0, pk_fc1b = 0, pk_fc2w = 0, pk_fc2b = 0, pk_b0 = 0;   // packed
  std::vector<void*> pool;
  float *pk = nullptr, *pkd = nullptr, *gpk = nullptr;
  long long pk_elems = 0, pkd_elems = 0;
  PackDesc *tab_fwd = nullptr, *tab_dgrad = nullptr;
  int n_fwd = 0, n_dgrad = 0, max_fwd = 0, max_dgrad = 0;
  bf16 *pk_u = nullptr, *pkd_u = nullptr;
  bf16 *pk_ig = nullptr, *pkd_ig = nullptr;  // K-major images [tap][CoP][Ci] for the streaming implicit-GEMM kernel
  UmmaPackDesc *utab_fwd = nullptr, *utab_dgrad = nullptr, *utab_igdgrad = nullptr;
  int n_ufwd = 0, n_udgrad = 0, max_ufwd = 1, max_udgrad = 1, n_igdgrad = 0;
  bool packed = false;
  const float* pending_params = nullptr;  // dg_critic_pack_lazy: flat parameters to pack before the next use of the weights
  // activations for up to NBmax samples
  float* a0 = nullptr;       // NHWC fp32 input batch
  void* a[9] = {nullptr};    // a[1..8]
  float *a9 = nullptr, *scores = nullptr, *seed = nullptr, *dz9 = nullptr, *vfc = nullptr;
  void* dz[9] = {nullptr};   // dz[1..8]
  unsigned short* bits[9] = {nullptr};  // bits[l]: LeakyReLU sign bits of a[l] (ConvOp::bits_out), bf16 mode, l = 1..8
  bool bits_ok[9] = {false};            // the last forward over these samples really wrote bits[l] (depends on the kernel chosen)
  // sign bits of a[l] for samples starting at s0, or null (not written / switched off by dg_set_tuning(21, 0))
  const unsigned short* bits_at(int l, int s0) const {
    if (!bits[l] || !bits_ok[l] || !g_tune[21]) return nullptr;
    return bits[l] + (size_t)s0 * pix(l - 1) * (size_t)(L[l - 1].Co >> 4);
  }
  float *g = nullptr, *u = nullptr;        // (maxB,Hf,Hf,nc) fp32
  void *v0 = nullptr, *v1 = nullptr;       // JVP ping-pong
  double* metric_scratch = nullptr;        // block partials of the MAE / MSE reduction (dg_metrics)
  float *fs_real = nullptr, *fs_fake = nullptr;  // frequency separation: high-pass real (NCHW) / fake (NHWC), allocated on first use
  void* keep[9] = {nullptr};               // parity instrumentation (dg_set_tuning(15, 1)): the interpolates' activations a[1..8]
  int keep_batch = 0;
  float *sumsq = nullptr, *coef = nullptr, *norms = nullptr, *scal = nullptr;
  SideStream side, side2;
  int defer_conv = 0;           // dg_critic_defer_conv_grads: the fused iteration returns before its conv weight gradients are final
  bool pending_finish = false;  // ... and dg_critic_step_finish has not been called yet
  static constexpr int N_CONV_ENTRIES = 9;  // tab_fwd order: 8 conv weights, features.0.bias, then the 4 classifier tensors
  int saved_batch = 0;
  TV act(void* p, int pitch, int coff = 0) const { return tv(p, bf, pitch, coff); }
  size_t pix(int l) const { return (size_t)Hout[l] * Hout[l]; }  // l = 0..7 -> a[l+1]
};

static void critic_enumerate(const dg_critic_config& c, Layer* L, long long& fc1w, long long& fc1b, long long& fc2w,
                             long long& fc2b, long long& n_params, int& fc_in) {
  const int w = c.coarse_dim;
  const int ci[8] = {c.nc, w, w, 2 * w, 2 * w, 4 * w, 4 * w, 8 * w};
  const int co[8] = {w, w, 2 * w, 2 * w, 4 * w, 4 * w, 8 * w, 8 * w};
  long long off = 0;
  for (int i = 0; i < 8; ++i) {
    L[i].Ci = ci[i]; L[i].Co = co[i]; L[i].stride = (i & 1) ? 2 : 1;
    L[i].w_off = off; off += (long long)co[i] * ci[i] * 9;
    if (i == 0) { L[i].b_off = off; off += co[i]; } else L[i].b_off = -1;
  }
  const int hl = c.fine_dim / 16;
  fc_in = 8 * w * hl * hl;
  fc1w = off; off += (long long)FC_HIDDEN * fc_in;
  fc1b = off; off += FC_HIDDEN;
  fc2w = off; off += FC_HIDDEN;
  fc2b = off; off += 1;
  n_params = off;
}
extern "C" int64_t dg_critic_param_count(const dg_critic_config* cfg) {
  if (!cfg) return -1;
  Layer L[8]; long long a, b, c, d, n; int k;
  critic_enumerate(*cfg, L, a, b, c, d, n, k);
  return n;
}
extern "C" int64_t dg_critic_param_offset(const dg_critic_config* cfg, int index) {
  if (!cfg || index < 0) return -1;
  Layer L[8]; long long a, b, c, d, n; int k;
  critic_enumerate(*cfg, L, a, b, c, d, n, k);
  // order: features.0.weight, features.0.bias, features.2.weight ... features.14.weight, classifier.0.{w,b}, classifier.2.{w,b}
  if (index == 0) return L[0].w_off;
  if (index == 1) return L[0].b_off;
  if (index <= 8) return L[index - 1].w_off;
  if (index == 9) return a;
  if (index == 10) return b;
  if (index == 11) return c;
  if (index == 12) return d;
  return -1;
}

extern "C" int dg_critic_create(const dg_critic_config* cfg, dg_critic** out) {
  DG_CHECK(cfg && out, "dg_critic_create: null argument");
  DG_CHECK(cfg->coarse_dim >= 1 && cfg->nc >= 1 && cfg->max_batch >= 1 && cfg->fine_dim >= 16 && cfg->fine_dim % 16 == 0,
           "dg_critic_create: bad config (fine_dim must be a multiple of 16)");
  DG_CHECK(cfg->precision == DG_FP32 || cfg->precision == DG_BF16, "dg_critic_create: precision %d", cfg->precision);
  dg_critic* c = new dg_critic();
  c->cfg = *cfg;
  c->bf = cfg->precision == DG_BF16;
  c->esz = c->bf ? 2 : 4;
  c->W = cfg->coarse_dim; c->Hf = cfg->fine_dim; c->nc = cfg->nc; c->maxB = cfg->max_batch; c->NBmax = 3 * cfg->max_batch;
  critic_enumerate(*cfg, c->L, c->fc1w_off, c->fc1b_off, c->fc2w_off, c->fc2b_off, c->n_params, c->fc_in);
  int H = c->Hf;
  for (int i = 0; i < 8; ++i) { c->Hin[i] = H; H = (c->L[i].stride == 2) ? H / 2 : H; c->Hout[i] = H; }
  c->Hlast = H; c->Clast = 8 * c->W;
  std::vector<PackDesc> tf, td;
  long long pk = 0, pkd = 0;
  int maxf = 1, maxd = 1;
  for (int i = 0; i < 8; ++i) {
    Layer& l = c->L[i];
    l.pk_off = pk; pk += (long long)packed_w_elems(l.Ci, l.Co);
    PackDesc d{}; d.src_off = l.w_off; d.dst_off = l.pk_off; d.Ci = l.Ci; d.Co = l.Co; d.CoP = round_up(l.Co, 16); d.mode = 0;
    d.umma = c->bf && umma_img_ok(l.Ci, l.Co);
    tf.push_back(d);
    maxf = std::max(maxf, l.Ci * l.Co * 9);
    l.pkd_off = pkd; pkd += (long long)packed_w_elems(l.Co, l.Ci);
    PackDesc e{}; e.src_off = l.w_off; e.dst_off = l.pkd_off; e.Ci = l.Ci; e.Co = l.Co; e.CoP = round_up(l.Ci, 16);
    e.mode = (l.stride == 2) ? 2 : 1;
    e.umma = c->bf && umma_img_ok(l.Co, l.Ci);
    td.push_back(e);
    maxd = std::max(maxd, l.Ci * l.Co * 9);
  }
  auto add_copy = [&](long long src, long long& dst, int n) {
    dst = pk; pk += round_up(n, 16);
    PackDesc d{}; d.src_off = src; d.dst_off = dst; d.Co = n; d.Ci = 1; d.mode = 5;
    tf.push_back(d);
  };
  add_copy(c->L[0].b_off, c->pk_b0, c->L[0].Co);
  c->L[0].pkb_off = c->pk_b0;
  {  // classifier.0.weight with NCHW -> NHWC column permutation (critic.py:103 flattens NCHW)
    c->pk_fc1w = pk; pk += (long long)FC_HIDDEN * c->fc_in;
    PackDesc d{}; d.src_off = c->fc1w_off; d.dst_off = c->pk_fc1w; d.Co = FC_HIDDEN; d.Ci = c->fc_in; d.mode = 4;
    d.slice_off = c->Clast;
    tf.push_back(d);
    maxf = std::max(maxf, FC_HIDDEN * c->fc_in);
  }
  add_copy(c->fc1b_off, c->pk_fc1b, FC_HIDDEN);
  add_copy(c->fc2w_off, c->pk_fc2w, FC_HIDDEN);
  add_copy(c->fc2b_off, c->pk_fc2b, 1);
  c->pk_elems = pk; c->pkd_elems = pkd;
  c->n_fwd = (int)tf.size(); c->n_dgrad = (int)td.size(); c->max_fwd = maxf; c->max_dgrad = maxd;
  std::vector<UmmaPackDesc> uf, ud;
  if (c->bf)
    for (int i = 0; i < 8; ++i) {
      const Layer& l = c->L[i];
      if (umma_ok(l.Ci, l.Co)) { uf.push_back({l.pk_off, l.Ci, round_up(l.Co, 16)}); c->max_ufwd = std::max(c->max_ufwd, 9 * l.Ci * l.Co); }
      if (umma_ok(l.Co, l.Ci)) { ud.push_back({l.pkd_off, l.Co, round_up(l.Ci, 16)}); c->max_udgrad = std::max(c->max_udgrad, 9 * l.Ci * l.Co); }
    }
  c->n_ufwd = (int)uf.size(); c->n_udgrad = (int)ud.size();
  // the streaming kernel also takes the narrow layer-1 data gradient (W -> nc as N = 16 with zero pad columns)
  std::vector<UmmaPackDesc> igd = ud;
  if (c->bf)
    for (int i = 0; i < 8; ++i) {
      const Layer& l = c->L[i];
      if (!umma_ok(l.Co, l.Ci) && umma_img_ok(l.Co, l.Ci)) igd.push_back({l.pkd_off, l.Co, round_up(l.Ci, 16)});
    }
  c->n_igdgrad = (int)igd.size();
  int s = 0;
#define CA(ptr, bytes) if ((s = dev_alloc(c->pool, (void**)&(ptr), (bytes))) != 0) { dg_critic_destroy(c); return s; }
  CA(c->pk, sizeof(float) * pk);
  CA(c->gpk, sizeof(float) * pk);
  CA(c->pkd, sizeof(float) * pkd);
  CA(c->pk_u, sizeof(bf16) * (pk + 64));
  CA(c->pkd_u, sizeof(bf16) * (pkd + 64));
  CA(c->pk_ig, sizeof(bf16) * (pk + 64));
  CA(c->pkd_ig, sizeof(bf16) * (pkd + 64));
  if ((s = upload_table(c->pool, tf, &c->tab_fwd)) != 0) { dg_critic_destroy(c); return s; }
  if ((s = upload_table(c->pool, td, &c->tab_dgrad)) != 0) { dg_critic_destroy(c); return s; }
  if ((s = upload_utable(c->pool, uf, &c->utab_fwd)) != 0) { dg_critic_destroy(c); return s; }
  if ((s = upload_utable(c->pool, ud, &c->utab_dgrad)) != 0) { dg_critic_destroy(c); return s; }
  if ((s = upload_utable(c->pool, igd, &c->utab_igdgrad)) != 0) { dg_critic_destroy(c); return s; }
  const size_t NB = c->NBmax, B = c->maxB, pf = (size_t)c->Hf * c->Hf;
  CA(c->a0, NB * pf * c->nc * sizeof(float));
  size_t vmax = 0;
  for (int i = 0; i < 8; ++i) {
    const size_t e = c->pix(i) * c->L[i].Co;
    CA(c->a[i + 1], NB * e * c->esz);
    CA(c->dz[i + 1], NB * e * c->esz);
    if (c->bf && c->L[i].Co % 16 == 0) CA(c->bits[i + 1], NB * (e / 16) * sizeof(unsigned short));
    vmax = std::max(vmax, e);
  }
  CA(c->a9, NB * FC_HIDDEN * sizeof(float));
  CA(c->dz9, NB * FC_HIDDEN * sizeof(float));
  CA(c->vfc, NB * FC_HIDDEN * sizeof(float));
  CA(c->scores, NB * sizeof(float));
  CA(c->seed, NB * sizeof(float));
  CA(c->g, B * pf * c->nc * sizeof(float));
  CA(c->u, B * pf * c->nc * sizeof(float));
  CA(c->v0, B * vmax * c->esz);
  CA(c->v1, B * vmax * c->esz);
  CA(c->sumsq, B * sizeof(float));
  CA(c->coef, B * sizeof(float));
  CA(c->norms, B * sizeof(float));
  CA(c->scal, 64);
  CA(c->metric_scratch, metric_scratch_bytes());
#undef CA
  if ((s = c->side.create(1)) != 0) { dg_critic_destroy(c); return s; }
  if ((s = c->side2.create(0)) != 0) { dg_critic_destroy(c); return s; }  // (no column sums run on the second side stream)
  *out = c;
  return 0;
}
extern "C" int dg_critic_destroy(dg_critic* c) {
  if (!c) return 0;
  c->side.destroy();
  c->side2.destroy();
  for (void* p : c->pool) cudaFree(p);
  delete c;
  return 0;
}
extern "C" int dg_critic_pack(dg_critic* c, const float* params, void* stream) {
  if (c && c->pending_finish) { set_error("dg_critic_pack: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(c && params, "dg_critic_pack: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool c_ig = c->bf && g_tune[16];  // K-major images of the streaming kernel ride on the same launch
  DG_TRY(pack_weights2(params, c->pk, c->pk_u, c->tab_fwd, c->n_fwd, c->max_fwd, c->pkd, c->pkd_u, c->tab_dgrad, c->n_dgrad,
                       c->max_dgrad, st, c_ig ? c->pk_ig : nullptr, c_ig ? c->pkd_ig : nullptr));
  c->packed = true;
  c->pending_params = nullptr;
  return 0;
}
// Records that `params` must be packed before the weights are used next; the launch itself is issued by the next entry point
// that needs the weights - the fused critic iteration runs it beside its batch assembly (which does not read weights) instead
// of in front of it.  `params` must stay valid and unchanged until then (the trainer passes its flat parameter buffer).
extern "C" int dg_critic_pack_lazy(dg_critic* c, const float* params) {
  DG_CHECK(c && params, "dg_critic_pack_lazy: null argument");
  c->pending_params = params;
  return 0;
}
static int critic_flush_pack(dg_critic* c, cudaStream_t st) {
  if (!c->pending_params) return 0;
  return dg_critic_pack(c, c->pending_params, (void*)st);
}

// Critic.forward (critic.py:101-106) over samples [s0, s0 + NB) of c->a0.
static int critic_forward_internal(dg_critic* c, int NB, cudaStream_t st, int s0 = 0, bool raw_head = false) {
  TV x = tv_batch(tv(c->a0, 0, c->nc), (size_t)c->Hf * c->Hf, s0);
  for (int i = 0; i < 8; ++i) {
    const Layer& l = c->L[i];
    ConvOp op;
    op.x = x; op.Hin = c->Hin[i]; op.Win = c->Hin[i]; op.Ci = l.Ci;
    op.y = tv_batch(c->act(c->a[i + 1], l.Co), c->pix(i), s0); op.Hout = c->Hout[i]; op.Wout = c->Hout[i]; op.Co = l.Co;
    op.B = NB; op.w = c->pk + l.pk_off; op.bias = (i == 0) ? c->pk + c->pk_b0 : nullptr;
    if (c->bf) { op.w_umma = c->pk_u + l.pk_off; op.w_ig = c->pk_ig + l.pk_off; }
    op.stride = l.stride; op.act = ACT_LRELU; op.slope = C_SLOPE;
    if (c->bits[i + 1] && g_tune[21]) op.bits_out = c->bits[i + 1] + (size_t)s0 * c->pix(i) * (size_t)(l.Co >> 4);
    bool wrote = false;
    DG_TRY(run_conv(op, st, &wrote));
    c->bits_ok[i + 1] = wrote && (s0 == 0 || c->bits_ok[i + 1]);  // a later sample range keeps the flag only if every range wrote
    x = op.y;
  }
  float* a9 = c->a9 + (size_t)s0 * FC_HIDDEN;
  if (raw_head) return fc_fwd_raw(x.p, c->bf, c->pk + c->pk_fc1w, a9, NB, c->fc_in, FC_HIDDEN, st);  // critic_head finishes
  DG_TRY(fc_fwd(x.p, c->bf, c->pk + c->pk_fc1w, c->pk + c->pk_fc1b, a9, NB, c->fc_in, FC_HIDDEN, ACT_LRELU, C_SLOPE, nullptr, st));
  DG_TRY(fc2_fwd(a9, c->pk + c->pk_fc2w, c->pk + c->pk_fc2b, c->scores + s0, NB, FC_HIDDEN, st));
  return 0;
}

// Weight gradient of conv layer i over samples [s0, s0 + n) of the saved activations / dz buffers (x = a[i], dy = dz[i+1]),
// accumulated into c->gpk; `bias_n` > 0 adds the features.0 bias gradient over the first bias_n of those samples.
// With `side` the launch goes to the critic's side stream, ordered after everything enqueued on `st` so far.
static int critic_layer_wgrad(dg_critic* c, int i, int s0, int n, int bias_n, bool side, cudaStream_t st) {
  const Layer& l = c->L[i];
  WgradOp w;
  const size_t pin = (size_t)c->Hin[i] * c->Hin[i];
  w.x = tv_batch((i == 0) ? tv(c->a0, 0, c->nc) : c->act(c->a[i], l.Ci), pin, s0);
  w.Hin = c->Hin[i]; w.Win = c->Hin[i]; w.Ci = l.Ci;
  w.dy = tv_batch(c->act(c->dz[i + 1], l.Co), c->pix(i), s0); w.Hout = c->Hout[i]; w.Wout = c->Hout[i]; w.Co = l.Co;
  w.B = n; w.stride = l.stride;
  w.dw = c->gpk + l.pk_off;
  if (i == 0 && bias_n > 0) { w.dbias = c->gpk + c->pk_b0; w.dbias_B = bias_n; }
  if (!side) return run_wgrad(w, st);
  // tuning key 19: the late layers' weight gradients alternate between two side streams, so two of these latency-bound
  // launches run beside each other (and beside the JVP chain on the caller's stream)
  SideStream& ss = (g_tune[19] && c->side2.s && i >= 3 && (i & 1)) ? c->side2 : c->side;
  DG_TRY(ss.fork(st));
  return run_wgrad(w, ss.s);
}

// Input-gradient chain for samples [s0, s0 + NB) seeded by c->seed; layer-1 data
// gradient only for samples [n0, n0+n1) into g_out (NHWC fp32) when n1 > 0.
// early != 0: the weight gradient of every conv layer over the chain's samples is enqueued as soon as its dz exists
// (c->gpk must already be zeroed): early == 1 on the side stream (forked from `st`), early == 2 on `st` itself (the
// chain already runs on a stream of its own).
// early_rows > 0 (with early == 1): only rows [s0, s0 + early_rows) and only layers <= early_max_layer go early (the
// interpolates' rows of the fused 3B batch need the JVP chain first).
static int critic_backward_chain(dg_critic* c, int NB, int n0, int n1, float* g_out, cudaStream_t st, int early = 0, int s0 = 0,
                                 bool seeded = false, int early_rows = 0, int early_max_layer = 7) {
  const int en = early_rows > 0 ? early_rows : NB;
  float* dz9 = c->dz9 + (size_t)s0 * FC_HIDDEN;
  const size_t e8 = (size_t)s0 * c->fc_in;
  void* dz8 = c->bf ? (void*)((bf16*)c->dz[8] + e8) : (void*)((float*)c->dz[8] + e8);
  const void* a8 = c->bf ? (const void*)((const bf16*)c->a[8] + e8) : (const void*)((const float*)c->a[8] + e8);
  if (!seeded) DG_TRY(fc2_seed(c->a9 + (size_t)s0 * FC_HIDDEN, c->pk + c->pk_fc2w, c->seed + s0, dz9, NB, FC_HIDDEN, C_SLOPE, st));
  DG_TRY(fc_dgrad(dz9, c->pk + c->pk_fc1w, dz8, c->bf, NB, c->fc_in, FC_HIDDEN, a8, c->bf, C_SLOPE, st));
  if (early && early_max_layer >= 7) DG_TRY(critic_layer_wgrad(c, 7, s0, en, 0, early == 1, st));
  for (int i = 7; i >= 1; --i) {  // dz[i] = dgrad_{i+1}(dz[i+1]) * lrelu'(a[i])
    const Layer& l = c->L[i];
    ConvOp op;
    op.x = tv_batch(c->act(c->dz[i + 1], l.Co), c->pix(i), s0); op.Hin = c->Hout[i]; op.Win = c->Hout[i]; op.Ci = l.Co;
    op.y = tv_batch(c->act(c->dz[i], l.Ci), c->pix(i - 1), s0); op.Hout = c->Hin[i]; op.Wout = c->Hin[i]; op.Co = l.Ci;
    op.B = NB; op.w = c->pkd + l.pkd_off;
    if (c->bf) { op.w_umma = c->pkd_u + l.pkd_off; op.w_ig = c->pkd_ig + l.pkd_off; }
    op.transposed = (l.stride == 2);
    op.act = ACT_MASK; op.slope = C_SLOPE; op.mask = tv_batch(c->act(c->a[i], l.Ci), c->pix(i - 1), s0);
    op.bits_in = c->bits_at(i, s0);
    DG_TRY(run_conv(op, st));
    if (early && i - 1 <= early_max_layer) DG_TRY(critic_layer_wgrad(c, i - 1, s0, en, en, early == 1, st));
  }
  if (n1 > 0) {
    const Layer& l = c->L[0];
    ConvOp op;
    op.x = tv_batch(c->act(c->dz[1], l.Co), c->pix(0), n0); op.Hin = c->Hout[0]; op.Win = c->Hout[0]; op.Ci = l.Co;
    op.y = tv(g_out, 0, l.Ci); op.Hout = c->Hin[0]; op.Wout = c->Hin[0]; op.Co = l.Ci;
    op.B = n1; op.w = c->pkd + l.pkd_off;
    if (c->bf && umma_img_ok(l.Co, l.Ci)) { op.w_umma = c->pkd_u + l.pkd_off; op.w_ig = c->pkd_ig + l.pkd_off; op.narrow_ok = 1; }
    DG_TRY(run_conv(op, st));
  }
  return 0;
}

// Weight gradients with x = saved activations, dy = dz, samples [n0, n0+n).
static int critic_wgrads_acts(dg_critic* c, int n0, int n, cudaStream_t st) {
  for (int i = 0; i < 8; ++i) {
    const Layer& l = c->L[i];
    WgradOp w;
    const size_t pin = (size_t)c->Hin[i] * c->Hin[i];
    w.x = (i == 0) ? tv_batch(tv(c->a0, 0, c->nc), pin, n0) : tv_batch(c->act(c->a[i], l.Ci), pin, n0);
    w.Hin = c->Hin[i]; w.Win = c->Hin[i]; w.Ci = l.Ci;
    w.dy = tv_batch(c->act(c->dz[i + 1], l.Co), c->pix(i), n0); w.Hout = c->Hout[i]; w.Wout = c->Hout[i]; w.Co = l.Co;
    w.B = n; w.stride = l.stride;
    w.dw = c->gpk + l.pk_off; w.dbias = (i == 0) ? c->gpk + c->pk_b0 : nullptr;
    DG_TRY(run_wgrad(w, st));
  }
  const size_t e8 = (size_t)c->fc_in;
  const void* a8 = c->bf ? (const void*)((bf16*)c->a[8] + (size_t)n0 * e8) : (const void*)((float*)c->a[8] + (size_t)n0 * e8);
  DG_TRY(fc_wgrad(c->dz9 + (size_t)n0 * FC_HIDDEN, a8, c->bf, c->gpk + c->pk_fc1w, n, c->fc_in, FC_HIDDEN, st));
  DG_TRY(colsum(tv(c->dz9 + (size_t)n0 * FC_HIDDEN, 0, FC_HIDDEN), (size_t)n, FC_HIDDEN, c->gpk + c->pk_fc1b, st));
  // classifier.2: dW2 = sum_b seed[b] * a9[b], db2 = sum_b seed[b]   (expressed as a 1-row fc_wgrad / colsum)
  DG_TRY(fc_wgrad(c->seed + n0, c->a9 + (size_t)n0 * FC_HIDDEN, 0, c->gpk + c->pk_fc2w, n, FC_HIDDEN, 1, st));
  DG_TRY(colsum(tv(c->seed + n0, 0, 1), (size_t)n, 1, c->gpk + c->pk_fc2b, st));
  return 0;
}

// Gradient-penalty double backward: v_0 = u, dW_l += wgrad(v_{l-1}, dz_l),
// v_l = m_l * conv_l(v_{l-1}); the interpolates are samples [n0, n0+B) of the saved batch.
static int critic_gp_second_order(dg_critic* c, int n0, int B, cudaStream_t st) {
  TV v = tv(c->u, 0, c->nc);
  void* pp[2] = {c->v0, c->v1};
  for (int i = 0; i < 8; ++i) {
    const Layer& l = c->L[i];
    WgradOp w;
    w.x = v; w.Hin = c->Hin[i]; w.Win = c->Hin[i]; w.Ci = l.Ci;
    w.dy = tv_batch(c->act(c->dz[i + 1], l.Co), c->pix(i), n0); w.Hout = c->Hout[i]; w.Wout = c->Hout[i]; w.Co = l.Co;
    w.B = B; w.stride = l.stride; w.dw = c->gpk + l.pk_off; w.dbias = nullptr;
    DG_TRY(run_wgrad(w, st));
    ConvOp op;
    op.x = v; op.Hin = c->Hin[i]; op.Win = c->Hin[i]; op.Ci = l.Ci;
    op.y = c->act(pp[i & 1], l.Co); op.Hout = c->Hout[i]; op.Wout = c->Hout[i]; op.Co = l.Co;
    op.B = B; op.w = c->pk + l.pk_off; op.bias = nullptr; op.stride = l.stride;
    if (c->bf) { op.w_umma = c->pk_u + l.pk_off; op.w_ig = c->pk_ig + l.pk_off; }
    op.act = ACT_MASK; op.slope = C_SLOPE; op.mask = tv_batch(c->act(c->a[i + 1], l.Co), c->pix(i), n0);
    op.bits_in = c->bits_at(i + 1, n0);
    DG_TRY(run_conv(op, st));
    v = op.y;
  }
  DG_TRY(fc_wgrad(c->dz9 + (size_t)n0 * FC_HIDDEN, v.p, c->bf, c->gpk + c->pk_fc1w, B, c->fc_in, FC_HIDDEN, st));
  DG_TRY(fc_fwd(v.p, c->bf, c->pk + c->pk_fc1w, nullptr, c->vfc, B, c->fc_in, FC_HIDDEN, ACT_MASK, C_SLOPE,
                c->a9 + (size_t)n0 * FC_HIDDEN, st));
  DG_TRY(colsum(tv(c->vfc, 0, FC_HIDDEN), (size_t)B, FC_HIDDEN, c->gpk + c->pk_fc2w, st));
  return 0;
}

static int critic_gp_first_order(dg_critic* c, const dg_hyper* hp, int n0, int B, float* scalars, int write_loss,
                                 float* norms_out, cudaStream_t st, float* u_out = nullptr) {
  const size_t per = (size_t)c->Hf * c->Hf * c->nc;
  DG_TRY(gp_norms_finish(c->g, B, per, hp->gp_lambda, c->sumsq, norms_out ? norms_out : c->norms, c->coef, scalars, write_loss, st));
  DG_TRY(gp_scale(c->g, c->coef, u_out ? u_out : c->u, B, per, st));
  (void)n0;
  return 0;
}

// Fused-step variant of the two functions above.  The interpolates occupy samples [2B, 3B) of every
// saved activation; once the input-gradient chain is done their activations are only needed as
// LeakyReLU masks, so the JVP chain v_l = m_l * conv_l(v_{l-1}) is written IN PLACE over them (each
// epilogue thread reads the mask element and overwrites the same element).  Afterwards
// a[l] = [real acts ; fake acts ; v_l] and dz[l+1] = [dz real ; dz fake ; dz interpolates], so ONE
// weight-gradient launch per layer over all 3B samples yields  d(E[C(fake)] - E[C(real)])/dW + dGP/dW.
static int critic_second_order_and_wgrads(dg_critic* c, int B, cudaStream_t st, bool two_chain, int early_max_layer = -1,
                                          bool defer_join = false) {
  const int n0 = 2 * B;
  // The weight gradient of layer i reads a[i] = [real ; fake ; v_{i-1}] and dz[i+1]; nothing later in the call
  // overwrites either, so with the side stream on it is enqueued there as soon as v_{i-1} exists and overlaps the
  // rest of the JVP chain and the classifier kernels (joined here, before the caller's unpack).
  // two_chain: the real + fake rows (forward, input-gradient chain, their weight-gradient rows, classifier rows) are
  // running on the side stream; this stream only adds the interpolates' rows and joins at the end.
  const bool side_on = !two_chain && g_tune[9] && c->side.s != nullptr;
  auto layer_wgrad = [&](int i) {
    // features.0.bias: real + fake samples only (the gradient penalty contributes exactly zero to biases)
    if (two_chain) return critic_layer_wgrad(c, i, n0, B, 0, false, st);
    // layers whose real + fake rows already went early (during the input-gradient chain): the interpolates' rows remain
    if (i <= early_max_layer) return critic_layer_wgrad(c, i, n0, B, 0, side_on, st);
    return critic_layer_wgrad(c, i, 0, 3 * B, n0, side_on, st);
  };
  TV v = tv_batch(tv(c->a0, 0, c->nc), (size_t)c->Hf * c->Hf, n0);  // u was written here by gp_scale
  c->keep_batch = 0;
  if (g_tune[15]) {  // parity instrumentation: the JVP chain below overwrites the interpolates' activations in place
    for (int i = 0; i < 8; ++i) {
      const size_t e = c->pix(i) * c->L[i].Co * c->esz;
      if (!c->keep[i + 1]) DG_TRY(dev_alloc(c->pool, &c->keep[i + 1], (size_t)c->maxB * e));
      DG_CUDA(cudaMemcpyAsync(c->keep[i + 1], (const char*)c->a[i + 1] + (size_t)n0 * e, (size_t)B * e, cudaMemcpyDeviceToDevice, st));
    }
    c->keep_batch = B;
  }
  for (int i = 0; i < 8; ++i) {
    if (side_on) DG_TRY(layer_wgrad(i));
    const Layer& l = c->L[i];
    ConvOp op;
    op.x = v; op.Hin = c->Hin[i]; op.Win = c->Hin[i]; op.Ci = l.Ci;
    op.y = tv_batch(c->act(c->a[i + 1], l.Co), c->pix(i), n0); op.Hout = c->Hout[i]; op.Wout = c->Hout[i]; op.Co = l.Co;
    op.B = B; op.w = c->pk + l.pk_off; op.bias = nullptr; op.stride = l.stride;
    if (c->bf) { op.w_umma = c->pk_u + l.pk_off; op.w_ig = c->pk_ig + l.pk_off; }
    op.act = ACT_MASK; op.slope = C_SLOPE; op.mask = op.y;
    op.bits_in = c->bits_at(i + 1, n0);  // (the in-place JVP then never reads the activation it overwrites)
    DG_TRY(run_conv(op, st));
    v = op.y;
  }
  // classifier: dW_fc1 over all 3B rows at once (two_chain: the interpolates' rows); v_fc and dW_fc2 as in critic_gp_second_order
  if (two_chain) DG_TRY(fc_wgrad(c->dz9 + (size_t)n0 * FC_HIDDEN, v.p, c->bf, c->gpk + c->pk_fc1w, B, c->fc_in, FC_HIDDEN, st));
  else DG_TRY(fc_wgrad(c->dz9, c->a[8], c->bf, c->gpk + c->pk_fc1w, 3 * B, c->fc_in, FC_HIDDEN, st));
  DG_TRY(fc_fwd(v.p, c->bf, c->pk + c->pk_fc1w, nullptr, c->vfc, B, c->fc_in, FC_HIDDEN, ACT_MASK, C_SLOPE,
                c->a9 + (size_t)n0 * FC_HIDDEN, st));
  if (two_chain) {
    for (int i = 0; i < 8; ++i) DG_TRY(layer_wgrad(i));
    DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));
  }
  // classifier.0.bias, classifier.2.weight (incl. the GP term sum_b v_fc) and classifier.2.bias in one launch
  DG_TRY(critic_small_grads(c->dz9, c->seed, c->a9, c->vfc, n0, B, FC_HIDDEN, c->gpk + c->pk_fc1b, c->gpk + c->pk_fc2w,
                            c->gpk + c->pk_fc2b, st));
  if (!two_chain) {
    if (!side_on)
      for (int i = 0; i < 8; ++i) DG_TRY(layer_wgrad(i));
    if (!defer_join) DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));  // deferred: dg_critic_step_finish joins
  }
  return 0;
}

extern "C" int dg_critic_fwd(dg_critic* c, const float* x, int batch, float* scores, void* stream) {
  if (c && c->pending_finish) { set_error("dg_critic_fwd: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(c && x && scores, "dg_critic_fwd: null argument");
  DG_CHECK(batch >= 1 && batch <= c->NBmax, "dg_critic_fwd: batch %d outside [1,%d]", batch, c->NBmax);
  if (c) DG_TRY(critic_flush_pack(c, (cudaStream_t)stream));  // a pending dg_critic_pack_lazy
  if (!c->packed) { set_error("dg_critic_fwd: dg_critic_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(nchw_to_nhwc(x, tv(c->a0, 0, c->nc), batch, c->nc, c->Hf, c->Hf, st));
  DG_TRY(critic_forward_internal(c, batch, st));
  DG_CUDA(cudaMemcpyAsync(scores, c->scores, sizeof(float) * batch, cudaMemcpyDeviceToDevice, st));
  c->saved_batch = batch;
  return 0;
}

extern "C" int dg_critic_activation(dg_critic* c, int which, int s0, int batch, float* out, void* stream) {
  DG_CHECK(c && out, "dg_critic_activation: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));
  if (which >= 101 && which <= 108) {
    const int i = which - 101;
    DG_CHECK(s0 >= 0 && batch >= 1 && s0 + batch <= c->keep_batch,
             "dg_critic_activation: samples [%d,%d) outside the %d interpolates kept (dg_set_tuning(15, 1) before the iteration)", s0,
             s0 + batch, c->keep_batch);
    return nhwc_to_nchw(tv_batch(c->act(c->keep[i + 1], c->L[i].Co), c->pix(i), s0), out, batch, c->L[i].Co, c->Hout[i], c->Hout[i], st);
  }
  DG_CHECK(s0 >= 0 && batch >= 1 && s0 + batch <= c->NBmax, "dg_critic_activation: samples [%d,%d) outside [0,%d)", s0, s0 + batch,
           c->NBmax);
  if (which >= 1 && which <= 8) {
    const int i = which - 1;
    return nhwc_to_nchw(tv_batch(c->act(c->a[i + 1], c->L[i].Co), c->pix(i), s0), out, batch, c->L[i].Co, c->Hout[i], c->Hout[i], st);
  }
  if (which == 9) {
    DG_CUDA(cudaMemcpyAsync(out, c->a9 + (size_t)s0 * FC_HIDDEN, sizeof(float) * batch * FC_HIDDEN, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  set_error("dg_critic_activation: unknown activation %d", which);
  return DG_ERR_INVALID;
}

extern "C" int dg_critic_bwd(dg_critic* c, const float* d_scores, float* grads_flat, float* d_x, void* stream) {
  if (c && c->pending_finish) { set_error("dg_critic_bwd: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(c && d_scores, "dg_critic_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(critic_flush_pack(c, st));  // a pending dg_critic_pack_lazy
  const int B = c->saved_batch;
  if (B <= 0) { set_error("dg_critic_bwd: no saved forward"); return DG_ERR_STATE; }
  DG_CHECK(!d_x || B <= c->maxB, "dg_critic_bwd: d_x needs batch <= max_batch");
  DG_CUDA(cudaMemcpyAsync(c->seed, d_scores, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  DG_TRY(critic_backward_chain(c, B, 0, d_x ? B : 0, c->g, st));
  if (d_x) DG_TRY(nhwc_to_nchw(tv(c->g, 0, c->nc), d_x, B, c->nc, c->Hf, c->Hf, st));
  if (grads_flat) {
    DG_CUDA(cudaMemsetAsync(c->gpk, 0, sizeof(float) * c->pk_elems, st));
    DG_TRY(critic_wgrads_acts(c, 0, B, st));
    DG_TRY(unpack_wgrads(c->gpk, grads_flat, c->tab_fwd, c->n_fwd, c->max_fwd, st));
  }
  return 0;
}

// WassersteinGAN._gp (wasserstein.py:87-117).
extern "C" int dg_gp(dg_critic* c, const dg_hyper* hp, const float* real, const float* fake, const float* alpha, int batch,
                     float* gp_out, float* norms, float* grads_flat, void* stream) {
  if (c && c->pending_finish) { set_error("dg_gp: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(c && hp && real && fake && alpha && gp_out, "dg_gp: null argument");
  DG_CHECK(batch >= 1 && batch <= c->maxB, "dg_gp: batch %d outside [1,%d]", batch, c->maxB);
  if (c) DG_TRY(critic_flush_pack(c, (cudaStream_t)stream));  // a pending dg_critic_pack_lazy
  if (!c->packed) { set_error("dg_gp: dg_critic_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(build_critic_input(real, fake, 1, alpha, c->a0, batch, c->nc, c->Hf, c->Hf, 1, st));
  DG_TRY(critic_forward_internal(c, batch, st));
  DG_TRY(fill(c->seed, 1.f, batch, st));
  DG_TRY(critic_backward_chain(c, batch, 0, batch, c->g, st));
  DG_TRY(critic_gp_first_order(c, hp, 0, batch, c->scal, 0, norms, st));
  DG_CUDA(cudaMemcpyAsync(gp_out, c->scal + 3, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (grads_flat) {
    DG_CUDA(cudaMemsetAsync(c->gpk, 0, sizeof(float) * c->pk_elems, st));
    DG_TRY(critic_gp_second_order(c, 0, batch, st));
    DG_TRY(unpack_wgrads(c->gpk, grads_flat, c->tab_fwd, c->n_fwd, c->max_fwd, st));
  }
  c->saved_batch = 0;
  return 0;
}

// ===========================================================================
// fused iterations
// ===========================================================================
__global__ void critic_seed_kernel(float* seed, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * B) seed[i] = i < B ? -1.f / B : (i < 2 * B ? 1.f / B : 1.f);
}

// _critic_train_iteration, wasserstein.py:35-52 (everything except C_optimizer.step()), given fake = G(coarse)
// as an NHWC fp32 tensor.
static int critic_step_body(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* fake_nhwc, const float* fine,
                            const float* alpha, int B, float* c_grads_flat, float* scalars, cudaStream_t st) {
  if (hp->freq_sep) {
    // frequency separation (GAN/wasserstein_fs.py:41-51): the critic and the penalty see the high-pass parts
    // x - low(x), low = AvgPool2d(filter_size, 1) o ReplicationPad2d(filter_size // 2)   (config/hyperparams.py:31-35)
    DG_CHECK(hp->filter_size >= 1 && (hp->filter_size & 1), "frequency separation: filter_size %d must be odd", hp->filter_size);
    const size_t fb = (size_t)c->maxB * c->Hf * c->Hf * c->nc * sizeof(float);
    if (!c->fs_real) { DG_TRY(dev_alloc(c->pool, (void**)&c->fs_real, fb)); DG_TRY(dev_alloc(c->pool, (void**)&c->fs_fake, fb)); }
    DG_TRY(lowpass_replicate(fine, c->fs_real, (long long)B * c->nc, c->Hf, c->Hf, 1, hp->filter_size / 2, 1, st));       // NCHW planes
    DG_TRY(lowpass_replicate(fake_nhwc, c->fs_fake, B, c->Hf, c->Hf, c->nc, hp->filter_size / 2, 1, st));                 // NHWC
    fine = c->fs_real;
    fake_nhwc = c->fs_fake;
  }
  // one 3B critic batch: [real ; fake ; alpha*real + (1-alpha)*fake]   (:37-38, :91-97).  A pending weight pack
  // (dg_critic_pack_lazy) runs beside it: the assembly reads no weights, so it goes to the side stream while the pack launch
  // occupies the caller's stream (dg_set_tuning(25, 0): pack first, then assemble).
  if (c->pending_params && g_tune[25] && c->side.s != nullptr) {
    DG_TRY(c->side.fork(st));
    DG_TRY(build_critic_input(fine, fake_nhwc, 0, alpha, c->a0, B, c->nc, c->Hf, c->Hf, 0, c->side.s));
    DG_TRY(critic_flush_pack(c, st));
    DG_TRY(c->side.join(st));
  } else {
    DG_TRY(critic_flush_pack(c, st));
    DG_TRY(build_critic_input(fine, fake_nhwc, 0, alpha, c->a0, B, c->nc, c->Hf, c->Hf, 0, st));
  }
  const bool two_chain = g_tune[9] >= 2 && c->side.s != nullptr;
  const bool head = !two_chain && g_tune[11] && critic_head_supported(B) && fc_fwd_raw_supported(c->fc_in, FC_HIDDEN);
  // g_tune[13] = k > 0: the real + fake rows of the weight gradients of layers 0 .. k-1 (the HBM-bound ones, which scale with
  // the rows) run on the side stream during the input-gradient chain, where that stream is otherwise idle
  const int early_ml = (!two_chain && g_tune[9] == 1 && c->side.s != nullptr) ? g_tune[13] - 1 : -1;
  if (!head) {
    critic_seed_kernel<<<(3 * B + 255) / 256, 256, 0, st>>>(c->seed, B);
    DG_LAUNCH_CHECK();
  }
  DG_CUDA(cudaMemsetAsync(c->gpk, 0, sizeof(float) * c->pk_elems, st));
  // g_tune[9] == 2: two chains over disjoint sample ranges.  The real + fake rows [0, 2B) run forward, input-gradient
  // chain (layers 8..2) and their weight-gradient rows on the side stream; the interpolates [2B, 3B) run forward,
  // input-gradient chain down to the input, penalty, JVP chain and their weight-gradient rows on the caller's stream.
  // The kernels are latency- rather than bandwidth-bound, so the two chains fill each other's ramp-up and tail.
  if (two_chain) {
    cudaStream_t sa = c->side.s;
    DG_TRY(c->side.fork(st));
    DG_TRY(critic_forward_internal(c, 2 * B, sa, 0));
    DG_TRY(critic_means(c->scores, B, scalars, sa));
    DG_CUDA(cudaEventRecord(c->side.aux_ev, sa));
    DG_TRY(critic_backward_chain(c, 2 * B, 0, 0, nullptr, sa, 2, 0));
    {
      DG_TRY(fc_wgrad(c->dz9, c->a[8], c->bf, c->gpk + c->pk_fc1w, 2 * B, c->fc_in, FC_HIDDEN, sa));
    }
    DG_TRY(critic_forward_internal(c, B, st, 2 * B));
    DG_TRY(critic_backward_chain(c, B, 2 * B, B, c->g, st, 0, 2 * B));
    DG_CUDA(cudaStreamWaitEvent(st, c->side.aux_ev, 0));  // the loss scalar needs the real / fake means
  } else if (head) {
    DG_TRY(critic_forward_internal(c, 3 * B, st, 0, true));
    DG_TRY(critic_head(c->a9, c->pk + c->pk_fc1b, c->pk + c->pk_fc2w, c->pk + c->pk_fc2b, c->scores, c->seed, c->dz9, scalars, B,
                       FC_HIDDEN, C_SLOPE, st));
    DG_TRY(critic_backward_chain(c, 3 * B, 2 * B, B, c->g, st, early_ml >= 0 ? 1 : 0, 0, true, 2 * B, early_ml));
  } else {
    DG_TRY(critic_forward_internal(c, 3 * B, st));
    DG_TRY(critic_means(c->scores, B, scalars, st));
    DG_TRY(critic_backward_chain(c, 3 * B, 2 * B, B, c->g, st, early_ml >= 0 ? 1 : 0, 0, false, 2 * B, early_ml));
  }
  DG_TRY(critic_gp_first_order(c, hp, 2 * B, B, scalars, 1, nullptr, st,
                               c->a0 + (size_t)2 * B * c->Hf * c->Hf * c->nc));  // u overwrites the interpolates
  const bool defer = c->defer_conv && !two_chain && c->n_fwd > dg_critic::N_CONV_ENTRIES;
  // tuning key 24: the classifier gradients (74 % of the flat buffer, a transposing unpack) are final on this stream while the
  // conv weight gradients still run on the side stream - unpack them now, join, then unpack the nine conv entries
  const bool split_unpack = !defer && g_tune[24] && !two_chain && g_tune[9] && c->side.s != nullptr && c->n_fwd > dg_critic::N_CONV_ENTRIES;
  DG_TRY(critic_second_order_and_wgrads(c, B, st, two_chain, early_ml, defer || split_unpack));
  if (split_unpack) {
    DG_TRY(unpack_wgrads(c->gpk, c_grads_flat, c->tab_fwd + dg_critic::N_CONV_ENTRIES, c->n_fwd - dg_critic::N_CONV_ENTRIES, c->max_fwd, st));
    DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));
    DG_TRY(unpack_wgrads(c->gpk, c_grads_flat, c->tab_fwd, dg_critic::N_CONV_ENTRIES, c->max_fwd, st));
  } else if (defer) {
    // the classifier gradients (74 % of the bucket) are final on this stream: unpack them now so that the caller can start
    // their all-reduce while the conv weight gradients are still running on the side stream
    DG_TRY(unpack_wgrads(c->gpk, c_grads_flat, c->tab_fwd + dg_critic::N_CONV_ENTRIES, c->n_fwd - dg_critic::N_CONV_ENTRIES, c->max_fwd, st));
    c->pending_finish = true;
  } else {
    DG_TRY(unpack_wgrads(c->gpk, c_grads_flat, c->tab_fwd, c->n_fwd, c->max_fwd, st));
  }
  c->saved_batch = 0;
  (void)g;
  return 0;
}

extern "C" int dg_critic_step(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* coarse, const float* fine,
                              const float* alpha, int batch, float* c_grads_flat, float* scalars, void* stream) {
  if (c && c->pending_finish) { set_error("dg_critic_step: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(g && c && hp && coarse && fine && alpha && c_grads_flat && scalars, "dg_critic_step: null argument");
  DG_CHECK(batch >= 1 && batch <= g->maxB && batch <= c->maxB, "dg_critic_step: batch %d too large", batch);
  DG_CHECK(g->Hf == c->Hf && g->Cout == c->nc, "dg_critic_step: generator output does not match critic input");
  if (!g->packed || !(c->packed || c->pending_params)) { set_error("dg_critic_step: weights not packed"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  const int B = batch;
  DG_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(float), st));
  // fake = G(coarse): the reference keeps the graph (:35) but discards the generator gradients (:65)
  DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), B, g->Cin, g->Hc, g->Hc, st));
  DG_TRY(gen_forward_internal(g, B, 0, st));
  g->saved_batch = 0;
  g->lookahead = 0;
  return critic_step_body(g, c, hp, g->fake, fine, alpha, B, c_grads_flat, scalars, st);
}

// Data-parallel overlap: with the switch on, dg_critic_step / dg_critic_step_fake return once the classifier gradients are
// unpacked (the conv weight gradients still run on the side stream); the caller starts the all-reduce of
// c_grads_flat[dg_critic_param_offset(cfg, 9) ..] and then calls dg_critic_step_finish, which joins the side stream and
// unpacks the conv gradients c_grads_flat[.. offset).
extern "C" int dg_critic_defer_conv_grads(dg_critic* c, int on) {
  DG_CHECK(c, "dg_critic_defer_conv_grads: null argument");
  if (c->pending_finish) { set_error("dg_critic_defer_conv_grads: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  c->defer_conv = on ? 1 : 0;
  return 0;
}
extern "C" int dg_critic_step_finish(dg_critic* c, float* c_grads_flat, void* stream) {
  DG_CHECK(c && c_grads_flat, "dg_critic_step_finish: null argument");
  if (!c->pending_finish) { set_error("dg_critic_step_finish: no deferred critic iteration"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));
  DG_TRY(unpack_wgrads(c->gpk, c_grads_flat, c->tab_fwd, dg_critic::N_CONV_ENTRIES, c->max_fwd, st));
  c->pending_finish = false;
  return 0;
}

// Look-ahead generator forward: the critic iterations between two generator updates all see the SAME generator
// weights (wasserstein.py:131-147), so fake = G(coarse) of the next `total` samples (several batches, concatenated)
// is computed in one pass: the persistent trunk kernel then has one CTA per sample for ALL of them (a single batch
// of 64 fills 64 of the 148 SMs).  The fakes stay in the generator's NHWC output buffer until the next
// generator forward; dg_critic_step_fake consumes them by sample offset.
extern "C" int dg_generator_lookahead(dg_generator* g, const float* coarse, int total, int save_first, void* stream) {
  DG_CHECK(g && coarse, "dg_generator_lookahead: null argument");
  DG_CHECK(total >= 1 && total <= g->maxB, "dg_generator_lookahead: %d samples outside [1,%d]", total, g->maxB);
  DG_CHECK(save_first >= 0 && save_first <= total, "dg_generator_lookahead: save_first %d outside [0,%d]", save_first, total);
  if (!g->packed) { set_error("dg_generator_lookahead: dg_generator_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), total, g->Cin, g->Hc, g->Hc, st));
  DG_TRY(gen_forward_internal(g, total, save_first, st));  // sets saved_batch = save_first
  g->lookahead = total;
  return 0;
}

extern "C" int dg_generator_lookahead_first(dg_generator* g, const float* coarse, int total, int save_first, int first_offset,
                                            int first_count, void* stream) {
  DG_CHECK(g && coarse, "dg_generator_lookahead_first: null argument");
  DG_CHECK(total >= 1 && total <= g->maxB, "dg_generator_lookahead_first: %d samples outside [1,%d]", total, g->maxB);
  DG_CHECK(first_offset >= 0 && first_count >= 1 && first_offset + first_count <= total,
           "dg_generator_lookahead_first: first range [%d,%d) outside [0,%d)", first_offset, first_offset + first_count, total);
  DG_CHECK(save_first >= 0 && save_first <= first_offset, "dg_generator_lookahead_first: save_first %d > first_offset %d", save_first,
           first_offset);
  if (!g_tune[12] || g->side.s == nullptr || first_count == total)
    return dg_generator_lookahead(g, coarse, total, save_first, stream);
  if (!g->packed) { set_error("dg_generator_lookahead_first: dg_generator_pack has not been called"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));
  DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), total, g->Cin, g->Hc, g->Hc, st));
  DG_TRY(g->side.fork(st));
  // the next critic iteration's fakes in stream order ...
  DG_TRY(gen_forward_range(g, first_offset, first_count, 0, st));
  // ... everything else (the generator iteration's batch with its saved activations first) behind it on the side stream
  if (first_offset > 0) DG_TRY(gen_forward_range(g, 0, first_offset, save_first, g->side.s));
  if (first_offset + first_count < total)
    DG_TRY(gen_forward_range(g, first_offset + first_count, total - first_offset - first_count, 0, g->side.s));
  g->saved_batch = save_first;
  g->lookahead = total;
  g->ready_lo = first_offset; g->ready_hi = first_offset + first_count;
  return 0;  // side.dirty stays set: consumers join
}

extern "C" int dg_critic_step_fake(dg_generator* g, dg_critic* c, const dg_hyper* hp, int fake_offset, const float* fine,
                                   const float* alpha, int batch, float* c_grads_flat, float* scalars, void* stream) {
  if (c && c->pending_finish) { set_error("dg_critic_step_fake: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(g && c && hp && fine && alpha && c_grads_flat && scalars, "dg_critic_step_fake: null argument");
  DG_CHECK(batch >= 1 && batch <= c->maxB, "dg_critic_step_fake: batch %d too large", batch);
  DG_CHECK(g->Hf == c->Hf && g->Cout == c->nc, "dg_critic_step_fake: generator output does not match critic input");
  if (fake_offset < 0 || fake_offset + batch > g->lookahead) {
    set_error("dg_critic_step_fake: samples [%d,%d) are not covered by the last dg_generator_lookahead (%d samples)", fake_offset,
              fake_offset + batch, g->lookahead);
    return DG_ERR_STATE;
  }
  if (!(c->packed || c->pending_params)) { set_error("dg_critic_step_fake: weights not packed"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  // fakes outside the range dg_generator_lookahead_first computed in stream order: wait for the deferred chain
  if (g->side.dirty && !(fake_offset >= g->ready_lo && fake_offset + batch <= g->ready_hi)) DG_TRY(g->side.join(st));
  DG_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(float), st));
  const float* fake = g->fake + (size_t)fake_offset * g->Hf * g->Hf * g->Cout;
  return critic_step_body(g, c, hp, fake, fine, alpha, batch, c_grads_flat, scalars, st);
}

// _generator_train_iteration, wasserstein.py:65-80 (everything except G_optimizer.step()).
// everything of _generator_train_iteration after fake = G(coarse) (activations of `batch` samples saved in g)
static int generator_step_body(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* fine, int B, float* g_grads_flat,
                               float* scalars, cudaStream_t st) {
  const long long n = (long long)B * g->Hf * g->Hf * g->Cout;
  const int R = hp->filter_size / 2;
  if (hp->freq_sep) {
    DG_CHECK(hp->filter_size >= 1 && (hp->filter_size & 1), "frequency separation: filter_size %d must be odd", hp->filter_size);
    const size_t fb = (size_t)c->maxB * c->Hf * c->Hf * c->nc * sizeof(float);
    if (!c->fs_real) { DG_TRY(dev_alloc(c->pool, (void**)&c->fs_real, fb)); DG_TRY(dev_alloc(c->pool, (void**)&c->fs_fake, fb)); }
    // c_fake = C(fake - low(fake))   (wasserstein_fs.py:75-82)
    DG_TRY(lowpass_replicate(g->fake, c->a0, B, g->Hf, g->Hf, g->Cout, R, 1, st));
  } else {
    // c_fake = C(fake)
    DG_CUDA(cudaMemcpyAsync(c->a0, g->fake, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  }
  // adversarial seed d(-gamma*mean)/dscore = -gamma/B
  DG_TRY(critic_forward_internal(c, B, st));
  DG_TRY(fill(c->seed, -hp->gamma / B, B, st));
  DG_TRY(critic_backward_chain(c, B, 0, B, c->g, st));
  // content loss + its seed, added to the adversarial input-gradient   (wasserstein.py:78, losses.py:51-53)
  DG_TRY(nchw_to_nhwc(fine, tv(g->fine_nhwc, 0, g->Cout), B, g->Cout, g->Hf, g->Hf, st));
  if (hp->freq_sep) {
    // L1 on the low-pass parts (wasserstein_fs.py:88).  With L = low (linear) and gh = dAdv/d(fake_high):
    //   dLoss/dfake = (I - L^T) gh + L^T s,  s = content_lambda * sign(low(fake) - low(fine)) / n   =  gh + L^T (s - gh)
    DG_TRY(lowpass_replicate(g->fake, c->fs_fake, B, g->Hf, g->Hf, g->Cout, R, 0, st));
    DG_TRY(lowpass_replicate(g->fine_nhwc, c->fs_real, B, g->Hf, g->Hf, g->Cout, R, 0, st));
    DG_TRY(l1_loss(c->fs_fake, c->fs_real, n, hp->content_lambda, g->l1, g->dfake, nullptr, st));                    // dfake = s
    DG_TRY(scale_add(tv(c->fs_fake, 0, 1), tv(g->dfake, 0, 1), 1.f, tv(c->g, 0, 1), -1.f, (size_t)n, 1, st));       // s - gh
    DG_TRY(lowpass_replicate(c->fs_fake, c->fs_real, B, g->Hf, g->Hf, g->Cout, R, 2, st));                           // L^T (s - gh)
    DG_TRY(scale_add(tv(g->dfake, 0, 1), tv(c->g, 0, 1), 1.f, tv(c->fs_real, 0, 1), 1.f, (size_t)n, 1, st));        // gh + ...
  } else {
    DG_TRY(l1_loss(g->fake, g->fine_nhwc, n, hp->content_lambda, g->l1, g->dfake, c->g, st));
  }
  DG_TRY(gen_scalars(c->scores, B, g->l1, hp->gamma, hp->content_lambda, scalars, st));
  DG_TRY(gen_backward_internal(g, g_grads_flat, nullptr, st));
  c->saved_batch = 0;
  g->lookahead = 0;  // the backward reuses activation buffers of the look-ahead pass
  return 0;
}

extern "C" int dg_generator_step(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* coarse, const float* fine,
                                 int batch, float* g_grads_flat, float* scalars, void* stream) {
  if (c && c->pending_finish) { set_error("dg_generator_step: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(g && c && hp && coarse && fine && g_grads_flat && scalars, "dg_generator_step: null argument");
  DG_CHECK(batch >= 1 && batch <= g->maxB && batch <= c->maxB, "dg_generator_step: batch %d too large", batch);
  DG_CHECK(g->Hf == c->Hf && g->Cout == c->nc, "dg_generator_step: generator output does not match critic input");
  if (c) DG_TRY(critic_flush_pack(c, (cudaStream_t)stream));  // a pending dg_critic_pack_lazy
  if (!g->packed || !c->packed) { set_error("dg_generator_step: weights not packed"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  const int B = batch;
  DG_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(float), st));
  DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), B, g->Cin, g->Hc, g->Hc, st));
  g->lookahead = 0;
  DG_TRY(gen_forward_internal(g, B, B, st));
  return generator_step_body(g, c, hp, fine, B, g_grads_flat, scalars, st);
}

// _generator_train_iteration on the batch whose forward (samples [0, batch) of the last dg_generator_lookahead,
// save_first == batch) is still resident: the generator weights have not changed since, so G(coarse) is not recomputed.
extern "C" int dg_generator_step_saved(dg_generator* g, dg_critic* c, const dg_hyper* hp, const float* fine, int batch,
                                       float* g_grads_flat, float* scalars, void* stream) {
  if (c && c->pending_finish) { set_error("dg_generator_step_saved: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(g && c && hp && fine && g_grads_flat && scalars, "dg_generator_step_saved: null argument");
  DG_CHECK(batch >= 1 && batch <= c->maxB, "dg_generator_step_saved: batch %d too large", batch);
  DG_CHECK(g->Hf == c->Hf && g->Cout == c->nc, "dg_generator_step_saved: generator output does not match critic input");
  if (c) DG_TRY(critic_flush_pack(c, (cudaStream_t)stream));  // a pending dg_critic_pack_lazy
  if (!g->packed || !c->packed) { set_error("dg_generator_step_saved: weights not packed"); return DG_ERR_STATE; }
  if (g->saved_batch != batch || g->lookahead < batch) {
    set_error("dg_generator_step_saved: no saved look-ahead forward for %d samples (saved %d, look-ahead %d)", batch, g->saved_batch,
              g->lookahead);
    return DG_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DG_TRY(g->side.join(st));  // a deferred look-ahead chain (dg_generator_lookahead_first) may still be running
  DG_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(float), st));
  return generator_step_body(g, c, hp, fine, batch, g_grads_flat, scalars, st);
}

// Per-batch metric pass, mlflow_tools/mlflow_epoch.py:53-63 (`gen_batch_and_log_metrics`): fake = G(coarse) with the CURRENT
// generator weights, mean C(real) and mean C(fake) with the CURRENT critic weights, MAE = content_loss(real, fake)
// (losses.py:40-55), MSE = content_MSELoss (:58-68), Wass = wass_loss(creal, cfake) = creal - cfake (:8-9).  MS-SSIM
// (losses.py:12-38) comes from the third-party pytorch_msssim package and is not part of this path.
// One generator forward (skipped when `coarse` is NULL: the fake is taken from the look-ahead buffer, valid while the
// generator weights have not changed since that pass), ONE critic forward over [real ; fake] (2B rows), one fused
// MAE + MSE reduction over the two fp32 fields.
extern "C" int dg_metrics(dg_generator* g, dg_critic* c, const float* coarse, int fake_offset, const float* fine, int batch,
                          float* out8, void* stream) {
  if (c && c->pending_finish) { set_error("dg_metrics: dg_critic_step_finish has not been called"); return DG_ERR_STATE; }
  DG_CHECK(g && c && fine && out8, "dg_metrics: null argument");
  DG_CHECK(batch >= 1 && batch <= c->maxB && batch <= g->maxB, "dg_metrics: batch %d too large", batch);
  DG_CHECK(g->Hf == c->Hf && g->Cout == c->nc, "dg_metrics: generator output does not match critic input");
  if (c) DG_TRY(critic_flush_pack(c, (cudaStream_t)stream));  // a pending dg_critic_pack_lazy
  if (!g->packed || !c->packed) { set_error("dg_metrics: weights not packed"); return DG_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  const float* fake = nullptr;
  if (coarse) {
    DG_TRY(g->side.join(st));
    DG_TRY(nchw_to_nhwc(coarse, g->act(g->x0, g->Cin), batch, g->Cin, g->Hc, g->Hc, st));
    DG_TRY(gen_forward_internal(g, batch, 0, st));
    g->saved_batch = 0;
    g->lookahead = 0;
    fake = g->fake;
  } else {
    if (fake_offset < 0 || fake_offset + batch > g->lookahead) {
      set_error("dg_metrics: samples [%d,%d) are not covered by the last look-ahead pass (%d samples)", fake_offset, fake_offset + batch,
                g->lookahead);
      return DG_ERR_STATE;
    }
    if (g->side.dirty && !(fake_offset >= g->ready_lo && fake_offset + batch <= g->ready_hi)) DG_TRY(g->side.join(st));
    fake = g->fake + (size_t)fake_offset * g->Hf * g->Hf * g->Cout;
  }
  DG_TRY(c->side.join(st)); DG_TRY(c->side2.join(st));
  DG_TRY(build_critic_input(fine, fake, 0, nullptr, c->a0, batch, c->nc, c->Hf, c->Hf, 2, st));  // [real ; fake], no interpolates
  DG_TRY(critic_forward_internal(c, 2 * batch, st));
  const long long n = (long long)batch * c->Hf * c->Hf * c->nc;
  DG_TRY(metric_sums(c->a0, c->a0 + n, n, c->scores, batch, c->metric_scratch, out8, st));
  c->saved_batch = 0;
  return 0;
}

// ===========================================================================
// misc exports
// ===========================================================================
extern "C" int dg_abi_version(void) { return DG_ABI_VERSION; }

extern "C" int dg_l1_loss(const float* a, const float* b, int64_t n, float scale, float* loss_out, float* d_a, void* stream) {
  DG_CHECK(a && b && loss_out && n > 0, "dg_l1_loss: bad argument");
  return l1_loss(a, b, n, scale, loss_out, d_a, nullptr, (cudaStream_t)stream);
}

extern "C" int dg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  DG_CHECK(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "dg_adam_step: bad argument");
  return adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

// ---- conv primitives on NCHW fp32 tensors (unit tests) ---------------------
namespace {
struct Scratch {
  std::vector<void*> pool;
  ~Scratch() { for (void* p : pool) cudaFree(p); }
};
}  // namespace

static int prim_setup(Scratch& s, int precision, const float* w, int ci, int co, int mode, float** pk, cudaStream_t st,
                      bf16** pk_u = nullptr, bf16** pk_ig = nullptr) {
  const bool dgrad = mode != 0;
  const size_t elems = dgrad ? packed_w_elems(co, ci) : packed_w_elems(ci, co);
  DG_TRY(dev_alloc(s.pool, (void**)pk, elems * sizeof(float)));
  PackDesc d{}; d.src_off = 0; d.dst_off = 0; d.Ci = ci; d.Co = co; d.mode = mode;
  d.CoP = dgrad ? round_up(ci, 16) : round_up(co, 16);
  PackDesc* dev;
  DG_TRY(dev_alloc(s.pool, (void**)&dev, sizeof(PackDesc)));
  DG_CUDA(cudaMemcpyAsync(dev, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  DG_CUDA(cudaStreamSynchronize(st));
  DG_TRY(pack_weights(w, *pk, nullptr, dev, 1, ci * co * 9, st));
  if (pk_u) {
    *pk_u = nullptr;
    const int oci = dgrad ? co : ci, oco = dgrad ? ci : co;
    if (precision == DG_BF16 && umma_ok(oci, oco)) {
      DG_TRY(dev_alloc(s.pool, (void**)pk_u, (elems + 64) * sizeof(bf16)));
      UmmaPackDesc u{0, oci, round_up(oco, 16)};
      UmmaPackDesc* udev;
      DG_TRY(dev_alloc(s.pool, (void**)&udev, sizeof(u)));
      DG_CUDA(cudaMemcpy(udev, &u, sizeof(u), cudaMemcpyHostToDevice));
      DG_TRY(pack_umma(*pk, *pk_u, udev, 1, (int)elems, st));
      if (pk_ig) {
        DG_TRY(dev_alloc(s.pool, (void**)pk_ig, (elems + 64) * sizeof(bf16)));
        DG_TRY(pack_ig(*pk, *pk_ig, udev, 1, (int)elems, st));
      }
    }
  }
  return 0;
}

extern "C" int dg_conv3x3_fwd(const float* x, const float* w, const float* bias, float* y, int batch, int ci, int co,
                              int hin, int win, int stride, float slope, int precision, void* stream) {
  DG_CHECK(x && w && y && (stride == 1 || stride == 2), "dg_conv3x3_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch s;
  const int bf = precision == DG_BF16;
  const size_t esz = bf ? 2 : 4;
  const int ho = (hin - 1) / stride + 1, wo = (win - 1) / stride + 1;
  float* pk; void *xi, *yo; bf16 *pku, *pkig = nullptr;
  DG_TRY(prim_setup(s, precision, w, ci, co, 0, &pk, st, &pku, &pkig));
  const int xbf = bf && ci > 3;  // 1- to 3-channel inputs stay fp32 as at the networks' boundaries (first-layer kernels)
  DG_TRY(dev_alloc(s.pool, &xi, (size_t)batch * hin * win * ci * (xbf ? 2 : 4)));
  DG_TRY(dev_alloc(s.pool, &yo, (size_t)batch * ho * wo * co * esz));
  DG_TRY(nchw_to_nhwc(x, tv(xi, xbf, ci), batch, ci, hin, win, st));
  ConvOp op;
  op.x = tv(xi, xbf, ci); op.Hin = hin; op.Win = win; op.Ci = ci;
  op.y = tv(yo, bf, co); op.Hout = ho; op.Wout = wo; op.Co = co;
  op.B = batch; op.w = pk; op.bias = bias; op.stride = stride; op.w_umma = pku; op.w_ig = pkig;
  if (slope != 1.f) { op.act = ACT_LRELU; op.slope = slope; }
  DG_TRY(run_conv(op, st));
  DG_TRY(nhwc_to_nchw(tv(yo, bf, co), y, batch, co, ho, wo, st));
  DG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int dg_conv3x3_dgrad(const float* dy, const float* w, float* dx, int batch, int ci, int co, int hin, int win,
                                int stride, int precision, void* stream) {
  DG_CHECK(dy && w && dx && (stride == 1 || stride == 2), "dg_conv3x3_dgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch s;
  const int bf = precision == DG_BF16;
  const size_t esz = bf ? 2 : 4;
  const int ho = (hin - 1) / stride + 1, wo = (win - 1) / stride + 1;
  float* pk; void *dyi, *dxo; bf16 *pku, *pkig = nullptr;
  DG_TRY(prim_setup(s, precision, w, ci, co, stride == 2 ? 2 : 1, &pk, st, &pku, &pkig));
  DG_TRY(dev_alloc(s.pool, &dyi, (size_t)batch * ho * wo * co * esz));
  DG_TRY(dev_alloc(s.pool, &dxo, (size_t)batch * hin * win * ci * esz));
  DG_TRY(nchw_to_nhwc(dy, tv(dyi, bf, co), batch, co, ho, wo, st));
  ConvOp op;
  op.x = tv(dyi, bf, co); op.Hin = ho; op.Win = wo; op.Ci = co;
  op.y = tv(dxo, bf, ci); op.Hout = hin; op.Wout = win; op.Co = ci;
  op.B = batch; op.w = pk; op.transposed = (stride == 2); op.w_umma = pku; op.w_ig = pkig;
  DG_TRY(run_conv(op, st));
  DG_TRY(nhwc_to_nchw(tv(dxo, bf, ci), dx, batch, ci, hin, win, st));
  DG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int dg_conv3x3_wgrad(const float* x, const float* dy, float* dw, float* dbias, int batch, int ci, int co, int hin,
                                int win, int stride, int precision, void* stream) {
  DG_CHECK(x && dy && dw && (stride == 1 || stride == 2), "dg_conv3x3_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch s;
  const int bf = precision == DG_BF16;
  const size_t esz = bf ? 2 : 4;
  const int ho = (hin - 1) / stride + 1, wo = (win - 1) / stride + 1;
  void *xi, *dyi; float *gpk, *gb;
  DG_TRY(dev_alloc(s.pool, &xi, (size_t)batch * hin * win * ci * esz));
  DG_TRY(dev_alloc(s.pool, &dyi, (size_t)batch * ho * wo * co * esz));
  DG_TRY(dev_alloc(s.pool, (void**)&gpk, packed_w_elems(ci, co) * sizeof(float)));
  DG_TRY(dev_alloc(s.pool, (void**)&gb, round_up(co, 16) * sizeof(float)));
  DG_TRY(nchw_to_nhwc(x, tv(xi, bf, ci), batch, ci, hin, win, st));
  DG_TRY(nchw_to_nhwc(dy, tv(dyi, bf, co), batch, co, ho, wo, st));
  WgradOp w;
  w.x = tv(xi, bf, ci); w.Hin = hin; w.Win = win; w.Ci = ci;
  w.dy = tv(dyi, bf, co); w.Hout = ho; w.Wout = wo; w.Co = co; w.B = batch; w.stride = stride;
  w.dw = gpk; w.dbias = dbias ? gb : nullptr;
  DG_TRY(run_wgrad(w, st));
  PackDesc d{}; d.src_off = 0; d.dst_off = 0; d.Ci = ci; d.Co = co; d.CoP = round_up(co, 16); d.mode = 0;
  PackDesc* dev;
  DG_TRY(dev_alloc(s.pool, (void**)&dev, sizeof(PackDesc)));
  DG_CUDA(cudaMemcpyAsync(dev, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  DG_CUDA(cudaStreamSynchronize(st));
  DG_TRY(unpack_wgrads(gpk, dw, dev, 1, ci * co * 9, st));
  if (dbias) DG_CUDA(cudaMemcpyAsync(dbias, gb, sizeof(float) * co, cudaMemcpyDeviceToDevice, st));
  DG_CUDA(cudaStreamSynchronize(st));
  return 0;
}
