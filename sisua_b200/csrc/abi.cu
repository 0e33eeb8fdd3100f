// C ABI + step orchestration (include/sisua_b200.h).  Host code here only sequences kernel launches on
// the caller's stream; all arithmetic is in the .cuh kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sisua_b200.h"
#include "adam.cuh"
#include "dp_adam.cuh"
#include "device_math.cuh"
#include "kernels_mid.cuh"
#include "kernels_unfused.cuh"
#ifdef SISUA_WITH_TC
#include "kernels_tc.cuh"
#include "kernels_tc_enc.cuh"
#include "kernels_tc_selftest.cuh"
#endif

using namespace sisua;

namespace {

struct ParamRef {
  std::string name;
  long long off;
  int rows, cols, ld, kind;   // kind: 0 weight 1 bias 2 gamma 3 beta
  long long size() const { return cols == 0 ? rows : (long long)rows * ld; }
};

struct Layer {          // one Dense -> (BN | bias) -> ReLU unit
  long long w_off;      // weight [H, Kin] offset
  int ldw, Kin;
  long long g_off, b_off;   // gamma / beta (bias in b_off when no BN)
  int bn_index;         // index into bn_moving, -1 without BN
  int stat_index;       // index into the per-step statistics buffer
  float* A;             // pre-activation buffer
  int lda;
  float dropout;
};

}  // namespace

struct sisua_model {
  sisua_step_config cfg;
  int device;
  std::string err;
  std::vector<ParamRef> params;
  long long total_floats = 0;
  // bound buffers
  float *P = nullptr, *Gd = nullptr, *M = nullptr, *V = nullptr, *moving = nullptr;
  // layers
  std::vector<Layer> enc, encl, dec;
  long long lat_w, lat_b, lib_w, lib_b, out_w, out_b, y_w, y_b;
  int n_bn = 0;
  int NO = 0;          // output head columns (nheads * G)
  int Gp = 0;
  int ld0 = 0;         // leading dim of the first-layer pre-activation buffer (64 or 128)
  // workspace
  std::vector<void*> allocs;
  float *A0 = nullptr, *PL = nullptr, *loc = nullptr, *scale = nullptr, *Zs = nullptr, *PLIB = nullptr,
        *lib_loc = nullptr, *lib_scale = nullptr, *lib = nullptr, *dPLIB = nullptr,
        *dLib = nullptr, *lse_part = nullptr, *Trow = nullptr, *dlibsum = nullptr, *D = nullptr, *dD = nullptr, *dHa = nullptr, *dHb = nullptr, *delta1 = nullptr,
        *PY = nullptr, *dPY = nullptr, *OUT = nullptr, *mask_scale = nullptr, *scratch_terms = nullptr;
  // host-buffer entry point (sisua_train_step_host): double-buffered device staging filled on a private copy stream
  struct HostStage {
    bool ready = false;
    cudaStream_t copy = nullptr;
    cudaEvent_t filled[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    float* x[2] = {nullptr, nullptr};
    uint16_t* x16[2] = {nullptr, nullptr};
    int32_t* indptr[2] = {nullptr, nullptr};
    uint16_t* cols[2] = {nullptr, nullptr}; uint16_t* vals[2] = {nullptr, nullptr};
    size_t csr_cap[2] = {0, 0};
    float* y[2] = {nullptr, nullptr}; float* lib[2] = {nullptr, nullptr}; float* eps_z[2] = {nullptr, nullptr};
    float* eps_l[2] = {nullptr, nullptr};
    uint8_t* mask[2] = {nullptr, nullptr};
    float* terms = nullptr; float* loss = nullptr;
    long long calls = 0;
  } hs;
  bool wout_packed = false;    // output-head tiles already re-packed in this step (pack_weights_kernel)
  int n_units = 0;             // hidden units (layers) that own a statistics slot
  double* stats = nullptr;     // [n_units][4][H]: sum, sumsq, sdy, sdyx
  double* sq = nullptr;        // [kMaxSegments]
  long long* d_step = nullptr;
  int* nf_host = nullptr; int* nf_dev = nullptr;      // sticky "a training loss was not finite" word (mapped host memory)
  float* d_lr_t = nullptr;
  SegTable seg;
  uint8_t* packed_wout = nullptr;   // pre-packed fp16 (hi | lo | bias) output-head weight tiles (tcgen05 path)
  int n_gene_tiles = 0;
  uint8_t* packed_w1 = nullptr;     // pre-packed fp16 (hi | lo) first-layer weight k-blocks
  int n_kblocks = 0;
  uint8_t* xt_tiles = nullptr;      // fp16 tiles of dropout(log1p(x)) written by the forward first-layer kernel for its backward
  int xt_kblocks = 0;
  int num_sms = 148;
  // dropout stream of the current training step
  uint64_t drop_seed = 0;
  uint32_t drop_step = 0;
  bool drop_step_on_device = false;   // step < 0: kernels read the device-side optimiser counter (graph replays)
  cudaEvent_t ev_out_grads = nullptr;   // caller-owned: recorded once d loss / d (out.W, out.b) is final
  long long launches = 0;       // kernels launched through this handle (bench.py reports it)
  const int* ridx = nullptr;    // row gather of the current sisua_train_step_gather call (consumed by the tcgen05 kernels)
  bool x_u16 = false;           // ... whose resident count matrix is uint16 (sisua_train_step_gather_u16)
  // data-parallel optimiser step over peer memory (sisua_dp_bind / sisua_adam_step_dp)
  DpArgs dp;
  bool dp_bound = false;
  int dp_grid = 0;
  float* mw_logw = nullptr;     // importance weights of sisua_marginal_llk
  float* gx = nullptr;          // staging for gathered rows: counts (only the un-fused cross-check path needs them dense),
  float *gy = nullptr, *glib = nullptr;   // proteins, library statistics,
  uint8_t* gmask = nullptr;     // label mask
  bool fwd_no_grads = false;       // training-mode forward only (sisua_forward_train_mode): batch statistics and dropout, no gradients
  bool decode_only = false;        // sisua_decode: the latent samples are already in the workspace, start at the decoder
  // one-call options of the posterior fast paths (set by sisua_infer_ex around forward_common)
  const float* x_eval = nullptr;   // counts the likelihood is evaluated on (the encoder still reads x)
  int nozi = 0;                    // likelihood without zero inflation
  float* out_mean_avg = nullptr;   // [B, G] mean over Monte-Carlo samples of the NB mean
  float* logw = nullptr;           // [S*B] importance weights log p(z) - log q(z | x)
  uint64_t infer_seed = 0;      // sisua_set_infer_seed
  long long infer_calls = 0;
  float gscale = 1.0f;          // scale of the fp16 gradient operand tiles (sisua_set_count_bound)
  int force_out_chunks = 0, force_enc_chunks = 0, force_bwd_chunks = 0;   // tests: override the split heuristics (0 = automatic)
  // optional per-section device timing (CUDA events on the caller's stream)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sec_events[8];
  size_t sec_used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

enum Section { SEC_ENC_FIRST = 0, SEC_MID_FWD = 1, SEC_OUT_HEADS = 2, SEC_MID_BWD = 3, SEC_ENC_FIRST_BWD = 4, SEC_ADAM = 5, SEC_COUNT = 6 };

static void sec_begin(sisua_model* h, cudaStream_t st, int id) {
  if (!h->profiling) return;
  if (h->sec_used[id] == h->sec_events[id].size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    h->sec_events[id].push_back({a, b});
  }
  cudaEventRecord(h->sec_events[id][h->sec_used[id]].first, st);
}
static void sec_end(sisua_model* h, cudaStream_t st, int id) {
  if (!h->profiling) return;
  cudaEventRecord(h->sec_events[id][h->sec_used[id]].second, st);
  h->sec_used[id] += 1;
}


#define SET_ERR(h, code, ...)                         \
  do {                                                \
    char _b[512];                                     \
    snprintf(_b, sizeof(_b), __VA_ARGS__);            \
    (h)->err = _b;                                    \
    return code;                                      \
  } while (0)

#define CUDA_OK(h, expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) SET_ERR(h, SISUA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_OK(h, what)                                                                      \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) SET_ERR(h, SISUA_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(_e)); \
  } while (0)


// Launch with Programmatic Dependent Launch: the kernel may be scheduled while its predecessor on the stream is
// still running; every kernel launched this way executes `griddepcontrol.wait` (pdl_wait()) before it touches
// anything a predecessor produced, so only its launch latency and parameter-only prologue overlap.
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  static const bool no_pdl = getenv("SISUA_NO_PDL") != nullptr;     // debugging aid: plain stream-ordered launches
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int n_heads(const sisua_step_config& c) { return (c.x_dist == SISUA_XDIST_ZINBD || c.x_dist == SISUA_XDIST_ZINB) ? 3 : 2; }
static bool tfp_links(const sisua_step_config& c) { return c.x_dist == SISUA_XDIST_ZINB || c.x_dist == SISUA_XDIST_NB; }
static long long align64(long long n) { return (n + 63) / 64 * 64; }

// ---- parameter table: must agree with sisua_b200/config.py:param_layout -------------------------
static void build_layout(sisua_model* h) {
  const sisua_step_config& c = h->cfg;
  const int H = c.n_hidden, G = c.n_genes, Z = c.n_latent, P = c.n_proteins;
  const int Gp = (G + 3) / 4 * 4;
  h->Gp = Gp;
  long long off = 0;
  auto add = [&](const std::string& name, int rows, int cols, int ld, int kind) {
    ParamRef p{name, off, rows, cols, ld, kind};
    h->params.push_back(p);
    off = align64(off + p.size());
    return p.off;
  };
  const bool bn = c.batchnorm != 0;
  const bool scvi = c.model_kind == SISUA_MODEL_SCVI;
  int bn_counter = 0;
  auto add_norm = [&](const std::string& prefix, Layer& L) {
    if (bn) {
      L.g_off = add(prefix + ".gamma", H, 0, H, 2);
      L.b_off = add(prefix + ".beta", H, 0, H, 3);
    } else {
      L.g_off = -1;
      L.b_off = add(prefix + ".b", H, 0, H, 1);
    }
  };
  h->enc.resize(c.n_enc_layers);
  h->encl.resize(scvi ? c.n_encl_layers : 0);
  h->dec.resize(c.n_dec_layers);
  h->enc[0].w_off = add("enc.0.W", H, G, Gp, 0); h->enc[0].ldw = Gp; h->enc[0].Kin = G;
  if (scvi) { h->encl[0].w_off = add("encl.0.W", H, G, Gp, 0); h->encl[0].ldw = Gp; h->encl[0].Kin = G; }
  add_norm("enc.0", h->enc[0]);
  for (int i = 1; i < c.n_enc_layers; ++i) {
    std::string p = "enc." + std::to_string(i);
    h->enc[i].w_off = add(p + ".W", H, H, H, 0); h->enc[i].ldw = H; h->enc[i].Kin = H;
    add_norm(p, h->enc[i]);
  }
  if (scvi) {
    add_norm("encl.0", h->encl[0]);
    for (int i = 1; i < c.n_encl_layers; ++i) {
      std::string p = "encl." + std::to_string(i);
      h->encl[i].w_off = add(p + ".W", H, H, H, 0); h->encl[i].ldw = H; h->encl[i].Kin = H;
      add_norm(p, h->encl[i]);
    }
  }
  const int ZP = c.model_kind == SISUA_MODEL_DCA ? Z : 2 * Z;
  h->lat_w = add("lat.W", ZP, H, H, 0);
  h->lat_b = add("lat.b", ZP, 0, ZP, 1);
  h->lib_w = h->lib_b = -1;
  if (scvi) { h->lib_w = add("lib.W", 2, H, H, 0); h->lib_b = add("lib.b", 2, 0, 2, 1); }
  h->dec[0].w_off = add("dec.0.W", H, Z, Z, 0); h->dec[0].ldw = Z; h->dec[0].Kin = Z;
  add_norm("dec.0", h->dec[0]);
  for (int i = 1; i < c.n_dec_layers; ++i) {
    std::string p = "dec." + std::to_string(i);
    h->dec[i].w_off = add(p + ".W", H, H, H, 0); h->dec[i].ldw = H; h->dec[i].Kin = H;
    add_norm(p, h->dec[i]);
  }
  h->y_w = h->y_b = -1;
  if (P > 0) { h->y_w = add("y.W", 2 * P, H, H, 0); h->y_b = add("y.b", 2 * P, 0, 2 * P, 1); }
  const int nheads = n_heads(c);
  h->NO = nheads * G;
  h->out_w = add("out.W", h->NO, H, H, 0);     // last: one contiguous early-ready all-reduce bucket
  h->out_b = add("out.b", h->NO, 0, h->NO, 1);
  h->total_floats = off;
  // BN indices follow config.py:bn_layer_names (enc, encl, dec)
  for (auto& L : h->enc) L.bn_index = bn ? bn_counter++ : -1;
  for (auto& L : h->encl) L.bn_index = bn ? bn_counter++ : -1;
  for (auto& L : h->dec) L.bn_index = bn ? bn_counter++ : -1;
  h->n_bn = bn_counter;
  int sc = 0;
  for (auto& L : h->enc) L.stat_index = sc++;
  for (auto& L : h->encl) L.stat_index = sc++;
  for (auto& L : h->dec) L.stat_index = sc++;
  h->n_units = sc;
  for (auto& L : h->enc) L.dropout = c.enc_dropout;
  for (auto& L : h->encl) L.dropout = c.encl_dropout;
  for (auto& L : h->dec) L.dropout = c.dec_dropout;
  h->seg.n = (int)h->params.size();
  for (int i = 0; i < h->seg.n; ++i) { h->seg.off[i] = h->params[i].off; h->seg.size[i] = h->params[i].size(); }
}

template <typename T>
static int ws_alloc(sisua_model* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) SET_ERR(h, SISUA_ERR_CUDA, "workspace cudaMalloc(%zu B) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = (T*)q;
  return SISUA_OK;
}

static NoiseSpec make_noise(const sisua_model* h) {     // in-kernel reparameterisation noise of the current step / call
  NoiseSpec n;
  n.seed_lo = (uint32_t)(h->drop_seed & 0xffffffffu); n.seed_hi = (uint32_t)(h->drop_seed >> 32);
  n.step = h->drop_step;
  n.step_ptr = h->drop_step_on_device ? h->d_step : nullptr;
  return n;
}

static DropSpec make_drop(sisua_model* h, float rate, uint32_t stream, bool training) {
  DropSpec d;
  memset(&d, 0, sizeof(d));
  if (training && rate > 0.f) {
    d.rate = rate; d.scale = 1.0f / (1.0f - rate);
    d.seed_lo = (uint32_t)(h->drop_seed & 0xffffffffu); d.seed_hi = (uint32_t)(h->drop_seed >> 32);
    d.step = h->drop_step; d.stream = stream;
    d.step_ptr = h->drop_step_on_device ? h->d_step : nullptr;
  }
  return d;
}

#ifdef SISUA_WITH_TC
static bool tc_heads_enabled(const sisua_model* h) {
  // the literal reading of Q2 (activations applied again on scVI's positive parameters) keeps the un-fused row kernel
  return h->cfg.gemm_mode != SISUA_GEMM_FP32_UNFUSED &&
         !(h->cfg.model_kind == SISUA_MODEL_SCVI && h->cfg.scvi_reapply_act);
}

constexpr int kLseChunks = 16;   // gene chunks of the scVI logsumexp pass (x 4 slices = partials per row)

template <int NH, bool VEC>
static int tc_set_attr_scvi(sisua_model* h) {
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, false, true, tc::LINK_SOFTPLUS, tc::MODE_SCVI_LSE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, false)));
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, false, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_EVAL>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, false)));
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, false, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_SUMS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, false)));
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, true, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_TRAIN>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, true)));
  return SISUA_OK;
}

template <int NH, bool TRAIN, bool VEC>
static int tc_set_attr(sisua_model* h) {
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  tc::OutSmem::total(NH, TRAIN)));
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  tc::OutSmem::total(NH, TRAIN)));
  CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_TFP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  tc::OutSmem::total(NH, TRAIN)));
  if (TRAIN && VEC) {      // the uint16-count instances (sisua_train_step_gather_u16)
    CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, true, true, tc::LINK_SOFTPLUS, tc::MODE_PLAIN, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, true)));
    CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, true, true, tc::LINK_GENERIC, tc::MODE_PLAIN, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, true)));
    CUDA_OK(h, cudaFuncSetAttribute(tc::out_heads_kernel<NH, true, true, tc::LINK_TFP, tc::MODE_PLAIN, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, tc::OutSmem::total(NH, true)));
  }
  return SISUA_OK;
}

template <int N0>
static int tc_enc_attr(sisua_model* h) {
  CUDA_OK(h, cudaFuncSetAttribute(tc::enc_first_fwd_kernel<N0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EncFwdSmem<N0, true>::total));
  CUDA_OK(h, cudaFuncSetAttribute(tc::enc_first_fwd_kernel<N0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EncFwdSmem<N0, false>::total));
  CUDA_OK(h, cudaFuncSetAttribute(tc::enc_first_fwd_kernel<N0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EncFwdSmem<N0, true>::total));
  CUDA_OK(h, cudaFuncSetAttribute(tc::enc_first_bwd_kernel<N0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EncBwdSmem<N0>::total));
  CUDA_OK(h, cudaFuncSetAttribute(tc::enc_first_bwd_kernel<N0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::EncBwdSmem<N0>::total));
  return SISUA_OK;
}

// ---- launch geometry of the three tcgen05 kernels (also reported by sisua_debug_geometry so the tests can assert the
// walk depth a batch size exercises) ----
struct TcGeometry {
  int enc_cell_tiles, enc_chunks, enc_kblocks_per_chunk;
  int out_cell_tiles, out_chunks, out_tiles_per_chunk;
  int bwd_gene_tiles, bwd_chunks, bwd_cell_tiles_per_chunk;
};
static TcGeometry tc_geometry(const sisua_model* h, int B, int R) {
  TcGeometry g;
  g.enc_cell_tiles = (B + 127) / 128;
  int chunks = std::max(1, std::min(h->n_kblocks, h->num_sms / g.enc_cell_tiles));
  if (h->force_enc_chunks > 0) chunks = std::min(h->n_kblocks, h->force_enc_chunks);
  g.enc_kblocks_per_chunk = (h->n_kblocks + chunks - 1) / chunks;
  g.enc_chunks = (h->n_kblocks + g.enc_kblocks_per_chunk - 1) / g.enc_kblocks_per_chunk;
  g.out_cell_tiles = (R + tc::kCellTile - 1) / tc::kCellTile;
  const int ngt = std::max(1, h->n_gene_tiles);
  chunks = std::max(1, std::min(ngt, (h->num_sms + g.out_cell_tiles / 2) / g.out_cell_tiles));
  if (h->force_out_chunks > 0) chunks = std::min(ngt, h->force_out_chunks);
  g.out_tiles_per_chunk = (ngt + chunks - 1) / chunks;
  g.out_chunks = (ngt + g.out_tiles_per_chunk - 1) / g.out_tiles_per_chunk;
  g.bwd_gene_tiles = (h->cfg.n_genes + 127) / 128;
  chunks = std::max(1, std::min(g.enc_cell_tiles, h->num_sms / g.bwd_gene_tiles));
  if (h->force_bwd_chunks > 0) chunks = std::min(g.enc_cell_tiles, h->force_bwd_chunks);
  g.bwd_cell_tiles_per_chunk = (g.enc_cell_tiles + chunks - 1) / chunks;
  g.bwd_chunks = (g.enc_cell_tiles + g.bwd_cell_tiles_per_chunk - 1) / g.bwd_cell_tiles_per_chunk;
  return g;
}

static int tc_encoder_first(sisua_model* h, cudaStream_t st, const float* x, int B, int N0, bool training, bool grads) {
  const sisua_step_config& c = h->cfg;
  {
    // both packed operands in one launch (the output-head tiles are consumed later in the same step)
    const bool heads = tc_heads_enabled(h);
    const int nh = n_heads(c);
    ++h->launches;
    tc::pack_weights_kernel<<<h->n_kblocks + (heads ? h->n_gene_tiles : 0), 256, 0, st>>>(
        h->P + h->enc[0].w_off, h->Gp, h->packed_w1, c.n_genes, N0, h->n_kblocks, heads ? h->P + h->out_w : nullptr,
        heads ? h->P + h->out_b : nullptr, h->packed_wout, nh);
    h->wout_packed = heads;
  }
  tc::EncFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.ridx = h->ridx; a.packed = h->packed_w1; a.A0 = h->A0; a.B = B; a.G = c.n_genes; a.ld0 = h->ld0; a.n_kblocks = h->n_kblocks;
  a.log_norm = c.log_norm; a.drop = make_drop(h, c.input_dropout, 0u, training);
  a.xt = grads ? h->xt_tiles : nullptr; a.xt_kblocks = h->xt_kblocks;
  const TcGeometry geo = tc_geometry(h, B, B);
  const int cell_tiles = geo.enc_cell_tiles, chunks = geo.enc_chunks;
  a.kblocks_per_chunk = geo.enc_kblocks_per_chunk;
  a.atomic_out = chunks > 1 ? 1 : 0;
  if (a.atomic_out) CUDA_OK(h, cudaMemsetAsync(h->A0, 0, (size_t)B * h->ld0 * sizeof(float), st));
  const bool vec = (c.n_genes % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  dim3 grid(cell_tiles, chunks);
  ++h->launches;
  if (h->x_u16) {      // (the entry point has checked the 16-byte geometry)
    if (N0 == 64) tc::enc_first_fwd_kernel<64, true, true><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<64, true>::total, st>>>(a);
    else tc::enc_first_fwd_kernel<128, true, true><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<128, true>::total, st>>>(a);
  } else if (N0 == 64) {
    if (vec) tc::enc_first_fwd_kernel<64, true><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<64, true>::total, st>>>(a);
    else tc::enc_first_fwd_kernel<64, false><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<64, false>::total, st>>>(a);
  } else {
    if (vec) tc::enc_first_fwd_kernel<128, true><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<128, true>::total, st>>>(a);
    else tc::enc_first_fwd_kernel<128, false><<<grid, tc::kEncFwdThreads, tc::EncFwdSmem<128, false>::total, st>>>(a);
  }
  LAUNCH_OK(h, "enc_first_fwd_kernel (tcgen05)");
  return SISUA_OK;
}

static int tc_encoder_first_bwd(sisua_model* h, cudaStream_t st, const float* x, int B, int N0) {
  const sisua_step_config& c = h->cfg;
  tc::EncBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.xt = h->xt_tiles; a.xt_kblocks = h->xt_kblocks;
  a.delta = h->delta1; a.dW = h->Gd + h->enc[0].w_off; a.B = B; a.G = c.n_genes; a.Gp = h->Gp; a.ld0 = h->ld0;
  a.n_cell_tiles = (B + 127) / 128; a.log_norm = c.log_norm;
  a.in_scale = (float)B; a.out_scale = 1.0f / (float)B;
  a.drop = make_drop(h, c.input_dropout, 0u, true);
  const TcGeometry geo = tc_geometry(h, B, B);
  const int gene_tiles = geo.bwd_gene_tiles, chunks = geo.bwd_chunks;
  a.tiles_per_chunk = geo.bwd_cell_tiles_per_chunk;
  const bool vec = (c.n_genes % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  dim3 grid(gene_tiles, chunks);
  ++h->launches;
  if (N0 == 64) {
    if (vec) tc::enc_first_bwd_kernel<64, true><<<grid, tc::kEncThreads, tc::EncBwdSmem<64>::total, st>>>(a);
    else tc::enc_first_bwd_kernel<64, false><<<grid, tc::kEncThreads, tc::EncBwdSmem<64>::total, st>>>(a);
  } else {
    if (vec) tc::enc_first_bwd_kernel<128, true><<<grid, tc::kEncThreads, tc::EncBwdSmem<128>::total, st>>>(a);
    else tc::enc_first_bwd_kernel<128, false><<<grid, tc::kEncThreads, tc::EncBwdSmem<128>::total, st>>>(a);
  }
  LAUNCH_OK(h, "enc_first_bwd_kernel (tcgen05)");
  return SISUA_OK;
}

static int tc_create(sisua_model* h) {
  {
    const int n0 = h->cfg.model_kind == SISUA_MODEL_SCVI ? 128 : 64;
    h->n_kblocks = (h->cfg.n_genes + 63) / 64;
    int rc0 = ws_alloc(h, &h->packed_w1, (size_t)h->n_kblocks * tc::w1_block_bytes(n0));
    if (rc0 != SISUA_OK) return rc0;
    h->xt_kblocks = (h->n_kblocks + 1) / 2 * 2;
    const size_t xt_bytes = (size_t)((h->cfg.max_batch + 127) / 128) * h->xt_kblocks * tc::kXtTile;
    rc0 = ws_alloc(h, &h->xt_tiles, xt_bytes);
    if (rc0 != SISUA_OK) return rc0;
    CUDA_OK(h, cudaMemset(h->xt_tiles, 0, xt_bytes));   // the padding k-block of an odd gene count is never written
    rc0 = n0 == 64 ? tc_enc_attr<64>(h) : tc_enc_attr<128>(h);
    if (rc0 != SISUA_OK) return rc0;
  }
  if (!tc_heads_enabled(h)) return SISUA_OK;
  const int nh = n_heads(h->cfg);
  h->n_gene_tiles = (h->cfg.n_genes + tc::kGeneTile - 1) / tc::kGeneTile;
  int rc = ws_alloc(h, &h->packed_wout, (size_t)h->n_gene_tiles * tc::packed_tile_stride(nh));
  if (rc != SISUA_OK) return rc;
  if (h->cfg.model_kind == SISUA_MODEL_SCVI) {
    const size_t R = h->cfg.max_batch;
    if ((rc = ws_alloc(h, &h->lse_part, R * kLseChunks * 4 * 2)) || (rc = ws_alloc(h, &h->Trow, R)) ||
        (rc = ws_alloc(h, &h->dlibsum, R))) return rc;
    if (nh == 3) { if ((rc = tc_set_attr_scvi<3, true>(h)) || (rc = tc_set_attr_scvi<3, false>(h))) return rc; }
    else { if ((rc = tc_set_attr_scvi<2, true>(h)) || (rc = tc_set_attr_scvi<2, false>(h))) return rc; }
    return SISUA_OK;
  }
#define TC_ATTR(NH) \
  if ((rc = tc_set_attr<NH, true, true>(h)) || (rc = tc_set_attr<NH, true, false>(h)) || \
      (rc = tc_set_attr<NH, false, true>(h)) || (rc = tc_set_attr<NH, false, false>(h))) return rc
  if (nh == 3) { TC_ATTR(3); } else { TC_ATTR(2); }
#undef TC_ATTR
  return SISUA_OK;
}

// scVI heads on the same fused kernel (gene softmax needs row sums before the Jacobian can be applied):
//   logsumexp pass -> [inference] evaluation pass
//                  -> [training] row-sum pass (llk, T = sum s t, d llk / d library) -> gradient pass
template <int NH, bool VEC>
static int tc_scvi_passes(sisua_model* h, cudaStream_t st, bool training, tc::OutHeadsArgs a, dim3 grid) {
  dim3 grid_lse(grid.x, std::min<int>(grid.y, kLseChunks));
  tc::OutHeadsArgs l = a;
  l.tiles_per_chunk = (a.n_tiles + (int)grid_lse.y - 1) / (int)grid_lse.y;
  grid_lse.y = (a.n_tiles + l.tiles_per_chunk - 1) / l.tiles_per_chunk;
  l.n_lse_parts = a.n_lse_parts = (int)grid_lse.y * 4;
  ++h->launches;
  tc::out_heads_kernel<NH, false, true, tc::LINK_SOFTPLUS, tc::MODE_SCVI_LSE><<<grid_lse, tc::kOutThreads, tc::OutSmem::total(NH, false), st>>>(l);
  LAUNCH_OK(h, "out_heads_kernel (scVI logsumexp)");
  ++h->launches;
  if (!training) {
    tc::out_heads_kernel<NH, false, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_EVAL><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, false), st>>>(a);
    LAUNCH_OK(h, "out_heads_kernel (scVI evaluation)");
    return SISUA_OK;
  }
  CUDA_OK(h, cudaMemsetAsync(a.Trow, 0, (size_t)a.R * sizeof(float), st));
  CUDA_OK(h, cudaMemsetAsync(a.dlibsum, 0, (size_t)a.R * sizeof(float), st));
  tc::out_heads_kernel<NH, false, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_SUMS><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, false), st>>>(a);
  LAUNCH_OK(h, "out_heads_kernel (scVI row sums)");
  ++h->launches;
  tc::out_heads_kernel<NH, true, VEC, tc::LINK_SOFTPLUS, tc::MODE_SCVI_TRAIN><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, true), st>>>(a);
  LAUNCH_OK(h, "out_heads_kernel (scVI gradients)");
  return SISUA_OK;
}

// fused output heads + count likelihood (+ backward of the heads when training)
static int tc_output_heads(sisua_model* h, cudaStream_t st, bool training, const float* x, int B, int S, float* llk_x,
                           float* out_mean, float* out_disp, float* out_pi, const NormSpec* fused_norm) {
  const sisua_step_config& c = h->cfg;
  const int nh = n_heads(c);
  const int R = S * B, G = c.n_genes;
  if (!h->wout_packed) {      // normally packed together with the first-layer operand at the start of the step
    ++h->launches;
    tc::pack_wout_kernel<<<h->n_gene_tiles, 256, 0, st>>>(h->P + h->out_w, h->P + h->out_b, h->packed_wout, G, nh, h->n_gene_tiles);
  }
  h->wout_packed = false;
  CUDA_OK(h, cudaMemsetAsync(llk_x, 0, (size_t)R * sizeof(float), st));
  tc::OutHeadsArgs a;
  memset(&a, 0, sizeof(a));
  a.D = h->D; a.ldD = kH; a.x = x; a.ridx = h->ridx; a.packed = h->packed_wout; a.llk_x = llk_x;
  if (fused_norm) { a.D = h->dec.back().A; a.ldD = h->dec.back().lda; a.fuse_norm = 1; a.ns = *fused_norm; }
  a.out_mean = out_mean; a.out_disp = out_disp; a.out_pi = out_pi;
  if (!training) { a.nozi = h->nozi; a.out_mean_avg = h->out_mean_avg; a.inv_S = 1.0f / (float)S; }
  a.dD = h->dD; a.dW = h->Gd ? h->Gd + h->out_w : nullptr; a.db = h->Gd ? h->Gd + h->out_b : nullptr;
  a.R = R; a.B = B; a.G = G; a.n_tiles = h->n_gene_tiles;
  a.mean_act = c.mean_act; a.disp_act = c.disp_act; a.upstream = -1.0f / (float)R;
  a.gscale = h->gscale; a.inv_gscale = 1.0f / h->gscale;
  const TcGeometry geo = tc_geometry(h, B, R);
  const int cell_tiles = geo.out_cell_tiles, chunks = geo.out_chunks;
  a.tiles_per_chunk = geo.out_tiles_per_chunk;
  dim3 grid(cell_tiles, chunks);
  const bool vec = (G % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (c.model_kind == SISUA_MODEL_SCVI) {
    a.lse_part = reinterpret_cast<float2*>(h->lse_part); a.lib = h->lib; a.clip_library = c.clip_library;
    a.Trow = h->Trow; a.dlibsum = h->dlibsum; a.dLib = h->dLib;
    if (nh == 3) return vec ? tc_scvi_passes<3, true>(h, st, training, a, grid) : tc_scvi_passes<3, false>(h, st, training, a, grid);
    return vec ? tc_scvi_passes<2, true>(h, st, training, a, grid) : tc_scvi_passes<2, false>(h, st, training, a, grid);
  }
  ++h->launches;
  const int link = tfp_links(c) ? tc::LINK_TFP : ((c.mean_act == SISUA_ACT_SOFTPLUS && c.disp_act == SISUA_ACT_SOFTPLUS1) ? tc::LINK_SOFTPLUS : tc::LINK_GENERIC);
#define TC_LAUNCH(NH, TRAIN, VEC)                                                                                   \
  do {                                                                                                              \
    if (link == tc::LINK_SOFTPLUS) tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_SOFTPLUS><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, TRAIN), st>>>(a);  \
    else if (link == tc::LINK_TFP) tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_TFP><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, TRAIN), st>>>(a);  \
    else tc::out_heads_kernel<NH, TRAIN, VEC, tc::LINK_GENERIC><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, TRAIN), st>>>(a);      \
  } while (0)
#define TC_LAUNCH_U16(NH)                                                                                            \
  do {                                                                                                              \
    if (link == tc::LINK_SOFTPLUS) tc::out_heads_kernel<NH, true, true, tc::LINK_SOFTPLUS, tc::MODE_PLAIN, true><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, true), st>>>(a);  \
    else if (link == tc::LINK_TFP) tc::out_heads_kernel<NH, true, true, tc::LINK_TFP, tc::MODE_PLAIN, true><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, true), st>>>(a);  \
    else tc::out_heads_kernel<NH, true, true, tc::LINK_GENERIC, tc::MODE_PLAIN, true><<<grid, tc::kOutThreads, tc::OutSmem::total(NH, true), st>>>(a);      \
  } while (0)
  if (h->x_u16) {      // training step on uint16 resident counts (the entry point has checked the geometry)
    if (nh == 3) TC_LAUNCH_U16(3); else TC_LAUNCH_U16(2);
  } else if (nh == 3) {
    if (training) { if (vec) TC_LAUNCH(3, true, true); else TC_LAUNCH(3, true, false); }
    else { if (vec) TC_LAUNCH(3, false, true); else TC_LAUNCH(3, false, false); }
  } else {
    if (training) { if (vec) TC_LAUNCH(2, true, true); else TC_LAUNCH(2, true, false); }
    else { if (vec) TC_LAUNCH(2, false, true); else TC_LAUNCH(2, false, false); }
  }
#undef TC_LAUNCH
#undef TC_LAUNCH_U16
  LAUNCH_OK(h, "out_heads_kernel (tcgen05)");
  return SISUA_OK;
}
#endif  // SISUA_WITH_TC

extern "C" const char* sisua_version(void) { return "sisua_b200 0.1 (sm_100a)"; }

extern "C" const char* sisua_last_error(sisua_handle h) { return h ? h->err.c_str() : "null handle"; }

static std::string g_create_err;

extern "C" int sisua_create(const sisua_step_config* cfg, int device, sisua_handle* out) {
  if (!cfg || !out) return SISUA_ERR_INVALID;
  *out = nullptr;
  sisua_model* h = new sisua_model();
  h->cfg = *cfg;
  h->device = device;
  *out = h;   // returned even on failure so the caller can read last_error, then destroy
  const sisua_step_config& c = h->cfg;
  if (c.n_hidden != kH) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "n_hidden=%d: kernels are built for 64 hidden units", c.n_hidden);
  if (c.n_latent < 1 || c.n_latent > 32) SET_ERR(h, SISUA_ERR_INVALID, "n_latent must be in [1,32]");
  if (c.n_proteins < 0 || c.n_proteins > 32) SET_ERR(h, SISUA_ERR_INVALID, "n_proteins must be in [0,32]");
  if (c.n_genes < 1) SET_ERR(h, SISUA_ERR_INVALID, "n_genes must be positive");
  if (c.n_enc_layers < 1 || c.n_enc_layers > 4 || c.n_dec_layers < 1 || c.n_dec_layers > 4)
    SET_ERR(h, SISUA_ERR_INVALID, "hidden layer counts must be in [1,4]");
  if (c.model_kind == SISUA_MODEL_SCVI && (c.n_encl_layers < 1 || c.n_encl_layers > 4))
    SET_ERR(h, SISUA_ERR_INVALID, "scVI needs 1..4 library-encoder layers");
  if (c.model_kind == SISUA_MODEL_SISUA && c.n_proteins < 1) SET_ERR(h, SISUA_ERR_INVALID, "SISUA needs proteins");
  if (c.model_kind != SISUA_MODEL_SISUA && c.n_proteins != 0) SET_ERR(h, SISUA_ERR_INVALID, "proteins only with SISUA");
  if (c.max_batch < 1) SET_ERR(h, SISUA_ERR_INVALID, "max_batch must be positive");
  for (float r : {c.input_dropout, c.enc_dropout, c.dec_dropout, c.encl_dropout})
    if (r < 0.f || r >= 1.f) SET_ERR(h, SISUA_ERR_INVALID, "dropout rates must be in [0, 1)");
  if (c.gemm_mode != SISUA_GEMM_FP32_UNFUSED && c.gemm_mode != SISUA_GEMM_TC_3XFP16) SET_ERR(h, SISUA_ERR_INVALID, "unknown gemm_mode %d", c.gemm_mode);
#ifndef SISUA_WITH_TC
  if (c.gemm_mode != SISUA_GEMM_FP32_UNFUSED) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "library built without the tcgen05 kernels");
#endif
  CUDA_OK(h, cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(h, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  build_layout(h);
  const size_t R = c.max_batch;
  const int H = kH, Z = c.n_latent, P = c.n_proteins;
  const bool scvi = c.model_kind == SISUA_MODEL_SCVI;
  h->ld0 = scvi ? 2 * H : H;
  int rc;
#define WS(ptr, count) if ((rc = ws_alloc(h, &(ptr), (count))) != SISUA_OK) return rc
  WS(h->A0, R * h->ld0);
  h->enc[0].A = h->A0; h->enc[0].lda = h->ld0;
  for (int i = 1; i < c.n_enc_layers; ++i) { WS(h->enc[i].A, R * H); h->enc[i].lda = H; }
  if (scvi) {
    h->encl[0].A = h->A0 + H; h->encl[0].lda = h->ld0;
    for (int i = 1; i < c.n_encl_layers; ++i) { WS(h->encl[i].A, R * H); h->encl[i].lda = H; }
  }
  for (int i = 0; i < c.n_dec_layers; ++i) { WS(h->dec[i].A, R * H); h->dec[i].lda = H; }
  WS(h->PL, R * 2 * Z); WS(h->loc, R * Z); WS(h->scale, R * Z); WS(h->Zs, R * Z);
  if (scvi) {
    WS(h->PLIB, R * 2); WS(h->lib_loc, R); WS(h->lib_scale, R); WS(h->lib, R); WS(h->dPLIB, R * 2); WS(h->dLib, R);
  }
  WS(h->D, R * H); WS(h->dD, R * H); WS(h->dHa, R * H); WS(h->dHb, R * H); WS(h->delta1, R * h->ld0);
  if (P > 0) { WS(h->PY, R * 2 * P); WS(h->dPY, R * 2 * P); }
  if (c.gemm_mode == SISUA_GEMM_FP32_UNFUSED || (scvi && c.scvi_reapply_act)) WS(h->OUT, R * (size_t)h->NO);
  WS(h->mask_scale, 1);
  WS(h->scratch_terms, 5 * R);
  WS(h->stats, (size_t)h->n_units * 4 * H);
  WS(h->sq, kMaxSegments);
  WS(h->d_step, 1);
  WS(h->d_lr_t, 1);
#undef WS
  CUDA_OK(h, cudaMemset(h->d_step, 0, sizeof(long long)));
  CUDA_OK(h, cudaHostAlloc((void**)&h->nf_host, sizeof(int), cudaHostAllocMapped));
  *h->nf_host = 0;
  CUDA_OK(h, cudaHostGetDevicePointer((void**)&h->nf_dev, h->nf_host, 0));
  CUDA_OK(h, cudaFuncSetAttribute(dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseBwdSmem));
  CUDA_OK(h, cudaFuncSetAttribute(dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseFwdSmem));
  CUDA_OK(h, cudaFuncSetAttribute(latent_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatentFwdSmem));
  CUDA_OK(h, cudaFuncSetAttribute(latent_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLatentBwdSmem));
  CUDA_OK(h, cudaFuncSetAttribute(count_row_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_OK(h, cudaFuncSetAttribute(count_row_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (scvi && (c.gemm_mode == SISUA_GEMM_FP32_UNFUSED || c.scvi_reapply_act) && c.n_genes * sizeof(float) > 200 * 1024)
    SET_ERR(h, SISUA_ERR_UNSUPPORTED, "un-fused scVI row kernel caches one softmax row in shared memory: n_genes <= 51200");
#ifdef SISUA_WITH_TC
  if (c.gemm_mode != SISUA_GEMM_FP32_UNFUSED) {
    rc = tc_create(h);
    if (rc != SISUA_OK) return rc;
  }
#endif
  return SISUA_OK;
}

extern "C" int sisua_destroy(sisua_handle h) {
  if (!h) return SISUA_ERR_INVALID;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->nf_host) cudaFreeHost(h->nf_host);
  for (auto& v : h->sec_events)
    for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (h->hs.copy) {
    cudaStreamSynchronize(h->hs.copy);
    for (int s = 0; s < 2; ++s) {
      if (h->hs.filled[s]) cudaEventDestroy(h->hs.filled[s]);
      if (h->hs.consumed[s]) cudaEventDestroy(h->hs.consumed[s]);
      cudaFree(h->hs.indptr[s]); cudaFree(h->hs.cols[s]); cudaFree(h->hs.vals[s]);
    }
    cudaStreamDestroy(h->hs.copy);
  }
  delete h;
  return SISUA_OK;
}

extern "C" int sisua_param_layout(sisua_handle h, sisua_param_desc* out, int* n, int64_t* total_floats) {
  if (!h || !n) return SISUA_ERR_INVALID;
  int cap = *n;
  *n = (int)h->params.size();
  if (total_floats) *total_floats = h->total_floats;
  if (!out) return SISUA_OK;
  if (cap < (int)h->params.size()) SET_ERR(h, SISUA_ERR_INVALID, "param_layout: capacity %d < %zu", cap, h->params.size());
  for (size_t i = 0; i < h->params.size(); ++i) {
    const ParamRef& p = h->params[i];
    memset(&out[i], 0, sizeof(out[i]));
    strncpy(out[i].name, p.name.c_str(), sizeof(out[i].name) - 1);
    out[i].offset = p.off; out[i].rows = p.rows; out[i].cols = p.cols; out[i].ld = p.ld; out[i].kind = p.kind;
  }
  return SISUA_OK;
}

extern "C" int sisua_bind_buffers(sisua_handle h, float* params, float* grads, float* adam_m, float* adam_v,
                                  float* bn_moving) {
  if (!h) return SISUA_ERR_INVALID;
  if (!params) SET_ERR(h, SISUA_ERR_INVALID, "bind_buffers: params is null");
  if (h->cfg.batchnorm && !bn_moving) SET_ERR(h, SISUA_ERR_INVALID, "bind_buffers: bn_moving is null but batchnorm is on");
  h->P = params; h->Gd = grads; h->M = adam_m; h->V = adam_v; h->moving = bn_moving;
  return SISUA_OK;
}

extern "C" const float* sisua_debug_buffer(sisua_handle h, const char* name) {
  if (!h || !name) return nullptr;
  std::string n(name);
  if (n == "d") return h->D;
  if (n == "delta1") return h->delta1;
  if (n == "out") return h->OUT;
  if (n == "dD") return h->dD;
  if (n == "a0") return h->A0;
  if (n == "z") return h->Zs;
  return nullptr;
}

extern "C" int sisua_debug_copy(sisua_handle h, const char* name, float* dst, int64_t n_floats, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  const float* src = sisua_debug_buffer(h, name);
  if (!src) SET_ERR(h, SISUA_ERR_INVALID, "debug_copy: unknown buffer '%s'", name ? name : "");
  CUDA_OK(h, cudaMemcpyAsync(dst, src, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SISUA_OK;
}

// Launch geometry of the tcgen05 kernels for a train step of B cells (tests assert the walk depths they claim to cover).
// out[9] = first layer (cell tiles, k-chunks, k-blocks per chunk) | output heads (cell tiles, gene chunks, gene tiles per
// chunk) | first-layer weight gradient (gene tiles, cell chunks, cell tiles per chunk).  Zeros without the tcgen05 path.
#ifdef SISUA_OUT_TRACE
// development aid, only in -DSISUA_OUT_TRACE builds (not declared in include/sisua_b200.h): copies the clock stamps of CTA 0
extern "C" int sisua_debug_out_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, sisua::tc::g_out_trace, sizeof(long long) * 4 * 64 * 8) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int sisua_debug_geometry(sisua_handle h, int B, int32_t* out) {
  if (!h || !out || B < 1) return SISUA_ERR_INVALID;
  for (int i = 0; i < 9; ++i) out[i] = 0;
#ifdef SISUA_WITH_TC
  if (h->cfg.gemm_mode != SISUA_GEMM_FP32_UNFUSED) {
    const TcGeometry g = tc_geometry(h, B, B);
    const int v[9] = {g.enc_cell_tiles, g.enc_chunks, g.enc_kblocks_per_chunk, g.out_cell_tiles, g.out_chunks, g.out_tiles_per_chunk,
                      g.bwd_gene_tiles, g.bwd_chunks, g.bwd_cell_tiles_per_chunk};
    for (int i = 0; i < 9; ++i) out[i] = v[i];
    if (!tc_heads_enabled(h)) out[3] = out[4] = out[5] = 0;
  }
#endif
  return SISUA_OK;
}

// Tests only: force the number of chunks the three tcgen05 kernels split their walk into (0 = automatic), so that a
// small batch can exercise the single-chunk, deep-pipeline geometry a full-size minibatch gets.
extern "C" int sisua_debug_force_chunks(sisua_handle h, int out_chunks, int enc_chunks, int bwd_chunks) {
  if (!h || out_chunks < 0 || enc_chunks < 0 || bwd_chunks < 0) return SISUA_ERR_INVALID;
  h->force_out_chunks = out_chunks; h->force_enc_chunks = enc_chunks; h->force_bwd_chunks = bwd_chunks;
  return SISUA_OK;
}

// ---- helpers ------------------------------------------------------------------------------------
static NormSpec make_norm(sisua_model* h, const Layer& L, bool training, int rows) {
  NormSpec ns;
  memset(&ns, 0, sizeof(ns));
  ns.eps = h->cfg.bn_eps;
  ns.drop = make_drop(h, L.dropout, 1u + (uint32_t)L.stat_index, training);
  if (L.bn_index >= 0) {
    ns.gamma = h->P + L.g_off; ns.beta = h->P + L.b_off;
    if (training) {
      ns.mode = NORM_BN_BATCH;
      ns.sum = h->stats + (size_t)L.stat_index * 4 * kH;
      ns.sumsq = ns.sum + kH;
      ns.inv_count = 1.0f / (float)rows;
    } else {
      ns.mode = NORM_BN_MOVING;
      ns.moving = h->moving + (size_t)L.bn_index * 2 * kH;
    }
  } else {
    ns.mode = NORM_BIAS;
    ns.beta = h->P + L.b_off;
  }
  return ns;
}
static NormSpec raw_norm() { NormSpec ns; memset(&ns, 0, sizeof(ns)); ns.mode = NORM_RAW; return ns; }

static int mid_grid(sisua_model* h, int rows) { return std::max(1, std::min((rows + kTileR - 1) / kTileR, 2 * h->num_sms)); }

template <int AOP, int BOP>
static void launch_sgemm(sisua_model* h, cudaStream_t st, const float* A, long long a_rs, long long a_cs, const float* B,
                         long long b_rs, long long b_cs, float* C, long long ldc, const float* bias, int M, int N, int K,
                         bool accumulate, DropSpec drop = DropSpec{0.f, 1.f, 0u, 0u, 0u, 0u, nullptr}) {
  int tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + kGemmBN - 1) / kGemmBN);
  int want = (2 * h->num_sms + tiles - 1) / tiles;
  int max_splits = std::max(1, K / 256);
  int splits = accumulate ? std::max(1, std::min(want, max_splits)) : 1;   // split-K adds atomically: needs a zeroed C
  int k_chunk = ((K + splits - 1) / splits + kGemmBK - 1) / kGemmBK * kGemmBK;
  splits = (K + k_chunk - 1) / k_chunk;
  dim3 grid((N + kGemmBN - 1) / kGemmBN, (M + kGemmBM - 1) / kGemmBM, splits);
  int atomic_out = (accumulate || splits > 1) ? 1 : 0;
  ++h->launches;
  sgemm_kernel<AOP, BOP><<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, M, N, K, k_chunk, atomic_out, drop);
}

static void launch_dense_fwd(sisua_model* h, cudaStream_t st, const float* A_in, int lda, int Kin, const NormSpec& ns,
                             const float* W, int ldw, const float* bias, int Nout, float* A_out, int ldo, int R,
                             double* out_sum = nullptr) {
  ++h->launches;
  launch_pdl(dense_fwd_kernel, dim3(mid_grid(h, R)), dim3(kMidThreads), kDenseFwdSmem, st, A_in, lda, Kin, ns, W, ldw, bias, Nout, A_out, ldo, R, out_sum,
                                                                   out_sum ? out_sum + kH : nullptr);
}

static void launch_col_stats(sisua_model* h, cudaStream_t st, const Layer& L, int R) {
  double* s = h->stats + (size_t)L.stat_index * 4 * kH;
  int grid = std::max(1, std::min((R + 15) / 16, 2 * h->num_sms));
  ++h->launches;
  launch_pdl(col_stats_kernel, dim3(grid), dim3(256), 0, st, L.A, L.lda, R, kH, s, s + kH);
}

// hidden stack forward: layer 0's pre-activation is already in L[0].A; leaves the last layer's
// pre-activation (plus statistics when training with BN) ready for the consumer.
static void stack_forward(sisua_model* h, cudaStream_t st, std::vector<Layer>& Ls, bool training, int R,
                          bool first_has_stats = false) {
  for (size_t i = 0; i < Ls.size(); ++i) {
    if (i == 0 && !first_has_stats && training && Ls[i].bn_index >= 0) launch_col_stats(h, st, Ls[i], R);
    if (i + 1 < Ls.size()) {
      NormSpec ns = make_norm(h, Ls[i], training, R);
      double* next_stats = (training && Ls[i + 1].bn_index >= 0) ? h->stats + (size_t)Ls[i + 1].stat_index * 4 * kH : nullptr;
      launch_dense_fwd(h, st, Ls[i].A, Ls[i].lda, kH, ns, h->P + Ls[i + 1].w_off, Ls[i + 1].ldw, nullptr, kH,
                       Ls[i + 1].A, Ls[i + 1].lda, R, next_stats);
    }
  }
}

// d loss / d D starts from the protein head's backward (SISUA) or from zero; the output heads add to it
static int init_dD(sisua_model* h, cudaStream_t st, int R) {
  const int H = kH, P = h->cfg.n_proteins;
  if (P > 0) {
    DenseBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.out_mode = 0; a.dOut = h->dPY; a.ldd = 2 * P; a.Nout = 2 * P; a.A_in = h->D; a.lda_in = H; a.Kin = H;
    a.ns_in = raw_norm(); a.W = h->P + h->y_w; a.ldw = H; a.dW = h->Gd + h->y_w; a.db = h->Gd + h->y_b;
    a.dIn = h->dD; a.ldi = H; a.accumulate_dIn = 0; a.R = R;
    ++h->launches;
    launch_pdl(dense_bwd_kernel, dim3(mid_grid(h, R)), dim3(kMidThreads), kDenseBwdSmem, st, a);
    LAUNCH_OK(h, "protein head backward");
  } else {
    CUDA_OK(h, cudaMemsetAsync(h->dD, 0, (size_t)R * H * sizeof(float), st));
  }
  return SISUA_OK;
}

// shared forward; rows_dec = S*B
static int forward_common(sisua_model* h, cudaStream_t st, bool training, const float* x, const float* y,
                          const float* library, const uint8_t* mask, const float* eps_z, const float* eps_l, int B,
                          int S, float* terms, float* loss, float* out_mean, float* out_disp, float* out_pi,
                          float* y_mean) {
  const sisua_step_config& c = h->cfg;
  const int H = kH, G = c.n_genes, Z = c.n_latent, P = c.n_proteins;
  const bool scvi = c.model_kind == SISUA_MODEL_SCVI, dca = c.model_kind == SISUA_MODEL_DCA;
  const int R = S * B;
  if (!h->P) SET_ERR(h, SISUA_ERR_STATE, "bind_buffers has not been called");
  if (B < 1 || S < 1 || R > c.max_batch) SET_ERR(h, SISUA_ERR_INVALID, "rows S*B=%d exceed max_batch=%d", R, c.max_batch);
  const bool decode_only = h->decode_only;
  const bool grads = training && !h->fwd_no_grads;      // gradient side of the fused kernels
  if ((!x && !decode_only) || !terms) SET_ERR(h, SISUA_ERR_INVALID, "x / terms must not be null");
  if (scvi && !library && !decode_only) SET_ERR(h, SISUA_ERR_INVALID, "scVI needs library [B,2]");      // eps_z / eps_l NULL: Philox noise in-kernel
  if (P > 0 && !y && !decode_only) SET_ERR(h, SISUA_ERR_INVALID, "SISUA needs y [B,P]");
  if (training) CUDA_OK(h, cudaMemsetAsync(h->stats, 0, (size_t)h->n_units * 4 * H * sizeof(double), st));
  if (loss) CUDA_OK(h, cudaMemsetAsync(loss, 0, sizeof(float), st));

  const bool fused_latent = (S == 1) && !decode_only;    // one kernel: latent projection -> reparameterisation / KL -> first decoder layer
  if (decode_only) {
    CUDA_OK(h, cudaMemsetAsync(terms + (size_t)3 * R, 0, (size_t)2 * R * sizeof(float), st));   // kl_z = kl_l = 0
#ifdef SISUA_WITH_TC
    if (tc_heads_enabled(h)) h->wout_packed = false;      // (the first-layer launch that also packs the head tiles is skipped)
#endif
  } else {
  // ---- first layer: log1p(x) . W1^T  (z encoder and, for scVI, the library encoder in one pass)
  const int N0 = scvi ? 2 * H : H;
  bool first_done = false;
  sec_begin(h, st, SEC_ENC_FIRST);
#ifdef SISUA_WITH_TC
  if (c.gemm_mode != SISUA_GEMM_FP32_UNFUSED) {
    int rc = tc_encoder_first(h, st, x, B, N0, training, grads);
    if (rc != SISUA_OK) return rc;
    first_done = true;
  }
#endif
  if (!first_done) {
    DropSpec din = make_drop(h, c.input_dropout, 0u, training);
    if (c.log_norm)
      launch_sgemm<LOAD_LOG1P, LOAD_NONE>(h, st, x, G, 1, h->P + h->enc[0].w_off, 1, h->Gp, h->A0, h->ld0, nullptr, B, N0, G, false, din);
    else
      launch_sgemm<LOAD_RAW_DROP, LOAD_NONE>(h, st, x, G, 1, h->P + h->enc[0].w_off, 1, h->Gp, h->A0, h->ld0, nullptr, B, N0, G, false, din);
  }
  LAUNCH_OK(h, "first-layer gemm");
  sec_end(h, st, SEC_ENC_FIRST);
  sec_begin(h, st, SEC_MID_FWD);
  stack_forward(h, st, h->enc, training, B);
  if (fused_latent) {
    Layer& L = h->enc.back();
    LatentBlockFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.A_enc = L.A; a.lda = L.lda; a.ns_enc = make_norm(h, L, training, B);
    a.W_lat = h->P + h->lat_w; a.b_lat = h->P + h->lat_b; a.ZP = dca ? Z : 2 * Z;
    a.eps_z = eps_z; a.noise = make_noise(h); a.W_d0 = h->P + h->dec[0].w_off;
    a.PL = h->PL; a.loc = h->loc; a.scale = h->scale; a.z = h->Zs; a.kl_z = terms + (size_t)3 * R;
    a.logw = h->logw;
    a.A_d0 = h->dec[0].A; a.ldd0 = h->dec[0].lda;
    if (training && h->dec[0].bn_index >= 0) { a.out_sum = h->stats + (size_t)h->dec[0].stat_index * 4 * kH; a.out_sumsq = a.out_sum + kH; }
    a.B = B; a.Z = Z; a.deterministic = dca ? (c.latent_linear ? 2 : 1) : 0; a.scale_act = c.scale_act;
    ++h->launches;
    launch_pdl(latent_block_fwd_kernel, dim3(mid_grid(h, B)), dim3(kMidThreads), kLatentFwdSmem, st, a);
    LAUNCH_OK(h, "latent_block_fwd_kernel");
  } else {
    Layer& L = h->enc.back();
    NormSpec ns = make_norm(h, L, training, B);
    const int ZP = dca ? Z : 2 * Z;
    launch_dense_fwd(h, st, L.A, L.lda, H, ns, h->P + h->lat_w, H, h->P + h->lat_b, ZP, h->PL, ZP, B);
  }
  if (scvi) {
    stack_forward(h, st, h->encl, training, B);
    Layer& L = h->encl.back();
    NormSpec ns = make_norm(h, L, training, B);
    launch_dense_fwd(h, st, L.A, L.lda, H, ns, h->P + h->lib_w, H, h->P + h->lib_b, 2, h->PLIB, 2, B);
  }
  LAUNCH_OK(h, "encoder stack");
  if (!fused_latent || scvi) {
    LatentArgs a;
    memset(&a, 0, sizeof(a));
    a.PL = fused_latent ? nullptr : h->PL; a.eps_z = eps_z; a.noise = make_noise(h); a.loc = h->loc; a.scale = h->scale; a.z = h->Zs;
    a.kl_z = terms + (size_t)3 * R; a.kl_l = terms + (size_t)4 * R;
    a.logw = h->logw;
    if (scvi) {
      a.PLIB = h->PLIB; a.eps_l = eps_l; a.library = library; a.lib_loc = h->lib_loc; a.lib_scale = h->lib_scale;
      a.lib = h->lib;
    }
    a.B = B; a.S = S; a.Z = Z; a.deterministic = dca ? (c.latent_linear ? 2 : 1) : 0; a.scale_act = c.scale_act;
    ++h->launches;
    launch_pdl(latent_fwd_kernel, dim3((B + 127) / 128), dim3(128), 0, st, a);
    LAUNCH_OK(h, "latent_fwd_kernel");
  } else {
    CUDA_OK(h, cudaMemsetAsync(terms + (size_t)4 * R, 0, (size_t)R * sizeof(float), st));   // kl_l = 0
  }
  }   // !decode_only
  // ---- decoder
  if (!fused_latent) {
    double* st0 = (training && h->dec[0].bn_index >= 0) ? h->stats + (size_t)h->dec[0].stat_index * 4 * kH : nullptr;
    launch_dense_fwd(h, st, h->Zs, Z, Z, raw_norm(), h->P + h->dec[0].w_off, h->dec[0].ldw, nullptr, H, h->dec[0].A,
                     h->dec[0].lda, R, st0);
  }
  stack_forward(h, st, h->dec, training, R, true);
  NormSpec ns_d = make_norm(h, h->dec.back(), training, R);
  // the fused output-head kernel applies the last unit's norm + ReLU + dropout while it loads its cell tile; the
  // activated matrix is only materialised for the protein head and the un-fused cross-check path
  bool fuse_dec_norm = false;
#ifdef SISUA_WITH_TC
  fuse_dec_norm = tc_heads_enabled(h) && P == 0;
#endif
  if (!fuse_dec_norm) {
    ++h->launches;
    launch_pdl(norm_relu_kernel, dim3(std::max(1, std::min((R * H + 255) / 256, 4 * h->num_sms))), dim3(256), 0, st,
        h->dec.back().A, h->dec.back().lda, ns_d, h->D, R);
    LAUNCH_OK(h, "decoder stack");
  }
  // ---- protein head (before the output layer so dD can be initialised by its backward)
  if (P > 0) {
    launch_dense_fwd(h, st, h->D, H, H, raw_norm(), h->P + h->y_w, H, h->P + h->y_b, 2 * P, h->PY, 2 * P, R);
    if (c.mask_norm == 1) { ++h->launches; mask_scale_kernel<<<1, 256, 0, st>>>(mask, B, h->mask_scale); }
    YHeadArgs a;
    memset(&a, 0, sizeof(a));
    a.PY = h->PY; a.y = y; a.mask = mask; a.llk_y = terms + (size_t)2 * R; a.dPY = grads ? h->dPY : nullptr;
    a.y_mean = y_mean; a.R = R; a.B = B; a.P = P; a.y_dist = c.y_dist; a.mean_act = c.mean_act; a.disp_act = c.disp_act;
    a.upstream = -c.alpha / (float)R;
    a.mask_scale = c.mask_norm == 1 ? h->mask_scale : nullptr;
    ++h->launches;
    launch_pdl(yhead_kernel, dim3((R + 127) / 128), dim3(128), 0, st, a);
    LAUNCH_OK(h, "yhead_kernel");
  } else {
    CUDA_OK(h, cudaMemsetAsync(terms + (size_t)2 * R, 0, (size_t)R * sizeof(float), st));
  }
  sec_end(h, st, SEC_MID_FWD);
  // ---- output heads + count likelihood
  sec_begin(h, st, SEC_OUT_HEADS);
  bool out_done = false;
#ifdef SISUA_WITH_TC
  if (tc_heads_enabled(h)) {
    if (grads) {   // the fused kernel adds its d loss / d D on top of the protein head's contribution
      int rc0 = init_dD(h, st, R);
      if (rc0 != SISUA_OK) return rc0;
    }
    int rc = tc_output_heads(h, st, grads, (!training && h->x_eval) ? h->x_eval : x, B, S, terms + (size_t)R, out_mean, out_disp, out_pi,
                             fuse_dec_norm ? &ns_d : nullptr);
    if (rc != SISUA_OK) return rc;
    if (grads && h->ev_out_grads) CUDA_OK(h, cudaEventRecord(h->ev_out_grads, st));
    out_done = true;
  }
#endif
  if (!out_done) {
    launch_sgemm<LOAD_NONE, LOAD_NONE>(h, st, h->D, H, 1, h->P + h->out_w, 1, H, h->OUT, h->NO, h->P + h->out_b, R, h->NO, H, false);
    CountRowArgs a;
    memset(&a, 0, sizeof(a));
    if (!training && (h->nozi || h->out_mean_avg)) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "infer_ex options need the fused tcgen05 heads (gemm_mode 1)");
    a.OUT = h->OUT; a.ldo = h->NO; a.x = (!training && h->x_eval) ? h->x_eval : x; a.lib = scvi ? h->lib : nullptr; a.llk_x = terms + (size_t)R;
    a.dlib = (scvi && grads) ? h->dLib : nullptr;
    a.out_mean = out_mean; a.out_disp = out_disp; a.out_pi = out_pi;
    a.R = R; a.B = B; a.G = G; a.scvi = scvi ? 1 : 0; a.zero_inflated = n_heads(c) == 3; a.tfp = tfp_links(c) ? 1 : 0;
    a.train = grads ? 1 : 0; a.mean_act = c.mean_act; a.disp_act = c.disp_act; a.reapply = c.scvi_reapply_act;
    a.upstream = -1.0f / (float)R; a.clip_library = c.clip_library;
    size_t smem = scvi ? (size_t)G * sizeof(float) : 0;
    ++h->launches;
    if (a.zero_inflated) count_row_kernel<true><<<R, 256, smem, st>>>(a);
    else count_row_kernel<false><<<R, 256, smem, st>>>(a);
    LAUNCH_OK(h, "count_row_kernel");
  }
  sec_end(h, st, SEC_OUT_HEADS);
  // ---- ELBO (+ moving BatchNorm statistics, folded into the same launch)
  {
    ElboArgs a;
    memset(&a, 0, sizeof(a));
    a.terms = terms; a.mask = P > 0 ? mask : nullptr; a.mask_scale = (P > 0 && c.mask_norm == 1) ? h->mask_scale : nullptr;
    a.R = R; a.B = B; a.alpha = c.alpha; a.beta = c.beta; a.loss = loss;
    a.nonfinite = training ? h->nf_dev : nullptr;
    if (training && c.batchnorm) {
      auto reg = [&](const Layer& L, int rows) {
        a.mu.sum[L.bn_index] = h->stats + (size_t)L.stat_index * 4 * kH;
        a.mu.sumsq[L.bn_index] = a.mu.sum[L.bn_index] + kH;
        a.mu.inv_count[L.bn_index] = 1.0f / (float)rows;
      };
      for (auto& L : h->enc) reg(L, B);
      for (auto& L : h->encl) reg(L, B);
      for (auto& L : h->dec) reg(L, R);
      a.n_bn = h->n_bn; a.moving = h->moving; a.momentum = c.bn_momentum;
    }
    ++h->launches;
    launch_pdl(elbo_kernel, dim3(std::max(1, std::min((R + 255) / 256, h->num_sms)) + a.n_bn), dim3(256), 0, st, a);
    LAUNCH_OK(h, "elbo_kernel");
  }
  return SISUA_OK;
}

// backward of a hidden stack. dH_top: gradient wrt the activated output of the last layer.
// in0: source of layer 0's input for dW_0 (null -> dW_0 comes from the big first-layer GEMM),
// dIn0: where the gradient wrt layer 0's input goes (null -> not needed), dA0: pre-activation
// gradient of layer 0 (delta1) when requested.
static int stack_backward(sisua_model* h, cudaStream_t st, std::vector<Layer>& Ls, int R, float* dH_top,
                          const float* in0, int ld_in0, int Kin0, float* dIn0, int ld_dIn0, float* dA0, int ld_dA0,
                          bool top_reduced = false, int stop_at = 0, float** dH_out = nullptr) {
  float* dH = dH_top;
  if (dH_out) *dH_out = dH_top;
  for (int i = (int)Ls.size() - 1; i >= stop_at; --i) {
    Layer& L = Ls[i];
    NormSpec ns = make_norm(h, L, true, R);
    double* sdy = h->stats + (size_t)L.stat_index * 4 * kH + 2 * kH;
    double* sdyx = sdy + kH;
    if (i == (int)Ls.size() - 1 && !top_reduced) {   // top unit: no consumer kernel produced its reductions
      float* dgamma = L.g_off >= 0 ? h->Gd + L.g_off : nullptr;
      float* dbeta = h->Gd + L.b_off;
      int grid = std::max(1, std::min((R + 15) / 16, 2 * h->num_sms));
      ++h->launches;
      launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(256), 0, st, dH, kH, L.A, L.lda, ns, R, sdy, sdyx, dgamma, dbeta);
    }
    if (i == 0 && !in0 && !dIn0 && dA0) {   // only the pre-activation gradient is wanted: light element-wise kernel
      ++h->launches;
      launch_pdl(bn_bwd_apply_kernel, dim3(std::max(1, std::min((R * (kH / 4) + 255) / 256, 2 * h->num_sms))), dim3(256), 0, st,
                 (const float*)dH, kH, (const float*)L.A, L.lda, ns, (const double*)sdy, (const double*)sdyx, dA0, ld_dA0, R);
      LAUNCH_OK(h, "bn_bwd_apply_kernel");
      if (dH_out) *dH_out = dH;
      break;
    }
    DenseBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.out_mode = 1; a.dOut = dH; a.ldd = kH; a.A_out = L.A; a.lda_out = L.lda; a.ns_out = ns; a.sdy = sdy; a.sdyx = sdyx;
    a.Nout = kH; a.R = R;
    float* dH_next = (dH == h->dHa) ? h->dHb : h->dHa;
    if (i > 0) {
      a.A_in = Ls[i - 1].A; a.lda_in = Ls[i - 1].lda; a.Kin = kH; a.ns_in = make_norm(h, Ls[i - 1], true, R);
      a.W = h->P + L.w_off; a.ldw = L.ldw; a.dW = h->Gd + L.w_off;
      a.dIn = dH_next; a.ldi = kH; a.accumulate_dIn = 0;
      // the unit below gets its norm-backward reductions from this kernel
      Layer& Lp = Ls[i - 1];
      a.prev_sdy = h->stats + (size_t)Lp.stat_index * 4 * kH + 2 * kH; a.prev_sdyx = a.prev_sdy + kH;
      a.prev_dgamma = Lp.g_off >= 0 ? h->Gd + Lp.g_off : nullptr; a.prev_dbeta = h->Gd + Lp.b_off;
    } else {
      a.A_in = in0; a.lda_in = ld_in0; a.Kin = in0 ? Kin0 : kH; a.ns_in = raw_norm();
      a.W = in0 ? h->P + L.w_off : nullptr; a.ldw = L.ldw; a.dW = in0 ? h->Gd + L.w_off : nullptr;
      a.dIn = dIn0; a.ldi = ld_dIn0; a.accumulate_dIn = 0;
      a.dA = dA0; a.ldda = ld_dA0;
    }
    ++h->launches;
    launch_pdl(dense_bwd_kernel, dim3(mid_grid(h, R)), dim3(kMidThreads), kDenseBwdSmem, st, a);
    LAUNCH_OK(h, "hidden backward");
    dH = dH_next;
    if (dH_out) *dH_out = dH;
  }
  return SISUA_OK;
}

extern "C" int sisua_train_step(sisua_handle h, const float* x, const float* y, const float* library,
                                const uint8_t* mask, const float* eps_z, const float* eps_l, int B, uint64_t seed,
                                int64_t step, float* terms, float* loss, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const sisua_step_config& c = h->cfg;
  h->drop_seed = seed;
  h->drop_step = (uint32_t)(step >= 0 ? step : 0);
  h->drop_step_on_device = step < 0;
  if (!h->Gd) SET_ERR(h, SISUA_ERR_STATE, "train_step needs a bound grads buffer");
  if (c.batchnorm && B < 2) SET_ERR(h, SISUA_ERR_INVALID, "training-mode batch norm needs B >= 2");
  const int H = kH, G = c.n_genes, Z = c.n_latent, P = c.n_proteins, R = B;
  const bool scvi = c.model_kind == SISUA_MODEL_SCVI, dca = c.model_kind == SISUA_MODEL_DCA;
  CUDA_OK(h, cudaMemsetAsync(h->Gd, 0, (size_t)h->total_floats * sizeof(float), st));
  int rc = forward_common(h, st, true, x, y, library, mask, eps_z, eps_l, B, 1, terms, loss, nullptr, nullptr, nullptr, nullptr);
  if (rc != SISUA_OK) return rc;
  bool out_done = false;
#ifdef SISUA_WITH_TC
  if (tc_heads_enabled(h)) out_done = true;   // fused kernel already produced dW_out, db_out, dD
#endif
  if (!out_done) {
    rc = init_dD(h, st, R);
    if (rc != SISUA_OK) return rc;
    sec_begin(h, st, SEC_OUT_HEADS);
    // dW_out[NO,H] += G_out^T . D ; db_out += colsum(G_out) ; dD += G_out . W_out
    launch_sgemm<LOAD_NONE, LOAD_NONE>(h, st, h->OUT, 1, h->NO, h->D, H, 1, h->Gd + h->out_w, H, nullptr, h->NO, H, R, true);
    int rows_per_block = std::max(64, (R + 63) / 64);
    dim3 g((h->NO + 255) / 256, (R + rows_per_block - 1) / rows_per_block);
    ++h->launches;
    col_sum_kernel<<<g, 256, 0, st>>>(h->OUT, h->NO, R, h->NO, rows_per_block, h->Gd + h->out_b);
    launch_sgemm<LOAD_NONE, LOAD_NONE>(h, st, h->OUT, h->NO, 1, h->P + h->out_w, H, 1, h->dD, H, nullptr, R, H, h->NO, true);
    LAUNCH_OK(h, "output-layer backward");
    if (h->ev_out_grads) CUDA_OK(h, cudaEventRecord(h->ev_out_grads, st));
    sec_end(h, st, SEC_OUT_HEADS);
  }
  // ---- decoder stack, latent, encoder stack(s)
  sec_begin(h, st, SEC_MID_BWD);
  // decoder units 1.. (unit 0 is handled by the fused latent block below)
  float* dH_d0 = h->dD;
  rc = stack_backward(h, st, h->dec, R, h->dD, nullptr, 0, 0, nullptr, 0, nullptr, 0, false, 1, &dH_d0);
  if (rc != SISUA_OK) return rc;
  {
    Layer& L0 = h->dec[0];
    Layer& Le = h->enc.back();
    NormSpec ns0 = make_norm(h, L0, true, R);
    double* sdy0 = h->stats + (size_t)L0.stat_index * 4 * kH + 2 * kH;
    if (h->dec.size() == 1) {     // no decoder unit above produced unit 0's norm-backward reductions
      int grid = std::max(1, std::min((R + 15) / 16, 2 * h->num_sms));
      ++h->launches;
      launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(256), 0, st, (const float*)dH_d0, kH, (const float*)L0.A, L0.lda, ns0, R, sdy0,
                 sdy0 + kH, L0.g_off >= 0 ? h->Gd + L0.g_off : (float*)nullptr, h->Gd + L0.b_off);
    }
    // the staging buffer for dH_enc must differ from dH_d0 (both ping-pong buffers may be in use)
    float* dH_enc = (dH_d0 == h->dHa) ? h->dHb : h->dHa;
    LatentBlockBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.dH_d0 = dH_d0; a.A_d0 = L0.A; a.ldd0 = L0.lda; a.ns_d0 = ns0; a.sdy = sdy0; a.sdyx = sdy0 + kH;
    a.W_d0 = h->P + L0.w_off; a.dW_d0 = h->Gd + L0.w_off;
    a.z = h->Zs; a.PL = h->PL; a.eps_z = eps_z; a.noise = make_noise(h); a.loc = h->loc; a.scale = h->scale;
    a.W_lat = h->P + h->lat_w; a.dW_lat = h->Gd + h->lat_w; a.db_lat = h->Gd + h->lat_b; a.ZP = dca ? Z : 2 * Z;
    a.A_enc = Le.A; a.lda = Le.lda; a.ns_enc = make_norm(h, Le, true, B);
    a.dH_enc = dH_enc;
    a.prev_sdy = h->stats + (size_t)Le.stat_index * 4 * kH + 2 * kH; a.prev_sdyx = a.prev_sdy + kH;
    a.prev_dgamma = Le.g_off >= 0 ? h->Gd + Le.g_off : nullptr; a.prev_dbeta = h->Gd + Le.b_off;
    a.B = B; a.Z = Z; a.deterministic = dca ? (c.latent_linear ? 2 : 1) : 0; a.scale_act = c.scale_act; a.kl_weight = c.beta / (float)B;
    ++h->launches;
    launch_pdl(latent_block_bwd_kernel, dim3(mid_grid(h, B)), dim3(kMidThreads), Z <= 16 ? kLatentBwdSmemNarrow : kLatentBwdSmem, st, a);
    LAUNCH_OK(h, "latent_block_bwd_kernel");
    if (scvi) {     // library latent: d lib -> d(raw loc, raw scale)
      LibraryBwdArgs lb;
      memset(&lb, 0, sizeof(lb));
      lb.dLib = h->dLib; lb.PLIB = h->PLIB; lb.eps_l = eps_l; lb.noise = make_noise(h); lb.library = library; lb.lib_loc = h->lib_loc;
      lb.lib_scale = h->lib_scale; lb.dPLIB = h->dPLIB;
      lb.B = B; lb.scale_act = c.scale_act; lb.kl_weight = c.beta / (float)B;
      ++h->launches;
      launch_pdl(library_bwd_kernel, dim3((B + 127) / 128), dim3(128), 0, st, lb);
      LAUNCH_OK(h, "library_bwd_kernel");
    }
    rc = stack_backward(h, st, h->enc, B, dH_enc, nullptr, 0, 0, nullptr, 0, h->delta1, h->ld0, true);
    if (rc != SISUA_OK) return rc;
  }
  if (scvi) {
    Layer& L = h->encl.back();
    DenseBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.out_mode = 0; a.dOut = h->dPLIB; a.ldd = 2; a.Nout = 2; a.A_in = L.A; a.lda_in = L.lda; a.Kin = H;
    a.ns_in = make_norm(h, L, true, B); a.W = h->P + h->lib_w; a.ldw = H; a.dW = h->Gd + h->lib_w; a.db = h->Gd + h->lib_b;
    a.dIn = h->dHa; a.ldi = H; a.R = B;
    a.prev_sdy = h->stats + (size_t)L.stat_index * 4 * kH + 2 * kH; a.prev_sdyx = a.prev_sdy + kH;
    a.prev_dgamma = L.g_off >= 0 ? h->Gd + L.g_off : nullptr; a.prev_dbeta = h->Gd + L.b_off;
    ++h->launches;
    launch_pdl(dense_bwd_kernel, dim3(mid_grid(h, B)), dim3(kMidThreads), kDenseBwdSmem, st, a);
    LAUNCH_OK(h, "library projection backward");
    rc = stack_backward(h, st, h->encl, B, h->dHa, nullptr, 0, 0, nullptr, 0, h->delta1 + H, h->ld0, true);
    if (rc != SISUA_OK) return rc;
  }
  sec_end(h, st, SEC_MID_BWD);
  // ---- first-layer weight gradient: dW1[N0, G] = delta1^T . log1p(x)
  const int N0 = scvi ? 2 * H : H;
  bool w1_done = false;
  sec_begin(h, st, SEC_ENC_FIRST_BWD);
#ifdef SISUA_WITH_TC
  if (c.gemm_mode != SISUA_GEMM_FP32_UNFUSED) {
    rc = tc_encoder_first_bwd(h, st, x, B, N0);
    if (rc != SISUA_OK) return rc;
    w1_done = true;
  }
#endif
  if (!w1_done) {
    DropSpec din = make_drop(h, c.input_dropout, 0u, true);
    if (c.log_norm)
      launch_sgemm<LOAD_NONE, LOAD_LOG1P>(h, st, h->delta1, 1, h->ld0, x, G, 1, h->Gd + h->enc[0].w_off, h->Gp, nullptr, N0, G, B, true, din);
    else
      launch_sgemm<LOAD_NONE, LOAD_RAW_DROP>(h, st, h->delta1, 1, h->ld0, x, G, 1, h->Gd + h->enc[0].w_off, h->Gp, nullptr, N0, G, B, true, din);
    LAUNCH_OK(h, "first-layer weight gradient");
  }
  sec_end(h, st, SEC_ENC_FIRST_BWD);
  return SISUA_OK;
}

// rows of a resident matrix -> dense minibatch (small per-cell side inputs; counts only for the un-fused path)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ ridx, float* __restrict__ dst,
                                                          int B, int width) {
  const long long n = (long long)B * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / width), c = (int)(i % width);
    dst[i] = src[(size_t)ridx[b] * width + c];
  }
}
__global__ void __launch_bounds__(256) gather_bytes_kernel(const uint8_t* __restrict__ src, const int* __restrict__ ridx,
                                                           uint8_t* __restrict__ dst, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dst[b] = src[ridx[b]];
}

// Same step on a minibatch given as ROW INDICES into matrices that stay resident in HBM (what `fit(shuffle=True)` draws
// every step: sisua/data/_single_cell_base.py:593-601 shuffle -> batch): x_all [N,G], y_all [N,P], library_all [N,2],
// mask_all [N], rows [B] int32 (device).  The tcgen05 kernels read the count rows through the index (no gathered copy
// of the minibatch is ever written); the small per-cell side inputs are gathered by one tiny kernel each.
// gather + widen rows of a uint16 count matrix into fp32 (rows == NULL: rows 0 .. n_rows-1)
__global__ void __launch_bounds__(256) widen_rows_u16_kernel(const uint16_t* __restrict__ src, const int* __restrict__ ridx,
                                                             float* __restrict__ dst, int B, int G) {
  const long long n = (long long)B * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / G), g = (int)(i - (long long)b * G);
    dst[i] = (float)src[(size_t)(ridx ? ridx[b] : b) * G + g];
  }
}

extern "C" int sisua_widen_rows_u16(sisua_handle h, const uint16_t* x_all, const int32_t* rows, float* dst, int n_rows, void* stream) {
  if (!h || !x_all || !dst || n_rows < 0) return SISUA_ERR_INVALID;
  if (n_rows == 0) return SISUA_OK;
  const int G = h->cfg.n_genes;
  ++h->launches;
  widen_rows_u16_kernel<<<std::max(1, std::min(8 * h->num_sms, (int)(((long long)n_rows * G + 255) / 256))), 256, 0, (cudaStream_t)stream>>>(
      x_all, rows, dst, n_rows, G);
  LAUNCH_OK(h, "widen_rows_u16_kernel");
  return SISUA_OK;
}

static int train_step_gather_impl(sisua_handle h, const void* x_any, bool u16, const float* y_all, const float* library_all,
                                  const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                                  uint64_t seed, int64_t step, float* terms, float* loss, void* stream);

extern "C" int sisua_train_step_gather(sisua_handle h, const float* x_all, const float* y_all, const float* library_all,
                                       const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                                       uint64_t seed, int64_t step, float* terms, float* loss, void* stream) {
  return train_step_gather_impl(h, x_all, false, y_all, library_all, mask_all, rows, eps_z, eps_l, B, seed, step, terms, loss, stream);
}

// The same step with the resident count matrix stored as uint16 (exact for count data below 65 536: half the HBM of the
// shard and half the bytes of both streaming reads).  The two tcgen05 kernels widen the counts themselves when the rows
// are 16-byte aligned (n_genes a multiple of 8, aligned base) and the model's heads are the plain (non-scVI) ones;
// otherwise the minibatch rows are widened into an fp32 staging buffer first.
extern "C" int sisua_train_step_gather_u16(sisua_handle h, const uint16_t* x_all, const float* y_all, const float* library_all,
                                           const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                                           uint64_t seed, int64_t step, float* terms, float* loss, void* stream) {
  return train_step_gather_impl(h, x_all, true, y_all, library_all, mask_all, rows, eps_z, eps_l, B, seed, step, terms, loss, stream);
}

static int train_step_gather_impl(sisua_handle h, const void* x_any, bool u16, const float* y_all, const float* library_all,
                                  const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                                  uint64_t seed, int64_t step, float* terms, float* loss, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  const float* x_all = reinterpret_cast<const float*>(x_any);
  if (!rows && !u16) return sisua_train_step(h, x_all, y_all, library_all, mask_all, eps_z, eps_l, B, seed, step, terms, loss, stream);
  const sisua_step_config& c = h->cfg;
  if (B < 1 || B > c.max_batch) SET_ERR(h, SISUA_ERR_INVALID, "train_step_gather: B=%d outside [1, max_batch=%d]", B, c.max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t R = c.max_batch;
  int rc;
  const float* x = x_all; const float* y = y_all; const float* lib = library_all; const uint8_t* mask = mask_all;
  bool tc_path = false;
#ifdef SISUA_WITH_TC
  tc_path = tc_heads_enabled(h);
#endif
  // uint16 counts stay uint16 all the way into the kernels when the geometry allows 16-byte copies
  const int32_t* x_rows = rows;      // the index the tcgen05 kernels read the counts through
  const bool u16_direct = u16 && tc_path && c.model_kind != SISUA_MODEL_SCVI && c.n_genes % 8 == 0 &&
                          (reinterpret_cast<uintptr_t>(x_any) & 15) == 0;
  if (u16 && !u16_direct) {
    if (!h->gx && (rc = ws_alloc(h, &h->gx, R * c.n_genes)) != SISUA_OK) return rc;
    if ((rc = sisua_widen_rows_u16(h, reinterpret_cast<const uint16_t*>(x_any), rows, h->gx, B, stream)) != SISUA_OK) return rc;
    x = h->gx; x_rows = nullptr;      // (the minibatch is now a dense fp32 copy)
  } else if (!tc_path) {       // cross-check path: dense copy of the counts
    if (!h->gx && (rc = ws_alloc(h, &h->gx, R * c.n_genes)) != SISUA_OK) return rc;
    ++h->launches;
    gather_rows_kernel<<<std::max(1, std::min(4 * h->num_sms, (int)(((long long)B * c.n_genes + 255) / 256))), 256, 0, st>>>(x_all, rows, h->gx, B, c.n_genes);
    x = h->gx; x_rows = nullptr;
  }
  if (rows && y_all && c.n_proteins > 0) {
    if (!h->gy && (rc = ws_alloc(h, &h->gy, R * c.n_proteins)) != SISUA_OK) return rc;
    ++h->launches;
    gather_rows_kernel<<<(B * c.n_proteins + 255) / 256, 256, 0, st>>>(y_all, rows, h->gy, B, c.n_proteins);
    y = h->gy;
  }
  if (rows && library_all) {
    if (!h->glib && (rc = ws_alloc(h, &h->glib, R * 2)) != SISUA_OK) return rc;
    ++h->launches;
    gather_rows_kernel<<<(B * 2 + 255) / 256, 256, 0, st>>>(library_all, rows, h->glib, B, 2);
    lib = h->glib;
  }
  if (rows && mask_all) {
    if (!h->gmask && (rc = ws_alloc(h, &h->gmask, R)) != SISUA_OK) return rc;
    ++h->launches;
    gather_bytes_kernel<<<(B + 255) / 256, 256, 0, st>>>(mask_all, rows, h->gmask, B);
    mask = h->gmask;
  }
  LAUNCH_OK(h, "row gather");
  h->ridx = tc_path ? x_rows : nullptr;
  h->x_u16 = u16_direct;
  rc = sisua_train_step(h, x, y, lib, mask, eps_z, eps_l, B, seed, step, terms, loss, stream);
  h->ridx = nullptr;
  h->x_u16 = false;
  return rc;
}

extern "C" int sisua_infer(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                           const float* eps_z, const float* eps_l, int B, int S, float* terms, float* z_loc,
                           float* z_scale, float* lib_loc, float* lib_scale, float* out_mean, float* out_disp,
                           float* out_pi, float* y_mean, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const sisua_step_config& c = h->cfg;
  float* t = terms ? terms : h->scratch_terms;
  // noise of a call without injected eps: Philox(infer seed; row, column, call index, stream)
  h->drop_seed = h->infer_seed; h->drop_step = (uint32_t)(h->infer_calls++); h->drop_step_on_device = false;
  int rc = forward_common(h, st, false, x, y, library, mask, eps_z, eps_l, B, S, t, nullptr, out_mean, out_disp, out_pi, y_mean);
  if (rc != SISUA_OK) return rc;
  const size_t zb = (size_t)B * c.n_latent * sizeof(float);
  if (z_loc) CUDA_OK(h, cudaMemcpyAsync(z_loc, h->loc, zb, cudaMemcpyDeviceToDevice, st));
  if (z_scale) CUDA_OK(h, cudaMemcpyAsync(z_scale, h->scale, zb, cudaMemcpyDeviceToDevice, st));
  if (c.model_kind == SISUA_MODEL_SCVI) {
    if (lib_loc) CUDA_OK(h, cudaMemcpyAsync(lib_loc, h->lib_loc, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (lib_scale) CUDA_OK(h, cudaMemcpyAsync(lib_scale, h->lib_scale, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return SISUA_OK;
}

// sisua_infer with the options the Posterior fast paths need (sisua/analysis/posterior.py:210-220,919-976):
//   x_eval       counts the log-likelihood is evaluated on while the encoder reads x (llk of ORIGINAL counts under the
//                model of the CORRUPTED ones); NULL = x
//   strip_zi     likelihood / parameters of the count distribution without its zero inflation ("imputed")
//   out_mean_avg [B,G] mean over the S Monte-Carlo samples of the NB mean, accumulated in the fused epilogue
//   logw         [S*B] log p(z_s) - log q(z_s | x) (+ the library latent's): importance weights
extern "C" int sisua_infer_ex(sisua_handle h, const float* x, const float* x_eval, const float* y, const float* library,
                              const uint8_t* mask, const float* eps_z, const float* eps_l, int B, int S, int strip_zi,
                              float* terms, float* z_loc, float* z_scale, float* lib_loc, float* lib_scale, float* out_mean,
                              float* out_disp, float* out_pi, float* y_mean, float* out_mean_avg, float* logw, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  if (out_mean_avg) {
    cudaError_t e = cudaMemsetAsync(out_mean_avg, 0, (size_t)B * h->cfg.n_genes * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) SET_ERR(h, SISUA_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
  }
  h->x_eval = x_eval; h->nozi = strip_zi ? 1 : 0; h->out_mean_avg = out_mean_avg; h->logw = logw;
  const int rc = sisua_infer(h, x, y, library, mask, eps_z, eps_l, B, S, terms, z_loc, z_scale, lib_loc, lib_scale, out_mean, out_disp,
                             out_pi, y_mean, stream);
  h->x_eval = nullptr; h->nozi = 0; h->out_mean_avg = nullptr; h->logw = nullptr;
  return rc;
}

// Training-mode forward pass WITHOUT gradients or weight update (`model(..., training=True)` outside fit,
// single_cell_model.py:178): BatchNorm uses the batch statistics and folds them into the moving ones, dropout masks are
// drawn from (seed, step); outputs as sisua_infer with S = 1.
extern "C" int sisua_forward_train_mode(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                                        const float* eps_z, const float* eps_l, int B, uint64_t seed, int64_t step, float* terms,
                                        float* z_loc, float* z_scale, float* lib_loc, float* lib_scale, float* out_mean,
                                        float* out_disp, float* out_pi, float* y_mean, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const sisua_step_config& c = h->cfg;
  if (c.batchnorm && B < 2) SET_ERR(h, SISUA_ERR_INVALID, "training-mode batch norm needs B >= 2");
  h->drop_seed = seed; h->drop_step = (uint32_t)(step >= 0 ? step : 0); h->drop_step_on_device = step < 0;
  h->fwd_no_grads = true;
  const int rc = forward_common(h, st, true, x, y, library, mask, eps_z, eps_l, B, 1, terms ? terms : h->scratch_terms, nullptr, out_mean,
                                out_disp, out_pi, y_mean);
  h->fwd_no_grads = false;
  if (rc != SISUA_OK) return rc;
  const size_t zb = (size_t)B * c.n_latent * sizeof(float);
  if (z_loc) CUDA_OK(h, cudaMemcpyAsync(z_loc, h->loc, zb, cudaMemcpyDeviceToDevice, st));
  if (z_scale) CUDA_OK(h, cudaMemcpyAsync(z_scale, h->scale, zb, cudaMemcpyDeviceToDevice, st));
  if (c.model_kind == SISUA_MODEL_SCVI) {
    if (lib_loc) CUDA_OK(h, cudaMemcpyAsync(lib_loc, h->lib_loc, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (lib_scale) CUDA_OK(h, cudaMemcpyAsync(lib_scale, h->lib_scale, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return SISUA_OK;
}

// Decoder only (SingleCellModel.decode(latents), single_cell_model.py:141-151; scvi.py:108-171): latent samples z [R, Z]
// (and, for scVI, sampled log-library sizes lib [R]) -> parameters of the output distribution(s) with the moving
// BatchNorm statistics: out_mean / out_disp / out_pi [R, G], y_mean [R, P]; any output may be NULL.
extern "C" int sisua_decode(sisua_handle h, const float* z, const float* lib, int R, float* out_mean, float* out_disp, float* out_pi,
                            float* y_mean, void* stream) {
  if (!h || !z) return SISUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const sisua_step_config& c = h->cfg;
  if (R < 1 || R > c.max_batch) SET_ERR(h, SISUA_ERR_INVALID, "decode: R=%d outside [1, max_batch=%d]", R, c.max_batch);
  if (c.model_kind == SISUA_MODEL_SCVI && !lib) SET_ERR(h, SISUA_ERR_INVALID, "decode: scVI needs the sampled log-library sizes");
  CUDA_OK(h, cudaMemcpyAsync(h->Zs, z, (size_t)R * c.n_latent * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (c.model_kind == SISUA_MODEL_SCVI) CUDA_OK(h, cudaMemcpyAsync(h->lib, lib, (size_t)R * sizeof(float), cudaMemcpyDeviceToDevice, st));
#ifdef SISUA_WITH_TC
  if (!tc_heads_enabled(h)) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "decode needs the fused tcgen05 output heads (gemm_mode 1)");
#else
  SET_ERR(h, SISUA_ERR_UNSUPPORTED, "decode needs the fused tcgen05 output heads");
#endif
  h->decode_only = true;
  // S = 1, B = R: the count likelihood of the fused heads is evaluated against zeros (x == NULL) and discarded
  const int rc = forward_common(h, st, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, R, 1, h->scratch_terms, nullptr, out_mean,
                                out_disp, out_pi, y_mean);
  h->decode_only = false;
  return rc;
}

// Importance-weighted marginal log-likelihood of one minibatch (SingleCellModel.marginal_log_prob,
// posterior.py:964-968): S Monte-Carlo samples through sisua_infer_ex, then per cell
//   mllk = logsumexp_s(llk_x + alpha * mask * llk_y + log p(z_s) - log q(z_s | x)) - log S
// and llk_x / llk_y = logsumexp_s(.) - log S.  Outputs [B] each; llk_y may be NULL.
extern "C" int sisua_marginal_llk(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                                  const float* eps_z, const float* eps_l, int B, int S, float* mllk, float* llk_x, float* llk_y,
                                  void* stream) {
  if (!h || !mllk || !llk_x) return SISUA_ERR_INVALID;
  const sisua_step_config& c = h->cfg;
  if (B < 1 || S < 1 || (long long)B * S > c.max_batch) SET_ERR(h, SISUA_ERR_INVALID, "marginal_llk: S*B=%lld exceeds max_batch=%d", (long long)B * S, c.max_batch);
  int rc;
  if (!h->mw_logw && (rc = ws_alloc(h, &h->mw_logw, (size_t)c.max_batch)) != SISUA_OK) return rc;
  rc = sisua_infer_ex(h, x, nullptr, y, library, mask, eps_z, eps_l, B, S, 0, h->scratch_terms, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr, nullptr, nullptr, nullptr, h->mw_logw, stream);
  if (rc != SISUA_OK) return rc;
  MarginalArgs a;
  memset(&a, 0, sizeof(a));
  a.terms = h->scratch_terms; a.logw = h->mw_logw; a.mask = mask; a.alpha = c.alpha; a.B = B; a.S = S; a.has_y = c.n_proteins > 0 ? 1 : 0;
  a.mllk = mllk; a.llk_x = llk_x; a.llk_y = llk_y;
  ++h->launches;
  marginal_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  LAUNCH_OK(h, "marginal_kernel");
  return SISUA_OK;
}

extern "C" int sisua_adam_step(sisua_handle h, float lr, float beta1, float beta2, float eps_hat, float clipnorm,
                               float grad_scale, int64_t t, void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->P || !h->Gd || !h->M || !h->V) SET_ERR(h, SISUA_ERR_STATE, "adam_step needs params, grads, m and v bound");
  sec_begin(h, st, SEC_ADAM);
  CUDA_OK(h, cudaMemsetAsync(h->sq, 0, kMaxSegments * sizeof(double), st));
  dim3 grid((unsigned)((h->total_floats + kAdamChunk - 1) / kAdamChunk));
  ++h->launches;
  grad_sqnorm_kernel<<<grid, 256, 0, st>>>(h->Gd, h->seg, h->total_floats, h->sq, h->d_step, (long long)t, lr, beta1, beta2, h->d_lr_t);
  ++h->launches;
  adam_kernel<<<grid, 256, 0, st>>>(h->P, h->Gd, h->M, h->V, h->seg, h->total_floats, h->sq, h->d_lr_t, beta1, beta2, eps_hat,
                                     clipnorm, h->cfg.clip_mode, grad_scale);
  LAUNCH_OK(h, "adam");
  sec_end(h, st, SEC_ADAM);
  return SISUA_OK;
}

// ---- data-parallel optimiser step as one kernel over NVLink peer memory (dp_adam.cuh) -------------------------------
// The host allocates four SYMMETRIC buffers per rank (same size on every rank, peer-mapped; torch's symmetric memory or
// cudaIpc / VMM -- the library only sees pointers): the flat gradient buffer, the flat parameter buffer (both also bound
// through sisua_bind_buffers), a table of kMaxRanks x 48 doubles and 3 x 8 flag words (zero-initialised).  peer_*[r] is
// rank r's copy as mapped into THIS process.  grid: CTAs of the exchange kernel (0 = one per SM); all of them must be
// resident at once, and every rank must use the same value.
extern "C" int sisua_dp_bind(sisua_handle h, int rank, int world, void* const* peer_grads, void* const* peer_params,
                             void* const* peer_sq, void* const* peer_flags, int grid) {
  if (!h || !peer_grads || !peer_params || !peer_sq || !peer_flags) return SISUA_ERR_INVALID;
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) SET_ERR(h, SISUA_ERR_INVALID, "dp_bind: rank %d / world %d (at most %d ranks)", rank, world, kMaxRanks);
  if (!h->M || !h->V) SET_ERR(h, SISUA_ERR_STATE, "dp_bind needs bound optimiser state");
  if (peer_grads[rank] != (void*)h->Gd || peer_params[rank] != (void*)h->P)
    SET_ERR(h, SISUA_ERR_INVALID, "dp_bind: this rank's symmetric gradient / parameter buffers must be the ones bound with sisua_bind_buffers");
  memset(&h->dp, 0, sizeof(h->dp));
  h->dp.rank = rank; h->dp.world = world;
  for (int r = 0; r < world; ++r) {
    h->dp.grads[r] = (float*)peer_grads[r]; h->dp.params[r] = (float*)peer_params[r];
    h->dp.sqp[r] = (double*)peer_sq[r]; h->dp.flags[r] = (unsigned int*)peer_flags[r];
  }
  unsigned int* loc = nullptr; double* sql = nullptr;
  int rc;
  if ((rc = ws_alloc(h, &loc, 8)) != SISUA_OK || (rc = ws_alloc(h, &sql, kMaxSegments)) != SISUA_OK) return rc;
  CUDA_OK(h, cudaMemset(loc, 0, 8 * sizeof(unsigned int)));
  CUDA_OK(h, cudaMemset(sql, 0, kMaxSegments * sizeof(double)));
  h->dp.local = loc; h->dp.sq_local = sql;
  h->dp_grid = grid > 0 ? grid : h->num_sms;
  int per_sm = 0;
  CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_adam_kernel, 256, 0));
  if (h->dp_grid > per_sm * h->num_sms) SET_ERR(h, SISUA_ERR_INVALID, "dp_bind: grid %d cannot be resident at once (%d CTAs fit)", h->dp_grid, per_sm * h->num_sms);
  h->dp_bound = true;
  return SISUA_OK;
}

// Replaces all-reduce + Adam.apply_gradients for data-parallel training: reduce-scatter over peer memory, clipnorm, Adam on
// this rank's shard, all-gather of the updated parameters -- one kernel, graph-capturable.  Every rank must call it once
// per step with the same arguments.  Optimiser state (m, v) is maintained for this rank's shard only.
extern "C" int sisua_adam_step_dp(sisua_handle h, float lr, float beta1, float beta2, float eps_hat, float clipnorm, int64_t t,
                                  void* stream) {
  if (!h) return SISUA_ERR_INVALID;
  if (!h->dp_bound) SET_ERR(h, SISUA_ERR_STATE, "adam_step_dp needs sisua_dp_bind");
  cudaStream_t st = (cudaStream_t)stream;
  sec_begin(h, st, SEC_ADAM);
  DpArgs a = h->dp;
  a.m = h->M; a.v = h->V; a.st = h->seg; a.total = h->total_floats; a.step = h->d_step; a.step_override = (long long)t;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps_hat = eps_hat; a.clipnorm = clipnorm; a.clip_mode = h->cfg.clip_mode;
  ++h->launches;
  dp_adam_kernel<<<h->dp_grid, 256, 0, st>>>(a);
  LAUNCH_OK(h, "dp_adam_kernel");
  sec_end(h, st, SEC_ADAM);
  return SISUA_OK;
}

// [shard_begin, shard_end) of the flat buffers whose optimiser state this rank maintains under sisua_adam_step_dp
extern "C" int sisua_dp_shard(sisua_handle h, int64_t* begin, int64_t* end) {
  if (!h || !begin || !end) return SISUA_ERR_INVALID;
  if (!h->dp_bound) SET_ERR(h, SISUA_ERR_STATE, "dp_shard needs sisua_dp_bind");
  const long long n4 = (h->total_floats + 3) / 4, per4 = (n4 + h->dp.world - 1) / h->dp.world;
  *begin = (int64_t)h->dp.rank * per4 * 4;
  *end = h->dp.rank == h->dp.world - 1 ? h->total_floats : std::min<long long>(h->total_floats, *begin + per4 * 4);
  return SISUA_OK;
}

// tcgen05 descriptor self-test (tests only): D[128,N] = A[128,K] . B[N,K]^T, fp16 operands, fp32 accumulate.
extern "C" int sisua_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn_major, int b_mn_major,
                                 void* stream) {
#ifdef SISUA_WITH_TC
  if (N % 16 != 0 || N < 16 || N > 256 || K % 16 != 0 || K < 16 || K > 256) return SISUA_ERR_INVALID;
  size_t smem = (size_t)(128 + N) * K * 2;
  if (cudaFuncSetAttribute(tc::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return SISUA_ERR_CUDA;
  tc::tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, a_mn_major, b_mn_major);
  return cudaGetLastError() == cudaSuccess ? SISUA_OK : SISUA_ERR_CUDA;
#else
  return SISUA_ERR_UNSUPPORTED;
#endif
}

// widen uint16 counts (what the host pipeline ships over PCIe for integer count matrices) to the fp32 layout
__global__ void __launch_bounds__(256) unpack_u16_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    uint4 v = *reinterpret_cast<const uint4*>(src + i);
    float4 a = make_float4((float)(v.x & 0xffffu), (float)(v.x >> 16), (float)(v.y & 0xffffu), (float)(v.y >> 16));
    float4 b = make_float4((float)(v.z & 0xffffu), (float)(v.z >> 16), (float)(v.w & 0xffffu), (float)(v.w >> 16));
    *reinterpret_cast<float4*>(dst + i) = a;
    *reinterpret_cast<float4*>(dst + i + 4) = b;
  } else {
    for (; i < n; ++i) dst[i] = (float)src[i];
  }
}

// CSR minibatch (row pointers, uint16 column ids, uint16 counts) -> dense fp32 [rows, G]: one warp per row zeroes the
// row with 16-byte stores, then scatters its non-zeros.  Single-cell count matrices are 70-96 % zeros
// (description/dataset.html:32,158,188), so this is what the host pipeline ships over PCIe.
__global__ void __launch_bounds__(256) unpack_csr_kernel(const int* __restrict__ indptr, const uint16_t* __restrict__ cols,
                                                         const uint16_t* __restrict__ vals, float* __restrict__ dst, int rows, int G) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* row = dst + (size_t)warp * G;
  if ((G & 3) == 0) {
    float4* r4 = reinterpret_cast<float4*>(row);
    for (int i = lane; i < G / 4; i += 32) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int i = lane; i < G; i += 32) row[i] = 0.f;
  }
  __syncwarp();
  const int b = indptr[warp], e = indptr[warp + 1];
  for (int i = b + lane; i < e; i += 32) {
    const int c = cols[i];
    if (c < G) row[c] = (float)vals[i];        // a batch built for a wider matrix must not write outside its row
  }
}

// same, into a uint16 matrix (what sisua_train_step_gather_u16 reads): half the bytes written here and read by the step
__global__ void __launch_bounds__(256) unpack_csr_u16_kernel(const int* __restrict__ indptr, const uint16_t* __restrict__ cols,
                                                             const uint16_t* __restrict__ vals, uint16_t* __restrict__ dst, int rows, int G) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  uint16_t* row = dst + (size_t)warp * G;
  if ((G & 7) == 0) {
    uint4* r4 = reinterpret_cast<uint4*>(row);
    for (int i = lane; i < G / 8; i += 32) r4[i] = make_uint4(0u, 0u, 0u, 0u);
  } else {
    for (int i = lane; i < G; i += 32) row[i] = 0;
  }
  __syncwarp();
  const int b = indptr[warp], e = indptr[warp + 1];
  for (int i = b + lane; i < e; i += 32) {
    const int c = cols[i];
    if (c < G) row[c] = vals[i];
  }
}

extern "C" int sisua_unpack_counts_csr_u16(sisua_handle h, const int32_t* indptr, const uint16_t* cols, const uint16_t* vals,
                                           uint16_t* dst, int rows, void* stream) {
  if (!h || !indptr || !dst || rows < 0) return SISUA_ERR_INVALID;
  if (h->cfg.n_genes > 65536) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "unpack_counts_csr_u16: uint16 column ids need n_genes <= 65536");
  if (reinterpret_cast<uintptr_t>(dst) & 15) SET_ERR(h, SISUA_ERR_INVALID, "unpack_counts_csr_u16: dst must be 16-byte aligned");
  ++h->launches;
  unpack_csr_u16_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(indptr, cols, vals, dst, rows, h->cfg.n_genes);
  LAUNCH_OK(h, "unpack_csr_u16_kernel");
  return SISUA_OK;
}

// Packed CSR ("delta-8"): one 16-bit word per stored entry, low byte = column advance from the previous entry of the row
// (the first from column 0), high byte = count.  An advance of 255 with count 0 is a pure skip (gaps >= 255); count 255
// is an escape whose real value is the next unread element of `big` for that row (big_ptr[row] .. big_ptr[row+1]).
// Half the PCIe bytes of the (uint16 column, uint16 count) form.  One warp per row: zero-fill, then 32 entries at a time
// with a warp scan over the advances and a ballot over the escapes.
__global__ void __launch_bounds__(256) unpack_csr8_u16_kernel(const int* __restrict__ indptr, const int* __restrict__ big_ptr,
                                                              const uint16_t* __restrict__ ents, const uint16_t* __restrict__ big,
                                                              uint16_t* __restrict__ dst, int rows, int G) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  uint16_t* row = dst + (size_t)warp * G;
  if ((G & 7) == 0) {
    uint4* r4 = reinterpret_cast<uint4*>(row);
    for (int i = lane; i < G / 8; i += 32) r4[i] = make_uint4(0u, 0u, 0u, 0u);
  } else {
    for (int i = lane; i < G; i += 32) row[i] = 0;
  }
  __syncwarp();
  const int b = indptr[warp], e = indptr[warp + 1];
  int col0 = 0, big0 = big_ptr[warp];
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t w = i < e ? ents[i] : 0u;
    int adv = (int)(w & 0xffu);
    const uint32_t v8 = w >> 8;
    // inclusive scan of the advances: this entry's column
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, adv, d);
      if (lane >= d) adv += o;
    }
    const int col = col0 + adv;
    const unsigned esc = __ballot_sync(0xffffffffu, i < e && v8 == 255u);
    uint32_t val = v8;
    if (v8 == 255u && i < e) val = big[big0 + __popc(esc & ((1u << lane) - 1u))];
    if (i < e && val != 0u && col < G) row[col] = (uint16_t)val;
    col0 += __shfl_sync(0xffffffffu, adv, 31);
    big0 += __popc(esc);
  }
}

extern "C" int sisua_unpack_counts_csr8_u16(sisua_handle h, const int32_t* indptr, const int32_t* big_ptr, const uint16_t* ents,
                                            const uint16_t* big, uint16_t* dst, int rows, void* stream) {
  if (!h || !indptr || !big_ptr || !dst || rows < 0) return SISUA_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(dst) & 15) SET_ERR(h, SISUA_ERR_INVALID, "unpack_counts_csr8_u16: dst must be 16-byte aligned");
  ++h->launches;
  unpack_csr8_u16_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(indptr, big_ptr, ents, big, dst, rows, h->cfg.n_genes);
  LAUNCH_OK(h, "unpack_csr8_u16_kernel");
  return SISUA_OK;
}

extern "C" int sisua_unpack_counts_csr(sisua_handle h, const int32_t* indptr, const uint16_t* cols, const uint16_t* vals,
                                       float* dst, int rows, void* stream) {
  if (!h || !indptr || !dst || rows < 0) return SISUA_ERR_INVALID;
  if (h->cfg.n_genes > 65536) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "unpack_counts_csr: uint16 column ids need n_genes <= 65536");
  if (reinterpret_cast<uintptr_t>(dst) & 15) SET_ERR(h, SISUA_ERR_INVALID, "unpack_counts_csr: dst must be 16-byte aligned");
  ++h->launches;
  unpack_csr_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(indptr, cols, vals, dst, rows, h->cfg.n_genes);
  LAUNCH_OK(h, "unpack_csr_kernel");
  return SISUA_OK;
}

extern "C" int sisua_unpack_counts_u16(sisua_handle h, const uint16_t* src, float* dst, int64_t n, void* stream) {
  if (!h || !src || !dst || n < 0) return SISUA_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    SET_ERR(h, SISUA_ERR_INVALID, "unpack_counts_u16: pointers must be 16-byte aligned");
  long long blocks = (n / 8 + 255) / 256 + 1;
  ++h->launches;
  unpack_u16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, (long long)n);
  LAUNCH_OK(h, "unpack_u16_kernel");
  return SISUA_OK;
}

// ---- artificial corruption of a count matrix (the imputation benchmark's input, sisua/data/utils.py:168-228) on the GPU
// Entry (row, col) > 0 is selected with probability `dropout` (Philox word 0 of call (row, col, 0, 0x200) against
// floor(dropout 2^32)); a selected count n becomes Binomial(n, retain) ('binomial') or n Bernoulli(retain) ('uniform').
// The binomial is the exact sum of n Bernoulli trials: trial t takes 16 bits of Philox call (row, col, 1 + t / 8, 0x200)
// (low half of word (t % 8) / 2 first) and succeeds when they are below floor(retain 65536) -- integers only, so
// oracle/philox.py:corrupt_counts reproduces the matrix bit for bit.  Counts in the thousands cost n / 8 Philox calls;
// they are a vanishing share of a single-cell matrix.
__global__ void __launch_bounds__(256) corrupt_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int cols,
                                                      long long ld_src, long long ld_dst, uint32_t sel_thr, uint32_t keep_thr,
                                                      int uniform, uint32_t seed_lo, uint32_t seed_hi) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols; const int c = (int)(i - r * cols);
    const float x = src[r * ld_src + c];
    float y = x;
    if (x > 0.f) {
      const uint2 key = make_uint2(seed_lo, seed_hi);
      const uint4 s = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)c, 0u, 0x200u), key);
      if (s.x < sel_thr) {
        const uint32_t nn = (uint32_t)x, trials = uniform ? 1u : nn;
        uint32_t k = 0;
        for (uint32_t t0 = 0; t0 < trials; t0 += 8) {
          const uint4 w = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)c, 1u + t0 / 8, 0x200u), key);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t u = (j & 1) ? (ww[j >> 1] >> 16) : (ww[j >> 1] & 0xffffu);
            if (t0 + j < trials && u < keep_thr) ++k;
          }
        }
        y = uniform ? (k ? x : 0.f) : (float)k;
      }
    }
    dst[r * ld_dst + c] = y;
  }
}

extern "C" int sisua_corrupt_counts(sisua_handle h, const float* src, float* dst, int64_t rows, int cols, int64_t ld_src,
                                    int64_t ld_dst, float dropout, float retain_rate, int distribution, uint64_t seed, void* stream) {
  if (!h || !src || !dst || rows < 0 || cols <= 0 || ld_src < cols || ld_dst < cols) return SISUA_ERR_INVALID;
  if (!(dropout >= 0.f && dropout < 1.f) || !(retain_rate >= 0.f && retain_rate <= 1.f) || (distribution != 0 && distribution != 1))
    SET_ERR(h, SISUA_ERR_INVALID, "corrupt_counts: dropout in [0, 1), retain_rate in [0, 1], distribution 0 (binomial) or 1 (uniform)");
  if (rows >= (1ll << 32)) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "corrupt_counts: more than 2^32 rows");
  if (rows == 0) return SISUA_OK;
  const uint32_t sel_thr = (uint32_t)std::min(4294967295.0, floor((double)dropout * 4294967296.0));
  const uint32_t keep_thr = (uint32_t)floor((double)retain_rate * 65536.0);
  const long long n = (long long)rows * cols;
  ++h->launches;
  corrupt_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 16ll * h->num_sms), 256, 0, (cudaStream_t)stream>>>(
      src, dst, rows, cols, ld_src, ld_dst, sel_thr, keep_thr, distribution, (uint32_t)seed, (uint32_t)(seed >> 32));
  LAUNCH_OK(h, "corrupt_kernel");
  return SISUA_OK;
}

// ---- host-buffer train step ---------------------------------------------------------------------
static int host_stage_init(sisua_model* h) {
  auto& S = h->hs;
  if (S.ready) return SISUA_OK;
  const sisua_step_config& c = h->cfg;
  const size_t R = c.max_batch;
  CUDA_OK(h, cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
  int rc;
  for (int s = 0; s < 2; ++s) {
    CUDA_OK(h, cudaEventCreateWithFlags(&S.filled[s], cudaEventDisableTiming));
    CUDA_OK(h, cudaEventCreateWithFlags(&S.consumed[s], cudaEventDisableTiming));
    if ((rc = ws_alloc(h, &S.x[s], R * c.n_genes)) || (rc = ws_alloc(h, &S.x16[s], R * c.n_genes + 8)) ||
        (rc = ws_alloc(h, &S.eps_z[s], R * c.n_latent)) || (rc = ws_alloc(h, &S.eps_l[s], R)) ||
        (rc = ws_alloc(h, &S.lib[s], R * 2)) || (rc = ws_alloc(h, &S.mask[s], R)) ||
        (rc = ws_alloc(h, &S.y[s], R * std::max(1, c.n_proteins)))) return rc;
  }
  if ((rc = ws_alloc(h, &S.terms, 5 * R)) || (rc = ws_alloc(h, &S.loss, 1))) return rc;
  S.ready = true;
  return SISUA_OK;
}

extern "C" int sisua_train_step_host(sisua_handle h, const sisua_host_batch* hb, uint64_t seed, int64_t step, float* host_loss,
                                     float* host_terms, void* stream) {
  if (!h || !hb) return SISUA_ERR_INVALID;
  const sisua_step_config& c = h->cfg;
  const int B = hb->B, G = c.n_genes;
  if (B < 1 || B > c.max_batch) SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: B=%d outside [1, max_batch=%d]", B, c.max_batch);
  if (hb->format < SISUA_HOST_F32 || hb->format > SISUA_HOST_CSR) SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: unknown format %d", hb->format);
  if (hb->format == SISUA_HOST_CSR) {
    if (!hb->indptr || hb->nnz < 0 || (hb->nnz > 0 && (!hb->cols || !hb->vals))) SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: incomplete CSR batch");
    if (G > 65536) SET_ERR(h, SISUA_ERR_UNSUPPORTED, "train_step_host: uint16 column ids need n_genes <= 65536");
    // the row pointers are host memory: check them before anything is enqueued (they drive a device-side scatter)
    if (hb->indptr[0] != 0 || (int64_t)hb->indptr[B] != hb->nnz) SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: indptr[0] = %d, indptr[B] = %d but nnz = %lld", hb->indptr[0], hb->indptr[B], (long long)hb->nnz);
    for (int i = 0; i < B; ++i)
      if (hb->indptr[i + 1] < hb->indptr[i]) SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: indptr is not monotone at row %d", i);
  } else if (!hb->x) {
    SET_ERR(h, SISUA_ERR_INVALID, "train_step_host: x is NULL");
  }
  int rc = host_stage_init(h);
  if (rc != SISUA_OK) return rc;
  auto& S = h->hs;
  cudaStream_t st = (cudaStream_t)stream;
  const int s = (int)(S.calls++ & 1);
  // ---- stage on the copy stream (after the kernels that last read this slot)
  if (S.calls > 2) CUDA_OK(h, cudaStreamWaitEvent(S.copy, S.consumed[s], 0));
  const size_t n = (size_t)B * G;
  if (hb->format == SISUA_HOST_F32) {
    CUDA_OK(h, cudaMemcpyAsync(S.x[s], hb->x, n * sizeof(float), cudaMemcpyHostToDevice, S.copy));
  } else if (hb->format == SISUA_HOST_U16) {
    CUDA_OK(h, cudaMemcpyAsync(S.x16[s], hb->x, n * sizeof(uint16_t), cudaMemcpyHostToDevice, S.copy));
  } else {
    if (S.csr_cap[s] < (size_t)hb->nnz || !S.indptr[s]) {      // grow (rare): nothing on the copy stream may still use the old buffers
      CUDA_OK(h, cudaStreamSynchronize(S.copy));
      CUDA_OK(h, cudaStreamSynchronize(st));
      cudaFree(S.cols[s]); cudaFree(S.vals[s]);
      S.cols[s] = S.vals[s] = nullptr;
      const size_t cap = (size_t)hb->nnz + (size_t)hb->nnz / 4 + 1024;
      CUDA_OK(h, cudaMalloc((void**)&S.cols[s], cap * sizeof(uint16_t)));
      CUDA_OK(h, cudaMalloc((void**)&S.vals[s], cap * sizeof(uint16_t)));
      if (!S.indptr[s]) CUDA_OK(h, cudaMalloc((void**)&S.indptr[s], ((size_t)c.max_batch + 1) * sizeof(int32_t)));
      S.csr_cap[s] = cap;
    }
    CUDA_OK(h, cudaMemcpyAsync(S.indptr[s], hb->indptr, ((size_t)B + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, S.copy));
    if (hb->nnz > 0) {
      CUDA_OK(h, cudaMemcpyAsync(S.cols[s], hb->cols, (size_t)hb->nnz * sizeof(uint16_t), cudaMemcpyHostToDevice, S.copy));
      CUDA_OK(h, cudaMemcpyAsync(S.vals[s], hb->vals, (size_t)hb->nnz * sizeof(uint16_t), cudaMemcpyHostToDevice, S.copy));
    }
  }
  if (hb->eps_z) CUDA_OK(h, cudaMemcpyAsync(S.eps_z[s], hb->eps_z, (size_t)B * c.n_latent * sizeof(float), cudaMemcpyHostToDevice, S.copy));
  if (hb->eps_l) CUDA_OK(h, cudaMemcpyAsync(S.eps_l[s], hb->eps_l, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, S.copy));
  if (hb->library) CUDA_OK(h, cudaMemcpyAsync(S.lib[s], hb->library, (size_t)B * 2 * sizeof(float), cudaMemcpyHostToDevice, S.copy));
  if (hb->mask) CUDA_OK(h, cudaMemcpyAsync(S.mask[s], hb->mask, (size_t)B, cudaMemcpyHostToDevice, S.copy));
  if (hb->y && c.n_proteins > 0) CUDA_OK(h, cudaMemcpyAsync(S.y[s], hb->y, (size_t)B * c.n_proteins * sizeof(float), cudaMemcpyHostToDevice, S.copy));
  CUDA_OK(h, cudaEventRecord(S.filled[s], S.copy));
  // ---- compute stream: widen the counts, run the step, read the loss back
  CUDA_OK(h, cudaStreamWaitEvent(st, S.filled[s], 0));
  if (hb->format == SISUA_HOST_U16) {
    rc = sisua_unpack_counts_u16(h, S.x16[s], S.x[s], (int64_t)n, stream);
  } else if (hb->format == SISUA_HOST_CSR) {
    rc = sisua_unpack_counts_csr(h, S.indptr[s], S.cols[s], S.vals[s], S.x[s], B, stream);
  }
  if (rc != SISUA_OK) return rc;
  rc = sisua_train_step(h, S.x[s], (hb->y && c.n_proteins > 0) ? S.y[s] : nullptr, hb->library ? S.lib[s] : nullptr,
                        hb->mask ? S.mask[s] : nullptr, hb->eps_z ? S.eps_z[s] : nullptr, hb->eps_l ? S.eps_l[s] : nullptr, B, seed,
                        step, S.terms, S.loss, stream);
  if (rc != SISUA_OK) return rc;
  CUDA_OK(h, cudaEventRecord(S.consumed[s], st));
  if (host_loss) CUDA_OK(h, cudaMemcpyAsync(host_loss, S.loss, sizeof(float), cudaMemcpyDeviceToHost, st));
  if (host_terms) CUDA_OK(h, cudaMemcpyAsync(host_terms, S.terms, (size_t)5 * B * sizeof(float), cudaMemcpyDeviceToHost, st));
  return SISUA_OK;
}

// Multi-GPU overlap hook: `cuda_event` (a cudaEvent_t owned by the caller, or NULL to clear) is recorded on the
// step's stream as soon as the gradients of the output heads (out.W, out.b: ~3/4 of all gradient bytes) are final,
// so the host can start their all-reduce on another stream while the rest of the backward pass runs.
extern "C" int sisua_set_grad_ready_event(sisua_handle h, void* cuda_event) {
  if (!h) return SISUA_ERR_INVALID;
  h->ev_out_grads = (cudaEvent_t)cuda_event;
  return SISUA_OK;
}

// The fused kernels hand d llk / d (head outputs) to the tensor cores as fp16 tiles; its entries are bounded by the
// largest count (or predicted mean), and fp16 ends at 65504.  Declaring a bound above 2^15 makes the kernels scale the
// tiles by a power of two (and the results back), instead of overflowing to inf -> NaN gradients.
extern "C" int sisua_nonfinite_flag(sisua_handle h, int reset) {
  if (!h || !h->nf_host) return -1;
  const int v = *(volatile int*)h->nf_host;
  if (reset) *(volatile int*)h->nf_host = 0;
  return v != 0;
}

extern "C" int sisua_set_count_bound(sisua_handle h, float max_count) {
  if (!h || !(max_count >= 0.f)) return SISUA_ERR_INVALID;
  float s = 1.0f;
  while (max_count * s > 32768.f && s > 1e-30f) s *= 0.5f;
  h->gscale = s;
  return SISUA_OK;
}

// Seed (and call index) of the reparameterisation noise sisua_infer draws in-kernel when eps_z / eps_l are NULL: call k
// after this uses Philox(seed; row, column, call_index + k, stream), so a fixed seed makes predict reproducible.
extern "C" int sisua_set_infer_seed(sisua_handle h, uint64_t seed, int64_t call_index) {
  if (!h || call_index < 0) return SISUA_ERR_INVALID;
  h->infer_seed = seed; h->infer_calls = call_index;
  return SISUA_OK;
}

// device-side optimiser step counter (number of Adam steps applied so far)
extern "C" int sisua_set_step(sisua_handle h, int64_t t, void* stream) {
  if (!h || t < 0) return SISUA_ERR_INVALID;
  long long v = (long long)t;
  CUDA_OK(h, cudaMemcpyAsync(h->d_step, &v, sizeof(v), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CUDA_OK(h, cudaStreamSynchronize((cudaStream_t)stream));     // `v` lives on this stack frame
  return SISUA_OK;
}

extern "C" int64_t sisua_launch_count(sisua_handle h) { return h ? h->launches : -1; }

extern "C" int sisua_profile_enable(sisua_handle h, int on) {
  if (!h) return SISUA_ERR_INVALID;
  h->profiling = on != 0;
  for (int i = 0; i < 8; ++i) h->sec_used[i] = 0;
  return SISUA_OK;
}

// Sum of the device time (ms) spent in each section since profile_enable(1); synchronises the device.
// ms_out[6] = enc_first, mid_fwd, out_heads, mid_bwd, enc_first_bwd, adam; counts_out[6] = timed intervals.
extern "C" int sisua_profile_read(sisua_handle h, float* ms_out, int* counts_out) {
  if (!h || !ms_out) return SISUA_ERR_INVALID;
  CUDA_OK(h, cudaDeviceSynchronize());
  for (int i = 0; i < SEC_COUNT; ++i) {
    double tot = 0.0;
    for (size_t j = 0; j < h->sec_used[i]; ++j) {
      float ms = 0.f;
      CUDA_OK(h, cudaEventElapsedTime(&ms, h->sec_events[i][j].first, h->sec_events[i][j].second));
      tot += ms;
    }
    ms_out[i] = (float)tot;
    if (counts_out) counts_out[i] = (int)h->sec_used[i];
  }
  return SISUA_OK;
}
