/* sisua_b200 — C ABI of the B200-native SISUA ELBO train / infer step.
 *
 * The reference (trungnt13/sisua) is pure Python and has no FFI of its own: the operator boundary of
 * its hot path is the Python class surface of sisua.models.SingleCellModel
 * (sisua/models/single_cell_model.py:67-306), whose arithmetic is delegated to odin-ai / TensorFlow.
 * This library is what that class binds instead (via ctypes); each entry point
 * names the reference call it replaces.  Conventions: plain pointers and sizes, every data pointer is
 * caller-owned DEVICE memory (fp32 unless stated), all work is enqueued on the caller's CUDA stream
 * with no hidden synchronisation, integer return codes (0 = ok) and sisua_last_error() for the text.
 * One handle per GPU per host thread.
 */
#ifndef SISUA_B200_H_
#define SISUA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SISUA_OK = 0, SISUA_ERR_INVALID = 1, SISUA_ERR_CUDA = 2, SISUA_ERR_UNSUPPORTED = 3, SISUA_ERR_STATE = 4 };

enum { SISUA_MODEL_VAE = 0, SISUA_MODEL_SCVI = 1, SISUA_MODEL_DCA = 2, SISUA_MODEL_SISUA = 3 };
/* gene-count likelihood: 'zinbd' / 'nbd' = (zero-inflated) NB by mean and inverse dispersion (odin-ai NegativeBinomialDisp,
 * configs/base.yaml:32-34); 'zinb' / 'nb' = TFP NegativeBinomial(total_count = exp(.), logits = .)
 * (tests/test_singlecell_models.py:60-80): the same likelihood with other links */
enum { SISUA_XDIST_ZINBD = 0, SISUA_XDIST_NBD = 1, SISUA_XDIST_ZINB = 2, SISUA_XDIST_NB = 3 };
enum { SISUA_YDIST_NB = 0, SISUA_YDIST_NBD = 1 };
enum { SISUA_ACT_SOFTPLUS = 0, SISUA_ACT_SOFTPLUS1 = 1, SISUA_ACT_SOFTPLUS_P1 = 2, SISUA_ACT_EXP = 3,
       SISUA_ACT_IDENTITY = 4 };
enum { SISUA_GEMM_FP32_UNFUSED = 0,   /* CUDA-core fp32 GEMMs, decoder output materialised (cross-check path) */
       SISUA_GEMM_TC_3XFP16 = 1 };    /* fused tcgen05 kernels: 3xFP16 compensated forward, fp16 gradient GEMMs */

/* Everything the step needs to know; built on the host from RVmeta / NetConf / configs/base.yaml
 * (sisua/train.py:71-106).  Field-for-field identical to sisua_b200.config.StepConfig. */
typedef struct sisua_step_config {
  int32_t model_kind;        /* SISUA_MODEL_*  (sisua/models/{vae,scvi,dca}.py)                      */
  int32_t n_genes;           /* G: event_shape of the transcriptomic RVmeta (train.py:79-89)         */
  int32_t n_proteins;        /* P: label RV of SISUA (vae.py:40), 0 otherwise                        */
  int32_t n_latent;          /* z: RVmeta(10,'diag') single_cell_model.py:77                         */
  int32_t n_hidden;          /* NetConf units, 64 (single_cell_model.py:78-81)                       */
  int32_t n_enc_layers, n_dec_layers, n_encl_layers;
  int32_t batchnorm;         /* NetConf(batchnorm=True)                                              */
  int32_t log_norm;          /* log1p on the counts (single_cell_model.py:126-131)                   */
  int32_t x_dist, y_dist;    /* 'zinbd'|'nbd' ; 'nb'|'nbd' (configs/base.yaml:32-40)                 */
  int32_t mean_act, disp_act, scale_act;   /* Q1 of SURVEY.md section 8a                             */
  int32_t scvi_reapply_act;  /* Q2 */
  int32_t mask_norm;         /* Q3: 0 mean over all cells, 1 over labelled cells                     */
  int32_t clip_mode;         /* Q5: 0 per-variable clipnorm (Keras), 1 global norm                   */
  int32_t gemm_mode;         /* SISUA_GEMM_*                                                         */
  int32_t max_batch;         /* rows the private workspace is sized for (S*B for inference)          */
  int32_t latent_linear;     /* deterministic (DCA) latent without the ReLU: RVmeta(.., 'linear'), dca.py:22-27 */
  float bn_eps, bn_momentum; /* Keras BatchNormalization defaults 1e-3 / 0.99                        */
  float input_dropout, enc_dropout, dec_dropout, encl_dropout;
  float beta, alpha;         /* KL weight (single_cell_model.py:83), label weight (base.yaml:6)      */
  float clip_library;        /* scvi.py:46,117                                                        */
} sisua_step_config;

typedef struct sisua_param_desc {
  char name[32];
  int64_t offset;            /* floats from the start of the flat buffer */
  int32_t rows, cols, ld;    /* cols == 0 for vectors; ld = leading dimension in floats */
  int32_t kind;              /* 0 weight, 1 bias, 2 gamma, 3 beta */
} sisua_param_desc;

typedef struct sisua_model* sisua_handle;

/* Replaces SingleCellModel.__init__ -> BetaVAE/NetConf/DenseDistribution construction
 * (single_cell_model.py:74-97, scvi.py:33-86).  Allocates only the private workspace. */
int sisua_create(const sisua_step_config* cfg, int device, sisua_handle* out);
int sisua_destroy(sisua_handle h);

/* Names / offsets of every trainable tensor inside ONE flat fp32 buffer (what keras `model.weights`
 * enumerates for save_weights/load_weights, single_cell_model.py:283-306). *n: in = capacity, out = count. */
int sisua_param_layout(sisua_handle h, sisua_param_desc* out, int* n, int64_t* total_floats);

/* Caller-owned device buffers: params/grads/adam_m/adam_v [total_floats]; bn_moving [n_bn,2,H]. */
int sisua_bind_buffers(sisua_handle h, float* params, float* grads, float* adam_m, float* adam_v, float* bn_moving);

/* One minibatch forward + backward: replaces the body of odin-ai's `optimize` under GradientTape —
 * train_steps -> encode -> decode -> _elbo -> tape.gradient (call site
 * single_cell_model.py:236).  x [B,G] counts; y [B,P] or NULL; library [B,2] (mean,var) or NULL;
 * mask [B] bytes or NULL; eps_z [B,z] and eps_l [B] (scVI): injected reparameterisation noise (tests), or NULL to draw
 * it in-kernel: standard normals by Box-Muller on Philox4x32-10(seed; row, column / 4, step, stream 0x100 (z) / 0x101
 * (library)), regenerated by the backward pass and reproduced by oracle/philox.py.  Outputs: terms [5,B] =
 * (elbo | llk_x | llk_y | kl_z | kl_l), loss [1].  Gradients land in the bound `grads` buffer and
 * BN moving statistics are updated.  Dropout masks (NetConf input_dropout / dropout) are the pure function
 * Philox4x32-10(seed; row, col/8, step, stream), 16 bits per column, regenerated in forward and backward; step < 0 makes the kernels
 * read the device-side optimiser step counter (+1), so a captured CUDA graph of train_step + adam_step can be replayed.  The optimiser is a separate call so the host can all-reduce
 * `grads` across GPUs in between. */
int sisua_train_step(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                     const float* eps_z, const float* eps_l, int B, uint64_t seed, int64_t step, float* terms,
                     float* loss, void* stream);

/* The same step on a minibatch given as ROW INDICES into matrices that stay resident in HBM -- what the reference's input
 * pipeline produces every step (shuffle -> batch, sisua/data/_single_cell_base.py:593-601): x_all [N,G], y_all [N,P],
 * library_all [N,2], mask_all [N], rows [B] int32 (device memory).  The fused kernels read the count rows through the index;
 * no gathered copy of the minibatch is written.  rows == NULL behaves like sisua_train_step. */
int sisua_train_step_gather(sisua_handle h, const float* x_all, const float* y_all, const float* library_all,
                            const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                            uint64_t seed, int64_t step, float* terms, float* loss, void* stream);

/* sisua_train_step_gather with the resident count matrix stored as uint16 [N, G] -- exact for count data below 65 536
 * (what sisua/data/utils.py:427-431 keeps as float32): half the HBM of the shard and half the bytes of both streaming
 * reads of a step.  The tcgen05 kernels widen the counts themselves when rows are 16-byte aligned (n_genes % 8 == 0,
 * aligned base) and the heads are the plain (non-scVI) ones; otherwise the minibatch rows are first widened into an fp32
 * staging buffer (same results, one extra pass).  rows may be NULL (rows 0 .. B-1).  Results are identical to the fp32
 * entry point on the widened matrix. */
int sisua_train_step_gather_u16(sisua_handle h, const uint16_t* x_all, const float* y_all, const float* library_all,
                                const uint8_t* mask_all, const int32_t* rows, const float* eps_z, const float* eps_l, int B,
                                uint64_t seed, int64_t step, float* terms, float* loss, void* stream);

/* sisua_unpack_counts_csr into a uint16 [rows, G] matrix (the input type of sisua_train_step_gather_u16): a host pipeline
 * that ships CSR minibatches never materialises them in fp32. */
int sisua_unpack_counts_csr_u16(sisua_handle h, const int32_t* indptr, const uint16_t* cols, const uint16_t* vals, uint16_t* dst,
                                int rows, void* stream);

/* Packed CSR minibatch ("delta-8") -> uint16 [rows, G]: one 16-bit word per stored entry, low byte = column advance from
 * the row's previous entry (the first from column 0), high byte = count.  (advance 255, count 0) is a pure skip for gaps of
 * 255 and more; count 255 escapes to the row's next unread element of `big` (big_ptr [rows + 1] offsets into it).  indptr
 * [rows + 1] offsets into `ents`.  2 bytes per non-zero over PCIe instead of 4 (CSR) or 4 G per cell (float32 rows).
 * sisua_b200/pipeline.py:Csr8Batch builds it. */
int sisua_unpack_counts_csr8_u16(sisua_handle h, const int32_t* indptr, const int32_t* big_ptr, const uint16_t* ents,
                                 const uint16_t* big, uint16_t* dst, int rows, void* stream);

/* dst[b, :] = (float) x_all[rows[b], :] for b < n_rows (rows == NULL: rows 0 .. n_rows-1): feeds the fp32 entry points
 * (sisua_infer*, validation) from a uint16 resident matrix, chunk by chunk. */
int sisua_widen_rows_u16(sisua_handle h, const uint16_t* x_all, const int32_t* rows, float* dst, int n_rows, void* stream);

/* Replaces one `self(**data, training=False, sample_shape=S)` call of SingleCellModel.predict
 * (single_cell_model.py:176-181) plus the parameter tensors the returned distributions hold.
 * eps_z [S,B,z], eps_l [S,B]; terms [5,S*B]; z_loc/z_scale [B,z]; out_mean/out_disp/out_pi [S*B,G]
 * (NB mean = "imputed" mean of posterior.py:210-220, inverse dispersion, dropout logit), y_mean
 * [S*B,P]; any output may be NULL. */
int sisua_infer(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                const float* eps_z, const float* eps_l, int B, int S, float* terms, float* z_loc, float* z_scale,
                float* lib_loc, float* lib_scale, float* out_mean, float* out_disp, float* out_pi, float* y_mean,
                void* stream);

/* sisua_infer with the options the reference's Posterior needs (sisua/analysis/posterior.py:210-220, 919-976), all inside
 * the fused kernels: x_eval [B,G] = counts the log-likelihood is evaluated on while the encoder reads x (NULL: x itself)
 * -- llk of the ORIGINAL counts under the model of the CORRUPTED ones; strip_zi != 0 = likelihood of the count
 * distribution without its zero inflation (the "imputed" distribution); out_mean_avg [B,G] = mean over the S samples of
 * the NB mean (the imputed matrix, posterior.py:986-988); logw [S*B] = log p(z_s) - log q(z_s | x) (+ library latent). */
int sisua_infer_ex(sisua_handle h, const float* x, const float* x_eval, const float* y, const float* library, const uint8_t* mask,
                   const float* eps_z, const float* eps_l, int B, int S, int strip_zi, float* terms, float* z_loc, float* z_scale,
                   float* lib_loc, float* lib_scale, float* out_mean, float* out_disp, float* out_pi, float* y_mean,
                   float* out_mean_avg, float* logw, void* stream);

/* Replaces SingleCellModel.marginal_log_prob(**batch, sample_shape=S) (call site posterior.py:964): importance-weighted
 * bound per cell, mllk [B] = logsumexp_s(llk_x + alpha mask llk_y + log p(z_s) - log q(z_s | x)) - log S, and the
 * per-output entries llk_x [B] (llk_y [B], nullable) = logsumexp_s(llk) - log S.  S*B <= max_batch. */
int sisua_marginal_llk(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                       const float* eps_z, const float* eps_l, int B, int S, float* mllk, float* llk_x, float* llk_y, void* stream);

/* `model(**batch, training=True)` outside fit (single_cell_model.py:178): training-mode forward pass -- BatchNorm batch
 * statistics (folded into the moving ones), dropout masks and noise from (seed, step) -- without gradients or update.
 * Outputs as sisua_infer with S = 1. */
int sisua_forward_train_mode(sisua_handle h, const float* x, const float* y, const float* library, const uint8_t* mask,
                             const float* eps_z, const float* eps_l, int B, uint64_t seed, int64_t step, float* terms,
                             float* z_loc, float* z_scale, float* lib_loc, float* lib_scale, float* out_mean, float* out_disp,
                             float* out_pi, float* y_mean, void* stream);

/* Replaces SingleCellModel.decode(latents) (single_cell_model.py:141-151, scvi.py:108-171): latent samples z [R, z] (and
 * scVI's sampled log-library sizes lib [R]) -> parameters of the output distributions under the moving BatchNorm
 * statistics.  out_mean / out_disp / out_pi [R, G], y_mean [R, P]; any may be NULL. */
int sisua_decode(sisua_handle h, const float* z, const float* lib, int R, float* out_mean, float* out_disp, float* out_pi,
                 float* y_mean, void* stream);

/* Replaces keras Adam.apply_gradients with clipnorm (configs/base.yaml:46-50). grad_scale multiplies
 * the gradients first (1/world_size after a sum all-reduce). t >= 1 sets the step index; t <= 0
 * advances the device-side counter (CUDA-graph friendly). */
int sisua_adam_step(sisua_handle h, float lr, float beta1, float beta2, float eps_hat, float clipnorm,
                    float grad_scale, int64_t t, void* stream);

/* Data-parallel optimiser step as ONE kernel over NVLink / NVSwitch peer memory: reduce-scatter of
 * the flat gradient buffer by peer loads, per-variable clipnorm, Adam on this rank's shard (optimiser state is sharded),
 * all-gather of the updated parameters by peer stores; the ranks meet at device-side barriers in symmetric memory.
 * sisua_dp_bind: peer_*[r] = rank r's SYMMETRIC buffer as mapped into this process -- gradients and parameters
 * [total_floats] (this rank's own must be the ones given to sisua_bind_buffers), a table of 8 x 48 doubles and 3 x 8 flag
 * words (zero-initialised).  grid = CTAs of the kernel (0: one per SM; all must be resident; equal on every rank).
 * sisua_adam_step_dp: every rank calls it once per step after sisua_train_step; no other collective is needed.
 * sisua_dp_shard: the [begin, end) range of the flat buffers whose Adam moments this rank maintains. */
int sisua_dp_bind(sisua_handle h, int rank, int world, void* const* peer_grads, void* const* peer_params, void* const* peer_sq,
                  void* const* peer_flags, int grid);
int sisua_adam_step_dp(sisua_handle h, float lr, float beta1, float beta2, float eps_hat, float clipnorm, int64_t t, void* stream);
int sisua_dp_shard(sisua_handle h, int64_t* begin, int64_t* end);

/* Workspace peeks for tests (device pointers valid until destroy): "d" decoder output [rows,H],
 * "delta1" first-layer pre-activation gradient. Returns NULL for unknown names. */
const float* sisua_debug_buffer(sisua_handle h, const char* name);

int sisua_debug_copy(sisua_handle h, const char* name, float* dst, int64_t n_floats, void* stream);

/* Launch geometry the tcgen05 kernels would use for a train step of B cells (tests assert the pipeline depth they
 * exercise): out[9] = first layer (cell tiles, k-chunks, k-blocks per chunk) | output heads (cell tiles, gene chunks,
 * gene tiles per chunk) | first-layer weight gradient (gene tiles, cell chunks, cell tiles per chunk). */
int sisua_debug_geometry(sisua_handle h, int B, int32_t* out);
/* Tests only: force how many chunks the output-head / first-layer / weight-gradient kernels split their gene or cell
 * walk into (0 = automatic heuristic), so a small batch can run the single-chunk deep pipeline of a full-size one. */
int sisua_debug_force_chunks(sisua_handle h, int out_chunks, int enc_chunks, int bwd_chunks);

/* Declares the largest count (or predicted mean) the step will see.  The fused kernels pass d llk / d (head outputs) to
 * the tensor cores as fp16 tiles whose entries are bounded by that value; above 2^15 they are scaled by a power of two
 * (and the products scaled back) so that nothing overflows fp16.  Default bound: 32768.  Without the declaration a
 * larger count produces inf -> NaN gradients (loud), never a silently clamped one. */
int sisua_set_count_bound(sisua_handle h, float max_count);

/* Artificial corruption of a count matrix on the GPU -- the input of the imputation benchmark (apply_artificial_corruption,
 * sisua/data/utils.py:168-228; Posterior corrupts its test set with it, sisua/analysis/posterior.py:150-170).  src / dst:
 * device fp32 [rows, cols] with row strides ld_src / ld_dst (floats); dst may alias src.  Every positive entry is selected
 * with probability `dropout`; a selected count n becomes Binomial(n, retain_rate) (distribution 0, 'binomial') or
 * n * Bernoulli(retain_rate) (distribution 1, 'uniform').  Draws are Philox4x32-10(seed; row, col, call, 0x200), integer
 * thresholds only: oracle/philox.py:corrupt_counts reproduces the result bit for bit.  Deviation from the reference,
 * which is a NumPy host routine: it corrupts EXACTLY floor(dropout * nnz) entries drawn without replacement from one
 * RandomState stream (sisua_b200.posterior.apply_artificial_corruption restates that on the host); here the selection is
 * independent per entry, so the corrupted share is dropout in expectation. */
int sisua_corrupt_counts(sisua_handle h, const float* src, float* dst, int64_t rows, int cols, int64_t ld_src, int64_t ld_dst,
                         float dropout, float retain_rate, int distribution, uint64_t seed, void* stream);

/* Per-step finiteness of the training loss (the reference's terminate_on_nan callback, configs/base.yaml:59, acts on
 * every step): every training step whose loss sum is not finite sets a sticky word in mapped host memory.  Returns 1 if
 * it is set, 0 if not, -1 on a bad handle; `reset` != 0 clears it.  No stream synchronisation: the word becomes visible
 * when the step that produced it has run, so polling once per step sees a failure within the depth of the launch queue. */
int sisua_nonfinite_flag(sisua_handle h, int reset);

/* Noise of sisua_infer calls that pass eps_z / eps_l == NULL: call k after this draws Philox(seed; row, column / 4,
 * call_index + k, stream 0x100 + 2 s (z) / 0x101 + 2 s (library) for Monte-Carlo sample s). */
int sisua_set_infer_seed(sisua_handle h, uint64_t seed, int64_t call_index);

/* Sets the device-side optimiser step counter (t = Adam steps applied so far); synchronises the stream. */
int sisua_set_step(sisua_handle h, int64_t t, void* stream);

/* Multi-GPU overlap hook: the caller's cudaEvent_t (NULL clears) is recorded on the step's stream as soon as the
 * gradients of the output heads (out.W, out.b — about 3/4 of the gradient bytes) are final, so their all-reduce can
 * run on another stream under the remaining backward pass. */
int sisua_set_grad_ready_event(sisua_handle h, void* cuda_event);

/* Host pipelines ship integer count matrices over PCIe as uint16 (half the bytes of the reference's float32
 * storage, sisua/data/utils.py:427-431); this widens n values to the fp32 layout the step consumes. */
int sisua_unpack_counts_u16(sisua_handle h, const uint16_t* src, float* dst, int64_t n, void* stream);
/* Same for a CSR minibatch: indptr [rows+1] (offsets into cols/vals, starting at 0), uint16 column ids and counts;
 * dst [rows, n_genes] is fully overwritten. */
int sisua_unpack_counts_csr(sisua_handle h, const int32_t* indptr, const uint16_t* cols, const uint16_t* vals, float* dst,
                            int rows, void* stream);

/* One train step from HOST buffers — the call a reference-side data pipeline makes per minibatch
 * (the dict of sisua/data/_single_cell_base.py:582-591 as it comes out of tf.data, on the host).
 * The library stages the batch on a private copy stream into double-buffered device memory (so the H2D copy of
 * call i+1 overlaps the kernels of call i), widens 16-bit / CSR count formats to fp32, runs sisua_train_step on
 * `stream`, and copies the scalar loss (and, if host_terms != NULL, the [5,B] terms) back to the host.
 * Asynchronous: host buffers must stay valid and host_loss / host_terms are defined once `stream` has drained
 * (pinned host memory keeps the copies asynchronous).  Follow with sisua_adam_step (after the gradient all-reduce
 * when data-parallel).  format: 0 = float32 [B,G], 1 = uint16 [B,G], 2 = CSR (indptr int32 [B+1] from 0,
 * uint16 gene ids, uint16 counts, nnz entries). */
#define SISUA_HOST_F32 0
#define SISUA_HOST_U16 1
#define SISUA_HOST_CSR 2
typedef struct sisua_host_batch {
  int32_t format;
  int32_t B;
  const void* x;            /* F32 / U16 dense counts; unused for CSR */
  const int32_t* indptr;    /* CSR */
  const uint16_t* cols;     /* CSR */
  const uint16_t* vals;     /* CSR */
  int64_t nnz;              /* CSR */
  const float* y;           /* [B,P] or NULL    (arguments as in sisua_train_step, host memory) */
  const float* library;     /* [B,2] or NULL */
  const uint8_t* mask;      /* [B] or NULL */
  const float* eps_z;       /* [B,Z] or NULL */
  const float* eps_l;       /* [B] or NULL */
} sisua_host_batch;
int sisua_train_step_host(sisua_handle h, const sisua_host_batch* batch, uint64_t seed, int64_t step, float* host_loss,
                          float* host_terms, void* stream);

/* Measurement hooks for bench.py: kernels launched so far through this handle; per-section device time
 * (CUDA events on the caller's stream). Sections: 0 first encoder layer, 1 mid forward, 2 output heads +
 * count likelihood (fused: one kernel), 3 mid backward, 4 first-layer weight gradient, 5 Adam. */
int64_t sisua_launch_count(sisua_handle h);
int sisua_profile_enable(sisua_handle h, int on);
int sisua_profile_read(sisua_handle h, float* ms_out, int* counts_out);

/* tcgen05 descriptor self-test used by the GPU tests: D[128,N] = A[128,K] . B[N,K]^T (device pointers, fp32 in/out,
 * fp16 operands inside); a_mn_major / b_mn_major pick the shared-memory operand layout handed to the MMA. */
int sisua_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn_major, int b_mn_major, void* stream);

const char* sisua_last_error(sisua_handle h);
const char* sisua_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SISUA_B200_H_ */
