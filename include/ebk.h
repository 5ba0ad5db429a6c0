/*
 * ebk.h -- C-ABI of the B200-native NRMS-family hot path ("ebk" = EB-NeRD kernels).
 *
 * The reference (ebanalyse/ebnerd-benchmark) has NO FFI/plugin interface: its hot
 * path is Keras graph code executed by TensorFlow.  Each entry point below
 * therefore cites the reference *Python* lines whose arithmetic it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the
 * ctypes binding a maintainer adds under src/ebrec/models/newsrec/.
 *
 * Conventions
 *  - every function returns 0 on success, a negative ebk_status otherwise, and never
 *    throws; ebk_last_error() returns a thread-local message for the last failure;
 *  - all tensor arguments are raw DEVICE pointers owned by the caller (row-major,
 *    contiguous, 16-byte aligned), fp32 unless noted, indices int32;
 *  - kernels never allocate: scratch and saved activations live in a caller-provided
 *    workspace whose size comes from the matching *_workspace_bytes() query;
 *    a forward and its backward must be given the SAME workspace;
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous and ordered only by
 *    the stream.  Entry points keep no modes between calls: per-call behaviour comes from the
 *    descriptor / options structs.  (Process-wide state: the tensor-map cache, the measurement
 *    hooks at the end of this header, and "a deferred weight gradient is outstanding", which
 *    ebk_join_deferred clears.)
 */
#ifndef EBK_H_
#define EBK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  EBK_OK = 0,
  EBK_ERR_INVALID = -1,   /* bad argument / shape / alignment */
  EBK_ERR_WORKSPACE = -2, /* workspace too small */
  EBK_ERR_CUDA = -3,      /* CUDA runtime error (message in ebk_last_error) */
  EBK_ERR_UNSUPPORTED = -4
} ebk_status;

/* Arithmetic mode of the dense contractions. */
typedef enum {
  EBK_MATH_FP32 = 0, /* CUDA-core fp32 FMA (exact-ish; small shapes, debugging) */
  EBK_MATH_TF32 = 1, /* tcgen05 kind::tf32 tensor cores (operands rounded to nearest), fp32 accumulate in TMEM */
  EBK_MATH_TF32X3 = 2 /* error-compensated 3xTF32 (hi/lo operand split, 3 MMAs): ~fp32 accuracy; inference default */
} ebk_math;

const char* ebk_last_error(void);
int ebk_version(void);
/* 1 when the library was built for sm_100a and the current device is CC 10.x. */
int ebk_device_ok(void);

/* ------------------------------------------------------------------------------------
 * Sequence encoder = [gather] -> Dropout -> SelfAttention -> Dropout -> AttLayer2.
 * One description serves the news encoder (token ids gathered from the table,
 * nrms.py:116-159) and the user encoder (dense [n_seq, L, Din] input, no dropout,
 * nrms.py:92-114).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_seq;   /* sequences: articles N = B*(H+C) (news) or impressions B (user)      */
  int32_t L;       /* sequence length: title_size T (news) or history_size H (user), <=64 */
  int32_t Din;     /* input width: table width E (news) or D (user); multiple of 4        */
  int32_t nh, dh;  /* heads, head dim (layers.py:137-141); D = nh*dh, dh <= 32            */
  int32_t att;     /* attention_hidden_dim of AttLayer2 (layers.py:14-22); 0 = stop after the
                      SelfAttention: out / d_out are [n_seq*L, D] and no Dropout follows it -- the
                      optional Dense/BN stack of nrms.py:142-152 continues with ebk_dense_* and
                      ebk_attlayer_*; attW/attb/attq and their gradients may then be NULL        */
  int32_t V;       /* table rows when gathering; ids outside [0,V) -> zero row, no grad   */
  float dropout;   /* Keras Dropout rate (nrms.py:136,153); used only when training != 0  */
  int32_t math;    /* ebk_math                                                             */
} ebk_seqenc_desc;

size_t ebk_seqenc_workspace_bytes(const ebk_seqenc_desc* d);

/* Forward.  Replaces Embedding+Dropout+SelfAttention+Dropout+AttLayer2
 * (nrms.py:125-156; layers.py:200-254 with the adjoint_a=True product of layers.py:249;
 * layers.py:55-81 incl. exp without max-subtraction and the +1e-7 of layers.py:75-77).
 *   tok   [n_seq, L] int32 or NULL; table/x: [V, Din] when tok != NULL else [n_seq*L, Din]
 *   Wqkv  [Din, 3*D]  columns = WQ | WK | WV (layers.py:155-172, no bias)
 *   attW  [D, att], attb [att], attq [att]   (layers.py:35-52)
 *   out   [n_seq, D]
 *   training != 0: inverted dropout with the counter-based mask of DESIGN.md
 *   (seed1: embedded tokens, element index r*Din+e; seed2: attention output, r*D+d). */
int ebk_seqenc_fwd(const ebk_seqenc_desc* d, const int32_t* tok, const float* table_or_x,
                   const float* Wqkv, const float* attW, const float* attb, const float* attq,
                   int training, uint64_t seed1, uint64_t seed2,
                   void* workspace, size_t workspace_bytes, float* out, void* stream);

/* Backward of ebk_seqenc_fwd (the reference gets it from TF autodiff).  Gradients are
 * ACCUMULATED (+=) into dWqkv/dattW/dattb/dattq.  Input gradient:
 *   tok != NULL: rows of dX are scatter-added into d_table [V, Din] (the Embedding's
 *                IndexedSlices gradient, nrms.py:125-134) when d_table != NULL; or, when d_x != NULL
 *                (d_table NULL), the UNMASKED per-row gradients dX [n_seq*L, Din] are written to d_x for
 *                ebk_embed_adam_step, which applies the dropout mask and sums rows per token itself;
 *   tok == NULL: d_x [n_seq*L, Din] is overwritten (may be NULL to skip). */
int ebk_seqenc_bwd(const ebk_seqenc_desc* d, const int32_t* tok, const float* table_or_x,
                   const float* Wqkv, const float* attW, const float* attb, const float* attq,
                   int training, uint64_t seed1, uint64_t seed2,
                   void* workspace, size_t workspace_bytes, const float* d_out,
                   float* dWqkv, float* dattW, float* dattb, float* dattq,
                   float* d_table, float* d_x, void* stream);

/* Per-call options of the sequence encoder (NULL = all off).  They replace the thread-local setters of the first
 * version of this ABI: nothing armed by one call can leak into the next.
 *   defer_wgrad       backward WITH token ids: the QKV weight-gradient GEMM (tensor-pipe bound) is launched on a
 *                     library-owned side stream, forked behind the input-gradient GEMM, and the call returns
 *                     without joining it: the caller may then enqueue work that does not touch dWqkv or the call's
 *                     workspace (the HBM-bound table pass of the optimizer) on its own stream and MUST call
 *                     ebk_join_deferred(stream) before anything that does (and before the next forward).
 *   table_grad_event  backward WITH token ids (data parallel): cudaEvent_t (as void*) recorded on the stream right
 *                     after the embedding-gradient scatter -- the table gradient is then final, so its collective
 *                     can start while the remaining backward kernels still run.
 *   peer_tables       forward WITH token ids (data parallel, rank-sharded embedding table; all ranks on one NVSwitch
 *                     box, <= 8): peer_tables[r] = this process's mapping of rank r's parameter buffer (own pointer
 *                     for r == rank).  Every 16-byte chunk of a gathered table row is read from the rank that owns
 *                     float index f (owner = f / peer_shard_floats) -- the all-gather of the updated table is fused
 *                     into the Embedding gather (nrms.py:125-134) over NVLink.  Only the all-TMA path
 *                     (ebk_seqenc_uses_tma(desc) == 1) gathers through peers; any other path returns
 *                     EBK_ERR_INVALID rather than read a stale local shard. */
/* Per-step scalars kept in DEVICE memory so that a training step captured in a CUDA graph can be replayed with new
 * values: the host writes the struct (one 24-byte copy) before every replay instead of passing seeds / alpha as
 * kernel arguments.  seed1 / seed2: the two Dropout seeds of ebk_seqenc_*; alpha: Adam's
 * lr*sqrt(1-b2^t)/(1-b1^t) of ebk_adam_keras_step_p / ebk_embed_adam_step_p (whose dropout seed is seed1). */
typedef struct {
  uint64_t seed1, seed2;
  float alpha;
  float reserved;
} ebk_step_params;

typedef struct {
  int32_t defer_wgrad;
  void* table_grad_event;
  const void* const* peer_tables;
  int32_t peer_world;
  size_t peer_shard_floats;
  const ebk_step_params* step_dev; /* DEVICE pointer or NULL: when set, seed1 / seed2 come from it, not from the arguments */
  void* token_csr_ws;              /* DEVICE scratch of ebk_token_csr_bytes(n_seq * L, V) bytes or NULL: when set (forward on the
                                      all-TMA path, see ebk_seqenc_uses_tma), the token positions are grouped by id first and
                                      every DISTINCT table row is read once -- what pays when rows come from peer tables */
  size_t token_csr_ws_bytes;
} ebk_seqenc_opts;
size_t ebk_token_csr_bytes(int32_t R, int32_t V);

int ebk_seqenc_fwd_opts(const ebk_seqenc_desc* d, const ebk_seqenc_opts* opts, const int32_t* tok,
                        const float* table_or_x, const float* Wqkv, const float* attW, const float* attb,
                        const float* attq, int training, uint64_t seed1, uint64_t seed2, void* workspace,
                        size_t workspace_bytes, float* out, void* stream);
int ebk_seqenc_bwd_opts(const ebk_seqenc_desc* d, const ebk_seqenc_opts* opts, const int32_t* tok,
                        const float* table_or_x, const float* Wqkv, const float* attW, const float* attb,
                        const float* attq, int training, uint64_t seed1, uint64_t seed2, void* workspace,
                        size_t workspace_bytes, const float* d_out, float* dWqkv, float* dattW, float* dattb,
                        float* dattq, float* d_table, float* d_x, void* stream);
/* joins the side stream of a defer_wgrad backward into `stream` (no-op when nothing is outstanding) */
int ebk_join_deferred(void* stream);
/* 1 when this descriptor runs on the all-TMA tcgen05 path (EBK_MATH_TF32, att % 4 == 0, TMA-encodable strides) */
int ebk_seqenc_uses_tma(const ebk_seqenc_desc* d);

/* CUDA IPC plumbing for peer_tables: ebk_ipc_export / ebk_ipc_open wrap cudaIpcGetMemHandle / cudaIpcOpenMemHandle
 * (handle64: 64 bytes; offset of ptr inside its allocation) so that the host code can exchange the mappings with any
 * byte transport. */
int ebk_ipc_export(const void* ptr, void* handle64, size_t* offset);
int ebk_ipc_open(const void* handle64, size_t offset, void** out);
/* cudaMemcpyAsync(device to device) on `stream`; src may be a peer mapping from ebk_ipc_open: a rank can refresh its
 * replica of the rank-sharded table from its peers' memory without entering a collective. */
int ebk_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Dense(+ReLU) -> [BatchNormalization] -> [Dropout] layer of the NRMSDocVec news encoder
 * (nrms_docvec.py:118-130: Dense(units, relu, l2) + BatchNormalization() + Dropout(p); :130 the
 * output Dense(D, relu) is the same call with bn = 0, dropout = 0).
 *   x [N, K], W [K, U], b [U], gamma/beta/mov_mean/mov_var [U], y [N, U]
 *   training != 0: batch statistics over the N rows of THIS call (biased variance) and the Keras
 *   moving-average update mov = mov*momentum + batch*(1-momentum) written to mov_mean/mov_var;
 *   training == 0: moving statistics, no dropout.
 * Backward ACCUMULATES dW, db, dgamma, dbeta; dW also receives 2*l2*l2_grad_scale*W; dx (nullable) is
 * overwritten.  `y` (the forward output) is only read when bn == 0.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t N, K, U;        /* rows, input width, units (K, U multiples of 4) */
  int32_t relu;           /* 1: relu after the bias */
  int32_t bn;             /* 1: BatchNormalization after the activation */
  float bn_momentum;      /* Keras default 0.99 */
  float bn_eps;           /* Keras default 1e-3 */
  float dropout;          /* after BN; training only */
  float l2;               /* kernel_regularizer l2 factor (nrms_docvec.py:122-124) */
  int32_t math;           /* ebk_math */
} ebk_dense_desc;

size_t ebk_dense_workspace_bytes(const ebk_dense_desc* d);
int ebk_dense_fwd(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                  const float* beta, float* mov_mean, float* mov_var, int training, uint64_t seed,
                  void* workspace, size_t workspace_bytes, float* y, void* stream);
int ebk_dense_bwd(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y,
                  int training, uint64_t seed, void* workspace, size_t workspace_bytes, const float* dy,
                  float l2_grad_scale, float* dW, float* db, float* dgamma, float* dbeta, float* dx, void* stream);
/* The same two calls for CUDA-graph replay: the dropout seed of the layer is
 *   (seed_sel == 0 ? step_dev->seed1 : step_dev->seed2) + seed_add, read from device memory by the kernels, so a captured
 *   launch bakes no seed in (NRMSDocVec: seed1 = history call, seed2 = candidate call, seed_add = layer index). */
int ebk_dense_fwd_p(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                    const float* beta, float* mov_mean, float* mov_var, int training, const ebk_step_params* step_dev,
                    int seed_sel, uint64_t seed_add, void* workspace, size_t workspace_bytes, float* y, void* stream);
int ebk_dense_bwd_p(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y,
                    int training, const ebk_step_params* step_dev, int seed_sel, uint64_t seed_add, void* workspace,
                    size_t workspace_bytes, const float* dy, float l2_grad_scale, float* dW, float* db, float* dgamma,
                    float* dbeta, float* dx, void* stream);
/* out[0] += scale * sum_i x[i]^2   (the l2 kernel-regulariser term of the loss) */
int ebk_sumsq_accum(const float* x, size_t n, float scale, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * AttLayer2 on its own (layers.py:7-104; call 55-81) with an optional Dropout on its input:
 * the NAML title/body pooling (naml.py:167-168, 198-199), the 4-view fusion (naml.py:133-138) and the
 * NAML user encoder (naml.py:79-84).
 *   x [n_seq*L, D], W [D, att], b [att], q [att]; out row n is written at out + n*out_ld (out_ld >= D,
 *   so several views can fill one [n_seq, 4, D] buffer).  L <= 64.
 *   training != 0 and dropout > 0: x is masked on read with the counter-based mask (element r*D + d).
 * Backward ACCUMULATES dW, db, dq and OVERWRITES dx [n_seq*L, D] with the gradient w.r.t. the MASKED
 * input dropout(x); the layer that produced x applies the same mask (ebk_conv1d_bwd: seed_out).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_seq, L, D, att;
  float dropout;
  int32_t math; /* ebk_math */
} ebk_attlayer_desc;

size_t ebk_attlayer_workspace_bytes(const ebk_attlayer_desc* d);
int ebk_attlayer_fwd(const ebk_attlayer_desc* d, const float* x, const float* W, const float* b, const float* q,
                     int training, uint64_t seed, void* workspace, size_t workspace_bytes, float* out,
                     int32_t out_ld, void* stream);
int ebk_attlayer_bwd(const ebk_attlayer_desc* d, const float* x, const float* W, const float* q, int training,
                     uint64_t seed, void* workspace, size_t workspace_bytes, const float* d_out, int32_t d_out_ld,
                     float* dW, float* db, float* dq, float* dx, void* stream);

/* ------------------------------------------------------------------------------------
 * NAML text view up to the pooling: Embedding -> Dropout -> Conv1D(F, window, padding="same",
 * activation) (naml.py:155-166 title, 186-197 body; table shared, naml.py:318-323).
 *   tok [n_seq, L] int32 (ids outside [0,V) -> zero row, no gradient), table [V, E],
 *   Wc [window*E, F] = the Keras kernel [window, E, F] flattened, bc [F],
 *   y [n_seq*L, F] = act(conv + bc)  -- BEFORE the output Dropout of naml.py:167/198, which the
 *   consumer (ebk_attlayer_*) applies on read with seed_out.
 *   "same" padding: (window-1)/2 zero rows on the left, the rest on the right (TF convention).
 * The conv is ONE tensor-core GEMM with K = window*E over a zero-padded copy of the dropped embeddings
 * (rows of one article are contiguous, so a window is a contiguous K-slice).
 * Backward: dy [n_seq*L, F] = gradient w.r.t. dropout_out(y) (unmasked, from ebk_attlayer_bwd);
 * ACCUMULATES dWc, dbc and scatter-adds the embedding gradient into d_table [V, E] (nullable).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_seq, L, E, F, window, V; /* E, F multiples of 4 */
  float dropout;                     /* rate of both Dropout layers; training only */
  int32_t relu;                      /* cnn_activation == "relu" */
  int32_t math;                      /* ebk_math */
} ebk_conv1d_desc;

size_t ebk_conv1d_workspace_bytes(const ebk_conv1d_desc* d);
int ebk_conv1d_fwd(const ebk_conv1d_desc* d, const int32_t* tok, const float* table, const float* Wc,
                   const float* bc, int training, uint64_t seed_in, void* workspace, size_t workspace_bytes,
                   float* y, void* stream);
int ebk_conv1d_bwd(const ebk_conv1d_desc* d, const int32_t* tok, const float* Wc, const float* y, int training,
                   uint64_t seed_in, uint64_t seed_out, void* workspace, size_t workspace_bytes, const float* dy,
                   float* dWc, float* dbc, float* d_table, void* stream);

/* ------------------------------------------------------------------------------------
 * NAML categorical view: Embedding(n_cat, dim) -> Dense(F, activation) (naml.py:205-252).
 * Only n_cat distinct inputs exist, so the layer is evaluated once per category
 * (T = act(emb W + b), [n_cat+1, F]; row n_cat = act(b) serves ids outside [0, n_cat)) and gathered;
 * the backward segment-sums the row gradients per category first.
 *   ids [N] int32, emb [n_cat, dim], W [dim, F], b [F]; out row n at out + n*out_ld.
 *   workspace: ebk_catview_workspace_bytes(n_cat, F); backward needs the forward's workspace.
 * Backward ACCUMULATES d_emb, dW, db.
 * ---------------------------------------------------------------------------------- */
size_t ebk_catview_workspace_bytes(int32_t n_cat, int32_t F);
int ebk_catview_fwd(int32_t N, int32_t n_cat, int32_t dim, int32_t F, int32_t relu, const int32_t* ids,
                    const float* emb, const float* W, const float* b, void* workspace, size_t workspace_bytes,
                    float* out, int32_t out_ld, void* stream);
int ebk_catview_bwd(int32_t N, int32_t n_cat, int32_t dim, int32_t F, int32_t relu, const int32_t* ids,
                    const float* emb, const float* W, void* workspace, size_t workspace_bytes, const float* d_out,
                    int32_t d_out_ld, float* d_emb, float* dW, float* db, void* stream);

/* ------------------------------------------------------------------------------------
 * Click score + loss.  Replaces Dot(axes=-1) + Activation("softmax") +
 * categorical_crossentropy (nrms.py:201-202, 61-62) and its gradient.
 *   news [B, C, D], user [B, D], labels [B, C] fp32 (one-hot)
 *   probs [B, C] out; loss_sum: 1 float, += sum_b CE_b * loss_scale  (caller zeroes)
 *   d_news [B, C, D], d_user [B, D] out = gradient of (loss_scale * sum_b CE_b);
 *   pass loss_scale = 1/B_global for Keras' batch-mean reduction.
 *   d_news/d_user may be NULL (forward only).
 * ---------------------------------------------------------------------------------- */
int ebk_score_softmax_ce(int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                         const float* labels, float loss_scale, float* probs, float* loss_sum,
                         float* d_news, float* d_user, void* stream);

/* The same with the loss selected by `kind` and separate scales for the gradient and for the value added to
 * loss_sum (data parallel: gradient scaled by 1/(B*world), reported loss by 1/B):
 *   EBK_LOSS_CATEGORICAL_CE  hparams.loss == "cross_entropy_loss" -> "categorical_crossentropy" (nrms.py:61-62)
 *   EBK_LOSS_BINARY_CE       hparams.loss == "log_loss" -> "binary_crossentropy" (nrms.py:63-64, base_model.py:63-66).
 *       The model output is still the softmax Activation (nrms.py:202); Keras' backend.binary_crossentropy picks up
 *       the logits cached on that output (`_keras_logits`) and evaluates sigmoid_cross_entropy_with_logits on
 *       them: loss_b = mean_c [max(z,0) - z*y + log(1 + exp(-|z|))], dz = (sigmoid(z) - y) / C.  probs stay softmax. */
typedef enum { EBK_LOSS_CATEGORICAL_CE = 0, EBK_LOSS_BINARY_CE = 1 } ebk_loss_kind;
int ebk_score_loss(int32_t kind, int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                   const float* labels, float grad_scale, float loss_scale, float* probs, float* loss_sum,
                   float* d_news, float* d_user, void* stream);

/* scorer head: sigmoid(news . user), nrms.py:204-205.  news [B, C, D] -> out [B, C]. */
int ebk_score_sigmoid(int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                      float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * tf.keras.optimizers.Adam (nrms.py:76-77), dense (non-lazy) Keras form:
 *   m += (g-m)(1-b1); v += (g*g-v)(1-b2); theta -= (m*alpha)/(sqrt(v)+eps)
 *   with alpha = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller; (1-b1), (1-b2) are formed
 *   in double and rounded to fp32 once, as Keras' Python-float constants are.
 * If zero_grad != 0 the gradient buffer is cleared in the same pass.
 * ---------------------------------------------------------------------------------- */
int ebk_adam_keras_step(float* theta, float* g, float* m, float* v, size_t n, float alpha,
                        double beta1, double beta2, float eps, int zero_grad, void* stream);
/* the same with alpha read from step_dev->alpha (device memory) when step_dev != NULL */
int ebk_adam_keras_step_p(float* theta, float* g, float* m, float* v, size_t n, float alpha,
                          const ebk_step_params* step_dev, double beta1, double beta2, float eps, int zero_grad,
                          void* stream);

/* ------------------------------------------------------------------------------------
 * Embedding table: IndexedSlices gradient + the same Keras-form Adam, fused (single-GPU training path).
 * Replaces the Embedding backward (nrms.py:125-134) and the table's share of Adam (nrms.py:76-77) without
 * materialising a dense [V, E] gradient: g[v, :] = sum over rows r with tok[r] == v of
 * dX[r, :] * dropout'(r, :) (mask of seed drop_seed, element index r*E + e, rate drop_p; 0 = off) is summed per
 * table row (ascending r: bit-reproducible) inside the optimizer pass; EVERY row's m, v, theta are updated
 * (non-lazy Adam, identical arithmetic to ebk_adam_keras_step).
 *   tok [R] int32 (ids outside [0, V) ignored), dX [R, E] = ebk_seqenc_bwd's d_x output, theta/m/v [V, E],
 *   d_table [V, E]: dense gradient buffer, must be all-zero on entry and is all-zero on exit (rows referenced
 *   more than 32 times are pre-reduced into it).  E % 4 == 0, E <= 1024.
 * ---------------------------------------------------------------------------------- */
size_t ebk_embed_adam_workspace_bytes(int32_t R, int32_t V);
/* ebk_embed_adam_step with drop_seed / alpha read from step_dev->seed1 / ->alpha (device memory) when step_dev != NULL */
int ebk_embed_adam_step_p(int32_t R, int32_t E, int32_t V, const int32_t* tok, const float* dX, float drop_p,
                          uint64_t drop_seed, float* theta, float* d_table, float* m, float* v, float alpha,
                          const ebk_step_params* step_dev, double beta1, double beta2, float eps, void* workspace,
                          size_t workspace_bytes, void* stream);
int ebk_embed_adam_step(int32_t R, int32_t E, int32_t V, const int32_t* tok, const float* dX, float drop_p,
                        uint64_t drop_seed, float* theta, float* d_table, float* m, float* v, float alpha,
                        double beta1, double beta2, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Data parallel, rank-sharded embedding table: gradient reduction FUSED into the Adam pass over NVLink peer memory
 * (no reference counterpart -- the reference is single-device; replaces reduce-scatter + ebk_adam_keras_step on the
 * shard).  Every rank scatter-adds its table gradient into its own dense [V, E] buffer (ebk_seqenc_bwd) and marks the
 * rows it touched (ebk_dp_token_flags: flags[v] = 1 for every token id in tok, v_pad >= V bytes, zeroed first); the
 * flags of all ranks are all-gathered into flags_all [world, v_pad].  ebk_adam_pull_step then updates floats
 * [lo_float, lo_float + n_float) of theta / m / v (this rank's shard of the table): the gradient of every 16-byte chunk
 * is the sum, in rank order, of the chunks of exactly those ranks whose flag for the row is set, read from
 * grads[p] (this process's CUDA-IPC mapping of rank p's gradient buffer; grads[rank] is local and the consumed local
 * chunks are cleared).  Same Keras-form arithmetic as ebk_adam_keras_step.  The caller orders it after every rank's
 * scatter (the flags all-gather does) and keeps peers from clearing their buffers until every owner has pulled.
 * ---------------------------------------------------------------------------------- */
int ebk_dp_token_flags(int32_t R, int32_t V, const int32_t* tok, uint8_t* flags, size_t v_pad, void* stream);
int ebk_adam_pull_step(float* theta, float* m, float* v, const void* const* grads, const uint8_t* flags_all,
                       size_t v_pad, int32_t world, int32_t rank, int32_t E, size_t lo_float, size_t n_float,
                       float alpha, const ebk_step_params* step_dev, double beta1, double beta2, float eps,
                       void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): number of kernels this library has launched so far, and an
 * optional per-kernel timer (CUDA events recorded on the launching stream around every
 * internal launch while enabled).  Slots [0, ebk_prof_num_tags()) are named by
 * ebk_prof_tag_name(); ebk_prof_collect() synchronises the recorded events and returns the
 * summed milliseconds and launch-group counts per slot since the last ebk_prof_enable(1).
 * ---------------------------------------------------------------------------------- */
long long ebk_launch_count(void);
int ebk_prof_enable(int on);
int ebk_prof_is_enabled(void);
int ebk_prof_num_tags(void);
const char* ebk_prof_tag_name(int slot);
int ebk_prof_collect(double* ms_out, long long* count_out);

/* ------------------------------------------------------------------------------------
 * Building blocks exported for the parity tests.
 * ---------------------------------------------------------------------------------- */
/* C[M,N] = (beta ? C : 0) + opA(A)[M,K] * opB(B)[K,N]; row-major; ld* in elements. */
int ebk_gemm(int32_t math, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
             const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
             float beta, void* stream);

/* The training-path GEMM on its own: all-TMA tcgen05 kind::tf32, C (+)= alpha * op(A) . op(B), beta in {0,1}.
 * A and B must hold tf32-representable values (the tensor core truncates); transA: A stored [K, M];
 * transB: B stored [N, K]; tall bits 0-3: 0 = 128-row tiles, 1 = 256-row tiles, 15 = auto; bits 4-7: thread-block
 * CTA pairs (tcgen05 cta_group::2, each CTA stages half of B), 0 = off, 1 = on, 15 = auto.  Replaces the K.dot / tf.matmul contractions of
 * layers.py:65, 214-230 and their autodiff transposes. */
int ebk_gemm_tma(int32_t transA, int32_t transB, int32_t tall, int32_t M, int32_t N, int32_t K,
                 const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                 float beta, float alpha, void* stream);

/* Multi-head attention core on packed projections (layers.py:231-252).
 *   qkv [n_seq*L, 3*D] -> y [n_seq*L, D] */
int ebk_attention_core_fwd(int32_t n_seq, int32_t L, int32_t nh, int32_t dh, const float* qkv,
                           float* y, void* stream);
/*  dy [n_seq*L, D] (optionally dropout-masked on read) -> dqkv [n_seq*L, 3*D] */
int ebk_attention_core_bwd(int32_t n_seq, int32_t L, int32_t nh, int32_t dh, const float* qkv,
                           const float* dy, float drop_p, uint64_t drop_seed, float* dqkv,
                           void* stream);

/* Keep-mask of the counter-based dropout, for tests: out[i] = 1.0f/0.0f, i in [0,n). */
int ebk_dropout_mask(uint64_t seed, float p, size_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EBK_H_ */
