/*
 * dmvae_b200.h - C ABI of the B200-native DMVAE / VaDE / MoE training-step kernels.
 *
 * The reference (ffs97/deep-mixture-vae) has no FFI of its own: its seam is the Python model API
 * plus TensorFlow's session.run(fetches, feed_dict) (code/base_models.py:126).  This library is
 * what sits under that seam instead of the TensorFlow runtime.  Each entry point names the reference
 * graph ops (file:line under /root/reference/code) that it replaces.
 *
 * Conventions
 *  - plain C: pointers, sizes, enums; no C++/torch types.
 *  - every pointer is a caller-owned DEVICE pointer unless the name ends in _host.
 *  - every call is asynchronous on the caller-supplied cudaStream_t (passed as void*); nothing
 *    allocates or synchronises after dmvae_ctx_create, so calls may be captured in a CUDA graph.
 *  - return value 0 = OK, non-zero = error code; dmvae_last_error() gives the message of the last
 *    failing call on this host thread.  No exception crosses the boundary.
 *  - matrices are row-major; leading dimensions (ld*) are in ELEMENTS.
 *
 * Padded activation layout ("ones column"): the dense layers fold the bias into the GEMM.  An
 * activation matrix for a layer with n inputs is stored [rows, n_pad] with column n equal to 1.0
 * and columns > n equal to 0; the weight matrix is stored [n_pad, out_pad] with row n holding the
 * bias and rows > n zero.  The weight-gradient GEMM then produces the bias gradient in row n for
 * free.  n_valid / n_block below describe where a kernel must write those 1.0 / 0 columns:
 * column j is data when (j % n_block) < n_valid, 1.0 when == n_valid, 0 when > n_valid.
 */
#ifndef DMVAE_B200_H
#define DMVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMVAE_B200_ABI_VERSION 2   /* 2: dmvae_gemm_epilogue.recon, dmvae_elbo_args.r_part, dmvae_reparam_args.fold_*, x_scale arguments */

typedef struct dmvae_ctx dmvae_ctx;

enum dmvae_status {
  DMVAE_OK = 0,
  DMVAE_ERR_INVALID = 1,      /* bad argument (shape, alignment, enum) */
  DMVAE_ERR_CUDA = 2,         /* a CUDA runtime / driver call failed */
  DMVAE_ERR_UNSUPPORTED = 3,  /* valid request this build cannot serve (e.g. no sm_100 device) */
  DMVAE_ERR_NCCL = 4
};

enum dmvae_dtype { DMVAE_F32 = 0, DMVAE_BF16 = 1, DMVAE_U8 = 2 };
enum dmvae_act { DMVAE_ACT_NONE = 0, DMVAE_ACT_RELU = 1, DMVAE_ACT_SIGMOID = 2 /* reconstructed_X, base_models.py:295-296 */ };
enum dmvae_input_type { DMVAE_INPUT_BINARY = 0, DMVAE_INPUT_REAL = 1 };   /* base_models.py:72-85 */
enum dmvae_elbo_mode {
  DMVAE_MODE_DMVAE = 0,          /* cluster_sample=False, w = softmax(logits)   priors.py:130-145 */
  DMVAE_MODE_DMVAE_SAMPLED = 1,  /* cluster_sample=True,  w = zeta (concrete)   priors.py:118-128 */
  DMVAE_MODE_VADE = 2            /* w = gamma = get_cluster_probs(Z)            priors.py:91-102  */
};

/* ---- context -------------------------------------------------------------------------------- */
int dmvae_abi_version(void);
const char* dmvae_last_error(void);
int dmvae_ctx_create(int device, dmvae_ctx** out);
int dmvae_ctx_destroy(dmvae_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t dmvae_ctx_launch_count(const dmvae_ctx* ctx);
/* 1 if the tcgen05/TMA bf16 GEMM path is usable on this device (sm_100), else 0 */
int dmvae_ctx_has_tcgen05(const dmvae_ctx* ctx);

/* ---- dense layers: replaces tf.layers.dense / FullyConnected and their autodiff ---------------
 * (base_models.py:221-248, :280-293; includes/layers.py:30-36; gradient ops from :110).
 * dtype selects the operand element type AND the engine:
 *   DMVAE_BF16: tcgen05.mma (kind::f16, fp32 accumulate in TMEM) fed by TMA   - 2e-2 tier
 *   DMVAE_F32 : fp32 SIMT FFMA tiles                                          - 1e-4 tier
 */
/* Reconstruction term fused into the OUTPUT layer's GEMM epilogue (define_recon_loss, base_models.py:72-85, and its
 * gradient): the epilogue holds the fp32 decoder logits d of a tile in registers, reads the targets x of the same
 * elements, and stores  scale * (sigmoid(d) - x)  (binary) or  scale * (d - x)  (real) in bf16 - the operand of the
 * decoder's backward GEMMs - INSTEAD of the logits; the per-row sum of the reconstruction term over the column range
 * one epilogue warp covers (32 to 256 columns, the kernel's choice) goes to its own slot of r_part[row, :], to be added
 * in a fixed order by dmvae_elbo_reduce.  The caller zero-fills r_part before the GEMM (slots no warp owns stay 0).
 * bf16 tcgen05 path only, act = NONE, no split-K, not inside dmvae_gemm_chain; columns >= D are stored as 0. */
typedef struct dmvae_recon_fuse {
  const void* X; int32_t x_dtype; int64_t ldx;   /* targets [rows, D]: DMVAE_U8 (with x_scale) or DMVAE_F32 */
  float x_scale;
  int32_t input_type;                              /* dmvae_input_type */
  float scale;                                     /* inv_global_batch * recon_scale */
  int32_t D;
  float* r_part; int32_t r_parts;                  /* fp32 [rows, r_parts], zero-filled; r_parts >= ceil(N / 32) */
} dmvae_recon_fuse;

typedef struct dmvae_gemm_epilogue {
  int32_t out_dtype;      /* DMVAE_F32 or DMVAE_BF16 (BF16 only with dtype=BF16) */
  int32_t act;            /* dmvae_act, applied to data columns */
  int32_t n_valid;        /* see "ones column" above; n_valid >= n_block disables it */
  int32_t n_block;
  float pad_one;          /* value written at column n_valid of every block: 1.0 (forward) or 0.0 (gradients) */
  const void* relu_mask;  /* optional [M, N] matrix of operand dtype: out *= (mask > 0)   (dgrad) */
  int64_t ld_mask;
  const float* bias;      /* optional fp32 [N] added before the activation (NULL in the padded layout) */
  int32_t accumulate;     /* 1: C += result (fp32 C only; required when split_k > 1) */
  int32_t split_k;        /* >= 1; partial sums are combined with fp32 red.global.add */
  const dmvae_recon_fuse* recon;   /* optional (host pointer, read during the call): fuse the reconstruction term */
} dmvae_gemm_epilogue;

/* C[M,N] = epilogue( op(A) . op(B) );  op(A) is [M,K]: trans_a=0 -> A stored [M,K], 1 -> stored [K,M].
 * op(B) is [K,N]: trans_b=0 -> B stored [K,N], 1 -> stored [N,K]. */
int dmvae_gemm(dmvae_ctx* ctx, int dtype, int trans_a, int trans_b,
               const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
               int M, int N, int K, const dmvae_gemm_epilogue* epi, void* stream);

/* Y = act(X.W) with the ones/zero padding columns written          (forward of one dense layer) */
int dmvae_linear_fwd(dmvae_ctx* ctx, int dtype, const void* X, int64_t ldx, const void* W, int64_t ldw,
                     void* Y, int64_t ldy, int out_dtype, int rows, int n_out_pad, int n_in_pad,
                     int act, int n_valid, int n_block, void* stream);
/* dX = (dY.W^T) * (act_in > 0), padding columns zeroed             (data gradient) */
int dmvae_linear_dgrad(dmvae_ctx* ctx, int dtype, const void* dY, int64_t lddy, const void* W, int64_t ldw,
                       const void* act_in, int64_t ld_act, void* dX, int64_t lddx, int out_dtype,
                       int rows, int n_in_pad, int n_out_pad, int n_valid, int n_block, void* stream);
/* dW (+)= X^T.dY  (fp32 [n_in_pad, n_out_pad]; row n_in is the bias gradient) */
int dmvae_linear_wgrad(dmvae_ctx* ctx, int dtype, const void* X, int64_t ldx, const void* dY, int64_t lddy,
                       float* dW, int64_t lddw, int rows, int n_in_pad, int n_out_pad,
                       int accumulate, int split_k, void* stream);

/* ---- input staging ------------------------------------------------------------------------- */
/* X (f32 / u8 / bf16, [rows, D], ldx) -> A0 (out_dtype, [rows, ld_out]) with the ones column at D.
 * x_scale: value of one unit of a uint8 input (1 for binarised data stored as 0/1, 1/255 for 8-bit intensities such as
 * the CIFAR pixels of includes/utils.py:204-210, which then travel and are re-read at 1 byte per element); ignored for
 * the other dtypes; 0 means 1. */
int dmvae_stage_input(dmvae_ctx* ctx, const void* X, int x_dtype, int64_t ldx, void* A0, int out_dtype,
                      int64_t ld_out, int rows, int D, float x_scale, void* stream);

/* Shuffled minibatch: dst[i, :] = src[idx[i], :], row_bytes per row (replaces the per-row Python append of
 * Dataset.get_batches, includes/utils.py:449-463).  idx is a DEVICE int32 array; src may be a PINNED HOST buffer
 * (unified addressing: the kernel reads it over PCIe / NVLink-C2C, so the shuffle costs no host-side gather and the
 * rows cross the bus exactly once); dst is device memory.  Rows and pitches must be 4-byte multiples (16 for speed). */
int dmvae_gather_rows(dmvae_ctx* ctx, const void* src, int64_t src_pitch_bytes, const int32_t* idx, void* dst,
                      int64_t dst_pitch_bytes, int rows, int row_bytes, void* stream);
/* The same for {0,1}-valued rows kept ONE BIT per element on the host (numpy.packbits(..., bitorder="little"), each row
 * padded with zero bytes to a multiple of 16 bytes): dst[r, :D] = the bits of packed row idx[r] as uint8 0/1.  8x fewer bytes
 * over the bus. */
int dmvae_gather_rows_bits(dmvae_ctx* ctx, const void* src_bits, int64_t src_pitch_bytes, const int32_t* idx, void* dst,
                           int64_t dst_pitch_bytes, int rows, int D, void* stream);

/* ---- reparameterisation: priors.py:86-89 (Z), :170-181 (concrete), utils.py:17-19 (Gumbel),
 *      host RNG of priors.py:67-68 replaced by Philox4x32-10 ------------------------------------ */
typedef struct dmvae_reparam_args {
  int32_t rows, L, K;
  const float* mean; const float* log_var; int64_t ld_zh;   /* fp32 [rows, L] each */
  const float* logits; int64_t ld_logits;                   /* fp32 [rows, K] or NULL (no concrete sample) */
  const float* eps_in;                                       /* injected N(0,1) [rows, L] (ld = L) or NULL -> Philox */
  const float* gumbel_in;                                    /* injected Gumbel [rows, K] (ld = K) or NULL -> Philox */
  uint64_t seed; uint64_t step; uint64_t row_offset;         /* Philox key / counter; row_offset = global row of row 0 */
  const uint64_t* step_dev;                                   /* optional device-resident step counter (overrides step; CUDA-graph replay) */
  float tau;                                                  /* concrete temperature */
  void* Z_out; int32_t z_dtype; int64_t ld_z; int32_t z_cols; /* Z in operand dtype, [rows, ld_z]; ones column at L, zeros to z_cols */
  float* eps_out;                                             /* fp32 [rows, L]: the eps actually used (kept for the backward) */
  float* zeta_out;                                            /* fp32 [rows, K] or NULL */
  /* optional: fold the three column groups of a split-weight logits GEMM first (see dmvae_split3_bf16):
   * fold[r, k] += fold[r, stride + k] + fold[r, 2 stride + k], k < fold_K.  `logits` may alias `fold`. */
  float* fold; int64_t ld_fold; int32_t fold_K; int32_t fold_stride;
} dmvae_reparam_args;
int dmvae_reparam_fwd(dmvae_ctx* ctx, const dmvae_reparam_args* a, void* stream);

/* dmu = dmu_kl + dZ (+ d_mean_extra) ; dlv = dlv_kl + 1/2 dZ eps exp(lv/2)   -> [rows, out_cols] operand-dtype matrix
 * (columns [0,L) = dmu, [L,2L) = dlv, rest 0) that feeds the head layer's dgrad / wgrad. */
int dmvae_reparam_bwd(dmvae_ctx* ctx, int rows, int L, const float* d_mean_kl, const float* d_log_var_kl,
                      int64_t ld_kl, const float* dZ, int64_t ld_dz, const float* dZ_extra, int64_t ld_dze,
                      const float* eps, const float* log_var, int64_t ld_lv,
                      const float* d_mean_extra, int64_t ld_dme,
                      void* out, int out_dtype, int64_t ld_out, int out_cols, void* stream);

/* ---- a chain of dependent dense-layer GEMMs in ONE persistent launch (bf16 tcgen05 only) ---------
 * The layers of one pass (base_models.py:218-293 forward, or the data-gradient / weight-gradient GEMMs of its
 * autodiff) are walked as one work list by persistent CTA pairs; a tile of layer e starts as soon as the 256-row
 * block(s) of the earlier layers it reads are complete (per-row-block counters in `counters`), so consecutive
 * layers overlap and there is one launch instead of one per layer.  Results are identical to calling dmvae_gemm
 * per entry in order.  dep[d] names an EARLIER entry whose output C this entry reads (as A, B or ReLU mask), -1 for
 * none; dep_all[d] = 1 when every row block of that output is needed (operands reduced over rows: weight
 * gradients), 0 when only the rows of the tile itself are (A operand / mask of a same-M layer).
 * fuse = 1 on the latent head entry (fp32 output [mean | log_var], 2L <= 32) applies dmvae_reparam_fwd's Gaussian
 * part in the epilogue (arguments in `reparam`; mean / log_var / ld_zh of it are ignored). */
typedef struct dmvae_chain_gemm {
  int32_t trans_a, trans_b;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  int32_t M, N, K;
  dmvae_gemm_epilogue epi;
  int32_t dep[2];
  int32_t dep_all[2];
  int32_t fuse;
} dmvae_chain_gemm;
/* number of int32 counters a chain of n entries over at most max_rows rows needs */
int64_t dmvae_gemm_chain_counters(int n, int max_rows);
/* zero_counters = 1: the call clears the counters itself (a memset node before the kernel); 0: the caller
 * guarantees they are zero when the kernel starts (e.g. cleared once per step for all chains). */
int dmvae_gemm_chain(dmvae_ctx* ctx, const dmvae_chain_gemm* g, int n, int32_t* counters, int64_t counters_len,
                     int zero_counters, const dmvae_reparam_args* reparam, void* stream);

/* ---- fused ELBO forward + backward ------------------------------------------------------------
 * replaces define_recon_loss (base_models.py:72-85), DiscreteFactorial.kl_from_prior
 * (priors.py:183-201), NormalMixtureFactorial.kl_from_prior (priors.py:104-147), get_cluster_probs
 * (priors.py:91-102), softmax (base_models.py:249), the loss sum (base_models.py:91-93) and the
 * autodiff of all of them.  Never materialises [B,K,L]. */
typedef struct dmvae_elbo_args {
  int32_t mode;            /* dmvae_elbo_mode */
  int32_t input_type;      /* dmvae_input_type */
  int32_t rows, D, L, K;
  const void* X; int32_t x_dtype; int64_t ldx;                 /* targets [rows, D] */
  const void* decoded; int32_t dec_dtype; int64_t ld_dec;      /* decoder logits [rows, >=D] */
  const float* mean; const float* log_var; int64_t ld_zh;      /* fp32 */
  const float* logits; int64_t ld_logits;                      /* fp32 [rows,K]   (DMVAE modes) */
  const float* eps; int64_t ld_eps;                            /* fp32 [rows,L]   (VADE: Z = mu + e^{lv/2} eps) */
  const float* zeta; int64_t ld_zeta;                          /* fp32 [rows,K]   (SAMPLED) */
  float tau;
  const float* prior_means; const float* prior_log_vars;       /* fp32 [K,L] dense */
  float kl_ratio; float inv_global_batch;
  const float* kl_ratio_dev;                                   /* optional device-resident kl_ratio (overrides kl_ratio; CUDA-graph replay) */
  float recon_scale;       /* weight of the reconstruction term in the loss and its gradient (1; 0 for prior pre-training) */
  /* outputs */
  float* per_sample;       /* [rows,4]: recon, KL_c, KL_z, recon_scale recon + kl_ratio (KL_c + KL_z) */
  float* qc;               /* [rows,K] q(c|x) (softmax(logits) or gamma) */
  int32_t* argmax;         /* [rows]   argmax_k q(c|x) */
  void* d_decoded; int64_t ld_ddec; int32_t ddec_cols;         /* dec_dtype [rows, ld]; columns [D, ddec_cols) zeroed */
  float* d_mean_kl; float* d_log_var_kl; int64_t ld_dkl;       /* fp32 [rows,L]: KL-side gradients (before reparam backward) */
  void* d_logits; int32_t dlogits_dtype; int64_t ld_dlogits; int32_t dlogits_cols; /* [rows, ld]; cols [K, dlogits_cols) zeroed (DMVAE modes) */
  float* d_Z_gamma; int64_t ld_dzg;                            /* fp32 [rows,L] gradient reaching Z through gamma (VADE) */
  float* w_scratch;        /* fp32 [rows,K]: d_s (VADE) */
  float* f_scratch;        /* fp32 [rows,2L]: [g_mbar | g_pbar] (SAMPLED) */
  /* VADE only, optional: d(other loss)/d gamma [rows,K] (already scaled by that loss's 1/batch), e.g. the MoE's
   * supervised loss gated by gamma (models.py:74).  Added through gamma's softmax Jacobian to d_s, so it reaches Z,
   * mean / log_var and the prior tables together with the ELBO's own gradient. */
  const float* d_gate_extra; int64_t ld_dge;
  float x_scale;           /* uint8 targets: value of one unit (see dmvae_stage_input); 0 means 1 */
  /* Reconstruction term computed elsewhere (dmvae_recon_fuse in the output layer's GEMM epilogue): with r_parts > 0
   * dmvae_elbo_fwd_bwd does the LATENT part only (X / decoded / d_decoded are not touched and may be NULL; it can run
   * before the decoder), and dmvae_elbo_reduce completes per_sample: recon = sum of r_part[row, :] in order. */
  const float* r_part; int32_t r_parts;
} dmvae_elbo_args;
int dmvae_elbo_fwd_bwd(dmvae_ctx* ctx, const dmvae_elbo_args* a, void* stream);

/* cross-sample reductions of the same pass: prior-table gradients (added into d_prior_*), and the
 * three loss terms.  workspace: fp32, at least dmvae_elbo_reduce_workspace(...) floats. */
int64_t dmvae_elbo_reduce_workspace(int rows, int L, int K);
int dmvae_elbo_reduce(dmvae_ctx* ctx, const dmvae_elbo_args* a, float* d_prior_means, float* d_prior_log_vars,
                      int accumulate, float* loss_out /* [4]: recon, KL_c, KL_z, loss (x inv_global_batch) */,
                      float* workspace, void* stream);
/* The same reduction in two launches, for the fused reconstruction term (r_part): stage 1 = the prior-table partial sums
 * (needs the latent part's outputs only: it can run beside the decoder), stage 2 = complete per_sample from r_part, the
 * loss terms and the final table sums (after the output layer's GEMM).  stage 0 = dmvae_elbo_reduce. */
int dmvae_elbo_reduce_stage(dmvae_ctx* ctx, const dmvae_elbo_args* a, float* d_prior_means, float* d_prior_log_vars,
                            int accumulate, float* loss_out, float* workspace, int stage, void* stream);

/* ---- MoE expert head (models.py:76-111, :149-163) ------------------------------------------- */
typedef struct dmvae_moe_args {
  int32_t classification;  /* 1: softmax experts + NLL x1000, 0: regression */
  int32_t rows, E, O;
  const float* pred; int64_t ld_pred;       /* fp32 [rows, >= E*O] expert outputs, column e*O + o (bias already added) */
  const float* gate; int64_t ld_gate;       /* fp32 [rows,E] = vae.cluster_probs (models.py:74) */
  const float* Y; int64_t ldy;              /* fp32 [rows,O] */
  float inv_global_batch;
  float* per_sample;       /* [rows,2]: supervised loss summand, error summand */
  float* y_soft;           /* [rows,O] reconstructed_Y_soft / reconstructed_Y */
  int32_t* pred_class;     /* [rows] (classification) */
  void* d_pred; int32_t dpred_dtype; int64_t ld_dpred; int32_t dpred_cols;   /* [rows, ld] */
  float* d_gate; int64_t ld_dgate;          /* fp32 [rows,E] */
} dmvae_moe_args;
int dmvae_moe_fwd_bwd(dmvae_ctx* ctx, const dmvae_moe_args* a, void* stream);
/* d_logits[k] (+)= q_k (d_gate_k - sum_j q_j d_gate_j): the gate's softmax Jacobian, written in operand dtype
 * for the c-head backward.  accumulate=0 overwrites and zeroes columns [K, cols). */
int dmvae_softmax_bwd_add(dmvae_ctx* ctx, int rows, int K, const float* q, const float* d_gate, int64_t ld_dgate,
                          void* d_logits, int dtype, int64_t ld_dlogits, int accumulate, int cols, void* stream);

/* q = softmax(scores) row-wise (base_models.py:249), fp32 [rows,K] dense output */
int dmvae_softmax_rows(dmvae_ctx* ctx, const float* scores, int64_t ld, int rows, int K, float* q, void* stream);
/* out[c] = scale * sum_r src[r,c]  (deterministic batch reductions of per-sample terms) */
int dmvae_reduce_columns(dmvae_ctx* ctx, const float* src, int64_t ld, int rows, int cols, float scale, float* out, void* stream);
/* fp32 features [rows,n] -> operand matrix [rows, out_cols] with optional ReLU and the ones column at n
 * (models.py:58-61: inp2cls = relu(vae.mean) when featLearn) */
int dmvae_stage_features(dmvae_ctx* ctx, const float* src, int64_t ld, int rows, int n, int relu, void* out, int out_dtype,
                         int64_t ld_out, int out_cols, void* stream);

/* ---- TF-semantics Adam (tf.train.AdamOptimizer, base_models.py:102-110) ---------------------
 * theta -= lr_t m/(sqrt(v)+eps), lr_t = lr sqrt(1-b2^t)/(1-b1^t) computed by the caller in double.
 * Flat over n fp32 parameters.  Optionally writes the bf16 operand copy and clears the gradient. */
#define DMVAE_ADAM_ZERO_GRADS 1 /* clear the gradient after the update */
#define DMVAE_ADAM_BACKGROUND 2 /* small-block launch shape that co-resides with a running GEMM kernel (streamed update) */
int dmvae_adam(dmvae_ctx* ctx, float* params, float* grads, float* m, float* v, void* params_bf16 /* or NULL */,
               int64_t n, float lr_t, const float* lr_t_dev /* optional device scalar overriding lr_t */,
               float beta1, float beta2, float eps, float grad_scale, int flags /* DMVAE_ADAM_* */, void* stream);
/* Per-step device state for CUDA-graph replay: state = {uint64 step; uint32 t; float lr_t}.  One tiny kernel:
 * step += 1, t += 1, lr_t = lr sqrt(1-beta2^t)/(1-beta1^t) (double precision). */
int dmvae_step_tick(dmvae_ctx* ctx, void* state_dev, float lr, float beta1, float beta2, void* stream);
/* ring[(step % cap) * n + j] = src[j], j < n <= 32: step read from state_dev (the block dmvae_step_tick advances) when it is
 * not NULL, else the argument.  Lets a captured step log its loss terms without a copy between two graph launches. */
int dmvae_log_append(dmvae_ctx* ctx, const float* src, int n, float* ring, int cap, const void* state_dev, uint64_t step,
                     void* stream);

/* ---- evaluation helpers (get_accuracy, base_models.py:425-432; utils.py:22-34) -------------- */
/* argmax over K of fp32 [rows,K] and contingency counts d[cluster, class] += 1 (int32 [K, n_labels]) */
int dmvae_argmax_contingency(dmvae_ctx* ctx, const float* scores, int64_t ld, int rows, int K,
                             const int32_t* classes, int n_labels, int32_t* argmax_out, int32_t* counts, void* stream);

/* ---- data parallel (new; the reference is single-device) ------------------------------------
 * One kernel per rank: sum the N ranks' gradient shards through NVLink peer pointers, apply Adam to
 * the owned 1/N shard, store the updated fp32 + bf16 parameters into every rank's replica and clear the gradient
 * shard it consumed in every replica.  The caller brackets the launch with a cross-rank barrier on each side. */
int dmvae_dp_reduce_adam(dmvae_ctx* ctx, int rank, int world, float* const* grads_peers_host,
                         float* const* params_peers_host, void* const* params_bf16_peers_host,
                         float* const* params_rep_peers_host /* every replica's fp32 master, or NULL */,
                         const int64_t* rep_ranges /* [2 n_rep] float index ranges [lo, hi) whose fp32 master is written into
                                                      EVERY replica even where params_peers_host[r] is NULL (prior tables: read in
                                                      fp32 by the ELBO kernel; logits layer: dmvae_split3_bf16 needs the master) */,
                         int n_rep,
                         float* m, float* v, int64_t n, int64_t shard_begin, int64_t shard_end,
                         float lr_t, const float* lr_t_dev, float beta1, float beta2, float eps,
                         int flags /* DMVAE_ADAM_ZERO_GRADS: clear the consumed shard in every replica (0: every rank clears its
                                      own gradient buffer after the closing barrier); DMVAE_ADAM_BACKGROUND: small blocks */,
                         void* stream);
/* NVSwitch form of dmvae_dp_reduce_adam: mc_* are MULTICAST addresses of the ranks' symmetric buffers (gradients, fp32
 * parameters or NULL, bf16 operand copy or NULL).  The gradient shard is summed in the switch (multimem.ld_reduce), the
 * update is applied to the owned shard [shard_begin, shard_end) with the local Adam slots m, v (indexed from the shard's
 * start), and the new parameters are broadcast through the switch (multimem.st).  With mc_params == NULL the fp32 master
 * of the shard is written to params_local only.  Bracket with cross-rank barriers like dmvae_dp_reduce_adam. */
int dmvae_dp_reduce_adam_mc(dmvae_ctx* ctx, const float* mc_grads, float* mc_params, void* mc_params_bf16,
                            float* params_local, float* m, float* v, int64_t n, int64_t shard_begin, int64_t shard_end,
                            float lr_t, const float* lr_t_dev, float beta1, float beta2, float eps, void* stream);
/* Cross-GPU barrier over peer-mapped flag pads: pads_host[r] = rank r's pad, uint32 [DMVAE_DP_CHANNELS][8], zero-initialised,
 * in symmetric memory; epochs = uint32 [DMVAE_DP_CHANNELS] in ordinary device memory, zero-initialised.  Every rank must
 * issue the same sequence of (channel) barriers; barriers on different channels may be in flight on different streams. */
#define DMVAE_DP_CHANNELS 32
int dmvae_dp_barrier(dmvae_ctx* ctx, int rank, int world, uint32_t* const* pads_host, uint32_t* epochs, int channel,
                     void* stream);
/* Peer-mapped buffers WITHOUT torch (one process per GPU): allocate the flat parameter / gradient buffer with
 * dmvae_dp_alloc, send the 64-byte handle to the other ranks over whatever the host has (a file, a socket, MPI), map the
 * peers' buffers with dmvae_dp_open, and pass the pointers to dmvae_dp_reduce_adam / dmvae_dp_barrier.  (The Python host
 * of this repository uses torch.distributed's symmetric memory for the same purpose.)  CUDA IPC under the hood: the
 * buffers are NVLink peer-mapped on a single node. */
typedef struct dmvae_ipc_handle { unsigned char bytes[64]; } dmvae_ipc_handle;
int dmvae_dp_alloc(dmvae_ctx* ctx, int64_t bytes, void** local_ptr, dmvae_ipc_handle* handle_out);
int dmvae_dp_open(dmvae_ctx* ctx, const dmvae_ipc_handle* peer_handle, void** peer_ptr);
int dmvae_dp_close(dmvae_ctx* ctx, void* peer_ptr);
int dmvae_dp_free(dmvae_ctx* ctx, void* local_ptr);

/* params_peers_host[r] may be NULL for r != rank when params_bf16_peers_host[r] is given: the fp32 master copy of a
 * shard then lives on its owner only (the bf16 operand copy, which is all the GEMMs read, is still replicated). */
/* zero a fp32 buffer (gradient accumulators) */
int dmvae_zero_f32(dmvae_ctx* ctx, float* p, int64_t n, void* stream);
/* fp32 -> bf16 copy (operand copy of the parameters) */
int dmvae_cast_bf16(dmvae_ctx* ctx, const float* src, void* dst, int64_t n, void* stream);

/* ---- exact cluster assignments with bf16 tensor cores (base_models.py:241-249, :425-432) -------------------------
 * The logits layer has K <= stride valid output columns inside a zero-padded block of >= 3 stride columns.  Its bf16
 * OPERAND copy is written as three column groups  W = hi + lo + lo2  (8 + 8 + 8 mantissa bits: the fp32 weight exactly):
 *   op[r, k] = bf16(W[r,k]),  op[r, stride+k] = bf16(W - hi),  op[r, 2 stride+k] = bf16(W - hi - lo),   k < K,
 * so ONE tcgen05 GEMM of unchanged shape yields three partial logits whose sum (dmvae_fold3, or the fold of
 * dmvae_reparam_fwd) is the fp32-exact product of the bf16 activations with the fp32 weights: argmax q(c|x) no longer
 * depends on the rounding of the weights.  The padding columns carry zero gradients, so dgrad / wgrad are unaffected. */
int dmvae_split3_bf16(dmvae_ctx* ctx, const float* W, void* W_bf16, int rows, int64_t ld, int K, int stride, void* stream);
int dmvae_fold3(dmvae_ctx* ctx, float* Y, int64_t ld, int rows, int K, int stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMVAE_B200_H */
